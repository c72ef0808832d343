"""GPU parity of the EIP-4844 blob commitment (types.Blob.ComputeCommitment mirror) against the
committed known-answer vectors (tests/golden/kzg_kat.json, made from the reference's SRS and test
blobs by tools/make_golden.py) and against the oracle on random blobs."""
import json
import os
import random

import pytest

from oracle import kzg as OK
from oracle import params as OP

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


@pytest.fixture(scope="module")
def kz():
    from davinci_node_b200 import kzg
    kzg.load_trusted_setup(open(os.path.join(GOLD, "kzg_g1_lagrange.bin"), "rb").read())
    return kzg


def _blob(name):
    if name.startswith("seed"):
        seed = int(name[4:])
        b = bytearray(4096 * 32)
        for i in range(50):
            b[i * 32:(i + 1) * 32] = (seed + i).to_bytes(32, "big")
        return bytes(b)
    if name == "all_ones":
        return b"".join((1).to_bytes(32, "big") for _ in range(4096))
    if name == "zero":
        return bytes(4096 * 32)
    return open(os.path.join(GOLD, name + ".bin"), "rb").read()


def test_known_answers(kz):
    kat = json.load(open(os.path.join(GOLD, "kzg_kat.json")))
    for case in kat["cases"]:
        got = kz.Blob(_blob(case["name"])).ComputeCommitment()
        assert got.hex() == case["commitment"], case["name"]


def test_random_blob_vs_oracle(kz):
    raw = open(os.path.join(GOLD, "kzg_g1_lagrange.bin"), "rb").read()
    lag = [OK.g1_decompress(raw[i:i + 48]) for i in range(0, len(raw), 48)]
    rnd = random.Random(8)
    r = OP.BLS12_381.r
    # statetransition-shaped blob: 2193 populated cells then zeros (state/blobs.go:58-96)
    cells = [rnd.randrange(OP.BN254.r) for _ in range(2193)] + [0] * (4096 - 2193)
    blob = b"".join(v.to_bytes(32, "big") for v in cells)
    assert kz.Blob(blob).ComputeCommitment() == OK.blob_to_commitment(blob, lag)
    cells = [rnd.randrange(r) for _ in range(4096)]
    cells[5], cells[6] = r - 1, 1
    blob = b"".join(v.to_bytes(32, "big") for v in cells)
    assert kz.Blob(blob).ComputeCommitment() == OK.blob_to_commitment(blob, lag)


def test_non_canonical_blob_is_rejected(kz):
    blob = bytearray(4096 * 32)
    blob[0:32] = OP.BLS12_381.r.to_bytes(32, "big")
    with pytest.raises(kz.BlobError):
        kz.Blob(bytes(blob)).ComputeCommitment()
    with pytest.raises(kz.BlobError):
        kz.Blob(b"\x00" * 100)
