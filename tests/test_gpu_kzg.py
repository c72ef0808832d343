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


def _lag():
    raw = open(os.path.join(GOLD, "kzg_g1_lagrange.bin"), "rb").read()
    return [OK.g1_decompress(raw[i:i + 48]) for i in range(0, len(raw), 48)]


def test_opening_proof_vs_oracle(kz):
    """Blob.ComputeProof (types/blobs.go:123): proof and claim equal the oracle's for a point outside the
    domain, a point inside it (EIP-4844 compute_quotient_eval_within_domain) and z = 0."""
    lag = _lag()
    rnd = random.Random(21)
    r = OP.BLS12_381.r
    cells = [rnd.randrange(OP.BN254.r) for _ in range(2193)] + [0] * (4096 - 2193)
    blob = b"".join(v.to_bytes(32, "big") for v in cells)
    roots = OK.roots_of_unity_brp(4096)
    for z in (rnd.randrange(r), roots[1234], 0):
        proof, y = kz.Blob(blob).ComputeProof(z)
        want_proof, want_y = OK.compute_proof(blob, z, lag)
        assert y == want_y, hex(z)
        assert proof == want_proof, hex(z)


def test_blob_proof_vs_oracle_and_linearity(kz):
    """ComputeBlobProof (types/blobs.go:111) against the oracle; and, as a size-independent property, the opening
    proof is linear in the blob: proof(p1 + p2, z) = proof(p1, z) + proof(p2, z), y likewise."""
    from oracle import curve as OC
    lag = _lag()
    cx = OC.ctx("bls12_381")
    rnd = random.Random(22)
    r = OP.BLS12_381.r
    c1 = [rnd.randrange(r) for _ in range(4096)]
    c2 = [rnd.randrange(r) for _ in range(4096)]
    enc = lambda cs: b"".join(v.to_bytes(32, "big") for v in cs)
    b1, b2, b12 = enc(c1), enc(c2), enc([(a + b) % r for a, b in zip(c1, c2)])
    com = kz.Blob(b1).ComputeCommitment()
    assert kz.Blob(b1).ComputeBlobProof(com) == OK.compute_blob_proof(b1, com, lag)
    z = rnd.randrange(r)
    (p1, y1), (p2, y2), (p12, y12) = (kz.Blob(b).ComputeProof(z) for b in (b1, b2, b12))
    assert y12 == (y1 + y2) % r
    assert OK.g1_decompress(p12) == cx.G1.add(OK.g1_decompress(p1), OK.g1_decompress(p2))


def test_opening_point_must_be_canonical(kz):
    blob = bytes(4096 * 32)
    with pytest.raises(kz.BlobError):
        kz.Blob(blob).ComputeProof(OP.BLS12_381.r)
    with pytest.raises(kz.BlobError):
        kz.Blob(blob).ComputeProof(1 << 256)
    proof, y = kz.Blob(blob).ComputeProof(5)           # zero polynomial: quotient 0 -> point at infinity, y = 0
    assert y == 0 and proof == bytes([0xC0]) + bytes(47)


def test_cell_proofs_vs_oracle(kz):
    """Blob.ComputeCellProofs (types/blobs.go:99): 128 proofs; the committed vectors (cells 0, 1, 77, 127 of the
    reference's sample blob) and the oracle on a statetransition-shaped blob (cells 5 and 126) must match, and the
    zero blob gives 128 points at infinity."""
    mono_raw = open(os.path.join(GOLD, "kzg_g1_monomial.bin"), "rb").read()
    kz.load_trusted_setup(open(os.path.join(GOLD, "kzg_g1_lagrange.bin"), "rb").read(), mono_raw)
    kat = json.load(open(os.path.join(GOLD, "kzg_cell_kat.json")))
    proofs = kz.Blob(_blob("blobdata1")).ComputeCellProofs()
    assert len(proofs) == 128
    for k, want in kat["proofs"].items():
        assert proofs[int(k)].hex() == want, k
    mono = [OK.g1_decompress(mono_raw[i:i + 48]) for i in range(0, len(mono_raw), 48)]
    rnd = random.Random(23)
    cells = [rnd.randrange(OP.BN254.r) for _ in range(2193)] + [0] * (4096 - 2193)
    blob = b"".join(v.to_bytes(32, "big") for v in cells)
    got = kz.Blob(blob).ComputeCellProofs()
    want = OK.compute_cell_proofs(blob, mono, cells=[5, 126])
    assert got[5] == want[5] and got[126] == want[126]
    assert kz.Blob(bytes(4096 * 32)).ComputeCellProofs() == [bytes([0xC0]) + bytes(47)] * 128


def test_gpu_openings_and_cell_proofs_verify_against_reference_tau_g2(kz):
    """Reference-held check no oracle output enters: the GPU's opening proofs satisfy
    e(C - [y]_1, G_2) = e(pi, [tau]_2 - [z]_2) with the [tau]_2 of /root/reference/crypto/blobs/kzg.go:26-45 and its cell
    proofs e(C - [I_k(tau)]_1, G_2) = e(pi_k, [tau^64]_2 - [h_k^64]_2) with the SRS file's [tau^64]_2, for a random
    statetransition-shaped blob (points outside and inside the evaluation domain, cells on both halves of the extension)."""
    g2raw = open(os.path.join(GOLD, "kzg_g2_monomial.bin"), "rb").read()
    g2 = lambda j: OK.g2_decompress(g2raw[96 * j:96 * (j + 1)])
    mono_raw = open(os.path.join(GOLD, "kzg_g1_monomial.bin"), "rb").read()
    kz.load_trusted_setup(open(os.path.join(GOLD, "kzg_g1_lagrange.bin"), "rb").read(), mono_raw)
    mono64 = [OK.g1_decompress(mono_raw[48 * j:48 * (j + 1)]) for j in range(64)]
    rnd = random.Random(31)
    r = OP.BLS12_381.r
    cells = [rnd.randrange(OP.BN254.r) for _ in range(2193)] + [0] * (4096 - 2193)
    blob = b"".join(v.to_bytes(32, "big") for v in cells)
    B = kz.Blob(blob)
    commitment, cell_proofs = B.ComputeCommitmentAndCellProofs()
    roots = OK.roots_of_unity_brp(4096)
    for z in (rnd.randrange(r), roots[77]):
        proof, y = B.ComputeProof(z)
        assert OK.verify_kzg_proof(commitment, z, y, proof, g2(1)), hex(z)
        assert not OK.verify_kzg_proof(commitment, z, (y + 1) % r, proof, g2(1))
    # ComputeBlobProof = the opening at the Fiat-Shamir challenge
    zc = OK.compute_challenge(blob, commitment)
    bp = B.ComputeBlobProof(commitment)
    assert OK.verify_kzg_proof(commitment, zc, OK.evaluate_in_evaluation_form(cells, zc, roots), bp, g2(1))
    for k in (3, 100):
        vals = OK.extended_cell_values(blob, k)
        assert OK.verify_cell_proof(commitment, k, vals, cell_proofs[k], mono64, g2(64)), k
    assert not OK.verify_cell_proof(commitment, 3, OK.extended_cell_values(blob, 4), cell_proofs[3], mono64, g2(64))
