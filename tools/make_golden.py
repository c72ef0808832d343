#!/usr/bin/env python3
"""Builds the committed fixtures under tests/golden/ from the reference's own data files.
Run in the build container only (reads /root/reference); the GPU box uses the committed outputs.

  kzg_g1_lagrange.bin   4096 x 48-byte compressed G1 Lagrange points, file order
                        (/root/reference/config/kzg_trusted_setup.txt lines 3..4098)
  kzg_g1_monomial_64.bin first 64 monomial-basis G1 points (lines 4164..) for the cross-check
  kzg_g1_monomial.bin   all 4096 monomial-basis G1 points (EIP-7594 cell proofs)
  kzg_g2_monomial.bin   the 65 monomial-basis G2 points [tau^j]_2 (lines 4099..4163); [tau^0]_2, [tau^1]_2 are checked here
                        against the bytes embedded in /root/reference/crypto/blobs/kzg.go:26-45
  statetransition_vk.json the BN254 Groth16 + Pedersen verifying key hard-coded in
                        /root/reference/config/statetransition_vkey.sol:60-115 (decimal constants, as ints in hex)
  kzg_cell_kat.json     oracle cell proofs (cells 0, 1, 77, 127) of the first sample blob
  kzg_kat.json          known-answer commitments computed by the oracle (oracle/kzg.py) for the
                        reference's deterministic test blobs (crypto/blobs/testdata.go:101-132) and
                        the first 128 KiB sample blob (crypto/blobs/testdata/blobdata1.txt)
"""
import hashlib
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import kzg  # noqa: E402

REF = "/root/reference"
OUT = os.path.join(ROOT, "tests", "golden")


def mono_hex_first(lines, n1, n2):
    return lines[2 + n1 + n2]


def main():
    os.makedirs(OUT, exist_ok=True)
    lines = open(os.path.join(REF, "config", "kzg_trusted_setup.txt")).read().split()
    n1, n2 = int(lines[0]), int(lines[1])
    assert (n1, n2) == (4096, 65)
    lag_hex = lines[2:2 + n1]
    mono_hex = lines[2 + n1 + n2:2 + n1 + n2 + n1]
    lag_bytes = b"".join(bytes.fromhex(h) for h in lag_hex)
    open(os.path.join(OUT, "kzg_g1_lagrange.bin"), "wb").write(lag_bytes)
    open(os.path.join(OUT, "kzg_g1_monomial_64.bin"), "wb").write(b"".join(bytes.fromhex(h) for h in mono_hex[:64]))
    lag = [kzg.g1_decompress(bytes.fromhex(h)) for h in lag_hex]
    g2_hex = lines[2 + n1:2 + n1 + n2]
    g2_bytes = b"".join(bytes.fromhex(h) for h in g2_hex)
    assert len(g2_bytes) == 65 * 96
    open(os.path.join(OUT, "kzg_g2_monomial.bin"), "wb").write(g2_bytes)
    # the verification key the reference embeds for its in-circuit KZG check: G1 || G2[0] || G2[1]
    import re
    src = open(os.path.join(REF, "crypto", "blobs", "kzg.go")).read()
    body = src[src.index("var srsData = []byte{"):src.index("// initVerificationKey")]
    vk = bytes(int(x, 16) for x in re.findall(r"0x([0-9a-fA-F]{2})", body))
    assert len(vk) == 48 + 96 + 96
    assert vk[48:144] == g2_bytes[:96] and vk[144:] == g2_bytes[96:192], "kzg.go's [tau]_2 differs from the SRS file"
    assert vk[:48] == bytes.fromhex(mono_hex_first(lines, n1, n2)), "kzg.go's G1 differs from [tau^0]_1"

    kat = {"srs_sha256": hashlib.sha256(lag_bytes).hexdigest(), "cases": []}

    def seed_blob(seed):
        b = bytearray(4096 * 32)
        for i in range(50):
            b[i * 32:(i + 1) * 32] = (seed + i).to_bytes(32, "big")
        return bytes(b)

    ones = b"".join((1).to_bytes(32, "big") for _ in range(4096))
    sample = bytes.fromhex(open(os.path.join(REF, "crypto", "blobs", "testdata", "blobdata1.txt")).read().strip())
    open(os.path.join(OUT, "blobdata1.bin"), "wb").write(sample)
    for name, blob in [("seed1", seed_blob(1)), ("seed2", seed_blob(2)), ("all_ones", ones), ("zero", bytes(4096 * 32)),
                       ("blobdata1", sample)]:
        c = kzg.blob_to_commitment(blob, lag)
        kat["cases"].append({"name": name, "blob_sha256": hashlib.sha256(blob).hexdigest(), "commitment": c.hex()})
        print(name, c.hex())
    json.dump(kat, open(os.path.join(OUT, "kzg_kat.json"), "w"), indent=1)
    import re as _re
    sol = open(os.path.join(REF, "config", "statetransition_vkey.sol")).read()
    consts = {m.group(1): hex(int(m.group(2), 0)) for m in
              _re.finditer(r"uint256 constant ((?:ALPHA|BETA_NEG|GAMMA_NEG|DELTA_NEG|PEDERSEN_G|PEDERSEN_GSIGMANEG|CONSTANT|PUB_\d+)_[XY](?:_[01])?) = (\d+);", sol)}
    assert len(consts) == 2 + 4 * 5 + 2 + 2 * 9, len(consts)
    json.dump({"source": "config/statetransition_vkey.sol:60-115", "constants": consts},
              open(os.path.join(OUT, "statetransition_vk.json"), "w"), indent=1, sort_keys=True)
    mono_bytes = b"".join(bytes.fromhex(h) for h in mono_hex)
    open(os.path.join(OUT, "kzg_g1_monomial.bin"), "wb").write(mono_bytes)
    mono = [kzg.g1_decompress(bytes.fromhex(h)) for h in mono_hex]
    cells = kzg.compute_cell_proofs(sample, mono, cells=[0, 1, 77, 127])
    json.dump({"blob": "blobdata1", "mono_sha256": hashlib.sha256(mono_bytes).hexdigest(),
               "proofs": {str(k): v.hex() for k, v in cells.items()}},
              open(os.path.join(OUT, "kzg_cell_kat.json"), "w"), indent=1)


if __name__ == "__main__":
    main()
