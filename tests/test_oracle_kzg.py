"""CPU: the KZG oracle against the fixtures taken from the reference."""
import json
import os

from oracle import curve as C
from oracle import kzg as K
from oracle import ntt as N
from oracle import params as P

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def _lagrange():
    raw = open(os.path.join(GOLD, "kzg_g1_lagrange.bin"), "rb").read()
    return [K.g1_decompress(raw[i:i + 48]) for i in range(0, len(raw), 48)]


def test_srs_sum_is_generator_and_kats():
    cx = C.ctx("bls12_381")
    lag = _lagrange()
    assert len(lag) == 4096 and all(cx.G1.on_curve(p) for p in lag[:64])
    assert cx.G1.sum(lag) == cx.g1                     # sum_j L_j(tau) = 1
    kat = {c["name"]: c["commitment"] for c in json.load(open(os.path.join(GOLD, "kzg_kat.json")))["cases"]}
    # value derived independently during the survey (seed 1 test blob)
    assert kat["seed1"] == "b9ae6e8a29d0159952e7682f798cc9560976d20d7ca9e032d138736758e2a86bfc38a11f1c0660b159f332098441764c"
    assert kat["all_ones"] == K.g1_compress(cx.g1).hex()
    assert K.g1_decompress(K.g1_compress(lag[3])) == lag[3]


def test_lagrange_route_equals_monomial_route():
    """commitment(blob) == sum_j coeff_j [tau^j]_1 for a blob whose polynomial has degree < 64."""
    cx = C.ctx("bls12_381")
    q = P.BLS12_381.r
    lag = _lagrange()
    raw = open(os.path.join(GOLD, "kzg_g1_monomial_64.bin"), "rb").read()
    mono = [K.g1_decompress(raw[i:i + 48]) for i in range(0, len(raw), 48)]
    assert mono[0] == cx.g1
    coeff = [(7 * j + 3) % q for j in range(64)]
    dom = N.Domain(P.BLS12_381, 4096)
    evals_nat = N.dft_natural(coeff + [0] * (4096 - 64), dom.omega, q)      # p(omega^j)
    cells = N.bit_reverse_list(evals_nat)                                   # blob cell i = p(omega^brp(i))
    blob = b"".join(v.to_bytes(32, "big") for v in cells)
    assert K.blob_to_commitment(blob, lag) == K.g1_compress(cx.G1.msm(mono, coeff))


def test_opening_proof_lagrange_route_equals_monomial_route():
    """KZG opening (types/blobs.go:123 ComputeProof): for a polynomial of degree < 64 the proof computed in the
    evaluation basis (EIP-4844 compute_kzg_proof_impl, oracle restatement) must equal the commitment to the
    synthetic-division quotient in the monomial basis of the same ceremony, and y must equal p(z) - for a point
    outside and a point inside the evaluation domain."""
    cx = C.ctx("bls12_381")
    q = P.BLS12_381.r
    lag = _lagrange()
    raw = open(os.path.join(GOLD, "kzg_g1_monomial_64.bin"), "rb").read()
    mono = [K.g1_decompress(raw[i:i + 48]) for i in range(0, len(raw), 48)]
    coeff = [(11 * j * j + 5 * j + 1) % q for j in range(64)]
    dom = N.Domain(P.BLS12_381, 4096)
    cells = N.bit_reverse_list(N.dft_natural(coeff + [0] * (4096 - 64), dom.omega, q))
    blob = b"".join(v.to_bytes(32, "big") for v in cells)
    roots = K.roots_of_unity_brp(4096)
    assert roots[1] == q - 1 and pow(roots[2], 4, q) == 1            # brp order: 1, -1, i, -i, ...
    for z in (0x1234567890ABCDEF << 100 | 77, roots[9]):
        proof, y = K.compute_proof(blob, z, lag)
        assert y == sum(c * pow(z, k, q) for k, c in enumerate(coeff)) % q
        quot = [0] * 63
        rem = 0
        for k in range(63, -1, -1):
            rem = (rem * z + coeff[k]) % q
            if k > 0:
                quot[k - 1] = rem
        assert proof == K.g1_compress(cx.G1.msm(mono[:63], quot))


def test_fiat_shamir_challenge_layout():
    """compute_challenge: sha256(domain || u128_be(4096) || blob || commitment) mod r."""
    import hashlib
    blob = bytes(4096 * 32)
    com = bytes([0xC0]) + bytes(47)
    want = int.from_bytes(hashlib.sha256(b"FSBLOBVERIFY_V1_" + bytes(14) + b"\x10\x00" + blob + com).digest(), "big")
    assert K.compute_challenge(blob, com) == want % P.BLS12_381.r
