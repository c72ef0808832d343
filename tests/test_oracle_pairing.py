"""The pairing oracle (oracle/pairing.py) and what it pins: bilinearity on all four curves, the EIP-4844 ceremony
relation between the reference's SRS file and the verification key embedded in crypto/blobs/kzg.go, the subgroup
membership of every constant of config/statetransition_vkey.sol, and pairing verification of the KZG opening / cell
proofs (golden vectors and fresh oracle outputs)."""
import json
import os

import pytest

from oracle import curve as OC
from oracle import kzg as OK
from oracle import pairing

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


@pytest.mark.parametrize("name", ["bn254", "bls12_381", "bls12_377", "bw6_761"])
def test_pairing_bilinear_nondegenerate(name):
    pr = pairing.get(name)
    cx = pr.cx
    P, Q = cx.g1, cx.g2
    e1 = pr.pair(P, Q)
    assert e1 != pr.F.one and pr.F.pow(e1, cx.r) == pr.F.one
    a, b = 0x1234567, 0xfedcba9
    assert pr.pair(cx.G1.mul(P, a), cx.G2.mul(Q, b)) == pr.F.pow(e1, a * b % cx.r)
    assert pr.product_is_one([(cx.G1.mul(P, a), Q), (cx.G1.neg(P), cx.G2.mul(Q, a))])
    assert not pr.product_is_one([(cx.G1.mul(P, a), Q), (cx.G1.neg(P), cx.G2.mul(Q, a + 1))])
    assert pr.pair(None, Q) == pr.F.one and pr.pair(P, None) == pr.F.one


@pytest.fixture(scope="module")
def srs():
    raw1 = open(os.path.join(GOLD, "kzg_g1_monomial.bin"), "rb").read()
    raw2 = open(os.path.join(GOLD, "kzg_g2_monomial.bin"), "rb").read()
    g1 = lambda j: OK.g1_decompress(raw1[48 * j:48 * (j + 1)])
    g2 = lambda j: OK.g2_decompress(raw2[96 * j:96 * (j + 1)])
    return g1, g2


def test_ceremony_relation_pins_the_bls12_381_pairing(srs):
    """e([tau^j]_1, G_2) == e(G_1, [tau^j]_2): G1 powers from the SRS file's monomial block, G2 powers from its G2
    block, whose first two entries are byte-identical to crypto/blobs/kzg.go:26-45 (tools/make_golden.py asserts it)."""
    g1, g2 = srs
    pr = pairing.get("bls12_381")
    cx = pr.cx
    assert g1(0) == cx.g1 and g2(0) == OK.cx_g2_generator()
    assert cx.G2.on_curve(g2(1)) and cx.G2.mul(g2(1), cx.r) is None
    for j in (1, 64):
        assert pr.product_is_one([(g1(j), g2(0)), (cx.G1.neg(g1(0)), g2(j))]), j
    assert not pr.product_is_one([(g1(2), g2(0)), (cx.G1.neg(g1(0)), g2(1))])


def test_statetransition_vk_constants_are_subgroup_points():
    """Every point of the verifying key the reference deploys on-chain (config/statetransition_vkey.sol:60-115) lies on
    the oracle's BN254 G1 / twist curve and in the order-r subgroup - pins the twist (3/(9+u)), the Fp2 element order of
    EIP-197 (X_1 = imaginary part) and the group order the Solidity-verifier port relies on."""
    c = {k: int(v, 16) for k, v in json.load(open(os.path.join(GOLD, "statetransition_vk.json")))["constants"].items()}
    cx = OC.ctx("bn254")
    g1_names = ["ALPHA", "CONSTANT"] + ["PUB_%d" % i for i in range(9)]
    for n in g1_names:
        pt = (c[n + "_X"], c[n + "_Y"])
        assert cx.G1.on_curve(pt), n
    for n in ["BETA_NEG", "GAMMA_NEG", "DELTA_NEG", "PEDERSEN_G", "PEDERSEN_GSIGMANEG"]:
        pt = ((c[n + "_X_0"], c[n + "_X_1"]), (c[n + "_Y_0"], c[n + "_Y_1"]))
        assert cx.G2.on_curve(pt), n
        assert cx.G2.mul(pt, cx.r) is None, n


@pytest.fixture(scope="module")
def lagrange():
    raw = open(os.path.join(GOLD, "kzg_g1_lagrange.bin"), "rb").read()
    return [OK.g1_decompress(raw[48 * i:48 * (i + 1)]) for i in range(4096)]


def test_opening_proofs_verify_against_tau_g2(srs, lagrange):
    """oracle ComputeProof outputs (out-of-domain and in-domain points) pass e(C - [y], G2) = e(pi, [tau]_2 - [z]_2)
    with the reference's [tau]_2; a wrong claim does not."""
    g1, g2 = srs
    blob = open(os.path.join(GOLD, "blobdata1.bin"), "rb").read()
    kat = {c["name"]: c for c in json.load(open(os.path.join(GOLD, "kzg_kat.json")))["cases"]}
    commitment = bytes.fromhex(kat["blobdata1"]["commitment"])
    roots = OK.roots_of_unity_brp(4096)
    for z in (0x1234567890abcdef1234567890abcdef, roots[5]):
        proof, y = OK.compute_proof(blob, z, lagrange)
        assert OK.verify_kzg_proof(commitment, z, y, proof, g2(1))
        assert not OK.verify_kzg_proof(commitment, z, (y + 1) % OK.P.BLS12_381.r, proof, g2(1))


@pytest.mark.parametrize("k", [0, 1, 77, 127])
def test_golden_cell_proofs_verify_against_tau64_g2(srs, k):
    g1, g2 = srs
    blob = open(os.path.join(GOLD, "blobdata1.bin"), "rb").read()
    kat = {c["name"]: c for c in json.load(open(os.path.join(GOLD, "kzg_kat.json")))["cases"]}
    commitment = bytes.fromhex(kat["blobdata1"]["commitment"])
    proofs = json.load(open(os.path.join(GOLD, "kzg_cell_kat.json")))["proofs"]
    mono64 = [g1(j) for j in range(64)]
    vals = OK.extended_cell_values(blob, k)
    assert OK.verify_cell_proof(commitment, k, vals, bytes.fromhex(proofs[str(k)]), mono64, g2(64))
    bad = list(vals)
    bad[3] = (bad[3] + 1) % OK.P.BLS12_381.r
    assert not OK.verify_cell_proof(commitment, k, bad, bytes.fromhex(proofs[str(k)]), mono64, g2(64))
