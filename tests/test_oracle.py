"""CPU tests of the oracle itself: curve parameters, NTT semantics, quotient identity, and the
Groth16 restatement (proof points == closed-form exponents, verifier equation holds)."""
import random

import pytest

from oracle import curve as C
from oracle import groth16 as G
from oracle import ntt as N
from oracle import params as P

CURVES = ["bn254", "bls12_377", "bls12_381", "bw6_761"]


@pytest.mark.parametrize("name", CURVES)
def test_curve_parameters(name):
    c = P.CURVES[name]
    cx = C.ctx(name)
    assert cx.G1.on_curve(cx.g1) and cx.G2.on_curve(cx.g2)
    assert cx.G1.mul(cx.g1, c.r) is None and cx.G2.mul(cx.g2, c.r) is None
    # 2-adic root of unity and non-residue coset generator
    assert pow(c.root_of_unity, 1 << c.two_adicity, c.r) == 1
    assert pow(c.root_of_unity, 1 << (c.two_adicity - 1), c.r) == c.r - 1
    assert pow(c.mult_gen, (c.r - 1) // 2, c.r) == c.r - 1
    assert c.p.bit_length() <= 64 * c.fp_limbs64 and c.r.bit_length() <= 64 * c.fr_limbs64


def test_bw6_is_two_chain_over_bls12_377():
    assert P.BW6_761.r == P.BLS12_377.p


@pytest.mark.parametrize("name", CURVES)
def test_group_law_consistency(name):
    cx = C.ctx(name)
    rnd = random.Random(3)
    for G_, g in ((cx.G1, cx.g1), (cx.G2, cx.g2)):
        a, b = rnd.randrange(cx.r), rnd.randrange(cx.r)
        pa, pb = G_.mul(g, a), G_.mul(g, b)
        assert G_.add(pa, pb) == G_.mul(g, (a + b) % cx.r)
        assert G_.add(pa, G_.neg(pa)) is None
        assert G_.add(pa, pa) == G_.mul(g, 2 * a % cx.r)
        pts = [G_.mul(g, rnd.randrange(cx.r)) for _ in range(40)]
        sc = [rnd.randrange(cx.r) for _ in range(40)]
        assert G_.msm(pts, sc) == G_.msm_naive(pts, sc)


@pytest.mark.parametrize("name", CURVES)
def test_fft_matches_definition(name):
    c = P.CURVES[name]
    dom = N.Domain(c, 16)
    rnd = random.Random(4)
    a = [rnd.randrange(c.r) for _ in range(16)]
    want = N.dft_definition(a, dom.omega, c.r)
    assert N.bit_reverse_list(N.fft(a, dom)) == want                       # DIF: natural -> bit-reversed
    assert N.fft(N.bit_reverse_list(a), dom, dit=True) == want             # DIT: bit-reversed -> natural
    assert N.fft(N.fft(a, dom), dom, inverse=True, dit=True) == a          # round trip
    co = N.fft(a, dom, coset=True)
    shifted = [x * pow(dom.g, j, c.r) % c.r for j, x in enumerate(a)]
    assert N.bit_reverse_list(co) == N.dft_definition(shifted, dom.omega, c.r)
    assert N.fft(co, dom, inverse=True, dit=True, coset=True) == a


@pytest.mark.parametrize("name", ["bn254", "bls12_377", "bw6_761"])
def test_compute_h_polynomial_identity(name):
    c = P.CURVES[name]
    q = c.r
    n = 32
    dom = N.Domain(c, n)
    rnd = random.Random(5)
    a = [rnd.randrange(q) for _ in range(n - 3)]
    b = [rnd.randrange(q) for _ in range(n - 3)]
    cc = [x * y % q for x, y in zip(a, b)]
    h = N.bit_reverse_list(N.compute_h(a, b, cc, dom))          # natural-order coefficients
    assert h[n - 1] == 0                                         # deg h <= n - 2
    pad = lambda v: v + [0] * (n - len(v))
    ca = N.fft(N.fft(pad(a), dom, inverse=True), dom, dit=True)  # sanity: interpolation round trip
    assert ca == pad(a)
    coef = lambda v: N.bit_reverse_list(N.fft(pad(v), dom, inverse=True))
    pa, pb, pc = coef(a), coef(b), coef(cc)
    ev = lambda poly, x: sum(cf * pow(x, i, q) for i, cf in enumerate(poly)) % q
    for _ in range(3):
        x = rnd.randrange(q)
        assert (ev(pa, x) * ev(pb, x) - ev(pc, x)) % q == ev(h, x) * (pow(x, n, q) - 1) % q


@pytest.mark.parametrize("name", ["bn254", "bls12_377"])
def test_groth16_prove_matches_closed_form_and_verifies(name):
    cx = C.ctx(name)
    q = cx.r
    rnd = random.Random(6)
    cs, W = G.synthetic_circuit(27, 3, q, seed=11, n_commit=1, n_private_committed=3)
    tox = G.Toxic(*(rnd.randrange(1, q) for _ in range(5)), sigmas=[rnd.randrange(1, q)])
    pk, ex = G.setup(cs, cx, tox)
    r, s = rnd.randrange(q), rnd.randrange(q)
    proof = G.prove(cs, pk, W, r, s, cx)
    A, B, Cx = G.proof_exponents(cs, ex, tox, W, r, s, q)
    assert proof["Ar"] == cx.G1.mul(cx.g1, A)
    assert proof["Bs"] == cx.G2.mul(cx.g2, B)
    assert proof["Krs"] == cx.G1.mul(cx.g1, Cx)
    assert G.verify_exponent(cs, ex, tox, W, A, B, Cx, q)
    # Pedersen proof of knowledge: Pok = sigma * Commitment
    assert proof["CommitmentPok"] == cx.G1.mul(proof["Commitments"][0], tox.sigmas[0])
    # a wrong witness must not verify
    W2 = list(W)
    W2[1] = (W2[1] + 1) % q
    assert not G.verify_exponent(cs, ex, tox, W2, A, B, Cx, q)


@pytest.mark.parametrize("cname,kind,ncommit,npubc", [("bn254", "solidity", 1, 1), ("bls12_377", "default", 2, 0),
                                                       ("bw6_761", "default", 1, 0)])
def test_pairing_verifier_accepts_oracle_proofs(cname, kind, ncommit, npubc):
    """gnark's verifier (pairings + real commitment hashes) and the Solidity-verifier port accept the oracle's proofs and
    reject a tampered proof / public input."""
    import random
    from oracle import hashes as H
    OC, OG = C, G
    cx = OC.ctx(cname)
    q = cx.r
    rnd = random.Random(5)
    cs, W = OG.synthetic_circuit(20, 4, q, seed=3, n_commit=ncommit, n_private_committed=3, n_public_committed=npubc)
    tox = OG.Toxic(*(rnd.randrange(1, q) for _ in range(5)), sigmas=[rnd.randrange(1, q) for _ in range(ncommit)])
    pk, ex = OG.setup(cs, cx, tox)
    vk = OG.verifying_key(cs, cx, tox, ex)
    W = list(W)
    # what the solver does: commitment -> challenge -> commitment wire, then the dependent wires
    for cm, key in zip(cs.commitments, pk["CommitmentKeys"]):
        cpt = cx.G1.msm(key["Basis"], [W[w] for w in cm["private_committed"]])
        W[cm["commitment_index"]] = H.commitment_challenge(kind, cpt, [W[w] for w in cm["public_committed"]], q, cx.p)
    ev = lambda t: sum(W[w] * cf for w, cf in t) % q
    for k in range(cs.nb_constraints):
        W[cs.O[k][0][0]] = ev(cs.L[k]) * ev(cs.R[k]) % q
    r, s = rnd.randrange(q), rnd.randrange(q)
    fold = H.fold_challenge([W[cm["commitment_index"]] for cm in cs.commitments], q) if ncommit > 1 else None
    proof = OG.prove(cs, pk, W, r, s, cx, fold_challenge=fold)
    public = W[1:cs.nb_public]
    assert OG.verify(vk, proof, public, cx, kind)
    bad = dict(proof)
    bad["Krs"] = cx.G1.add(proof["Krs"], cx.g1)
    assert not OG.verify(vk, bad, public, cx, kind)
    wrong = list(public)
    wrong[0] = (wrong[0] + 1) % q
    assert not OG.verify(vk, proof, wrong, cx, kind)
    assert not OG.verify(vk, proof, public, cx, "default" if kind == "solidity" else "solidity") or not cs.commitments
    if cname == "bn254":
        c = OG.solidity_constants(vk, cx)
        A, B, Cc = proof["Ar"], proof["Bs"], proof["Krs"]
        p8 = [A[0], A[1], B[0][1], B[0][0], B[1][1], B[1][0], Cc[0], Cc[1]]
        idx = [wire - 1 for wire in cs.commitments[0]["public_committed"]]
        assert OG.solidity_verify_proof(c, p8, proof["Commitments"][0], proof["CommitmentPok"], public, idx, cx)
        assert not OG.solidity_verify_proof(c, p8, proof["Commitments"][0], proof["CommitmentPok"], wrong, idx, cx)
        assert not OG.solidity_verify_proof(c, p8, proof["Commitments"][0], proof["CommitmentPok"], public + [q], idx, cx)


def test_hash_known_answers():
    from oracle import hashes as H
    OC = C
    assert H.keccak256(b"").hex() == "c5d2460186f7233c927e7db2dcc703c0e500b653ca82273b7bfad8045d85a470"
    assert H.keccak256(b"abc").hex() == "4e03657aea45a94fc7d47ba826c8d667c0d1e6e33a64a036ec44f58fa12d6c45"
    dst = b"QUUX-V01-CS02-with-expander-SHA256-128"          # RFC 9380 appendix K.1
    assert H.expand_message_xmd(b"", dst, 0x20).hex() == "68a985b87eb6b46952128911f2a4412bbc302a9d759667f87f7a21d803f07235"
    assert H.expand_message_xmd(b"abc", dst, 0x20).hex() == "d8ccab23b5985ccea865c6c97b6e5b8350e794e603b4b97902f53a8a0d605615"
    # the product's host-side implementation (derived tables) agrees with the oracle's (tabulated constants)
    import random
    from davinci_node_b200 import hash_to_field as P
    rnd = random.Random(1)
    for n in (0, 1, 135, 136, 137, 300, 1000):
        d = bytes(rnd.randrange(256) for _ in range(n))
        assert P.keccak256(d) == H.keccak256(d)
    r = OC.ctx("bw6_761").r
    assert P.hash_to_fr(b"hello", b"G16-BSB22", 2, r) == H.fr_hash(b"hello", b"G16-BSB22", 2, r)
    assert P.commitment_challenge("solidity", (5, 7), [1, 2], r, 96) == H.commitment_challenge("solidity", (5, 7), [1, 2], r, OC.ctx("bw6_761").p)
