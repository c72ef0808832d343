"""Host logic of davinci_node_b200.verifier (the mirror of gnark's groth16.Verify) without a GPU: the two C-ABI calls it
makes - the G1 MSM and the product-of-pairings check - are replaced by oracle-backed stand-ins, so what is tested is the
part that runs on the host in the product as well: witness-size checks, commitment challenges with the real hash
functions, the fold of several commitments, which points get paired with which, the negation of Ar.  The kernels behind
the two calls are covered by tests/test_host_pairing.py (same template on the CPU) and tests/test_gpu_zz_verify.py."""
import random

import numpy as np
import pytest

from oracle import curve as OC
from oracle import groth16 as OG
from oracle import hashes as H
from oracle import pairing


def _twist_point_outside_subgroup(cx):
    G2 = cx.G2
    rnd = random.Random(17)
    while True:
        x = (rnd.randrange(cx.p), rnd.randrange(cx.p)) if isinstance(cx.g2[0], tuple) else rnd.randrange(cx.p)
        F = G2.F
        y = F.sqrt(F.add(F.mul(F.sqr(x), x), G2.b))
        if y is not None and G2.on_curve((x, y)) and G2.mul((x, y), cx.r) is not None:
            return (x, y)


@pytest.fixture
def fake_gpu(monkeypatch):
    from davinci_node_b200 import capi, verifier
    from davinci_node_b200.layout import Layout
    calls = {"msm": 0, "pairing": 0}

    def msm(L, group, point_bufs, scalars, device):
        cx = OC.ctx(L.name)
        G = cx.G1 if group == 1 else cx.G2
        pts = [L.dec_affine(b, group)[0] for b in point_bufs]
        calls["msm"] += 1
        return L.enc_affine([G.msm_naive(pts, [int(s) % L.r for s in scalars])], group)

    def check(curve_id, g1_bufs, g2_bufs, device=-1, want_gt=False):
        L = Layout(curve_id)
        pr = pairing.get(L.name)
        calls["pairing"] += 1
        assert len(g1_bufs) == len(g2_bufs)
        return pr.product_is_one([(L.dec_affine(a, 1)[0], L.dec_affine(b, 2)[0]) for a, b in zip(g1_bufs, g2_bufs)])

    monkeypatch.setattr(verifier, "_msm", msm)
    monkeypatch.setattr(verifier, "pairing_check", check)
    monkeypatch.setattr(capi, "init_once", lambda: None)
    return calls


@pytest.mark.parametrize("cname,kind,ncommit,npubc", [("bn254", "solidity", 1, 2), ("bn254", "default", 0, 0),
                                                      ("bls12_377", "default", 2, 1), ("bw6_761", "default", 1, 0)])
def test_verify_mirror_agrees_with_gnark_verifier_restatement(fake_gpu, cname, kind, ncommit, npubc):
    from oracle_bridge import proof_from_oracle, vk_from_oracle
    from davinci_node_b200 import prover, verifier
    from davinci_node_b200.layout import Layout
    cx = OC.ctx(cname)
    q = cx.r
    L = Layout(cname)
    rnd = random.Random(11)
    cs, W = OG.synthetic_circuit(20, 4, q, seed=3, n_commit=ncommit, n_private_committed=3, n_public_committed=npubc)
    tox = OG.Toxic(*(rnd.randrange(1, q) for _ in range(5)), sigmas=[rnd.randrange(1, q) for _ in range(ncommit)])
    pk, ex = OG.setup(cs, cx, tox)
    ovk = OG.verifying_key(cs, cx, tox, ex)
    W = list(W)
    for cm, key in zip(cs.commitments, pk["CommitmentKeys"]):
        cpt = cx.G1.msm(key["Basis"], [W[w] for w in cm["private_committed"]])
        W[cm["commitment_index"]] = H.commitment_challenge(kind, cpt, [W[w] for w in cm["public_committed"]], q, cx.p)
    ev = lambda t: sum(W[w] * cf for w, cf in t) % q
    for k in range(cs.nb_constraints):
        W[cs.O[k][0][0]] = ev(cs.L[k]) * ev(cs.R[k]) % q
    r, s = rnd.randrange(q), rnd.randrange(q)
    fold = H.fold_challenge([W[cm["commitment_index"]] for cm in cs.commitments], q) if ncommit > 1 else None
    oproof = OG.prove(cs, pk, W, r, s, cx, fold_challenge=fold)
    public = W[1:cs.nb_public]
    assert OG.verify(ovk, oproof, public, cx, kind)

    vk = vk_from_oracle(ovk, L.id)
    proof = proof_from_oracle(oproof, L.id)
    opts = [prover.WithProverTargetSolidityVerifier()] if kind == "solidity" else []
    assert verifier.Verify(proof, vk, public, *opts) is None
    assert fake_gpu["pairing"] == (2 if ncommit else 1)
    # tampered proof, wrong public input, wrong hash, wrong witness size, wrong commitment count
    bad = dict(oproof)
    bad["Krs"] = cx.G1.add(oproof["Krs"], cx.g1)
    assert not verifier.verify(proof_from_oracle(bad, L.id), vk, public, *opts)
    wrong = list(public)
    wrong[0] = (wrong[0] + 1) % q
    assert not verifier.verify(proof, vk, wrong, *opts)
    assert OG.verify(ovk, oproof, wrong, cx, kind) is False
    if ncommit:
        other = [] if kind == "solidity" else [prover.WithProverTargetSolidityVerifier()]
        assert not verifier.verify(proof, vk, public, *other)
        bad = dict(oproof)
        bad["CommitmentPok"] = cx.G1.add(oproof["CommitmentPok"], cx.g1)
        with pytest.raises(verifier.VerificationError, match="proof of knowledge"):
            verifier.Verify(proof_from_oracle(bad, L.id), vk, public, *opts)
        short = proof_from_oracle(oproof, L.id)
        short.Commitments = short.Commitments[:-1]
        with pytest.raises(verifier.VerificationError, match="number of commitments"):
            verifier.Verify(short, vk, public, *opts)
    # Bs on the twist but outside the order-r subgroup (gnark: proof.isValid())
    rogue = proof_from_oracle(oproof, L.id)
    rogue.Bs = L.enc_affine([_twist_point_outside_subgroup(cx)], 2)
    with pytest.raises(verifier.VerificationError, match="subgroup"):
        verifier.Verify(rogue, vk, public, *opts)
    with pytest.raises(verifier.VerificationError, match="invalid witness size"):
        verifier.Verify(proof, vk, public[:-1], *opts)
    other_curve = proof_from_oracle(oproof, L.id)
    other_curve.curve_id = 3 if L.id != 3 else 1
    with pytest.raises(verifier.VerificationError, match="different curves"):
        verifier.Verify(other_curve, vk, public, *opts)


def test_pairing_check_refuses_to_run_without_the_gpu():
    """No CPU fallback: without a device the C-ABI call fails loudly."""
    from davinci_node_b200 import capi, verifier
    from davinci_node_b200.layout import Layout
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    L = Layout("bn254")
    with pytest.raises(capi.B200Error):
        verifier.pairing_check(L.id, [np.zeros(L.affine_bytes(1), dtype=np.uint8)], [np.zeros(L.affine_bytes(2), dtype=np.uint8)])
