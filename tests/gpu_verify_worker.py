"""Worker of tests/test_gpu_zz_verify.py (own process: a fault in the new pairing kernels must not poison the CUDA context
of the rest of the GPU suite).  Prints one JSON line {"ok": true, ...} or raises."""
import json
import os
import random
import sys
import time

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
sys.path.insert(0, HERE)


def main():
    import numpy as np
    from davinci_node_b200 import capi, gnark_types as T, layout, prover, verifier
    from oracle import curve as OC
    from oracle import groth16 as OG
    from oracle import pairing
    from oracle_bridge import ccs_from_oracle, pk_from_oracle, vk_from_oracle
    capi.init(1)
    out = {"pairing_ms": {}, "verify_ms": {}}

    # ---- 1. the pairing itself: value bit-identical to the oracle's reduced Tate pairing, product checks, edge cases
    for cname in ("bn254", "bls12_377", "bls12_381", "bw6_761"):
        pr = pairing.get(cname)
        cx = pr.cx
        L = layout.Layout(cname)
        a, b = 0x1234567, 0xfedcba9
        P, Q = cx.G1.mul(cx.g1, a), cx.G2.mul(cx.g2, b)
        e1 = lambda pt: L.enc_affine([pt], 1)
        e2 = lambda pt: L.enc_affine([pt], 2)
        t0 = time.time()
        ok, gt = verifier.pairing_check(L.id, [e1(P)], [e2(Q)], want_gt=True)
        out["pairing_ms"][cname] = round(1e3 * (time.time() - t0), 1)
        assert gt == pr.pair(P, Q), cname + ": reduced pairing value differs from the oracle"
        assert not ok
        good = [(cx.G1.mul(cx.g1, a), cx.g2), (cx.G1.neg(cx.g1), cx.G2.mul(cx.g2, a))]
        bad = [(cx.G1.mul(cx.g1, a), cx.g2), (cx.G1.neg(cx.g1), cx.G2.mul(cx.g2, a + 1))]
        chk = lambda pairs: verifier.pairing_check(L.id, [e1(p) for p, _ in pairs], [e2(q) for _, q in pairs])
        assert chk(good) and pr.product_is_one(good), cname
        assert not chk(bad), cname
        assert chk(good + [(None, cx.g2), (cx.g1, None)]), cname
        assert chk([]), cname
        if cname != "bn254":          # BN254's G1 has cofactor 1
            rnd = random.Random(3)
            while True:
                x = rnd.randrange(cx.p)
                y = OC.sqrt_mod((x * x * x + cx.G1.b) % cx.p, cx.p)
                if y is not None and cx.G1.mul((x, y), cx.r) is not None:
                    break
            try:
                chk([((x, y), cx.g2)])
                raise AssertionError(cname + ": a point outside the subgroup was accepted")
            except capi.B200Error as e:
                assert "subgroup" in str(e)

    # ---- 2. Groth16: GPU proofs verified on the GPU, decisions equal to the gnark-verifier restatement
    cases = [("bn254", 40, 5, 1, 6, 2, "solidity"), ("bls12_377", 40, 5, 2, 5, 1, "default"), ("bw6_761", 24, 4, 1, 4, 0, "default"),
             ("bn254", 24, 4, 0, 0, 0, "default")]
    for cname, ncons, npub, ncommit, npc, npubc, kind in cases:
        cx = OC.ctx(cname)
        q = cx.r
        L = layout.Layout(cname)
        rnd = random.Random(ncons * 7 + ncommit)
        cs, W0 = OG.synthetic_circuit(ncons, npub, q, seed=ncons, n_commit=ncommit, n_private_committed=npc,
                                      n_public_committed=npubc)
        tox = OG.Toxic(*(rnd.randrange(1, q) for _ in range(5)), sigmas=[rnd.randrange(1, q) for _ in range(ncommit)])
        opk, ex = OG.setup(cs, cx, tox)
        ovk = OG.verifying_key(cs, cx, tox, ex)
        ccs, pk, vk = ccs_from_oracle(cs, L.id), pk_from_oracle(opk, L.id), vk_from_oracle(ovk, L.id)
        w = T.Witness(L.id, W0[1:cs.nb_public], W0[cs.nb_public:cs.nb_public + ccs.nb_secret])
        opts = [prover.WithProverTargetSolidityVerifier()] if kind == "solidity" else []
        try:
            proof = prover.ProveWithWitness(L.id, ccs, pk, w, *opts)       # un-pinned randomness
            sol = ccs.solve(w, lambda i, v: proof.Commitments[i], kind)
            public = sol.values[1:cs.nb_public]
            assert OG.verify(ovk, proof.points(), public, cx, kind)
            t0 = time.time()
            verifier.Verify(proof, vk, public, *opts)
            out["verify_ms"]["%s/%d" % (cname, ncommit)] = round(1e3 * (time.time() - t0), 1)
            rejected = 0
            for idx in (0, -1):         # a public wire may be unused by the circuit: decisions must AGREE with the oracle
                wrong = list(public)
                wrong[idx] = (wrong[idx] + 1) % q
                want = OG.verify(ovk, proof.points(), wrong, cx, kind)
                assert verifier.verify(proof, vk, wrong, *opts) == want, (cname, ncommit, idx)
                rejected += not want
            assert rejected, "no wrong public input was rejected"
            bad = T.Proof(L.id)
            bad.Ar, bad.Bs, bad.Commitments, bad.CommitmentPok = proof.Ar, proof.Bs, proof.Commitments, proof.CommitmentPok
            bad.Krs = L.enc_affine([cx.G1.add(proof.points()["Krs"], cx.g1)], 1)
            assert not verifier.verify(bad, vk, public, *opts)
            # Bs on the twist but outside the order-r subgroup (gnark: proof.isValid())
            from test_verifier_cpu import _twist_point_outside_subgroup
            bad.Krs, bad.Bs = proof.Krs, L.enc_affine([_twist_point_outside_subgroup(cx)], 2)
            try:
                verifier.Verify(bad, vk, public, *opts)
                raise AssertionError("a Bs outside the subgroup was accepted")
            except verifier.VerificationError as e:
                assert "subgroup" in str(e)
        finally:
            prover.release_proving_key(pk)
    out["ok"] = True
    out["launches"] = int(capi.lib.b200_launch_count())
    print(json.dumps(out))


def _point_outside_subgroup(cx):
    from oracle import curve as OC
    rnd = random.Random(3)
    while True:
        x = rnd.randrange(cx.p)
        y = OC.sqrt_mod((x * x * x + cx.G1.b) % cx.p, cx.p)
        if y is not None and cx.G1.mul((x, y), cx.r) is not None:
            return (x, y)


BATCH_CASES = (("bn254", 96), ("bls12_377", 40), ("bw6_761", 40))


def main_batch(edge=False):
    """b200_pairing_check_batch: many independent checks, one thread each; decisions known by construction (and a few
    confirmed by the oracle), agreement with the single-check entry point, throughput printed."""
    import numpy as np
    from davinci_node_b200 import capi, layout, verifier
    from oracle import pairing
    capi.init(1)
    out = {"checks_per_s": {}}
    for cname, n_checks in BATCH_CASES:
        pr = pairing.get(cname)
        cx = pr.cx
        L = layout.Layout(cname)
        rnd = random.Random(5)
        e1 = lambda pt: L.enc_affine([pt], 1)
        e2 = lambda pt: L.enc_affine([pt], 2)
        g1s, g2s, want, pairs_of = [], [], [], []
        for c in range(n_checks):
            a = rnd.randrange(1, 1 << 40)
            good = c % 3 != 1
            pairs = [(cx.G1.mul(cx.g1, a), cx.g2), (cx.G1.neg(cx.g1), cx.G2.mul(cx.g2, a if good else a + 1))]
            if edge and c % 7 == 3:
                pairs[0], pairs[1] = (None, cx.g2), (cx.g1, None)        # infinities: the product is one
                good = True
            if edge and c % 11 == 5 and cname != "bn254":
                pairs[0] = (_point_outside_subgroup(cx), cx.g2)          # reported as None, the other checks unaffected
                good = None
            pairs_of.append(pairs)
            want.append(good)
            for P, Q in pairs:
                g1s.append(e1(P))
                g2s.append(e2(Q))
        t0 = time.time()
        got = verifier.pairing_check_batch(L.id, g1s, g2s, 2)
        dt = time.time() - t0
        out["checks_per_s"][cname] = round(n_checks / dt, 1)
        assert got == want, (cname, [i for i in range(n_checks) if got[i] != want[i]][:8])
        for c in (0, 1):              # the construction itself, confirmed by the oracle and by the single-check kernel
            if want[c] is None:
                continue
            assert pr.product_is_one(pairs_of[c]) == want[c]
            assert verifier.pairing_check(L.id, g1s[2 * c:2 * c + 2], g2s[2 * c:2 * c + 2]) == want[c]
        assert verifier.pairing_check_batch(L.id, [], [], 2) == []
    out["ok"] = True
    print(json.dumps(out))


if __name__ == "__main__":
    if "--batch-edge" in sys.argv:
        main_batch(edge=True)
    elif "--batch" in sys.argv:
        main_batch()
    else:
        main()
