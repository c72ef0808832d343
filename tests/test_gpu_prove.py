"""GPU parity of the full Groth16 prove through the reference-shaped interface
(davinci_node_b200.prover.ProveWithWitness == prover/prover_gpu.go:111): with (r, s) pinned the
proof is bit-identical to the oracle's restatement of gnark's Prove, and it satisfies the verifier
equation in the exponent."""
import random

import numpy as np
import pytest

from oracle import curve as OC
from oracle import groth16 as OG

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def env():
    from davinci_node_b200 import capi, layout, prover, gnark_types
    capi.init()
    return capi, layout, prover, gnark_types


CASES = [
    # curve, constraints, public wires, commitments, private committed, mix, commitment hash, public committed wires
    ("bls12_377", 60, 6, 1, 5, "witness", "default", 0),     # voteverifier-shaped: 6 public wires, 1 commitment
    ("bn254", 45, 9, 1, 4, "witness", "solidity", 1),        # statetransition-shaped: keccak target, input committed
    ("bw6_761", 37, 2, 1, 3, "uniform", "default", 0),       # aggregator-shaped
    ("bls12_377", 130, 3, 0, 0, "uniform", "default", 0),    # no commitment
    ("bn254", 40, 3, 2, 3, "witness", "default", 1),         # two commitments (folded proof of knowledge)
]


@pytest.mark.parametrize("cname,ncons,npub,ncommit,npc,mix,hash_kind,npubc", CASES)
def test_prove_bit_exact_and_verifies(env, cname, ncons, npub, ncommit, npc, mix, hash_kind, npubc):
    """With (r, s) pinned the GPU proof equals the oracle's restatement of gnark's Prove bit for bit; pinned or not it
    passes gnark's verifier (pairings, real commitment hashes) and - BN254 with one commitment - the port of the
    Solidity verifier the reference deploys."""
    from oracle_bridge import ccs_from_oracle, pk_from_oracle
    capi, layout, prover, T = env
    cx = OC.ctx(cname)
    q = cx.r
    L = layout.Layout(cname)
    rnd = random.Random(hash((cname, ncons)) & 0xFFFF)
    cs, W0 = OG.synthetic_circuit(ncons, npub, q, seed=ncons, n_commit=ncommit, n_private_committed=npc, mix=mix,
                                  n_public_committed=npubc)
    tox = OG.Toxic(*(rnd.randrange(1, q) for _ in range(5)), sigmas=[rnd.randrange(1, q) for _ in range(ncommit)])
    opk, ex = OG.setup(cs, cx, tox)
    vk = OG.verifying_key(cs, cx, tox, ex)
    ccs = ccs_from_oracle(cs, L.id)
    pk = pk_from_oracle(opk, L.id)
    w = T.Witness(L.id, W0[1:cs.nb_public], W0[cs.nb_public:cs.nb_public + ccs.nb_secret])
    opts = [prover.WithProverTargetSolidityVerifier()] if hash_kind == "solidity" else []
    r, s = rnd.randrange(q), rnd.randrange(q)
    prover.SetRandomness(lambda cid: (r, s))
    try:
        proof = prover.ProveWithWitness(L.id, ccs, pk, w, *opts)
        # the solver (host) fixed the commitment-wire values; replay the same assignment in the oracle
        sol = ccs.solve(w, lambda i, v: proof.Commitments[i], hash_kind)
        W = sol.values
        want = OG.prove(cs, opk, W, r, s, cx, fold_challenge=sol.fold_challenge)
        got = proof.points()
        assert got["Ar"] == want["Ar"]
        assert got["Bs"] == want["Bs"]
        assert got["Krs"] == want["Krs"]
        assert got["Commitments"] == want["Commitments"]
        if ncommit:
            assert got["CommitmentPok"] == want["CommitmentPok"]
        A, B, Cx = OG.proof_exponents(cs, ex, tox, W, r, s, q)
        assert got["Ar"] == cx.G1.mul(cx.g1, A) and got["Bs"] == cx.G2.mul(cx.g2, B)
        assert OG.verify_exponent(cs, ex, tox, W, A, B, Cx, q)
        public = W[1:cs.nb_public]
        assert OG.verify(vk, got, public, cx, hash_kind)
        # un-pinned randomness: a different proof that passes the same verifiers
        prover.SetRandomness(None)
        p2 = prover.ProveWithWitness(L.id, ccs, pk, w, *opts).points()
        assert p2["Ar"] != got["Ar"]
        assert OG.verify(vk, p2, public, cx, hash_kind)
        wrong = list(public)
        wrong[-1] = (wrong[-1] + 1) % q
        assert not OG.verify(vk, p2, wrong, cx, hash_kind)
        if cname == "bn254" and ncommit == 1 and hash_kind == "solidity":
            consts = OG.solidity_constants(vk, cx)
            Ar, Bs, Krs = p2["Ar"], p2["Bs"], p2["Krs"]
            proof8 = [Ar[0], Ar[1], Bs[0][1], Bs[0][0], Bs[1][1], Bs[1][0], Krs[0], Krs[1]]
            committed_inputs = [wire - 1 for wire in cs.commitments[0]["public_committed"]]
            assert OG.solidity_verify_proof(consts, proof8, p2["Commitments"][0], p2["CommitmentPok"], public,
                                            committed_inputs, cx)
            # the on-chain calldata (solidity/solidity.go:85-116 mirror) carries exactly these words
            from davinci_node_b200 import solidity
            prover.SetRandomness(lambda cid: (r, s))
            sp = solidity.Groth16CommitmentProof().FromGnarkProof(prover.ProveWithWitness(L.id, ccs, pk, w, *opts))
            prover.SetRandomness(None)
            data = sp.ABIEncode()
            assert len(data) == 384 and solidity.Groth16CommitmentProof.ABIDecode(data).words() == sp.words()
            wds = sp.words()
            assert OG.solidity_verify_proof(consts, wds[:8], wds[8:10], wds[10:12], public, committed_inputs, cx)
            assert not OG.solidity_verify_proof(consts, proof8, p2["Commitments"][0], p2["CommitmentPok"], wrong,
                                                committed_inputs, cx)
    finally:
        prover.SetRandomness(None)
        prover.release_proving_key(pk)


@pytest.mark.parametrize("case,stride", [(0, 2), (2, 3), (4, 2), (1, 32)])
def test_prove_with_table_stride(env, monkeypatch, case, stride):
    """HBM budget knob: with table stride s only every s-th window table is resident (s bucket sets per base set and a
    final Horner); the proof bytes do not change."""
    monkeypatch.setenv("B200_TABLE_STRIDE", str(stride))
    test_prove_bit_exact_and_verifies(env, *CASES[case])


def test_pk_table_budget_policy(env):
    """b200_set_pk_table_budget: a key registered under a tiny budget gets a large table stride (few resident tables),
    under a generous one stride 1; both prove the same bytes."""
    from oracle_bridge import ccs_from_oracle, pk_from_oracle
    capi, layout, prover, T = env
    cx = OC.ctx("bn254")
    q = cx.r
    L = layout.Layout("bn254")
    rnd = random.Random(77)
    cs, W0 = OG.synthetic_circuit(40, 4, q, seed=9, n_commit=1, n_private_committed=3)
    tox = OG.Toxic(*(rnd.randrange(1, q) for _ in range(5)), sigmas=[rnd.randrange(1, q)])
    opk, ex = OG.setup(cs, cx, tox)
    ccs = ccs_from_oracle(cs, L.id)
    w = T.Witness(L.id, W0[1:cs.nb_public], W0[cs.nb_public:cs.nb_public + ccs.nb_secret])
    r, s = rnd.randrange(q), rnd.randrange(q)
    prover.SetRandomness(lambda cid: (r, s))
    proofs, infos = [], []
    try:
        for budget in (1, 1 << 40):
            capi.check(capi.lib.b200_set_pk_table_budget(budget))
            pk = pk_from_oracle(opk, L.id)
            try:
                proofs.append(prover.ProveWithWitness(L.id, ccs, pk, w).points())
                infos.append(capi.pk_info(prover.register_proving_key(pk, ccs)))
            finally:
                prover.release_proving_key(pk)
    finally:
        capi.check(capi.lib.b200_set_pk_table_budget(0))
        prover.SetRandomness(None)
    assert infos[0]["table_stride"] > 1 and infos[1]["table_stride"] == 1
    assert infos[0]["table_bytes"] < infos[1]["table_bytes"]
    assert proofs[0] == proofs[1]


def test_prove_rejects_unsatisfied_witness(env):
    from oracle_bridge import ccs_from_oracle, pk_from_oracle
    capi, layout, prover, T = env
    cx = OC.ctx("bn254")
    L = layout.Layout("bn254")
    cs, W0 = OG.synthetic_circuit(10, 3, cx.r, seed=1)
    ccs = ccs_from_oracle(cs, L.id)
    # make constraint 3's output a constant wire so a wrong witness cannot be "solved around"
    ccs.O[3] = [(1, 1)]
    w = T.Witness(L.id, W0[1:cs.nb_public], W0[cs.nb_public:cs.nb_public + ccs.nb_secret])
    with pytest.raises(T.UnsatisfiedConstraintError):
        ccs.solve(w)


def test_cpu_prover_is_absent(env):
    capi, layout, prover, T = env
    with pytest.raises(prover.ProverError):
        prover.CPUProver(1, None, None, None)
    with pytest.raises(prover.ProverError):
        prover.CPUProverWithWitness(1, None, None, None)


@pytest.mark.parametrize("cname", ["bn254", "bls12_377"])
def test_setup_mirror_matches_oracle_and_proves(env, cname):
    """prover.Setup mirror (prover/setup.go:15): with the toxic waste pinned the GPU-built key equals
    the oracle's key point for point, and a proof made with it satisfies the verifier equation."""
    from oracle_bridge import ccs_from_oracle, pk_from_oracle
    capi, layout, prover, T = env
    cx = OC.ctx(cname)
    q = cx.r
    L = layout.Layout(cname)
    rnd = random.Random(99)
    cs, W0 = OG.synthetic_circuit(33, 4, q, seed=5, n_commit=1, n_private_committed=3)
    tox = OG.Toxic(*(rnd.randrange(1, q) for _ in range(5)), sigmas=[rnd.randrange(1, q)])
    opk, ex = OG.setup(cs, cx, tox)
    want = pk_from_oracle(opk, L.id)
    ccs = ccs_from_oracle(cs, L.id)
    prover.SetSetupRandomness(lambda cid, k: dict(tau=tox.tau, alpha=tox.alpha, beta=tox.beta, gamma=tox.gamma,
                                                  delta=tox.delta, sigmas=tox.sigmas))
    try:
        pk, vk = prover.Setup(ccs)
    finally:
        prover.SetSetupRandomness(None)
    for name in ("g1_alpha", "g1_beta", "g1_delta", "g1_A", "g1_B", "g1_Z", "g1_K", "g2_beta", "g2_delta", "g2_B",
                 "infinity_a", "infinity_b", "domain_generator", "domain_coset_gen"):
        assert np.array_equal(getattr(pk, name), getattr(want, name)), name
    assert np.array_equal(pk.commitment_keys[0]["Basis"], want.commitment_keys[0]["Basis"])
    assert np.array_equal(pk.commitment_keys[0]["BasisExpSigma"], want.commitment_keys[0]["BasisExpSigma"])
    assert L.dec_affine(vk.g1_K, 1) == [cx.G1.mul(cx.g1, k) for k in ex["vk_K"]]
    assert L.dec_affine(vk.g2_gamma, 2)[0] == cx.G2.mul(cx.g2, tox.gamma)
    # prove with the GPU-built key
    w = T.Witness(L.id, W0[1:cs.nb_public], W0[cs.nb_public:cs.nb_public + ccs.nb_secret])
    r, s = rnd.randrange(q), rnd.randrange(q)
    prover.SetRandomness(lambda cid: (r, s))
    try:
        proof = prover.ProveWithWitness(L.id, ccs, pk, w)
        sol = ccs.solve(w, lambda i, v: proof.Commitments[i])
        A, B, Cx = OG.proof_exponents(cs, ex, tox, sol.values, r, s, q)
        got = proof.points()
        assert got["Ar"] == cx.G1.mul(cx.g1, A) and got["Krs"] == cx.G1.mul(cx.g1, Cx)
        assert OG.verify_exponent(cs, ex, tox, sol.values, A, B, Cx, q)
    finally:
        prover.SetRandomness(None)
        prover.release_proving_key(pk)


@pytest.mark.parametrize("cname,parts", [("bls12_377", 2), ("bw6_761", 3)])
def test_range_split_prove_single_gpu(env, cname, parts):
    """Range-split prove with all key slices on ONE device: per-slice partial sums,
    concatenated as an all-gather would deliver them, assembled - bit-identical to the oracle's proof.
    (tests/test_gpu_multi.py runs the same through NCCL on 2+ GPUs.)"""
    import torch
    from oracle_bridge import ccs_from_oracle, pk_from_oracle
    from davinci_node_b200 import multi
    capi, layout, prover, T = env
    cx = OC.ctx(cname)
    q = cx.r
    L = layout.Layout(cname)
    rnd = random.Random(2024 + parts)
    cs, W = OG.synthetic_circuit(50, 5, q, seed=parts, n_commit=1, n_private_committed=4)
    tox = OG.Toxic(*(rnd.randrange(1, q) for _ in range(5)), sigmas=[rnd.randrange(1, q)])
    opk, ex = OG.setup(cs, cx, tox)
    ccs, pk = ccs_from_oracle(cs, L.id), pk_from_oracle(opk, L.id)
    r, s = rnd.randrange(q), rnd.randrange(q)
    want = OG.prove(cs, opk, W, r, s, cx)
    a, b, c = OG.constraint_values(cs, W, q)
    dev = lambda v: torch.from_numpy(L.enc_fr(v)).cuda()
    Wd, ad, bd, cd = dev(W), dev(a), dev(b), dev(c)
    committed = cs.commitments[0]["private_committed"]
    pc = [(dev([W[i] for i in committed]), len(committed))]
    partials, subs, handles, infos = [], [], [], []
    try:
        for rank in range(parts):
            sub, sub_ccs, info = multi.slice_proving_key(pk, ccs, parts, rank)
            h = multi.register_key_slice(sub, sub_ccs, info)
            subs.append(sub)
            handles.append(h)
            infos.append(info)
            partials.append(multi.prove_partial(h, L, info, Wd, ad, bd, cd, len(a), r, s, pc))
        got = multi.assemble(L, torch.cat(partials), parts, r, s, have_pok=True)
        assert L.dec_affine(got["Ar"], 1)[0] == want["Ar"]
        assert L.dec_affine(got["Bs"], 2)[0] == want["Bs"]
        assert L.dec_affine(got["Krs"], 1)[0] == want["Krs"]
        assert L.dec_affine(got["CommitmentPok"], 1)[0] == want["CommitmentPok"]
        # sharded quotient: a, b, c arrive as coset evaluations (b200_pk_coset_evals_dev, abc_form = 1) - same proof
        n = pk.domain_cardinality
        pad = lambda v: torch.cat([v, torch.zeros((n - len(a)) * L.fr_bytes, dtype=torch.uint8, device="cuda")])
        ea, eb, ec = pad(ad), pad(bd), pad(cd)
        for v in (ea, eb, ec):
            multi.coset_evals_inplace(handles[0], v)
        partials2 = [multi.prove_partial(h, L, info, Wd, ea, eb, ec, n, r, s, pc, abc_form=1)
                     for h, info in zip(handles, infos)]
        got2 = multi.assemble(L, torch.cat(partials2), parts, r, s, have_pok=True)
        for key in ("Ar", "Bs", "Krs", "CommitmentPok"):
            assert np.array_equal(got2[key], got[key]), key
    finally:
        for sub in subs:
            prover.release_proving_key(sub)
