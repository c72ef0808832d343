"""GPU point (de)compression for every curve / group and the gnark proving-key stream reader / writer
(davinci-node_b200/artifacts.py = pk.UnsafeReadFrom / pk.WriteTo mirrors, circuits/artifacts.go:391-406), against the
oracle's independent big-int serializer (oracle/serialize.py) and, for BLS12-381, the reference's own SRS bytes."""
import os
import random

import numpy as np
import pytest

from oracle import curve as OC
from oracle import groth16 as OG
from oracle import serialize as OS

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
CURVES = ["bn254", "bls12_377", "bls12_381", "bw6_761"]


@pytest.fixture(scope="module")
def env():
    from davinci_node_b200 import artifacts, capi, layout
    capi.init()
    return artifacts, capi, layout


@pytest.mark.parametrize("group", [1, 2])
@pytest.mark.parametrize("cname", CURVES)
def test_points_roundtrip(env, cname, group):
    from gpu_util import rand_points
    art, capi, layout = env
    cx = OC.ctx(cname)
    L = layout.Layout(cname)
    rnd = random.Random(17 + group)
    G = cx.group(group)
    pts = rand_points(cx, group, 40, rnd)
    pts += [G.neg(q) for q in pts[:10]] + [None, cx.gen(group), G.neg(cx.gen(group))]
    raw = b"".join(OS.compress_point(cx, group, q) for q in pts)
    assert len(raw) == len(pts) * art.compressed_bytes(L, group)
    assert capi.lib.b200_compressed_bytes(L.id, group) == art.compressed_bytes(L, group)
    aff = art.decompress_points(L, group, raw, len(pts))
    assert L.dec_affine(aff, group) == pts
    assert art.compress_points(L, group, L.enc_affine(pts, group)) == raw


@pytest.mark.parametrize("cname", CURVES)
def test_field_sqrt(env, cname):
    """Tonelli-Shanks in Fp and the norm method in Fp2, including the purely real / purely imaginary roots of real
    inputs and non-squares (the decompression kernels' building block)."""
    import ctypes as C
    import torch
    art, capi, layout = env
    cx = OC.ctx(cname)
    L = layout.Layout(cname)
    rnd = random.Random(5)
    p = cx.p
    st = torch.cuda.current_stream().cuda_stream

    def run(field, vals_flat, n):
        d = torch.from_numpy(L.enc_fp(vals_flat)).cuda()
        out = torch.empty_like(d)
        capi.check(capi.lib.b200_dbg_field_op_dev(L.id, field, 8, d.data_ptr(), None, out.data_ptr(), n, st))
        return L.dec_fp(out.cpu().numpy())

    xs = [rnd.randrange(1, p) for _ in range(24)]
    sq = [x * x % p for x in xs]
    nonsq = [v for v in (rnd.randrange(1, p) for _ in range(40)) if pow(v, (p - 1) // 2, p) == p - 1][:8]
    got = run(0, sq + nonsq + [0], len(sq) + len(nonsq) + 1)
    for g, s_ in zip(got[:len(sq)], sq):
        assert g * g % p == s_
    assert all(g == 0 for g in got[len(sq):])
    if cx.c.g2_degree == 2:
        F2 = cx.F2
        els = [(rnd.randrange(p), rnd.randrange(1, p)) for _ in range(12)] + [(rnd.randrange(1, p), 0), (0, rnd.randrange(1, p))]
        squares = [F2.sqr(e) for e in els]
        # a non-square of Fp2: u * (square) is a square iff u is; pick by Euler's criterion through the norm
        flat = [c for e in squares for c in e]
        got = run(2, flat, len(squares))
        for k, s_ in enumerate(squares):
            root = (got[2 * k], got[2 * k + 1])
            assert F2.sqr(root) == s_, k
        bad = []
        while len(bad) < 4:
            e = (rnd.randrange(p), rnd.randrange(p))
            if F2.sqrt(e) is None:
                bad.append(e)
        got = run(2, [c for e in bad for c in e], len(bad))
        assert all(v == 0 for v in got)


def test_srs_bytes_from_the_reference(env):
    """BLS12-381 G1 and G2 decompression against bytes the reference ships: the ceremony's monomial points and the
    65 G2 points (kzg_trusted_setup.txt; the first two also embedded in crypto/blobs/kzg.go:26-45)."""
    from oracle import kzg as OK
    art, capi, layout = env
    L = layout.Layout("bls12_381")
    raw1 = open(os.path.join(GOLD, "kzg_g1_monomial_64.bin"), "rb").read()
    raw2 = open(os.path.join(GOLD, "kzg_g2_monomial.bin"), "rb").read()
    g1 = L.dec_affine(art.decompress_points(L, 1, raw1, 64), 1)
    g2 = L.dec_affine(art.decompress_points(L, 2, raw2, 65), 2)
    assert g1 == [OK.g1_decompress(raw1[48 * j:48 * (j + 1)]) for j in range(64)]
    assert g2 == [OK.g2_decompress(raw2[96 * j:96 * (j + 1)]) for j in range(65)]
    assert art.compress_points(L, 2, L.enc_affine(g2, 2)) == raw2


@pytest.mark.parametrize("cname", ["bn254", "bls12_377", "bw6_761"])
def test_proving_key_stream_roundtrip_and_prove(env, cname):
    """oracle key -> oracle serializer -> read_proving_key (GPU decompression) == the key; write_proving_key gives the
    same bytes back; a proof made with the key that went through the file format is bit-identical to the oracle's."""
    from oracle_bridge import ccs_from_oracle, pk_from_oracle
    from davinci_node_b200 import gnark_types as T, prover
    art, capi, layout = env
    cx = OC.ctx(cname)
    q = cx.r
    L = layout.Layout(cname)
    rnd = random.Random(41)
    cs, W0 = OG.synthetic_circuit(30, 4, q, seed=6, n_commit=1, n_private_committed=3)
    tox = OG.Toxic(*(rnd.randrange(1, q) for _ in range(5)), sigmas=[rnd.randrange(1, q)])
    opk, ex = OG.setup(cs, cx, tox)
    stream = OS.write_proving_key(cx, opk)
    pk = art.read_proving_key(stream, L.id)
    want = pk_from_oracle(opk, L.id)
    for name in ("g1_alpha", "g1_beta", "g1_delta", "g1_A", "g1_B", "g1_Z", "g1_K", "g2_beta", "g2_delta", "g2_B",
                 "infinity_a", "infinity_b", "domain_generator", "domain_coset_gen"):
        assert np.array_equal(getattr(pk, name), getattr(want, name)), name
    assert pk.domain_cardinality == want.domain_cardinality
    assert np.array_equal(pk.commitment_keys[0]["BasisExpSigma"], want.commitment_keys[0]["BasisExpSigma"])
    assert art.write_proving_key(pk) == stream
    ccs = ccs_from_oracle(cs, L.id)
    w = T.Witness(L.id, W0[1:cs.nb_public], W0[cs.nb_public:cs.nb_public + ccs.nb_secret])
    r, s = rnd.randrange(q), rnd.randrange(q)
    prover.SetRandomness(lambda cid: (r, s))
    try:
        proof = prover.ProveWithWitness(L.id, ccs, pk, w)
        sol = ccs.solve(w, lambda i, v: proof.Commitments[i])
        wantp = OG.prove(cs, opk, sol.values, r, s, cx)
        got = proof.points()
        assert got["Ar"] == wantp["Ar"] and got["Bs"] == wantp["Bs"] and got["Krs"] == wantp["Krs"]
    finally:
        prover.SetRandomness(None)
        prover.release_proving_key(pk)
    # corrupted streams are rejected, not mis-read
    bad = bytearray(stream)
    off = 8 + 5 * L.fr_bytes + 1
    bad[off] &= 0x1F if cname != "bn254" else 0x3F          # alpha: flags cleared = "uncompressed"
    with pytest.raises(art.ArtifactError):
        art.read_proving_key(bytes(bad), L.id)
    with pytest.raises(art.ArtifactError):
        art.read_proving_key(stream[:-3], L.id)
    with pytest.raises(art.ArtifactError):
        art.read_proving_key(stream + b"\x00", L.id)


@pytest.mark.parametrize("cname", CURVES)
def test_binary_gcd_inversion(env, cname):
    """FpT::inv_bin (the one-thread inversion of the proof's final affine normalisation) == a^-1, with 0 -> 0."""
    import torch
    art, capi, layout = env
    cx = OC.ctx(cname)
    L = layout.Layout(cname)
    rnd = random.Random(9)
    p = cx.p
    vals = [0, 1, 2, p - 1, p - 2, (p - 1) // 2, (p + 1) // 2] + [rnd.randrange(p) for _ in range(40)] + [1 << k for k in (31, 32, 63, 64, 200)]
    d = torch.from_numpy(L.enc_fp(vals)).cuda()
    out = torch.empty_like(d)
    capi.check(capi.lib.b200_dbg_field_op_dev(L.id, 0, 9, d.data_ptr(), None, out.data_ptr(), len(vals),
                                              torch.cuda.current_stream().cuda_stream))
    got = L.dec_fp(out.cpu().numpy())
    assert got == [pow(v, -1, p) if v else 0 for v in vals]
