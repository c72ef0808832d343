"""GPU parity: Pippenger MSM (G1 and G2, all four curves) vs the oracle's definition
sum_i [s_i] P_i, bit-exact on the affine result, through the C ABI."""
import random

import numpy as np
import pytest

from oracle import curve as ocurve

pytestmark = pytest.mark.gpu

CURVES = ["bn254", "bls12_377", "bls12_381", "bw6_761"]


@pytest.fixture(scope="module")
def env():
    from davinci_node_b200 import capi, layout
    capi.init()
    return capi, layout


def gpu_msm_host(capi, L, group, pts, scalars):
    """through the host-pointer entry point (what the Go shim calls)"""
    ep = L.enc_affine(pts, group)
    es = L.enc_fr(scalars)
    out = np.zeros(L.affine_bytes(group), dtype=np.uint8)
    capi.check(capi.lib.b200_msm(L.id, group, ep.ctypes.data, es.ctypes.data, len(pts), out.ctypes.data, 0))
    return L.dec_affine(out, group)[0]


def gpu_msm_dev(capi, L, group, pts, scalars, c=0):
    from gpu_util import to_dev, dev_empty, ptr, stream, sync
    dp = to_dev(L.enc_affine(pts, group))
    ds = to_dev(L.enc_fr(scalars))
    ox = dev_empty(L.xyzz_bytes(group))
    oa = dev_empty(L.affine_bytes(group))
    capi.check(capi.lib.b200_msm_dev(L.id, group, ptr(dp), ptr(ds), len(pts), ptr(ox), c, stream()))
    capi.check(capi.lib.b200_to_affine_dev(L.id, group, ptr(ox), ptr(oa), 1, stream()))
    sync()
    return L.dec_affine(oa.cpu().numpy(), group)[0]


@pytest.mark.parametrize("cname", CURVES)
@pytest.mark.parametrize("group", [1, 2])
def test_msm_random(env, cname, group):
    from gpu_util import rand_points
    capi, layout = env
    L = layout.Layout(cname)
    cx = ocurve.ctx(cname)
    G = cx.group(group)
    rnd = random.Random(99 + group)
    sizes = [1, 2, 33, 700] if group == 1 else [1, 33, 300]
    for n in sizes:
        pts = rand_points(cx, group, n, rnd)
        sc = [rnd.randrange(cx.r) for _ in range(n)]
        want = G.msm(pts, sc)
        assert gpu_msm_host(capi, L, group, pts, sc) == want, (cname, group, n)
    # empty input -> infinity
    assert gpu_msm_dev(capi, L, group, [], []) is None


@pytest.mark.parametrize("cname", CURVES)
def test_msm_edge_scalars_and_windows(env, cname):
    """0 / 1 / r-1 / small scalars, infinity points, repeated points, several window widths."""
    from gpu_util import rand_points
    capi, layout = env
    L = layout.Layout(cname)
    cx = ocurve.ctx(cname)
    G = cx.G1
    rnd = random.Random(5)
    n = 400
    pts = rand_points(cx, 1, n, rnd)
    pts[7] = None
    pts[8] = pts[9]
    sc = []
    for i in range(n):
        k = i % 8
        sc.append([0, 1, cx.r - 1, 2, rnd.randrange(1 << 64), rnd.randrange(cx.r), (1 << (cx.r.bit_length() - 1)),
                   cx.r - 2][k])
    want = G.msm(pts, sc)
    for c in (0, 3, 8, 13, 16):
        assert gpu_msm_dev(capi, L, 1, pts, sc, c) == want, (cname, c)


@pytest.mark.parametrize("cname", ["bls12_377", "bn254"])
def test_msm_skewed_buckets_overflow_path(env, cname):
    """All scalars equal / witness-like 0-1 heavy mix: one bucket per window receives everything,
    which exercises the oversized-bucket task split and block merge."""
    from gpu_util import rand_points
    capi, layout = env
    L = layout.Layout(cname)
    cx = ocurve.ctx(cname)
    G = cx.G1
    rnd = random.Random(6)
    n = 3000
    pts = rand_points(cx, 1, n, rnd)
    k = rnd.randrange(cx.r)
    want = G.mul(G.sum(pts), k)
    assert gpu_msm_dev(capi, L, 1, pts, [k] * n, 8) == want
    sc = [1 if rnd.random() < 0.5 else (0 if rnd.random() < 0.5 else rnd.randrange(1 << 64)) for _ in range(n)]
    assert gpu_msm_dev(capi, L, 1, pts, sc, 0) == G.msm(pts, sc)


@pytest.mark.parametrize("cname,group", [("bls12_377", 1), ("bls12_377", 2), ("bn254", 1), ("bw6_761", 1), ("bls12_381", 1)])
def test_msm_table_mode(env, cname, group):
    """Base set registered once (precomputed window multiples), with and without an index map,
    several window widths: same affine result as the oracle's definition."""
    import ctypes as C
    from gpu_util import rand_points, to_dev, dev_empty, ptr, stream, sync
    capi, layout = env
    L = layout.Layout(cname)
    cx = ocurve.ctx(cname)
    G = cx.group(group)
    rnd = random.Random(17 + group)
    n = 260
    pts = rand_points(cx, group, n, rnd)
    pts[5] = None
    sc = [rnd.randrange(cx.r) for _ in range(n)]
    sc[0], sc[1], sc[2], sc[3] = 0, 1, cx.r - 1, rnd.randrange(1 << 64)
    want = G.msm(pts, sc)
    dp = to_dev(L.enc_affine(pts, group))
    ds = to_dev(L.enc_fr(sc))
    for c in (0, 5, 11):
        h = C.c_uint64(0)
        capi.check(capi.lib.b200_bases_create_dev(L.id, group, ptr(dp), n, c, C.byref(h), stream()))
        try:
            ox, oa = dev_empty(L.xyzz_bytes(group)), dev_empty(L.affine_bytes(group))
            capi.check(capi.lib.b200_msm_bases_dev(h.value, ptr(ds), n, None, ptr(ox), stream()))
            capi.check(capi.lib.b200_to_affine_dev(L.id, group, ptr(ox), ptr(oa), 1, stream()))
            sync()
            assert L.dec_affine(oa.cpu().numpy(), group)[0] == want, (cname, group, c)
            # index map: reversed bases, one scalar skipped, fewer scalars than bases
            k = n - 7
            m = np.arange(n - 1, n - 1 - k, -1, dtype=np.uint32)
            m[9] = 0xFFFFFFFF
            dm = to_dev(m.view(np.uint8))
            capi.check(capi.lib.b200_msm_bases_dev(h.value, ptr(ds), k, ptr(dm), ptr(ox), stream()))
            capi.check(capi.lib.b200_to_affine_dev(L.id, group, ptr(ox), ptr(oa), 1, stream()))
            sync()
            want2 = G.msm([pts[n - 1 - i] for i in range(k) if i != 9], [sc[i] for i in range(k) if i != 9])
            assert L.dec_affine(oa.cpu().numpy(), group)[0] == want2
        finally:
            capi.check(capi.lib.b200_bases_release(h.value))


@pytest.mark.parametrize("levels", [1, 2, 3])
@pytest.mark.parametrize("cname,group", [("bls12_377", 1), ("bls12_377", 2), ("bn254", 1), ("bw6_761", 1)])
def test_msm_affine_pre_reduction(env, cname, group, levels, monkeypatch):
    """Table-mode MSM with the sorted entries halved `levels` times by batched affine additions (shared inversion)
    before the bucket accumulation: same result as the oracle, including repeated points (doubling inside a pair),
    P + (-P) pairs, infinity points, a heavily loaded bucket and buckets shorter than the padding."""
    import ctypes as C
    from gpu_util import rand_points, to_dev, dev_empty, ptr, stream, sync
    capi, layout = env
    L = layout.Layout(cname)
    cx = ocurve.ctx(cname)
    G = cx.group(group)
    rnd = random.Random(100 * levels + group)
    n = 700
    pts = rand_points(cx, group, 40, rnd)
    pts = [pts[rnd.randrange(40)] for _ in range(n)]          # many repeated points
    pts[11] = None
    sc = [rnd.randrange(cx.r) for _ in range(n)]
    for i in range(0, 200):                                     # equal scalars on repeated points: P + P and P + (-P)
        sc[i] = 5 if i % 3 else cx.r - 5
    for i in range(200, 300):
        sc[i] = 1
    sc[300], sc[301] = 0, cx.r - 1
    want = G.msm(pts, sc)
    dp = to_dev(L.enc_affine(pts, group))
    ds = to_dev(L.enc_fr(sc))
    monkeypatch.setenv("B200_MSM_PRE", str(levels))
    for c in (4, 9):
        h = C.c_uint64(0)
        capi.check(capi.lib.b200_bases_create_dev(L.id, group, ptr(dp), n, c, C.byref(h), stream()))
        try:
            ox, oa = dev_empty(L.xyzz_bytes(group)), dev_empty(L.affine_bytes(group))
            capi.check(capi.lib.b200_msm_bases_dev(h.value, ptr(ds), n, None, ptr(ox), stream()))
            capi.check(capi.lib.b200_to_affine_dev(L.id, group, ptr(ox), ptr(oa), 1, stream()))
            sync()
            assert L.dec_affine(oa.cpu().numpy(), group)[0] == want, (cname, group, c, levels)
        finally:
            capi.check(capi.lib.b200_bases_release(h.value))
