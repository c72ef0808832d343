"""GPU parity: NTT (all gnark decimation / coset / inverse variants) and the fused quotient
computeH vs the big-int oracle, bit-exact, through the C ABI."""
import ctypes as C
import random

import numpy as np
import pytest

from oracle import ntt as N
from oracle import params as OP

pytestmark = pytest.mark.gpu

CURVES = ["bn254", "bls12_377", "bls12_381", "bw6_761"]


@pytest.fixture(scope="module")
def env():
    from davinci_node_b200 import capi, layout
    capi.init()
    return capi, layout


def make_domain(capi, L, dom):
    h = C.c_uint64(0)
    g = L.enc_fr([dom.omega])
    cg = L.enc_fr([dom.g])
    capi.check(capi.lib.b200_domain_create(L.id, dom.n, g.ctypes.data, cg.ctypes.data, C.byref(h)))
    return h.value


@pytest.mark.parametrize("cname", CURVES)
@pytest.mark.parametrize("logn", [1, 4, 9, 12, 14])
def test_ntt_variants(env, cname, logn):
    from gpu_util import to_dev, ptr, stream, sync
    capi, layout = env
    L = layout.Layout(cname)
    c = OP.CURVES[cname]
    n = 1 << logn
    dom = N.Domain(c, n)
    h = make_domain(capi, L, dom)
    rnd = random.Random(logn * 7 + 1)
    a = [rnd.randrange(c.r) for _ in range(n)]
    a[0], a[-1] = 0, c.r - 1
    try:
        variants = [(inv, dit, coset) for inv in (0, 1) for dit in (0, 1) for coset in (0, 1)]
        if logn >= 12:
            variants = [(0, 0, 0), (1, 0, 1), (0, 1, 1), (1, 1, 0)]
        for inv, dit, coset in variants:
            d = to_dev(L.enc_fr(a))
            capi.check(capi.lib.b200_ntt_dev(h, ptr(d), inv, dit, coset, stream()))
            sync()
            got = L.dec_fr(d.cpu().numpy())
            assert got == N.fft(a, dom, inverse=bool(inv), dit=bool(dit), coset=bool(coset)), (cname, logn, inv, dit, coset)
    finally:
        capi.check(capi.lib.b200_domain_release(h))


@pytest.mark.parametrize("cname", ["bn254", "bls12_377", "bw6_761"])
@pytest.mark.parametrize("logn", [3, 10, 13])
def test_compute_h(env, cname, logn):
    from gpu_util import to_dev, ptr, stream, sync
    capi, layout = env
    L = layout.Layout(cname)
    c = OP.CURVES[cname]
    q = c.r
    n = 1 << logn
    dom = N.Domain(c, n)
    h = make_domain(capi, L, dom)
    rnd = random.Random(logn)
    nc = n - 3
    a = [rnd.randrange(q) for _ in range(nc)]
    b = [rnd.randrange(q) for _ in range(nc)]
    cc = [x * y % q for x, y in zip(a, b)]
    pad = lambda v: v + [0] * (n - len(v))
    try:
        da, db, dc = (to_dev(L.enc_fr(pad(v))) for v in (a, b, cc))
        capi.check(capi.lib.b200_compute_h_dev(h, ptr(da), ptr(db), ptr(dc), stream()))
        sync()
        got = L.dec_fr(da.cpu().numpy())
        assert got == N.compute_h(a, b, cc, dom)
    finally:
        capi.check(capi.lib.b200_domain_release(h))


def test_ntt_round_trip_large(env):
    """2^20 BLS12-377: size-independent property  iFFT_DIT(FFT_DIF(a)) == a  (no big-int oracle)."""
    from gpu_util import to_dev, ptr, stream, sync
    capi, layout = env
    L = layout.Layout("bls12_377")
    c = OP.CURVES["bls12_377"]
    n = 1 << 20
    dom = N.Domain(c, n)
    h = make_domain(capi, L, dom)
    rng = np.random.default_rng(3)
    a = rng.integers(0, 1 << 62, size=(n, 4), dtype=np.uint64)
    a[:, 3] &= np.uint64((1 << 59) - 1)            # < r
    d = to_dev(a.view(np.uint8).reshape(-1))
    try:
        capi.check(capi.lib.b200_ntt_dev(h, ptr(d), 0, 0, 1, stream()))
        capi.check(capi.lib.b200_ntt_dev(h, ptr(d), 1, 1, 1, stream()))
        sync()
        assert np.array_equal(d.cpu().numpy().view(np.uint64).reshape(n, 4), a)
    finally:
        capi.check(capi.lib.b200_domain_release(h))
