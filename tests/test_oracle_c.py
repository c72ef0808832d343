"""CPU: the C++ CPU restatement (oracle/c/oracle.cpp, the cpu_baseline 'port') must agree bit-for-bit
with the big-int oracle on MSM (G1/G2, index maps), FFT variants, computeH and the prove schedule."""
import ctypes as C
import random

import numpy as np
import pytest

from davinci_node_b200.layout import Layout
from oracle import cport
from oracle import curve as OC
from oracle import groth16 as OG
from oracle import ntt as N
from oracle import params as OP

CURVES = ["bn254", "bls12_377", "bls12_381", "bw6_761"]


def _rand_points(cx, group, n, rnd):
    G, g = cx.group(group), cx.gen(group)
    base = [G.mul(g, rnd.randrange(1, cx.r)) for _ in range(6)]
    pts = list(base)
    while len(pts) < n:
        pts.append(G.add(pts[rnd.randrange(len(pts))], base[rnd.randrange(6)]))
    return pts[:n]


@pytest.mark.parametrize("cname", CURVES)
@pytest.mark.parametrize("group", [1, 2])
def test_c_msm(cname, group):
    lib = cport.lib()
    L = Layout(cname)
    cx = OC.ctx(cname)
    rnd = random.Random(21 + group)
    n = 150 if group == 1 else 60
    pts = _rand_points(cx, group, n, rnd)
    pts[3] = None
    sc = [rnd.randrange(cx.r) for _ in range(n)]
    sc[0], sc[1], sc[2] = 0, 1, cx.r - 1
    ep, es = L.enc_affine(pts, group), L.enc_fr(sc)
    out = np.zeros(L.affine_bytes(group), dtype=np.uint8)
    assert lib.oc_msm(L.id, group, cport.p(ep), cport.p(es), n, None, cport.p(out), 0) == 0
    assert L.dec_affine(out, group)[0] == cx.group(group).msm(pts, sc)
    # index map: reversed order with a skipped scalar
    m = np.arange(n - 1, -1, -1, dtype=np.uint32)
    m[5] = 0xFFFFFFFF
    assert lib.oc_msm(L.id, group, cport.p(ep), cport.p(es), n, cport.p(m), cport.p(out), 2) == 0
    want = cx.group(group).msm([pts[n - 1 - i] for i in range(n) if i != 5], [sc[i] for i in range(n) if i != 5])
    assert L.dec_affine(out, group)[0] == want


@pytest.mark.parametrize("cname", CURVES)
def test_c_fft_and_h(cname):
    lib = cport.lib()
    L = Layout(cname)
    c = OP.CURVES[cname]
    q = c.r
    logn = 6
    n = 1 << logn
    dom = N.Domain(c, n)
    rnd = random.Random(31)
    a = [rnd.randrange(q) for _ in range(n)]
    w, g = L.enc_fr([dom.omega]), L.enc_fr([dom.g])
    for inv in (0, 1):
        for dit in (0, 1):
            for coset in (0, 1):
                buf = L.enc_fr(a)
                assert lib.oc_fft(L.id, cport.p(buf), logn, cport.p(w), cport.p(g), inv, dit, coset, 0) == 0
                assert L.dec_fr(buf) == N.fft(a, dom, inverse=bool(inv), dit=bool(dit), coset=bool(coset))
    aa = [rnd.randrange(q) for _ in range(n - 2)] + [0, 0]
    bb = [rnd.randrange(q) for _ in range(n - 2)] + [0, 0]
    cc = [x * y % q for x, y in zip(aa, bb)]
    ba, bb_, bc = L.enc_fr(aa), L.enc_fr(bb), L.enc_fr(cc)
    assert lib.oc_compute_h(L.id, cport.p(ba), cport.p(bb_), cport.p(bc), logn, cport.p(w), cport.p(g), 0) == 0
    assert L.dec_fr(ba) == N.compute_h(aa, bb, cc, dom)


def test_c_prove_matches_python_oracle():
    """The C prove schedule (extended arrays + index maps, as the product lays the key out) against
    the Python restatement of gnark's Prove."""
    from cpu_prove_util import c_oracle_prove
    cname = "bls12_377"
    cx = OC.ctx(cname)
    q = cx.r
    rnd = random.Random(41)
    cs, W = OG.synthetic_circuit(40, 4, q, seed=9, n_commit=1, n_private_committed=3)
    tox = OG.Toxic(*(rnd.randrange(1, q) for _ in range(5)), sigmas=[rnd.randrange(1, q)])
    pk, ex = OG.setup(cs, cx, tox)
    r, s = rnd.randrange(q), rnd.randrange(q)
    want = OG.prove(cs, pk, W, r, s, cx)
    got = c_oracle_prove(cname, cs, pk, W, r, s)
    assert got["Ar"] == want["Ar"] and got["Bs"] == want["Bs"] and got["Krs"] == want["Krs"]
