"""GPU parity: generated Montgomery field kernels and the XYZZ group law vs the big-int oracle,
bit-exact, through the C ABI debug entry points (include/b200_groth16.h)."""
import random

import numpy as np
import pytest

from oracle import curve as ocurve
from oracle import params as OP

pytestmark = pytest.mark.gpu

CURVES = ["bn254", "bls12_377", "bls12_381", "bw6_761"]
OPS = {"add": 0, "sub": 1, "mul": 2, "sqr": 3, "from_mont": 4, "to_mont": 5, "inv": 6, "neg": 7}


@pytest.fixture(scope="module")
def env():
    from davinci_node_b200 import capi, layout
    capi.init()
    return capi, layout


def _run_field(capi, cid, field, op, a, b, nbytes_el, n):
    from gpu_util import to_dev, dev_empty, ptr, stream, sync
    da = to_dev(a)
    db = to_dev(b) if b is not None else None
    out = dev_empty(n * nbytes_el)
    capi.check(capi.lib.b200_dbg_field_op_dev(cid, field, op, ptr(da), ptr(db), ptr(out), n, stream()))
    sync()
    return out.cpu().numpy()


@pytest.mark.parametrize("cname", CURVES)
@pytest.mark.parametrize("which", ["fp", "fr"])
def test_field_ops(env, cname, which):
    capi, layout = env
    L = layout.Layout(cname)
    q = L.p if which == "fp" else L.r
    enc = L.enc_fp if which == "fp" else L.enc_fr
    dec = L.dec_fp if which == "fp" else L.dec_fr
    nb = L.fp_bytes if which == "fp" else L.fr_bytes
    fsel = 0 if which == "fp" else 1
    rnd = random.Random(hash((cname, which)) & 0xFFFF)
    edge = [0, 1, 2, q - 1, q - 2, (q - 1) // 2, (1 << (q.bit_length() - 1)), (1 << (8 * nb)) % q]
    A = [a for a in edge for _ in edge] + [rnd.randrange(q) for _ in range(300)]
    B = [b for _ in edge for b in edge] + [rnd.randrange(q) for _ in range(300)]
    n = len(A)
    ea, eb = enc(A), enc(B)
    got = dec(_run_field(capi, L.id, fsel, OPS["add"], ea, eb, nb, n))
    assert got == [(a + b) % q for a, b in zip(A, B)]
    got = dec(_run_field(capi, L.id, fsel, OPS["sub"], ea, eb, nb, n))
    assert got == [(a - b) % q for a, b in zip(A, B)]
    got = dec(_run_field(capi, L.id, fsel, OPS["mul"], ea, eb, nb, n))
    assert got == [(a * b) % q for a, b in zip(A, B)]
    got = dec(_run_field(capi, L.id, fsel, OPS["sqr"], ea, None, nb, n))
    assert got == [(a * a) % q for a in A]
    got = dec(_run_field(capi, L.id, fsel, OPS["neg"], ea, None, nb, n))
    assert got == [(-a) % q for a in A]
    got = dec(_run_field(capi, L.id, fsel, OPS["inv"], ea, None, nb, n))
    assert got == [pow(a, -1, q) if a else 0 for a in A]
    # from_mont / to_mont are checked on raw limbs
    raw = dec(_run_field(capi, L.id, fsel, OPS["from_mont"], ea, None, nb, n), mont=False)
    assert raw == A
    raw = dec(_run_field(capi, L.id, fsel, OPS["to_mont"], enc(A, mont=False), None, nb, n))
    assert raw == A


@pytest.mark.parametrize("cname", ["bn254", "bls12_377", "bls12_381"])
def test_fp2_ops(env, cname):
    capi, layout = env
    L = layout.Layout(cname)
    cx = ocurve.ctx(cname)
    F = cx.F2
    rnd = random.Random(77)
    p = L.p
    edge = [(0, 0), (1, 0), (0, 1), (p - 1, p - 1), (p - 1, 0), (0, p - 1)]
    A = [a for a in edge for _ in edge] + [(rnd.randrange(p), rnd.randrange(p)) for _ in range(150)]
    B = [b for _ in edge for b in edge] + [(rnd.randrange(p), rnd.randrange(p)) for _ in range(150)]
    n = len(A)
    ea = L.enc_fp([c for a in A for c in a])
    eb = L.enc_fp([c for b in B for c in b])

    def dec2(buf):
        v = L.dec_fp(buf)
        return [(v[2 * i], v[2 * i + 1]) for i in range(len(v) // 2)]

    nb = 2 * L.fp_bytes
    assert dec2(_run_field(capi, L.id, 2, OPS["add"], ea, eb, nb, n)) == [F.add(a, b) for a, b in zip(A, B)]
    assert dec2(_run_field(capi, L.id, 2, OPS["sub"], ea, eb, nb, n)) == [F.sub(a, b) for a, b in zip(A, B)]
    assert dec2(_run_field(capi, L.id, 2, OPS["mul"], ea, eb, nb, n)) == [F.mul(a, b) for a, b in zip(A, B)]
    assert dec2(_run_field(capi, L.id, 2, OPS["sqr"], ea, None, nb, n)) == [F.sqr(a) for a in A]
    assert dec2(_run_field(capi, L.id, 2, OPS["neg"], ea, None, nb, n)) == [F.neg(a) for a in A]
    want = [F.inv(a) if not F.is_zero(a) else (0, 0) for a in A]
    assert dec2(_run_field(capi, L.id, 2, OPS["inv"], ea, None, nb, n)) == want


@pytest.mark.parametrize("cname", CURVES)
@pytest.mark.parametrize("group", [1, 2])
def test_ec_ops(env, cname, group):
    from gpu_util import to_dev, dev_empty, ptr, stream, sync, rand_points, xyzz_of, affine_of_xyzz
    capi, layout = env
    L = layout.Layout(cname)
    cx = ocurve.ctx(cname)
    G = cx.group(group)
    rnd = random.Random(1234 + group)
    n = 40
    P = rand_points(cx, group, n, rnd)
    Q = rand_points(cx, group, n, rnd)
    # special cases: Q == P (doubling), Q == -P (infinity), acc infinity, Q infinity
    Q[0] = P[0]
    Q[1] = G.neg(P[1])
    P[2] = None
    Q[3] = None
    P[4] = None
    Q[4] = None
    PX = [xyzz_of(cx, group, p, rnd) for p in P]
    QX = [xyzz_of(cx, group, q, rnd) for q in Q]
    da = to_dev(L.enc_xyzz(PX, group))
    xb, ab = L.xyzz_bytes(group), L.affine_bytes(group)
    assert xb == capi.lib.b200_xyzz_bytes(L.id, group) and ab == capi.lib.b200_affine_bytes(L.id, group)

    def run(op, db, out_bytes):
        out = dev_empty(n * out_bytes)
        capi.check(capi.lib.b200_dbg_ec_op_dev(L.id, group, op, ptr(da), ptr(db), ptr(out), n, stream()))
        sync()
        return out.cpu().numpy()

    want_add = [G.add(p, q) for p, q in zip(P, Q)]
    got = L.dec_xyzz(run(0, to_dev(L.enc_affine(Q, group)), xb), group)
    assert [affine_of_xyzz(cx, group, g) for g in got] == want_add
    got = L.dec_xyzz(run(1, to_dev(L.enc_xyzz(QX, group)), xb), group)
    assert [affine_of_xyzz(cx, group, g) for g in got] == want_add
    got = L.dec_xyzz(run(2, None, xb), group)
    assert [affine_of_xyzz(cx, group, g) for g in got] == [G.add(p, p) for p in P]
    got = L.dec_affine(run(3, None, ab), group)
    assert got == P
    ks = [0, 1, 2, cx.r - 1] + [rnd.randrange(cx.r) for _ in range(n - 4)]
    got = L.dec_xyzz(run(4, to_dev(L.enc_fr(ks)), xb), group)
    assert [affine_of_xyzz(cx, group, g) for g in got] == [G.mul(p, k) for p, k in zip(P, ks)]
