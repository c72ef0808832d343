import os, sys, random
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
from davinci_node_b200 import capi, layout
from oracle import curve as ocurve
from gpu_util import to_dev, dev_empty, ptr, stream, sync, rand_points, xyzz_of, affine_of_xyzz
capi.init()
cname, group = "bn254", 1
L = layout.Layout(cname); cx = ocurve.ctx(cname); G = cx.group(group)
rnd = random.Random(1235); n = 40
P = rand_points(cx, group, n, rnd)
PX = [xyzz_of(cx, group, p, rnd) for p in P]
da = to_dev(L.enc_xyzz(PX, group)); xb = L.xyzz_bytes(group)
def run(ks):
    out = dev_empty(n * xb)
    capi.check(capi.lib.b200_dbg_ec_op_dev(L.id, group, 4, ptr(da), ptr(to_dev(L.enc_fr(ks))), ptr(out), n, stream())); sync()
    return L.dec_xyzz(out.cpu().numpy(), group)
for name, ks in (("all1", [1]*n), ("all3", [3]*n), ("mixed01", [i % 2 for i in range(n)]), ("1_then_rand", [1] + [rnd.randrange(cx.r) for _ in range(n-1)]),
                 ("small", list(range(n)))):
    got = run(ks)
    bad = [i for i in range(n) if affine_of_xyzz(cx, group, got[i]) != G.mul(P[i], ks[i])]
    print(name, "bad:", bad[:12])
    for i in bad[:2]:
        g, p = got[i], PX[i]
        print("  i", i, "k", ks[i], "coords equal to input:", [g[c] == p[c] for c in range(4)])
print("---- details")
got = run([1]*n)
for i in (0, 1, 33):
    print(i, "gotX", hex(got[i][0])[:20], "inX", hex(PX[i][0])[:20], "inY", hex(PX[i][1])[:20], "inZZ", hex(PX[i][2])[:20], "inZZZ", hex(PX[i][3])[:20])
    others = {("X%d" % j): PX[j][0] for j in range(n)}
    print("   equals some input coord:", [k for k, v in others.items() if v == got[i][0]][:3], "zero" if got[i][0] == 0 else "")
# op 1 (add) with acc = inf for all
INF = [xyzz_of(cx, group, None, rnd) for _ in range(n)]
dinf = to_dev(L.enc_xyzz(INF, group))
out = dev_empty(n * xb)
capi.check(capi.lib.b200_dbg_ec_op_dev(L.id, group, 1, ptr(dinf), ptr(da), ptr(out), n, stream())); sync()
g2 = L.dec_xyzz(out.cpu().numpy(), group)
print("add(inf, P) all-copy: bad", [i for i in range(n) if g2[i] != PX[i]][:10])
if g2[0] != PX[0]:
    print("  coords equal:", [g2[0][c] == PX[0][c] for c in range(4)], hex(g2[0][0])[:20])
