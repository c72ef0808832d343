"""Helpers shared by the -m gpu parity tests: torch is used only as the device-memory allocator."""
import random

import numpy as np

from oracle import curve as ocurve


def torch_mod():
    import torch
    return torch


def to_dev(arr):
    torch = torch_mod()
    return torch.from_numpy(np.ascontiguousarray(arr)).cuda()


def dev_empty(nbytes):
    torch = torch_mod()
    return torch.zeros(max(int(nbytes), 16), dtype=torch.uint8, device="cuda")


def ptr(t):
    return t.data_ptr() if t is not None else None


def stream():
    return torch_mod().cuda.current_stream().cuda_stream


def sync():
    torch_mod().cuda.synchronize()


def rand_points(cx, group, n, rnd):
    """n random subgroup points (multiples of the generator), affine."""
    G = cx.group(group)
    g = cx.gen(group)
    base = [G.mul(g, rnd.randrange(1, cx.r)) for _ in range(min(n, 24))]
    pts = list(base)
    while len(pts) < n:       # cheap extra points: sums of existing ones
        pts.append(G.add(pts[rnd.randrange(len(pts))], base[rnd.randrange(len(base))]))
    return pts[:n]


def xyzz_of(cx, group, pt, rnd):
    """A random XYZZ representative of an affine point (None -> ZZ = 0)."""
    F = cx.group(group).F
    if pt is None:
        return (F.from_int(rnd.randrange(cx.p)) if F.deg == 1 else (rnd.randrange(cx.p), 1), F.one, F.zero, F.zero)
    z = F.from_int(rnd.randrange(1, cx.p)) if F.deg == 1 else (rnd.randrange(1, cx.p), rnd.randrange(cx.p))
    zz = F.sqr(z)
    zzz = F.mul(zz, z)
    return (F.mul(pt[0], zz), F.mul(pt[1], zzz), zz, zzz)


def affine_of_xyzz(cx, group, q):
    F = cx.group(group).F
    X, Y, ZZ, ZZZ = q
    if F.is_zero(ZZ):
        return None
    return (F.mul(X, F.inv(ZZ)), F.mul(Y, F.inv(ZZZ)))
