#!/usr/bin/env python3
"""End-to-end proof rate from pageable host buffers vs the same buffers page-locked with b200_host_register
(what a Go shim sees with heap slices): python tools/pageable_probe.py [--logn 22] [--proofs 8] [--inflight 2]"""
import argparse
import ctypes as C
import os
import sys
import time
from concurrent.futures import ThreadPoolExecutor

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from davinci_node_b200 import capi, synthetic  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--logn", type=int, default=22)
    ap.add_argument("--proofs", type=int, default=8)
    ap.add_argument("--inflight", type=int, default=2)
    args = ap.parse_args()
    capi.init(1)
    wl = synthetic.SyntheticWorkload("bls12_377", args.logn, seed=0xD0A1)
    h = wl.register()
    L = wl.L
    r, s = 0x5EED5EED5EED5EED % L.r, (0x5EED << 64 | 0xABCDEF) % L.r
    sols = []
    for i in range(args.inflight):
        sol = wl.solution(seed=i, pinned=False)
        bufs = {k: np.array(sol[k].numpy(), copy=True) for k in ("W", "a", "b", "c")}      # plain pageable memory
        fake = dict(sol)
        fake.update({k: torch.from_numpy(v) for k, v in bufs.items()})
        sols.append((fake, bufs))
    argsets = [wl.prove_args(sols[j % args.inflight][0], r, s, on_device=False) for j in range(args.proofs)]
    pool = ThreadPoolExecutor(max_workers=args.inflight)

    def one(j):
        torch.cuda.set_device(0)
        pin, pout, out, keep = argsets[j]
        capi.check(capi.lib.b200_prove(h, C.byref(pin), C.byref(pout), 0))

    def rate():
        list(pool.map(one, range(args.inflight)))
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        list(pool.map(one, range(args.proofs)))
        torch.cuda.synchronize()
        return args.proofs / (time.perf_counter() - t0)

    pageable = rate()
    for _, bufs in sols:
        for v in bufs.values():
            capi.check(capi.lib.b200_host_register(v.ctypes.data, v.nbytes))
    registered = rate()
    print("e2e proofs/s from pageable host buffers: %.2f ; after b200_host_register: %.2f (n = 2^%d, %d in flight)"
          % (pageable, registered, args.logn, args.inflight))
    for _, bufs in sols:
        for v in bufs.values():
            capi.check(capi.lib.b200_host_unregister(v.ctypes.data))


if __name__ == "__main__":
    main()
