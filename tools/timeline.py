#!/usr/bin/env python3
"""Phase timeline of the proving schedule under real stream concurrency (no profiler attached):

  python tools/timeline.py [--logn 22] [--inflight 4] [--proofs 8] [--out gpurun_out/timeline.json]

Every instrumented phase (b200_profile_timeline) is bracketed by CUDA events on its own stream; the
tool prints, per phase family, the summed duration and the time covered by the union of its intervals,
plus the idle gaps of the compute-bound families (bucket accumulation + NTT passes)."""
import argparse
import ctypes as C
import json
import os
import sys
import time
from concurrent.futures import ThreadPoolExecutor

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

TAGS = {0: "acc_g1", 1: "acc_g2", 2: "ntt_pass", 3: "msm_total_g1", 4: "msm_total_g2", 5: "sort", 6: "sched",
        7: "ovf", 8: "bucket_reduce", 9: "sums", 10: "inputs", 11: "assemble", 12: "pre_reduce"}


def union(iv):
    iv = sorted(iv)
    tot, cur_s, cur_e = 0.0, None, None
    for s, e in iv:
        if cur_e is None or s > cur_e:
            if cur_e is not None:
                tot += cur_e - cur_s
            cur_s, cur_e = s, e
        else:
            cur_e = max(cur_e, e)
    if cur_e is not None:
        tot += cur_e - cur_s
    return tot


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--logn", type=int, default=22)
    ap.add_argument("--curve", default="bls12_377")
    ap.add_argument("--inflight", type=int, default=4)
    ap.add_argument("--proofs", type=int, default=8)
    ap.add_argument("--out", default=None)
    ap.add_argument("--dump", action="store_true")
    args = ap.parse_args()
    import torch
    from davinci_node_b200 import capi, synthetic
    capi.init(1)
    lib = capi.lib
    wl = synthetic.SyntheticWorkload(args.curve, args.logn, seed=0xD0A1)
    h = wl.register()
    L = wl.L
    sols = [wl.solution(seed=i) for i in range(2)]
    r, s = 0x5EED5EED5EED5EED % L.r, (0x5EED << 64 | 0xABCDEF) % L.r
    argsets = [wl.prove_args(sols[j % 2], r, s, on_device=True) for j in range(args.proofs)]

    def one(j):
        torch.cuda.set_device(0)
        pin, pout, out, keep = argsets[j]
        capi.check(lib.b200_prove_dev(h, C.byref(pin), C.byref(pout), 0))

    pool = ThreadPoolExecutor(max_workers=args.inflight)
    list(pool.map(one, range(args.proofs)))          # warm-up
    torch.cuda.synchronize()
    capi.check(lib.b200_profile_enable(1))
    t0 = time.perf_counter()
    list(pool.map(one, range(args.proofs)))
    torch.cuda.synchronize()
    wall = (time.perf_counter() - t0) * 1e3
    cap = 200000
    buf = (C.c_double * (4 * cap))()
    n = C.c_uint64(0)
    capi.check(lib.b200_profile_timeline(buf, cap, C.byref(n)))
    capi.check(lib.b200_profile_enable(0))
    recs = [(int(buf[4 * i]), int(buf[4 * i + 1]), buf[4 * i + 2], buf[4 * i + 3]) for i in range(n.value)]
    span = max(r[3] for r in recs) - min(r[2] for r in recs)
    fam = {}
    for tag, st, a, b in recs:
        fam.setdefault(tag, []).append((a, b))
    print("proofs=%d inflight=%d wall=%.1f ms  span=%.1f ms  (%.2f ms/proof)" % (args.proofs, args.inflight, wall, span,
                                                                              wall / args.proofs))
    print("%-16s %6s %10s %10s" % ("phase", "n", "sum ms", "union ms"))
    summary = {}
    for tag in sorted(fam):
        iv = fam[tag]
        sm, un = sum(b - a for a, b in iv), union(iv)
        summary[TAGS.get(tag, str(tag))] = {"n": len(iv), "sum_ms": sm, "union_ms": un}
        print("%-16s %6d %10.2f %10.2f" % (TAGS.get(tag, str(tag)), len(iv), sm, un))
    heavy = fam.get(0, []) + fam.get(1, []) + fam.get(2, [])
    print("compute-bound families (acc + ntt) cover %.1f of %.1f ms" % (union(heavy), span))
    if args.dump:
        for tag, st, a, b in sorted(recs, key=lambda r: r[2]):
            if tag in (3, 4):
                continue
            print("%9.3f %9.3f  s%-2d %s" % (a, b, st, TAGS.get(tag, str(tag))))
    if args.out:
        json.dump({"proofs": args.proofs, "inflight": args.inflight, "wall_ms": wall, "span_ms": span,
                   "summary": summary, "records": recs}, open(args.out, "w"))


if __name__ == "__main__":
    import signal
    signal.signal(signal.SIGPIPE, signal.SIG_DFL)      # `| head` must not end in a traceback
    main()
