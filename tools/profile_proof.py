#!/usr/bin/env python3
"""One proof of the bench workload between cudaProfilerStart/Stop, for a per-kernel launch list:

  ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv \
      --log-file gpurun_out/launches_proof.csv python tools/profile_proof.py [--logn 22]

Run without ncu it prints the single-proof latency (1 in flight) and the 2-in-flight throughput."""
import argparse
import ctypes as C
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--logn", type=int, default=22)
    ap.add_argument("--curve", default="bls12_377")
    ap.add_argument("--reps", type=int, default=3)
    ap.add_argument("--mix", default="witness")
    args = ap.parse_args()
    import torch
    from davinci_node_b200 import capi, synthetic
    capi.init(1)
    lib = capi.lib
    wl = synthetic.SyntheticWorkload(args.curve, args.logn, seed=0xD0A1)
    h = wl.register()
    sol = wl.solution(seed=1, mix=args.mix)
    L = wl.L
    r, s = 0x5EED5EED5EED5EED % L.r, (0x5EED << 64 | 0xABCDEF) % L.r
    pin, pout, out, keep = wl.prove_args(sol, r, s, on_device=True)
    for _ in range(2):
        capi.check(lib.b200_prove_dev(h, C.byref(pin), C.byref(pout), 0))
    torch.cuda.synchronize()
    torch.cuda.profiler.start()
    t0 = time.perf_counter()
    capi.check(lib.b200_prove_dev(h, C.byref(pin), C.byref(pout), 0))
    torch.cuda.synchronize()
    t1 = time.perf_counter()
    torch.cuda.profiler.stop()
    lat = []
    for _ in range(args.reps):
        t0 = time.perf_counter()
        capi.check(lib.b200_prove_dev(h, C.byref(pin), C.byref(pout), 0))
        torch.cuda.synchronize()
        lat.append((time.perf_counter() - t0) * 1e3)
    # kernel-family timers (CUDA events inside the library)
    capi.check(lib.b200_profile_enable(1))
    capi.check(lib.b200_prove_dev(h, C.byref(pin), C.byref(pout), 0))
    ms = (C.c_double * 8)()
    cnt = (C.c_uint64 * 8)()
    capi.check(lib.b200_profile_collect(ms, cnt))
    capi.check(lib.b200_profile_enable(0))
    print(json.dumps({"logn": args.logn, "curve": args.curve, "mix": args.mix, "profiled_ms": (t1 - t0) * 1e3,
                      "latency_ms": lat, "timers_ms": list(ms), "timers_n": list(cnt)}))


if __name__ == "__main__":
    main()
