#!/usr/bin/env python3
"""Ad-hoc GPU probe (not a test, not the bench): Montgomery-multiply issue-rate calibration and raw
MSM timings on random data.  Writes gpurun_out/probe.json."""
import ctypes as C
import json
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from davinci_node_b200 import capi, layout  # noqa: E402


def rand_elems(n, limbs64, bits, rng):
    a = rng.integers(0, 1 << 63, size=(n, limbs64), dtype=np.uint64) * 2 + rng.integers(0, 2, size=(n, limbs64), dtype=np.uint64)
    top_bits = bits - 1 - 64 * (limbs64 - 1)
    a[:, -1] &= np.uint64((1 << top_bits) - 1)
    return a


def timed(fn, reps=3):
    fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(reps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        fn()
        e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    return min(ts), ts


def main():
    capi.init()
    rng = np.random.default_rng(1)
    res = {"calib": {}, "msm": []}
    st = torch.cuda.current_stream().cuda_stream
    # ---- calibration
    for cid, name in [(1, "bn254"), (2, "bls12_377"), (4, "bw6_761")]:
        L = layout.Layout(cid)
        nthreads = 148 * 2048
        iters = 2000 if cid != 4 else 500
        buf = torch.from_numpy(rand_elems(nthreads, L.fp_l, L.p.bit_length(), rng).view(np.uint8)).cuda()
        ms, _ = timed(lambda: capi.check(capi.lib.b200_calib_mul_dev(cid, 0, buf.data_ptr(), nthreads, iters, st)))
        N = 2 * L.fp_l
        macs = nthreads * iters * (2 * N * N + N)
        res["calib"][name] = {"ms": ms, "mul_per_s": nthreads * iters / ms * 1e3, "wide_mac_per_s": macs / ms * 1e3}
        print("calib", name, res["calib"][name], flush=True)
    # ---- MSM timings
    cases = [(2, 1, 16), (2, 1, 18), (2, 1, 20), (2, 1, 22), (2, 2, 18), (2, 2, 20), (1, 1, 20), (1, 1, 22), (1, 2, 20),
             (4, 1, 18), (4, 1, 20), (4, 2, 18), (3, 1, 12)]
    if len(sys.argv) > 1:
        cases = [tuple(int(x) for x in a.split(",")) for a in sys.argv[1:]]
    for cid, group, lg in cases:
        L = layout.Layout(cid)
        n = 1 << lg
        w = L.coord_width(group)
        pts = torch.from_numpy(rand_elems(n * 2 * w, L.fp_l, L.p.bit_length(), rng).view(np.uint8)).cuda()
        sc = torch.from_numpy(rand_elems(n, L.fr_l, L.r.bit_length(), rng).view(np.uint8)).cuda()
        out = torch.zeros(L.xyzz_bytes(group), dtype=torch.uint8, device="cuda")
        plan = capi.msm_plan(cid, n)
        ms, all_ms = timed(lambda: capi.check(capi.lib.b200_msm_dev(cid, group, pts.data_ptr(), sc.data_ptr(), n,
                                                                     out.data_ptr(), 0, st)))
        rec = {"curve": L.name, "group": group, "log_n": lg, "ms": ms, "all_ms": all_ms, "plan": plan,
               "points_per_s": n / ms * 1e3}
        # table mode (precomputed window multiples resident in HBM)
        hb = C.c_uint64(0)
        t0 = time.time()
        capi.check(capi.lib.b200_bases_create_dev(cid, group, pts.data_ptr(), n, 0, C.byref(hb), st))
        rec["table_build_s"] = time.time() - t0
        tms, _ = timed(lambda: capi.check(capi.lib.b200_msm_bases_dev(hb.value, sc.data_ptr(), n, None, out.data_ptr(), st)))
        rec["table_ms"] = tms
        capi.check(capi.lib.b200_bases_release(hb.value))
        res["msm"].append(rec)
        print("msm", {k: rec[k] for k in ("curve", "group", "log_n", "ms", "table_ms", "table_build_s")}, flush=True)
        del pts, sc
        torch.cuda.empty_cache()
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    with open(os.path.join(ROOT, "gpurun_out", "probe.json"), "w") as fh:
        json.dump(res, fh, indent=1)


if __name__ == "__main__":
    main()
