#!/usr/bin/env python3
"""BASELINE.json configs[4]: MSM (G1/G2, table mode and windowed) and NTT microbench sweep per curve on
one B200, next to the CPU restatement (oracle/c) on the host cores for the sizes it finishes quickly.
Writes gpurun_out/sweep.json and a markdown table (gpurun_out/sweep.md).

  python tools/sweep.py [--max-log 26] [--cpu-max-log 20]
"""
import argparse
import ctypes as C
import json
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from davinci_node_b200 import capi, layout, synthetic  # noqa: E402
from davinci_node_b200.curve_consts import domain_constants  # noqa: E402


def p_mul(n32):
    return 2 * n32 * n32 + n32


def adds_star(n, bits):
    return min(n * (-(-(bits + 1) // c)) + 2 * (-(-(bits + 1) // c)) * (1 << (c - 1)) for c in range(4, 25))


def timed(fn, reps=3):
    fn()
    torch.cuda.synchronize()
    best = 1e30
    for _ in range(reps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        fn()
        e1.record()
        torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1))
    return best


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--max-log", type=int, default=26)
    ap.add_argument("--cpu-max-log", type=int, default=20)
    ap.add_argument("--curves", default="bn254,bls12_377,bw6_761")
    args = ap.parse_args()
    capi.init(1)
    lib = capi.lib
    st = torch.cuda.current_stream().cuda_stream
    rng = np.random.default_rng(7)
    peak = None
    rows = []
    cpu = None
    try:
        from oracle import cport
        cpu = cport.lib()
    except Exception as e:  # pragma: no cover
        print("cpu port unavailable:", e)
    for cname in args.curves.split(","):
        L = layout.Layout(cname)
        nfp, nfr = 2 * L.fp_l, 2 * L.fr_l
        bits = L.r.bit_length()
        # calibration: measured IMAD.WIDE rate on this field
        nthreads, iters = 148 * 2048, 600 if cname != "bw6_761" else 200
        cb = torch.from_numpy(synthetic.rand_canonical(rng, nthreads, L.fp_l, L.p.bit_length()).view(np.uint8).reshape(-1)).cuda()
        ms = timed(lambda: capi.check(lib.b200_calib_mul_dev(L.id, 0, cb.data_ptr(), nthreads, iters, st)), 2)
        peak = nthreads * iters * p_mul(nfp) / (ms / 1e3)
        for lg in range(16, args.max_log + 1, 2):
            n = 1 << lg
            sc = torch.from_numpy(synthetic.rand_canonical(rng, n, L.fr_l, bits).view(np.uint8).reshape(-1)).cuda()
            for grp in (1, 2):
                w = L.coord_width(grp)
                pb = 2 * w * L.fp_bytes
                if n * pb > 9e9:                             # keep points + sort workspaces inside one GPU
                    continue
                # table mode needs ~13 tables plus 1.5x that as transient build scratch
                with_tables = n * pb * 13 * 2.6 < 90e9
                pts = torch.from_numpy(synthetic.rand_canonical(rng, n * 2 * w, L.fp_l, L.p.bit_length()).view(np.uint8).reshape(-1)).cuda()
                out = torch.zeros(L.xyzz_bytes(grp), dtype=torch.uint8, device="cuda")
                wms = timed(lambda: capi.check(lib.b200_msm_dev(L.id, grp, pts.data_ptr(), sc.data_ptr(), n, out.data_ptr(), 0, st)), 2)
                tms = None
                if with_tables:
                    hb = C.c_uint64(0)
                    capi.check(lib.b200_bases_create_dev(L.id, grp, pts.data_ptr(), n, 0, C.byref(hb), st))
                    tms = timed(lambda: capi.check(lib.b200_msm_bases_dev(hb.value, sc.data_ptr(), n, None, out.data_ptr(), st)), 2)
                    capi.check(lib.b200_bases_release(hb.value))
                best = min(wms, tms) if tms else wms
                macs = adds_star(n, bits) * 10 * p_mul(nfp) * (3 if (grp == 2 and L.g2_deg == 2) else 1)
                rec = {"curve": cname, "op": "msm_g%d" % grp, "log_n": lg, "gpu_ms_table": tms, "gpu_ms_windowed": wms,
                       "imad_frac_measured_peak": macs / (best / 1e3) / peak, "points_per_s": n / (best / 1e3)}
                if cpu is not None and lg <= args.cpu_max_log:
                    hp, hs = pts.cpu().numpy(), sc.cpu().numpy()
                    ho = np.zeros(L.affine_bytes(grp), dtype=np.uint8)
                    t0 = time.perf_counter()
                    cpu.oc_msm(L.id, grp, hp.ctypes.data, hs.ctypes.data, n, None, ho.ctypes.data, 0)
                    rec["cpu_ms"] = (time.perf_counter() - t0) * 1e3
                    rec["cpu_threads"] = cpu.oc_num_threads()
                rows.append(rec)
                print(rec, flush=True)
                del pts
                torch.cuda.empty_cache()
            # NTT forward DIF + inverse coset DIF
            omega, g = domain_constants(L.id, lg)
            gw, gc = L.enc_fr([omega]), L.enc_fr([g])
            dom = C.c_uint64(0)
            capi.check(lib.b200_domain_create(L.id, n, gw.ctypes.data, gc.ctypes.data, C.byref(dom)))
            nms = timed(lambda: capi.check(lib.b200_ntt_dev(dom.value, sc.data_ptr(), 0, 0, 0, st)), 3)
            cms = timed(lambda: capi.check(lib.b200_ntt_dev(dom.value, sc.data_ptr(), 1, 0, 1, st)), 3)
            capi.check(lib.b200_domain_release(dom.value))
            rec = {"curve": cname, "op": "ntt", "log_n": lg, "gpu_ms_fwd": nms, "gpu_ms_inv_coset": cms,
                   "hbm_gbs": 2 * n * L.fr_bytes / (nms / 1e3) / 1e9,
                   "imad_frac_measured_peak": (n // 2) * lg * p_mul(nfr) / (nms / 1e3) / peak}
            if cpu is not None and lg <= args.cpu_max_log + 2:
                hs = sc.cpu().numpy().copy()
                t0 = time.perf_counter()
                cpu.oc_fft(L.id, hs.ctypes.data, lg, gw.ctypes.data, gc.ctypes.data, 0, 0, 0, 0)
                rec["cpu_ms"] = (time.perf_counter() - t0) * 1e3     # includes the port's twiddle generation
            rows.append(rec)
            print(rec, flush=True)
            del sc
            torch.cuda.empty_cache()
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    json.dump({"imad_wide_peak_last_curve": peak, "rows": rows}, open(os.path.join(ROOT, "gpurun_out", "sweep.json"), "w"), indent=1)
    with open(os.path.join(ROOT, "gpurun_out", "sweep.md"), "w") as fh:
        fh.write("| curve | op | log2 n | B200 ms (table) | B200 ms (windowed) | frac of measured IMAD.WIDE peak | CPU port ms (threads) |\n|---|---|---:|---:|---:|---:|---:|\n")
        for r in rows:
            if r["op"].startswith("msm"):
                fh.write("| %s | %s | %d | %s | %.2f | %.2f | %s |\n" % (r["curve"], r["op"], r["log_n"], ("%.2f" % r["gpu_ms_table"]) if r["gpu_ms_table"] else "- (tables > HBM budget)", r["gpu_ms_windowed"],
                         r["imad_frac_measured_peak"], ("%.0f (%d)" % (r["cpu_ms"], r["cpu_threads"])) if "cpu_ms" in r else "-"))
            else:
                fh.write("| %s | ntt fwd / inv-coset | %d | %.3f / %.3f | - | %.2f (%.0f GB/s) | %s |\n" % (r["curve"], r["log_n"], r["gpu_ms_fwd"], r["gpu_ms_inv_coset"],
                         r["imad_frac_measured_peak"], r["hbm_gbs"], ("%.0f" % r["cpu_ms"]) if "cpu_ms" in r else "-"))


if __name__ == "__main__":
    main()
