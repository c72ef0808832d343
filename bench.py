#!/usr/bin/env python3
"""bench.py - voteverifier-shaped Groth16 proofs/s on B200 (BASELINE.json metric).

  python bench.py [--gpus N] [--steps K] [--warmup W] [--logn 22] [--impl reference]

A "step" is one proof (quotient H + 5 MSMs + commitment PoK + assembly) of one solved synthetic
ballot circuit per GPU: BLS12-377, n = m = 2^logn (default 2^22), witness-like scalar
mix, one BSB22 commitment over 2^(logn-4) wires.  N GPUs = N independent proofs per step, no
collective (weak scaling).  `value` times b200_prove_dev (inputs resident in HBM); `e2e` times
b200_prove (pinned host buffers in, proof bytes out, copies inside the timed region).  One JSON line
is printed by rank 0.  `--impl reference` times the CPU restatement of the reference path
(oracle/c, all host threads) on a bounded sample of the same workload.
"""
import argparse
import ctypes as C
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "voteverifier Groth16 proofs/s"
UNIT = "proofs/s"
CURVE = "bls12_377"


def p_mul(n32):
    return 2 * n32 * n32 + n32


def msm_adds_star(n, bits):
    """min over c in [4,24] of n*W(c) + 2*W(c)*2^(c-1), W(c) = ceil((b+1)/c)."""
    best = None
    for c in range(4, 25):
        w = -(-(bits + 1) // c)
        adds = n * w + 2 * w * (1 << (c - 1))
        best = adds if best is None else min(best, adds)
    return best


def workload_macs(logn, n_a, n_b, n_k, n_z, n_c):
    """Algorithmic 32x32->64 MACs of one BLS12-377 proof (SURVEY.md 8d definitions)."""
    pm_fp, pm_fr, bits = p_mul(12), p_mul(8), 253
    g1 = sum(msm_adds_star(x, bits) for x in (n_a, n_b, n_k, n_z, n_c)) * 10 * pm_fp
    g2 = msm_adds_star(n_b, bits) * 10 * pm_fp * 3
    ntt = 7 * ((1 << logn) // 2) * logn * pm_fr
    return g1, g2, ntt


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled during the timed region."""

    def __init__(self, index):
        self.index = index
        self.rows = []
        self.proc = None

    def start(self):
        q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
             "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
             "clocks_event_reasons.sw_power_cap")
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + q,
                                          "--format=csv,noheader,nounits", "-lms", "200"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except OSError:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([x.strip() for x in line.split(",")])

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            pass
        sm, mx, reasons = [], None, set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            try:
                sm.append(float(r[0]))
                mx = float(r[1])
            except (ValueError, IndexError):
                continue
            for k, nm in enumerate(names):
                if len(r) > 3 + k and r[3 + k].lower().startswith("active"):
                    reasons.add(nm)
        sm.sort()
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": mx, "reasons": sorted(reasons),
                "samples": len(sm)}


def cpu_reference_run(logn_sample, logn_full, steps, warmup, threads=0):
    """CPU restatement of the reference path (oracle/c: gnark's computeH + 5 MSMs schedule) on a
    2^logn_sample proof; proofs/s is extrapolated linearly to the 2^logn_full workload."""
    import numpy as np
    from davinci_node_b200 import synthetic
    from davinci_node_b200.curve_consts import domain_constants
    from davinci_node_b200.layout import Layout
    from oracle import cport
    lib = cport.lib()
    L = Layout(CURVE)
    n = 1 << logn_sample
    rng = np.random.default_rng(5)
    g1b, g2b, frb = L.affine_bytes(1), L.affine_bytes(2), L.fr_bytes
    # random field elements as coordinates: identical arithmetic cost, no setup needed on the CPU side
    pts = lambda cnt, w: synthetic.rand_canonical(rng, cnt * 2 * w, L.fp_l, L.p.bit_length())
    m, nb_public = n, 6
    n_c = 1 << max(1, logn_sample - 4)
    infA, infB = rng.random(m) < 0.30, rng.random(m) < 0.40
    SKIP = 0xFFFFFFFF
    mapA = np.full(m + 4, SKIP, dtype=np.uint32)
    mapB = np.full(m + 4, SKIP, dtype=np.uint32)
    mapA[:m][~infA] = np.arange(int((~infA).sum()), dtype=np.uint32)
    mapB[:m][~infB] = np.arange(int((~infB).sum()), dtype=np.uint32)
    nA, nB = int((~infA).sum()), int((~infB).sum())
    mapA[m], mapA[m + 2] = nA, nA + 1
    mapB[m + 1], mapB[m + 2] = nB, nB + 1
    npriv = m - nb_public
    mapK = np.full(npriv + 4, SKIP, dtype=np.uint32)
    keep = np.ones(npriv, dtype=bool)
    keep[:n_c + 1] = False
    nK = int(keep.sum())
    mapK[:npriv][keep] = np.arange(nK, dtype=np.uint32)
    mapK[npriv + 3] = nK
    A, B1, B2, K, Z = pts(nA + 2, 1), pts(nB + 2, 1), pts(nB + 2, 2), pts(nK + 1, 1), pts(n - 1, 1)
    W = synthetic.witness_like(rng, m + 4, L.fr_l, L.r.bit_length())
    omega, g = domain_constants(L.id, logn_sample)
    om, gg = L.enc_fr([omega]), L.enc_fr([g])
    outs = [np.zeros(g1b, dtype=np.uint8), np.zeros(g2b, dtype=np.uint8), np.zeros(g1b, dtype=np.uint8)]
    if threads <= 0:
        threads = lib.oc_num_threads()
    times = []
    for it in range(warmup + steps):
        a, b, c = (synthetic.rand_canonical(rng, n, L.fr_l, L.r.bit_length()) for _ in range(3))
        args = cport.ProveArgs(curve=L.id, logn=logn_sample, omega=cport.p(om), g=cport.p(gg), A_ext=cport.p(A),
                               B1_ext=cport.p(B1), B2_ext=cport.p(B2), K_ext=cport.p(K), Z=cport.p(Z),
                               mapA=cport.p(mapA), mapB=cport.p(mapB), mapK=cport.p(mapK), m=m, nb_public=nb_public,
                               nZ=n - 1, W_ext=cport.p(W), a=cport.p(a), b=cport.p(b), c=cport.p(c),
                               out_ar=cport.p(outs[0]), out_bs=cport.p(outs[1]), out_krs=cport.p(outs[2]),
                               threads=threads)
        t0 = time.perf_counter()
        assert lib.oc_prove(args) == 0
        dt = time.perf_counter() - t0
        if it >= warmup:
            times.append(dt)
    sample_s = sum(times) / len(times)
    scale = float(1 << (logn_full - logn_sample))
    return {"value": 1.0 / (sample_s * scale), "unit": UNIT, "cores": threads, "kind": "port",
            "sample": "one BLS12-377 proof at n=m=2^%d (1/%d of the workload), %.2f s on %d threads; value = "
                      "measured proofs/s / %d (linear extrapolation)" % (logn_sample, int(scale), sample_s, threads, int(scale)),
            "sample_seconds": sample_s}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--logn", type=int, default=22)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--batch", type=int, default=4, help="proofs per step per GPU")
    ap.add_argument("--inflight", type=int, default=4, help="proofs in flight per GPU (host threads, <= 4 slots)")
    ap.add_argument("--cpu-logn", type=int, default=17)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()

    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    config = {"workload": "voteverifier-shaped Groth16 proof, BLS12-377, n=m=2^%d, witness-like scalars, 1 BSB22 "
                          "commitment (2^%d wires), structured synthetic key" % (args.logn, max(1, args.logn - 4)),
              "proofs_per_step_per_gpu": args.batch,
              "parallelism": "independent proofs per GPU (%d in flight per GPU), no collective" % args.inflight,
              "l2": "inputs (%.0f MB/proof + 1.9 GB key) larger than L2" % ((4 * (1 << args.logn) * 32) / 1e6)}

    # ------------------------------------------------------------------ reference arm (CPU)
    if args.impl == "reference":
        if rank != 0:
            return
        steps = max(1, min(args.steps, 3))
        cb = cpu_reference_run(args.cpu_logn, args.logn, steps, min(args.warmup, 1))
        line = {"metric": METRIC, "value": cb["value"], "unit": UNIT, "n_gpus": args.gpus, "steps": steps,
                "warmup": min(args.warmup, 1), "ms_per_step": 1e3 / cb["value"], "higher_is_better": True,
                "scaling": "weak", "vs_baseline": None, "dtype": "u32x12 Montgomery (BLS12-377 fp) / u32x8 (fr)",
                "data": "synthetic", "config": config, "impl": "reference", "cpu_baseline": cb,
                "e2e": {"value": cb["value"], "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
                "gpu_launches": 0}
        print(json.dumps(line))
        return

    # ------------------------------------------------------------------ B200 arm
    # NCCL (used only for the barrier / max-over-ranks timing) prints its version banner on stdout at the
    # VERSION debug level, which would precede the one JSON line
    if os.environ.get("NCCL_DEBUG", "").upper() not in ("INFO", "TRACE"):
        os.environ["NCCL_DEBUG"] = "NONE"
    import numpy as np
    import torch
    import torch.distributed as dist
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device - this backend has no CPU fallback")
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    elif args.gpus > 1:
        raise SystemExit("bench.py: launch with torch.distributed.run for --gpus > 1 (one rank per GPU)")
    torch.cuda.set_device(local_rank)
    from davinci_node_b200 import capi, synthetic
    capi.init(1 << local_rank)
    lib = capi.lib

    wl = synthetic.SyntheticWorkload(CURVE, args.logn, seed=0xD0A1)
    h = wl.register()
    L = wl.L
    from concurrent.futures import ThreadPoolExecutor
    nsol = 2
    sols = [wl.solution(seed=1000 * rank + i) for i in range(nsol)]
    r, s = 0x5EED5EED5EED5EED % L.r, (0x5EED << 64 | 0xABCDEF) % L.r
    # one argument set (own output buffers) per proof of a step; witnesses cycle over `nsol` solutions
    dev_args = [wl.prove_args(sols[j % nsol], r, s, on_device=True) for j in range(args.batch)]
    host_args = [wl.prove_args(sols[j % nsol], r, s, on_device=False) for j in range(args.batch)]
    pool = ThreadPoolExecutor(max_workers=max(1, args.inflight))

    def one_dev(j):
        torch.cuda.set_device(local_rank)
        pin, pout, out, keep = dev_args[j]
        capi.check(lib.b200_prove_dev(h, C.byref(pin), C.byref(pout), local_rank))

    def one_host(j):
        torch.cuda.set_device(local_rank)
        pin, pout, out, keep = host_args[j]
        capi.check(lib.b200_prove(h, C.byref(pin), C.byref(pout), local_rank))

    def step_dev(i):
        list(pool.map(one_dev, range(args.batch)))

    def step_host(i):
        list(pool.map(one_host, range(args.batch)))

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(fn, steps):
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for i in range(steps):
            fn(i)
        e1.record()
        barrier()
        ms = e0.elapsed_time(e1)
        if world > 1:
            t = torch.tensor([ms], device="cuda", dtype=torch.float64)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms = float(t.item())
        return ms

    for i in range(max(args.warmup, 1)):
        step_dev(i)
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    launches0 = lib.b200_launch_count()
    ms_total = timed(step_dev, args.steps)
    launches = lib.b200_launch_count() - launches0
    step_host(0)
    ms_e2e = timed(step_host, args.steps)
    clocks = sampler.stop() if rank == 0 else None

    value = world * args.batch * args.steps / (ms_total / 1e3)
    e2e_value = world * args.batch * args.steps / (ms_e2e / 1e3)

    if rank != 0:
        if world > 1:
            dist.barrier()
            dist.destroy_process_group()
        return

    # ------------------------------------------------------------------ rank 0: roofline detail (serialised, untimed)
    st = torch.cuda.current_stream().cuda_stream
    # measured IMAD.WIDE issue rate (calibration kernel, bls12-377 fp)
    nthreads, iters = 148 * 2048, 1000
    cbuf = torch.from_numpy(synthetic.rand_canonical(np.random.default_rng(1), nthreads, L.fp_l, L.p.bit_length()).view(np.uint8).reshape(-1)).cuda()
    capi.check(lib.b200_calib_mul_dev(L.id, 0, cbuf.data_ptr(), nthreads, iters, st))
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    capi.check(lib.b200_calib_mul_dev(L.id, 0, cbuf.data_ptr(), nthreads, iters, st))
    e1.record()
    torch.cuda.synchronize()
    peak_meas = nthreads * iters * p_mul(12) / (e0.elapsed_time(e1) / 1e3)
    sm_max = (clocks or {}).get("sm_max_mhz") or 1965.0
    peak_nominal = 148 * 64 * sm_max * 1e6

    # dominant kernel: k_msm_accumulate<G1> of the quotient (Z) MSM - dense 253-bit scalars over n-1 points, one
    # third of a proof's GPU time.  Run alone, in table mode as the key is held, with the CUDA-event kernel timers
    # on; the G2 MSM (Bs) is measured the same way as `roofline_g2`.
    sol = sols[0]
    nB = len(wl.kB)

    def msm_alone(points_np, group, npts):
        pts = torch.from_numpy(points_np).cuda()
        outx = torch.zeros(L.xyzz_bytes(group), dtype=torch.uint8, device="cuda")
        hb = C.c_uint64(0)
        capi.check(lib.b200_bases_create_dev(L.id, group, pts.data_ptr(), npts, 0, C.byref(hb), st))
        run = lambda: capi.check(lib.b200_msm_bases_dev(hb.value, sol["a_dev"].data_ptr(), npts, None, outx.data_ptr(), st))
        run()
        torch.cuda.synchronize()
        capi.check(lib.b200_profile_enable(1))
        for _ in range(3):
            run()
        ms = (C.c_double * 5)()
        cnt = (C.c_uint64 * 5)()
        capi.check(lib.b200_profile_collect(ms, cnt))
        capi.check(lib.b200_profile_enable(0))
        capi.check(lib.b200_bases_release(hb.value))
        tot, acc = (3, 0) if group == 1 else (4, 1)
        return ms[tot] / max(cnt[tot], 1), ms[acc] / max(cnt[acc], 1)

    # uniform full-width scalars (the quotient's a-vector): the SURVEY 8d work formula is exact for them,
    # whereas the witness-like wire vector skips ~60% of the points and would flatter the fraction
    nZ = min(wl.nc, wl.n - 1)
    g1_total_ms, g1_acc_ms = msm_alone(wl.pk.g1_Z, 1, nZ)
    g2_total_ms, g2_acc_ms = msm_alone(wl.pk.g2_B, 2, nB)
    ms = (C.c_double * 5)()
    cnt = (C.c_uint64 * 5)()
    reps = 3
    # NTT passes alone (quotient on resident buffers)
    dom = C.c_uint64(0)
    from davinci_node_b200.curve_consts import domain_constants
    omega, cg = domain_constants(L.id, args.logn)
    gw, gc = L.enc_fr([omega]), L.enc_fr([cg])
    capi.check(lib.b200_domain_create(L.id, wl.n, gw.ctypes.data, gc.ctypes.data, C.byref(dom)))
    buf = torch.zeros(wl.n * L.fr_bytes, dtype=torch.uint8, device="cuda")
    buf[:sol["a_dev"].numel()] = sol["a_dev"]
    capi.check(lib.b200_ntt_dev(dom.value, buf.data_ptr(), 0, 0, 0, st))
    torch.cuda.synchronize()
    capi.check(lib.b200_profile_enable(1))
    for _ in range(reps):
        capi.check(lib.b200_ntt_dev(dom.value, buf.data_ptr(), 0, 0, 0, st))
    capi.check(lib.b200_profile_collect(ms, cnt))
    ntt_ms = ms[2] / reps
    ntt_passes = cnt[2] // reps
    capi.check(lib.b200_profile_enable(0))
    capi.check(lib.b200_domain_release(dom.value))

    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    hbm_peak = peaks.get("hbm_gbs", 6650.0)
    g1_macs, g2_macs, ntt_macs = workload_macs(args.logn, len(wl.kA) + 2, nB + 2, len(wl.kK) + 1, wl.n - 1, wl.n_c)
    g1_alg = msm_adds_star(nZ, 253) * 10 * p_mul(12)
    g2_alg = msm_adds_star(nB, 253) * 10 * p_mul(12) * 3
    peak_note = "measured in this run (b200_calib_mul_dev); nominal 148*64*f = %.2f" % (peak_nominal / 1e12)
    roofline = {"bound": "imad", "kernel": "k_msm_accumulate<G1> inside the quotient (Z) MSM: %d points, uniform 253-bit "
                                            "scalars, table mode c=20 (the kernel is %.0f%% of the MSM)" % (nZ, 100 * g1_acc_ms / g1_total_ms),
                "achieved": g1_alg / (g1_total_ms / 1e3) / 1e12, "peak": peak_meas / 1e12,
                "unit": "T wide-MAC/s (32x32->64 IMAD.WIDE)", "frac": g1_alg / (g1_total_ms / 1e3) / peak_meas,
                "peak_source": peak_note, "frac_of_nominal": g1_alg / (g1_total_ms / 1e3) / peak_nominal,
                # dram__bytes_read+write of the kernel from the committed ncu capture at 2^22 points
                # (profiles/r1_ncu_msm_accumulate_g1.md: 10.69 GB + 0.17 GB), scaled to this launch's point count
                "traffic": (10.69e9 + 0.17e9) * nZ / float(1 << 22),
                "traffic_note": "algorithmic point bytes = adds* x 96 B = %.2f GB; gathered 96-byte points straddle "
                                "64-byte DRAM bursts" % (msm_adds_star(nZ, 253) * 96 / 1e9),
                "launch_ms": g1_total_ms, "kernel_ms": g1_acc_ms, "algorithmic_macs_per_launch": g1_alg,
                "ncu_pipe_fmaheavy_pct": 88.6}
    roofline_g2 = {"bound": "imad", "kernel": "G2 MSM (Bs), %d points, uniform 253-bit scalars (k_msm_accumulate<Fp2> = %.0f%% of it)" % (nB, 100 * g2_acc_ms / g2_total_ms),
                   "achieved": g2_alg / (g2_total_ms / 1e3) / 1e12, "peak": peak_meas / 1e12,
                   "unit": "T wide-MAC/s (32x32->64 IMAD.WIDE)", "frac": g2_alg / (g2_total_ms / 1e3) / peak_meas,
                   "peak_source": peak_note, "frac_of_nominal": g2_alg / (g2_total_ms / 1e3) / peak_nominal,
                   # profiles/r1_ncu_msm_accumulate_g2.md: 9.80 GB + 16.61 GB at 2^21 points
                   "traffic": (9.80e9 + 16.61e9) * nB / float(1 << 21),
                   "traffic_note": "algorithmic point bytes = adds* x 192 B; the excess is local-memory traffic of the "
                                   "out-of-line Fp2 multiply",
                   "launch_ms": g2_total_ms, "kernel_ms": g2_acc_ms, "algorithmic_macs_per_launch": g2_alg}
    ntt_bytes = 2 * wl.n * L.fr_bytes
    roofline_ntt = {"bound": "hbm", "kernel": "k_ntt_pass (one 2^%d transform = %d passes)" % (args.logn, ntt_passes),
                    "achieved": ntt_bytes / (ntt_ms / 1e3) / 1e9, "peak": hbm_peak, "unit": "GB/s",
                    "frac": ntt_bytes / (ntt_ms / 1e3) / 1e9 / hbm_peak,
                    "peak_source": "MEASURED_PEAKS.json" if peaks else "fallback", "traffic": None,
                    "imad_frac": ((wl.n // 2) * args.logn * p_mul(8)) / (ntt_ms / 1e3) / peak_meas, "launch_ms": ntt_ms}
    step_macs = (g1_macs + g2_macs + ntt_macs) * args.batch
    cb = None
    if not args.no_cpu_baseline and world == 1:
        cb = cpu_reference_run(args.cpu_logn, args.logn, 1, 0)
    line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms_total / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "u32x12 Montgomery (BLS12-377 fp) / u32x8 (fr)", "data": "synthetic", "config": config,
            "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": wl.h2d_bytes() * args.batch, "d2h_bytes_per_step": wl.d2h_bytes() * args.batch,
                    "ms_per_step": ms_e2e / args.steps},
            "gpu_launches": int(launches), "clocks": clocks, "roofline": roofline, "roofline_g2": roofline_g2,
            "roofline_ntt": roofline_ntt,
            "step_imad_frac_dense_formula": step_macs / (ms_total / args.steps / 1e3) / peak_meas,
            "step_imad_note": "SURVEY 8d work formula assumes dense scalars; the witness-like wire vector skips ~60% of "
                              "the A/B/K points, so this can exceed 1 - the kernel-level `roofline` uses dense scalars",
            "cpu_baseline": cb}
    print(json.dumps(line))
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
