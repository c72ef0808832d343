#!/usr/bin/env python3
"""bench.py - Groth16 proofs/s on B200 (BASELINE.json metric) and the CPU reference beside it.

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference]
                  [--config voteverifier|aggregator|statetransition|blob] [--mode independent|range-split]
                  [--logn L] [--mix witness|uniform]

Default (BASELINE.json configs[1], the config the metric is quoted on): voteverifier-shaped proofs, BLS12-377,
n = m = 2^22, witness-like scalar mix, one BSB22 commitment over 2^18 wires.  A "step" is `--batch` proofs (quotient H +
5 MSMs + commitment PoK + assembly) per GPU; N GPUs = N independent proof streams, no collective (weak scaling).
`value` times b200_prove_dev (inputs resident in HBM); `e2e` times b200_prove (pinned host buffers in, proof bytes out,
copies inside the timed region).  One JSON line is printed by rank 0.

  --config aggregator       BW6-761, n = 2^22 (configs[2]); with --gpus N > 1 ONE proof is split by point range over the
                            N GPUs (`--mode range-split`: partial sums per GPU, NCCL all-gather of a few hundred bytes,
                            assembly; strong scaling)
  --config statetransition  BN254, n = 2^24, followed by the blob KZG work of state/blobs.go:29-117 (configs[3])
  --config blob             EIP-4844 blob commitments alone
  --impl reference          the reference's CPU path (oracle/c restatement of gnark's prover, every host thread) on the
                            SAME workload: the first warm-up step is one FULL proof (timed, reported), every other step
                            is a bounded sample = one component of a full-size proof (see oracle/c/oracle.cpp).
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

import bench_workload as BW  # noqa: E402  (numpy only: shared by both arms)

DTYPES = {"bls12_377": "u32x12 Montgomery (BLS12-377 fp) / u32x8 (fr)", "bn254": "u32x8 Montgomery (BN254 fp, fr)",
          "bw6_761": "u32x24 Montgomery (BW6-761 fp) / u32x12 (fr)", "bls12_381": "u32x12 Montgomery (BLS12-381 fp) / u32x8 (fr)"}


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


def workload_text(cfg_name, shape, mix):
    cfg = BW.CONFIGS[cfg_name]
    curve = getattr(shape, "curve", None) or shape.L.name
    return ("%s, %s, n=m=2^%d, %s scalars, 1 BSB22 commitment (2^%d wires), structured synthetic key"
            % (cfg["title"], curve.upper().replace("_", "-"), shape.logn,
               "witness-like (40%% 0 / 20%% 1 / 25%% <2^64 / 15%% full)" if mix == "witness" else "uniform-random",
               max(1, shape.n_c.bit_length() - 1)))


# ====================================================================================== CPU reference arm
class CpuProver:
    """The reference's CPU path on one workload shape: oracle/c's restatement of gnark's computeH + MultiExp schedule
    (kind "port": gnark itself cannot be built here - no Go toolchain), all host threads, valid curve points.
    Imports only the oracle and numpy."""

    def __init__(self, shape, mix="witness", threads=0):
        import numpy as np
        from oracle import cport
        from oracle import curve as OC
        self.np, self.cport, self.lib, self.shape = np, cport, cport.lib(), shape
        self.threads = threads if threads > 0 else cport.host_threads()
        cx = OC.ctx(shape.curve)
        sh, T, lib = shape, self.threads, self.lib
        fp_l = sh.fp_l

        def gen(group, count, k0):
            w = 1 if (group == 1 or sh.g2_deg == 1) else 2
            g = cx.g1 if group == 1 else cx.g2
            flat = [g[0], g[1]] if w == 1 else [g[0][0], g[0][1], g[1][0], g[1][1]]
            base = BW.enc_mont(flat, sh.p, fp_l)
            out = np.zeros((count, 2 * w * fp_l), dtype=np.uint64)
            assert lib.oc_gen_points(sh.cid, group, cport.p(base), k0, count, cport.p(out), T) == 0
            return out

        t0 = time.perf_counter()
        self.A, self.B1, self.B2 = gen(1, sh.nA + 2, 1000003), gen(1, sh.nB + 2, 2000003), gen(2, sh.nB + 2, 2000003)
        self.K, self.Z, self.sigma = gen(1, sh.nK + 1, 3000017), gen(1, sh.nZ, 5000011), gen(1, sh.n_c, 7000003)
        self.keygen_s = time.perf_counter() - t0
        self.mapA, self.mapB, self.mapK = sh.maps()
        rng = np.random.default_rng(5)
        bits = sh.fr_bits()

        def mont(canon):
            out = np.empty_like(canon)
            assert lib.oc_fr_to_mont(sh.cid, cport.p(canon), cport.p(out), canon.shape[0], T) == 0
            return out

        Wc = BW.witness_like(rng, sh.m + 4, sh.fr_l, bits, mix)
        Wc[0] = 0
        Wc[0, 0] = 1
        self.W = mont(Wc)
        self.cvals = np.ascontiguousarray(self.W[sh.committed[0]:sh.committed[0] + sh.n_c])
        self.a0 = mont(BW.rand_canonical(rng, sh.n, sh.fr_l, bits))
        self.b0 = mont(BW.rand_canonical(rng, sh.n, sh.fr_l, bits))
        self.a0[sh.nc:] = 0
        self.b0[sh.nc:] = 0
        self.c0 = np.empty_like(self.a0)
        assert lib.oc_fr_mul(sh.cid, cport.p(self.a0), cport.p(self.b0), cport.p(self.c0), sh.n) == 0
        omega, g = BW.domain_constants(sh.curve, sh.logn)
        self.om, self.gg = BW.enc_mont([omega], sh.r, sh.fr_l), BW.enc_mont([g], sh.r, sh.fr_l)
        w1 = 2 * fp_l
        w2 = w1 if sh.g2_deg == 1 else 2 * w1
        self.outs = [np.zeros(w1, dtype=np.uint64), np.zeros(w2, dtype=np.uint64), np.zeros(w1, dtype=np.uint64),
                     np.zeros(w1, dtype=np.uint64)]
        self.a, self.b, self.c = self.a0.copy(), self.b0.copy(), self.c0.copy()
        self.comp_buf = np.zeros(7, dtype=np.float64)
        self.comp = None
        # key / domain set-up (gnark: fft.NewDomain at key load): builds the cached twiddle tables, untimed
        t0 = time.perf_counter()
        self.run(0, 10)
        self.domain_s = time.perf_counter() - t0

    COMPONENTS = ("H", "A", "B1", "B2", "K", "Z", "PoK+assembly")
    # component index (into comp_seconds) and share of it that sample part p runs
    PARTS = [((0, 1.0), (6, 1.0)), ((1, 1.0),), ((2, 1.0),), ((3, 0.5),), ((3, 0.5),), ((4, 1.0),),
             ((5, 0.25),), ((5, 0.25),), ((5, 0.25),), ((5, 0.25),)]

    def run(self, part=0, nparts=1):
        """One full proof (nparts == 1, fresh a / b / c; fills self.comp) or sample component `part` of 10; returns
        seconds."""
        cp, sh = self.cport, self.shape
        if nparts == 1:
            self.np.copyto(self.a, self.a0)
            self.np.copyto(self.b, self.b0)
            self.np.copyto(self.c, self.c0)
        args = cp.ProveArgs(curve=sh.cid, logn=sh.logn, omega=cp.p(self.om), g=cp.p(self.gg), A_ext=cp.p(self.A),
                            B1_ext=cp.p(self.B1), B2_ext=cp.p(self.B2), K_ext=cp.p(self.K), Z=cp.p(self.Z),
                            mapA=cp.p(self.mapA), mapB=cp.p(self.mapB), mapK=cp.p(self.mapK), m=sh.m,
                            nb_public=sh.nb_public, nZ=sh.nZ, W_ext=cp.p(self.W), a=cp.p(self.a), b=cp.p(self.b),
                            c=cp.p(self.c), out_ar=cp.p(self.outs[0]), out_bs=cp.p(self.outs[1]),
                            out_krs=cp.p(self.outs[2]), threads=self.threads, sigma=cp.p(self.sigma),
                            cvals=cp.p(self.cvals), n_commit=sh.n_c, out_pok=cp.p(self.outs[3]), part=part, nparts=nparts,
                            comp_seconds=cp.p(self.comp_buf))
        t0 = time.perf_counter()
        assert self.lib.oc_prove(args) == 0
        dt = time.perf_counter() - t0
        if nparts == 1:
            self.comp = [float(x) for x in self.comp_buf]
        return dt


def cpu_blob_commit_baseline(reps=20):
    """CPU leg of the blob config: the 4096-point BLS12-381 G1 MSM a commitment is (types/blobs.go:90 ->
    gokzg4844 BlobToKZGCommitment), timed through oracle/c's Pippenger on every host thread, valid subgroup points,
    uniform 255-bit scalars.  kind "port" (go-eth-kzg cannot be built here)."""
    import numpy as np
    from oracle import cport
    from oracle import curve as OC
    lib, T = cport.lib(), cport.host_threads()
    p_ = 0x1a0111ea397fe69a4b1ba7b6434bacd764774b84f38512bf6730d2a0f6b0f6241eabfffeb153ffffb9feffffffffaaab
    r_ = 0x73eda753299d7d483339d80809a1d80553bda402fffe5bfeffffffff00000001
    fp_l, fr_l = 6, 4
    cx = OC.ctx("bls12_381")
    base = BW.enc_mont([cx.g1[0], cx.g1[1]], p_, fp_l)
    pts = np.zeros((4096, 2 * fp_l), dtype=np.uint64)
    assert lib.oc_gen_points(3, 1, cport.p(base), 7000003, 4096, cport.p(pts), T) == 0
    rng = np.random.default_rng(9)
    canon = BW.rand_canonical(rng, 4096, fr_l, r_.bit_length())
    sc = np.empty_like(canon)
    assert lib.oc_fr_to_mont(3, cport.p(canon), cport.p(sc), 4096, T) == 0
    out = np.zeros(2 * fp_l, dtype=np.uint64)
    assert lib.oc_msm(3, 1, cport.p(pts), cport.p(sc), 4096, None, cport.p(out), T) == 0
    t0 = time.perf_counter()
    for _ in range(reps):
        assert lib.oc_msm(3, 1, cport.p(pts), cport.p(sc), 4096, None, cport.p(out), T) == 0
    dt = (time.perf_counter() - t0) / reps
    return {"value": 1.0 / dt, "unit": "commitments/s", "cores": T, "kind": "port",
            "sample": "%d x the 4096-point BLS12-381 G1 MSM of one commitment (%.2f ms each) on %d threads; the blob -> "
                      "scalar decoding (4096 reductions) is not included" % (reps, dt * 1e3, T)}


def cpu_reference_run(cfg_name, logn, mix, steps, warmup, nparts=10, full_proof=True):
    """Times the CPU arm.  Warm-up step 0 is ONE FULL proof of the workload (when full_proof); every other step is a
    bounded sample: component i mod 10 of a full-size proof (oracle/c/oracle.cpp prove_impl), which runs exactly the code
    the full proof runs for that component.  A step counts as the share of a proof its component took inside the full
    proof (1/10 each when no full proof was run).  Returns (cpu_baseline dict, seconds of the timed steps)."""
    cfg = BW.CONFIGS[cfg_name]
    shape = BW.Shape(cfg["curve"], logn, cfg["seed"], nb_public=cfg["nb_public"])
    cp = CpuProver(shape, mix)
    nparts = 10
    full_s = None
    done_warm = 0
    if full_proof and warmup >= 1:
        full_s = cp.run()
        done_warm = 1
    if cp.comp:
        tot = sum(cp.comp)
        weights = [sum(cp.comp[ci] * share for ci, share in cp.PARTS[p]) / tot for p in range(nparts)]
    else:
        weights = [1.0 / nparts] * nparts
    for i in range(done_warm, warmup):
        cp.run(i % nparts, nparts)
    t_steps, proofs_equiv = [], 0.0
    for i in range(steps):
        t_steps.append(cp.run(i % nparts, nparts))
        proofs_equiv += weights[i % nparts]
    total = sum(t_steps)
    unit = BW.METRICS[cfg_name][1]
    value = proofs_equiv / total if steps else (1.0 / full_s if full_s else None)
    sample = ("each step = component i mod 10 of ONE %s proof at full size n=m=2^%d (H + PoK + assembly | A | B1 | G2 halves | K | "
              "Z quarters by window range), counted as that component's share of the full proof's time: %.2f s per step on %d "
              "threads" % (shape.curve, logn, total / max(steps, 1), cp.threads))
    if full_s is not None:
        sample += "; warm-up step 0 = one FULL proof: %.1f s (= %.4f %s)" % (full_s, 1.0 / full_s, unit)
    cb = {"value": value, "unit": unit, "cores": cp.threads, "kind": "port", "sample": sample,
          "full_proof_seconds": full_s, "sample_step_seconds": total / max(steps, 1), "proofs_equivalent_timed": proofs_equiv,
          "component_seconds_full_proof": dict(zip(cp.COMPONENTS, cp.comp)) if cp.comp else None,
          "keygen_seconds": cp.keygen_s,
          "note": "oracle/c: C++ restatement of gnark's prover (64-bit CIOS, XYZZ buckets, c=16 Pippenger, OpenMP); "
                  "gnark's assembly field arithmetic is faster than this port"}
    return cb, total


def reference_arm(args, cfg_name, logn):
    """`--impl reference`: rank 0 alone; same metric / unit / config as the B200 arm."""
    metric, unit = BW.METRICS[cfg_name]
    cfg = BW.CONFIGS[cfg_name]
    shape = BW.Shape(cfg["curve"], logn, cfg["seed"], nb_public=cfg["nb_public"])
    cb, total = cpu_reference_run(cfg_name, logn, args.mix, args.steps, args.warmup, args.cpu_parts)
    config = {"workload": workload_text(cfg_name, shape, args.mix),
              "step": "one component of a proof (bounded sample, see cpu_baseline.sample)"}
    line = {"metric": metric, "value": cb["value"], "unit": unit, "n_gpus": args.gpus, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": 1e3 * total / max(args.steps, 1), "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": DTYPES[shape.curve], "data": "synthetic",
            "config": config, "impl": "reference", "cpu_baseline": cb,
            "e2e": {"value": cb["value"], "unit": unit, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line))


# ====================================================================================== B200 arm helpers
def load_profile_facts():
    """ncu-derived per-launch facts committed under profiles/ (written by tools/summarize_ncu.py --facts); the bench
    cites them with their source instead of carrying literals."""
    path = os.path.join(ROOT, "profiles", "r2_ncu_facts.json")
    try:
        return json.load(open(path)), "profiles/r2_ncu_facts.json"
    except Exception:
        return {}, None


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--logn", type=int, default=0, help="domain size override (default: the config's)")
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--config", default="voteverifier", choices=["voteverifier", "aggregator", "statetransition", "blob"])
    ap.add_argument("--mode", default="auto", choices=["auto", "independent", "range-split"])
    ap.add_argument("--mix", default="witness", choices=["witness", "uniform"])
    ap.add_argument("--batch", type=int, default=0, help="proofs per step per GPU (default 4; 1 for the big configs)")
    ap.add_argument("--inflight", type=int, default=0, help="proofs in flight per GPU (host threads)")
    ap.add_argument("--cpu-parts", type=int, default=10, help=argparse.SUPPRESS)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-roofline", action="store_true")
    ap.add_argument("--no-uniform", action="store_true", help="skip the uniform-random variant of the headline line")
    ap.add_argument("--no-shard-quotient", action="store_true", help="range-split: every GPU computes the whole quotient")
    ap.add_argument("--dump-timeline", default=None, help="range-split: write rank 0's phase timeline of one proof (JSON)")
    args = ap.parse_args()

    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    cfg_name = args.config
    if cfg_name == "blob":
        if args.impl == "reference":
            if rank == 0:
                print(json.dumps({"impl": "reference", "unavailable": "the blob-only config has no CPU restatement arm "
                                  "(the KZG oracle is big-int Python); use --config statetransition"}))
            return
        return blob_arm(args, rank, local_rank, world)
    logn = args.logn or BW.CONFIGS[cfg_name]["logn"]

    if args.impl == "reference":
        if rank != 0:
            return
        return reference_arm(args, cfg_name, logn)

    mode = args.mode
    if mode == "auto":
        mode = "range-split" if (cfg_name in ("aggregator", "statetransition") and world > 1) else "independent"
    if mode == "range-split":
        return range_split_arm(args, cfg_name, logn, rank, local_rank, world)
    return independent_arm(args, cfg_name, logn, rank, local_rank, world)


def _dist_setup(local_rank, world, gpus):
    # NCCL (barrier / max-over-ranks timing / the range-split all-gather) prints its version banner on stdout at the
    # VERSION debug level, which would precede the one JSON line
    if os.environ.get("NCCL_DEBUG", "").upper() not in ("INFO", "TRACE"):
        os.environ["NCCL_DEBUG"] = "NONE"
    import torch
    import torch.distributed as dist
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device - this backend has no CPU fallback")
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    elif gpus > 1:
        raise SystemExit("bench.py: launch with torch.distributed.run for --gpus > 1 (one rank per GPU)")
    torch.cuda.set_device(local_rank)
    return torch, dist


def _timers(torch, dist, world):
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

    return barrier, timed


def _measured_peak(torch, capi, lib, L, synthetic, np):
    """IMAD.WIDE issue rate measured in this run (dependent Montgomery products, b200_calib_mul_dev)."""
    st = torch.cuda.current_stream().cuda_stream
    nthreads, iters = 148 * 2048, 1000
    n32 = 2 * L.fp_l
    cbuf = torch.from_numpy(synthetic.rand_canonical(np.random.default_rng(1), nthreads, L.fp_l, L.p.bit_length()).view(np.uint8).reshape(-1)).cuda()
    capi.check(lib.b200_calib_mul_dev(L.id, 0, cbuf.data_ptr(), nthreads, iters, st))
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    capi.check(lib.b200_calib_mul_dev(L.id, 0, cbuf.data_ptr(), nthreads, iters, st))
    e1.record()
    torch.cuda.synchronize()
    return nthreads * iters * BW.p_mul(n32) / (e0.elapsed_time(e1) / 1e3)


def _blob_flow(kzg_mod, blob, z_int):
    """state/blobs.go:29-117 + crypto/blobs/blob.go:41-49: commitment, commitment + 128 cell proofs, opening proof."""
    c1 = blob.ComputeCommitment()
    c2, cells = blob.ComputeCommitmentAndCellProofs()
    proof, y = blob.ComputeProof(z_int)
    return c1, c2, cells, proof, y


def _st_blob(np):
    """statetransition-shaped blob: 32 + 1 + 60 * 36 = 2193 populated cells of BN254-sized values (state/blobs.go:58-96)."""
    rnd = np.random.default_rng(3)
    r254 = BW.CURVES["bn254"][2]
    cells = [int(rnd.integers(1, 1 << 62)) * int(rnd.integers(1, 1 << 62)) * int(rnd.integers(1, 1 << 62)) * int(rnd.integers(1, 1 << 62)) % r254
             for _ in range(2193)] + [0] * (4096 - 2193)
    return b"".join(v.to_bytes(32, "big") for v in cells)


def _load_srs(kzg_mod):
    g = os.path.join(ROOT, "tests", "golden")
    kzg_mod.load_trusted_setup(open(os.path.join(g, "kzg_g1_lagrange.bin"), "rb").read(),
                               open(os.path.join(g, "kzg_g1_monomial.bin"), "rb").read())


# ====================================================================================== independent proofs per GPU
def independent_arm(args, cfg_name, logn, rank, local_rank, world):
    torch, dist = _dist_setup(local_rank, world, args.gpus)
    import numpy as np
    from concurrent.futures import ThreadPoolExecutor
    from davinci_node_b200 import capi, synthetic
    capi.init(1 << local_rank)
    lib = capi.lib
    cfg = BW.CONFIGS[cfg_name]
    metric, unit = BW.METRICS[cfg_name]
    big = logn >= 23 or cfg["curve"] == "bw6_761"
    batch = args.batch or (1 if big else 4)
    inflight = args.inflight or (2 if big else 4)

    wl = synthetic.SyntheticWorkload(cfg["curve"], logn, seed=cfg["seed"], nb_public=cfg["nb_public"])
    t0 = time.time()
    h = wl.register()
    register_s = time.time() - t0
    info = capi.pk_info(h)
    L = wl.L
    nsol = 2
    sols = [wl.solution(seed=1000 * rank + i, mix=args.mix) for i in range(nsol)]
    r, s = 0x5EED5EED5EED5EED % L.r, (0x5EED << 64 | 0xABCDEF) % L.r
    dev_args = [wl.prove_args(sols[j % nsol], r, s, on_device=True) for j in range(batch)]
    host_args = [wl.prove_args(sols[j % nsol], r, s, on_device=False) for j in range(batch)]
    pool = ThreadPoolExecutor(max_workers=max(1, inflight))
    with_blob = cfg_name == "statetransition"
    if with_blob:
        from davinci_node_b200 import kzg
        _load_srs(kzg)
        blob = kzg.Blob(_st_blob(np))
        zpt = 0x1234567890ABCDEF1234567890ABCDEF % BW.CURVES["bn254"][2]

    def one(j, table, fn):
        torch.cuda.set_device(local_rank)
        pin, pout, out, keep = table[j]
        capi.check(fn(h, C.byref(pin), C.byref(pout), local_rank))
        if with_blob:
            _blob_flow(kzg, blob, zpt)

    def step_dev(i):
        list(pool.map(lambda j: one(j, dev_args, lib.b200_prove_dev), range(batch)))

    def step_host(i):
        list(pool.map(lambda j: one(j, host_args, lib.b200_prove), range(batch)))

    barrier, timed = _timers(torch, dist, world)
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
    value = world * batch * args.steps / (ms_total / 1e3)
    e2e_value = world * batch * args.steps / (ms_e2e / 1e3)

    # the uniform-random variant of the same workload (worst case, no zero / one skipping), rank 0's GPU
    uniform = None
    if args.mix == "witness" and not args.no_uniform and cfg_name == "voteverifier":
        usol = wl.solution(seed=777 + rank, mix="uniform")
        uargs = [wl.prove_args(usol, r, s, on_device=True) for j in range(batch)]

        def step_uni(i):
            list(pool.map(lambda j: one(j, uargs, lib.b200_prove_dev), range(batch)))

        step_uni(0)
        ms_uni = timed(step_uni, max(2, args.steps // 2))
        uniform = {"value": world * batch * max(2, args.steps // 2) / (ms_uni / 1e3), "unit": unit,
                   "workload": "same key, uniform-random wire vector (every A / B / K point is added in all 13 windows)"}
        del usol, uargs

    if rank != 0:
        if world > 1:
            dist.barrier()
            dist.destroy_process_group()
        return

    shape_cfg = {"workload": workload_text(cfg_name, wl, args.mix), "proofs_per_step_per_gpu": batch,
                 "parallelism": "independent proofs per GPU (%d in flight per GPU), no collective" % inflight,
                 "l2": "inputs (%.0f MB/proof + %.1f GB of key tables) larger than L2" % (wl.h2d_bytes() / 1e6, info["table_bytes"] / 1e9),
                 "key": {"table_stride": info["table_stride"], "table_gb": info["table_bytes"] / 1e9,
                         "wire_window_bits": info["wire_window"], "z_window_bits": info["z_window"],
                         "slots": info["slots"], "register_s": register_s}}
    line = {"metric": metric, "value": value, "unit": unit, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms_total / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": DTYPES[wl.L.name], "data": "synthetic", "config": shape_cfg,
            "e2e": {"value": e2e_value, "unit": unit, "h2d_bytes_per_step": (wl.h2d_bytes() + (131072 * 3 if with_blob else 0)) * batch,
                    "d2h_bytes_per_step": (wl.d2h_bytes() + (48 * 131 + 32 if with_blob else 0)) * batch, "ms_per_step": ms_e2e / args.steps},
            "gpu_launches": int(launches), "clocks": clocks}
    if uniform:
        line["uniform_variant"] = uniform

    if not args.no_roofline:
        line.update(roofline_detail(torch, capi, lib, synthetic, np, wl, sols[0], logn, info, ms_total / args.steps, batch))
    if not args.no_cpu_baseline and world == 1:
        # one FULL proof of the same workload on the host cores (plus two sample steps), product-free code path
        cb, _ = cpu_reference_run(cfg_name, logn, args.mix, 0, 1, args.cpu_parts)
        cb["sample"] = "ONE full %s proof at n=m=2^%d on %d threads: %.1f s" % (wl.L.name, logn, cb["cores"], cb["full_proof_seconds"])
        line["cpu_baseline"] = cb
    else:
        line["cpu_baseline"] = None
    print(json.dumps(line))
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


def roofline_detail(torch, capi, lib, synthetic, np, wl, sol, logn, info, ms_step, batch):
    """Rank 0, after the timed region (serialised, untimed): the dominant kernel timed alone with CUDA events."""
    L = wl.L
    st = torch.cuda.current_stream().cuda_stream
    n32 = 2 * L.fp_l
    bits = L.r.bit_length()
    peak_meas = _measured_peak(torch, capi, lib, L, synthetic, np)
    peak_nominal = 148 * 64 * 1965.0e6
    facts, facts_src = load_profile_facts()

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
        del pts
        tot, acc = (3, 0) if group == 1 else (4, 1)
        return ms[tot] / max(cnt[tot], 1), ms[acc] / max(cnt[acc], 1)

    # dominant kernel: k_msm_accumulate<G1> of the quotient (Z) MSM - dense full-width scalars over n-1 points.  Uniform
    # scalars: the SURVEY 8d work formula is exact for them.
    nZ = min(wl.nc, wl.n - 1)
    nB = len(wl.kB)
    g1_total_ms, g1_acc_ms = msm_alone(wl.pk.g1_Z, 1, nZ)
    g2_total_ms, g2_acc_ms = msm_alone(wl.pk.g2_B, 2, nB)
    g2mul = 3 if L.g2_deg == 2 else 1
    g1_alg = BW.msm_adds_star(nZ, bits) * 10 * BW.p_mul(n32)
    g2_alg = BW.msm_adds_star(nB, bits) * 10 * BW.p_mul(n32) * g2mul
    # the work the accumulate kernel actually does in table mode: one mixed addition per non-zero signed digit
    cz = info["z_window"]
    nwin_z = -(-(bits + 1) // cz)
    g1_actual = nZ * nwin_z * (1 - 2.0 ** -cz) * 10 * BW.p_mul(n32)
    peak_note = "IMAD.WIDE issue rate measured in this run (b200_calib_mul_dev); nominal 148*64*f = %.2f" % (peak_nominal / 1e12)
    fz = facts.get("k_msm_accumulate_g1", {})
    roofline = {"bound": "imad", "kernel": "k_msm_accumulate<G1> inside the quotient (Z) MSM: %d points, uniform %d-bit scalars, table "
                                            "mode c=%d (the kernel is %.0f%% of the MSM)" % (nZ, bits, cz, 100 * g1_acc_ms / g1_total_ms),
                "achieved": g1_alg / (g1_total_ms / 1e3) / 1e12, "peak": peak_meas / 1e12,
                "unit": "T wide-MAC/s (32x32->64 IMAD.WIDE)", "frac": g1_alg / (g1_total_ms / 1e3) / peak_meas,
                "work_count": "adds* of SURVEY 8d (min over c of n W(c) + 2 W(c) 2^(c-1)) x 10 field products x %d wide MACs, over "
                              "the WHOLE MSM's duration (sort + accumulate + tails)" % BW.p_mul(n32),
                "kernel_frac_actual_work": g1_actual / (g1_acc_ms / 1e3) / peak_meas,
                "kernel_frac_note": "the %d x %d x (1 - 2^-%d) mixed additions the kernel really performs, over the kernel's own "
                                    "duration" % (nZ, nwin_z, cz),
                "peak_source": peak_note, "frac_of_nominal": g1_alg / (g1_total_ms / 1e3) / peak_nominal,
                "traffic": (fz["dram_bytes"] * (nZ * nwin_z * (1 - 2.0 ** -cz)) / fz["adds"]) if fz else None,
                "traffic_source": ("%s (ncu --set full of the kernel inside one proof: %d mixed additions at c=%d, scaled by "
                                   "the additions of this launch)" % (facts_src, fz["adds"], fz["window_bits"])) if fz else None,
                "ncu_pipe_fmaheavy_pct": fz.get("fmaheavy_pct") if fz else None,
                "algorithmic_point_bytes": BW.msm_adds_star(nZ, bits) * L.affine_bytes(1),
                "launch_ms": g1_total_ms, "kernel_ms": g1_acc_ms, "algorithmic_macs_per_launch": g1_alg}
    f2 = facts.get("k_msm_accumulate_g2", {})
    g2_c, g2_nwin = cz, nwin_z     # the model's window of this set is taken to be the Z set's (same order of magnitude of points)
    roofline_g2 = {"bound": "imad", "kernel": "G2 MSM (Bs), %d points, uniform scalars (k_msm_accumulate<G2> = %.0f%% of it)" % (nB, 100 * g2_acc_ms / g2_total_ms),
                   "achieved": g2_alg / (g2_total_ms / 1e3) / 1e12, "peak": peak_meas / 1e12,
                   "unit": "T wide-MAC/s (32x32->64 IMAD.WIDE)", "frac": g2_alg / (g2_total_ms / 1e3) / peak_meas,
                   "peak_source": peak_note, "frac_of_nominal": g2_alg / (g2_total_ms / 1e3) / peak_nominal,
                   "traffic": (f2["dram_bytes"] * (nB * g2_nwin * (1 - 2.0 ** -g2_c)) / f2["adds"]) if f2 else None,
                   "traffic_source": ("%s (scaled by additions; includes the kernel's spill stores)" % facts_src) if f2 else None,
                   "launch_ms": g2_total_ms, "kernel_ms": g2_acc_ms, "algorithmic_macs_per_launch": g2_alg}
    # NTT passes alone
    ms = (C.c_double * 5)()
    cnt = (C.c_uint64 * 5)()
    reps = 3
    dom = C.c_uint64(0)
    omega, cg = BW.domain_constants(L.name, logn)
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
    # the whole quotient (7 transforms, fused pointwise work)
    b2, c2 = torch.zeros_like(buf), torch.zeros_like(buf)
    capi.check(lib.b200_compute_h_dev(dom.value, buf.data_ptr(), b2.data_ptr(), c2.data_ptr(), st))
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        capi.check(lib.b200_compute_h_dev(dom.value, buf.data_ptr(), b2.data_ptr(), c2.data_ptr(), st))
    e1.record()
    torch.cuda.synchronize()
    h_ms = e0.elapsed_time(e1) / reps
    capi.check(lib.b200_domain_release(dom.value))
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    hbm_peak = peaks.get("hbm_gbs", 6650.0)
    nfr = 2 * L.fr_l
    ntt_bytes = 2 * wl.n * L.fr_bytes
    roofline_ntt = {"bound": "hbm", "kernel": "k_ntt_pass (one 2^%d transform = %d passes)" % (logn, ntt_passes),
                    "achieved": ntt_bytes / (ntt_ms / 1e3) / 1e9, "peak": hbm_peak, "unit": "GB/s",
                    "frac": ntt_bytes / (ntt_ms / 1e3) / 1e9 / hbm_peak,
                    "peak_source": "MEASURED_PEAKS.json" if peaks else "fallback", "traffic": None,
                    "imad_frac": ((wl.n // 2) * logn * BW.p_mul(nfr)) / (ntt_ms / 1e3) / peak_meas, "launch_ms": ntt_ms,
                    "quotient_ms": h_ms, "quotient_hbm_frac": 7 * ntt_bytes / (h_ms / 1e3) / 1e9 / hbm_peak,
                    "quotient_imad_frac": 7 * ((wl.n // 2) * logn * BW.p_mul(nfr)) / (h_ms / 1e3) / peak_meas}
    # proof-level fraction with the scalar mix the generator really produced (zeros free, ones one addition ...)
    Wc = sol["W_canon"]
    sets = {"A": ~wl.infA, "B": ~wl.infB}
    kmask = np.ones(wl.m, dtype=bool)
    kmask[:wl.nb_public] = False
    kmask[wl.krs_skip] = False
    sets["K"] = kmask
    macs = 0
    detail = {}
    for name, mask in sets.items():
        z, o, sm, fu = BW.mix_stats(Wc[mask])
        adds = BW.msm_adds_sparse(z, o, sm, fu, bits)
        detail[name] = {"zero": z, "one": o, "lt_2_64": sm, "full": fu, "adds": adds}
        macs += adds * 10 * BW.p_mul(n32) * (1 + g2mul if name == "B" else 1)
    zc, oc, sc, fc = BW.mix_stats(Wc[wl.committed])
    macs += BW.msm_adds_sparse(zc, oc, sc, fc, bits) * 10 * BW.p_mul(n32)
    macs += BW.msm_adds_star(wl.n - 1, bits) * 10 * BW.p_mul(n32)
    macs += 7 * (wl.n // 2) * logn * BW.p_mul(nfr)
    return {"roofline": roofline, "roofline_g2": roofline_g2, "roofline_ntt": roofline_ntt,
            "step_imad_frac": macs * batch / (ms_step / 1e3) / peak_meas,
            "step_imad_note": "sparsity-aware SURVEY 8d formula (zeros free, ones 1 addition, <2^64 values ceil(65/c) windows) "
                              "summed over the 6 MSMs + 7 transforms of a proof, x proofs per step, over the step time",
            "step_scalar_mix": detail}


# ====================================================================================== one proof over N GPUs
def range_split_arm(args, cfg_name, logn, rank, local_rank, world):
    torch, dist = _dist_setup(local_rank, world, args.gpus)
    import numpy as np
    from davinci_node_b200 import capi, multi, synthetic
    from davinci_node_b200.gnark_types import ConstraintSystem
    capi.init(1 << local_rank)
    lib = capi.lib
    cfg = BW.CONFIGS[cfg_name]
    metric, unit = BW.METRICS[cfg_name]
    wl = synthetic.SyntheticWorkload(cfg["curve"], logn, seed=cfg["seed"], nb_public=cfg["nb_public"])
    pk = wl.build()
    L = wl.L
    ccs = ConstraintSystem(curve_id=L.id, nb_wires=wl.m, nb_public=wl.nb_public, nb_secret=0, L=[], R=[], O=[],
                           commitments=[{"private_committed": wl.committed.tolist(), "commitment_index": wl.commit_wire}])
    t0 = time.time()
    if world > 1:
        sub, sub_ccs, info = multi.slice_proving_key(pk, ccs, world, rank)
        h = multi.register_key_slice(sub, sub_ccs, info)
    else:
        h = wl.register()
        info = None
    register_s = time.time() - t0
    kinfo = capi.pk_info(h)
    wl.pk = None
    del pk
    sol = wl.solution(seed=77, mix=args.mix)
    r, s = 0x5EED5EED5EED5EED % L.r, (0x5EED << 64 | 0xABCDEF) % L.r
    frb = L.fr_bytes
    c0 = int(wl.committed[0])
    with_blob = cfg_name == "statetransition"
    if with_blob:
        from davinci_node_b200 import kzg
        _load_srs(kzg)
        blob = kzg.Blob(_st_blob(np))
        zpt = 0x1234567890ABCDEF1234567890ABCDEF % BW.CURVES["bn254"][2]

    def prove(Wd, ad, bd, cd):
        if world == 1:
            pin, pout, out, keep = prove.single
            capi.check(lib.b200_prove_dev(h, C.byref(pin), C.byref(pout), local_rank))
            return out
        pc = [(Wd[c0 * frb:(c0 + wl.n_c) * frb], wl.n_c)]
        return multi.prove_range_split(h, L, info, Wd, ad, bd, cd, wl.nc, r, s, True, pc,
                                       shard_quotient=not args.no_shard_quotient)

    if world == 1:
        prove.single = wl.prove_args(sol, r, s, on_device=True)
        host_single = wl.prove_args(sol, r, s, on_device=False)

    def step_dev(i):
        prove(sol["W_dev"], sol["a_dev"], sol["b_dev"], sol["c_dev"])
        if with_blob and rank == 0:
            _blob_flow(kzg, blob, zpt)

    # end to end: every rank copies the solver's output (pinned host) to its GPU, proves, and rank 0 reads the proof back
    stage = {k: torch.empty_like(sol[k + "_dev"]) for k in ("W", "a", "b", "c")}

    def step_host(i):
        if world == 1:
            pin, pout, out, keep = host_single
            capi.check(lib.b200_prove(h, C.byref(pin), C.byref(pout), local_rank))
        else:
            for k in ("W", "a", "b", "c"):
                stage[k].copy_(sol[k], non_blocking=True)
            proof = prove(stage["W"], stage["a"], stage["b"], stage["c"])
            assert len(proof["Ar"])       # host numpy arrays: the D2H read of the result
        if with_blob and rank == 0:
            _blob_flow(kzg, blob, zpt)

    barrier, timed = _timers(torch, dist, world)
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
    timeline = None
    if args.dump_timeline:
        TAGS = {0: "acc_g1", 1: "acc_g2", 2: "ntt_pass", 3: "msm_total_g1", 4: "msm_total_g2", 5: "sort", 6: "sched",
                7: "ovf", 8: "bucket_reduce", 9: "sums", 10: "inputs", 11: "assemble", 12: "pre_reduce"}
        barrier()
        capi.check(lib.b200_profile_enable(1))
        t0 = time.perf_counter()
        step_dev(0)
        torch.cuda.synchronize()
        wall_ms = (time.perf_counter() - t0) * 1e3
        cap = 4096
        buf = (C.c_double * (4 * cap))()
        nrec = C.c_uint64(0)
        capi.check(lib.b200_profile_timeline(buf, cap, C.byref(nrec)))
        capi.check(lib.b200_profile_enable(0))
        recs = [{"phase": TAGS.get(int(buf[4 * i]), str(int(buf[4 * i]))), "stream": int(buf[4 * i + 1]),
                 "start_ms": buf[4 * i + 2], "end_ms": buf[4 * i + 3]} for i in range(nrec.value)]
        timeline = {"rank": rank, "wall_ms": wall_ms, "records": recs}
        if rank == 0:
            json.dump(timeline, open(args.dump_timeline, "w"))
        barrier()
    if rank == 0:
        part_bytes = 5 * L.xyzz_bytes(1) + L.xyzz_bytes(2)
        config = {"workload": workload_text(cfg_name, wl, args.mix),
                  "parallelism": ("ONE proof split by point range over %d GPUs: per-GPU partial sums, NCCL all-gather of %d B per "
                                  "rank over NVLink, assembly" % (world, part_bytes)) if world > 1 else "one proof at a time on one GPU",
                  "l2": "inputs (%.0f MB/proof) and key tables (%.1f GB per GPU) larger than L2" % (wl.h2d_bytes() / 1e6, kinfo["table_bytes"] / 1e9),
                  "key": {"table_stride": kinfo["table_stride"], "table_gb_per_gpu": kinfo["table_bytes"] / 1e9,
                          "wire_window_bits": kinfo["wire_window"], "z_window_bits": kinfo["z_window"], "register_s": register_s}}
        line = {"metric": metric, "value": args.steps / (ms_total / 1e3), "unit": unit, "n_gpus": world, "steps": args.steps,
                "warmup": args.warmup, "ms_per_step": ms_total / args.steps, "higher_is_better": True,
                "scaling": "strong" if world > 1 else "weak", "vs_baseline": None, "dtype": DTYPES[L.name], "data": "synthetic",
                "config": config,
                "e2e": {"value": args.steps / (ms_e2e / 1e3), "unit": unit,
                        "h2d_bytes_per_step": wl.h2d_bytes() * world + (131072 * 3 if with_blob else 0),
                        "d2h_bytes_per_step": wl.d2h_bytes() + (48 * 131 + 32 if with_blob else 0), "ms_per_step": ms_e2e / args.steps},
                "gpu_launches": int(launches), "clocks": clocks,
                "collective": ({"op": "all_gather", "bytes_per_rank": part_bytes, "per_step": 1,
                                "quotient": ("sharded: 3 NCCL broadcasts of %d MB (coset evaluations of a, b, c) per proof" % (wl.n * frb // 1000000))
                                if not args.no_shard_quotient else "every GPU computes the whole quotient"} if world > 1 else None)}
        if not args.no_cpu_baseline and world == 1:
            cb, _ = cpu_reference_run(cfg_name, logn, args.mix, 0, 1, args.cpu_parts)
            line["cpu_baseline"] = cb
        else:
            line["cpu_baseline"] = None
        print(json.dumps(line))
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


# ====================================================================================== blob KZG alone
def blob_arm(args, rank, local_rank, world):
    torch, dist = _dist_setup(local_rank, world, args.gpus)
    import numpy as np
    from davinci_node_b200 import capi, kzg
    capi.init(1 << local_rank)
    _load_srs(kzg)
    blob = kzg.Blob(_st_blob(np))
    metric, unit = BW.METRICS["blob"]
    batch = args.batch or 16
    barrier, timed = _timers(torch, dist, world)

    def step(i):
        for _ in range(batch):
            blob.ComputeCommitment()

    for i in range(max(args.warmup, 1)):
        step(i)
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    l0 = capi.lib.b200_launch_count()
    ms = timed(step, args.steps)
    launches = capi.lib.b200_launch_count() - l0
    t0 = time.perf_counter()
    _blob_flow(kzg, blob, 0x1234567890ABCDEF)
    flow_ms = (time.perf_counter() - t0) * 1e3
    clocks = sampler.stop() if rank == 0 else None
    if rank == 0:
        v = world * batch * args.steps / (ms / 1e3)
        line = {"metric": metric, "value": v, "unit": unit, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
                "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
                "dtype": DTYPES["bls12_381"], "data": "synthetic",
                "config": {"workload": "EIP-4844 blob KZG commitment, 4096-point BLS12-381 MSM, statetransition-shaped blob "
                                       "(2193 populated cells), real ceremony SRS", "commitments_per_step_per_gpu": batch,
                           "l2": "latency-bound (128 KiB in, 48 B out per call): the SRS tables stay L2-resident by design"},
                "e2e": {"value": v, "unit": unit, "h2d_bytes_per_step": 131072 * batch, "d2h_bytes_per_step": 48 * batch},
                "gpu_launches": int(launches), "clocks": clocks,
                "blob_eval_data_flow_ms": flow_ms, "cpu_baseline": None}
        if not args.no_cpu_baseline and world == 1:
            line["cpu_baseline"] = cpu_blob_commit_baseline()
        print(json.dumps(line))
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
