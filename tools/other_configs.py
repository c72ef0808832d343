#!/usr/bin/env python3
"""BASELINE.json configs[2] and [3] on one B200: aggregator-shaped proof (BW6-761), statetransition-shaped
proof (BN254) and the EIP-4844 blob KZG commitment.  Prints one JSON object per config and writes
gpurun_out/other_configs.json.

  python tools/other_configs.py [--agg-logn 20] [--st-logn 22]
"""
import argparse
import ctypes as C
import json
import os
import sys
import time
from concurrent.futures import ThreadPoolExecutor

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from davinci_node_b200 import capi, kzg, synthetic  # noqa: E402


def run_proofs(cname, logn, nb_public, proofs=4, inflight=2):
    t0 = time.time()
    wl = synthetic.SyntheticWorkload(cname, logn, seed=logn * 3 + 1, nb_public=nb_public)
    h = wl.register()
    t_reg = time.time() - t0
    L = wl.L
    sols = [wl.solution(seed=50 + i) for i in range(2)]
    r, s = 12345678901234567890 % L.r, 98765432109876543210 % L.r
    dev = [wl.prove_args(sols[j % 2], r, s, on_device=True) for j in range(proofs)]
    host = [wl.prove_args(sols[j % 2], r, s, on_device=False) for j in range(proofs)]
    pool = ThreadPoolExecutor(max_workers=inflight)

    def one(args, fn):
        torch.cuda.set_device(0)
        pin, pout, out, keep = args
        capi.check(fn(h, C.byref(pin), C.byref(pout), 0))

    res = {}
    for name, arglist, fn in (("resident", dev, capi.lib.b200_prove_dev), ("e2e", host, capi.lib.b200_prove)):
        list(pool.map(lambda a: one(a, fn), arglist[:inflight]))          # warm-up
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        list(pool.map(lambda a: one(a, fn), arglist))
        torch.cuda.synchronize()
        res[name + "_proofs_per_s"] = proofs / (time.perf_counter() - t0)
    # single proof latency
    t0 = time.perf_counter()
    one(dev[0], capi.lib.b200_prove_dev)
    res["single_proof_ms"] = (time.perf_counter() - t0) * 1e3
    res.update({"curve": cname, "log_n": logn, "register_s": t_reg, "h2d_bytes_per_proof": wl.h2d_bytes()})
    from davinci_node_b200 import prover
    prover.release_proving_key(wl.pk)
    del wl, sols, dev, host
    torch.cuda.empty_cache()
    return res


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--agg-logn", type=int, default=20)
    ap.add_argument("--st-logn", type=int, default=22)
    args = ap.parse_args()
    capi.init(1)
    out = {}
    out["aggregator_bw6_761"] = run_proofs("bw6_761", args.agg_logn, nb_public=2)
    print(json.dumps(out["aggregator_bw6_761"]), flush=True)
    out["statetransition_bn254"] = run_proofs("bn254", args.st_logn, nb_public=9)
    print(json.dumps(out["statetransition_bn254"]), flush=True)
    # blob commitment: statetransition-shaped blob (2193 populated cells), real EIP-4844 SRS fixture
    srs = open(os.path.join(ROOT, "tests", "golden", "kzg_g1_lagrange.bin"), "rb").read()
    t0 = time.perf_counter()
    kzg.load_trusted_setup(srs)
    t_srs = time.perf_counter() - t0
    rnd = np.random.default_rng(3)
    cells = [int(x) for x in rnd.integers(1, 1 << 62, size=2193)] + [0] * (4096 - 2193)
    blob = kzg.Blob(b"".join(v.to_bytes(32, "big") for v in cells))
    blob.ComputeCommitment()
    ts = []
    for _ in range(20):
        t0 = time.perf_counter()
        c = blob.ComputeCommitment()
        ts.append((time.perf_counter() - t0) * 1e3)
    out["blob_commit"] = {"srs_register_s": t_srs, "commit_ms_median": float(np.median(ts)), "commit_ms_min": min(ts),
                          "note": "host blob in (128 KiB), 48-byte commitment out, 4096-point BLS12-381 MSM in table mode"}
    print(json.dumps(out["blob_commit"]), flush=True)
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    json.dump(out, open(os.path.join(ROOT, "gpurun_out", "other_configs.json"), "w"), indent=1)


if __name__ == "__main__":
    main()
