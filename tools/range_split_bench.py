#!/usr/bin/env python3
"""BASELINE.json configs[2]: ONE aggregator-shaped proof (BW6-761) with its MSMs split by point range over the GPUs
of a box.  Launch with torchrun, one rank per GPU:

  python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 \
      tools/range_split_bench.py [--curve bw6_761] [--logn 20] [--proofs 4]

Every rank builds the same structured synthetic key (same seed), registers ITS slice, and the ranks prove together:
partial sums per GPU, NCCL all-gather of the six partial points, assembly.  Rank 0 also registers the whole key and
checks that the range-split proof is bit-identical to the single-GPU proof.  Prints one JSON line on rank 0."""
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
    ap.add_argument("--curve", default="bw6_761")
    ap.add_argument("--logn", type=int, default=20)
    ap.add_argument("--proofs", type=int, default=4)
    args = ap.parse_args()
    if os.environ.get("NCCL_DEBUG", "").upper() not in ("INFO", "TRACE"):
        os.environ["NCCL_DEBUG"] = "NONE"
    import numpy as np
    import torch
    import torch.distributed as dist
    rank, world = int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    from davinci_node_b200 import capi, multi, prover, synthetic
    from davinci_node_b200.gnark_types import ConstraintSystem
    capi.init(1 << local)
    wl = synthetic.SyntheticWorkload(args.curve, args.logn, seed=0xA66)
    pk = wl.build()
    L = wl.L
    ccs = ConstraintSystem(curve_id=L.id, nb_wires=wl.m, nb_public=wl.nb_public, nb_secret=0, L=[], R=[], O=[],
                           commitments=[{"private_committed": wl.committed.tolist(), "commitment_index": wl.commit_wire}])
    sub, sub_ccs, info = multi.slice_proving_key(pk, ccs, world, rank)
    t0 = time.time()
    h = multi.register_key_slice(sub, sub_ccs, info)
    t_reg = time.time() - t0
    sol = wl.solution(seed=77)
    r, s = 0x5EED5EED5EED5EED % L.r, (0x5EED << 64 | 0xABCDEF) % L.r
    frb = L.fr_bytes
    Wd, ad, bd, cd = sol["W_dev"], sol["a_dev"], sol["b_dev"], sol["c_dev"]
    pc = [(Wd[int(wl.committed[0]) * frb:(int(wl.committed[0]) + wl.n_c) * frb], wl.n_c)]

    def one():
        return multi.prove_range_split(h, L, info, Wd, ad, bd, cd, wl.nc, r, s, True, pc)

    proof = one()
    torch.cuda.synchronize()
    dist.barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(args.proofs):
        proof = one()
    e1.record()
    torch.cuda.synchronize()
    dist.barrier()
    ms = multi.max_over_ranks(e0.elapsed_time(e1) / args.proofs, device="cuda")
    out = {"config": "range-split proof, %s, n=m=2^%d, %d GPUs" % (args.curve, args.logn, world), "n_gpus": world,
           "ms_per_proof": ms, "proofs_per_s": 1e3 / ms, "slice_register_s": t_reg,
           "collective": "NCCL all_gather of %d bytes per rank" % (5 * L.xyzz_bytes(1) + L.xyzz_bytes(2))}
    if rank == 0:
        # single-GPU reference on the whole key: must be bit-identical
        prover.release_proving_key(sub)
        wl.pk = pk
        hf = wl.register()
        pin, pout, outbuf, keep = wl.prove_args(sol, r, s, on_device=True)
        capi.check(capi.lib.b200_prove_dev(hf, C.byref(pin), C.byref(pout), local))
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for _ in range(args.proofs):
            capi.check(capi.lib.b200_prove_dev(hf, C.byref(pin), C.byref(pout), local))
        torch.cuda.synchronize()
        out["single_gpu_ms_per_proof"] = (time.perf_counter() - t0) / args.proofs * 1e3
        g1b = L.affine_bytes(1)
        buf = outbuf.cpu().numpy()
        same = (np.array_equal(buf[:g1b], proof["Ar"]) and np.array_equal(buf[g1b:2 * g1b], proof["Krs"]) and
                np.array_equal(buf[2 * g1b:3 * g1b], proof["CommitmentPok"]) and np.array_equal(buf[3 * g1b:], proof["Bs"]))
        out["bit_identical_to_single_gpu"] = bool(same)
        print(json.dumps(out))
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
