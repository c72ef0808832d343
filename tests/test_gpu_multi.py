"""GPU, 2+ devices: range-split MSM with the per-GPU partials all-gathered over NCCL, bit-exact vs
the oracle.  Skipped on a single-GPU box (the gloo CPU test covers the plumbing)."""
import os
import random
import socket
import sys

import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _worker(rank, world, port, q):
    try:
        _worker_body(rank, world, port, q)
    except BaseException as e:      # a dead worker must not leave the parent waiting for its queue entry
        import traceback
        q.put((rank, {"exception": False, "trace": traceback.format_exc()[-1500:]}))
        raise


def _worker_body(rank, world, port, q):
    import torch
    import torch.distributed as dist
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    sys.path.insert(0, root)
    sys.path.insert(0, os.path.join(root, "tests"))
    from davinci_node_b200 import capi, layout, multi
    from oracle import curve as OC
    from gpu_util import rand_points
    capi.init(1 << rank)
    out = {}
    for cname, grp, n in (("bls12_377", 1, 301), ("bw6_761", 2, 90)):
        cx = OC.ctx(cname)
        L = layout.Layout(cname)
        rnd = random.Random(1234)                 # same data on every rank
        pts = rand_points(cx, grp, n, rnd)
        sc = [rnd.randrange(cx.r) for _ in range(n)]
        lo, hi = multi.shard_range(n, world, rank)
        dp = torch.from_numpy(L.enc_affine(pts[lo:hi], grp)).cuda()
        ds = torch.from_numpy(L.enc_fr(sc[lo:hi])).cuda()
        got = L.dec_affine(multi.msm_range_split(L.id, grp, dp, ds, hi - lo), grp)[0]
        out[cname] = (got == cx.group(grp).msm(pts, sc))
    # ---- range-split PROVE over NCCL: each rank holds a slice of an oracle-generated key
    from davinci_node_b200 import prover
    from oracle import groth16 as OG
    from oracle_bridge import ccs_from_oracle, pk_from_oracle
    cname = "bls12_377"
    cx = OC.ctx(cname)
    L = layout.Layout(cname)
    qf = cx.r
    rnd = random.Random(4321)
    cs, W = OG.synthetic_circuit(40, 4, qf, seed=8, n_commit=1, n_private_committed=3)
    tox = OG.Toxic(*(rnd.randrange(1, qf) for _ in range(5)), sigmas=[rnd.randrange(1, qf)])
    opk, ex = OG.setup(cs, cx, tox)
    ccs, pk = ccs_from_oracle(cs, L.id), pk_from_oracle(opk, L.id)
    r, s = rnd.randrange(qf), rnd.randrange(qf)
    want = OG.prove(cs, opk, W, r, s, cx)
    a, b, c = OG.constraint_values(cs, W, qf)
    dev = lambda v: torch.from_numpy(L.enc_fr(v)).cuda()
    Wd, ad, bd, cd = dev(W), dev(a), dev(b), dev(c)
    committed = cs.commitments[0]["private_committed"]
    pc = [(dev([W[i] for i in committed]), len(committed))]
    sub, sub_ccs, info = multi.slice_proving_key(pk, ccs, world, rank)
    h = multi.register_key_slice(sub, sub_ccs, info)
    got = multi.prove_range_split(h, L, info, Wd, ad, bd, cd, len(a), r, s, True, pc)
    out["prove_split"] = (L.dec_affine(got["Ar"], 1)[0] == want["Ar"] and L.dec_affine(got["Bs"], 2)[0] == want["Bs"]
                          and L.dec_affine(got["Krs"], 1)[0] == want["Krs"]
                          and L.dec_affine(got["CommitmentPok"], 1)[0] == want["CommitmentPok"])
    # the unsharded quotient (every GPU computes the whole of it) must give the same bytes
    got_u = multi.prove_range_split(h, L, info, Wd, ad, bd, cd, len(a), r, s, True, pc, shard_quotient=False)
    out["prove_split_unsharded_quotient"] = all(np.array_equal(got_u[k], got[k]) for k in ("Ar", "Bs", "Krs", "CommitmentPok"))
    prover.release_proving_key(sub)
    q.put((rank, out))
    dist.destroy_process_group()


def test_range_split_msm_nccl():
    import torch
    import torch.multiprocessing as mp
    world = min(torch.cuda.device_count(), 4)
    if world < 2:
        pytest.skip("needs >= 2 GPUs")
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = [q.get(timeout=240) for _ in range(world)]
    for rank, out in res:
        assert "trace" not in out, out["trace"]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    for rank, out in res:
        assert all(out.values()), (rank, out)
