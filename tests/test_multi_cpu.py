"""CPU (gloo, world_size 2): the host-side logic of the multi-GPU paths - range sharding, batch
sharding, rank-ordered gathering of fixed-size partials and max-over-ranks timing."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    import sys
    sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    from davinci_node_b200 import multi
    n = 1001
    lo, hi = multi.shard_range(n, world, rank)
    part = torch.full((384,), rank + 1, dtype=torch.uint8)          # stands in for one XYZZ partial
    part[0] = hi - lo & 0xFF
    allp = multi.gather_partials(part)
    tmax = multi.max_over_ranks(10.0 * (rank + 1))
    q.put((rank, lo, hi, allp.numpy().tolist(), tmax, multi.shard_items(7, world, rank)))
    dist.destroy_process_group()


def test_range_split_plumbing_gloo():
    world = 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=120) for _ in range(world))
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    (r0, lo0, hi0, all0, t0, it0), (r1, lo1, hi1, all1, t1, it1) = res
    assert (lo0, hi0, lo1, hi1) == (0, 500, 500, 1001)                 # disjoint cover
    assert all0 == all1 and len(all0) == 2 * 384                       # same gathered buffer everywhere
    assert all0[1] == 1 and all0[384 + 1] == 2                         # rank order
    assert all0[0] == 500 & 0xFF and all0[384] == 501 & 0xFF
    assert t0 == t1 == 20.0                                            # max over ranks
    assert sorted(it0 + it1) == list(range(7)) and not set(it0) & set(it1)


def test_shard_range_properties():
    from davinci_node_b200 import multi
    for n in (0, 1, 7, 4096, (1 << 22) - 1):
        for world in (1, 2, 4, 8):
            cover = []
            for r in range(world):
                lo, hi = multi.shard_range(n, world, r)
                assert 0 <= lo <= hi <= n
                cover.append((lo, hi))
            assert cover[0][0] == 0 and cover[-1][1] == n
            assert all(cover[i][1] == cover[i + 1][0] for i in range(world - 1))
            sizes = [hi - lo for lo, hi in cover]
            assert max(sizes) - min(sizes) <= 1
