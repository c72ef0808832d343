"""Multi-GPU plumbing (one process per GPU, torch.distributed): SURVEY.md 8e.

1. Independent proofs: `shard_items` gives each rank its share of a batch; no data-path collective.
2. Range-split MSM: rank g owns points/scalars [g*n/G, (g+1)*n/G); each GPU reduces its range to ONE
   XYZZ point (<= 384 bytes), the partials are all-gathered (NCCL over NVLink on GPUs, gloo in the CPU
   tests) and every rank adds the G partials on its device (`b200_sum_partials_dev`).  NCCL cannot
   reduce with the elliptic-curve group law, hence gather + local adds; the payload is a few hundred
   bytes so the collective is pure latency.
"""
import numpy as np


def shard_range(n, world, rank):
    """Contiguous [lo, hi) range of rank `rank` when n items are split over `world` ranks."""
    lo = n * rank // world
    hi = n * (rank + 1) // world
    return lo, hi


def shard_items(n_items, world, rank):
    """Indices of a batch of independent items (proofs) handled by `rank` (round-robin)."""
    return list(range(rank, n_items, world))


def gather_partials(partial, group=None):
    """all_gather of one fixed-size byte tensor per rank -> (world * nbytes) tensor, rank order."""
    import torch
    import torch.distributed as dist
    world = dist.get_world_size(group)
    out = [torch.empty_like(partial) for _ in range(world)]
    dist.all_gather(out, partial, group=group)
    return torch.cat(out)


def max_over_ranks(value, device="cpu", group=None):
    import torch
    import torch.distributed as dist
    t = torch.tensor([float(value)], dtype=torch.float64, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.MAX, group=group)
    return float(t.item())


def msm_range_split(curve_id, grp, d_points_local, d_scalars_local, n_local, group=None):
    """Range-split MSM: every rank passes ITS slice (device tensors); returns the affine result bytes
    (numpy, gnark layout), identical on every rank."""
    import torch
    from . import capi
    from .layout import Layout
    L = Layout(curve_id)
    st = torch.cuda.current_stream().cuda_stream
    partial = torch.zeros(L.xyzz_bytes(grp), dtype=torch.uint8, device="cuda")
    capi.check(capi.lib.b200_msm_dev(L.id, grp, d_points_local.data_ptr() if n_local else None,
                                     d_scalars_local.data_ptr() if n_local else None, n_local, partial.data_ptr(), 0, st))
    allp = gather_partials(partial, group)
    out = torch.zeros(L.affine_bytes(grp), dtype=torch.uint8, device="cuda")
    count = allp.numel() // L.xyzz_bytes(grp)
    capi.check(capi.lib.b200_sum_partials_dev(L.id, grp, allp.data_ptr(), count, out.data_ptr(), st))
    torch.cuda.synchronize()
    return out.cpu().numpy()
