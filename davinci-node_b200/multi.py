"""Multi-GPU plumbing (one process per GPU, torch.distributed): SURVEY.md 8e.

1. Independent proofs: `shard_items` gives each rank its share of a batch; no data-path collective.
2. Range-split MSM: rank g owns points/scalars [g*n/G, (g+1)*n/G); each GPU reduces its range to ONE
   XYZZ point (<= 384 bytes), the partials are all-gathered (NCCL over NVLink on GPUs, gloo in the CPU
   tests) and every rank adds the G partials on its device (`b200_sum_partials_dev`).  NCCL cannot
   reduce with the elliptic-curve group law, hence gather + local adds; the payload is a few hundred
   bytes so the collective is pure latency.
"""
import numpy as np


def shard_range(n, world, rank, first_weight=1.0):
    """Contiguous [lo, hi) range of rank `rank` when n items are split over `world` ranks.  first_weight < 1 gives
    rank 0 a smaller share (it also carries the commitment proof of knowledge and the alpha / beta / delta terms of a
    range-split proof); the other ranks split the rest evenly."""
    if first_weight == 1.0 or world == 1:
        return n * rank // world, n * (rank + 1) // world
    first = int(n * first_weight / world)
    if rank == 0:
        return 0, first
    rest = n - first
    return first + rest * (rank - 1) // (world - 1), first + rest * rank // (world - 1)


def first_rank_weight(world):
    """Share of rank 0 in a range-split proof relative to an even split: its extra work (the PoK MSM over the
    commitment's 2^(logn-4) wires, run before the quotient) is ~1.5% of a proof, i.e. 1.5% * world of a slice."""
    return max(0.7, 1.0 - 0.015 * world)


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


# ------------------------------------------------------------------------------------------------
# Range-split PROVE: one large proof over several GPUs (aggregator / statetransition)
def slice_proving_key(pk, ccs, world, rank):
    """The part of a gnark proving key GPU `rank` holds: the wire-indexed arrays (A, B, K and the
    infinity flags) restricted to its wire range, its range of Z, and - on rank 0 only - alpha / beta /
    delta and the commitment keys.  Returns (sub_pk, sub_ccs, info) ready for register_proving_key."""
    from .gnark_types import ConstraintSystem, ProvingKey
    from .layout import Layout
    L = Layout(pk.curve_id)
    g1b, g2b = L.affine_bytes(1), L.affine_bytes(2)
    m = len(pk.infinity_a)
    fw = first_rank_weight(world) if pk.commitment_keys else 1.0
    wlo, whi = shard_range(m, world, rank, fw)
    infA = np.asarray(pk.infinity_a, dtype=np.uint8)
    infB = np.asarray(pk.infinity_b, dtype=np.uint8)
    cumA = np.concatenate([[0], np.cumsum(infA == 0)])
    cumB = np.concatenate([[0], np.cumsum(infB == 0)])
    skip = np.zeros(m, dtype=bool)
    for w in ccs.krs_skip_wires():
        skip[w] = True
    in_k = np.ones(m, dtype=bool)
    in_k[:ccs.nb_public] = False
    in_k &= ~skip
    cumK = np.concatenate([[0], np.cumsum(in_k)])
    nz = len(pk.g1_Z) // g1b
    zlo, zhi = shard_range(nz, world, rank, fw)
    zero1, zero2 = np.zeros(g1b, dtype=np.uint8), np.zeros(g2b, dtype=np.uint8)
    first = rank == 0
    pts = lambda buf, lo, hi, sz: np.ascontiguousarray(buf[lo * sz:hi * sz])
    sub = ProvingKey(
        curve_id=pk.curve_id, domain_cardinality=pk.domain_cardinality,
        domain_generator=pk.domain_generator, domain_coset_gen=pk.domain_coset_gen,
        g1_alpha=pk.g1_alpha if first else zero1, g1_beta=pk.g1_beta if first else zero1,
        g1_delta=pk.g1_delta if first else zero1,
        g1_A=pts(pk.g1_A, cumA[wlo], cumA[whi], g1b), g1_B=pts(pk.g1_B, cumB[wlo], cumB[whi], g1b),
        g1_Z=pts(pk.g1_Z, zlo, zhi, g1b), g1_K=pts(pk.g1_K, cumK[wlo], cumK[whi], g1b),
        g2_beta=pk.g2_beta if first else zero2, g2_delta=pk.g2_delta if first else zero2,
        g2_B=pts(pk.g2_B, cumB[wlo], cumB[whi], g2b),
        infinity_a=infA[wlo:whi].copy(), infinity_b=infB[wlo:whi].copy(),
        commitment_keys=list(pk.commitment_keys) if first else [])
    nb_public_local = int(min(max(ccs.nb_public - wlo, 0), whi - wlo))
    local_skip = sorted(int(w - wlo) for w in ccs.krs_skip_wires() if wlo <= w < whi)
    # the sub constraint system only carries what registration reads (sizes and the K skip list)
    sub_ccs = ConstraintSystem(curve_id=pk.curve_id, nb_wires=whi - wlo, nb_public=nb_public_local, nb_secret=0,
                               L=[], R=[], O=[], commitments=[{"private_committed": local_skip, "commitment_index": None}]
                               if local_skip else [])
    sub_ccs.krs_skip_wires = lambda: set(local_skip)
    return sub, sub_ccs, {"wire_range": (wlo, whi), "z_offset": zlo, "first": first,
                          "domain_size": int(pk.domain_cardinality)}


def register_key_slice(sub_pk, sub_ccs, info):
    """b200_pk_register for one slice (z_offset set); returns the handle."""
    from . import prover
    return prover.register_proving_key(sub_pk, sub_ccs, z_offset=info["z_offset"])


def coset_evals_inplace(handle, buf):
    """buf (domain_size fr, natural order, zero padded, on the current device) <- its evaluations on the coset
    g<omega>: the per-vector half of the quotient (b200_pk_coset_evals_dev), enqueued on the current torch stream."""
    import torch
    from . import capi
    capi.check(capi.lib.b200_pk_coset_evals_dev(handle, buf.data_ptr(), torch.cuda.current_device(),
                                                torch.cuda.current_stream().cuda_stream))


def prove_partial(handle, L, info, W_dev, a_dev, b_dev, c_dev, nc, r, s, priv_committed_dev=None, fold_challenge=None,
                  abc_form=0):
    """Partial sums of one key slice: W_dev is the FULL wire vector on this device (the slice is taken
    here), a/b/c are full.  Returns a device tensor of 5*xyzz(1)+xyzz(2) bytes whose K slot already holds
    K_g + s*Ar_g + r*Bs1_g.  fold_challenge (int) is required with more than one commitment.

    The C ABI reads its device inputs on internal streams, so everything the caller produced on the current
    torch stream (W / a / b / c, the r,s upload below) is synchronised before the call."""
    import ctypes as C
    import torch
    from . import capi
    frb = L.fr_bytes
    wlo, whi = info["wire_range"]
    rs = torch.from_numpy(np.concatenate([L.enc_fr([r]), L.enc_fr([s])])).cuda()
    pin = capi.ProveIn()
    pin.wires = capi.Slice(W_dev.data_ptr() + wlo * frb, whi - wlo)
    pin.a, pin.b, pin.c = (capi.Slice(t.data_ptr(), nc) for t in (a_dev, b_dev, c_dev))
    pin.r, pin.s = rs.data_ptr(), rs.data_ptr() + frb
    k = len(priv_committed_dev) if (priv_committed_dev and info["first"]) else 0
    pin.nb_commitments = k
    pcs = (capi.Slice * max(k, 1))()
    for i in range(k):
        t, cnt = priv_committed_dev[i]
        pcs[i] = capi.Slice(t.data_ptr(), cnt)
    pin.priv_committed = pcs
    pin.fold_challenge = None
    pin.abc_form = abc_form
    if k > 1:
        if fold_challenge is None:
            raise ValueError("fold_challenge is required with more than one commitment")
        fc = torch.from_numpy(L.enc_fr([fold_challenge])).cuda()
        pin.fold_challenge = fc.data_ptr()
    out = torch.zeros(5 * L.xyzz_bytes(1) + L.xyzz_bytes(2), dtype=torch.uint8, device="cuda")
    torch.cuda.current_stream().synchronize()
    capi.check(capi.lib.b200_prove_partial_dev(handle, C.byref(pin), out.data_ptr(), torch.cuda.current_device()))
    return out


def assemble(L, partials_dev, nparts, r, s, have_pok):
    """Final assembly from the gathered partials; returns dict of affine byte arrays."""
    import torch
    from . import capi
    frb, g1b, g2b = L.fr_bytes, L.affine_bytes(1), L.affine_bytes(2)
    rs = torch.from_numpy(np.concatenate([L.enc_fr([r]), L.enc_fr([s])])).cuda()
    out = torch.zeros(3 * g1b + g2b, dtype=torch.uint8, device="cuda")
    st = torch.cuda.current_stream().cuda_stream
    capi.check(capi.lib.b200_assemble_dev(L.id, partials_dev.data_ptr(), nparts, rs.data_ptr(), rs.data_ptr() + frb,
                                          1 if have_pok else 0, out.data_ptr(), st))
    buf = out.cpu().numpy()
    return {"Ar": buf[:g1b], "Krs": buf[g1b:2 * g1b], "CommitmentPok": buf[2 * g1b:3 * g1b], "Bs": buf[3 * g1b:]}


def quotient_owners(world):
    """Which rank transforms a, b, c when the quotient is sharded (two transforms per vector)."""
    return (0, 1 % world, 2 % world) if world >= 3 else (0, 1 % world, 0)


def prove_range_split(handle, L, info, W_dev, a_dev, b_dev, c_dev, nc, r, s, have_pok, priv_committed_dev=None,
                      group=None, fold_challenge=None, shard_quotient=True):
    """One proof over all ranks of `group`: partial sums on every GPU, all-gather (NCCL / NVLink), local
    assembly.  Every rank returns the same proof.

    shard_quotient: instead of every GPU computing the whole quotient (7 transforms), the owner of each of a, b, c
    turns it into coset evaluations (2 transforms) and broadcasts them over NVLink (domain_size * fr_bytes each);
    every GPU then runs only the pointwise division and the last inverse transform."""
    import torch
    import torch.distributed as dist
    world = dist.get_world_size(group)
    if shard_quotient and world > 1:
        rank = dist.get_rank(group)
        n = info["domain_size"]
        frb = L.fr_bytes
        bufs = info.setdefault("_qbuf", [torch.empty(n * frb, dtype=torch.uint8, device="cuda") for _ in range(3)])
        owners = quotient_owners(world)
        for buf, src_vec, owner in zip(bufs, (a_dev, b_dev, c_dev), owners):
            if rank == owner:
                buf[:nc * frb].copy_(src_vec[:nc * frb])
                buf[nc * frb:].zero_()
                coset_evals_inplace(handle, buf)
        for buf, owner in zip(bufs, owners):
            dist.broadcast(buf, src=dist.get_global_rank(group, owner) if group is not None else owner, group=group)
        part = prove_partial(handle, L, info, W_dev, bufs[0], bufs[1], bufs[2], n, r, s, priv_committed_dev, fold_challenge,
                             abc_form=1)
    else:
        part = prove_partial(handle, L, info, W_dev, a_dev, b_dev, c_dev, nc, r, s, priv_committed_dev, fold_challenge)
    allp = gather_partials(part, group)
    return assemble(L, allp, world, r, s, have_pok)
