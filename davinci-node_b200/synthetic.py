"""Synthetic voteverifier / aggregator / statetransition-shaped proving workloads.

The real circuits' ccs/pk are CDN artifacts that cannot be fetched here, so the bench and the
full-size tests prove against a *structured* synthetic key: every key point is [k]G for a known
64-bit k (made on the GPU with the fixed-base kernel).  The prover's arithmetic - the quotient, the
five MSMs with infinity filtering, the commitment handling, the assembly - is exactly the production
path, and because the discrete logs are known every proof element has a closed form the tests check
bit-for-bit at full size.  (It is not a Groth16-valid key; keys from a real trusted setup are covered
by the small oracle-generated cases in tests/test_gpu_prove.py.)

torch is used only as the device / pinned-host allocator.
"""
import ctypes as C

import numpy as np

from . import capi
from .curve_consts import CONSTS, domain_constants
from .layout import Layout

def _torch():
    import torch
    return torch


def _stream():
    return _torch().cuda.current_stream().cuda_stream


def rand_canonical(rng, n, limbs64, bits):
    """n uniform-ish canonical field elements < 2^(bits-1) as (n, limbs64) uint64."""
    a = rng.integers(0, 1 << 63, size=(n, limbs64), dtype=np.uint64) * np.uint64(2) + \
        rng.integers(0, 2, size=(n, limbs64), dtype=np.uint64)
    top_bits = bits - 1 - 64 * (limbs64 - 1)
    a[:, -1] &= np.uint64((1 << top_bits) - 1)
    return a


def witness_like(rng, n, limbs64, bits, mix="witness"):
    """Canonical scalars with the solved-witness value mix of SURVEY.md 8d-1:
    40% zero, 20% one, 25% < 2^64, 15% full width ('uniform': all full width)."""
    a = rand_canonical(rng, n, limbs64, bits)
    if mix == "uniform":
        return a
    if mix == "ones":          # every wire 1: each key's whole MSM lands in ONE bucket (maximum skew)
        a[:] = 0
        a[:, 0] = 1
        return a
    if mix == "zeros":         # nothing but the constant wire (set by the caller)
        a[:] = 0
        return a
    u = rng.random(n)
    zero = u < 0.40
    one = (u >= 0.40) & (u < 0.60)
    small = (u >= 0.60) & (u < 0.85)
    a[zero] = 0
    a[one] = 0
    a[one, 0] = 1
    a[small, 1:] = 0
    return a


def to_mont_dev(L: Layout, canon_u64: np.ndarray):
    """canonical (n, limbs) uint64 -> device tensor of Montgomery elements (uint8)."""
    torch = _torch()
    d = torch.from_numpy(np.ascontiguousarray(canon_u64).view(np.uint8).reshape(-1)).cuda()
    out = torch.empty_like(d)
    n = canon_u64.shape[0]
    capi.check(capi.lib.b200_dbg_field_op_dev(L.id, 1, 5, d.data_ptr(), None, out.data_ptr(), n, _stream()))
    return out


def fixed_base_dev(L: Layout, group, base_affine_np, k_u64: np.ndarray):
    """device tensor of affine points [k_i] base for 64-bit scalars k_i."""
    torch = _torch()
    n = len(k_u64)
    canon = np.zeros((n, L.fr_l), dtype=np.uint64)
    canon[:, 0] = k_u64
    ks = to_mont_dev(L, canon)
    base = torch.from_numpy(base_affine_np).cuda()
    out = torch.empty(n * L.affine_bytes(group), dtype=torch.uint8, device="cuda")
    capi.check(capi.lib.b200_fixed_base_dev(L.id, group, base.data_ptr(), ks.data_ptr(), n, out.data_ptr(), _stream()))
    return out


class SyntheticWorkload:
    """One circuit-shaped proving key with known discrete logs plus generators for solved witnesses."""

    def __init__(self, curve, logn, seed=0xD0A1, nb_public=6, n_commit_log=None, inf_a=0.30, inf_b=0.40,
                 nb_constraints=None):
        self.L = L = Layout(curve)
        self.logn, self.n = logn, 1 << logn
        self.m = self.n                              # nbWires = domain size
        self.nc = nb_constraints if nb_constraints is not None else self.n - 3
        self.nb_public = nb_public
        n_c = 1 << (n_commit_log if n_commit_log is not None else max(1, logn - 4))
        self.n_c = min(n_c, self.m - nb_public - 2)
        self.g1, self.g2 = CONSTS[L.id]["g1"], CONSTS[L.id]["g2"]
        self.rng = np.random.default_rng(seed)
        self.seed = seed
        rng = self.rng
        m = self.m
        self.infA = rng.random(m) < inf_a
        self.infB = rng.random(m) < inf_b
        self.infA[:nb_public] = False
        self.committed = np.arange(nb_public, nb_public + self.n_c, dtype=np.uint32)
        self.commit_wire = nb_public + self.n_c
        self.krs_skip = np.concatenate([self.committed, np.array([self.commit_wire], dtype=np.uint32)])
        nA, nB = int((~self.infA).sum()), int((~self.infB).sum())
        nK = m - nb_public - len(self.krs_skip)
        k64 = lambda cnt: rng.integers(1, 1 << 62, size=cnt, dtype=np.uint64)
        # discrete logs (w.r.t. the generators) of every key element
        self.kA, self.kB, self.kK, self.kZ = k64(nA), k64(nB), k64(nK), k64(self.n - 1)
        self.kBasis, self.sigma = k64(self.n_c), int(rng.integers(2, 1 << 62))
        self.alpha, self.beta, self.delta = (int(x) for x in k64(3))
        self.pk = None
        self.handle = None

    # ---- key construction (GPU fixed-base) and registration
    def build(self):
        from .gnark_types import ProvingKey
        L = self.L
        g1 = L.enc_affine([self.g1], 1)
        g2 = L.enc_affine([self.g2], 2)
        host = lambda t: t.cpu().numpy()
        single = lambda grp, base, k: host(fixed_base_dev(L, grp, base, np.array([k], dtype=np.uint64)))
        omega, coset = domain_constants(L.id, self.logn)
        sig_basis = (self.kBasis.astype(object) * self.sigma)        # sigma * k exceeds 64 bits: reduce on host
        sig_canon = np.zeros((self.n_c, L.fr_l), dtype=np.uint64)
        for i, v in enumerate(sig_basis):
            v = int(v) % L.r
            for j in range(L.fr_l):
                sig_canon[i, j] = (v >> (64 * j)) & 0xFFFFFFFFFFFFFFFF
        torch = _torch()
        base1 = torch.from_numpy(g1).cuda()
        sig_s = to_mont_dev(L, sig_canon)
        sig_pts = torch.empty(self.n_c * L.affine_bytes(1), dtype=torch.uint8, device="cuda")
        capi.check(capi.lib.b200_fixed_base_dev(L.id, 1, base1.data_ptr(), sig_s.data_ptr(), self.n_c,
                                                sig_pts.data_ptr(), _stream()))
        self.pk = ProvingKey(
            curve_id=L.id, domain_cardinality=self.n,
            domain_generator=L.enc_fr([omega]), domain_coset_gen=L.enc_fr([coset]),
            g1_alpha=single(1, g1, self.alpha), g1_beta=single(1, g1, self.beta), g1_delta=single(1, g1, self.delta),
            g1_A=host(fixed_base_dev(L, 1, g1, self.kA)), g1_B=host(fixed_base_dev(L, 1, g1, self.kB)),
            g1_Z=host(fixed_base_dev(L, 1, g1, self.kZ)), g1_K=host(fixed_base_dev(L, 1, g1, self.kK)),
            g2_beta=single(2, g2, self.beta), g2_delta=single(2, g2, self.delta),
            g2_B=host(fixed_base_dev(L, 2, g2, self.kB)),
            infinity_a=self.infA.astype(np.uint8), infinity_b=self.infB.astype(np.uint8),
            commitment_keys=[{"Basis": host(fixed_base_dev(L, 1, g1, self.kBasis)), "BasisExpSigma": host(sig_pts)}],
        )
        return self.pk

    def register(self):
        """b200_pk_register with this key; returns the handle."""
        from . import prover
        from .gnark_types import ConstraintSystem
        if self.pk is None:
            self.build()
        ccs = ConstraintSystem(curve_id=self.L.id, nb_wires=self.m, nb_public=self.nb_public, nb_secret=0, L=[], R=[],
                               O=[], commitments=[{"private_committed": self.committed.tolist(),
                                                   "commitment_index": self.commit_wire}])
        self.handle = prover.register_proving_key(self.pk, ccs)
        return self.handle

    # ---- solved witnesses
    def solution(self, seed, mix="witness", pinned=True):
        """A solved assignment: W (m), a, b, c (nc) as Montgomery byte tensors on the host (pinned) and
        their canonical uint64 form (for closed-form checks).  c = a*b is computed on the GPU."""
        torch = _torch()
        L = self.L
        rng = np.random.default_rng(seed)
        Wc = witness_like(rng, self.m, L.fr_l, L.r.bit_length(), mix)
        Wc[0] = 0
        Wc[0, 0] = 1
        ac = rand_canonical(rng, self.nc, L.fr_l, L.r.bit_length())
        bc = rand_canonical(rng, self.nc, L.fr_l, L.r.bit_length())
        Wd, ad, bd = to_mont_dev(L, Wc), to_mont_dev(L, ac), to_mont_dev(L, bc)
        cd = torch.empty_like(ad)
        capi.check(capi.lib.b200_dbg_field_op_dev(L.id, 1, 2, ad.data_ptr(), bd.data_ptr(), cd.data_ptr(), self.nc, _stream()))
        torch.cuda.synchronize()

        def hostbuf(t):
            h = torch.empty(t.shape, dtype=torch.uint8, pin_memory=pinned)
            h.copy_(t)
            return h

        return {"W": hostbuf(Wd), "a": hostbuf(ad), "b": hostbuf(bd), "c": hostbuf(cd),
                "W_dev": Wd, "a_dev": ad, "b_dev": bd, "c_dev": cd, "W_canon": Wc, "a_canon": ac, "b_canon": bc}

    def prove_args(self, sol, r, s, on_device=False):
        """(b200_prove_in, b200_proof_out, keepalive) for one proof of `sol`."""
        torch = _torch()
        L = self.L
        frb = L.fr_bytes
        key = "_dev" if on_device else ""
        rb, sb = L.enc_fr([r]), L.enc_fr([s])
        keep = [rb, sb]
        W = sol["W" + key]
        pin = capi.ProveIn()
        pin.wires = capi.Slice(W.data_ptr(), self.m)
        for name in ("a", "b", "c"):
            setattr(pin, name, capi.Slice(sol[name + key].data_ptr(), self.nc))
        if on_device:
            rs = torch.from_numpy(np.concatenate([rb, sb])).cuda()
            keep.append(rs)
            pin.r, pin.s = rs.data_ptr(), rs.data_ptr() + frb
        else:
            pin.r, pin.s = rb.ctypes.data, sb.ctypes.data
        pin.nb_commitments = 1
        pcs = (capi.Slice * 1)()
        pcs[0] = capi.Slice(W.data_ptr() + int(self.committed[0]) * frb, self.n_c)   # committed wires are contiguous
        pin.priv_committed = pcs
        pin.fold_challenge = None
        g1b, g2b = L.affine_bytes(1), L.affine_bytes(2)
        if on_device:
            out = torch.zeros(3 * g1b + g2b, dtype=torch.uint8, device="cuda")
        else:
            out = torch.zeros(3 * g1b + g2b, dtype=torch.uint8, pin_memory=True)
        base = out.data_ptr()
        pout = capi.ProofOut(base, base + 3 * g1b, base + g1b, base + 2 * g1b)   # ar, bs, krs, pok
        keep += [pcs, out]
        return pin, pout, out, keep

    def decode_proof(self, out):
        L = self.L
        g1b = L.affine_bytes(1)
        buf = out.cpu().numpy()
        return {"Ar": L.dec_affine(buf[:g1b], 1)[0], "Krs": L.dec_affine(buf[g1b:2 * g1b], 1)[0],
                "CommitmentPok": L.dec_affine(buf[2 * g1b:3 * g1b], 1)[0], "Bs": L.dec_affine(buf[3 * g1b:], 2)[0]}

    def h2d_bytes(self):
        return (self.m + 3 * self.nc + 2) * self.L.fr_bytes

    def d2h_bytes(self):
        return 3 * self.L.affine_bytes(1) + self.L.affine_bytes(2)
