"""Mirror of the reference's `prover.Setup` (/root/reference/prover/setup.go:15-28 -> groth16.Setup).

    Setup(ccs) -> (pk, vk)

The toxic waste is sampled on the host (crypto-strength, or pinned by `SetSetupRandomness` in tests);
the per-wire polynomial evaluations A_i(tau), B_i(tau), C_i(tau) are O(nnz) host big-integer work
(Lagrange basis at tau); every key POINT - the part that dominates gnark's Setup, a
fixed-base scalar multiplication per wire and per domain element in G1 and G2 - is computed on the
GPU by `b200_fixed_base_dev`.  Key layout and conventions follow gnark: infinity points of A / B are
dropped and flagged in InfinityA / InfinityB, K excludes public, committed and commitment wires, Z is
stored bit-reversed, commitment keys are the 1/gamma-scaled K points of the committed wires.
"""
import secrets
from dataclasses import dataclass, field

import numpy as np

from . import capi
from .curve_consts import CONSTS, domain_constants
from .gnark_types import ConstraintSystem, ProvingKey
from .layout import Layout


@dataclass
class VerifyingKey:
    """groth16_<curve>.VerifyingKey (points as gnark-layout byte buffers)."""
    curve_id: int
    g1_alpha: np.ndarray
    g1_K: np.ndarray                      # public + commitment wires, 1/gamma scaled
    g2_beta: np.ndarray
    g2_gamma: np.ndarray
    g2_delta: np.ndarray
    commitment_keys: list = field(default_factory=list)   # [{'G': G2, 'GSigmaNeg': G2}]
    public_and_commitment_committed: list = field(default_factory=list)


_setup_randomness = None


def SetSetupRandomness(fn):
    """Test hook: fn(curve_id, n_commitments) -> dict(tau, alpha, beta, gamma, delta, sigmas=[...])."""
    global _setup_randomness
    _setup_randomness = fn


def _sample(curve_id, n_commit):
    if _setup_randomness is not None:
        return _setup_randomness(curve_id, n_commit)
    r = Layout(curve_id).r
    nz = lambda: 1 + secrets.randbelow(r - 1)
    return dict(tau=nz(), alpha=nz(), beta=nz(), gamma=nz(), delta=nz(), sigmas=[nz() for _ in range(n_commit)])


def _bitrev(i, logn):
    r = 0
    for _ in range(logn):
        r = (r << 1) | (i & 1)
        i >>= 1
    return r


def _fixed_base(L, group, base_pt, scalars):
    """GPU: affine bytes of [k] base for every k in scalars (Python ints mod r)."""
    import torch
    n = len(scalars)
    if n == 0:
        return np.zeros(0, dtype=np.uint8)
    st = torch.cuda.current_stream().cuda_stream
    ks = torch.from_numpy(L.enc_fr(scalars)).cuda()
    base = torch.from_numpy(L.enc_affine([base_pt], group)).cuda()
    out = torch.empty(n * L.affine_bytes(group), dtype=torch.uint8, device="cuda")
    capi.check(capi.lib.b200_fixed_base_dev(L.id, group, base.data_ptr(), ks.data_ptr(), n, out.data_ptr(), st))
    torch.cuda.synchronize()
    return out.cpu().numpy()


def Setup(ccs: ConstraintSystem):
    """prover/setup.go:15: trusted setup for `ccs`; returns (ProvingKey, VerifyingKey)."""
    capi.init_once()
    L = Layout(ccs.curve_id)
    q = L.r
    g1, g2 = CONSTS[L.id]["g1"], CONSTS[L.id]["g2"]
    tw = _sample(L.id, len(ccs.commitments))
    tau, alpha, beta, gamma, delta = (tw[k] % q for k in ("tau", "alpha", "beta", "gamma", "delta"))
    n = 1
    while n < max(ccs.nb_constraints, 1):
        n *= 2
    logn = n.bit_length() - 1
    omega, coset = domain_constants(L.id, logn)
    # Lagrange basis at tau: L_j(tau) = (tau^n - 1) w^j / (n (tau - w^j))  (batch inversion)
    zn = (pow(tau, n, q) - 1) % q
    ws, dens = [], []
    wj = 1
    for _ in range(n):
        ws.append(wj)
        dens.append(n * (tau - wj) % q)
        wj = wj * omega % q
    pref = [1] * (n + 1)
    for i, d in enumerate(dens):
        pref[i + 1] = pref[i] * d % q
    inv = pow(pref[n], -1, q)
    lag = [0] * n
    for i in range(n - 1, -1, -1):
        lag[i] = zn * ws[i] % q * (inv * pref[i] % q) % q
        inv = inv * dens[i] % q
    m = ccs.nb_wires
    A, B, Cc = [0] * m, [0] * m, [0] * m
    for k in range(ccs.nb_constraints):
        lk = lag[k]
        for w, cf in ccs.L[k]:
            A[w] = (A[w] + cf * lk) % q
        for w, cf in ccs.R[k]:
            B[w] = (B[w] + cf * lk) % q
        for w, cf in ccs.O[k]:
            Cc[w] = (Cc[w] + cf * lk) % q
    dinv, ginv = pow(delta, -1, q), pow(gamma, -1, q)
    Kfull = [(beta * A[i] + alpha * B[i] + Cc[i]) % q for i in range(m)]
    committed, commit_wires = set(), set()
    for cm in ccs.commitments:
        committed.update(cm["private_committed"])
        commit_wires.add(cm["commitment_index"])
    pub = [i for i in range(m) if i < ccs.nb_public or i in commit_wires]
    priv = [i for i in range(ccs.nb_public, m) if i not in committed and i not in commit_wires]
    z_nat = []
    tj = 1
    for _ in range(n):
        z_nat.append(tj * zn % q * dinv % q)
        tj = tj * tau % q
    Z = [z_nat[_bitrev(i, logn)] for i in range(n)][: n - 1]
    infA = np.array([a == 0 for a in A], dtype=np.uint8)
    infB = np.array([b == 0 for b in B], dtype=np.uint8)
    fb1 = lambda ks: _fixed_base(L, 1, g1, ks)
    fb2 = lambda ks: _fixed_base(L, 2, g2, ks)
    keys = []
    vkeys = []
    for cm, sg in zip(ccs.commitments, tw["sigmas"]):
        basis = [Kfull[i] * ginv % q for i in cm["private_committed"]]
        keys.append({"Basis": fb1(basis), "BasisExpSigma": fb1([b * sg % q for b in basis])})
        gk = 1 + secrets.randbelow(q - 1) if _setup_randomness is None else (sg * 7 + 3) % q
        vkeys.append({"G": fb2([gk]), "GSigmaNeg": fb2([(-sg * gk) % q])})
    pk = ProvingKey(
        curve_id=L.id, domain_cardinality=n, domain_generator=L.enc_fr([omega]), domain_coset_gen=L.enc_fr([coset]),
        g1_alpha=fb1([alpha]), g1_beta=fb1([beta]), g1_delta=fb1([delta]),
        g1_A=fb1([a for a in A if a]), g1_B=fb1([b for b in B if b]), g1_Z=fb1(Z),
        g1_K=fb1([Kfull[i] * dinv % q for i in priv]),
        g2_beta=fb2([beta]), g2_delta=fb2([delta]), g2_B=fb2([b for b in B if b]),
        infinity_a=infA, infinity_b=infB, commitment_keys=keys)
    vk = VerifyingKey(curve_id=L.id, g1_alpha=pk.g1_alpha, g1_K=fb1([Kfull[i] * ginv % q for i in pub]),
                      g2_beta=pk.g2_beta, g2_gamma=fb2([gamma]), g2_delta=pk.g2_delta, commitment_keys=vkeys,
                      public_and_commitment_committed=[list(cm.get("public_committed", [])) for cm in ccs.commitments])
    return pk, vk
