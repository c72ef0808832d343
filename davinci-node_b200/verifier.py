"""Host-side mirror of gnark's groth16.Verify over the C ABI (b200_pairing_check, b200_msm).

The reference verifies every proof right after proving it (/root/reference/circuits/artifacts.go:595-613,
`groth16.Verify(proof, vk, publicWitness, opts...)`) and deploys the same check on-chain
(/root/reference/config/statetransition_vkey.sol:653-746).  Split of the work, as in gnark:

  host (this file)   commitment challenges = hash-to-field over (commitment || public committed inputs), the fold
                     challenge of several commitments, the bookkeeping of which points are paired;
  GPU  (C ABI)       the subgroup check of Bs ((r - 1) Bs == -Bs, b200_msm on G2; Ar / Krs are checked by the Miller loop),
                     the public-input MSM  L = sum_i pub_i K_i + sum_j chal_j K_(np+j) + sum_j C_j   (b200_msm),
                     the folded commitment sum_j fold^j C_j (b200_msm), and the two product-of-pairings checks
                         e(C_fold, GSigmaNeg) e(PoK, G) == 1
                         e(-Ar, Bs) e(alpha, beta) e(L, gamma) e(Krs, delta) == 1            (b200_pairing_check).

Errors are raised (gnark returns an error); there is no CPU fallback.  Points are byte buffers in gnark memory layout
(layout.py), the public witness is the vector of public wire values WITHOUT the constant-one wire.
"""
import ctypes as C

import numpy as np

from . import capi
from . import hash_to_field as H2F
from .layout import Layout
from .prover import _hash_kind
from .setup import VerifyingKey


class VerificationError(RuntimeError):
    """groth16.Verify returned an error (pairing check failed, malformed proof, wrong witness size)."""


def _neg_g1(L: Layout, buf):
    pt = L.dec_affine(buf, 1)[0]
    return L.enc_affine([None if pt is None else (pt[0], (-pt[1]) % L.p)], 1)


def _msm(L: Layout, group, point_bufs, scalars, device):
    pts = np.concatenate([np.asarray(b, dtype=np.uint8) for b in point_bufs])
    sc = L.enc_fr(scalars)
    out = np.zeros(L.affine_bytes(group), dtype=np.uint8)
    capi.check(capi.lib.b200_msm(L.id, group, pts.ctypes.data, sc.ctypes.data, len(scalars), out.ctypes.data, device))
    return out


def _msm_g1(L: Layout, point_bufs, scalars, device):
    return _msm(L, 1, point_bufs, scalars, device)


def _neg_g2(L: Layout, buf):
    pt = L.dec_affine(buf, 2)[0]
    if pt is None:
        return L.enc_affine([None], 2)
    y = pt[1]
    ny = (-y) % L.p if L.coord_width(2) == 1 else ((-y[0]) % L.p, (-y[1]) % L.p)
    return L.enc_affine([(pt[0], ny)], 2)


def _g2_in_subgroup(L: Layout, buf, device):
    """r * Q == infinity, checked as (r - 1) * Q == -Q with the G2 MSM (gnark: G2Affine.IsInSubGroup)."""
    got = _msm(L, 2, [buf], [L.r - 1], device)
    return bytes(got) == bytes(_neg_g2(L, buf))


def pairing_check(curve_id, g1_bufs, g2_bufs, device=-1, want_gt=False):
    """prod_i e(P_i, Q_i) == 1 on the GPU (gnark-crypto <curve>.PairingCheck).  g1_bufs / g2_bufs: equally long lists of
    G1Affine / G2Affine byte buffers.  Returns the predicate (and, with want_gt, the F_{p^k} coefficients of the product's
    reduced Tate pairing as Python ints)."""
    L = Layout(curve_id)
    if len(g1_bufs) != len(g2_bufs):
        raise ValueError("pairing_check: %d G1 points for %d G2 points" % (len(g1_bufs), len(g2_bufs)))
    capi.init_once()
    n = len(g1_bufs)
    g1 = np.concatenate([np.asarray(b, dtype=np.uint8) for b in g1_bufs]) if n else np.zeros(1, dtype=np.uint8)
    g2 = np.concatenate([np.asarray(b, dtype=np.uint8) for b in g2_bufs]) if n else np.zeros(1, dtype=np.uint8)
    if n and (g1.size != n * L.affine_bytes(1) or g2.size != n * L.affine_bytes(2)):
        raise ValueError("pairing_check: point buffers of the wrong size")
    res = C.c_int(0)
    gt = np.zeros(int(capi.lib.b200_gt_bytes(L.id)), dtype=np.uint8)
    capi.check(capi.lib.b200_pairing_check(L.id, g1.ctypes.data, g2.ctypes.data, n, C.byref(res),
                                           gt.ctypes.data if want_gt else None, device))
    ok = bool(res.value)
    return (ok, tuple(L.dec_fp(gt))) if want_gt else ok


def pairing_check_batch(curve_id, g1_bufs, g2_bufs, pairs_per_check, device=-1):
    """n independent product checks in one call (b200_pairing_check_batch: one GPU thread per pair and per check).
    g1_bufs / g2_bufs: flat lists of len n * pairs_per_check point buffers, check c owning the slice
    [c * pairs_per_check, (c + 1) * pairs_per_check).  Returns a list of True / False / None (None: a G1 point of that
    check lies outside the order-r subgroup)."""
    L = Layout(curve_id)
    per = int(pairs_per_check)
    if per <= 0 or len(g1_bufs) != len(g2_bufs) or len(g1_bufs) % per:
        raise ValueError("pairing_check_batch: %d / %d points for checks of %d pairs" % (len(g1_bufs), len(g2_bufs), per))
    n_checks = len(g1_bufs) // per
    if not n_checks:
        return []
    capi.init_once()
    g1 = np.concatenate([np.asarray(b, dtype=np.uint8) for b in g1_bufs])
    g2 = np.concatenate([np.asarray(b, dtype=np.uint8) for b in g2_bufs])
    if g1.size != len(g1_bufs) * L.affine_bytes(1) or g2.size != len(g2_bufs) * L.affine_bytes(2):
        raise ValueError("pairing_check_batch: point buffers of the wrong size")
    res = np.zeros(n_checks, dtype=np.int32)
    capi.check(capi.lib.b200_pairing_check_batch(L.id, g1.ctypes.data, g2.ctypes.data, per, n_checks, res.ctypes.data, device))
    return [None if r < 0 else bool(r) for r in res]


def Verify(proof, vk: VerifyingKey, public_witness, *opts, device=-1):
    """groth16.Verify: returns None when the proof verifies, raises VerificationError otherwise."""
    L = Layout(vk.curve_id)
    if proof.curve_id != vk.curve_id:
        raise VerificationError("proof and verifying key are on different curves")
    hash_kind = _hash_kind(opts)
    q = L.r
    ab1 = L.affine_bytes(1)
    nK = len(vk.g1_K) // ab1
    ncm = len(vk.commitment_keys)
    pub = [1] + [int(v) % q for v in public_witness]
    if len(pub) + ncm != nK:
        raise VerificationError("invalid witness size, got %d, expected %d" % (len(pub) - 1, nK - ncm - 1))
    if len(proof.Commitments) != ncm:
        raise VerificationError("invalid number of commitments in the proof")
    capi.init_once()
    committed = vk.public_and_commitment_committed or [[] for _ in range(ncm)]
    chals = []
    for cm, wires in zip(proof.Commitments, committed):
        pt = L.dec_affine(cm, 1)[0]
        chals.append(H2F.commitment_challenge(hash_kind, pt, [pub[w] for w in wires], q, L.fp_bytes))
    if ncm:
        # pedersen.BatchVerifyMultiVk: sum_j e(fold^j C_j, GSigmaNeg_j) + e(PoK, G) (every key shares G)
        g1s = [np.asarray(proof.Commitments[0], dtype=np.uint8)]
        if ncm > 1:
            fold = H2F.fold_challenge(chals, q)
            g1s += [_msm_g1(L, [cm], [pow(fold, j, q)], device) for j, cm in enumerate(proof.Commitments) if j > 0]
        g1s.append(proof.CommitmentPok)
        g2s = [k["GSigmaNeg"] for k in vk.commitment_keys] + [vk.commitment_keys[0]["G"]]
        if not pairing_check(L.id, g1s, g2s, device):
            raise VerificationError("commitment proof of knowledge: pairing check failed")
    kbufs = [vk.g1_K[i * ab1:(i + 1) * ab1] for i in range(nK)]
    Lpt = _msm_g1(L, kbufs + list(proof.Commitments), pub + chals + [1] * ncm, device)
    # proof.isValid(): Ar, Krs, Bs in the order-r subgroups.  The Miller loop over r reports the G1 points (r * P must
    # end at infinity); Bs needs its own check
    if not _g2_in_subgroup(L, proof.Bs, device):
        raise VerificationError("proof is invalid: points not in the correct subgroup")
    try:
        ok = pairing_check(L.id, [_neg_g1(L, proof.Ar), vk.g1_alpha, Lpt, proof.Krs],
                           [proof.Bs, vk.g2_beta, vk.g2_gamma, vk.g2_delta], device)
    except capi.B200Error as e:
        if "subgroup" in str(e):
            raise VerificationError("proof is invalid: points not in the correct subgroup") from e
        raise
    if not ok:
        raise VerificationError("pairing doesn't match")


def verify(proof, vk, public_witness, *opts, device=-1) -> bool:
    """Boolean convenience around Verify."""
    try:
        Verify(proof, vk, public_witness, *opts, device=device)
        return True
    except VerificationError:
        return False
