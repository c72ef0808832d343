"""Host-side mirror of the reference's `prover` package (prover/config.go, prover_cpu.go,
prover_gpu.go, setup.go) over the C ABI of libb200groth16.so.

The reference is Go; this container has no Go toolchain, so the drop-in Go file
(davinci-node_b200/go/prover_b200.go) cannot be compiled here.  This module keeps the same names,
argument meaning and error behaviour so the parity tests read like the reference's own tests:

    Prove(curve, ccs, pk, assignment, *opts)          -> Proof          prover_cpu.go:19 / prover_gpu.go:66
    ProveWithWitness(curve, ccs, pk, witness, *opts)  -> Proof          prover_cpu.go:49 / prover_gpu.go:111
    GPUProver / GPUProverWithWitness                  (the B200 path)   prover_gpu.go:86,121
    CPUProver / CPUProverWithWitness                  raise: this backend has no CPU prover
    SetProver(fn), UseGPUProver                                         config.go:16-56
    SetRandomness(fn)   test-only hook pinning (r, s)                   SURVEY.md 8b "pinned-randomness hook"

`ccs`, `pk`, witness and proof objects mirror gnark's types as plain containers of byte buffers in
gnark-crypto memory layout (see gnark_types.py).  Errors surface as exceptions (Go: `(nil, err)`);
nothing falls back to a CPU prover (north_star).
"""
import ctypes as C
import os
import secrets
import threading

import numpy as np

from . import capi
from .gnark_types import ConstraintSystem, Proof, ProvingKey, Witness
from .layout import Layout
from .setup import Setup, SetSetupRandomness, VerifyingKey  # noqa: F401  (prover/setup.go:15)

# env GPU_PROVER, read at import like the reference's init() (prover/config.go:18-26)
UseGPUProver = os.environ.get("GPU_PROVER", "true").lower() in ("1", "true", "yes")


class ProverError(RuntimeError):
    pass


def _slice(arr, elem_bytes):
    if arr is None or len(arr) == 0:
        return capi.Slice(None, 0)
    assert arr.dtype == np.uint8 and arr.flags["C_CONTIGUOUS"]
    return capi.Slice(arr.ctypes.data, len(arr) // elem_bytes)


_lock = threading.Lock()
_registered = {}      # id(pk) -> (pk, handle)   device-resident copies keyed by proving-key identity


def _default_randomness(curve_id):
    L = Layout(curve_id)
    return secrets.randbelow(L.r), secrets.randbelow(L.r)


_randomness = _default_randomness


def SetRandomness(fn):
    """Test hook: fn(curve_id) -> (r, s) as Python ints; None restores crypto-strength sampling."""
    global _randomness
    _randomness = fn or _default_randomness


def register_proving_key(pk: ProvingKey, ccs: ConstraintSystem, z_offset=0):
    """Upload pk to the selected GPUs once; later proofs reuse the resident copy (the reference's
    icicle path does the same lazily inside gpugroth16.Prove, prover/prover_gpu.go:33-56)."""
    with _lock:
        ent = _registered.get(id(pk))
        if ent is not None:
            return ent[1]
        capi.init_once()
        L = Layout(pk.curve_id)
        g1b, g2b, frb = L.affine_bytes(1), L.affine_bytes(2), L.fr_bytes
        d = capi.PkDesc()
        d.curve = pk.curve_id
        d.domain_size = pk.domain_cardinality
        keep = [pk.domain_generator, pk.domain_coset_gen, pk.g1_alpha, pk.g1_beta, pk.g1_delta, pk.g2_beta, pk.g2_delta]
        d.generator, d.coset_gen = keep[0].ctypes.data, keep[1].ctypes.data
        d.g1_alpha, d.g1_beta, d.g1_delta = keep[2].ctypes.data, keep[3].ctypes.data, keep[4].ctypes.data
        d.g2_beta, d.g2_delta = keep[5].ctypes.data, keep[6].ctypes.data
        d.g1_A, d.g1_B = _slice(pk.g1_A, g1b), _slice(pk.g1_B, g1b)
        d.g1_Z, d.g1_K = _slice(pk.g1_Z, g1b), _slice(pk.g1_K, g1b)
        d.g2_B = _slice(pk.g2_B, g2b)
        infA = np.ascontiguousarray(pk.infinity_a, dtype=np.uint8)
        infB = np.ascontiguousarray(pk.infinity_b, dtype=np.uint8)
        d.infinity_a, d.infinity_b = _slice(infA, 1), _slice(infB, 1)
        d.nb_wires = len(infA)
        d.nb_public = ccs.nb_public
        skip = np.ascontiguousarray(sorted(ccs.krs_skip_wires()), dtype=np.uint32)
        d.krs_skip = capi.Slice(skip.ctypes.data if len(skip) else None, len(skip))
        k = len(pk.commitment_keys)
        d.nb_commitments = k
        basis = (capi.Slice * max(k, 1))()
        sigma = (capi.Slice * max(k, 1))()
        for i, key in enumerate(pk.commitment_keys):
            basis[i] = _slice(key["Basis"], g1b)
            sigma[i] = _slice(key["BasisExpSigma"], g1b)
        d.commit_basis = basis
        d.commit_basis_exp_sigma = sigma
        d.z_offset = z_offset
        h = C.c_uint64(0)
        capi.check(capi.lib.b200_pk_register(C.byref(d), C.byref(h)))
        _registered[id(pk)] = (pk, h.value)
        return h.value


def release_proving_key(pk: ProvingKey):
    with _lock:
        ent = _registered.pop(id(pk), None)
    if ent:
        capi.check(capi.lib.b200_pk_release(ent[1]))


class ProverOption:
    """backend.ProverOption mirror: only the option that changes proof bytes on this path is modelled - the
    hash-to-field function of the BSB22 commitment challenge (SURVEY.md 8a "prover options")."""

    def __init__(self, hash_kind):
        self.hash_kind = hash_kind


def WithProverHashToFieldFunction(kind):
    """backend.WithProverHashToFieldFunction: 'default' (hash_to_field "bsb22-commitment") or 'solidity' (keccak256)."""
    return ProverOption(kind)


def WithProverTargetSolidityVerifier():
    """solidity.WithProverTargetSolidityVerifier(backend.GROTH16) (circuits/statetransition/artifacts.go:18,
    circuits/results/artifacts.go:17): keccak256 commitment hash, matching config/statetransition_vkey.sol:668-677."""
    return ProverOption("solidity")


def _hash_kind(opts):
    kind = "default"
    for o in opts:
        if isinstance(o, ProverOption):
            kind = o.hash_kind
    return kind


def _gpu_prove(curve_id, ccs: ConstraintSystem, pk: ProvingKey, w: Witness, device=-1, hash_kind="default"):
    if pk.curve_id != curve_id:
        raise ProverError("proving key type mismatch for curve %s: got a %s key" % (curve_id, pk.curve_id))
    L = Layout(curve_id)
    handle = register_proving_key(pk, ccs)

    # --- host side of groth16.Prove: solve (with the BSB22 commitment hint calling the GPU)
    def commit_hint(i, values_bytes):
        out = np.zeros(L.affine_bytes(1), dtype=np.uint8)
        capi.check(capi.lib.b200_commit(handle, i, _slice(values_bytes, L.fr_bytes), out.ctypes.data, device))
        return out

    sol = ccs.solve(w, commit_hint, hash_kind)            # raises on an unsatisfied constraint
    r, s = _randomness(curve_id)
    rb, sb = L.enc_fr([r]), L.enc_fr([s])
    pin = capi.ProveIn()
    pin.wires = _slice(sol.W, L.fr_bytes)
    pin.a, pin.b, pin.c = _slice(sol.A, L.fr_bytes), _slice(sol.B, L.fr_bytes), _slice(sol.C, L.fr_bytes)
    pin.r, pin.s = rb.ctypes.data, sb.ctypes.data
    k = len(sol.private_committed)
    pin.nb_commitments = k
    pcs = (capi.Slice * max(k, 1))()
    for i, v in enumerate(sol.private_committed):
        pcs[i] = _slice(v, L.fr_bytes)
    pin.priv_committed = pcs
    fold = L.enc_fr([sol.fold_challenge]) if k > 1 else None
    pin.fold_challenge = fold.ctypes.data if fold is not None else None
    proof = Proof(curve_id)
    proof.Ar = np.zeros(L.affine_bytes(1), dtype=np.uint8)
    proof.Krs = np.zeros(L.affine_bytes(1), dtype=np.uint8)
    proof.Bs = np.zeros(L.affine_bytes(2), dtype=np.uint8)
    proof.CommitmentPok = np.zeros(L.affine_bytes(1), dtype=np.uint8)
    proof.Commitments = sol.commitments
    pout = capi.ProofOut(proof.Ar.ctypes.data, proof.Bs.ctypes.data, proof.Krs.ctypes.data,
                         proof.CommitmentPok.ctypes.data)
    capi.check(capi.lib.b200_prove(handle, C.byref(pin), C.byref(pout), device))
    return proof


# ---------------------------------------------------------------- reference-named entry points
def GPUProverWithWitness(curve, ccs, pk, w, *opts):
    """prover/prover_gpu.go:121-131 (the B200 path; errors are raised, never swallowed)."""
    try:
        return _gpu_prove(curve, ccs, pk, w, hash_kind=_hash_kind(opts))
    except capi.B200Error as e:
        raise ProverError(str(e)) from e


def GPUProver(curve, ccs, pk, assignment, *opts):
    """prover/prover_gpu.go:86-96: build the witness from the assignment, then prove."""
    w = Witness.from_assignment(assignment, curve)
    return GPUProverWithWitness(curve, ccs, pk, w, *opts)


def CPUProver(curve, ccs, pk, assignment, *opts):
    raise ProverError("b200 backend has no CPU prover (no CPU fallback by design)")


def CPUProverWithWitness(curve, ccs, pk, w, *opts):
    raise ProverError("b200 backend has no CPU prover (no CPU fallback by design)")


def _default_prover(curve, ccs, pk, assignment, *opts):
    if not UseGPUProver:
        return CPUProver(curve, ccs, pk, assignment, *opts)
    return GPUProver(curve, ccs, pk, assignment, *opts)


_prover = _default_prover


def SetProver(fn):
    """prover/config.go:54-56 (tests inject e.g. a debug prover)."""
    global _prover
    _prover = fn or _default_prover


def Prove(curve, ccs, pk, assignment, *opts):
    """prover/prover_cpu.go:19 / prover_gpu.go:66."""
    return _prover(curve, ccs, pk, assignment, *opts)


def ProveWithWitness(curve, ccs, pk, w, *opts):
    """prover/prover_cpu.go:49 / prover_gpu.go:111."""
    if not UseGPUProver:
        return CPUProverWithWitness(curve, ccs, pk, w, *opts)
    return GPUProverWithWitness(curve, ccs, pk, w, *opts)
