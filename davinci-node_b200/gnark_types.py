"""Plain-container mirrors of the gnark types that cross the prover boundary
(`constraint.ConstraintSystem`, `groth16.ProvingKey`, `witness.Witness`, `groth16.Proof`;
/root/reference/prover/config.go:34-50).  Buffers are numpy uint8 arrays in gnark-crypto's memory
layout, exactly what the Go shim hands to the C ABI from gnark's own structs.

Only the data the proving path reads is modelled.  `ConstraintSystem.solve` is the host-side R1CS
solver (gnark's `r1cs.Solve`, SURVEY.md A.1 step 4 - it stays on the CPU in the reference too and is
unchanged gnark code in a real deployment); here it handles the synthetic single-output-wire circuits
the tests and the bench use.
"""
from dataclasses import dataclass, field

import numpy as np

from . import hash_to_field as H2F
from .layout import Layout


class UnsatisfiedConstraintError(ValueError):
    pass


@dataclass
class Witness:
    """witness.Witness: public (without the constant one) then secret values, as Python ints."""
    curve_id: int
    public: list
    secret: list

    @staticmethod
    def from_assignment(assignment, curve_id):
        """frontend.NewWitness(assignment, field): `assignment` is a mapping with 'public' / 'secret'."""
        return Witness(curve_id, list(assignment["public"]), list(assignment["secret"]))


@dataclass
class Solution:
    W: np.ndarray
    A: np.ndarray
    B: np.ndarray
    C: np.ndarray
    private_committed: list
    commitments: list
    fold_challenge: int = 0
    values: list = None          # wire values as ints (kept for tests)


@dataclass
class ConstraintSystem:
    """R1CS: constraint k is <L_k,w> * <R_k,w> = <O_k,w>; wires = [one, public.., secret.., internal..]."""
    curve_id: int
    nb_wires: int
    nb_public: int               # includes the constant-one wire (GetNbPublicVariables)
    nb_secret: int
    L: list                      # per constraint list of (wire, coeff)
    R: list
    O: list
    commitments: list = field(default_factory=list)   # {'private_committed': [...], 'commitment_index': w}

    @property
    def nb_constraints(self):
        return len(self.L)

    def krs_skip_wires(self):
        s = set()
        for cm in self.commitments:
            s.update(cm["private_committed"])
            s.add(cm["commitment_index"])
        return s

    def commitment_challenge(self, i, commitment_bytes, L: Layout, vals, hash_kind="default"):
        """Value of commitment wire i: gnark's `opt.HashToFieldFn` over Commitment.Marshal() || the public committed
        values.  hash_kind 'default' = hash_to_field "bsb22-commitment" (RFC 9380 xmd),
        'solidity' = keccak256 (solidity.WithProverTargetSolidityVerifier)."""
        pt = L.dec_affine(commitment_bytes, 1)[0]
        pub = [vals[wire] for wire in self.commitments[i].get("public_committed", [])]
        return H2F.commitment_challenge(hash_kind, pt, pub, L.r, L.fp_bytes)

    def solve(self, w: Witness, commit_hint=None, hash_kind="default") -> Solution:
        L = Layout(self.curve_id)
        q = L.r
        if len(w.public) != self.nb_public - 1 or len(w.secret) != self.nb_secret:
            raise ValueError("witness size mismatch: want %d public / %d secret values"
                             % (self.nb_public - 1, self.nb_secret))
        vals = [None] * self.nb_wires
        vals[0] = 1
        for i, v in enumerate(w.public):
            vals[1 + i] = int(v) % q
        for i, v in enumerate(w.secret):
            vals[self.nb_public + i] = int(v) % q
        priv_committed, commitments = [], []
        for i, cm in enumerate(self.commitments):
            cv = L.enc_fr([vals[x] for x in cm["private_committed"]])
            priv_committed.append(cv)
            if commit_hint is None:
                raise ValueError("circuit has commitments but no commitment hint was supplied")
            cbytes = commit_hint(i, cv)
            commitments.append(cbytes)
            vals[cm["commitment_index"]] = self.commitment_challenge(i, cbytes, L, vals, hash_kind)

        def ev(terms):
            acc = 0
            for wire, cf in terms:
                v = vals[wire]
                if v is None:
                    raise UnsatisfiedConstraintError("wire %d used before it is solved" % wire)
                acc += v * cf
            return acc % q

        a, b, c = [], [], []
        for k in range(self.nb_constraints):
            x, y = ev(self.L[k]), ev(self.R[k])
            out = self.O[k]
            if len(out) == 1 and vals[out[0][0]] is None:
                wire, cf = out[0]
                vals[wire] = x * y % q * pow(cf, -1, q) % q
            z = ev(out)
            if (x * y - z) % q:
                raise UnsatisfiedConstraintError("constraint #%d is not satisfied" % k)
            a.append(x)
            b.append(y)
            c.append(z)
        fold = 0
        if len(self.commitments) > 1:
            # pedersen.BatchProve: Fiat-Shamir over the commitment wires' values
            fold = H2F.fold_challenge([vals[cm["commitment_index"]] for cm in self.commitments], q)
        return Solution(L.enc_fr(vals), L.enc_fr(a), L.enc_fr(b), L.enc_fr(c), priv_committed, commitments, fold, vals)


@dataclass
class ProvingKey:
    """groth16_<curve>.ProvingKey exported fields."""
    curve_id: int
    domain_cardinality: int
    domain_generator: np.ndarray
    domain_coset_gen: np.ndarray
    g1_alpha: np.ndarray
    g1_beta: np.ndarray
    g1_delta: np.ndarray
    g1_A: np.ndarray
    g1_B: np.ndarray
    g1_Z: np.ndarray
    g1_K: np.ndarray
    g2_beta: np.ndarray
    g2_delta: np.ndarray
    g2_B: np.ndarray
    infinity_a: np.ndarray
    infinity_b: np.ndarray
    commitment_keys: list = field(default_factory=list)   # [{'Basis': bytes, 'BasisExpSigma': bytes}]


class Proof:
    """groth16_<curve>.Proof: Ar, Krs (G1Affine), Bs (G2Affine), Commitments, CommitmentPok."""

    def __init__(self, curve_id):
        self.curve_id = curve_id
        self.Ar = self.Krs = self.Bs = self.CommitmentPok = None
        self.Commitments = []

    def points(self):
        L = Layout(self.curve_id)
        return {
            "Ar": L.dec_affine(self.Ar, 1)[0],
            "Krs": L.dec_affine(self.Krs, 1)[0],
            "Bs": L.dec_affine(self.Bs, 2)[0],
            "Commitments": [L.dec_affine(c, 1)[0] for c in self.Commitments],
            "CommitmentPok": L.dec_affine(self.CommitmentPok, 1)[0] if self.Commitments else None,
        }
