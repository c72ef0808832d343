"""Workload definitions shared by BOTH arms of bench.py (numpy only - imports neither the product package nor the
oracle, so `bench.py --impl reference` never maps libb200groth16.so).

The real circuits' ccs / pk are CDN artifacts (sizes are engineering estimates), so every config is a
*shape*: curve, domain size, public wires, infinity fractions of the A / B keys, one BSB22 commitment, and the solved
witness' value mix.  The B200 arm builds a structured key of that shape on the GPU
(davinci_node_b200.synthetic.SyntheticWorkload draws the same masks from the same seed); the CPU arm builds one with
oracle/c's point generator.
"""
import numpy as np

# (id, p, r, fp limbs64, fr limbs64, G2 extension degree, fr two-adicity, 2^s-th root of unity, gnark coset generator)
CURVES = {
    "bn254": (1,
              0x30644e72e131a029b85045b68181585d97816a916871ca8d3c208c16d87cfd47,
              0x30644e72e131a029b85045b68181585d2833e84879b9709143e1f593f0000001, 4, 4, 2, 28,
              19103219067921713944291392827692070036145651957329286315305642004821462161904, 5),
    "bls12_377": (2,
                  0x1ae3a4617c510eac63b05c06ca1493b1a22d9f300f5138f1ef3622fba094800170b5d44300000008508c00000000001,
                  0x12ab655e9a2ca55660b44d1e5c37b00159aa76fed00000010a11800000000001, 6, 4, 2, 47,
                  8065159656716812877374967518403273466521432693661810619979959746626482506078, 22),
    "bw6_761": (4,
                0x122e824fb83ce0ad187c94004faff3eb926186a81d14688528275ef8087be41707ba638e584e91903cebaff25b423048689c8ed12f9fd9071dcd3dc73ebff2e98a116c25667a8f8160cf8aeeaf0a437e6913e6870000082f49d00000000008b,
                0x1ae3a4617c510eac63b05c06ca1493b1a22d9f300f5138f1ef3622fba094800170b5d44300000008508c00000000001, 12, 6, 1, 46,
                32863578547254505029601261939868325669770508939375122462904745766352256812585773382134936404344547323199885654433, 15),
}

# BASELINE.json configs -> shapes (SURVEY.md 8d "concrete synthetic inputs per config")
CONFIGS = {
    # configs[0] / [1]: voteverifier, BLS12-377, 6 public wires (1 + IsValid + 4 limbs of BallotHash)
    "voteverifier": dict(curve="bls12_377", logn=22, nb_public=6, seed=0xD0A1,
                         title="voteverifier-shaped Groth16 proof"),
    # configs[2]: aggregator, BW6-761 (2 public wires), MSMs range-split across GPUs
    "aggregator": dict(curve="bw6_761", logn=22, nb_public=2, seed=0xA66,
                       title="aggregator-shaped Groth16 proof"),
    # configs[3]: statetransition, BN254, 9 public inputs + the constant wire, followed by the blob KZG work
    "statetransition": dict(curve="bn254", logn=24, nb_public=10, seed=0x57A7E,
                            title="statetransition-shaped Groth16 proof + EIP-4844 blob commitment"),
}

METRICS = {
    "voteverifier": ("voteverifier Groth16 proofs/s", "proofs/s"),
    "aggregator": ("aggregator Groth16 proofs/s (BW6-761)", "proofs/s"),
    "statetransition": ("statetransition Groth16 proofs/s (BN254, incl. blob KZG commitment)", "proofs/s"),
    "blob": ("EIP-4844 blob KZG commitments/s", "commitments/s"),
}


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
    u = rng.random(n)
    zero = u < 0.40
    one = (u >= 0.40) & (u < 0.60)
    small = (u >= 0.60) & (u < 0.85)
    a[zero] = 0
    a[one] = 0
    a[one, 0] = 1
    a[small, 1:] = 0
    return a


def mix_stats(canon):
    """(zero, one, < 2^64 but > 1, full-width) counts of a canonical scalar array - what the sparsity-aware
    work formula needs."""
    hi = (canon[:, 1:] != 0).any(axis=1)
    lo = canon[:, 0]
    zero = int((~hi & (lo == 0)).sum())
    one = int((~hi & (lo == 1)).sum())
    small = int((~hi & (lo > 1)).sum())
    return zero, one, small, int(hi.sum())


class Shape:
    """Index structure of a circuit-shaped proving key: which wires have A / B points, which are committed.
    Draws the infinity masks exactly as davinci_node_b200.synthetic.SyntheticWorkload does (same seed, same order)."""

    def __init__(self, curve, logn, seed, nb_public=6, n_commit_log=None, inf_a=0.30, inf_b=0.40, nb_constraints=None):
        self.curve = curve
        self.cid, self.p, self.r, self.fp_l, self.fr_l, self.g2_deg, _, _, _ = CURVES[curve]
        self.logn, self.n = logn, 1 << logn
        self.m = self.n
        self.nc = nb_constraints if nb_constraints is not None else self.n - 3
        self.nb_public = nb_public
        n_c = 1 << (n_commit_log if n_commit_log is not None else max(1, logn - 4))
        self.n_c = min(n_c, self.m - nb_public - 2)
        rng = np.random.default_rng(seed)
        m = self.m
        self.infA = rng.random(m) < inf_a
        self.infB = rng.random(m) < inf_b
        self.infA[:nb_public] = False
        self.committed = np.arange(nb_public, nb_public + self.n_c, dtype=np.uint32)
        self.commit_wire = nb_public + self.n_c
        self.krs_skip = np.concatenate([self.committed, np.array([self.commit_wire], dtype=np.uint32)])
        self.nA, self.nB = int((~self.infA).sum()), int((~self.infB).sum())
        self.nK = m - nb_public - len(self.krs_skip)
        self.nZ = self.n - 1

    def maps(self):
        """(mapA, mapB over W_ext = [w.., r, s, 1, -rs]; mapK over W_ext[nb_public:]) in the extended-array
        convention of prover.cu / oracle.cpp: 0xffffffff = skip, tail slots route to delta / alpha / beta."""
        SKIP = 0xFFFFFFFF
        m = self.m
        mapA = np.full(m + 4, SKIP, dtype=np.uint32)
        mapB = np.full(m + 4, SKIP, dtype=np.uint32)
        mapA[:m][~self.infA] = np.arange(self.nA, dtype=np.uint32)
        mapB[:m][~self.infB] = np.arange(self.nB, dtype=np.uint32)
        mapA[m], mapA[m + 2] = self.nA, self.nA + 1
        mapB[m + 1], mapB[m + 2] = self.nB, self.nB + 1
        npriv = m - self.nb_public
        mapK = np.full(npriv + 4, SKIP, dtype=np.uint32)
        keep = np.ones(npriv, dtype=bool)
        keep[self.krs_skip - self.nb_public] = False
        mapK[:npriv][keep] = np.arange(self.nK, dtype=np.uint32)
        mapK[npriv + 3] = self.nK
        return mapA, mapB, mapK

    def fr_bits(self):
        return self.r.bit_length()

    def h2d_bytes(self):
        return (self.m + 3 * self.nc + 2) * 8 * self.fr_l


def enc_mont(vals, modulus, limbs64):
    """ints -> gnark-crypto Montgomery limbs, (len, limbs64) uint64."""
    R = 1 << (64 * limbs64)
    out = np.zeros((len(vals), limbs64), dtype=np.uint64)
    for i, v in enumerate(vals):
        v = (int(v) % modulus) * R % modulus
        for j in range(limbs64):
            out[i, j] = (v >> (64 * j)) & 0xFFFFFFFFFFFFFFFF
    return out


def domain_constants(curve, logn):
    """(omega, coset generator) of gnark's fft.Domain of size 2^logn (canonical ints)."""
    _, _, r, _, _, _, s, root, gen = CURVES[curve]
    return pow(root, 1 << (s - logn), r), gen


# ------------------------------------------------------------------------------------ work formulas
def p_mul(n32):
    """32x32->64 multiply-accumulates of one Montgomery multiplication with n32 limbs."""
    return 2 * n32 * n32 + n32


def msm_adds_star(n, bits):
    """min over c in [4,24] of n*W(c) + 2*W(c)*2^(c-1), W(c) = ceil((b+1)/c): dense scalars."""
    best = None
    for c in range(4, 25):
        w = -(-(bits + 1) // c)
        adds = n * w + 2 * w * (1 << (c - 1))
        best = adds if best is None else min(best, adds)
    return best


def msm_adds_sparse(zero, one, small, full, bits):
    """Sparsity-aware counterpart for a solved witness: zeros cost nothing, ones one addition each (a plain sum),
    values < 2^64 fill ceil(65 / c) windows, full-width values W(c); same min over c, same bucket-reduction term."""
    best = None
    for c in range(4, 25):
        w = -(-(bits + 1) // c)
        ws = -(-65 // c)
        adds = full * w + small * ws + one + 2 * w * (1 << (c - 1))
        best = adds if best is None else min(best, adds)
    return best
