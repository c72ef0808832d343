"""EIP-4844 blob KZG commitment oracle (big-int).  TEST INFRASTRUCTURE ONLY.

Restates `kzg4844.BlobToCommitment` (go-ethereum v1.17.1 -> go-eth-kzg v1.5.0 / c-kzg-4844 v2.1.6,
go.mod:17,110,132; called at /root/reference/types/blobs.go:90-96): the blob is 4096 big-endian
32-byte canonical BLS12-381 scalars, cell i being the evaluation at omega^brp(i)
(/root/reference/crypto/blobs/omega.go, barycentric.go:47-73); with the SRS Lagrange points in the
natural order of /root/reference/config/kzg_trusted_setup.txt the commitment is
sum_i blob[i] * lag[brp(i)], serialised as a 48-byte compressed G1 point.

Pinning: the SRS is the real ceremony output (fixture tests/golden/kzg_g1_lagrange.bin, made by
tools/make_golden.py); sum of all Lagrange points == G1 generator (all-ones blob), and the
Lagrange-basis result equals the monomial-basis route.  No geth-produced
commitment bytes exist in the reference, so byte parity with geth is "unpinned" beyond those checks.
"""
from . import curve as C
from . import ntt as N
from . import params as P

FLAG_COMPRESSED, FLAG_INFINITY, FLAG_LARGEST = 0x80, 0x40, 0x20


def g1_decompress(b: bytes):
    cx = C.ctx("bls12_381")
    p = cx.p
    assert len(b) == 48 and b[0] & FLAG_COMPRESSED
    if b[0] & FLAG_INFINITY:
        return None
    x = int.from_bytes(bytes([b[0] & 0x1F]) + b[1:], "big")
    y = C.sqrt_mod((x * x * x + 4) % p, p)
    assert y is not None, "x not on curve"
    if (y > (p - 1) // 2) != bool(b[0] & FLAG_LARGEST):
        y = p - y
    return (x, y)


def g1_compress(pt) -> bytes:
    cx = C.ctx("bls12_381")
    if pt is None:
        return bytes([0xC0]) + bytes(47)
    x, y = pt
    b = bytearray(x.to_bytes(48, "big"))
    b[0] |= FLAG_COMPRESSED
    if y > (cx.p - 1) // 2:
        b[0] |= FLAG_LARGEST
    return bytes(b)


def blob_scalars(blob: bytes):
    assert len(blob) % 32 == 0
    vals = [int.from_bytes(blob[i:i + 32], "big") for i in range(0, len(blob), 32)]
    if any(v >= P.BLS12_381.r for v in vals):
        raise ValueError("non-canonical field element in blob")
    return vals


def blob_to_commitment(blob: bytes, lagrange_points) -> bytes:
    """lagrange_points: affine points in SRS-file (natural) order."""
    cx = C.ctx("bls12_381")
    vals = blob_scalars(blob)
    n = len(vals)
    logn = n.bit_length() - 1
    pts = [lagrange_points[N.bitrev(i, logn)] for i in range(n)]
    return g1_compress(cx.G1.msm(pts, vals))


# ----------------------------------------------------------------------------- opening proofs
# Restates go-ethereum `kzg4844.ComputeProof` / `ComputeBlobProof` (called at
# /root/reference/types/blobs.go:111-134) = EIP-4844 `compute_kzg_proof_impl` / `compute_blob_kzg_proof`
# (consensus-specs deneb/polynomial-commitments.md): the blob is the polynomial's evaluations over the
# 4096th roots of unity in bit-reversed order; the proof commits to q(X) = (p(X) - p(z)) / (X - z) in the same
# evaluation basis.
PRIMITIVE_ROOT_2_32 = 10238227357739495823651030575849232062558860180284477541189508159991286009131  # barycentric.go:52
FIAT_SHAMIR_PROTOCOL_DOMAIN = b"FSBLOBVERIFY_V1_"


def roots_of_unity_brp(n):
    r = P.BLS12_381.r
    logn = n.bit_length() - 1
    w = pow(PRIMITIVE_ROOT_2_32, 1 << (32 - logn), r)
    nat = [1] * n
    for i in range(1, n):
        nat[i] = nat[i - 1] * w % r
    return [nat[N.bitrev(i, logn)] for i in range(n)]


def evaluate_in_evaluation_form(poly, z, roots):
    r = P.BLS12_381.r
    n = len(poly)
    if z in roots:
        return poly[roots.index(z)]
    acc = 0
    for pi, wi in zip(poly, roots):
        acc += pi * wi % r * pow((z - wi) % r, -1, r)
    return acc % r * ((pow(z, n, r) - 1) % r) % r * pow(n, -1, r) % r


def quotient_evaluations(poly, z, roots):
    """(q evaluations, y): q_i = (p_i - y) / (w_i - z); the in-domain case follows the spec's
    compute_quotient_eval_within_domain."""
    r = P.BLS12_381.r
    y = evaluate_in_evaluation_form(poly, z, roots)
    q = [0] * len(poly)
    for i, (pi, wi) in enumerate(zip(poly, roots)):
        if wi == z:
            acc = 0
            for pj, wj in zip(poly, roots):
                if wj == z:
                    continue
                acc += (pj - y) % r * wj % r * pow(z * (z - wj) % r, -1, r)
            q[i] = acc % r
        else:
            q[i] = (pi - y) % r * pow((wi - z) % r, -1, r) % r
    return q, y


def compute_proof(blob: bytes, z: int, lagrange_points):
    """-> (48-byte proof, claimed value y as int).  lagrange_points in SRS-file (natural) order."""
    cx = C.ctx("bls12_381")
    vals = blob_scalars(blob)
    n = len(vals)
    logn = n.bit_length() - 1
    roots = roots_of_unity_brp(n)
    q, y = quotient_evaluations(vals, z % P.BLS12_381.r, roots)
    pts = [lagrange_points[N.bitrev(i, logn)] for i in range(n)]
    return g1_compress(cx.G1.msm(pts, q)), y


def compute_challenge(blob: bytes, commitment: bytes) -> int:
    import hashlib
    n = len(blob) // 32
    data = FIAT_SHAMIR_PROTOCOL_DOMAIN + n.to_bytes(16, "big") + blob + commitment
    return int.from_bytes(hashlib.sha256(data).digest(), "big") % P.BLS12_381.r


def compute_blob_proof(blob: bytes, commitment: bytes, lagrange_points) -> bytes:
    return compute_proof(blob, compute_challenge(blob, commitment), lagrange_points)[0]


# ----------------------------------------------------------------------------- cell proofs (EIP-7594)
# Restates go-ethereum `kzg4844.ComputeCellProofs` (called at /root/reference/types/blobs.go:99-105) =
# `compute_cells_and_kzg_proofs` of consensus-specs fulu/polynomial-commitments-sampling.md: the blob polynomial in
# coefficient form is divided by the vanishing polynomial X^64 - h_k^64 of each of the 128 cosets of the extended
# (8192-point) domain; proof_k commits to the quotient in the MONOMIAL basis of the ceremony.
CELLS_PER_EXT_BLOB = 128
FIELD_ELEMENTS_PER_CELL = 64


def blob_coefficients(vals):
    """evaluations over the 4096th roots of unity in bit-reversed order -> monomial coefficients."""
    dom = N.Domain(P.BLS12_381, len(vals))
    return N.fft(list(vals), dom, inverse=True, dit=True)


def cell_coset_shift(k, n=4096):
    """h_k: first element of coset k = roots_of_unity_brp(2n)[64 k]."""
    r = P.BLS12_381.r
    ext = 2 * n
    logext = ext.bit_length() - 1
    w = pow(PRIMITIVE_ROOT_2_32, 1 << (32 - logext), r)
    return pow(w, N.bitrev(FIELD_ELEMENTS_PER_CELL * k, logext), r)


def cell_quotient(coeffs, k):
    """quotient of p(X) by X^64 - h_k^64 (remainder = the interpolant of the cell, discarded)."""
    r = P.BLS12_381.r
    m = FIELD_ELEMENTS_PER_CELL
    a = pow(cell_coset_shift(k, len(coeffs)), m, r)
    n = len(coeffs)
    q = [0] * (n - m)
    for j in range(n - m - 1, -1, -1):
        q[j] = (coeffs[j + m] + (a * q[j + m] if j + m < n - m else 0)) % r
    return q


def compute_cell_proofs(blob: bytes, monomial_points, cells=None):
    """-> {cell index: 48-byte proof} for `cells` (default all 128); monomial_points = [tau^j]_1, j < 4096."""
    cx = C.ctx("bls12_381")
    coeffs = blob_coefficients(blob_scalars(blob))
    out = {}
    for k in (range(CELLS_PER_EXT_BLOB) if cells is None else cells):
        q = cell_quotient(coeffs, k)
        out[k] = g1_compress(cx.G1.msm(monomial_points[:len(q)], q))
    return out
