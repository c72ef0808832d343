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


# ----------------------------------------------------------------------------- pairing-based verification
# The reference verifies openings in-circuit against the ceremony's [tau]_2 (/root/reference/crypto/blobs/kzg.go:26-45,
# evaluation.go) and geth verifies them natively (kzg4844.VerifyProof / VerifyCellProofBatch); both are the check
#     e(C - [y]_1, G_2) = e(pi, [tau]_2 - [z]_2)          (cells: e(C - [I_k(tau)]_1, G_2) = e(pi_k, [tau^64]_2 - [a_k]_2))
# against reference-held material only - this pins every opening / cell proof the GPU or this oracle produces.
def g2_decompress(b: bytes):
    """96-byte compressed BLS12-381 G2 point (ZCash / gnark-crypto encoding: x.c1 first, flags in the top 3 bits)."""
    cx = C.ctx("bls12_381")
    p = cx.p
    assert len(b) == 96 and b[0] & FLAG_COMPRESSED
    if b[0] & FLAG_INFINITY:
        return None
    x1 = int.from_bytes(bytes([b[0] & 0x1F]) + b[1:48], "big")
    x0 = int.from_bytes(b[48:], "big")
    F2 = cx.F2
    x = (x0, x1)
    y = F2.sqrt(F2.add(F2.mul(F2.sqr(x), x), cx.G2.b))
    assert y is not None, "x not on the twist"
    largest = (y[1] > (p - 1) // 2) if y[1] else (y[0] > (p - 1) // 2)     # lexicographic: c1 first
    if largest != bool(b[0] & FLAG_LARGEST):
        y = F2.neg(y)
    return (x, y)


def verify_kzg_proof(commitment: bytes, z: int, y: int, proof: bytes, tau_g2, g2_gen=None) -> bool:
    """e(C - [y]_1, G_2) * e(-pi, [tau]_2 - [z]_2) == 1  (EIP-4844 verify_kzg_proof_impl)."""
    from . import pairing
    pr = pairing.get("bls12_381")
    cx = pr.cx
    r = cx.r
    G1, G2 = cx.G1, cx.G2
    g2 = g2_gen or cx_g2_generator()
    Cm, Pi = g1_decompress(commitment), g1_decompress(proof)
    lhs = G1.add(Cm, G1.neg(G1.mul(cx.g1, y % r)))
    rhs = G2.add(tau_g2, G2.neg(G2.mul(g2, z % r)))
    return pr.product_is_one([(lhs, g2), (G1.neg(Pi), rhs)])


def cx_g2_generator():
    """The ceremony's G_2 = the standard BLS12-381 G2 generator (kzg.go:33-38; equals [tau^0]_2 of the SRS file)."""
    return g2_decompress(bytes.fromhex(
        "93e02b6052719f607dacd3a088274f65596bd0d09920b61ab5da61bbdc7f5049334cf11213945d57e5ac7d055d042b7e"
        "024aa2b2f08f0a91260805272dc51051c6e47ad4fa403b02b4510b647ae3d1770bac0326a805bbefd48056c8c121bdb8"))


def extended_cell_values(blob: bytes, k: int):
    """The 64 evaluations that make cell k of the extended blob (cells < 64 are the blob's own chunks)."""
    r = P.BLS12_381.r
    vals = blob_scalars(blob)
    m = FIELD_ELEMENTS_PER_CELL
    if k < len(vals) // m:
        return vals[m * k:m * (k + 1)]
    coeffs = blob_coefficients(vals)
    n = len(vals)
    logext = (2 * n).bit_length() - 1
    w = pow(PRIMITIVE_ROOT_2_32, 1 << (32 - logext), r)
    out = []
    for j in range(m):
        x = pow(w, N.bitrev(m * k + j, logext), r)
        acc = 0
        for cf in reversed(coeffs):
            acc = (acc * x + cf) % r
        out.append(acc)
    return out


def cell_interpolant(cell_vals, k, n=4096):
    """Coefficients (degree < 64) of the polynomial through cell k's points x_j = h_k zeta^brp6(j)."""
    r = P.BLS12_381.r
    m = FIELD_ELEMENTS_PER_CELL
    logm = m.bit_length() - 1
    h = cell_coset_shift(k, n)
    zeta = pow(PRIMITIVE_ROOT_2_32, 1 << (32 - logm), r)
    zinv = pow(zeta, -1, r)
    minv = pow(m, -1, r)
    hinv = pow(h, -1, r)
    coeffs = []
    for i in range(m):
        acc = 0
        for j, v in enumerate(cell_vals):
            acc += v * pow(zinv, i * N.bitrev(j, logm) % m, r)
        coeffs.append(acc % r * minv % r * pow(hinv, i, r) % r)
    return coeffs


def verify_cell_proof(commitment: bytes, k: int, cell_vals, proof: bytes, mono_g1_64, tau64_g2, n=4096) -> bool:
    """e(C - [I_k(tau)]_1, G_2) * e(-pi_k, [tau^64]_2 - [h_k^64]_2) == 1  (EIP-7594 verify_cell_kzg_proof)."""
    from . import pairing
    pr = pairing.get("bls12_381")
    cx = pr.cx
    r = cx.r
    G1, G2 = cx.G1, cx.G2
    g2 = cx_g2_generator()
    coeffs = cell_interpolant(cell_vals, k, n)
    interp = G1.msm_naive(mono_g1_64, coeffs)
    Cm, Pi = g1_decompress(commitment), g1_decompress(proof)
    lhs = G1.add(Cm, G1.neg(interp))
    a = pow(cell_coset_shift(k, n), FIELD_ELEMENTS_PER_CELL, r)
    rhs = G2.add(tau64_g2, G2.neg(G2.mul(g2, a)))
    return pr.product_is_one([(lhs, g2), (G1.neg(Pi), rhs)])
