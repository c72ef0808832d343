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
