"""gnark artefact formats on the proving path.

  read_proving_key(data, curve_id)   `pk.UnsafeReadFrom` mirror (/root/reference/circuits/artifacts.go:391-406): parses
                                     the stream `pk.WriteTo` produced (/root/reference/cmd/circuit-compile/main.go:507-512)
                                     and decompresses every point on the GPU (b200_points_decompress_dev) - the reference
                                     spends the load time in ~10^7 CPU square roots.
  write_proving_key(pk)              `pk.WriteTo` mirror (compression on the GPU).

Format (restated from gnark v0.14 backend/groth16/<curve>/marshal.go and gnark-crypto's Encoder - the
source is not vendored, so the framing is "as surveyed", pinned only by the round trip with the oracle's independent
big-int serializer in tests/):
    fft.Domain   : u64 BE cardinality | 5 fr (BE canonical): cardinalityInv, generator, generatorInv, cosetGen,
                   cosetGenInv | 1 byte withPrecompute
    G1           : Alpha, Beta, Delta (compressed) | A, B, Z, K  (u32 BE count + compressed points each)
    G2           : Beta, Delta | B
    u64 BE nbWires | u64 NbInfinityA | u64 NbInfinityB | InfinityA, InfinityB (u32 BE count + one byte per bool)
    u32 BE number of commitment keys | per key: Basis, BasisExpSigma (u32 BE count + compressed points)
"""
import ctypes as C
import struct

import numpy as np

from . import capi
from .gnark_types import ProvingKey
from .layout import Layout


class ArtifactError(ValueError):
    pass


def _torch():
    import torch
    return torch


def compressed_bytes(L: Layout, group):
    return L.coord_width(group) * L.fp_bytes


def decompress_points(L: Layout, group, raw: bytes, count: int) -> np.ndarray:
    """count compressed points -> affine points in gnark memory layout (uint8 array), on the GPU."""
    torch = _torch()
    capi.init_once()
    cb = compressed_bytes(L, group)
    if len(raw) != count * cb:
        raise ArtifactError("truncated point block")
    if count == 0:
        return np.zeros(0, dtype=np.uint8)
    d_in = torch.frombuffer(bytearray(raw), dtype=torch.uint8).cuda()
    d_out = torch.empty(count * L.affine_bytes(group), dtype=torch.uint8, device="cuda")
    err = torch.zeros(1, dtype=torch.int32, device="cuda")
    st = torch.cuda.current_stream().cuda_stream
    capi.check(capi.lib.b200_points_decompress_dev(L.id, group, d_in.data_ptr(), count, d_out.data_ptr(), err.data_ptr(), st))
    flags = int(err.item())
    if flags & 1:
        raise ArtifactError("invalid point encoding (not a compressed point)")
    if flags & 2:
        raise ArtifactError("invalid point: x is not on the curve")
    return d_out.cpu().numpy()


def compress_points(L: Layout, group, affine: np.ndarray) -> bytes:
    torch = _torch()
    capi.init_once()
    count = len(affine) // L.affine_bytes(group)
    if count == 0:
        return b""
    d_in = torch.from_numpy(np.ascontiguousarray(affine)).cuda()
    d_out = torch.empty(count * compressed_bytes(L, group), dtype=torch.uint8, device="cuda")
    capi.check(capi.lib.b200_points_compress_dev(L.id, group, d_in.data_ptr(), count, d_out.data_ptr(),
                                                 torch.cuda.current_stream().cuda_stream))
    return d_out.cpu().numpy().tobytes()


class _Reader:
    def __init__(self, data):
        self.d, self.o = memoryview(data), 0

    def take(self, n):
        if self.o + n > len(self.d):
            raise ArtifactError("unexpected end of proving key")
        v = self.d[self.o:self.o + n]
        self.o += n
        return bytes(v)

    def u32(self):
        return struct.unpack(">I", self.take(4))[0]

    def u64(self):
        return struct.unpack(">Q", self.take(8))[0]


def read_proving_key(data: bytes, curve_id) -> ProvingKey:
    L = Layout(curve_id)
    rd = _Reader(data)
    card = rd.u64()
    if card == 0 or card & (card - 1):
        raise ArtifactError("domain cardinality is not a power of two")
    frs = [int.from_bytes(rd.take(L.fr_bytes), "big") for _ in range(5)]   # cardInv, gen, genInv, coset, cosetInv
    if any(v >= L.r for v in frs):
        raise ArtifactError("non-canonical field element in the domain")
    if pow(frs[1], card, L.r) != 1 or frs[0] * card % L.r != 1:
        raise ArtifactError("domain generator / cardinality mismatch")
    rd.take(1)                                                             # withPrecompute
    g1c, g2c = compressed_bytes(L, 1), compressed_bytes(L, 2)

    def point(group):
        return decompress_points(L, group, rd.take(g1c if group == 1 else g2c), 1)

    def points(group):
        n = rd.u32()
        return decompress_points(L, group, rd.take(n * (g1c if group == 1 else g2c)), n)

    alpha, beta, delta = point(1), point(1), point(1)
    A, B, Z, K = points(1), points(1), points(1), points(1)
    beta2, delta2 = point(2), point(2)
    B2 = points(2)
    nb_wires, nb_inf_a, nb_inf_b = rd.u64(), rd.u64(), rd.u64()

    def bools():
        n = rd.u32()
        v = np.frombuffer(rd.take(n), dtype=np.uint8).copy()
        if v.max(initial=0) > 1:
            raise ArtifactError("invalid bool")
        return v

    inf_a, inf_b = bools(), bools()
    if len(inf_a) != nb_wires or len(inf_b) != nb_wires or int(inf_a.sum()) != nb_inf_a or int(inf_b.sum()) != nb_inf_b:
        raise ArtifactError("infinity flags inconsistent with their counters")
    g1b = L.affine_bytes(1)
    if len(A) // g1b != nb_wires - nb_inf_a or len(B) // g1b != nb_wires - nb_inf_b or len(B2) // L.affine_bytes(2) != nb_wires - nb_inf_b:
        raise ArtifactError("A / B lengths inconsistent with the infinity flags")
    keys = []
    for _ in range(rd.u32()):
        basis = points(1)
        sigma = points(1)
        if len(basis) != len(sigma):
            raise ArtifactError("commitment key: Basis / BasisExpSigma length mismatch")
        keys.append({"Basis": basis, "BasisExpSigma": sigma})
    if rd.o != len(rd.d):
        raise ArtifactError("trailing bytes after the proving key")
    return ProvingKey(curve_id=L.id, domain_cardinality=card, domain_generator=L.enc_fr([frs[1]]),
                      domain_coset_gen=L.enc_fr([frs[3]]), g1_alpha=alpha, g1_beta=beta, g1_delta=delta, g1_A=A, g1_B=B,
                      g1_Z=Z, g1_K=K, g2_beta=beta2, g2_delta=delta2, g2_B=B2, infinity_a=inf_a, infinity_b=inf_b,
                      commitment_keys=keys)


def write_proving_key(pk: ProvingKey) -> bytes:
    L = Layout(pk.curve_id)
    out = bytearray()
    card = pk.domain_cardinality
    gen, coset = L.dec_fr(pk.domain_generator)[0], L.dec_fr(pk.domain_coset_gen)[0]
    out += struct.pack(">Q", card)
    for v in (pow(card, -1, L.r), gen, pow(gen, -1, L.r), coset, pow(coset, -1, L.r)):
        out += v.to_bytes(L.fr_bytes, "big")
    out += b"\x00"

    def points(group, arr, prefix=True):
        n = len(arr) // L.affine_bytes(group)
        return (struct.pack(">I", n) if prefix else b"") + compress_points(L, group, arr)

    for a in (pk.g1_alpha, pk.g1_beta, pk.g1_delta):
        out += points(1, a, False)
    for a in (pk.g1_A, pk.g1_B, pk.g1_Z, pk.g1_K):
        out += points(1, a)
    out += points(2, pk.g2_beta, False) + points(2, pk.g2_delta, False) + points(2, pk.g2_B)
    inf_a, inf_b = np.asarray(pk.infinity_a, dtype=np.uint8), np.asarray(pk.infinity_b, dtype=np.uint8)
    out += struct.pack(">QQQ", len(inf_a), int(inf_a.sum()), int(inf_b.sum()))
    out += struct.pack(">I", len(inf_a)) + inf_a.tobytes() + struct.pack(">I", len(inf_b)) + inf_b.tobytes()
    out += struct.pack(">I", len(pk.commitment_keys))
    for key in pk.commitment_keys:
        out += points(1, key["Basis"]) + points(1, key["BasisExpSigma"])
    return bytes(out)
