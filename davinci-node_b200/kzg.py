"""Mirror of the reference's blob-commitment boundary: `types.Blob.ComputeCommitment`
(/root/reference/types/blobs.go:90-96 -> gethkzg.BlobToCommitment), `ComputeProof` and `ComputeBlobProof`
(types/blobs.go:107-134), on the GPU.

The EIP-4844 SRS (4096 G1 Lagrange points, 48-byte compressed, the order of
/root/reference/config/kzg_trusted_setup.txt) is registered once, decompressed on the device and
kept resident; each commitment is one 4096-point BLS12-381 MSM."""
import ctypes as C
import threading

import numpy as np

from . import capi

BLOB_BYTES = 4096 * 32
BLS_MODULUS = 0x73eda753299d7d483339d80809a1d80553bda402fffe5bfeffffffff00000001
_lock = threading.Lock()
_srs_handle = None


class BlobError(ValueError):
    pass


def load_trusted_setup(g1_lagrange_compressed: bytes, g1_monomial_compressed: bytes = None):
    """g1_lagrange_compressed: 4096 * 48 bytes; g1_monomial_compressed (optional, same size): the monomial-basis
    points, needed by ComputeCellProofs.  (Go: parsed from config.KZGTrustedSetup.)"""
    global _srs_handle
    with _lock:
        capi.init_once()
        buf = np.frombuffer(g1_lagrange_compressed, dtype=np.uint8).copy()
        if len(buf) % 48:
            raise BlobError("SRS must be a whole number of 48-byte points")
        h = C.c_uint64(0)
        capi.check(capi.lib.b200_kzg_srs_register(buf.ctypes.data, len(buf) // 48, C.byref(h)))
        if g1_monomial_compressed is not None:
            mono = np.frombuffer(g1_monomial_compressed, dtype=np.uint8).copy()
            capi.check(capi.lib.b200_kzg_srs_add_monomial(h.value, mono.ctypes.data, len(mono) // 48))
        if _srs_handle is not None:
            capi.check(capi.lib.b200_kzg_srs_release(_srs_handle))
        _srs_handle = h.value
        return _srs_handle


class Blob:
    """types.Blob: 131072 bytes = 4096 big-endian 32-byte field elements."""

    def __init__(self, data: bytes):
        if len(data) != BLOB_BYTES:
            raise BlobError("blob must be %d bytes" % BLOB_BYTES)
        self.data = bytes(data)

    def ComputeCommitment(self, device=-1) -> bytes:
        """48-byte compressed commitment; raises BlobError on a non-canonical field element
        (gethkzg returns an error there)."""
        if _srs_handle is None:
            raise BlobError("trusted setup not loaded (call load_trusted_setup first)")
        blob = np.frombuffer(self.data, dtype=np.uint8)
        out = np.zeros(48, dtype=np.uint8)
        try:
            capi.check(capi.lib.b200_blob_commit(_srs_handle, blob.ctypes.data, out.ctypes.data, device))
        except capi.B200Error as e:
            raise BlobError(str(e)) from e
        return bytes(out)

    def ComputeProof(self, point: int, device=-1):
        """types/blobs.go:123: KZG proof at `point` for the blob polynomial -> (48-byte proof, claim y as int).
        Raises BlobError when |point| does not fit 32 bytes or is not a canonical field element."""
        if _srs_handle is None:
            raise BlobError("trusted setup not loaded (call load_trusted_setup first)")
        if point < 0 or point.bit_length() > 256:
            raise BlobError("point does not fit in 32 bytes")
        blob = np.frombuffer(self.data, dtype=np.uint8)
        z = np.frombuffer(point.to_bytes(32, "big"), dtype=np.uint8)
        proof = np.zeros(48, dtype=np.uint8)
        claim = np.zeros(32, dtype=np.uint8)
        try:
            capi.check(capi.lib.b200_blob_proof(_srs_handle, blob.ctypes.data, z.ctypes.data, proof.ctypes.data,
                                                claim.ctypes.data, device))
        except capi.B200Error as e:
            raise BlobError(str(e)) from e
        return bytes(proof), int.from_bytes(bytes(claim), "big")

    def ComputeBlobProof(self, commitment: bytes, device=-1) -> bytes:
        """types/blobs.go:111: the proof that verifies the blob against `commitment` (opening at the
        Fiat-Shamir challenge; the SHA-256 transcript hash is host work, as in geth)."""
        import hashlib
        if len(commitment) != 48:
            raise BlobError("commitment must be 48 bytes")
        data = b"FSBLOBVERIFY_V1_" + (BLOB_BYTES // 32).to_bytes(16, "big") + self.data + bytes(commitment)
        z = int.from_bytes(hashlib.sha256(data).digest(), "big") % BLS_MODULUS
        return self.ComputeProof(z, device)[0]

    def ComputeCellProofs(self, device=-1):
        """types/blobs.go:99: the 128 EIP-7594 cell proofs (48 bytes each)."""
        if _srs_handle is None:
            raise BlobError("trusted setup not loaded (call load_trusted_setup first)")
        blob = np.frombuffer(self.data, dtype=np.uint8)
        out = np.zeros(128 * 48, dtype=np.uint8)
        try:
            capi.check(capi.lib.b200_blob_cell_proofs(_srs_handle, blob.ctypes.data, out.ctypes.data, device))
        except capi.B200Error as e:
            raise BlobError(str(e)) from e
        return [bytes(out[48 * k:48 * (k + 1)]) for k in range(128)]

    def ComputeCommitmentAndProof(self, device=-1):
        """types/blobs.go:138-149: (commitment, blob proof) of a Version 0 BlobTxSidecar."""
        commitment = self.ComputeCommitment(device)
        return commitment, self.ComputeBlobProof(commitment, device)

    def ComputeCommitmentAndCellProofs(self, device=-1):
        """types/blobs.go:152-162: (commitment, 128 cell proofs) of a Version 1 BlobTxSidecar."""
        return self.ComputeCommitment(device), self.ComputeCellProofs(device)
