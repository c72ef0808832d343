"""Mirror of the reference's blob-commitment boundary: `types.Blob.ComputeCommitment`
(/root/reference/types/blobs.go:90-96 -> gethkzg.BlobToCommitment), on the GPU.

The EIP-4844 SRS (4096 G1 Lagrange points, 48-byte compressed, the order of
/root/reference/config/kzg_trusted_setup.txt) is registered once, decompressed on the device and
kept resident; each commitment is one 4096-point BLS12-381 MSM."""
import ctypes as C
import threading

import numpy as np

from . import capi

BLOB_BYTES = 4096 * 32
_lock = threading.Lock()
_srs_handle = None


class BlobError(ValueError):
    pass


def load_trusted_setup(g1_lagrange_compressed: bytes):
    """g1_lagrange_compressed: 4096 * 48 bytes.  (Go: parsed from config.KZGTrustedSetup.)"""
    global _srs_handle
    with _lock:
        capi.init_once()
        buf = np.frombuffer(g1_lagrange_compressed, dtype=np.uint8).copy()
        if len(buf) % 48:
            raise BlobError("SRS must be a whole number of 48-byte points")
        h = C.c_uint64(0)
        capi.check(capi.lib.b200_kzg_srs_register(buf.ctypes.data, len(buf) // 48, C.byref(h)))
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
