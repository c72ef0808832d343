"""ctypes loader for the C++ CPU restatement (oracle/c/oracle.cpp).  TEST INFRASTRUCTURE ONLY:
imported by tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs."""
import ctypes as C
import os
import subprocess

_HERE = os.path.dirname(os.path.abspath(__file__))
SRC = os.path.join(_HERE, "c", "oracle.cpp")
LIB = os.path.join(_HERE, "_build", "liboracle_c.so")


def build(force=False):
    os.makedirs(os.path.dirname(LIB), exist_ok=True)
    if force or not os.path.exists(LIB) or os.path.getmtime(LIB) < os.path.getmtime(SRC):
        cmd = ["g++", "-O3", "-fopenmp", "-std=c++17", "-shared", "-fPIC", "-o", LIB, SRC]
        subprocess.run(cmd, check=True)
    return LIB


class ProveArgs(C.Structure):
    _fields_ = [
        ("curve", C.c_int), ("logn", C.c_int), ("omega", C.c_void_p), ("g", C.c_void_p),
        ("A_ext", C.c_void_p), ("B1_ext", C.c_void_p), ("B2_ext", C.c_void_p), ("K_ext", C.c_void_p), ("Z", C.c_void_p),
        ("mapA", C.c_void_p), ("mapB", C.c_void_p), ("mapK", C.c_void_p),
        ("m", C.c_uint64), ("nb_public", C.c_uint64), ("nZ", C.c_uint64),
        ("W_ext", C.c_void_p), ("a", C.c_void_p), ("b", C.c_void_p), ("c", C.c_void_p),
        ("out_ar", C.c_void_p), ("out_bs", C.c_void_p), ("out_krs", C.c_void_p), ("threads", C.c_int),
    ]


_lib = None


def lib():
    global _lib
    if _lib is None:
        build()
        _lib = C.CDLL(LIB)
        vp, u64, i = C.c_void_p, C.c_uint64, C.c_int
        _lib.oc_msm.argtypes = [i, i, vp, vp, u64, vp, vp, i]
        _lib.oc_fft.argtypes = [i, vp, i, vp, vp, i, i, i, i]
        _lib.oc_compute_h.argtypes = [i, vp, vp, vp, i, vp, vp, i]
        _lib.oc_prove.argtypes = [C.POINTER(ProveArgs)]
        _lib.oc_fr_mul.argtypes = [i, vp, vp, vp, u64]
        _lib.oc_num_threads.restype = i
    return _lib


def p(arr):
    return arr.ctypes.data if arr is not None else None
