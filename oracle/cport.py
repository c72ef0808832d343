"""ctypes loader for the C++ CPU restatement (oracle/c/oracle.cpp).  TEST INFRASTRUCTURE ONLY:
imported by tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs."""
import ctypes as C
import os
import subprocess

_HERE = os.path.dirname(os.path.abspath(__file__))
SRC = os.path.join(_HERE, "c", "oracle.cpp")
LIB = os.path.join(_HERE, "_build", "liboracle_c.so")
STAMP = LIB + ".flags"


def _cpu_flags():
    try:
        with open("/proc/cpuinfo") as fh:
            for line in fh:
                if line.startswith("flags"):
                    return set(line.split(":", 1)[1].split())
    except OSError:
        pass
    return set()


def _arch_flags():
    """mulx / adx when THIS machine has them (the library is built in one container and may be loaded on another:
    the stamp file records the flags it was built with and lib() rebuilds when the host cannot run them)."""
    have = _cpu_flags()
    return [f for f, name in (("-mbmi2", "bmi2"), ("-madx", "adx")) if name in have]


def build(force=False):
    os.makedirs(os.path.dirname(LIB), exist_ok=True)
    flags = _arch_flags()
    stamp = open(STAMP).read().split() if os.path.exists(STAMP) else None
    stale = not os.path.exists(LIB) or os.path.getmtime(LIB) < os.path.getmtime(SRC)
    unusable = stamp is None or any(f not in flags for f in stamp)
    if force or stale or unusable:
        cmd = ["g++", "-O3", "-fopenmp", "-std=c++17", "-shared", "-fPIC"] + flags + ["-o", LIB, SRC]
        subprocess.run(cmd, check=True)
        with open(STAMP, "w") as fh:
            fh.write(" ".join(flags))
    return LIB


class ProveArgs(C.Structure):
    _fields_ = [
        ("curve", C.c_int), ("logn", C.c_int), ("omega", C.c_void_p), ("g", C.c_void_p),
        ("A_ext", C.c_void_p), ("B1_ext", C.c_void_p), ("B2_ext", C.c_void_p), ("K_ext", C.c_void_p), ("Z", C.c_void_p),
        ("mapA", C.c_void_p), ("mapB", C.c_void_p), ("mapK", C.c_void_p),
        ("m", C.c_uint64), ("nb_public", C.c_uint64), ("nZ", C.c_uint64),
        ("W_ext", C.c_void_p), ("a", C.c_void_p), ("b", C.c_void_p), ("c", C.c_void_p),
        ("out_ar", C.c_void_p), ("out_bs", C.c_void_p), ("out_krs", C.c_void_p), ("threads", C.c_int),
        ("sigma", C.c_void_p), ("cvals", C.c_void_p), ("n_commit", C.c_uint64), ("out_pok", C.c_void_p),
        ("part", C.c_int), ("nparts", C.c_int), ("comp_seconds", C.c_void_p),
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
        _lib.oc_fr_to_mont.argtypes = [i, vp, vp, u64, i]
        _lib.oc_gen_points.argtypes = [i, i, vp, u64, u64, vp, i]
        _lib.oc_num_threads.restype = i
        _lib.oc_num_procs.restype = i
    return _lib


def host_threads():
    """Threads the CPU baseline uses: the cores this process may run on (torchrun's OMP_NUM_THREADS=1 is ignored -
    every entry point takes an explicit thread count)."""
    try:
        return max(1, len(os.sched_getaffinity(0)))
    except AttributeError:
        return max(1, os.cpu_count() or 1)


def p(arr):
    return arr.ctypes.data if arr is not None else None
