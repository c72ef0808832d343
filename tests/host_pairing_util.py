"""Builds and loads tests/host_pairing.cu: the product's pairing template (csrc/pairing.cuh) compiled for the CPU."""
import ctypes as C
import os
import subprocess

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
SRC = os.path.join(HERE, "host_pairing.cu")
OUT = os.path.join(HERE, "_build", "libhost_pairing.so")
CSRC = os.path.join(ROOT, "davinci-node_b200", "csrc")
CURVE_ID = {"bn254": 1, "bls12_377": 2, "bls12_381": 3, "bw6_761": 4, "bn254_fermat": 101}
FP_LIMBS = {"bn254": 8, "bls12_377": 12, "bls12_381": 12, "bw6_761": 24, "bn254_fermat": 8}
_lib = None


def _ensure_generated():
    if not os.path.exists(os.path.join(CSRC, "gen", "pairing_consts.cuh")):
        import importlib.util
        spec = importlib.util.spec_from_file_location("b200_build", os.path.join(ROOT, "davinci-node_b200", "build.py"))
        mod = importlib.util.module_from_spec(spec)
        spec.loader.exec_module(mod)
        mod.generate_fields()


def load():
    global _lib
    if _lib is not None:
        return _lib
    _ensure_generated()
    deps = [SRC, os.path.join(CSRC, "pairing.cuh"), os.path.join(CSRC, "field.cuh"), os.path.join(CSRC, "gen", "pairing_consts.cuh")]
    if not os.path.exists(OUT) or os.path.getmtime(OUT) < max(os.path.getmtime(d) for d in deps):
        os.makedirs(os.path.dirname(OUT), exist_ok=True)
        nvcc = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
        cmd = [nvcc, "-O2", "-std=c++17", "-gencode", "arch=compute_100a,code=sm_100a", "--expt-relaxed-constexpr",
               "-Xcompiler", "-fPIC", "-shared", "-I", CSRC, "-I", os.path.join(ROOT, "include"), "-o", OUT, SRC]
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode:
            raise RuntimeError("host pairing harness failed to build:\n" + r.stdout + r.stderr)
    _lib = C.CDLL(OUT)
    for f in (_lib.hp_pair, _lib.hp_ext_mul, _lib.hp_check):
        f.restype = C.c_int
    return _lib


def limbs(x, n):
    return [(x >> (32 * i)) & 0xFFFFFFFF for i in range(n)]


def pack(vals, n):
    flat = []
    for v in vals:
        flat += limbs(int(v), n)
    return (C.c_uint32 * len(flat))(*flat)


def unpack(buf, count, n):
    return tuple(sum(buf[i * n + j] << (32 * j) for j in range(n)) for i in range(count))


def g1_coords(P):
    return [0, 0] if P is None else [P[0], P[1]]


def g2_coords(Q, name):
    if name == "bw6_761":
        return [0, 0] if Q is None else [Q[0], Q[1]]
    name = name.replace("_fermat", "")
    return [0, 0, 0, 0] if Q is None else [Q[0][0], Q[0][1], Q[1][0], Q[1][1]]


def pair(name, P, Q, which=1):
    """(value as a k-tuple of ints, P-in-subgroup flag) from the host build of PairingT."""
    lib = load()
    n = FP_LIMBS[name]
    k = 6 if name == "bw6_761" else 12
    out = (C.c_uint32 * (k * n))()
    rc = lib.hp_pair(CURVE_ID[name], pack(g1_coords(P), n), pack(g2_coords(Q, name), n), out, which)
    assert rc in (0, 1), rc
    return unpack(out, k, n), bool(rc)


def ext_mul(name, a, b):
    lib = load()
    n = FP_LIMBS[name]
    k = len(a)
    out = (C.c_uint32 * (k * n))()
    assert lib.hp_ext_mul(CURVE_ID[name], pack(a, n), pack(b, n), out) == 0
    return unpack(out, k, n)


def check(name, pairs):
    lib = load()
    n = FP_LIMBS[name]
    g1, g2 = [], []
    for P, Q in pairs:
        g1 += g1_coords(P)
        g2 += g2_coords(Q, name)
    return lib.hp_check(CURVE_ID[name], pack(g1, n), pack(g2, n), len(pairs))
