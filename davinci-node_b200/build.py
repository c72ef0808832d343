"""Build recipe: generate the field headers, compile every CUDA translation unit for sm_100a
(in parallel, in-tree) and link libb200groth16.so next to this file.  nvcc cross-compiles without a
GPU, so this runs on the CPU build box; the .so travels to the GPU box with the repo snapshot."""
import concurrent.futures as cf
import glob
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
CSRC = os.path.join(HERE, "csrc")
# B200_BUILD_TAG=<name> builds an experiment variant (extra flags from B200_EXTRA_NVCC_FLAGS) next to the product library
_TAG = os.environ.get("B200_BUILD_TAG", "")
OBJ = os.path.join(HERE, "build" + ("_" + _TAG if _TAG else ""))
LIB = os.path.join(HERE, "libb200groth16%s.so" % ("_" + _TAG if _TAG else ""))

NVCC = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-std=c++17",
    "-Xcompiler", "-fPIC", "-Xcompiler", "-fvisibility=hidden", "--expt-relaxed-constexpr",
    "-I", CSRC, "-I", os.path.join(ROOT, "include"),
] + os.environ.get("B200_EXTRA_NVCC_FLAGS", "").split()


def _newest(paths):
    return max(os.path.getmtime(p) for p in paths)


def generate_fields():
    sys.path.insert(0, os.path.join(ROOT, "tools"))
    sys.path.insert(0, ROOT)
    import gen_field
    argv, sys.argv = sys.argv, ["gen_field.py", os.path.join(CSRC, "gen")]
    try:
        import io
        import contextlib
        with contextlib.redirect_stdout(io.StringIO()):
            gen_field.main()
    finally:
        sys.argv = argv


def build(verbose=False, force=False):
    generate_fields()
    os.makedirs(OBJ, exist_ok=True)
    headers = glob.glob(os.path.join(CSRC, "*.cuh")) + glob.glob(os.path.join(CSRC, "*.h")) + \
        glob.glob(os.path.join(CSRC, "gen", "*.cuh")) + glob.glob(os.path.join(ROOT, "include", "*.h"))
    hdr_time = _newest(headers)
    sources = sorted(glob.glob(os.path.join(CSRC, "*.cu")))
    jobs = []
    objs = []
    for src in sources:
        obj = os.path.join(OBJ, os.path.basename(src)[:-3] + ".o")
        objs.append(obj)
        if force or not os.path.exists(obj) or os.path.getmtime(obj) < max(hdr_time, os.path.getmtime(src)):
            jobs.append((src, obj))

    def compile_one(job):
        src, obj = job
        cmd = [NVCC] + NVCC_FLAGS + ["-c", src, "-o", obj]
        if verbose:
            cmd.insert(1, "-Xptxas=-v")
        r = subprocess.run(cmd, capture_output=True, text=True)
        return src, r.returncode, r.stdout + r.stderr

    if jobs:
        with cf.ThreadPoolExecutor(max_workers=min(len(jobs), os.cpu_count() or 4)) as ex:
            for src, rc, out in ex.map(compile_one, jobs):
                if verbose or rc:
                    sys.stderr.write("== %s\n%s\n" % (src, out))
                if rc:
                    raise RuntimeError("nvcc failed for %s" % src)
    if jobs or not os.path.exists(LIB) or os.path.getmtime(LIB) < _newest(objs):
        cmd = [NVCC, "-shared", "-o", LIB] + objs + ["-gencode", "arch=compute_100a,code=sm_100a"]
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode:
            raise RuntimeError("link failed: " + r.stdout + r.stderr)
    return LIB


if __name__ == "__main__":
    print(build(verbose="-v" in sys.argv, force="-f" in sys.argv))
