"""Builds libuwtrack.so (the C-ABI CUDA library) in-tree with nvcc for sm_100a.

Every .cu is compiled to its own object (in parallel, rebuilt only when it or a shared header
changed) under build/obj/, then linked into uw_slam_b200/libuwtrack.so."""
import concurrent.futures
import glob
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
CSRC = os.path.join(HERE, "csrc")
SO = os.path.join(HERE, "libuwtrack.so")
OBJ = os.path.join(ROOT, "build", "obj")
NVCC = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")

# -fmad=false: the arithmetic spec (docs/ARITHMETIC.md) rounds every float operation once;
# fused multiply-adds are written explicitly where the spec calls for them.
CFLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
    "-fmad=false", "-Xcompiler", "-fPIC",
    # host code too (camera model, ROI): no FMA contraction on hosts where GCC contracts by
    # default (aarch64, -march=native), or the "same values as OpenCV" maps would drift
    "-Xcompiler", "-ffp-contract=off",
]
LFLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-shared", "-cudart", "static",
          "-Xcompiler", "-fPIC"]
EXTRA = os.environ.get("UWT_NVCC_EXTRA", "").split()  # e.g. -DUWT_FLOW_STATS for debug builds


def sources():
    return sorted(glob.glob(os.path.join(CSRC, "*.cu")))


def headers():
    return glob.glob(os.path.join(CSRC, "*.cuh")) + \
        [os.path.join(ROOT, "include", "uwtrack.h"), __file__]


def obj_of(src):
    return os.path.join(OBJ, os.path.basename(src)[:-3] + ".o")


def stale(target, deps):
    if not os.path.exists(target):
        return True
    t = os.path.getmtime(target)
    return any(os.path.getmtime(d) > t for d in deps)


def needs_build():
    return stale(SO, sources() + headers())


def compile_one(src, verbose):
    cmd = [NVCC] + CFLAGS + EXTRA + (["-Xptxas", "-v"] if verbose else []) + \
        ["-c", "-o", obj_of(src), src]
    r = subprocess.run(cmd, capture_output=True, text=True)
    return src, r.returncode, r.stdout + r.stderr


def build(force=False, verbose=False):
    if not force and not needs_build():
        return SO
    os.makedirs(OBJ, exist_ok=True)
    todo = [s for s in sources() if force or EXTRA or stale(obj_of(s), [s] + headers())]
    with concurrent.futures.ThreadPoolExecutor(max_workers=max(1, len(todo))) as ex:
        for src, rc, log in ex.map(lambda s: compile_one(s, verbose), todo):
            if log.strip():
                sys.stderr.write(log)
            if rc != 0:
                raise subprocess.CalledProcessError(rc, "nvcc " + src)
    subprocess.check_call([NVCC] + LFLAGS + ["-o", SO] + [obj_of(s) for s in sources()])
    return SO


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
