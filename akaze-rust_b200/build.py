"""Builds libakaze_b200.so (CUDA, sm_100a only) in-tree with nvcc.

--fmad=false: the reference (rustc) never contracts a*b+c into an FMA, and the parity target for the
stencil stages is bit-exactness, so nvcc must not contract either.

Every source is compiled to its own object file (in parallel, only when stale) and the objects are linked
into one shared library; no relocatable device code is needed (no cross-file device calls).
"""
import os
import shutil
import subprocess
from concurrent.futures import ThreadPoolExecutor

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
OBJ = os.path.join(HERE, "build")
LIB = os.path.join(HERE, "libakaze_b200.so")
LIB_FAST = os.path.join(HERE, "libakaze_b200_fast.so")  # opt-in AKZ_FAST_MATH build (fused multiply-adds in the stencils)
FAST_SOURCES = ("scale_space.cu", "detector.cu")        # the only sources that see -DAKZ_FAST_MATH
SOURCES = ["akaze_api.cu", "scale_space.cu", "detector.cu", "keypoints.cu", "matcher.cu", "matcher_tc.cu", "ransac.cu"]
HEADERS = ["common.cuh", "tile_util.cuh", "nccl_dyn.h", os.path.join("..", "..", "include", "akaze_b200.h")]

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-O3", "-lineinfo", "--fmad=false", "-std=c++17",
    "-Xcompiler", "-fPIC",
]


def nvcc_path():
    p = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    if not os.path.exists(p):
        raise RuntimeError("nvcc not found: the CUDA extension cannot be built")
    return p


def _deps(src):
    return [os.path.join(CSRC, src)] + [os.path.join(CSRC, h) for h in HEADERS] + [os.path.abspath(__file__)]


def _obj(src, fast=False):
    return os.path.join(OBJ, src.replace(".cu", "_fast.o" if fast and src in FAST_SOURCES else ".o"))


def _stale(target, deps):
    if not os.path.exists(target):
        return True
    t = os.path.getmtime(target)
    return any(os.path.getmtime(d) > t for d in deps)


def is_stale(fast=False):
    return _stale(LIB_FAST if fast else LIB, [d for s in SOURCES for d in _deps(s)])


def build(force=False, verbose=False, fast=False):
    """Compile every CUDA source for sm_100a into one shared library. Returns its path. fast=True builds the opt-in
    fused-multiply-add variant of the stencil kernels (libakaze_b200_fast.so; never loaded unless AKZ_FAST_MATH=1)."""
    lib = LIB_FAST if fast else LIB
    if not force and not is_stale(fast):
        return lib
    os.makedirs(OBJ, exist_ok=True)
    nvcc = nvcc_path()

    def compile_one(src):
        if not force and not _stale(_obj(src, fast), _deps(src)):
            return ""
        extra = ["-DAKZ_FAST_MATH"] if fast and src in FAST_SOURCES else []
        cmd = [nvcc] + NVCC_FLAGS + extra + (["-Xptxas", "-v"] if verbose else []) + ["-c", os.path.join(CSRC, src), "-o", _obj(src, fast)]
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError("nvcc failed on %s:\n%s%s" % (src, r.stdout, r.stderr))
        return r.stderr

    with ThreadPoolExecutor(max_workers=len(SOURCES)) as ex:
        logs = list(ex.map(compile_one, SOURCES))
    if verbose:
        print("\n".join(logs))
    r = subprocess.run([nvcc, "-shared", "-o", lib] + [_obj(s, fast) for s in SOURCES] + ["-ldl"], capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError("link failed:\n" + r.stdout + r.stderr)
    return lib


if __name__ == "__main__":
    import sys
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv, fast="--fast" in sys.argv))
