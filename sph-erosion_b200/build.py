"""Builds libsphe_b200.so (hand-written sm_100a CUDA + the C ABI) in-tree with nvcc.

The .so is git-ignored but travels to the GPU box with the gpurun snapshot.
"""
import glob
import os
import shutil
import subprocess

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIBDIR = os.path.join(HERE, "lib")
LIB = os.path.join(LIBDIR, "libsphe_b200.so")

NVCC_FLAGS = ["-O3", "-std=c++17", "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo",
              "--extended-lambda", "-Xcompiler", "-fPIC", "-shared", "-cudart", "static"]


def sources():
    return sorted(glob.glob(os.path.join(CSRC, "*.cu")))


def _deps():
    return sources() + glob.glob(os.path.join(CSRC, "*.cuh")) + glob.glob(os.path.join(CSRC, "experiments", "*.cuh")) + glob.glob(os.path.join(CSRC, "*.h")) + \
        glob.glob(os.path.join(HERE, "..", "include", "*.h"))


def needs_build():
    if not os.path.exists(LIB):
        return True
    t = os.path.getmtime(LIB)
    return any(os.path.getmtime(f) > t for f in _deps())


# per-file extra flags: the terrain contact search must round every multiply/add separately (terrain.cu)
EXTRA = {"terrain.cu": ["-fmad=false"]}


NVCC_FLAGS += os.environ.get("SPHE_NVCC_EXTRA", "").split()   # A/B builds, e.g. SPHE_NVCC_EXTRA="-DFL_THREADS=256 -DFL_MINB=3"
if os.environ.get("SPHE_WITH_EXPERIMENTS") == "1":
    NVCC_FLAGS.append("-DSPHE_WITH_EXPERIMENTS")   # also compile csrc/experiments/ (rejected round-1 kernel variants)


def build_library(force=False, verbose=False):
    """nvcc -gencode arch=compute_100a,code=sm_100a ... -> sph-erosion_b200/lib/libsphe_b200.so"""
    if not force and not needs_build():
        return LIB
    nvcc = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    if not os.path.exists(nvcc):
        raise RuntimeError("nvcc not found and %s is missing or stale" % LIB)
    os.makedirs(LIBDIR, exist_ok=True)
    objdir = os.path.join(HERE, "build")
    os.makedirs(objdir, exist_ok=True)
    compile_flags = [f for f in NVCC_FLAGS if f != "-shared"]
    procs, objs = [], []
    for src in sources():
        obj = os.path.join(objdir, os.path.basename(src)[:-3] + ".o")
        objs.append(obj)
        cmd = [nvcc] + compile_flags + EXTRA.get(os.path.basename(src), []) + (["-Xptxas", "-v"] if verbose else []) + ["-c", "-o", obj, src]
        procs.append((cmd, subprocess.Popen(cmd, cwd=HERE)))
    for cmd, p in procs:
        if p.wait() != 0:
            raise subprocess.CalledProcessError(p.returncode, cmd)
    subprocess.check_call([nvcc, "-shared", "-cudart", "static", "-gencode", "arch=compute_100a,code=sm_100a", "-o", LIB] + objs, cwd=HERE)
    return LIB


if __name__ == "__main__":
    print(build_library(force=True, verbose=True))
