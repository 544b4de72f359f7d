"""Build the CUDA library in-tree for sm_100a (nvcc cross-compiles without a GPU)."""
from __future__ import annotations

import glob
import os
import subprocess

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIBDIR = os.path.join(HERE, "_lib")
LIB = os.path.join(LIBDIR, "libmcfost_b200.so")

NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17", "-Xcompiler", "-fPIC"]
# translation units and their extra flags:
#   api.cu       host API + deterministic sub-kernels: --fmad=false so that geometry is bit-exact
#                against the reference's non-contracted arithmetic (DESIGN.md "precision")
#   mc_kernel.cu the Monte Carlo photon-loop kernel: FMA contraction allowed (statistical path)
#   multi.cu     multi-GPU driver (host code only; NCCL is dlopened at run time)
UNITS = [("api.cu", ["--fmad=false"]), ("mc_kernel.cu", []), ("multi.cu", [])]


def sources():
    return sorted(glob.glob(os.path.join(CSRC, "*.cu")) + glob.glob(os.path.join(CSRC, "*.cuh"))
                  + [os.path.join(HERE, "..", "include", "mcfost_b200.h")])


def stale():
    if not os.path.exists(LIB):
        return True
    t = os.path.getmtime(LIB)
    return any(os.path.getmtime(s) > t for s in sources())


def build(force=False, verbose=False, defs=(), out=None):
    """defs / out: development builds (e.g. defs=("MCB_DEV", "MCB_DEV_CYL2D_ONLY"), out="libdev.so": only the kernels of
    the headline configuration, with the profiling knobs; select it with MCFOST_B200_LIB)."""
    lib = os.path.join(LIBDIR, out) if out else LIB
    if not (force or out or stale()):
        return LIB
    os.makedirs(LIBDIR, exist_ok=True)
    objs = []
    procs = []
    tag = ("." + os.path.splitext(out)[0]) if out else ""
    for src, extra in UNITS:
        obj = os.path.join(LIBDIR, src.replace(".cu", tag + ".o"))
        cmd = ["nvcc"] + NVCC_FLAGS + extra + ["-D" + d for d in defs] + (["-Xptxas", "-v"] if verbose else []) + ["-c", "-o", obj, os.path.join(CSRC, src)]
        procs.append((cmd, subprocess.Popen(cmd, cwd=CSRC)))
        objs.append(obj)
    for cmd, pr in procs:
        if pr.wait() != 0:
            raise subprocess.CalledProcessError(pr.returncode, cmd)
    subprocess.run(["nvcc", "-shared", "-o", lib] + objs + ["-ldl"], check=True, cwd=CSRC)
    return lib


if __name__ == "__main__":
    import sys
    if len(sys.argv) > 1 and sys.argv[1] == "dev":
        print(build(force=True, verbose=True, defs=("MCB_DEV", "MCB_DEV_CYL2D_ONLY") + tuple(sys.argv[2:]),
                    out=os.environ.get("MCB_DEV_OUT", "libmcfost_b200_dev.so")))
    else:
        print(build(force=True, verbose=True))
