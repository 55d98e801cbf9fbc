"""Build libsessd_b200.so (hand-written CUDA for sm_100a + the C ABI of include/sessd_b200.h) IN-TREE.

    python se-ssd_b200/build.py [--force]

nvcc cross-compiles here without a GPU; the resulting .so is git-ignored but travels to the GPU box with gpurun.
Files that implement the rotated-box geometry are compiled with -fmad=false so that their fp32 arithmetic rounds
like the CPU twin of the reference (see csrc/rotbox.cuh).
"""
import os
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
OUT = os.path.join(HERE, "libsessd_b200.so")
OBJ = os.path.join(HERE, "build")
ARCH = ["-gencode", "arch=compute_100a,code=sm_100a"]
COMMON = ["-O3", "-lineinfo", "-std=c++17", "-Xcompiler", "-fPIC", "--expt-relaxed-constexpr"]
NO_FMA = {"iou3d.cu", "postproc.cu", "assign.cu", "odiou.cu"}


def sources():
    return sorted(f for f in os.listdir(CSRC) if f.endswith(".cu"))


def newest_dep():
    t = os.path.getmtime(os.path.join(os.path.dirname(HERE), "include", "sessd_b200.h"))
    for f in os.listdir(CSRC):
        t = max(t, os.path.getmtime(os.path.join(CSRC, f)))
    return t


def build(force=False, verbose=False):
    if (not force) and os.path.exists(OUT) and os.path.getmtime(OUT) >= newest_dep():
        return OUT
    os.makedirs(OBJ, exist_ok=True)
    nvcc = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")

    def compile_one(src):
        obj = os.path.join(OBJ, src[:-3] + ".o")
        cmd = [nvcc] + ARCH + COMMON + (["-fmad=false"] if src in NO_FMA else []) + ["-c", os.path.join(CSRC, src), "-o", obj]
        if verbose:
            print(" ".join(cmd))
        subprocess.check_call(cmd)
        return obj

    with ThreadPoolExecutor(max_workers=8) as ex:
        objs = list(ex.map(compile_one, sources()))
    cmd = [nvcc] + ARCH + ["-shared", "-o", OUT] + objs + ["-lcudart"]
    if verbose:
        print(" ".join(cmd))
    subprocess.check_call(cmd)
    return OUT


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose=True))
