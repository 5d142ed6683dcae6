"""Build libmicroaligner_b200.so (sm_100a) in-tree with nvcc. Usage: python build.py [--force]"""
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(ROOT, "microaligner_b200", "csrc")
OUT = os.path.join(ROOT, "microaligner_b200", "libmicroaligner_b200.so")
SOURCES = ["api.cu", "warp.cu", "pyramid.cu", "farneback.cu", "dog.cu", "nmi.cu", "hostcodec.cu"]
# -fmad=false: never contract a*b+c -- bit parity with OpenCV's SSE-baseline arithmetic depends on it;
# every fused multiply-add in the tree is an explicit intrinsic.
FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-std=c++17", "-fmad=false",
         "-Xcompiler", "-fPIC", "-Xcompiler", "-O2", "--shared", "-cudart", "shared"]


def needs_build():
    if not os.path.exists(OUT):
        return True
    t = os.path.getmtime(OUT)
    deps = [os.path.join(CSRC, f) for f in os.listdir(CSRC)] + [os.path.join(ROOT, "include", "microaligner_b200.h"), __file__]
    return any(os.path.getmtime(d) > t for d in deps)


def build(force=False, verbose=False):
    if not force and not needs_build():
        return OUT
    nvcc = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
    cmd = [nvcc] + FLAGS + (["-Xptxas", "-v"] if verbose else []) + [os.path.join(CSRC, s) for s in SOURCES] + ["-o", OUT]
    print(" ".join(cmd), flush=True)
    subprocess.check_call(cmd)
    return OUT


if __name__ == "__main__":
    build(force="--force" in sys.argv, verbose="-v" in sys.argv)
