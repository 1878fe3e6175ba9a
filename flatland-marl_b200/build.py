"""Builds the sm_100a CUDA library in-tree (flatland-marl_b200/csrc/libflatland_b200.so).

nvcc cross-compiles without a GPU; the built .so is git-ignored but travels to the GPU box."""
import os
import subprocess

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
SRC = os.path.join(HERE, "csrc", "flatland_b200.cu")
HDR = os.path.join(ROOT, "include", "flatland_b200.h")
LIB = os.path.join(HERE, "csrc", "libflatland_b200.so")

NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
              "--expt-extended-lambda", "-Xcompiler", "-fPIC", "-shared"]


def needs_build():
    if not os.path.exists(LIB):
        return True
    t = os.path.getmtime(LIB)
    srcs = [os.path.join(HERE, "csrc", f) for f in os.listdir(os.path.join(HERE, "csrc")) if f.endswith((".cu", ".cuh"))]
    return any(os.path.getmtime(p) > t for p in srcs + [HDR])


def build(force=False, verbose=False):
    if not (force or needs_build()):
        return LIB
    nvcc = os.environ.get("NVCC", "nvcc")
    cmd = [nvcc] + NVCC_FLAGS + (["-Xptxas", "-v"] if verbose else []) + \
          ["-I", os.path.join(ROOT, "include"), "-o", LIB, SRC]
    subprocess.check_call(cmd)
    return LIB


if __name__ == "__main__":
    print(build(force=True, verbose=True))
