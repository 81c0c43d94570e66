#!/usr/bin/env python3
"""Build libwap_b200.so (the product): hand-written sm_100a CUDA kernels plus
the C-ABI host engine, in-tree so the .so travels to the GPU box.

Numerics flags are part of the parity contract:
  -fmad=false      no implicit FMA contraction (explicit fmaf only where the
                   reference's AVX2 path fuses)
  -ftz=true        the reference runs with FTZ/DAZ (DenormalDisabler)
  -prec-div/-prec-sqrt=true   IEEE division and square root
"""
import glob
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.normpath(os.path.join(HERE, ".."))
CSRC = os.path.join(HERE, "csrc")
LIB = os.path.join(HERE, "libwap_b200.so")
NVCC = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")

FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
         "-fmad=false", "-ftz=true", "-prec-div=true", "-prec-sqrt=true",
         "-Xcompiler", "-fPIC", "-Xcompiler", "-fno-fast-math", "-shared",
         "-I", CSRC, "-I", os.path.join(ROOT, "include")]


def build(verbose=True, extra=()):
    srcs = sorted(glob.glob(os.path.join(CSRC, "*.cu")) + glob.glob(os.path.join(CSRC, "*.cc")))
    deps = srcs + glob.glob(os.path.join(CSRC, "*.cuh")) + glob.glob(os.path.join(CSRC, "*.h")) + \
        glob.glob(os.path.join(CSRC, "*.inc")) + glob.glob(os.path.join(ROOT, "include", "*.h")) + [__file__]
    if not extra and os.path.exists(LIB) and os.path.getmtime(LIB) >= max(os.path.getmtime(d) for d in deps):
        return LIB
    if not os.path.exists(NVCC):
        if os.path.exists(LIB):
            return LIB
        raise RuntimeError("nvcc not found and no prebuilt libwap_b200.so")
    cmd = [NVCC] + FLAGS + list(extra)
    for s in srcs:
        cmd += (["-x", "cu", s] if s.endswith(".cc") else [s])
    cmd += ["-o", LIB]
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode:
        sys.stderr.write((r.stdout + r.stderr)[-6000:])
        raise RuntimeError("nvcc build failed")
    if verbose:
        tail = (r.stdout + r.stderr).strip()
        if tail:
            print(tail[-4000:])
        print("built", LIB)
    return LIB


if __name__ == "__main__":
    build(extra=sys.argv[1:])
