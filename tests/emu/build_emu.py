#!/usr/bin/env python3
"""Build tests/emu/_build/libwap_emu.so: the product's .cu sources compiled by
g++ against the test-only CUDA emulator (cuda_emu.h).  TEST INFRASTRUCTURE:
lets the `-m "not gpu"` suite run the real kernel source on the CPU and compare
it with oracle/_ref.  The product (libwap_b200.so) never loads this."""
import glob
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.normpath(os.path.join(HERE, "..", ".."))
CSRC = os.path.join(ROOT, "webrtc-audio-processing_b200", "csrc")
OUT = os.path.join(HERE, "_build")
# WAP_EMU_O0=1: unoptimised variant (no code cloning), for WAP_EMU_CHECK_DIVERGENCE runs.
O0 = bool(os.environ.get("WAP_EMU_O0"))
LIB = os.path.join(OUT, "libwap_emu_O0.so" if O0 else "libwap_emu.so")


def build(verbose=True):
    os.makedirs(OUT, exist_ok=True)
    srcs = sorted(glob.glob(os.path.join(CSRC, "*.cu")) + glob.glob(os.path.join(CSRC, "*.cc")))
    deps = srcs + glob.glob(os.path.join(CSRC, "*.cuh")) + glob.glob(os.path.join(CSRC, "*.h")) + \
        glob.glob(os.path.join(CSRC, "*.inc")) + glob.glob(os.path.join(HERE, "cuda_emu.*")) + \
        glob.glob(os.path.join(ROOT, "include", "*.h"))
    if os.path.exists(LIB) and os.path.getmtime(LIB) >= max(os.path.getmtime(d) for d in deps):
        return LIB
    cmd = ["g++", "-std=c++17", "-O0" if O0 else "-O2", "-g", "-ffp-contract=off", "-fno-fast-math", "-fPIC", "-shared",
           "-DWAP_EMU=1", "-include", os.path.join(HERE, "cuda_emu.h"), "-I", HERE, "-I", CSRC,
           "-I", os.path.join(ROOT, "include"), "-Wall", "-Wno-unused-function", "-Wno-unknown-pragmas",
           "-Wno-unused-variable"]
    for s in srcs:
        cmd += ["-x", "c++", s]
    cmd += ["-x", "c++", os.path.join(HERE, "cuda_emu.cc"), "-o", LIB, "-lpthread", "-ldl"]
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode:
        sys.stderr.write(r.stdout + r.stderr)
        raise RuntimeError("emu build failed")
    if verbose:
        if r.stderr.strip():
            sys.stderr.write(r.stderr)
        print("built", LIB)
    return LIB


if __name__ == "__main__":
    build()
