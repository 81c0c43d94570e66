#!/usr/bin/env python3
"""Check the restated glibc routines (csrc/wap_libm.cuh: powf(2, p), tanhf) against the libm the
oracle links, through the debug entry point wapdbg_libm of the emulator build (default) or of the
CUDA library (--gpu).  usage: tools/check_libm_restatement.py [--gpu] [n]"""
import ctypes as C
import os
import sys

import numpy as np

ROOT = os.path.normpath(os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
sys.path.insert(0, os.path.join(ROOT, "webrtc-audio-processing_b200", "python"))
sys.path.insert(0, os.path.join(ROOT, "tests", "emu"))


def libm_reference(which, x):
    m = C.CDLL("libm.so.6")
    m.powf.restype = C.c_float
    m.powf.argtypes = [C.c_float, C.c_float]
    m.tanhf.restype = C.c_float
    m.tanhf.argtypes = [C.c_float]
    m.tanh.restype = C.c_double
    m.tanh.argtypes = [C.c_double]
    if which == 0:
        return np.array([m.powf(2.0, float(v)) for v in x], np.float32)
    if which == 1:
        return np.array([m.tanhf(float(v)) for v in x], np.float32)
    # the NS prior-model indicator: (float)(0.5 * (tanh((double)x) + 1)); the sum magnifies last-bit
    # differences of tanh near -1
    return np.array([np.float32(0.5 * (np.float64(m.tanh(float(v))) + 1.0)) for v in x], np.float32)


def arguments(which, n, seed=7):
    rng = np.random.default_rng(seed + which)
    if which == 0:
        parts = [rng.uniform(-30, 30, n // 2), rng.uniform(-2, 2, n // 4), rng.uniform(-125, 125, n // 4)]
    elif which == 1:
        parts = [rng.uniform(-1.2, 1.2, n // 2), rng.uniform(-9, 9, n // 4), rng.uniform(-1e-3, 1e-3, n // 8),
                 rng.uniform(-30, 30, n // 8)]
    else:
        parts = [rng.uniform(-1.2, 1.2, n // 4), rng.uniform(-20, 0, n // 2), rng.uniform(-40, 40, n // 4)]
    return np.concatenate(parts).astype(np.float32)


def mismatches(lib, which, n):
    x = arguments(which, n)
    y = x.copy()
    assert lib.wapdbg_libm(y.ctypes.data_as(C.c_void_p), len(y), which) == 0
    ref = libm_reference(which, x)
    return int(np.count_nonzero(y.view(np.uint32) != ref.view(np.uint32))), len(x)


if __name__ == "__main__":
    import wap_b200
    gpu = "--gpu" in sys.argv
    nums = [a for a in sys.argv[1:] if a.isdigit()]
    n = int(nums[0]) if nums else 400000
    if gpu:
        lib = wap_b200.load()
    else:
        import build_emu
        lib = wap_b200.load(build_emu.build())
    for which, name in ((0, "powf(2, p)"), (1, "tanhf"), (2, "tanh")):
        bad, tot = mismatches(lib, which, n)
        print("%-10s %d mismatches of %d" % (name, bad, tot))
