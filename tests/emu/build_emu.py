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
# WAP_EMU_EXTRA="-DWAP_ECHO_LOCKSTEP=1": experiment variants of the kernel source (separate library name)
EXTRA = os.environ.get("WAP_EMU_EXTRA", "").split()
LIB = os.path.join(OUT, "libwap_emu_O0.so" if O0 else ("libwap_emu_x.so" if EXTRA else "libwap_emu.so"))


def build(verbose=True):
    import concurrent.futures as cf
    os.makedirs(OUT, exist_ok=True)
    srcs = sorted(glob.glob(os.path.join(CSRC, "*.cu")) + glob.glob(os.path.join(CSRC, "*.cc")))
    deps = srcs + glob.glob(os.path.join(CSRC, "*.cuh")) + glob.glob(os.path.join(CSRC, "*.h")) + \
        glob.glob(os.path.join(CSRC, "*.inc")) + glob.glob(os.path.join(HERE, "cuda_emu.*")) + \
        glob.glob(os.path.join(ROOT, "include", "*.h"))
    if os.path.exists(LIB) and os.path.getmtime(LIB) >= max(os.path.getmtime(d) for d in deps):
        return LIB
    base = ["g++", "-std=c++17", "-O0" if O0 else "-O2", "-g", "-ffp-contract=off", "-fno-fast-math", "-fno-strict-aliasing", "-fPIC",
            "-DWAP_EMU=1", "-include", os.path.join(HERE, "cuda_emu.h"), "-I", HERE, "-I", CSRC,
            "-I", os.path.join(ROOT, "include"), "-Wall", "-Wno-unused-function", "-Wno-unknown-pragmas",
            "-Wno-unused-variable"] + EXTRA
    # the same translation units as the product build: wap_k_echo.cu once per config class
    tus = []
    tag = "_O0" if O0 else ("_x" if EXTRA else "")
    for s in srcs:
        name = os.path.basename(s).rsplit(".", 1)[0]
        if name == "wap_k_echo":
            for rt in (0, 1):
                tus += [(s, ["-DWAP_ECHO_CLASS=%d" % c, "-DWAP_EC3_RUNTIME=%d" % rt],
                         os.path.join(OUT, "%s_%d%s%s.o" % (name, c, "_rt" if rt else "", tag))) for c in range(6)]
        elif name == "wap_k_delay":
            tus += [(s, ["-DWAP_EC3_RUNTIME=%d" % rt], os.path.join(OUT, name + ("_rt" if rt else "") + tag + ".o")) for rt in (0, 1)]
        else:
            tus.append((s, [], os.path.join(OUT, name + tag + ".o")))
    tus.append((os.path.join(HERE, "cuda_emu.cc"), [], os.path.join(OUT, "cuda_emu" + tag + ".o")))

    def one(tu):
        src, defs, obj = tu
        r = subprocess.run(base + defs + ["-x", "c++", "-c", src, "-o", obj], capture_output=True, text=True)
        return obj, r.returncode, r.stdout + r.stderr

    objs, warn = [], []
    with cf.ThreadPoolExecutor(max_workers=os.cpu_count() or 4) as ex:
        for obj, rc, out in ex.map(one, tus):
            objs.append(obj)
            if rc:
                sys.stderr.write(out)
                raise RuntimeError("emu build failed")
            if out.strip():
                warn.append(out)
    # (sanitizer flags in WAP_EMU_EXTRA must reach the link as well)
    r = subprocess.run(["g++", "-shared", "-o", LIB] + [f for f in EXTRA if f.startswith("-fsanitize")] + objs +
                       ["-lpthread", "-ldl"], capture_output=True, text=True)
    if r.returncode:
        sys.stderr.write(r.stdout + r.stderr)
        raise RuntimeError("emu link failed")
    if verbose:
        if warn:
            sys.stderr.write("".join(warn))
        print("built", LIB)
    return LIB


if __name__ == "__main__":
    build()
