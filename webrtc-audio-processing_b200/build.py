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
         "-Xcompiler", "-fPIC", "-Xcompiler", "-fno-fast-math", "-Xcompiler", "-fno-strict-aliasing", "-shared",
         "-I", CSRC, "-I", os.path.join(ROOT, "include")]


ECHO_CLASSES = 6  # wap::EchoClass instances of k_echo (wap_k_echo.cu is compiled once per class)
OBJ = os.path.join(HERE, "_obj")


def translation_units():
    """(source, extra defines, object name): every .cu once, wap_k_echo.cu once per config class."""
    tus = []
    for s in sorted(glob.glob(os.path.join(CSRC, "*.cu")) + glob.glob(os.path.join(CSRC, "*.cc"))):
        base = os.path.basename(s).rsplit(".", 1)[0]
        if base == "wap_k_echo":
            # per config class, default EchoCanceller3Config (compile-time constants) and run-time parameters
            for rt in (0, 1):
                tus += [(s, ["-DWAP_ECHO_CLASS=%d" % c, "-DWAP_EC3_RUNTIME=%d" % rt], "%s_%d%s.o" % (base, c, "_rt" if rt else ""))
                        for c in range(ECHO_CLASSES)]
        elif base == "wap_k_delay":
            tus += [(s, ["-DWAP_EC3_RUNTIME=%d" % rt], base + ("_rt.o" if rt else ".o")) for rt in (0, 1)]
        else:
            tus.append((s, [], base + ".o"))
    return tus


def build(verbose=True, extra=(), lib=None):
    import concurrent.futures as cf
    lib = lib or LIB
    tus = translation_units()
    srcs = [t[0] for t in tus]
    deps = srcs + glob.glob(os.path.join(CSRC, "*.cuh")) + glob.glob(os.path.join(CSRC, "*.h")) + \
        glob.glob(os.path.join(CSRC, "*.inc")) + glob.glob(os.path.join(ROOT, "include", "*.h")) + [__file__]
    if not extra and os.path.exists(lib) and os.path.getmtime(lib) >= max(os.path.getmtime(d) for d in deps):
        return lib
    if not os.path.exists(NVCC):
        if os.path.exists(lib):
            return lib
        raise RuntimeError("nvcc not found and no prebuilt libwap_b200.so")
    os.makedirs(OBJ, exist_ok=True)
    compile_flags = [f for f in FLAGS if f != "-shared"]

    def one(tu):
        src, defs, name = tu
        obj = os.path.join(OBJ, name)
        cmd = [NVCC] + compile_flags + list(extra) + defs + (["-x", "cu"] if src.endswith(".cc") else []) + ["-c", src, "-o", obj]
        r = subprocess.run(cmd, capture_output=True, text=True)
        return obj, r.returncode, (r.stdout + r.stderr)

    objs, log = [], []
    with cf.ThreadPoolExecutor(max_workers=os.cpu_count() or 4) as ex:
        for obj, rc, out in ex.map(one, tus):
            objs.append(obj)
            if out.strip():
                log.append(out.strip())
            if rc:
                sys.stderr.write(out[-6000:])
                raise RuntimeError("nvcc build failed")
    r = subprocess.run([NVCC, "-shared", "-gencode", "arch=compute_100a,code=sm_100a", "-o", lib] + objs,
                       capture_output=True, text=True)
    if r.returncode:
        sys.stderr.write((r.stdout + r.stderr)[-6000:])
        raise RuntimeError("nvcc link failed")
    if verbose:
        if log:
            print("\n".join(log)[-6000:])
        print("built", lib)
    return lib


if __name__ == "__main__":
    build(extra=sys.argv[1:])
