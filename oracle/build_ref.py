#!/usr/bin/env python3
"""Build oracle/_ref/libwap_ref.so: the UNMODIFIED reference (webrtc-audio-processing
3.0 / WebRTC M145 AudioProcessing) compiled straight from the sources where they
lie under /root/reference, plus the harness driver oracle/ref_driver.cc.

Test infrastructure only (the checker for tests/, smoke() and bench.py's
cpu_baseline / --impl reference legs).  Nothing of the product links it.

Recipe: source list = every .c/.cc named by the
reference's meson.build files, minus MIPS/NEON variants, the Rust seam and two
TUs that need Android / video headers; rtc_base/cpu_info.cc is replaced by
oracle/cpu_info_stub.cc so the ISA path is selectable (AVX2 canonical).
abseil is replaced by the header-only shim in oracle/absl_shim.
Flags: -O2 -ffp-contract=off (GCC would otherwise fuse mul+add in the -mfma TUs
and silently change the reference's arithmetic).

Outputs go ONLY to oracle/_ref/ (git-ignored, travels to the GPU box).
If /root/reference is absent (GPU box) the prebuilt .so is used as is.
"""
import concurrent.futures as cf
import glob
import hashlib
import os
import re
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
REF = os.environ.get("WAP_REFERENCE_ROOT", "/root/reference")
OUT = os.path.join(HERE, "_ref")
OBJ = os.path.join(OUT, "obj")
LIB = os.path.join(OUT, "libwap_ref.so")

EXCLUDE_SUBSTR = ("_mips", "_neon", "rust_audio_processing",
                  "warn_current_thread_is_deadlocked", "api/video/video_timing",
                  "rtc_base/cpu_info.cc")
COMMON = ["-O2", "-ffp-contract=off", "-fPIC", "-DWEBRTC_LIBRARY_IMPL",
          "-DWEBRTC_ENABLE_SYMBOL_EXPORT", "-DNDEBUG", "-DWEBRTC_APM_DEBUG_DUMP=0",
          "-DWEBRTC_POSIX", "-DWEBRTC_LINUX", "-DWEBRTC_ENABLE_AVX2", "-D_GNU_SOURCE",
          "-w", "-I" + os.path.join(HERE, "absl_shim"),
          "-I" + os.path.join(REF, "webrtc"), "-I" + REF]


def reference_sources():
    srcs = set()
    for mb in glob.glob(os.path.join(REF, "webrtc", "**", "meson.build"), recursive=True):
        d = os.path.dirname(mb)
        for m in re.finditer(r"'([^']+\.(?:cc|c))'", open(mb).read()):
            p = os.path.normpath(os.path.join(d, m.group(1)))
            if os.path.exists(p) and not any(x in p for x in EXCLUDE_SUBSTR):
                srcs.add(p)
    return sorted(srcs)


DUMP = False  # second variant: -DWEBRTC_APM_DEBUG_DUMP=1 (stage taps via ApmDataDumper)


def compile_one(src):
    tag = hashlib.sha1(src.encode()).hexdigest()[:10]
    obj = os.path.join(OBJ + ("_dump" if DUMP else ""), os.path.basename(src).rsplit(".", 1)[0] + "_" + tag + ".o")
    if os.path.exists(obj) and os.path.getmtime(obj) >= os.path.getmtime(src):
        return obj, None
    cxx = src.endswith(".cc")
    cmd = (["g++", "-std=c++23"] if cxx else ["gcc", "-std=c11"]) + COMMON
    if DUMP:
        cmd = [c.replace("-DWEBRTC_APM_DEBUG_DUMP=0", "-DWEBRTC_APM_DEBUG_DUMP=1") for c in cmd]
    if "avx2" in os.path.basename(src):
        cmd += ["-mavx2", "-mfma"]
    cmd += ["-c", src, "-o", obj]
    r = subprocess.run(cmd, capture_output=True, text=True)
    return obj, (r.stderr if r.returncode else None)


def build(verbose=True, dump=False):
    """dump=True builds oracle/_ref/libwap_ref_dump.so: the same sources with the
    reference's own ApmDataDumper taps compiled in (used to localise divergences)."""
    global DUMP, LIB
    DUMP = dump
    LIB = os.path.join(OUT, "libwap_ref_dump.so" if dump else "libwap_ref.so")
    if not os.path.isdir(REF):
        if os.path.exists(LIB):
            if verbose:
                print("oracle/_ref: reference sources absent, using prebuilt", LIB)
            return LIB
        raise RuntimeError("no /root/reference and no prebuilt oracle/_ref/libwap_ref.so")
    os.makedirs(OBJ + ("_dump" if dump else ""), exist_ok=True)
    srcs = reference_sources() + [os.path.join(HERE, "cpu_info_stub.cc"),
                                  os.path.join(HERE, "ref_driver.cc")]
    if dump:  # ApmDataDumper::DumpWav needs the wav writer, which meson only builds for tests
        srcs += [os.path.join(REF, "webrtc", "common_audio", f) for f in ("wav_file.cc", "wav_header.cc")]
    objs, errs = [], []
    with cf.ThreadPoolExecutor(max_workers=os.cpu_count() or 4) as ex:
        for obj, err in ex.map(compile_one, srcs):
            objs.append(obj)
            if err:
                errs.append(err)
    if errs:
        sys.stderr.write("\n".join(errs[:5]))
        raise RuntimeError("oracle/_ref: %d translation units failed" % len(errs))
    newest = max(os.path.getmtime(o) for o in objs)
    if not os.path.exists(LIB) or os.path.getmtime(LIB) < newest:
        subprocess.check_call(["g++", "-shared", "-o", LIB] + objs + ["-lpthread", "-lm"])
    if verbose:
        print("oracle/_ref: built", LIB, "from", len(srcs), "TUs")
    return LIB


def build_seam(verbose=True):
    """oracle/_ref/libwap_seam.so: the reference's rust_audio_processing.cc, compiled UNMODIFIED with
    -DWEBRTC_USE_RUST_APM against this repo's include/wap_audio_processing.h, plus oracle/seam_driver.cc,
    linked against libwap_ref.so for the rest of webrtc.  Its wap_* references stay undefined: the
    test loads libwap_b200.so (or the emulator build) first with RTLD_GLOBAL."""
    lib = os.path.join(OUT, "libwap_seam.so")
    seam = os.path.join(REF, "webrtc", "modules", "audio_processing", "rust_audio_processing.cc")
    if not os.path.isdir(REF):
        if os.path.exists(lib):
            return lib
        raise RuntimeError("no /root/reference and no prebuilt oracle/_ref/libwap_seam.so")
    build(verbose=False)
    hdr = os.path.join(HERE, "..", "include", "wap_audio_processing.h")
    drv = os.path.join(HERE, "seam_driver.cc")
    if os.path.exists(lib) and os.path.getmtime(lib) >= max(os.path.getmtime(x) for x in (seam, hdr, drv)):
        return lib
    cmd = ["g++", "-std=c++23", "-shared"] + COMMON + ["-DWEBRTC_USE_RUST_APM", "-I" + os.path.join(HERE, "..", "include"),
           seam, drv, "-o", lib, "-L" + OUT, "-lwap_ref", "-Wl,-rpath,$ORIGIN"]
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode:
        sys.stderr.write(r.stderr[-4000:])
        raise RuntimeError("oracle/_ref: the reference seam does not compile against include/wap_audio_processing.h")
    if verbose:
        print("oracle/_ref: built", lib, "(reference seam, unmodified, against include/wap_audio_processing.h)")
    return lib


if __name__ == "__main__":
    build(dump="--dump" in sys.argv)
    build_seam()
