"""The SURVEY.md 8(d) synthetic call-leg generator (measurement / test infrastructure).

One generator for every arm: bench.py's B200 arm, its reference CPU arm, the in-bench parity spot
check and the tests all take their int16 frames from `cycle()`.  The fast path is
tools/wap_synth.c (built by __graft_entry__.build() into tools/_build/libwap_synth.so, which
travels to the GPU box); `cycle_numpy()` is the same algorithm in numpy (vectorised over legs) and
the two are compared bit for bit in tests/test_cpu.py.
"""
import ctypes as C
import os
import subprocess

import numpy as np

ROOT = os.path.normpath(os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
SRC = os.path.join(ROOT, "tools", "wap_synth.c")
LIB = os.path.join(ROOT, "tools", "_build", "libwap_synth.so")
_lib = None


def build(verbose=False):
    if os.path.exists(LIB) and os.path.getmtime(LIB) >= os.path.getmtime(SRC):
        return LIB
    os.makedirs(os.path.dirname(LIB), exist_ok=True)
    r = subprocess.run(["gcc", "-O2", "-shared", "-fPIC", "-fopenmp", SRC, "-o", LIB, "-lm"],
                       capture_output=True, text=True)
    if r.returncode:
        raise RuntimeError("wap_synth build failed: " + r.stderr)
    if verbose:
        print("built", LIB)
    return LIB


def _load():
    global _lib
    if _lib is None:
        try:
            if os.path.exists(SRC):
                build()
        except Exception:
            pass
        if os.path.exists(LIB):
            L = C.CDLL(LIB)
            L.wap_synth_cycle.argtypes = [C.c_int] * 5 + [C.c_void_p, C.c_void_p]
            _lib = L
    return _lib


def cycle(kind, rate, first_leg, legs, frames, render=None, capture=None, force_numpy=False):
    """int16 render / capture [frames][legs][rate // 100] of legs first_leg .. first_leg + legs - 1:
    one seamless cycle (see tools/wap_synth.c).  Pre-allocated (e.g. pinned) output arrays may be
    passed in."""
    fl = rate // 100 * (2 if kind == 2 else 1)   # kind 2: interleaved stereo frames
    if render is None:
        render = np.zeros((frames, legs, fl), np.int16)
    if capture is None:
        capture = np.zeros((frames, legs, fl), np.int16)
    L = None if force_numpy else _load()
    if L is None:
        r, c = cycle_numpy(kind, rate, first_leg, legs, frames)
        render[...] = r
        capture[...] = c
        return render, capture
    assert render.flags["C_CONTIGUOUS"] and capture.flags["C_CONTIGUOUS"]
    err = L.wap_synth_cycle(kind, rate, first_leg, legs, frames, render.ctypes.data_as(C.c_void_p),
                            capture.ctypes.data_as(C.c_void_p))
    assert err == 0
    return render, capture


class _VecRandom:
    """webrtc::Random (reference rtc_base/random.h:71-77, random.cc:52-56), one state per leg."""

    def __init__(self, seeds):
        self.s = np.asarray(seeds, dtype=np.uint64).copy()

    def rand_float(self):
        s = self.s
        s ^= s >> np.uint64(12)
        s ^= s << np.uint64(25)
        s ^= s >> np.uint64(27)
        out = s * np.uint64(2685821657736338717)
        v = (out - np.uint64(1)).astype(np.float64) / float(0xFFFFFFFFFFFFFFFF)
        return v.astype(np.float32)

    def sample(self, amplitude):
        a = np.float32(amplitude)
        return np.float32(2) * a * self.rand_float() - a


def cycle_numpy(kind, rate, first_leg, legs, frames):
    fl = rate // 100
    n = frames * fl
    i = np.arange(first_leg, first_leg + legs, dtype=np.int64)
    rr = _VecRandom(1000 + 2 * i)
    rn = _VecRandom(1001 + 2 * i)
    k = np.arange(n)
    y = np.zeros((legs, n), np.float64)
    x = np.zeros((legs, n), np.float32)
    with np.errstate(over="ignore"):
        if kind == 0:
            for t in range(n):
                v = rr.sample(8000.0)
                x[:, t] = v if (t % rate) < (rate // 10) * 9 else 0.0
            D = (rate // 16000) * (64 * (1 + (i % 48)) + (7 * i) % 64)
            rows = np.arange(legs)[:, None]
            xd = x.astype(np.float64)
            y = 0.5 * xd[rows, (k[None, :] - D[:, None]) % n] + 0.25 * xd[rows, (k[None, :] - D[:, None] - 37) % n] + \
                0.1 * xd[rows, (k[None, :] - D[:, None] - 160) % n]
            for t in range(n):
                floor_ = rn.sample(50.0)
                burst = rn.sample(3000.0)
                y[:, t] += floor_
                if (t % (2 * rate)) >= (rate // 10) * 17:
                    y[:, t] += burst
        elif kind == 2:
            rr2 = _VecRandom(1000 + 2 * i + 7919)
            rn2 = _VecRandom(1001 + 2 * i + 7919)
            xr = np.zeros((legs, n), np.float32)
            for t in range(n):
                v, w = rr.sample(8000.0), rr2.sample(8000.0)
                on = (t % rate) < (rate // 10) * 9
                x[:, t] = v if on else 0.0
                xr[:, t] = w if on else 0.0
            D = (rate // 16000) * (64 * (1 + (i % 48)) + (7 * i) % 64)
            rows = np.arange(legs)[:, None]
            xd, xrd = x.astype(np.float64), xr.astype(np.float64)
            at = lambda a, d: a[rows, (k[None, :] - d[:, None]) % n]
            ys = []
            for c in range(2):
                Dc = D + 3 * c
                ys.append(0.5 * at(xd, Dc) + 0.25 * at(xd, Dc + 37) + 0.1 * at(xd, Dc + 160) +
                          0.35 * at(xrd, Dc + 11) + 0.15 * at(xrd, Dc + 53))
            for t in range(n):
                floor0 = rn.sample(50.0)
                burst = rn.sample(3000.0)
                floor1 = rn2.sample(50.0)
                ys[0][:, t] += floor0
                ys[1][:, t] += floor1
                if (t % (2 * rate)) >= (rate // 10) * 17:
                    ys[0][:, t] += burst
                    ys[1][:, t] += burst
            q = lambda a: np.clip(np.rint(a), -32768, 32767).astype(np.int16)
            il = lambda a, b: np.stack([a, b], 2).reshape(legs, frames, fl * 2).transpose(1, 0, 2).copy()
            return il(q(x), q(xr)), il(q(ys[0]), q(ys[1]))
        else:
            for t in range(n):
                y[:, t] = rn.sample(300.0)
            tt = k / rate
            on = (k % rate) < rate // 2
            y += (4000.0 * np.sin(2 * np.pi * 1000.0 * tt) + 4000.0 * np.sin(2 * np.pi * 2300.0 * tt)) * on
    q = lambda a: np.clip(np.rint(a), -32768, 32767).astype(np.int16)
    to = lambda a: a.reshape(legs, frames, fl).transpose(1, 0, 2).copy()
    return to(q(x)), to(q(y))
