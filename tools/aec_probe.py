"""Emulator-vs-oracle AEC3 comparison (development tool; CPU only)."""
import os, sys, time
ROOT = os.path.normpath(os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
for p in ("oracle", "webrtc-audio-processing_b200/python", "tests", "tests/emu"):
    sys.path.insert(0, os.path.join(ROOT, p))
import numpy as np
import build_emu, wap_b200, ref
from common import golden, run_engine, synthetic_leg

nf = int(sys.argv[1]) if len(sys.argv) > 1 else 200
src = sys.argv[2] if len(sys.argv) > 2 else "speech"
ns = int(sys.argv[3]) if len(sys.argv) > 3 else 0
L = wap_b200.load(build_emu.build(verbose=False))
if src == "speech":
    sp = golden("speech_16k.npz")
    far, near = sp["far"][:nf * 160], sp["near"][:nf * 160]
else:
    far, near = synthetic_leg(int(src), nf)
t = time.time()
ref_out, stats, err = ref.RefApm(aec=True, ns=bool(ns), ns_level=1).run_i16(16000, far, near, stats_every=50)
print("ref", time.time() - t, "err", err)
t = time.time()
out = run_engine(L, 16000, far, near, n_streams=1, delay_ms=0, aec=True, ns=bool(ns), ns_level=1).reshape(-1)
print("emu", time.time() - t)
d = np.abs(out.astype(np.int32) - ref_out.astype(np.int32)).reshape(-1, 160).max(axis=1)
bad = np.nonzero(d > 3)[0]
print("max diff", d.max(), "first bad frame", bad[:10], "n bad", bad.size, "of", d.size)
print("per-50-frame max:", [int(d[i:i + 50].max()) for i in range(0, d.size, 50)])
print("ref stats (erl, erle, delay...):\n", stats[:8])
