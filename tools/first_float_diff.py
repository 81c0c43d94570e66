#!/usr/bin/env python3
"""First frame at which the float-interface output of the CUDA library and of the compiled
reference differ (GPU box): tools/first_float_diff.py <frames> <leg> "<config dict>" [rate]"""
import os
import sys

import numpy as np

ROOT = os.path.normpath(os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
for d in ("tests", "oracle", os.path.join("webrtc-audio-processing_b200", "python")):
    sys.path.insert(0, os.path.join(ROOT, d))
import ref  # noqa: E402
import wap_b200  # noqa: E402
from common import synthetic_leg_48k  # noqa: E402

nf, leg, kw = int(sys.argv[1]), int(sys.argv[2]), eval(sys.argv[3])
rate = int(sys.argv[4]) if len(sys.argv) > 4 else 16000
fl = rate // 100
far, near = synthetic_leg_48k(leg, nf, 2.0, rate=rate)
eng = wap_b200.Engine(1, rate, max_rate=32000, **kw)
r = ref.RefApm(max_rate=32000, **kw)
first, tot = None, 0
for f in range(nf):
    c = (near[f * fl:(f + 1) * fl].astype(np.float32) / 32768.0).reshape(1, fl)
    rr = (far[f * fl:(f + 1) * fl].astype(np.float32) / 32768.0).reshape(1, fl) if kw.get("aec") else None
    eng.set_stream_delay_ms(0)
    o = eng.process(rr, c).reshape(-1)
    ro, err = r.tick_f32(rate, None if rr is None else rr.reshape(-1), c.reshape(-1))
    nb = int(np.count_nonzero(o.view(np.uint32) != ro.view(np.uint32)))
    if nb and first is None:
        first = f
        print("first differing frame", f, "samples", nb, "max |d| (FloatS16)", float(np.abs(o - ro).max()) * 32768)
    tot += nb
print(kw, "differing samples", tot, "first frame", first)
