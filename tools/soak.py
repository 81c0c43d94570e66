#!/usr/bin/env python3
"""Float-identity soak on the GPU box: several legs per configuration through the float interface of
the CUDA library and of the compiled reference; prints the number of differing samples per config.
usage: tools/soak.py [frames] [legs]"""
import os
import sys

import numpy as np

ROOT = os.path.normpath(os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
for d in ("tests", "oracle", os.path.join("webrtc-audio-processing_b200", "python")):
    sys.path.insert(0, os.path.join(ROOT, d))
import ref  # noqa: E402
import wap_b200  # noqa: E402
from common import synthetic_leg, synthetic_leg_48k  # noqa: E402

nf = int(sys.argv[1]) if len(sys.argv) > 1 else 6000
nl = int(sys.argv[2]) if len(sys.argv) > 2 else 6
CONFIGS = [
    (16000, 32000, dict(aec=True, ns=True, ns_level=1)),
    (16000, 32000, dict(aec=True, ns=True, ns_level=3, agc2=True, agc2_fixed_gain_db=9.0)),
    (32000, 32000, dict(aec=True, ns=True, ns_level=2)),
    (48000, 48000, dict(aec=True, ns=True, ns_level=1)),
    (48000, 32000, dict(aec=True, ns=True, ns_level=1, agc2=True, agc2_fixed_gain_db=6.0)),
    (44100, 32000, dict(aec=True, ns=True, ns_level=0)),
    (8000, 32000, dict(aec=True, ns=True, ns_level=1)),
]
total_bad = 0
for rate, max_rate, kw in CONFIGS:
    fl = rate // 100
    legs = [synthetic_leg(3 + 5 * i, nf) if rate == 16000 else synthetic_leg_48k(3 + 5 * i, nf, 1.0 + 0.5 * i, rate=rate)
            for i in range(nl)]
    far = np.stack([l[0] for l in legs]).reshape(nl, nf, fl).astype(np.float32) / 32768.0
    near = np.stack([l[1] for l in legs]).reshape(nl, nf, fl).astype(np.float32) / 32768.0
    eng = wap_b200.Engine(nl, rate, max_rate=max_rate, **kw)
    refs = [ref.RefApm(max_rate=max_rate, **kw) for _ in range(nl)]
    bad, first = 0, None
    for f in range(nf):
        eng.set_stream_delay_ms(0)
        o = eng.process(np.ascontiguousarray(far[:, f]), np.ascontiguousarray(near[:, f]))
        for i in range(nl):
            ro, err = refs[i].tick_f32(rate, far[i, f], near[i, f])
            nb = int(np.count_nonzero(o[i].view(np.uint32) != ro.view(np.uint32)))
            if nb and first is None:
                first = (f, i)
            bad += nb
    eng.close()
    total_bad += bad
    print(rate, max_rate, kw, "differing samples", bad, "first (frame, leg)", first, flush=True)
print("TOTAL", total_bad)
