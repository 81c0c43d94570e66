#!/usr/bin/env python3
"""A short run of every kernel path (16 / 32 / 48 kHz, resampled, stereo, AGC2, level adjustment,
mute, ragged last CTA) -- a quick smoke of all config classes on the GPU box, and the workload to put
under compute-sanitizer where that tool is available (it is closed on this pool):
  [compute-sanitizer --tool memcheck|racecheck|synccheck] python tools/sanitize_run.py"""
import os
import sys

import numpy as np

ROOT = os.path.normpath(os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
for d in ("tests", os.path.join("webrtc-audio-processing_b200", "python")):
    sys.path.insert(0, os.path.join(ROOT, d))
import wap_b200  # noqa: E402
from common import stereo_leg, synthetic_leg, synthetic_leg_48k  # noqa: E402

NF = int(sys.argv[1]) if len(sys.argv) > 1 else 24
CONFIGS = [
    (16000, 1, 32000, dict(aec=True, ns=True, ns_level=1)),
    (16000, 1, 32000, dict(aec=False, ns=True, ns_level=3, agc2=True, agc2_fixed_gain_db=12.0)),
    (32000, 1, 32000, dict(aec=True, ns=True, ns_level=2, pre_gain=1.5, post_gain=0.7)),
    (48000, 1, 48000, dict(aec=True, ns=True, ns_level=1, agc2=True, agc2_fixed_gain_db=6.0)),
    (48000, 1, 32000, dict(aec=True, ns=True, ns_level=1, agc2=True, agc2_fixed_gain_db=6.0)),
    (44100, 1, 32000, dict(aec=True, ns=False)),
    (48000, 2, 32000, dict(aec=True, ns=True, ns_level=1)),
]
for rate, ch, max_rate, kw in CONFIGS:
    fl = rate // 100 * ch
    n = 5  # two CTAs, the second one ragged
    legs = []
    for i in range(n):
        if ch == 2:
            legs.append(stereo_leg(rate, NF, 3 + i))
        else:
            legs.append(synthetic_leg(3 + i, NF) if rate == 16000 else synthetic_leg_48k(3 + i, NF, 2.0, rate=rate))
    eng = wap_b200.Engine(n, rate, channels=ch, max_rate=max_rate, **kw)
    acc = 0
    for f in range(NF):
        if f == 10:
            eng.set_capture_output_used(False, legs=[1])
            eng.set_playout_volume(77)
        if f == 15:
            eng.set_capture_output_used(True, legs=[1])
        r = np.stack([l[0][f * fl:(f + 1) * fl] for l in legs])
        c = np.stack([l[1][f * fl:(f + 1) * fl] for l in legs])
        eng.set_stream_delay_ms(0)
        acc += int(np.abs(eng.process(r if kw["aec"] else None, c).astype(np.int32)).sum())
    eng.stats(0)
    eng.close()
    print(rate, ch, max_rate, kw, "ok", acc, flush=True)
