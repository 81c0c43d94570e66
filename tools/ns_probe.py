"""Quick throughput probe through the host-buffer API (not the bench)."""
import sys, time, os
ROOT = os.path.normpath(os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
sys.path.insert(0, os.path.join(ROOT, "webrtc-audio-processing_b200", "python"))
import numpy as np
import wap_b200
L = wap_b200.load()
for rate, n in ((16000, 4096), (16000, 32768), (48000, 16384)):
    fl = rate // 100
    e = wap_b200.Engine(n, rate, lib=L, aec=False, ns=True, ns_level=2)
    rng = np.random.default_rng(0)
    cap = (rng.standard_normal((n, fl)) * 2000).astype(np.int16)
    for _ in range(5):
        e.process(None, cap)
    t = time.time()
    K = 20
    for _ in range(K):
        e.process(None, cap)
    dt = (time.time() - t) / K
    print("NS-only rate=%d streams=%d: %.3f ms/tick -> %.0f real-time streams (host-buffer API)" % (rate, n, dt * 1e3, n * 0.010 / dt))
    e.close()
