"""Host-side cost of one tick's bookkeeping (no GPU work waited for): tools/host_overhead.py [legs]"""
import ctypes as C
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.join(os.path.dirname(__file__), "..", "webrtc-audio-processing_b200", "python"))
import torch  # noqa: E402  (device buffers)
import wap_b200  # noqa: E402

S = int(sys.argv[1]) if len(sys.argv) > 1 else 65536
eng = wap_b200.Engine(S, 16000, aec=True, ns=True)
L = eng.lib
r = torch.zeros(S, 160, dtype=torch.int16, device="cuda")
c = torch.zeros(S, 160, dtype=torch.int16, device="cuda")
o = torch.zeros(S, 160, dtype=torch.int16, device="cuda")
for it in range(6):
    L.wap_engine_synchronize(eng.h)
    t0 = time.perf_counter()
    L.wap_streams_set_delay_ms(eng.handles, S, 0)
    t1 = time.perf_counter()
    L.wap_process_streams_device(eng.h, eng.handles, S, r.data_ptr(), c.data_ptr(), o.data_ptr(), 0)
    t2 = time.perf_counter()
    L.wap_engine_synchronize(eng.h)
    t3 = time.perf_counter()
    print("set_delay %.3f ms, enqueue %.3f ms, wait %.3f ms" % ((t1 - t0) * 1e3, (t2 - t1) * 1e3, (t3 - t2) * 1e3))
