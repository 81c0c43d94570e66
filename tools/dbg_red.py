"""Debug helper: the residual echo detector's state after a few frames, emulator build vs device build."""
import ctypes as C, os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "webrtc-audio-processing_b200", "python")); sys.path.insert(0, os.path.join(ROOT, "tests"))
import wap_b200
import importlib.util
spec = importlib.util.spec_from_file_location("ted", os.path.join(ROOT, "tests", "test_echo_detector.py"))
sys.path.insert(0, os.path.join(ROOT, "oracle"))
ted = importlib.util.module_from_spec(spec); spec.loader.exec_module(ted)
far, near = ted._residual_echo_leg(40)
RED = 4 * (30 + 4 + 4 * 650 + 1 + 4 + 2 + 2 + 3 + 4)
names = [("render_buffer", 30), ("rb_next", 1), ("rb_count", 1), ("frames_since_zero", 1), ("seen_capture", 1), ("render_power", 650),
         ("render_power_mean", 650), ("render_power_std_dev", 650), ("covariance", 650), ("next_insertion_index", 1), ("render_mean", 1),
         ("render_variance", 1), ("capture_mean", 1), ("capture_variance", 1), ("reliability", 1), ("echo_likelihood", 1), ("mm_max", 1),
         ("mm_counter", 1), ("stats_valid", 1), ("stats_likelihood", 1), ("stats_recent_max", 1), ("slot_full", 1), ("slot_valid", 1),
         ("slot_likelihood", 1), ("slot_recent_max", 1)]
def f32_power(x):
    acc = np.float32(0)
    for v in x.astype(np.float32):
        acc = np.float32(acc + np.float32(v * v))
    return acc, np.float32(acc / np.float32(x.size))
blobs = {}
for tag, path in (("emu", os.path.join(ROOT, "tests", "emu", "_build", "libwap_emu.so")), ("gpu", None)):
    L = wap_b200.load(path)
    eng = wap_b200.Engine(1, 16000, lib=L, aec=True, ns=True, ns_level=1, echo_detector=True)
    per = []
    for f in range(int(sys.argv[1]) if len(sys.argv) > 1 else 8):
        sl = slice(f * 160, (f + 1) * 160)
        eng.set_stream_delay_ms(0)
        o = eng.process((far[sl].astype(np.float32) / 32768).reshape(1, -1), (near[sl].astype(np.float32) / 32768).reshape(1, -1))
        if f < 3:
            x = (o[0] * np.float32(32768)).astype(np.float32)
            acc, pw = f32_power(x)
            print(tag, f, "sum", acc.view(np.uint32), "power", pw, hex(pw.view(np.uint32)), "alpha*power", hex(np.float32(np.float32(0.001) * pw).view(np.uint32)))
        per.append(np.array(eng.export_state(0)[-RED:]).view(np.uint32).copy())
    blobs[tag] = per
    eng.close()
for f, (a, b) in enumerate(zip(blobs["emu"], blobs["gpu"])):
    d = np.nonzero(a != b)[0]
    if d.size:
        print("frame", f, "differing words", d[:20])
        off = 0
        for nm, n in names:
            for i in d:
                if off <= i < off + n:
                    print("  ", nm, i - off, a[i:i + 1].view(np.float32), b[i:i + 1].view(np.float32), hex(a[i]), hex(b[i]))
            off += n
        break
else:
    print("no differences")
