"""-m gpu: the CUDA path (libwap_b200.so through its C ABI) against the compiled
reference (oracle/_ref) on the same inputs.  Bars: FFTs bit-exact; audio
per-sample max |delta| <= 1e-4 of full scale (BASELINE.json north_star)."""
import ctypes as C

import numpy as np
import pytest

from common import golden, run_engine

pytestmark = pytest.mark.gpu
TOL_FS = 1e-4  # of full scale => 3.2768 int16 LSB


def test_fft128_bit_exact(gpu_lib, oracle):
    rng = np.random.default_rng(7)
    x = (rng.standard_normal((257, 128)) * 4000).astype(np.float32)
    for inv in (0, 1):
        y = x.copy()
        assert gpu_lib.wapdbg_fft128(y.ctypes.data_as(C.c_void_p), len(y), inv) == 0
        r = np.stack([oracle.fft128(v, bool(inv)) for v in x])
        assert np.array_equal(y.view(np.uint32), r.view(np.uint32))


def test_fft256_bit_exact(gpu_lib, oracle):
    rng = np.random.default_rng(8)
    x = (rng.standard_normal((129, 256)) * 4000).astype(np.float32)
    for inv in (0, 1):
        y = x.copy()
        assert gpu_lib.wapdbg_fft256(y.ctypes.data_as(C.c_void_p), len(y), inv) == 0
        r = np.stack([oracle.rdft256(v, -1 if inv else 1) for v in x])
        assert np.array_equal(y.view(np.uint32), r.view(np.uint32))


@pytest.mark.parametrize("rate,level,tag", [(16000, 1, "ns_mod_16k"), (48000, 2, "ns_high_48k")])
def test_ns_parity_speech(gpu_lib, oracle, rate, level, tag):
    near = golden("speech_%dk.npz" % (rate // 1000))["near"]
    fl = rate // 100
    nf = min(600, near.size // fl)
    ref_out, _, err = oracle.RefApm(aec=False, ns=True, ns_level=level, max_rate=48000).run_i16(rate, None, near[:nf * fl])
    assert err == 0
    # the oracle itself is pinned against its recorded run
    assert np.array_equal(ref_out, golden("ref_outputs.npz")[tag][:nf * fl])
    out = run_engine(gpu_lib, rate, None, near[:nf * fl], n_streams=3, aec=False, ns=True, ns_level=level)
    d = np.abs(out.reshape(-1).astype(np.int32) - ref_out.astype(np.int32))
    assert d.max() <= TOL_FS * 32768, (d.max(), int(np.argmax(d > TOL_FS * 32768)) // fl)
