"""-m gpu: the CUDA path (libwap_b200.so through its C ABI) against the compiled
reference (oracle/_ref) on the same inputs.  Bars: FFTs bit-exact; audio
per-sample max |delta| <= 1e-4 of full scale (BASELINE.json north_star)."""
import ctypes as C

import numpy as np
import pytest

from common import golden, loud_bursty_signal, run_engine, run_legs, synthetic_leg, synthetic_leg_48k, vanishing_echo_leg

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


def _check_aec(out, stats, ref_out, ref_stats, tag):
    d = np.abs(out.astype(np.int32) - ref_out.astype(np.int32))
    assert d.max() <= TOL_FS * 32768, (tag, int(d.max()), int(np.argmax(d > TOL_FS * 32768)) // 160)
    # ERLE within 0.1 dB (BASELINE.json north_star); ERL / delay as diagnostics
    assert np.abs(stats[:, 1] - ref_stats[:, 3]).max() <= 0.1, tag
    assert np.abs(stats[:, 0] - ref_stats[:, 1]).max() <= 0.1, tag
    assert np.array_equal(stats[:, 2], ref_stats[:, 5]), tag


def test_aec3_ns_parity_speech(gpu_lib, oracle):
    """BASELINE config 1: 16 kHz mono AEC3 + NS(moderate) on the reference's speech fixture (7 s)."""
    sp = golden("speech_16k.npz")
    nf = sp["near"].size // 160
    far, near = sp["far"][:nf * 160], sp["near"][:nf * 160]
    ref_out, ref_stats, err = oracle.RefApm(aec=True, ns=True, ns_level=1).run_i16(16000, far, near, stats_every=50)
    assert err == 0
    n_gold = min(ref_out.size, golden("ref_outputs.npz")["aec_ns_16k"].size)
    assert np.array_equal(ref_out[:n_gold], golden("ref_outputs.npz")["aec_ns_16k"][:n_gold])
    out, stats = run_legs(gpu_lib, 16000, [(far, near)] * 2, stats_every=50, aec=True, ns=True, ns_level=1)
    assert np.array_equal(out[0], out[1])
    _check_aec(out[0], stats[0], ref_out, ref_stats, "speech")


@pytest.mark.parametrize("ns", [False, True])
def test_aec3_parity_synthetic_legs_10s(gpu_lib, oracle, ns):
    """BASELINE config 2 shape at test size: 24 legs with echo-path delays spread over
    4-196 ms in ONE batched engine, 10 s each (passes the 2.5 s initial state, the 500-block
    ERLE start-up, the NS 200-frame start-up and one 500-frame NS histogram update)."""
    ids = list(range(0, 48, 2))
    legs = [synthetic_leg(i, 1000) for i in ids]
    out, stats = run_legs(gpu_lib, 16000, legs, stats_every=100, aec=True, ns=ns, ns_level=1)
    for k, (far, near) in enumerate(legs):
        ref_out, ref_stats, err = oracle.RefApm(aec=True, ns=ns, ns_level=1).run_i16(16000, far, near, stats_every=100)
        assert err == 0
        _check_aec(out[k], stats[k], ref_out, ref_stats, "leg %d" % ids[k])
    # the canceller did something: ERLE above 5 dB for most legs at the end
    assert np.median(stats[:, -1, 1]) > 5.0


def test_aec3_no_render_passes_capture_through(gpu_lib, oracle):
    """Edge case (block_processor.cc:123-131): no render data => capture is only delayed by one block."""
    far, near = synthetic_leg(3, 50)
    ref = oracle.RefApm(aec=True, ns=False)
    ref_out, _, err = ref.run_i16(16000, None, near)
    assert err == 0
    out = run_engine(gpu_lib, 16000, None, near, n_streams=2, delay_ms=0, aec=True, ns=False)
    assert np.abs(out.reshape(-1).astype(np.int32) - ref_out.astype(np.int32)).max() <= TOL_FS * 32768


def test_aec3_saturated_capture_and_silence(gpu_lib, oracle):
    """Edge cases: full-scale (saturating) capture, then digital silence on both sides."""
    rng = np.random.default_rng(5)
    n = 300 * 160
    far = (rng.uniform(-1, 1, n) * 20000).astype(np.int16)
    near = np.clip(np.roll(far.astype(np.int32), 200) * 2, -32768, 32767).astype(np.int16)  # clipped echo
    far[200 * 160:] = 0
    near[200 * 160:] = 0
    ref_out, ref_stats, err = oracle.RefApm(aec=True, ns=True, ns_level=1).run_i16(16000, far, near, stats_every=50)
    assert err == 0
    out, stats = run_legs(gpu_lib, 16000, [(far, near)], stats_every=50, aec=True, ns=True, ns_level=1)
    _check_aec(out[0], stats[0], ref_out, ref_stats, "saturated")


def test_aec3_echo_path_vanishes_loud_render(gpu_lib, oracle):
    """Misadjustment rescale + coarse re-seed paths (subtractor.cc:246-257,297-316,345-375)."""
    far, near = vanishing_echo_leg(800)
    ref_out, ref_stats, err = oracle.RefApm(aec=True, ns=True, ns_level=1).run_i16(16000, far, near, stats_every=50)
    assert err == 0
    out, stats = run_legs(gpu_lib, 16000, [(far, near)], stats_every=50, aec=True, ns=True, ns_level=1)
    _check_aec(out[0], stats[0], ref_out, ref_stats, "vanishing echo")


def test_aec3_ns_parity_48k_three_band(gpu_lib, oracle):
    """48 kHz mono AEC3 + NS (three bands, upper-band gain / comfort noise, PostFilter): the speech
    fixture and six synthetic legs, two of them with a clipped, HF-heavy render."""
    sp = golden("speech_48k.npz")
    n = 300 * 480
    legs = [(sp["far"][:n], sp["near"][:n])] + [synthetic_leg_48k(i, 300, 3.5 if i % 3 == 2 else 1.0) for i in range(6)]
    out, stats = run_legs(gpu_lib, 48000, legs, stats_every=50, aec=True, ns=True, ns_level=1)
    for k, (far, near) in enumerate(legs):
        ref_out, ref_stats, err = oracle.RefApm(aec=True, ns=True, ns_level=1, max_rate=48000).run_i16(
            48000, far, near, stats_every=50)
        assert err == 0
        d = np.abs(out[k].astype(np.int32) - ref_out.astype(np.int32))
        assert d.max() <= TOL_FS * 32768, (k, int(d.max()))
        assert np.abs(stats[k][:, 1] - ref_stats[:, 3]).max() <= 0.1, k
        assert np.array_equal(stats[k][:, 2], ref_stats[:, 5]), k


@pytest.mark.parametrize("rate", [16000, 32000, 48000])
def test_full_chain_aec3_ns_agc2(gpu_lib, oracle, rate):
    """BASELINE config 5 shape at test size: AEC3 + NS + AGC2 (fixed 6 dB gain + limiter), at the
    three native rates (one band, two-band QMF, three-band filter bank)."""
    if rate == 16000:
        legs = [synthetic_leg(i, 400) for i in range(4)]
    else:
        legs = [synthetic_leg_48k(i, 400, 1.0 + i, rate=rate) for i in range(4)]
    out, stats = run_legs(gpu_lib, rate, legs, stats_every=100, aec=True, ns=True, ns_level=1, agc2=True,
                          agc2_fixed_gain_db=6.0)
    for k, (far, near) in enumerate(legs):
        ref_out, ref_stats, err = oracle.RefApm(aec=True, ns=True, ns_level=1, max_rate=48000, agc2=True,
                                                agc2_fixed_gain_db=6.0).run_i16(rate, far, near, stats_every=100)
        assert err == 0
        d = np.abs(out[k].astype(np.int32) - ref_out.astype(np.int32))
        assert d.max() <= TOL_FS * 32768, (k, int(d.max()))
        assert np.abs(stats[k][:, 1] - ref_stats[:, 3]).max() <= 0.1, k


def test_agc2_limiter_regions(gpu_lib, oracle):
    x = loud_bursty_signal(16000, 300)
    ref_out, _, err = oracle.RefApm(aec=False, ns=False, agc2=True, agc2_fixed_gain_db=12.0).run_i16(16000, None, x)
    assert err == 0 and np.abs(ref_out).max() >= 32000
    out = run_engine(gpu_lib, 16000, None, x, n_streams=2, aec=False, ns=False, agc2=True, agc2_fixed_gain_db=12.0)
    assert np.abs(out.reshape(-1).astype(np.int32) - ref_out.astype(np.int32)).max() <= TOL_FS * 32768


@pytest.mark.parametrize("chunks", [3, 0])
def test_large_batch_pipelined_host_path(gpu_lib, oracle, chunks):
    """wap_process_streams on a large batch cuts the legs into ranges whose PCIe copies overlap
    the kernels of another range (3 forced ragged ranges of 8200 legs; automatic = 4 wave-aligned
    ranges of 47400 legs): 8 distinct legs tiled, every copy equals the reference output."""
    ids = [1, 6, 11, 18, 23, 30, 37, 44]
    nf = 150 if chunks else 40
    base = [synthetic_leg(i, nf) for i in ids]
    n = 8200 if chunks else 47400
    legs = [base[i % 8] for i in range(n)]
    out, _ = run_legs(gpu_lib, 16000, legs, pipeline_chunks=chunks, aec=True, ns=True, ns_level=1)
    for k in range(8):
        far, near = base[k]
        ref_out, _, err = oracle.RefApm(aec=True, ns=True, ns_level=1).run_i16(16000, far, near)
        assert err == 0
        d = np.abs(out[k].astype(np.int32) - ref_out.astype(np.int32)).max()
        assert d <= TOL_FS * 32768, (k, d)
        assert np.array_equal(out[k::8], np.broadcast_to(out[k], out[k::8].shape))


def test_ns_parity_32k_two_band(gpu_lib, oracle):
    """32 kHz NS-only (very high) with the two-band QMF, 6 s of held 16 kHz speech."""
    near = np.repeat(golden("speech_16k.npz")["near"][:600 * 160], 2)
    ref_out, _, err = oracle.RefApm(aec=False, ns=True, ns_level=3).run_i16(32000, None, near)
    assert err == 0
    out = run_engine(gpu_lib, 32000, None, near, n_streams=3, aec=False, ns=True, ns_level=3)
    d = np.abs(out.reshape(-1).astype(np.int32) - ref_out.astype(np.int32)).max()
    assert d <= TOL_FS * 32768, d


@pytest.mark.parametrize("rate,kw", [
    (48000, dict(aec=True, ns=True, ns_level=1, agc2=True, agc2_fixed_gain_db=6.0)),
    (48000, dict(aec=False, ns=True, ns_level=2)),
    (44100, dict(aec=True, ns=True, ns_level=1)),
    (8000, dict(aec=True, ns=True, ns_level=1)),
])
def test_resampled_rates(gpu_lib, oracle, rate, kw):
    """API rate != processing rate under the reference's default maximum_internal_processing_rate
    (32 kHz): sinc resamplers in and out, 48 kHz fullband side path; 6 legs x 4 s."""
    legs = [synthetic_leg_48k(i, 400, 1.0 + 0.5 * i, rate=rate) for i in range(6)]
    out, stats = run_legs(gpu_lib, rate, legs, stats_every=100 if kw["aec"] else 0, max_rate=32000, **kw)
    for k, (far, near) in enumerate(legs):
        ref_out, ref_stats, err = oracle.RefApm(max_rate=32000, **kw).run_i16(
            rate, far if kw["aec"] else None, near, stats_every=100 if kw["aec"] else 0)
        assert err == 0
        d = np.abs(out[k].astype(np.int32) - ref_out.astype(np.int32)).max()
        assert d <= TOL_FS * 32768, (k, int(d))
        if kw["aec"]:
            assert np.abs(stats[k][:, 1] - ref_stats[:, 3]).max() <= 0.1, k


def test_libm_restatement_on_device(gpu_lib):
    """powf(2, p) / tanhf as glibc evaluates them, on the device: identical bits."""
    import os
    import sys
    sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tools"))
    import check_libm_restatement as chk
    for which in (0, 1, 2):
        bad, tot = chk.mismatches(gpu_lib, which, 100000)
        assert bad == 0, (which, bad, tot)


@pytest.mark.parametrize("rate,max_rate,kw", [
    (16000, 32000, dict(aec=True, ns=True, ns_level=1)),
    (32000, 32000, dict(aec=False, ns=True, ns_level=3)),
    (48000, 48000, dict(aec=True, ns=True, ns_level=1)),
    (48000, 32000, dict(aec=True, ns=True, ns_level=1, agc2=True, agc2_fixed_gain_db=6.0)),
])
def test_float_interface_is_bit_identical(gpu_lib, oracle, rate, max_rate, kw):
    """Float interface, 5 s: every output sample has the reference's float32 bits."""
    from common import float_interface_max_diff
    differing, worst = float_interface_max_diff(gpu_lib, oracle, rate, 500, max_rate=max_rate, **kw)
    assert differing == 0, (differing, worst)


@pytest.mark.parametrize("rate,max_rate,right_gain", [(16000, 32000, 4.0), (48000, 32000, 0.6), (48000, 48000, 3.0)])
def test_stereo_default_pipeline(gpu_lib, oracle, rate, max_rate, right_gain):
    """Stereo in / out with the default pipeline and AEC3+NS: 4 s, bit-identical."""
    from common import run_stereo_i16
    out, ref_out = run_stereo_i16(gpu_lib, oracle, rate, 400, 9, max_rate=max_rate, right_gain=right_gain,
                                  aec=True, ns=True, ns_level=1)
    d = np.abs(out.astype(np.int32) - ref_out.astype(np.int32)).max()
    assert d <= TOL_FS * 32768, int(d)


def test_level_adjustment_and_runtime_settings(gpu_lib, oracle):
    """Pre / post gains, AGC2 fixed gain and playout volume changed at run time, 48 kHz default config."""
    from common import run_with_runtime_settings
    events = [(10, "pre_gain", 2.5), (20, "post_gain", 0.5), (30, "playout_volume", 100), (40, "playout_volume", 180),
              (50, "fixed_post_gain", 12.0), (70, "pre_gain", 0.7), (90, "fixed_post_gain", -3.0), (120, "post_gain", 3.0)]
    kw = dict(aec=True, ns=True, ns_level=1, agc2=True, agc2_fixed_gain_db=3.0, pre_amp=2.0, pre_gain=1.2, post_gain=1.1)
    for rate, max_rate in ((48000, 32000), (16000, 32000), (48000, 48000)):
        assert run_with_runtime_settings(gpu_lib, oracle, rate, 200, events, max_rate=max_rate, **kw) == 0


def test_long_run_60s_float_identity(gpu_lib, oracle):
    """One minute of AEC3 + NS at 16 kHz through the float interface (12 NS histogram periods, ERLE
    and reverb estimators long past start-up, several render-silence gaps): still every float32 bit."""
    from common import float_interface_max_diff
    differing, worst = float_interface_max_diff(gpu_lib, oracle, 16000, 6000, leg=13, aec=True, ns=True, ns_level=1)
    assert differing == 0, (differing, worst)


@pytest.mark.parametrize("rate", [16000, 48000])
def test_agc2_limiter_float_identity(gpu_lib, oracle, rate):
    """The limiter's attack interpolation (a power of 8 the reference takes from libm) at float level."""
    import wap_b200
    x = loud_bursty_signal(rate, 600)
    fl = rate // 100
    eng = wap_b200.Engine(1, rate, lib=gpu_lib, aec=False, ns=False, agc2=True, agc2_fixed_gain_db=12.0)
    ref = oracle.RefApm(aec=False, ns=False, agc2=True, agc2_fixed_gain_db=12.0)
    differing = 0
    for f in range(600):
        c = (x[f * fl:(f + 1) * fl].astype(np.float32) / 32768.0).reshape(1, fl)
        o = eng.process(None, c).reshape(-1)
        ro, err = ref.tick_f32(rate, None, c.reshape(-1))
        assert err == 0
        differing += int(np.count_nonzero(o.view(np.uint32) != ro.view(np.uint32)))
    eng.close()
    assert differing == 0


def test_bench_line_contract(gpu_lib):
    """`python bench.py` (small arguments): one JSON line with the bench contract's keys, the kernels
    really launched, a roofline object for the dominant kernel and a bounded CPU baseline."""
    import json
    import os
    import subprocess
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    r = subprocess.run([sys.executable, os.path.join(root, "bench.py"), "--streams", "512", "--steps", "6", "--warmup", "3",
                        "--cpu-seconds", "1.5"], capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [l for l in r.stdout.splitlines() if l.strip()]
    assert len(lines) == 1
    d = json.loads(lines[0])
    for key in ("metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling",
                "vs_baseline", "dtype", "data", "config", "e2e", "gpu_launches", "clocks", "roofline", "cpu_baseline"):
        assert key in d, key
    assert d["steps"] == 6 and d["warmup"] == 3 and d["n_gpus"] == 1 and d["scaling"] == "weak"
    assert d["gpu_launches"] == 18 and d["value"] > 0 and d["e2e"]["value"] > 0
    assert d["e2e"]["h2d_bytes_per_step"] == 2 * 512 * 160 * 2 and d["e2e"]["d2h_bytes_per_step"] == 512 * 160 * 2
    rf = d["roofline"]
    assert rf["bound"] == "hbm" and rf["kernel"] in ("k_echo", "k_delay") and 0 < rf["frac"] < 1 and rf["peak"] > 1000
    assert d["cpu_baseline"]["kind"] == "reference" and d["cpu_baseline"]["value"] > 0
    # steady state by construction, and the timed batch is checked against the compiled reference
    assert d["config"]["settle_ticks"] >= 300
    pc = d["parity_spot_check"]
    assert pc["legs"] == 16 and pc["pass"] and pc["max_abs_diff_lsb"] == 0, pc
    assert "profile-derived" in (rf["traffic_source"] or "profile-derived")
    names = [o["name"] for o in d["other_configs"]]
    assert any("config 3" in n for n in names) and any("config 5" in n for n in names)
    for o in d["other_configs"]:
        assert o["value"] > 0 and o["e2e"]["value"] > 0 and o["parity_spot_check"]["pass"], o
