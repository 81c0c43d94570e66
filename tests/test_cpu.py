"""-m "not gpu": (1) the oracle (compiled reference, oracle/_ref) is pinned against the
reference's own known-answer vectors and against its recorded outputs; (2) the C ABI
library loads and exports every symbol include/*.h declares; (3) the kernel SOURCE,
compiled for the CPU warp emulator (tests/emu, test infrastructure), matches the
oracle -- this is how kernels are debugged on the GPU-less build box; the parity
tests proper are the -m gpu ones."""
import ctypes as C
import os
import re

import numpy as np
import pytest

from common import ROOT, golden, run_engine

TOL_FS = 1e-4


# ------------------------------------------------------------------ oracle pinning
def test_oracle_isa_is_avx2(oracle):
    assert oracle.lib().ref_isa_level() == 2


def test_oracle_hpf_known_answer(oracle):
    """Reference KAT: tests/unit/high_pass_filter_unittest.cc:198-330 (kReferenceInput /
    kReference, tolerance 1/32768, last frame compared)."""
    kat = golden("hpf_kat.npz")
    L = oracle.lib()
    for name in ("MonoInitial", "MonoConverged"):
        x, want = kat[name + "_in"], kat[name + "_ref"]
        h = L.ref_hpf_create(16000, 1)
        last = None
        for f in range(x.size // 160):
            fr = x[f * 160:(f + 1) * 160].astype(np.float32).copy()
            L.ref_hpf_process(h, fr.ctypes.data_as(C.c_void_p), 1, 160)
            last = fr
        L.ref_hpf_destroy(h)
        assert np.abs(last[:want.size] - want).max() <= 1.0 / 32768


@pytest.mark.parametrize("tag,kw,rate,use_far", [
    ("ns_mod_16k", dict(aec=False, ns=True, ns_level=1), 16000, False),
    ("aec_ns_16k", dict(aec=True, ns=True, ns_level=1), 16000, True),
])
def test_oracle_matches_recorded_run(oracle, tag, kw, rate, use_far):
    """Detects a silently different oracle build (flags, ISA path)."""
    sp = golden("speech_%dk.npz" % (rate // 1000))
    n = 300 * rate // 100
    out, _, err = oracle.RefApm(max_rate=48000, **kw).run_i16(rate, sp["far"][:n] if use_far else None, sp["near"][:n])
    assert err == 0
    assert np.array_equal(out, golden("ref_outputs.npz")[tag][:n])


# ------------------------------------------------------------------ C ABI surface
def _declared_symbols():
    syms = set()
    for fn in os.listdir(os.path.join(ROOT, "include")):
        src = open(os.path.join(ROOT, "include", fn)).read()
        src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
        syms |= set(re.findall(r"\b(wap_[a-z0-9_]+)\s*\(", src))
    return syms


def test_abi_library_exports_every_declared_symbol():
    import wap_b200
    L = wap_b200.load()  # the CUDA build; loads without a GPU (cudart is linked statically)
    declared = _declared_symbols()
    assert declared and declared == set(wap_b200.EXPORTS), declared ^ set(wap_b200.EXPORTS)
    for s in declared:
        assert hasattr(L, s), s


def test_product_does_not_link_oracle_or_emulator():
    import subprocess
    import wap_b200
    out = subprocess.run(["ldd", wap_b200.DEFAULT_LIB], capture_output=True, text=True).stdout
    assert "wap_ref" not in out and "wap_emu" not in out
    for root, _, files in os.walk(os.path.join(ROOT, "webrtc-audio-processing_b200")):
        for f in files:
            if f.endswith((".cu", ".cuh", ".h", ".inc", ".py", ".cc")):
                txt = open(os.path.join(root, f)).read()
                assert "libwap_ref" not in txt and "import ref" not in txt, f


# ------------------------------------------------------------------ kernel source on the emulator
def test_emu_fft128_bit_exact(emu_lib, oracle):
    rng = np.random.default_rng(7)
    x = (rng.standard_normal((33, 128)) * 4000).astype(np.float32)
    for inv in (0, 1):
        y = x.copy()
        assert emu_lib.wapdbg_fft128(y.ctypes.data_as(C.c_void_p), len(y), inv) == 0
        r = np.stack([oracle.fft128(v, bool(inv)) for v in x])
        assert np.array_equal(y.view(np.uint32), r.view(np.uint32))


def test_emu_fft256_bit_exact(emu_lib, oracle):
    rng = np.random.default_rng(8)
    x = (rng.standard_normal((17, 256)) * 4000).astype(np.float32)
    for inv in (0, 1):
        y = x.copy()
        assert emu_lib.wapdbg_fft256(y.ctypes.data_as(C.c_void_p), len(y), inv) == 0
        r = np.stack([oracle.rdft256(v, -1 if inv else 1) for v in x])
        assert np.array_equal(y.view(np.uint32), r.view(np.uint32))


def test_emu_hpf_known_answer(emu_lib):
    """The device HPF against the reference's literal KAT (high_pass_filter_unittest.cc)."""
    import wap_b200
    kat = golden("hpf_kat.npz")
    for name in ("MonoInitial", "MonoConverged"):
        x, want = kat[name + "_in"], kat[name + "_ref"]
        eng = wap_b200.Engine(1, 16000, lib=emu_lib, aec=False, ns=False, hpf=True)
        last = None
        for f in range(x.size // 160):
            # KAT samples are fed in the float [-1,1] convention (x32768 is exact in fp32)
            last = eng.process(None, (x[f * 160:(f + 1) * 160] / 8.0).astype(np.float32)[None, :])[0] * 8.0
        eng.close()
        assert np.abs(last[:want.size] - want).max() <= 1.0 / 32768


def test_emu_ns_parity_16k(emu_lib, oracle):
    near = golden("speech_16k.npz")["near"][:250 * 160]
    ref_out, _, err = oracle.RefApm(aec=False, ns=True, ns_level=1).run_i16(16000, None, near)
    assert err == 0
    out = run_engine(emu_lib, 16000, None, near, n_streams=2, aec=False, ns=True, ns_level=1)
    d = np.abs(out.reshape(-1).astype(np.int32) - ref_out.astype(np.int32))
    assert d.max() <= TOL_FS * 32768, d.max()


def test_emu_ns_parity_48k_three_band(emu_lib, oracle):
    near = golden("speech_48k.npz")["near"][:120 * 480]
    ref_out, _, err = oracle.RefApm(aec=False, ns=True, ns_level=2, max_rate=48000).run_i16(48000, None, near)
    assert err == 0
    out = run_engine(emu_lib, 48000, None, near, n_streams=1, aec=False, ns=True, ns_level=2)
    d = np.abs(out.reshape(-1).astype(np.int32) - ref_out.astype(np.int32))
    assert d.max() <= TOL_FS * 32768, d.max()


# ------------------------------------------------------------------ AEC3 on the emulator
def _check_aec(out, stats, ref_out, ref_stats):
    d = np.abs(out.astype(np.int32) - ref_out.astype(np.int32))
    assert d.max() <= TOL_FS * 32768, (int(d.max()), int(np.argmax(d > TOL_FS * 32768)))
    # ERLE within 0.1 dB (north_star); ERL and the reported delay as diagnostics.
    assert np.abs(stats[:, 1] - ref_stats[:, 3]).max() <= 0.1
    assert np.abs(stats[:, 0] - ref_stats[:, 1]).max() <= 0.1
    assert np.array_equal(stats[:, 2], ref_stats[:, 5])


def test_emu_aec3_ns_parity_speech(emu_lib, oracle):
    """cfg 1 shape: AEC3 + NS(moderate), 16 kHz mono, the reference's own far/near speech fixture."""
    from common import run_legs
    sp = golden("speech_16k.npz")
    n = 300 * 160
    ref_out, ref_stats, err = oracle.RefApm(aec=True, ns=True, ns_level=1).run_i16(
        16000, sp["far"][:n], sp["near"][:n], stats_every=50)
    assert err == 0
    assert np.array_equal(ref_out, golden("ref_outputs.npz")["aec_ns_16k"][:n])
    out, stats = run_legs(emu_lib, 16000, [(sp["far"][:n], sp["near"][:n])], stats_every=50,
                          aec=True, ns=True, ns_level=1)
    _check_aec(out[0], stats[0], ref_out, ref_stats)


def test_emu_aec3_parity_synthetic_legs(emu_lib, oracle):
    """AEC3 only, three legs with different echo-path delays in ONE batched engine (the
    synthetic generator of SURVEY.md section 8d), 4 s: covers delay estimation, the
    alignment change, filter convergence and the exit from the initial state."""
    from common import run_legs, synthetic_leg
    ids = (1, 20, 47)
    legs = [synthetic_leg(i, 400) for i in ids]
    out, stats = run_legs(emu_lib, 16000, legs, stats_every=100, aec=True, ns=False)
    for k, (far, near) in enumerate(legs):
        ref_out, ref_stats, err = oracle.RefApm(aec=True, ns=False).run_i16(16000, far, near, stats_every=100)
        assert err == 0
        _check_aec(out[k], stats[k], ref_out, ref_stats)
    assert stats[:, -1, 2].min() > 0  # every leg found a non-zero render delay


def test_emu_render_before_first_capture_is_dropped_with_ns(emu_lib, oracle):
    """AudioProcessingImpl re-initialises on the first ProcessStream call when NS is enabled
    (audio_processing_impl.cc:558-559,894-925,1874-1881), so the first render frame never
    reaches AEC3; without NS it does.  The first two output frames expose the difference."""
    from common import run_legs, synthetic_leg
    far, near = synthetic_leg(0, 6)
    for ns in (False, True):
        ref_out, _, err = oracle.RefApm(aec=True, ns=ns, ns_level=1).run_i16(16000, far, near)
        assert err == 0
        out, _ = run_legs(emu_lib, 16000, [(far, near)], aec=True, ns=ns, ns_level=1)
        assert np.abs(out[0].astype(np.int32) - ref_out.astype(np.int32)).max() <= TOL_FS * 32768


def test_emu_aec3_echo_path_vanishes_loud_render(emu_lib, oracle):
    """Near-full-scale render whose echo path disappears after 3 s: drives the refined filter
    into the misadjustment rescale (subtractor.cc:246-257,345-375) and the coarse-filter
    re-seed (subtractor.cc:297-316)."""
    from common import run_legs, vanishing_echo_leg
    far, near = vanishing_echo_leg(450)
    ref_out, ref_stats, err = oracle.RefApm(aec=True, ns=False).run_i16(16000, far, near, stats_every=50)
    assert err == 0
    out, stats = run_legs(emu_lib, 16000, [(far, near)], stats_every=50, aec=True, ns=False)
    _check_aec(out[0], stats[0], ref_out, ref_stats)


def test_emu_aec3_ns_parity_48k_three_band(emu_lib, oracle):
    """48 kHz mono AEC3 + NS with maximum_internal_processing_rate = 48000: three-band split on both
    sides, AEC3 on band 0 with the upper-band gain / comfort noise / one-block delay
    (suppression_gain.cc:124-217, suppression_filter.cc:153-183), PostFilter after the merge."""
    from common import run_legs, synthetic_leg_48k
    sp = golden("speech_48k.npz")
    n = 150 * 480
    legs = [(sp["far"][:n], sp["near"][:n]), synthetic_leg_48k(11, 150, 3.5)]  # 2nd: clipped, HF-heavy render
    out, stats = run_legs(emu_lib, 48000, legs, stats_every=50, aec=True, ns=True, ns_level=1)
    for k, (far, near) in enumerate(legs):
        ref_out, ref_stats, err = oracle.RefApm(aec=True, ns=True, ns_level=1, max_rate=48000).run_i16(
            48000, far, near, stats_every=50)
        assert err == 0
        _check_aec(out[k], stats[k], ref_out, ref_stats)


def _speech_32k(n_frames):
    """The 16 kHz speech fixture held at 32 kHz (sample repetition): the images above 8 kHz give the
    upper band real content."""
    sp = golden("speech_16k.npz")
    n = n_frames * 160
    return np.repeat(sp["far"][:n], 2), np.repeat(sp["near"][:n], 2)


def test_emu_ns_parity_32k_two_band(emu_lib, oracle):
    """32 kHz: two-band QMF split / merge (splitting_filter.cc:68-101, signal_processing/
    splitting_filter.c:31-204), NS on band 0 and its upper-band gain on band 1, 32 kHz high-pass."""
    _, near = _speech_32k(150)
    ref_out, _, err = oracle.RefApm(aec=False, ns=True, ns_level=2).run_i16(32000, None, near)
    assert err == 0
    out = run_engine(emu_lib, 32000, None, near, n_streams=1, aec=False, ns=True, ns_level=2)
    d = np.abs(out.reshape(-1).astype(np.int32) - ref_out.astype(np.int32)).max()
    assert d <= TOL_FS * 32768, d


def test_emu_aec3_ns_parity_32k_two_band(emu_lib, oracle):
    """32 kHz mono AEC3 + NS: two bands on both sides, AEC3 on band 0 with one upper band
    (gain / comfort noise / one-block delay), no PostFilter (post_filter.cc:44-52)."""
    from common import run_legs, synthetic_leg_48k
    legs = [_speech_32k(150), synthetic_leg_48k(7, 150, 3.5, rate=32000)]  # 2nd: clipped, HF-heavy render
    out, stats = run_legs(emu_lib, 32000, legs, stats_every=50, aec=True, ns=True, ns_level=1)
    for k, (far, near) in enumerate(legs):
        ref_out, ref_stats, err = oracle.RefApm(aec=True, ns=True, ns_level=1).run_i16(32000, far, near, stats_every=50)
        assert err == 0
        _check_aec(out[k], stats[k], ref_out, ref_stats)
    # AEC3 alone (no re-initialisation-free path differences, no NS upper-band gain)
    far, near = legs[1]
    ref_out, _, err = oracle.RefApm(aec=True, ns=False).run_i16(32000, far, near)
    assert err == 0
    out, _ = run_legs(emu_lib, 32000, [legs[1]], aec=True, ns=False)
    assert np.abs(out[0].astype(np.int32) - ref_out.astype(np.int32)).max() <= TOL_FS * 32768


@pytest.mark.parametrize("rate,kw", [
    (48000, dict(aec=True, ns=True, ns_level=1, agc2=True, agc2_fixed_gain_db=6.0)),  # the reference's default
    (48000, dict(aec=False, ns=True, ns_level=2)),                                   # routing for 48 kHz clients
    (48000, dict(aec=False, ns=False, hpf=True)),
    (44100, dict(aec=True, ns=True, ns_level=1)),
    (44100, dict(aec=False, ns=False, agc2=True, agc2_fixed_gain_db=6.0)),            # no band split: 44.1 -> 48 kHz
    (8000, dict(aec=True, ns=True, ns_level=1)),
    (24000, dict(aec=False, ns=True, ns_level=1)),
])
def test_emu_resampled_rates(emu_lib, oracle, rate, kw):
    """API rate != processing rate (maximum_internal_processing_rate = 32000, the default):
    PushSincResampler on the way in (render and capture) and out, the 48 kHz
    capture_fullband_audio side path, the 48 kHz high-pass coefficients it selects for 32 kHz
    data (audio_processing_impl.cc:598-611,632-692,1451-1460,1892)."""
    from common import run_legs, synthetic_leg_48k
    legs = [synthetic_leg_48k(3 + k, 80, 2.0, rate=rate) for k in range(2)]
    out, stats = run_legs(emu_lib, rate, legs, stats_every=40 if kw["aec"] else 0, max_rate=32000, **kw)
    for k, (far, near) in enumerate(legs):
        ref_out, ref_stats, err = oracle.RefApm(max_rate=32000, **kw).run_i16(
            rate, far if kw["aec"] else None, near, stats_every=40 if kw["aec"] else 0)
        assert err == 0
        d = np.abs(out[k].astype(np.int32) - ref_out.astype(np.int32)).max()
        assert d <= TOL_FS * 32768, (k, d)
        assert np.abs(ref_out.astype(np.int32)).mean() > 1.0
        if kw["aec"]:
            assert np.abs(stats[k][:, 1] - ref_stats[:, 3]).max() <= 0.1


def test_emu_resampled_float_interface_and_mute(emu_lib, oracle):
    """48 kHz float frames under the default config: the float interface resamples before scaling
    to FloatS16 and scales back before the output resampler does not apply (fullband path); while
    the output is muted the fullband buffer keeps the unprocessed input and its resampler rests."""
    import wap_b200
    from common import synthetic_leg_48k
    far, near = synthetic_leg_48k(5, 60, 1.5, rate=48000)
    kw = dict(aec=True, ns=True, ns_level=1)
    eng = wap_b200.Engine(1, 48000, lib=emu_lib, max_rate=32000, **kw)
    ref = oracle.RefApm(max_rate=32000, **kw)
    worst = 0.0
    for f in range(60):
        if f == 20:
            eng.set_capture_output_used(False); ref.set_capture_output_used(False)
        if f == 35:
            eng.set_capture_output_used(True); ref.set_capture_output_used(True)
        c = (near[f * 480:(f + 1) * 480].astype(np.float32) / 32768.0).reshape(1, 480)
        r = (far[f * 480:(f + 1) * 480].astype(np.float32) / 32768.0).reshape(1, 480)
        eng.set_stream_delay_ms(0)
        o = eng.process(r, c)
        ro, err = ref.tick_f32(48000, r.reshape(-1), c.reshape(-1))
        assert err == 0
        worst = max(worst, float(np.abs(o.reshape(-1) - ro).max()))
        if 20 <= f < 35:
            assert np.array_equal(o.reshape(-1), c.reshape(-1))  # muted: the input comes back
    assert worst <= TOL_FS, worst
    eng.close()


@pytest.mark.parametrize("rate,gain_db", [(16000, 0.0), (16000, 12.0), (32000, 14.0), (48000, 20.0)])
def test_emu_agc2_fixed_gain_and_limiter(emu_lib, oracle, rate, gain_db):
    """GainController2, default sub-configuration (fixed digital gain + limiter) on its own
    (gain_controller2.cc:183-260, agc2/limiter.cc, fixed_digital_level_estimator.cc,
    interpolated_gain_curve.cc); at 48 kHz nothing is multi-band, so no band split either."""
    from common import loud_bursty_signal
    x = loud_bursty_signal(rate, 150)
    ref_out, _, err = oracle.RefApm(aec=False, ns=False, max_rate=48000, agc2=True,
                                    agc2_fixed_gain_db=gain_db).run_i16(rate, None, x)
    assert err == 0
    out = run_engine(emu_lib, rate, None, x, n_streams=1, aec=False, ns=False, agc2=True, agc2_fixed_gain_db=gain_db)
    assert np.abs(out.reshape(-1).astype(np.int32) - ref_out.astype(np.int32)).max() <= TOL_FS * 32768
    if gain_db > 0:
        assert np.abs(ref_out).max() >= 32000   # the limiter really worked


def test_emu_full_chain_aec3_ns_agc2(emu_lib, oracle):
    """BASELINE config 5 shape at test size: AEC3 + NS + AGC2 (fixed 6 dB + limiter), 16 kHz mono."""
    from common import run_legs, synthetic_leg
    legs = [synthetic_leg(i, 250) for i in (6, 29)]
    out, stats = run_legs(emu_lib, 16000, legs, stats_every=50, aec=True, ns=True, ns_level=1, agc2=True,
                          agc2_fixed_gain_db=6.0)
    for k, (far, near) in enumerate(legs):
        ref_out, ref_stats, err = oracle.RefApm(aec=True, ns=True, ns_level=1, agc2=True,
                                                agc2_fixed_gain_db=6.0).run_i16(16000, far, near, stats_every=50)
        assert err == 0
        _check_aec(out[k], stats[k], ref_out, ref_stats)


def test_emu_libm_restatement(emu_lib):
    """csrc/wap_libm.cuh restates glibc's powf(2, p) and tanhf (neither is correctly rounded, and the
    NS calls both): identical bits on 60 k random arguments each."""
    import sys
    sys.path.insert(0, os.path.join(ROOT, "tools"))
    import check_libm_restatement as chk
    for which in (0, 1, 2):
        bad, tot = chk.mismatches(emu_lib, which, 60000)
        assert bad == 0, (which, bad, tot)


@pytest.mark.parametrize("rate,max_rate,kw", [
    (16000, 32000, dict(aec=False, ns=True, ns_level=1)),
    (16000, 32000, dict(aec=True, ns=True, ns_level=2)),
    (32000, 32000, dict(aec=False, ns=True, ns_level=3)),
    (48000, 48000, dict(aec=False, ns=True, ns_level=1)),
    (48000, 32000, dict(aec=True, ns=True, ns_level=1, agc2=True, agc2_fixed_gain_db=6.0)),
    (44100, 32000, dict(aec=False, ns=True, ns_level=0)),
])
def test_emu_float_interface_is_bit_identical(emu_lib, oracle, rate, max_rate, kw):
    """Through the float interface nothing is rounded to int16 at the end, so this compares the
    float32 bits of every output sample: zero differing samples."""
    from common import float_interface_max_diff
    differing, worst = float_interface_max_diff(emu_lib, oracle, rate, 120, max_rate=max_rate, **kw)
    assert differing == 0, (differing, worst)


@pytest.mark.parametrize("rate,max_rate,right_gain,kw", [
    (16000, 32000, 0.6, dict(aec=True, ns=True, ns_level=1)),
    (16000, 32000, 4.0, dict(aec=True, ns=False)),                 # right channel clips, left does not
    (48000, 48000, 0.6, dict(aec=True, ns=True, ns_level=1)),
    (48000, 32000, 3.0, dict(aec=True, ns=True, ns_level=1, agc2=True, agc2_fixed_gain_db=6.0)),
    (44100, 32000, 0.6, dict(aec=True, ns=True, ns_level=2)),
])
def test_emu_stereo_default_pipeline(emu_lib, oracle, rate, max_rate, right_gain, kw):
    """Stereo frames with multi_channel_render / _capture off (the default) and AEC3 on: render is
    averaged to mono, capture continues with the first channel after AEC3's saturation test has seen
    both (audio_processing_impl.cc:585-594,1343,1365-1373), output on both channels."""
    from common import run_stereo_i16
    out, ref_out = run_stereo_i16(emu_lib, oracle, rate, 100, 9, max_rate=max_rate, right_gain=right_gain, **kw)
    assert np.array_equal(ref_out[0::2], ref_out[1::2])
    d = np.abs(out.astype(np.int32) - ref_out.astype(np.int32)).max()
    assert d <= TOL_FS * 32768, d
    if right_gain > 2 and rate == 16000:  # the saturation flag really differs from a left-only run
        left = np.repeat(__import__("common").stereo_leg(rate, 100, 9, right_gain)[1][0::2], 2)
        far = __import__("common").stereo_leg(rate, 100, 9, right_gain)[0]
        alt, _, _ = oracle.RefApm(max_rate=max_rate, **kw).run_i16(rate, far, left, render_ch=2, capture_ch=2)
        assert not np.array_equal(alt, ref_out)


def test_emu_unbuilt_channel_layouts_are_refused(emu_lib):
    """Stereo without AEC3 is built (tests/test_multichannel.py) at 16 / 48 kHz native; through the resamplers,
    at 32 kHz, or with more than two channels it is refused, not approximated."""
    import wap_b200
    with pytest.raises(RuntimeError):
        wap_b200.Engine(1, 32000, channels=2, lib=emu_lib, aec=False, ns=True)
    with pytest.raises(RuntimeError):
        wap_b200.Engine(1, 48000, channels=2, lib=emu_lib, aec=False, ns=True, max_rate=32000)
    with pytest.raises(RuntimeError):
        wap_b200.Engine(1, 48000, channels=2, lib=emu_lib, aec=False, ns=False, hpf=False, agc2=True)
    with pytest.raises(RuntimeError):
        wap_b200.Engine(1, 16000, channels=3, lib=emu_lib, aec=True, ns=True)


LEVEL_EVENTS = [(10, "pre_gain", 2.5), (20, "post_gain", 0.5), (30, "playout_volume", 100), (40, "playout_volume", 180),
                (50, "pre_gain", 0.7), (60, "post_gain", 3.0)]
AGC2_EVENTS = [(8, "fixed_post_gain", 12.0), (20, "fixed_post_gain", 12.0), (30, "fixed_post_gain", -3.0),
               (50, "fixed_post_gain", 20.0)]


@pytest.mark.parametrize("rate,max_rate,events,kw", [
    (16000, 32000, LEVEL_EVENTS, dict(aec=True, ns=True, ns_level=1, pre_gain=1.5, post_gain=0.8)),
    (16000, 32000, LEVEL_EVENTS, dict(aec=True, ns=False, pre_amp=3.0)),
    (48000, 48000, LEVEL_EVENTS, dict(aec=True, ns=True, ns_level=1, pre_gain=1.0, post_gain=1.0)),
    (48000, 32000, LEVEL_EVENTS, dict(aec=True, ns=True, ns_level=1, agc2=True, agc2_fixed_gain_db=3.0,
                                      pre_amp=2.0, pre_gain=1.2, post_gain=1.1)),
    (32000, 32000, LEVEL_EVENTS, dict(aec=False, ns=True, ns_level=2, pre_gain=4.0, post_gain=0.9)),
    (16000, 32000, LEVEL_EVENTS, dict(aec=True, ns=True, ns_level=1)),      # only the playout volume acts
    (16000, 32000, AGC2_EVENTS, dict(aec=False, ns=False, agc2=True, agc2_fixed_gain_db=0.0)),
    (48000, 32000, AGC2_EVENTS, dict(aec=True, ns=True, ns_level=1, agc2=True, agc2_fixed_gain_db=6.0)),
])
def test_emu_level_adjustment_and_runtime_settings(emu_lib, oracle, rate, max_rate, events, kw):
    """Pre-amplifier / capture level adjustment (ramped pre and post gains, clamp), the runtime
    settings kCapturePreGain / kCapturePostGain / kCaptureFixedPostGain / kPlayoutVolumeChange, and
    the echo-path gain-change flag AEC3 gets from them (audio_processing_impl.cc:970-1038,1289-1341,
    1526-1528; gain_controller2.cc:160-168): float32 bits identical."""
    from common import run_with_runtime_settings
    assert run_with_runtime_settings(emu_lib, oracle, rate, 80, events, max_rate=max_rate, **kw) == 0


def test_bench_reference_arm_contract(oracle):
    """`bench.py --impl reference` (the CPU arm the driver times beside the GPU arm): exactly one JSON
    line on stdout with the contract's keys; no CUDA device needed."""
    import json
    import subprocess
    import sys
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1",
                        "--warmup", "1", "--cpu-seconds", "1.5"], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [l for l in r.stdout.splitlines() if l.strip()]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["unit"] == "real-time streams" and d["higher_is_better"] is True
    assert d["value"] > 0 and d["gpu_launches"] == 0 and d["vs_baseline"] is None
    assert d["cpu_baseline"]["kind"] == "reference" and d["cpu_baseline"]["cores"] >= 1
    assert d["e2e"]["value"] == d["value"] and d["e2e"]["h2d_bytes_per_step"] == 0
    assert "workload" in d["config"]


def test_synthetic_generator_c_equals_numpy_restatement():
    """tools/wap_synth.c (what both bench arms use) against the numpy restatement of the same
    SURVEY 8(d) generator (xorshift64* as webrtc::Random, gated render, circular 3-tap echo path)."""
    import synth
    if synth._load() is None:
        pytest.skip("no C compiler / prebuilt generator")
    for kind, rate in ((0, 16000), (0, 48000), (1, 48000), (2, 16000), (2, 48000)):
        a = synth.cycle(kind, rate, 37, 5, 110)
        b = synth.cycle(kind, rate, 37, 5, 110, force_numpy=True)
        assert np.array_equal(a[0], b[0]) and np.array_equal(a[1], b[1])
    r, c = synth.cycle(0, 16000, 0, 2, 200)
    r = r.transpose(1, 0, 2).reshape(2, -1)
    assert np.all(r[:, 14400:16000] == 0) and np.abs(r[:, :14400]).max() > 7000   # 0.9 s on / 0.1 s off
    assert len(np.unique(r[0, :1000])) > 900 and not np.array_equal(r[0], r[1])
    # webrtc::Random known answers: first outputs of seed 1000 (xorshift64*, shifts 12/25/27)
    v = synth._VecRandom([1000])
    s = 1000
    s ^= s >> 12; s ^= (s << 25) & (2**64 - 1); s ^= s >> 27
    assert int(v.s[0]) == 1000 and abs(float(v.rand_float()[0]) - ((s * 2685821657736338717) % 2**64 - 1) / (2**64 - 1)) < 1e-7
