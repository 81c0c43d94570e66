"""Multi-channel AEC3 (BASELINE config 4): stereo frames with
pipeline.multi_channel_render and _capture on.  The reference then runs EchoCanceller3 with two capture
channels and one (downmixed) or two render channels, switching when MultiChannelContentDetector sees
persistent stereo content; the engine must follow it bit for bit: int16 output identical, float output
identical in every bit.  Runs on the CPU warp emulator and (marked gpu) on the product library."""
import numpy as np
import pytest

from common import stereo_leg


@pytest.fixture(params=["emu", pytest.param("gpu", marks=pytest.mark.gpu)])
def api_lib(request):
    return request.getfixturevalue("emu_lib" if request.param == "emu" else "gpu_lib")


MC = dict(aec=True, ns=False, mc_render=True, mc_capture=True, max_rate=48000)


def _kw_without_ns(kw):
    """engine kwargs on top of MC (which already fixes ns=False): an `ns` entry replaces it."""
    return kw


def render_variant(far, kind, rate):
    """far: interleaved stereo render.  'stereo': as is (persistent stereo content after 2 s); 'mono': both
    channels identical (the AEC keeps one downmixed render channel); 'burst': 0.5 s of stereo content, then
    identical channels (temporary stereo: average downmix + echo_path_gain_change, never persistent)."""
    far = far.copy()
    if kind == "mono":
        far[1::2] = far[0::2]
    elif kind == "burst":
        n = rate // 100 * 2 * 50
        far[n + 1::2] = far[n::2]
    return far


def run_pair(lib, oracle, rate, n_frames, kind, seed=9, right_gain=0.6, delay_ms=0, engine_kw=None, ref_kv=None):
    import wap_b200
    far, near = stereo_leg(rate, n_frames, seed, right_gain)
    far = render_variant(far, kind, rate)
    fl = rate // 100 * 2
    ekw = dict(MC)
    ekw.update(engine_kw or {})
    eng = wap_b200.Engine(1, rate, channels=2, lib=lib, **ekw)
    if ref_kv is None:
        refapm = oracle.RefApm(**MC)
    else:
        kv = dict(aec=1, ns=0, mc_render=1, mc_capture=1, max_rate=48000)
        kv.update(ref_kv)
        refapm = oracle.RefApm(kv=kv)
    ref_out, _, err = refapm.run_i16(rate, far, near, render_ch=2, capture_ch=2)
    assert err == 0
    out = np.zeros_like(near)
    for f in range(n_frames):
        eng.set_stream_delay_ms(delay_ms)
        out[f * fl:(f + 1) * fl] = eng.process(far[f * fl:(f + 1) * fl].reshape(1, fl),
                                               near[f * fl:(f + 1) * fl].reshape(1, fl)).reshape(-1)
    eng.close()
    return out, ref_out


def first_bad_frame(out, ref_out, fl):
    d = np.abs(out.astype(np.int32) - ref_out.astype(np.int32)).reshape(-1, fl).max(1)
    bad = np.nonzero(d)[0]
    return (int(bad[0]), int(d.max())) if bad.size else None


@pytest.mark.parametrize("rate,n_frames,kind", [
    (16000, 420, "stereo"),   # crosses the detector's 2 s hysteresis: Initialize() with 2 render channels + multichannel config
    (16000, 150, "mono"),
    (16000, 150, "burst"),
    (48000, 300, "stereo"),
    (48000, 100, "mono"),
    (48000, 100, "burst"),
])
def test_multichannel_matches_reference(api_lib, oracle, rate, n_frames, kind):
    out, ref_out = run_pair(api_lib, oracle, rate, n_frames, kind)
    assert first_bad_frame(out, ref_out, rate // 100 * 2) is None
    # the second channel is really processed on its own (not a copy of the first)
    assert np.any(out[0::2] != out[1::2])


@pytest.mark.parametrize("rate,n_frames,kind,level", [
    (16000, 330, "stereo", 1),    # NS moderate, crosses the switch to two render channels and NS's 200-frame start-up
    (16000, 120, "mono", 3),      # NS very high
    (48000, 260, "stereo", 2),    # NS high, three bands: upper-band gains and delays per channel, minimum over channels
    (48000, 90, "burst", 0),      # NS low
])
def test_multichannel_with_noise_suppressor(api_lib, oracle, rate, n_frames, kind, level):
    """NoiseSuppressor with two channels (noise_suppressor.cc:278-292,294-559): per-channel state, one frame
    counter and zero-frame test, Wiener filter / gain adjustment / upper-band gain = minimum over the channels."""
    import wap_b200
    far, near = stereo_leg(rate, n_frames, 13, 0.6)
    far = render_variant(far, kind, rate)
    fl = rate // 100 * 2
    kw = dict(MC, ns=True, ns_level=level)
    eng = wap_b200.Engine(1, rate, channels=2, lib=api_lib, **kw)
    ref_out, _, err = oracle.RefApm(**kw).run_i16(rate, far, near, render_ch=2, capture_ch=2)
    assert err == 0
    out = np.zeros_like(near)
    for f in range(n_frames):
        eng.set_stream_delay_ms(0)
        out[f * fl:(f + 1) * fl] = eng.process(far[f * fl:(f + 1) * fl].reshape(1, fl), near[f * fl:(f + 1) * fl].reshape(1, fl)).reshape(-1)
    eng.close()
    assert first_bad_frame(out, ref_out, fl) is None


def test_multichannel_noise_suppressor_muted_output_and_silence(api_lib, oracle):
    """Silent stretches (the zero-frame rule looks at every channel) and a muted output (Process stops after the
    per-channel filter update) with NS on two channels, float interface."""
    import wap_b200
    rate, n_frames, fl = 16000, 160, 160
    far, near = stereo_leg(rate, n_frames, 17, 0.6)
    near = near.copy()
    near[0:2 * fl * 6] = 0                      # six silent frames at the start (both channels)
    near[2 * fl * 40:2 * fl * 44:2] = 0         # left channel silent, right one not
    kw = dict(MC, ns=True, ns_level=1)
    eng = wap_b200.Engine(1, rate, channels=2, lib=api_lib, **kw)
    refapm = oracle.RefApm(**kw)
    differing = 0
    for f in range(n_frames):
        if f in (60, 90):
            eng.set_capture_output_used(f == 90)
            refapm.set_capture_output_used(f == 90)
        r = (far[f * 2 * fl:(f + 1) * 2 * fl].reshape(fl, 2).T.astype(np.float32) / 32768.0).copy()
        c = (near[f * 2 * fl:(f + 1) * 2 * fl].reshape(fl, 2).T.astype(np.float32) / 32768.0).copy()
        eng.set_stream_delay_ms(0)
        o = eng.process(r.reshape(1, -1), c.reshape(1, -1)).reshape(-1)
        ro, err = refapm.tick_f32(rate, r.reshape(-1), c.reshape(-1), render_ch=2, capture_ch=2)
        assert err == 0
        differing += int(np.count_nonzero(o.view(np.uint32) != ro.view(np.uint32)))
    eng.close()
    assert differing == 0


def test_multichannel_saturated_capture_channel(api_lib, oracle):
    # the right capture channel clips: AnalyzeCapture's saturation flag covers every channel
    out, ref_out = run_pair(api_lib, oracle, 16000, 260, "stereo", right_gain=2.6)
    assert first_bad_frame(out, ref_out, 320) is None


def test_multichannel_float_interface_and_statistics(api_lib, oracle):
    import wap_b200
    rate, n_frames = 16000, 330
    far, near = stereo_leg(rate, n_frames, 11, 0.6)
    fl = rate // 100
    eng = wap_b200.Engine(1, rate, channels=2, lib=api_lib, **MC)
    refapm = oracle.RefApm(**MC)
    differing = 0
    for f in range(n_frames):
        # planar float frames [channel][sample]
        r = (far[f * 2 * fl:(f + 1) * 2 * fl].reshape(fl, 2).T.astype(np.float32) / 32768.0).copy()
        c = (near[f * 2 * fl:(f + 1) * 2 * fl].reshape(fl, 2).T.astype(np.float32) / 32768.0).copy()
        eng.set_stream_delay_ms(0)
        o = eng.process(r.reshape(1, -1), c.reshape(1, -1)).reshape(-1)
        ro, err = refapm.tick_f32(rate, r.reshape(-1), c.reshape(-1), render_ch=2, capture_ch=2)
        assert err == 0
        differing += int(np.count_nonzero(o.view(np.uint32) != ro.view(np.uint32)))
        if f % 110 == 109:
            ours, theirs = eng.stats(0), refapm.stats()
            # (has_erl, erl, has_erle, erle, has_delay, delay_ms)
            assert ours.has_echo_return_loss and theirs[0] and ours.has_echo_return_loss_enhancement and theirs[2]
            assert ours.echo_return_loss == pytest.approx(float(theirs[1]), abs=1e-4)
            assert ours.echo_return_loss_enhancement == pytest.approx(float(theirs[3]), abs=1e-4)
            assert ours.delay_ms == int(theirs[5])
    eng.close()
    assert differing == 0


def test_multichannel_batch_legs_switch_independently(api_lib, oracle):
    """Three legs in one batch: stereo content, identical channels, a stereo burst.  Each leg's content
    detector re-initialises its own state; the others must not notice."""
    import wap_b200
    rate, n_frames, fl = 16000, 260, 320
    kinds = ["stereo", "mono", "burst"]
    legs = []
    for i, kind in enumerate(kinds):
        far, near = stereo_leg(rate, n_frames, 20 + i, 0.6)
        legs.append((render_variant(far, kind, rate), near))
    eng = wap_b200.Engine(len(kinds), rate, channels=2, lib=api_lib, **MC)
    outs = [np.zeros(n_frames * fl, np.int16) for _ in kinds]
    for f in range(n_frames):
        eng.set_stream_delay_ms(0)
        r = np.stack([l[0][f * fl:(f + 1) * fl] for l in legs])
        c = np.stack([l[1][f * fl:(f + 1) * fl] for l in legs])
        o = eng.process(r, c)
        for i in range(len(kinds)):
            outs[i][f * fl:(f + 1) * fl] = o[i]
    eng.close()
    for i, kind in enumerate(kinds):
        ref_out, _, err = oracle.RefApm(**MC).run_i16(rate, legs[i][0], legs[i][1], render_ch=2, capture_ch=2)
        assert err == 0
        assert first_bad_frame(outs[i], ref_out, fl) is None, kind


def test_multichannel_custom_configs(api_lib, oracle):
    """A mono config and a multichannel config of the user's own (SetEchoCancellerConfig(config,
    multichannel_config)): shorter detector hysteresis so that the switch happens early, a fixed-downmix
    capture mixer, other filter lengths for the multichannel AEC."""
    mono = {"multi_channel.stereo_detection_hysteresis_seconds": 0.5, "filter.refined.length_blocks": 11,
            "filter.refined_initial.length_blocks": 10}
    mc = {"filter.refined.length_blocks": 12, "filter.coarse.length_blocks": 9, "filter.coarse_initial.length_blocks": 8,
          "delay.capture_alignment_mixing.downmix": True, "delay.capture_alignment_mixing.adaptive_selection": False,
          "suppressor.normal_tuning.max_inc_factor": 1.7}
    ref_kv = {"ec3." + k: v for k, v in mono.items()}
    ref_kv.update({"ec3mc." + k: v for k, v in mc.items()})
    out, ref_out = run_pair(api_lib, oracle, 16000, 200, "stereo", engine_kw=dict(aec3=mono, aec3_multichannel=mc), ref_kv=ref_kv)
    assert first_bad_frame(out, ref_out, 320) is None


def test_multichannel_without_stereo_detection(api_lib, oracle):
    # detect_stereo_content off: two render channels and the multichannel config from the first frame
    mono = {"multi_channel.detect_stereo_content": False}
    ref_kv = {"ec3.multi_channel.detect_stereo_content": 0, "ec3mc.multi_channel.detect_stereo_content": 0}
    out, ref_out = run_pair(api_lib, oracle, 16000, 120, "mono", engine_kw=dict(aec3=mono, aec3_multichannel=mono), ref_kv=ref_kv)
    assert first_bad_frame(out, ref_out, 320) is None


def test_multichannel_state_moves_between_engines(api_lib, oracle):
    """Export a multi-channel leg after the switch to two render channels and continue it on another engine."""
    import wap_b200
    rate, n_frames, fl, cut = 16000, 300, 320, 240
    far, near = stereo_leg(rate, n_frames, 31, 0.6)
    ref_out, _, err = oracle.RefApm(**MC).run_i16(rate, far, near, render_ch=2, capture_ch=2)
    assert err == 0
    a = wap_b200.Engine(1, rate, channels=2, lib=api_lib, **MC)
    b = wap_b200.Engine(1, rate, channels=2, lib=api_lib, **MC)
    out = np.zeros_like(near)
    eng = a
    for f in range(n_frames):
        if f == cut:
            b.import_state(a.export_state(0), 0)
            eng = b
        eng.set_stream_delay_ms(0)
        out[f * fl:(f + 1) * fl] = eng.process(far[f * fl:(f + 1) * fl].reshape(1, fl), near[f * fl:(f + 1) * fl].reshape(1, fl)).reshape(-1)
    a.close()
    b.close()
    assert first_bad_frame(out, ref_out, fl) is None


def test_multichannel_unsupported_combinations_are_refused(api_lib):
    import wap_b200
    for kw in (dict(MC, mc_capture=False), dict(MC, mc_render=False),
               dict(MC, max_rate=32000)):
        rate = 48000 if kw.get("max_rate") == 32000 else 16000
        with pytest.raises(RuntimeError):
            wap_b200.Engine(1, rate, channels=2, lib=api_lib, **kw)


def test_multichannel_single_leg_entry_points(api_lib, oracle):
    """The seam's own calls (wap_create_with_aec3_config + ProcessReverseStream / ProcessStream) on stereo
    frames with both multi_channel flags: same engine underneath, user configs for both variants."""
    import ctypes as C
    import wap_b200
    L = api_lib
    rate, n_frames, fl = 16000, 150, 320
    mono = {"multi_channel.stereo_detection_hysteresis_seconds": 0.3}
    mc = {"filter.coarse.length_blocks": 10, "filter.coarse_initial.length_blocks": 9}
    far, near = stereo_leg(rate, n_frames, 41, 0.6)
    kv = dict(aec=1, ns=0, mc_render=1, mc_capture=1, max_rate=48000)
    kv.update({"ec3." + k: v for k, v in mono.items()})
    kv.update({"ec3mc." + k: v for k, v in mc.items()})
    ref_out, _, err = oracle.RefApm(kv=kv).run_i16(rate, far, near, render_ch=2, capture_ch=2)
    assert err == 0
    cfg = wap_b200.make_config(L, **MC)
    c_mono, c_mc = wap_b200.make_aec3_config(L, mono), wap_b200.make_aec3_config(L, mc, multichannel=True)
    h = L.wap_create_with_aec3_config(cfg, C.byref(c_mono), C.byref(c_mc))
    assert h
    sc = wap_b200.WapStreamConfig(rate, 2)
    out = np.zeros_like(near)
    tmp = np.zeros(fl, np.int16)
    for f in range(n_frames):
        r = np.ascontiguousarray(far[f * fl:(f + 1) * fl])
        c = np.ascontiguousarray(near[f * fl:(f + 1) * fl])
        p = lambda a: a.ctypes.data_as(C.c_void_p)
        assert L.wap_process_reverse_stream_i16(h, p(r), fl, sc, sc, p(tmp), fl) == 0
        L.wap_set_stream_delay_ms(h, 0)
        assert L.wap_process_stream_i16(h, p(c), fl, sc, sc, p(tmp), fl) == 0
        out[f * fl:(f + 1) * fl] = tmp
    L.wap_destroy(h)
    assert first_bad_frame(out, ref_out, fl) is None


@pytest.mark.parametrize("rate,n_frames,ns,gain_db", [
    (16000, 260, False, 30.0),   # +30 dB: the limiter works on the loud passages of both channels
    (48000, 230, True, 27.0),    # three bands, NS on both channels, AGC2 in front of the PostFilter
])
def test_multichannel_with_agc2_fixed_gain_and_limiter(api_lib, oracle, rate, n_frames, ns, gain_db):
    """GainController2 (fixed digital gain + limiter) on multi-channel legs: one level estimate over the
    channels (maximum envelope), one set of scaling factors for both."""
    kw = dict(agc2=True, agc2_fixed_gain_db=gain_db)
    ref = dict(agc2=1, agc2_gain_db=gain_db)
    if ns:
        kw.update(ns=True, ns_level=1)
        ref.update(ns=1, ns_level=1)
    out, ref_out = run_pair(api_lib, oracle, rate, n_frames, "stereo", seed=5, engine_kw=_kw_without_ns(kw), ref_kv=ref)
    assert first_bad_frame(out, ref_out, rate // 100 * 2) is None
    assert np.abs(ref_out.astype(np.int32)).max() > 15000      # loud enough for the limiter to matter


@pytest.mark.parametrize("rate,n_frames", [(16000, 200), (48000, 160)])
def test_multichannel_level_adjustment_and_runtime_settings(api_lib, oracle, rate, n_frames):
    """CaptureLevelsAdjuster on multi-channel legs (pre gain behind the high-pass filters, post gain behind AGC2 /
    the PostFilter, ramped, the same for both channels) and the echo-path gain-change flag a changed pre gain or
    playout volume raises -- with runtime changes in the middle of the call."""
    import wap_b200
    far, near = stereo_leg(rate, n_frames, 7, 0.6)
    fl = rate // 100 * 2
    kw = dict(MC)
    kw.update(pre_gain=1.5, post_gain=0.8, agc2=True, agc2_fixed_gain_db=3.0)
    eng = wap_b200.Engine(1, rate, channels=2, lib=api_lib, **kw)
    ref = oracle.RefApm(kv=dict(aec=1, ns=0, mc_render=1, mc_capture=1, max_rate=48000, cla=1, cla_pre=1.5, cla_post=0.8,
                                agc2=1, agc2_gain_db=3.0))
    events = {40: ("pre_gain", 2.5), 70: ("playout_volume", 90), 71: ("playout_volume", 140), 100: ("post_gain", 1.7),
              130: ("pre_gain", 0.6)}
    out = np.zeros_like(near)
    ref_out = np.zeros_like(near)
    for f in range(n_frames):
        if f in events:
            what, val = events[f]
            getattr(eng, "set_" + what)(val)
            getattr(ref, "set_" + what)(val)
        sl = slice(f * fl, (f + 1) * fl)
        eng.set_stream_delay_ms(0)
        out[sl] = eng.process(far[sl].reshape(1, fl), near[sl].reshape(1, fl)).reshape(-1)
        ro, _, err = ref.run_i16(rate, far[sl], near[sl], render_ch=2, capture_ch=2)
        assert err == 0
        ref_out[sl] = ro
    eng.close()
    assert first_bad_frame(out, ref_out, fl) is None


@pytest.mark.parametrize("rate,n_frames,kw,ref", [
    (16000, 260, dict(ns=True, ns_level=1), dict(ns=1, ns_level=1)),
    (48000, 240, dict(ns=True, ns_level=2, agc2=True, agc2_fixed_gain_db=24.0), dict(ns=1, ns_level=2, agc2=1, agc2_gain_db=24.0)),
    (16000, 120, dict(ns=False, hpf=True, agc2=True, agc2_fixed_gain_db=30.0, pre_gain=1.3, post_gain=0.9),
     dict(ns=0, hpf=1, agc2=1, agc2_gain_db=30.0, cla=1, cla_pre=1.3, cla_post=0.9)),
])
def test_stereo_without_echo_canceller(api_lib, oracle, rate, n_frames, kw, ref):
    """Stereo capture without AEC3: the reference keeps both channels (the reduction to one channel only happens
    next to an echo controller) -- high-pass filter per channel, noise suppressor with its minima over the
    channels, AGC2 with one level estimate, level adjustment -- whatever the pipeline flags say."""
    import wap_b200
    _, near = stereo_leg(rate, n_frames, 11, 0.7)
    fl = rate // 100 * 2
    for flags in (dict(mc_render=False, mc_capture=False), dict(mc_render=True, mc_capture=True)):
        ekw = dict(aec=False, max_rate=48000)
        ekw.update(kw)
        ekw.update(flags)
        eng = wap_b200.Engine(1, rate, channels=2, lib=api_lib, **ekw)
        rkv = dict(aec=0, max_rate=48000, mc_render=int(flags["mc_render"]), mc_capture=int(flags["mc_capture"]))
        rkv.update(ref)
        ref_out, _, err = oracle.RefApm(kv=rkv).run_i16(rate, None, near, render_ch=2, capture_ch=2)
        assert err == 0
        out = np.zeros_like(near)
        for f in range(n_frames):
            out[f * fl:(f + 1) * fl] = eng.process(None, near[f * fl:(f + 1) * fl].reshape(1, fl)).reshape(-1)
        eng.close()
        assert first_bad_frame(out, ref_out, fl) is None, flags
        assert np.any(out[0::2] != out[1::2])
