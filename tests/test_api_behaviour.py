"""The seam's contract exercised on the emulator build of the product sources and on the GPU build:
error codes, delay clamp, mute / un-mute, single-leg entry points, int16 / float equivalence.
Mirrors the reference's APM API tests (tests/unit/audio_processing_unittest.cc,
audio_processing_impl_unittest.cc) for the part of the surface the seam forwards."""
import ctypes as C
import os

import numpy as np
import pytest

from common import golden, synthetic_leg


@pytest.fixture(params=["emu", pytest.param("gpu", marks=pytest.mark.gpu)])
def api_lib(request):
    """Every test of this file runs twice: on the emulator build of the product sources (CPU box) and,
    marked gpu, on libwap_b200.so itself through the same C ABI."""
    return request.getfixturevalue("emu_lib" if request.param == "emu" else "gpu_lib")


ROOT = os.path.normpath(os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
TOL = 1e-4 * 32768


def _sc(wap_b200, rate, ch=1):
    return wap_b200.WapStreamConfig(rate, ch)


def test_error_codes_and_delay_clamp(api_lib):
    import wap_b200
    L = api_lib
    h = L.wap_create()
    assert h
    x = np.zeros(160, np.int16)
    y = np.zeros(160, np.int16)
    p = lambda a: a.ctypes.data_as(C.c_void_p)
    ok = _sc(wap_b200, 16000)
    # audio_processing_impl.cc:163-323 through rust_audio_processing.cc:21-40
    assert L.wap_process_stream_i16(h, p(x), 160, _sc(wap_b200, 7000), _sc(wap_b200, 7000), p(y), 160) == 3   # BadSampleRate
    assert L.wap_process_stream_i16(h, p(x), 160, _sc(wap_b200, 16000, 0), ok, p(y), 160) == 4               # BadNumberChannels
    assert L.wap_process_stream_i16(h, p(x), 100, ok, ok, p(y), 160) == 6                                     # BadDataLength
    assert L.wap_process_stream_i16(None, p(x), 160, ok, ok, p(y), 160) == 1                                  # NullPointer
    # set_stream_delay_ms clamps to [0, 500] and reports the warning (audio_processing_impl.cc:1689-1707)
    assert L.wap_set_stream_delay_ms(h, 20) == 0 and L.wap_stream_delay_ms(h) == 20
    assert L.wap_set_stream_delay_ms(h, -5) == 5 and L.wap_stream_delay_ms(h) == 0
    assert L.wap_set_stream_delay_ms(h, 900) == 5 and L.wap_stream_delay_ms(h) == 500
    # statistics are empty before the first capture frame (ApmStatsReporter::cached_stats_)
    st = wap_b200.WapStats()
    assert L.wap_get_statistics(h, C.byref(st)) == 0
    assert not st.has_echo_return_loss and not st.has_delay_ms
    L.wap_destroy(h)


def test_single_leg_entry_points_match_the_reference(api_lib, oracle):
    """wap_process_reverse_stream_i16 + wap_set_stream_delay_ms + wap_process_stream_i16 on one
    handle (what RustAudioProcessing forwards) == the reference, including GetStatistics."""
    import wap_b200
    L = api_lib
    far, near = synthetic_leg(9, 120)
    ref_out, ref_stats, err = oracle.RefApm(aec=True, ns=True, ns_level=1).run_i16(16000, far, near, stats_every=40)
    assert err == 0
    cfg = wap_b200.make_config(L, aec=True, ns=True, ns_level=1)
    h = L.wap_create_with_config(cfg)
    sc = _sc(wap_b200, 16000)
    out = np.zeros_like(near)
    scratch = np.zeros(160, np.int16)
    p = lambda a: a.ctypes.data_as(C.c_void_p)
    stats = []
    for f in range(120):
        r = np.ascontiguousarray(far[f * 160:(f + 1) * 160])
        c = np.ascontiguousarray(near[f * 160:(f + 1) * 160])
        o = np.zeros(160, np.int16)
        assert L.wap_process_reverse_stream_i16(h, p(r), 160, sc, sc, p(scratch), 160) == 0
        assert np.array_equal(scratch, r)          # render passes through unchanged
        assert L.wap_set_stream_delay_ms(h, 0) == 0
        assert L.wap_process_stream_i16(h, p(c), 160, sc, sc, p(o), 160) == 0
        out[f * 160:(f + 1) * 160] = o
        if (f + 1) % 40 == 0:
            st = wap_b200.WapStats()
            L.wap_get_statistics(h, C.byref(st))
            stats.append((st.echo_return_loss, st.echo_return_loss_enhancement, st.delay_ms))
    L.wap_destroy(h)
    assert np.abs(out.astype(np.int32) - ref_out.astype(np.int32)).max() <= TOL
    stats = np.array(stats)
    assert np.abs(stats[:, 1] - ref_stats[:, 3]).max() <= 0.1
    assert np.array_equal(stats[:, 2], ref_stats[:, 5])


def test_float_and_int16_entry_points_agree(api_lib):
    import wap_b200
    far, near = synthetic_leg(4, 60)
    outs = []
    for dt in (np.int16, np.float32):
        eng = wap_b200.Engine(1, 16000, lib=api_lib, aec=True, ns=True, ns_level=1)
        o = []
        for f in range(60):
            r = far[f * 160:(f + 1) * 160][None, :]
            c = near[f * 160:(f + 1) * 160][None, :]
            if dt is np.float32:
                r, c = (r / 32768.0).astype(np.float32), (c / 32768.0).astype(np.float32)
            eng.set_stream_delay_ms(0)
            y = eng.process(r, c)[0]
            o.append(y if dt is np.int16 else y * 32768.0)
        eng.close()
        outs.append(np.concatenate(o).astype(np.float64))
    # FloatS16ToS16 rounds; the float path keeps the fraction
    assert np.abs(outs[0] - outs[1]).max() <= 0.5 + 1e-3


def test_mute_unmute_matches_reference(api_lib, oracle):
    """set_output_will_be_muted: AEC3 skips the suppressor, NS its synthesis, and the first frame
    after un-muting is zeroed (audio_processing_impl.cc:818-838,1450,1540-1552)."""
    import wap_b200
    far, near = synthetic_leg(2, 90)
    ref = oracle.RefApm(aec=True, ns=True, ns_level=1)
    eng = wap_b200.Engine(2, 16000, lib=api_lib, aec=True, ns=True, ns_level=1)
    ref_out = np.zeros_like(near)
    out = np.zeros((2, near.size), np.int16)
    for f in range(90):
        if f == 30:
            ref.set_capture_output_used(False)
            eng.set_capture_output_used(False, legs=[0])   # leg 1 stays un-muted
        if f == 60:
            ref.set_capture_output_used(True)
            eng.set_capture_output_used(True, legs=[0])
        sl = slice(f * 160, (f + 1) * 160)
        o, _, err = ref.run_i16(16000, far[sl], near[sl])
        assert err == 0
        ref_out[sl] = o
        eng.set_stream_delay_ms(0)
        out[:, sl] = eng.process(np.stack([far[sl]] * 2), np.stack([near[sl]] * 2))
    eng.close()
    assert np.abs(out[0].astype(np.int32) - ref_out.astype(np.int32)).max() <= TOL
    assert np.all(out[0][60 * 160:61 * 160] == 0)             # zeroed frame after un-muting
    assert not np.array_equal(out[0], out[1])                  # the flag is per leg


def test_legs_joining_later_and_slot_reuse(api_lib, oracle):
    """Ragged batches: legs created on different ticks run different 2/3-block cadences inside the
    same launch; a destroyed leg's arena slot is re-initialised for the next leg."""
    import wap_b200
    L = api_lib
    nf = 50
    legs = [synthetic_leg(i, nf) for i in range(3)]
    refs = []
    for far, near in legs:
        o, _, err = oracle.RefApm(aec=True, ns=True, ns_level=1).run_i16(16000, far, near)
        assert err == 0
        refs.append(o)
    eng = wap_b200.Engine(1, 16000, lib=L, capacity=2, aec=True, ns=True, ns_level=1)
    p = lambda a: a.ctypes.data_as(C.c_void_p)
    hs = [eng.handles[0], None]
    out = [np.zeros(nf * 160, np.int16) for _ in range(3)]
    start = {0: 0, 1: 7, 2: 31}   # leg 1 joins at tick 7; leg 0 leaves at tick 30, leg 2 takes its slot
    live = {0: hs[0]}
    for t in range(nf + 31):
        if t == 7:
            h1 = (C.c_void_p * 1)()
            assert L.wap_engine_create_streams(eng.h, 1, h1) == 0
            live[1] = h1[0]
            # capacity is 2: a third simultaneous leg is refused
            h_extra = (C.c_void_p * 1)()
            assert L.wap_engine_create_streams(eng.h, 1, h_extra) == 5
        if t == 30:
            L.wap_destroy(live.pop(0))
        if t == 31:
            h2 = (C.c_void_p * 1)()
            assert L.wap_engine_create_streams(eng.h, 1, h2) == 0
            live[2] = h2[0]
        ids = [i for i in sorted(live) if 0 <= t - start[i] < nf]
        if not ids:
            continue
        hh = (C.c_void_p * len(ids))(*[live[i] for i in ids])
        r = np.stack([legs[i][0][(t - start[i]) * 160:(t - start[i] + 1) * 160] for i in ids])
        c = np.stack([legs[i][1][(t - start[i]) * 160:(t - start[i] + 1) * 160] for i in ids])
        o = np.zeros_like(c)
        assert L.wap_streams_set_delay_ms(hh, len(ids), 0) == 0
        assert L.wap_process_streams(hh, len(ids), p(r), p(c), p(o), 0, None) == 0
        for k, i in enumerate(ids):
            out[i][(t - start[i]) * 160:(t - start[i] + 1) * 160] = o[k]
    for i in range(3):
        n = min(nf, nf + 31 - start[i]) * 160
        if i == 0:
            n = 30 * 160   # leg 0 left after 30 ticks
        assert np.abs(out[i][:n].astype(np.int32) - refs[i][:n].astype(np.int32)).max() <= TOL, i
    for h in live.values():
        L.wap_destroy(h)
    eng.handles = (C.c_void_p * 0)()
    eng.n = 0
    eng.close()


@pytest.mark.parametrize("level", [0, 3])
def test_ns_levels_low_and_very_high(api_lib, oracle, level):
    """NoiseSuppression kLow / kVeryHigh (suppression_params.cc:18-48); kModerate / kHigh are covered
    by the parity tests."""
    from common import run_engine
    near = golden("speech_16k.npz")["near"][:120 * 160]
    ref_out, _, err = oracle.RefApm(aec=False, ns=True, ns_level=level).run_i16(16000, None, near)
    assert err == 0
    out = run_engine(api_lib, 16000, None, near, n_streams=1, aec=False, ns=True, ns_level=level)
    assert np.abs(out.reshape(-1).astype(np.int32) - ref_out.astype(np.int32)).max() <= TOL


def test_pipelined_host_tick_matches_plain(api_lib, oracle):
    """wap_engine_set_pipeline_chunks: cutting the batch into leg ranges (ragged last range)
    changes nothing in the output."""
    import wap_b200
    from common import run_legs
    legs = [synthetic_leg(i, 30) for i in range(11)]
    ref, _ = run_legs(api_lib, 16000, legs, pipeline_chunks=1, aec=True, ns=True, ns_level=1)
    out, _ = run_legs(api_lib, 16000, legs, pipeline_chunks=3, aec=True, ns=True, ns_level=1)
    assert np.array_equal(ref, out)
    far, near = legs[10]
    ref_out, _, err = oracle.RefApm(aec=True, ns=True, ns_level=1).run_i16(16000, far, near)
    assert err == 0 and np.array_equal(out[10], ref_out[:out[10].size])
    eng = wap_b200.Engine(2, 16000, lib=api_lib, aec=True, ns=False)
    with pytest.raises(RuntimeError):
        eng.set_pipeline_chunks(99)
    eng.close()


def test_rates_and_unsupported_formats(api_lib):
    """Mono engines exist for every API rate up to 48 kHz that has whole 10 ms frames (native 16 / 32 /
    48 kHz, the others through the sinc resamplers); true multi-channel processing and rates above 48 kHz are outside
    the built scope and are refused, not approximated."""
    import wap_b200
    for rate, max_rate in ((16000, 32000), (32000, 32000), (48000, 48000), (48000, 32000), (8000, 32000),
                           (44100, 32000), (24000, 48000)):
        eng = wap_b200.Engine(1, rate, lib=api_lib, aec=True, ns=True, max_rate=max_rate)
        x = np.zeros((1, rate // 100), np.int16)
        assert eng.process(x, x).shape == (1, rate // 100)
        eng.close()
    for rate in (96000, 22050):
        with pytest.raises(RuntimeError):
            wap_b200.Engine(1, rate, lib=api_lib, aec=True, ns=True)
    eng = wap_b200.Engine(1, 48000, channels=2, lib=api_lib, aec=True, ns=True)  # stereo in/out, mono processing
    x = np.zeros((1, 960), np.int16)
    assert eng.process(x, x).shape == (1, 960)
    eng.close()


@pytest.mark.parametrize("rate,max_rate,kw", [
    (16000, 32000, dict(aec=True, ns=True, ns_level=1)),
    (48000, 32000, dict(aec=True, ns=True, ns_level=1, agc2=True, agc2_fixed_gain_db=6.0, pre_gain=1.5, post_gain=0.8)),
])
def test_stream_state_export_import(api_lib, oracle, rate, max_rate, kw):
    """Stream lifecycle: a leg exported in the middle of a call and imported into a leg of another
    engine continues bit-identically (output and statistics); a blob does not fit another config."""
    import wap_b200
    from common import synthetic_leg, synthetic_leg_48k
    nf, cut = 70, 37
    fl = rate // 100
    far, near = synthetic_leg(4, nf) if rate == 16000 else synthetic_leg_48k(4, nf, 1.0, rate=rate)
    ref_out, ref_stats, err = oracle.RefApm(max_rate=max_rate, **kw).run_i16(rate, far, near, stats_every=nf)
    assert err == 0

    def tick(eng, f, slot=0):
        eng.set_stream_delay_ms(0)
        r = np.zeros((eng.n, fl), np.int16); c = np.zeros((eng.n, fl), np.int16)
        r[slot] = far[f * fl:(f + 1) * fl]; c[slot] = near[f * fl:(f + 1) * fl]
        return eng.process(r, c)[slot]

    a = wap_b200.Engine(1, rate, lib=api_lib, max_rate=max_rate, **kw)
    out = np.zeros(nf * fl, np.int16)
    for f in range(cut):
        out[f * fl:(f + 1) * fl] = tick(a, f)
    a.set_playout_volume(120)           # a pending runtime setting travels with the leg
    blob = a.export_state(0)
    a.close()
    b = wap_b200.Engine(3, rate, lib=api_lib, max_rate=max_rate, **kw)   # another engine, another slot
    b.import_state(blob, 2)
    for f in range(cut, nf):
        out[f * fl:(f + 1) * fl] = tick(b, f, slot=2)
    st = b.stats(2)
    ref2 = oracle.RefApm(max_rate=max_rate, **kw)
    ro = np.zeros(nf * fl, np.int16)
    for f in range(nf):
        if f == cut:
            ref2.set_playout_volume(120)
        o, _, e2 = ref2.run_i16(rate, far[f * fl:(f + 1) * fl], near[f * fl:(f + 1) * fl])
        ro[f * fl:(f + 1) * fl] = o
    assert np.array_equal(out, ro)
    assert abs(st.echo_return_loss_enhancement - float(ref2.stats()[3])) <= 1e-6
    other = wap_b200.Engine(1, rate, lib=api_lib, max_rate=max_rate, aec=True, ns=False)
    with pytest.raises(RuntimeError):
        other.import_state(blob, 0)
    other.close(); b.close()


def test_stream_migrates_between_engines_device_to_device(api_lib, oracle, request):
    """wap_stream_migrate: a live leg moves to a free slot of another engine of the same config class with
    device-to-device copies (another GPU of the box when there is one) and continues bit-identically next to
    the legs that were already there; the source slot is free again; engines of another class refuse."""
    import wap_b200
    L = api_lib
    nf, cut, fl = 90, 41, 160
    legs = [synthetic_leg(i, nf) for i in (5, 8)]
    refs = [oracle.RefApm(aec=True, ns=True, ns_level=1).run_i16(16000, far, near)[0] for far, near in legs]
    dst_device = 0
    if request.node.callspec.params["api_lib"] == "gpu":
        import torch
        dst_device = 1 if torch.cuda.device_count() > 1 else 0
    a = wap_b200.Engine(2, 16000, lib=L, aec=True, ns=True, ns_level=1)
    b = wap_b200.Engine(1, 16000, lib=L, aec=True, ns=True, ns_level=1, capacity=3, device=dst_device)
    out = [np.zeros(nf * fl, np.int16) for _ in legs]

    def tick(handles, which, f):
        n = len(handles)
        arr = (C.c_void_p * n)(*handles)
        L.wap_streams_set_delay_ms(arr, n, 0)
        r = np.zeros((n, fl), np.int16); c = np.zeros((n, fl), np.int16); o = np.zeros((n, fl), np.int16)
        for k, w in enumerate(which):
            if w is not None:
                r[k] = legs[w][0][f * fl:(f + 1) * fl]; c[k] = legs[w][1][f * fl:(f + 1) * fl]
        p = lambda x: x.ctypes.data_as(C.c_void_p)
        assert L.wap_process_streams(arr, n, p(r), p(c), p(o), 0, None) == 0
        for k, w in enumerate(which):
            if w is not None:
                out[w][f * fl:(f + 1) * fl] = o[k]

    ha = [a.handles[0], a.handles[1]]
    for f in range(cut):
        tick(ha, [0, 1], f)
    other = wap_b200.Engine(1, 16000, lib=L, aec=True, ns=False)
    assert L.wap_stream_migrate(ha[1], other.h) == 7          # UnsupportedConfig: another config class
    assert L.wap_stream_migrate(ha[1], b.h) == 0              # leg 1 of engine a -> engine b
    for f in range(cut, nf):
        tick([ha[0]], [0], f)                                  # engine a goes on with leg 0
        tick([b.handles[0], ha[1]], [None, 1], f)              # engine b: its own (idle) leg + the migrated one
    assert np.array_equal(out[0], refs[0]) and np.array_equal(out[1], refs[1])
    # the source slot is free again, the destination has two legs left to hand out
    extra = (C.c_void_p * 1)()
    assert L.wap_engine_create_streams(a.h, 1, extra) == 0
    L.wap_destroy(extra[0])
    L.wap_destroy(ha[1]); a.handles[1] = None
    a.handles = (C.c_void_p * 1)(ha[0]); a.n = 1
    other.close(); a.close(); b.close()


def test_stage_taps_match_reference_dumps(api_lib, tmp_path):
    """wap_stream_read_taps against the reference's own ApmDataDumper output (the dump variant of the
    compiled reference, -DWEBRTC_APM_DEBUG_DUMP=1, run in a helper process): the taps after the last
    block equal the last records of the reference's tap files bit for bit."""
    import glob
    import subprocess
    import sys
    import wap_b200
    nf, leg = 260, 6
    dump_lib = os.path.join(ROOT, "oracle", "_ref", "libwap_ref_dump.so")
    if not os.path.exists(dump_lib):
        pytest.skip("dump variant of the oracle not built (python -c 'import __graft_entry__ as g; g.build()')")
    subprocess.run([sys.executable, os.path.join(ROOT, "tests", "dump_ref_taps.py"), str(tmp_path), str(nf), str(leg)],
                   check=True)
    far, near = synthetic_leg(leg, nf)
    eng = wap_b200.Engine(1, 16000, lib=api_lib, max_rate=32000, aec=True, ns=True, ns_level=1)
    for f in range(nf):
        eng.set_stream_delay_ms(0)
        eng.process(far[f * 160:(f + 1) * 160].reshape(1, 160), near[f * 160:(f + 1) * 160].reshape(1, 160))
    t = eng.taps(0)
    eng.close()

    def last(name, n, dtype):
        files = sorted(glob.glob(os.path.join(str(tmp_path), name + "_[0-9]*-[0-9]*.dat")))
        assert files, name
        return np.fromfile(files[-1], dtype)[-n:]

    for name in ("aec3_erle", "aec3_erle_onset_compensated", "aec3_erl", "aec3_suppressor_gain", "aec3_N2",
                 "aec3_refined_gain_H_error"):
        assert np.array_equal(np.array(getattr(t, name), np.float32), last(name, 65, np.float32)), name
    assert t.aec3_erl_time_domain == last("aec3_erl_time_domain", 1, np.float32)[0]
    assert t.aec3_fullband_erle_log2 == last("aec3_fullband_erle_log2", 1, np.float32)[0]
    assert t.aec3_filter_delay == last("aec3_filter_delay", 1, np.int32)[0]
    assert t.aec3_render_delay_controller_buffer_delay == last("aec3_render_delay_controller_buffer_delay", 1, np.int64)[0]
    for name in ("aec3_usable_linear_estimate", "aec3_transparent_mode", "aec3_initial_state", "aec3_echo_saturation",
                 "aec3_dominant_nearend"):   # bools are dumped as int16
        assert getattr(t, name) == last(name, 1, np.int16)[0], name
    assert t.aec3_capture_saturation == last("aec3_capture_saturation", 1, np.int32)[0]
    assert np.all(np.array(t.ns_noise_spectrum) > 0) and 0.0 <= t.ns_prior_speech_probability <= 1.0
    assert np.all((np.array(t.ns_filter) >= 0.0) & (np.array(t.ns_filter) <= 1.0))


def _drive_single(L, wap_b200, h, far, near, f0, f1, out):
    """ProcessReverseStream + set_stream_delay_ms(0) + ProcessStream on one handle, frames [f0, f1)."""
    sc = _sc(wap_b200, 16000)
    p = lambda a: a.ctypes.data_as(C.c_void_p)
    scratch = np.zeros(160, np.int16)
    for f in range(f0, f1):
        r = np.ascontiguousarray(far[f * 160:(f + 1) * 160])
        c = np.ascontiguousarray(near[f * 160:(f + 1) * 160])
        o = np.zeros(160, np.int16)
        assert L.wap_process_reverse_stream_i16(h, p(r), 160, sc, sc, p(scratch), 160) == 0
        assert L.wap_set_stream_delay_ms(h, 0) == 0
        assert L.wap_process_stream_i16(h, p(c), 160, sc, sc, p(o), 160) == 0
        out[f * 160:(f + 1) * 160] = o


def test_apply_config_mid_call_matches_reference(api_lib, oracle):
    """ApplyConfig in the middle of a call (audio_processing_impl.cc:694-771): a new NS level rebuilds
    only the noise suppressor and a new AGC2 gain only GainController2 -- AEC3 keeps its converged
    filters -- while toggling a submodule re-initialises everything.  Bit-compared with the reference
    driven through the same ApplyConfig calls."""
    import wap_b200
    L = api_lib
    nf = 260
    far, near = synthetic_leg(5, nf)
    ref = oracle.RefApm(kv=dict(aec=1, ns=1, ns_level=1, max_rate=48000, agc2=1, agc2_gain_db=3.0))
    h = L.wap_create_with_config(wap_b200.make_config(L, aec=True, ns=True, ns_level=1, agc2=True, agc2_fixed_gain_db=3.0))
    out = np.zeros_like(near)
    ref_out = np.zeros_like(near)
    steps = [(0, 120, None), (120, 170, dict(ns_level=2)), (170, 215, dict(agc2_gain_db=9.0)), (215, nf, dict(ns=0))]
    for f0, f1, change in steps:
        if change:
            ref.apply_config(**change)
            cfg = wap_b200.WapConfig()
            assert L.wap_get_config(h, C.byref(cfg)) == 0
            if "ns_level" in change:
                cfg.noise_suppression_level = int(change["ns_level"])
            if "agc2_gain_db" in change:
                cfg.gain_controller2_fixed_digital_gain_db = change["agc2_gain_db"]
            if "ns" in change:
                cfg.noise_suppression_enabled = bool(change["ns"])
            assert L.wap_apply_config(h, cfg) == 0
        _drive_single(L, wap_b200, h, far, near, f0, f1, out)
        o, _, err = ref.run_i16(16000, far[f0 * 160:f1 * 160], near[f0 * 160:f1 * 160])
        assert err == 0
        ref_out[f0 * 160:f1 * 160] = o
    L.wap_destroy(h)
    d = np.abs(out.astype(np.int32) - ref_out.astype(np.int32))
    assert d.max() == 0, (int(d.max()), int(np.argmax(d)) // 160)


def test_apply_config_with_an_identical_config_keeps_all_state(api_lib):
    """A logically identical WapConfig whose padding bytes differ (a C caller that fills the struct
    field by field) must not reset anything."""
    import wap_b200
    L = api_lib
    far, near = synthetic_leg(6, 100)
    outs = []
    for reapply in (False, True):
        h = L.wap_create_with_config(wap_b200.make_config(L, aec=True, ns=True, ns_level=1))
        out = np.zeros_like(near)
        _drive_single(L, wap_b200, h, far, near, 0, 60, out)
        if reapply:
            cur = wap_b200.WapConfig()
            assert L.wap_get_config(h, C.byref(cur)) == 0
            dirty = wap_b200.WapConfig()
            C.memset(C.byref(dirty), 0xA5, C.sizeof(dirty))       # garbage in the padding
            for name, _ in wap_b200.WapConfig._fields_:
                setattr(dirty, name, getattr(cur, name))
            assert bytes(dirty) != bytes(cur)
            assert L.wap_apply_config(h, dirty) == 0
        _drive_single(L, wap_b200, h, far, near, 60, 100, out)
        L.wap_destroy(h)
        outs.append(out)
    assert np.array_equal(outs[0], outs[1])


def test_render_format_of_a_shared_engine_is_fixed_and_never_reinterpreted(api_lib):
    """Legs of a batched engine have the formats the engine was created with: a render frame of another
    format is refused (UnsupportedConfig) without touching the leg -- its AEC3 / NS state survives and
    nothing is copied into a buffer of another size.  (A private single-leg handle follows the reference's
    MaybeInitializeRender instead: tests/test_formats.py.)"""
    import wap_b200
    L = api_lib
    far, near = synthetic_leg(7, 80)
    p = lambda a: a.ctypes.data_as(C.c_void_p)
    ref_run = np.zeros_like(near)
    h = L.wap_create_with_config(wap_b200.make_config(L, aec=True, ns=True, ns_level=1))
    _drive_single(L, wap_b200, h, far, near, 0, 80, ref_run)
    L.wap_destroy(h)
    e = wap_b200.Engine(1, 16000, lib=L, aec=True, ns=True, ns_level=1)
    h = e.handles[0]
    out = np.zeros_like(near)
    _drive_single(L, wap_b200, h, far, near, 0, 40, out)
    big = np.zeros(480, np.int16)
    sc48 = _sc(wap_b200, 48000)
    for _ in range(3):
        assert L.wap_process_reverse_stream_i16(h, p(big), 480, sc48, sc48, p(big), 480) == 7   # UnsupportedConfig
    _drive_single(L, wap_b200, h, far, near, 40, 80, out)
    assert np.array_equal(out, ref_run)            # the refused calls changed nothing
    o = np.zeros(480, np.int16)
    assert L.wap_process_stream_i16(h, p(big), 480, sc48, sc48, p(o), 480) == 5   # BadStreamParameter: fixed formats
    e.close()


def test_process_streams_rejects_engine_less_and_duplicate_handles(api_lib):
    import wap_b200
    L = api_lib
    h = L.wap_create()
    hs = (C.c_void_p * 1)(h)
    x = np.zeros(160, np.int16)
    p = lambda a: a.ctypes.data_as(C.c_void_p)
    assert L.wap_process_streams(hs, 1, None, p(x), p(x), 0, None) == 5       # BadStreamParameter, no crash
    L.wap_destroy(h)
    eng = wap_b200.Engine(2, 16000, lib=L, aec=True, ns=True)
    dup = (C.c_void_p * 2)(eng.handles[0], eng.handles[0])
    x2 = np.zeros((2, 160), np.int16)
    assert L.wap_process_streams(dup, 2, p(x2), p(x2), p(x2.copy()), 0, None) == 5
    assert eng.process(x2, x2).shape == (2, 160)
    eng.close()


@pytest.mark.gpu
def test_render_and_capture_threads_on_one_handle(gpu_lib, oracle):
    """The render thread (ProcessReverseStream) and the capture thread (everything else) drive one
    handle at the same time, as webrtc::AudioProcessing allows: the reverse call only enqueues, the
    capture side owns the engine.  Frames are handed over in lock step so the result is comparable."""
    import threading
    import wap_b200
    L = gpu_lib
    nf = 60
    far, near = synthetic_leg(8, nf)
    ref_out, _, err = oracle.RefApm(aec=True, ns=True, ns_level=1).run_i16(16000, far, near)
    assert err == 0
    h = L.wap_create_with_config(wap_b200.make_config(L, aec=True, ns=True, ns_level=1))
    sc = _sc(wap_b200, 16000)
    p = lambda a: a.ctypes.data_as(C.c_void_p)
    rendered = [threading.Event() for _ in range(nf)]
    captured = [threading.Event() for _ in range(nf)]
    errs = []

    def render_thread():
        scratch = np.zeros(160, np.int16)
        for f in range(nf):
            if f:
                captured[f - 1].wait()
            r = np.ascontiguousarray(far[f * 160:(f + 1) * 160])
            errs.append(L.wap_process_reverse_stream_i16(h, p(r), 160, sc, sc, p(scratch), 160))
            rendered[f].set()

    out = np.zeros_like(near)
    t = threading.Thread(target=render_thread)
    t.start()
    for f in range(nf):
        rendered[f].wait()
        c = np.ascontiguousarray(near[f * 160:(f + 1) * 160])
        o = np.zeros(160, np.int16)
        L.wap_set_stream_delay_ms(h, 0)
        errs.append(L.wap_process_stream_i16(h, p(c), 160, sc, sc, p(o), 160))
        out[f * 160:(f + 1) * 160] = o
        captured[f].set()
    t.join()
    L.wap_destroy(h)
    assert not any(errs)
    assert np.array_equal(out, ref_out)


def test_reference_seam_class_runs_on_this_library(api_lib, oracle):
    """The reference's own alternate-backend class, webrtc::RustAudioProcessing
    (modules/audio_processing/rust_audio_processing.cc, compiled unmodified against
    include/wap_audio_processing.h by oracle/build_ref.py), linked at load time against this library's
    wap_* symbols and driven through the loop of examples/run-offline.cpp: its output must be the
    output of the reference's built-in implementation."""
    import build_ref
    seam_path = build_ref.build_seam(verbose=False)
    # the implementation under test provides the wap_* symbols the seam library leaves undefined
    C.CDLL(api_lib._name, mode=C.RTLD_GLOBAL)
    glob = C.CDLL(None)
    if C.cast(glob.wap_create, C.c_void_p).value != C.cast(api_lib.wap_create, C.c_void_p).value:
        pytest.skip("another implementation's wap_* symbols are already global in this process "
                    "(emulator and GPU builds in one pytest run); run with -m gpu or -m 'not gpu'")
    S = C.CDLL(seam_path)
    S.seam_run_offline_i16.argtypes = [C.c_int] * 5 + [C.c_float] + [C.c_int] * 3 + [C.c_void_p] * 3 + [C.c_int, C.c_void_p]
    nf = 130
    far, near = synthetic_leg(12, nf)
    for kw in (dict(aec=True, ns=True, ns_level=1), dict(aec=True, ns=True, ns_level=2, agc2=True, agc2_fixed_gain_db=4.0)):
        ref_out, ref_stats, err = oracle.RefApm(**kw).run_i16(16000, far, near, stats_every=65)
        assert err == 0
        out = np.zeros_like(near)
        stats = np.zeros((2, 3), np.float64)
        p = lambda a: a.ctypes.data_as(C.c_void_p)
        rc = S.seam_run_offline_i16(1, 1, kw["ns_level"], 48000, int(kw.get("agc2", False)), float(kw.get("agc2_fixed_gain_db", 0.0)),
                                    16000, 1, nf, p(far), p(near), p(out), 65, p(stats))
        assert rc == 0, rc
        assert np.array_equal(out, ref_out)
        assert np.abs(stats[:, 1] - ref_stats[:, 3]).max() <= 0.1 and np.array_equal(stats[:, 2], ref_stats[:, 5])


def test_multi_channel_pipeline_flags_on_mono_legs(api_lib, oracle):
    """pipeline.multi_channel_render / _capture with mono frames: the reference runs its mono configuration
    (config_selector.cc:44-58, one render and one capture channel), so the flags are accepted and change
    nothing; with stereo frames they select true multi-channel processing, which is refused."""
    import wap_b200
    far, near = synthetic_leg(10, 100)
    ro, _, err = oracle.RefApm(aec=True, ns=True, ns_level=1, mc_render=True, mc_capture=True).run_i16(16000, far, near)
    assert err == 0
    eng = wap_b200.Engine(1, 16000, lib=api_lib, aec=True, ns=True, ns_level=1, mc_render=True, mc_capture=True)
    out = np.zeros_like(near)
    for f in range(100):
        sl = slice(f * 160, (f + 1) * 160)
        eng.set_stream_delay_ms(0)
        out[sl] = eng.process(far[sl][None, :], near[sl][None, :])[0]
    eng.close()
    assert np.array_equal(out, ro)
    # stereo frames with the flags are the multi-channel class (tests/test_multichannel.py); with one flag only: refused
    with pytest.raises(RuntimeError):
        wap_b200.Engine(1, 16000, channels=2, lib=api_lib, aec=True, ns=True, mc_render=True, mc_capture=False)
