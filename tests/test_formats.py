"""Format negotiation: legs whose capture input, capture output and render streams
have formats of their own -- AudioProcessingImpl::InitializeLocked's rate choice, one PushSincResampler
per stream, the capture downmix, rates above 48 kHz -- against the compiled reference, through the
batched engine (wap_engine_create_with_formats) and the single-leg seam entry points."""
import ctypes as C

import numpy as np
import pytest

import wap_b200
from ref import RefApm


@pytest.fixture(params=["emu", pytest.param("gpu", marks=pytest.mark.gpu)])
def api_lib(request):
    return request.getfixturevalue("emu_lib" if request.param == "emu" else "gpu_lib")


def _at_rate(x48, rate):
    """A 48 kHz test signal at another rate (plain interpolation: test data only)."""
    n = x48.size * rate // 48000
    t = np.arange(n) * (48000.0 / rate)
    return np.interp(t, np.arange(x48.size), x48)


def _leg(seed, frames, render_fmt, in_fmt, echo_gain=0.5):
    """render [frames * rr/100 * rc] and capture [frames * ir/100 * ic] int16 interleaved, sharing one
    48 kHz scene: gated white render, 3-tap echo path, noise floor, near-end bursts."""
    rng = np.random.default_rng(seed)
    n = frames * 480
    t = np.arange(n) / 48000.0
    x = rng.uniform(-9000, 9000, n) * ((t % 1.0) < 0.85)
    x = np.convolve(x, np.ones(3) / 3.0, mode="same")        # keeps some energy below 4 kHz at every rate
    d = 480 + 37 * (seed % 7)
    y = np.zeros(n)
    for g, dd in ((echo_gain, d), (echo_gain / 2, d + 111), (echo_gain / 5, d + 480)):
        y[dd:] += g * x[:n - dd]
    y += rng.uniform(-60, 60, n) + rng.uniform(-6000, 6000, n) * ((t % 0.8) > 0.5) + 1500 * np.sin(2 * np.pi * 440 * t)
    (rr, rc), (ir, ic) = render_fmt, in_fmt
    xr, yi = _at_rate(x, rr), _at_rate(y, ir)
    r = np.stack([xr * (1.0 if c == 0 else 0.6) + (0 if c == 0 else 300 * np.sin(np.arange(xr.size) * 0.01)) for c in range(rc)], 1)
    c = np.stack([yi * (1.0 if k == 0 else 0.8) + (0 if k == 0 else rng.uniform(-40, 40, yi.size)) for k in range(ic)], 1)
    q = lambda a: np.clip(np.round(a), -32768, 32767).astype(np.int16).reshape(-1)
    return q(r), q(c)


def _to_planar_float(a, frames, rate, ch):
    """interleaved int16 -> per-frame planar float32 [-1, 1]."""
    n = rate // 100
    return (a.reshape(frames, n, ch).transpose(0, 2, 1).astype(np.float32) / 32768.0).reshape(-1)


def _run_engine(L, cfg, render_fmt, in_fmt, out_fmt, legs, frames, as_float):
    (rr, rc), (ir, ic), (orate, oc) = render_fmt, in_fmt, out_fmt
    e = wap_b200.Engine(len(legs), ir, channels=ic, lib=L, out_format=out_fmt, render_format=render_fmt, **cfg)
    nr, ni, no = rr // 100 * rc, ir // 100 * ic, orate // 100 * oc
    outs = [[] for _ in legs]
    e.set_stream_delay_ms(0)
    for f in range(frames):
        r = np.stack([l[0][f * nr:(f + 1) * nr] for l in legs])
        c = np.stack([l[1][f * ni:(f + 1) * ni] for l in legs])
        e.set_stream_delay_ms(0)
        o = e.process(r, c)
        assert o.shape == (len(legs), no)
        for i in range(len(legs)):
            outs[i].append(o[i].copy())
    e.close()
    return [np.concatenate(o) for o in outs]


def _ref_kv(cfg):
    kv = {"aec": int(cfg.get("aec", True)), "ns": int(cfg.get("ns", True)), "ns_level": cfg.get("ns_level", 1),
          "max_rate": cfg.get("max_rate", 48000), "hpf": int(cfg.get("hpf", False))}
    if "downmix" in cfg:
        kv["downmix"] = cfg["downmix"]
    if cfg.get("agc2"):
        kv["agc2"] = 1
        kv["agc2_gain_db"] = cfg.get("agc2_fixed_gain_db", 0.0)
    return kv


CASES = {
    # name: (render (rate, ch), capture in, capture out, config)
    "in48k_out16k": ((48000, 1), (48000, 1), (16000, 1), dict(aec=True, ns=True)),
    "render48k_stereo_capture16k": ((48000, 2), (16000, 1), (16000, 1), dict(aec=True, ns=True)),
    "in32k_out16k_render8k": ((8000, 1), (32000, 1), (16000, 1), dict(aec=True, ns=True)),
    "all_96k_default_max_rate": ((96000, 1), (96000, 1), (96000, 1), dict(aec=True, ns=True, max_rate=32000)),
    "in96k_out48k_render44k1_proc48k": ((44100, 1), (96000, 1), (48000, 1), dict(aec=False, ns=True, ns_level=2, max_rate=48000)),
    "in44k1_render48k": ((48000, 1), (44100, 1), (44100, 1), dict(aec=True, ns=False, max_rate=32000)),
    "stereo_in_mono_out_average": ((16000, 1), (16000, 2), (16000, 1), dict(aec=True, ns=True)),
    "stereo_in_mono_out_first": ((16000, 2), (16000, 2), (16000, 1), dict(aec=True, ns=True, downmix=1)),
    "stereo48k_in_mono16k_out": ((32000, 2), (48000, 2), (16000, 1), dict(aec=True, ns=True)),
    "in16k_out24k": ((16000, 1), (16000, 1), (24000, 1), dict(aec=True, ns=True)),
    "ns_only_in48k_out32k": ((48000, 1), (48000, 1), (32000, 1), dict(aec=False, ns=True, max_rate=48000)),
    "agc2_in48k_out8k": ((48000, 1), (48000, 1), (8000, 1), dict(aec=True, ns=True, agc2=True, agc2_fixed_gain_db=6.0)),
    "stereo_default_pipeline_render_mono": ((16000, 1), (16000, 2), (16000, 2), dict(aec=True, ns=True)),
    # the reference's default routing of a 48 kHz client (processing at 32 kHz, 48 kHz full-band side buffer) with a
    # stereo microphone and a mono uplink
    "stereo48k_in_mono48k_out_default_max_rate": ((48000, 2), (48000, 2), (48000, 1), dict(aec=True, ns=True, max_rate=32000)),
}


@pytest.mark.parametrize("name", sorted(CASES))
def test_engine_formats_match_the_reference_i16(api_lib, oracle, name):
    render_fmt, in_fmt, out_fmt, cfg = CASES[name]
    frames = 220
    legs = [_leg(3 + 5 * i, frames, render_fmt, in_fmt) for i in range(2)]
    got = _run_engine(api_lib, cfg, render_fmt, in_fmt, out_fmt, legs, frames, False)
    for i, (r, c) in enumerate(legs):
        want = RefApm(kv=_ref_kv(cfg)).run_formats(render_fmt, in_fmt, out_fmt, r, c)
        d = np.abs(got[i].astype(np.int32) - want.astype(np.int32))
        assert d.max() == 0, (name, i, int(d.max()), int((d > 0).sum()), int(np.flatnonzero(d)[0]) // (out_fmt[0] // 100 * out_fmt[1]))
        assert np.abs(want).max() > 300


@pytest.mark.parametrize("name", ["in48k_out16k", "render48k_stereo_capture16k", "all_96k_default_max_rate",
                                  "stereo_in_mono_out_average", "in16k_out24k"])
def test_engine_formats_match_the_reference_f32(api_lib, oracle, name):
    render_fmt, in_fmt, out_fmt, cfg = CASES[name]
    frames = 120
    r, c = _leg(11, frames, render_fmt, in_fmt)
    rf = _to_planar_float(r, frames, *render_fmt)
    cf = _to_planar_float(c, frames, *in_fmt)
    got = _run_engine(api_lib, cfg, render_fmt, in_fmt, out_fmt, [(rf, cf)], frames, True)[0]
    want = RefApm(kv=_ref_kv(cfg)).run_formats(render_fmt, in_fmt, out_fmt, rf, cf)
    assert got.dtype == np.float32 and np.array_equal(got.view(np.uint32), want.view(np.uint32)), \
        (name, float(np.abs(got - want).max()))


def test_unsupported_format_combinations_are_refused(api_lib):
    L = api_lib
    sc = wap_b200.WapStreamConfig
    cfg = wap_b200.make_config(L, aec=True, ns=True, max_rate=32000)
    mk = lambda c, i, o, r: L.wap_engine_create_with_formats(0, 1, c, sc(*i), sc(*o), sc(*r), None, None)
    # 48 kHz output above the processing rate with another input format: capture_fullband_audio from a resampled input
    assert not mk(cfg, (16000, 1), (48000, 1), (16000, 1))
    # 48 kHz AEC3 (PostFilter) with another output rate
    assert not mk(wap_b200.make_config(L, aec=True, ns=True, max_rate=48000), (48000, 1), (96000, 1), (48000, 1))
    # rates above 96 kHz, more output channels than input channels
    assert not mk(cfg, (192000, 1), (192000, 1), (192000, 1))
    assert not mk(cfg, (16000, 1), (16000, 2), (16000, 1))
    # multi-channel processing needs one format for all three streams
    mc = wap_b200.make_config(L, aec=True, ns=False, mc_render=True, mc_capture=True, max_rate=48000)
    assert not mk(mc, (48000, 2), (48000, 2), (16000, 2))
    e = mk(mc, (48000, 2), (48000, 2), (48000, 2))
    assert e
    L.wap_engine_destroy(e)


def _sc(rate, ch=1):
    return wap_b200.WapStreamConfig(rate, ch)


def test_single_leg_seam_with_differing_formats(api_lib, oracle):
    """ProcessStream(src, input_config, output_config, dest) / ProcessReverseStream with formats of their own
    through the seam's single-leg entry points, including a render format that changes mid-call
    (MaybeInitializeRender: everything is re-initialised) and the render pass-through output."""
    L = api_lib
    p = lambda a: a.ctypes.data_as(C.c_void_p)
    render_fmt, in_fmt, out_fmt = (48000, 2), (32000, 1), (16000, 1)
    frames = 150
    r, c = _leg(5, frames, render_fmt, in_fmt)
    r2, c2 = _leg(6, frames, (16000, 1), in_fmt)
    ref = RefApm(kv=_ref_kv(dict(aec=True, ns=True)))
    want = np.concatenate([ref.run_formats(render_fmt, in_fmt, out_fmt, r, c),
                           ref.run_formats((16000, 1), in_fmt, out_fmt, r2, c2)])
    h = L.wap_create_with_config(wap_b200.make_config(L, aec=True, ns=True))
    got = []
    for (rf, rs, cs) in ((render_fmt, r, c), ((16000, 1), r2, c2)):
        nr, ni, no = rf[0] // 100 * rf[1], in_fmt[0] // 100, out_fmt[0] // 100
        for f in range(frames):
            rr = rs[f * nr:(f + 1) * nr].copy()
            scratch = np.full(nr, 77, np.int16)
            assert L.wap_process_reverse_stream_i16(h, p(rr), nr, _sc(*rf), _sc(*rf), p(scratch), nr) == 0
            assert np.array_equal(scratch, rr)          # identical formats: the input comes back
            assert L.wap_set_stream_delay_ms(h, 0) == 0
            o = np.zeros(no, np.int16)
            assert L.wap_process_stream_i16(h, p(cs[f * ni:(f + 1) * ni].copy()), ni, _sc(*in_fmt), _sc(*out_fmt), p(o), no) == 0
            got.append(o)
    got = np.concatenate(got)
    d = np.abs(got.astype(np.int32) - want.astype(np.int32))
    assert d.max() == 0, (int(d.max()), int(np.flatnonzero(d)[0]) // 160)
    # int16 reverse stream with another output format: dest is not written (no render processing)
    rr = r[:960].copy()
    scratch = np.full(160, 77, np.int16)
    assert L.wap_process_reverse_stream_i16(h, p(rr), 960, _sc(48000, 2), _sc(16000, 1), p(scratch), 160) == 0
    assert (scratch == 77).all()
    # float reverse stream: channel conversions of the pass-through output like AudioConverter's
    x = (np.random.default_rng(1).uniform(-0.5, 0.5, (2, 480))).astype(np.float32)
    src = (C.POINTER(C.c_float) * 2)(*[x[i].ctypes.data_as(C.POINTER(C.c_float)) for i in range(2)])
    mono = np.zeros(480, np.float32)
    dst = (C.POINTER(C.c_float) * 1)(mono.ctypes.data_as(C.POINTER(C.c_float)))
    assert L.wap_process_reverse_stream_f32(h, src, _sc(48000, 2), _sc(48000, 1), dst) == 0
    want_mono, err = RefApm(kv=_ref_kv(dict(aec=True, ns=True))).reverse_f32((48000, 2), (48000, 1), x)
    assert err == 0 and np.array_equal(mono.view(np.uint32), want_mono.view(np.uint32))
    up = np.zeros((2, 480), np.float32)
    src1 = (C.POINTER(C.c_float) * 1)(x[0].ctypes.data_as(C.POINTER(C.c_float)))
    dst2 = (C.POINTER(C.c_float) * 2)(*[up[i].ctypes.data_as(C.POINTER(C.c_float)) for i in range(2)])
    # (mono -> stereo is not a valid reverse configuration for the reference: out channels must be 1 or in channels)
    assert L.wap_process_reverse_stream_f32(h, src1, _sc(48000, 1), _sc(48000, 2), dst2) != 0
    # a rate conversion of the pass-through output is refused, not approximated
    assert L.wap_process_reverse_stream_f32(h, src, _sc(48000, 2), _sc(16000, 2), dst2) == 7
    L.wap_destroy(h)


def test_fullband_side_buffer_with_channel_downmix_while_muted(api_lib, oracle):
    """48 kHz stereo in, 48 kHz mono out under the default maximum_internal_processing_rate: while the output is
    muted the capture_fullband_audio buffer -- the downmixed, otherwise unprocessed input -- comes back."""
    L = api_lib
    render_fmt, in_fmt, out_fmt = (48000, 1), (48000, 2), (48000, 1)
    cfg = dict(aec=True, ns=True, max_rate=32000)
    frames = 60
    r, c = _leg(8, frames, render_fmt, in_fmt)
    ref = RefApm(kv=_ref_kv(cfg))
    e = wap_b200.Engine(1, 48000, channels=2, lib=L, out_format=out_fmt, render_format=render_fmt, **cfg)
    for f in range(frames):
        if f == 20:
            e.set_capture_output_used(False); ref.set_capture_output_used(False)
        if f == 40:
            e.set_capture_output_used(True); ref.set_capture_output_used(True)
        rr, cc = r[f * 480:(f + 1) * 480], c[f * 960:(f + 1) * 960]
        e.set_stream_delay_ms(0)
        got = e.process(rr.reshape(1, -1), cc.reshape(1, -1)).reshape(-1)
        want = ref.run_formats(render_fmt, in_fmt, out_fmt, rr, cc)
        assert np.array_equal(got, want), f
    e.close()
