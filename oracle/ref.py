"""ctypes binding of oracle/_ref/libwap_ref.so (the compiled, unmodified reference).

Test infrastructure only: imported by tests/, __graft_entry__.smoke() and
bench.py's cpu_baseline / --impl reference legs -- never by the product path.
"""
import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "_ref", "libwap_ref.so")
_lib = None


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise RuntimeError("oracle/_ref/libwap_ref.so missing: run python oracle/build_ref.py")
        L = C.CDLL(LIB_PATH)
        L.ref_apm_create.restype = C.c_void_p
        L.ref_apm_create.argtypes = [C.c_int] * 7
        L.ref_apm_create_agc2.restype = C.c_void_p
        L.ref_apm_create_agc2.argtypes = [C.c_int] * 5 + [C.c_float]
        L.ref_apm_destroy.argtypes = [C.c_void_p]
        L.ref_apm_run_i16.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int,
                                      C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_void_p]
        L.ref_apm_tick_f32.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int,
                                       C.c_void_p, C.c_void_p, C.c_void_p]
        L.ref_apm_stats.argtypes = [C.c_void_p, C.c_void_p]
        L.ref_apm_stats_echo_detector.argtypes = [C.c_void_p, C.c_void_p]
        L.ref_apm_set_capture_output_used.argtypes = [C.c_void_p, C.c_int]
        L.ref_apm_bench.restype = C.c_double
        L.ref_apm_bench.argtypes = [C.c_int] * 8 + [C.c_void_p, C.c_void_p, C.c_size_t]
        L.ref_apm_bench_kv.restype = C.c_double
        L.ref_apm_bench_kv.argtypes = [C.c_char_p] + [C.c_int] * 6 + [C.c_void_p, C.c_void_p, C.c_size_t]
        L.ref_apm_create_kv.restype = C.c_void_p
        L.ref_apm_create_kv.argtypes = [C.c_char_p]
        L.ref_apm_apply_kv.argtypes = [C.c_void_p, C.c_char_p]
        L.ref_ec3_validate_kv.argtypes = [C.c_char_p, C.c_char_p, C.POINTER(C.c_double)]
        L.ref_fft128.argtypes = [C.c_void_p, C.c_int]
        L.ref_rdft256.argtypes = [C.c_void_p, C.c_int]
        L.ref_hpf_create.restype = C.c_void_p
        L.ref_hpf_create.argtypes = [C.c_int, C.c_int]
        L.ref_hpf_destroy.argtypes = [C.c_void_p]
        L.ref_hpf_process.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_int]
        L.ref_3band_create.restype = C.c_void_p
        L.ref_3band_destroy.argtypes = [C.c_void_p]
        L.ref_3band_analysis.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p]
        L.ref_3band_synthesis.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p]
        _lib = L
    return _lib


def _p(a):
    return a.ctypes.data_as(C.c_void_p) if a is not None else None


class RefApm:
    """One reference webrtc::AudioProcessing instance (= one call leg)."""

    def __init__(self, aec=True, ns=True, ns_level=1, max_rate=48000, hpf=False,
                 mc_render=False, mc_capture=False, agc2=False, agc2_fixed_gain_db=0.0,
                 pre_amp=None, pre_gain=None, post_gain=None, kv=None):
        if kv is not None:
            # generic "key=value;..." configuration (ref_driver.cc: ParseKv); ec3.* keys inject an
            # EchoCanceller3Config through BuiltinAudioProcessingBuilder::SetEchoCancellerConfig
            self.h = lib().ref_apm_create_kv(kv_string(kv))
            if not self.h:
                raise ValueError("ref_apm_create_kv: unknown key in %r" % (kv,))
        elif pre_amp is not None or pre_gain is not None or post_gain is not None:
            L = lib()
            L.ref_apm_create_levels.restype = C.c_void_p
            L.ref_apm_create_levels.argtypes = [C.c_int] * 6 + [C.c_float, C.c_int, C.c_float, C.c_int, C.c_float, C.c_float]
            cla = pre_gain is not None or post_gain is not None
            self.h = L.ref_apm_create_levels(int(aec), int(ns), int(ns_level), int(max_rate), int(hpf), int(agc2),
                                             float(agc2_fixed_gain_db), int(pre_amp is not None), float(pre_amp or 1.0),
                                             int(cla), float(1.0 if pre_gain is None else pre_gain),
                                             float(1.0 if post_gain is None else post_gain))
        elif agc2:
            self.h = lib().ref_apm_create_agc2(int(aec), int(ns), int(ns_level), int(max_rate), 1,
                                               float(agc2_fixed_gain_db))
        else:
            self.h = lib().ref_apm_create(int(aec), int(ns), int(ns_level), int(max_rate),
                                          int(hpf), int(mc_render), int(mc_capture))

    def __del__(self):
        if getattr(self, "h", None):
            lib().ref_apm_destroy(self.h)
            self.h = None

    def run_i16(self, rate, render, capture, render_ch=1, capture_ch=1, stats_every=0):
        """render/capture: int16 [nframes*rate/100*ch] interleaved. Returns (out, stats)."""
        n = rate // 100
        capture = np.ascontiguousarray(capture, dtype=np.int16)
        nframes = capture.size // (n * capture_ch)
        if render is not None:
            render = np.ascontiguousarray(render, dtype=np.int16)
            assert render.size // (n * render_ch) >= nframes
        out = np.zeros(nframes * n * capture_ch, dtype=np.int16)
        ns = nframes // stats_every if stats_every else 0
        stats = np.zeros((max(ns, 1), 6), dtype=np.float32)
        err = lib().ref_apm_run_i16(self.h, rate, render_ch, capture_ch, nframes, _p(render),
                                    _p(capture), _p(out), stats_every, _p(stats))
        return out, stats[:ns], err

    def tick_f32(self, rate, render, capture, render_ch=1, capture_ch=1):
        capture = np.ascontiguousarray(capture, dtype=np.float32)
        render = None if render is None else np.ascontiguousarray(render, dtype=np.float32)
        out = np.zeros_like(capture)
        err = lib().ref_apm_tick_f32(self.h, rate, render_ch, capture_ch, _p(render), _p(capture), _p(out))
        return out, err

    def run_formats(self, render_fmt, in_fmt, out_fmt, render, capture):
        """One format (rate, channels) per stream.  render / capture: whole signals, int16 interleaved or
        float32 planar per frame ([frames][ch][samples]); returns the output frames stacked the same way."""
        L = lib()
        (rr, rc), (ir, ic), (orate, oc) = render_fmt, in_fmt, out_fmt
        capture = np.ascontiguousarray(capture)
        i16 = capture.dtype == np.int16
        fn = L.ref_apm_tick_fmt_i16 if i16 else L.ref_apm_tick_fmt_f32
        fn.argtypes = [C.c_void_p] + [C.c_int] * 6 + [C.c_void_p] * 3
        nr, ni, no = rr // 100 * rc, ir // 100 * ic, orate // 100 * oc
        frames = capture.size // ni
        render = None if render is None else np.ascontiguousarray(render, dtype=capture.dtype)
        out = np.zeros(frames * no, capture.dtype)
        cap = capture.reshape(-1)
        for f in range(frames):
            r = None if render is None else render.reshape(-1)[f * nr:(f + 1) * nr]
            c = cap[f * ni:(f + 1) * ni]
            o = out[f * no:(f + 1) * no]
            err = fn(self.h, rr, rc, ir, ic, orate, oc, _p(r), _p(c), _p(o))
            assert err == 0, err
        return out

    def reverse_f32(self, in_fmt, out_fmt, render):
        L = lib()
        L.ref_apm_reverse_f32.argtypes = [C.c_void_p] + [C.c_int] * 4 + [C.c_void_p] * 2
        (rr, rc), (orr, orc) = in_fmt, out_fmt
        render = np.ascontiguousarray(render, dtype=np.float32)
        out = np.zeros(orr // 100 * orc, np.float32)
        err = L.ref_apm_reverse_f32(self.h, rr, rc, orr, orc, _p(render), _p(out))
        return out, err

    def apply_config(self, **kv):
        """AudioProcessing::ApplyConfig with the current config updated by the given APM keys."""
        err = lib().ref_apm_apply_kv(self.h, kv_string(kv))
        assert err == 0, err

    def set_capture_output_used(self, used):
        lib().ref_apm_set_capture_output_used(self.h, int(used))

    def set_pre_gain(self, g):
        L = lib(); L.ref_apm_set_pre_gain.argtypes = [C.c_void_p, C.c_float]; L.ref_apm_set_pre_gain(self.h, float(g))

    def set_post_gain(self, g):
        L = lib(); L.ref_apm_set_post_gain.argtypes = [C.c_void_p, C.c_float]; L.ref_apm_set_post_gain(self.h, float(g))

    def set_fixed_post_gain(self, db):
        L = lib(); L.ref_apm_set_fixed_post_gain.argtypes = [C.c_void_p, C.c_float]; L.ref_apm_set_fixed_post_gain(self.h, float(db))

    def set_playout_volume(self, v):
        L = lib(); L.ref_apm_set_playout_volume.argtypes = [C.c_void_p, C.c_int]; L.ref_apm_set_playout_volume(self.h, int(v))

    def stats(self):
        s = np.zeros(6, dtype=np.float32)
        lib().ref_apm_stats(self.h, _p(s))
        return s

    def set_stream_delay_ms(self, ms):
        """The value passed to AudioProcessing::set_stream_delay_ms before every ProcessStream call (default 0)."""
        L = lib(); L.ref_apm_set_stream_delay_ms.argtypes = [C.c_void_p, C.c_int]; L.ref_apm_set_stream_delay_ms(self.h, int(ms))

    def stats_echo_detector(self):
        """[has_likelihood, residual_echo_likelihood, has_recent_max, residual_echo_likelihood_recent_max]"""
        s = np.zeros(4, dtype=np.float64)
        lib().ref_apm_stats_echo_detector(self.h, _p(s))
        return s


def kv_string(kv):
    if isinstance(kv, str):
        return kv.encode()
    return ";".join("%s=%r" % (k, float(v)) for k, v in kv.items()).encode()


def ec3_validate(kv, probe=None):
    """EchoCanceller3Config::Validate on default + kv; returns (was_valid, value of `probe` afterwards)."""
    out = C.c_double(0.0)
    r = lib().ref_ec3_validate_kv(kv_string({"ec3." + k: v for k, v in kv.items()}), probe.encode() if probe else None,
                                  C.byref(out))
    assert r >= 0, r
    return bool(r), out.value


def fft128(a, inverse=False):
    a = np.array(a, dtype=np.float32).copy()
    lib().ref_fft128(_p(a), int(inverse))
    return a


def rdft256(a, isgn=1):
    a = np.array(a, dtype=np.float32).copy()
    lib().ref_rdft256(_p(a), int(isgn))
    return a


def cpu_bench_kv(kv, rate, channels, streams, threads, warm, nframes, render, capture, stride=0):
    """ref_apm_bench for any configuration ref_apm_create_kv understands; frames interleaved per channel."""
    render = np.ascontiguousarray(render, dtype=np.int16)
    capture = np.ascontiguousarray(capture, dtype=np.int16)
    secs = lib().ref_apm_bench_kv(kv_string(kv), rate, channels, streams, threads, warm, nframes, _p(render), _p(capture), stride)
    assert secs > 0, "ref_apm_bench_kv: bad configuration"
    return secs


def cpu_bench(aec, ns, ns_level, rate, streams, threads, warm, nframes, render, capture, stride=0):
    render = np.ascontiguousarray(render, dtype=np.int16)
    capture = np.ascontiguousarray(capture, dtype=np.int16)
    return lib().ref_apm_bench(int(aec), int(ns), int(ns_level), rate, streams, threads, warm,
                               nframes, _p(render), _p(capture), stride)
