"""Shared helpers of the parity tests."""
import os

import numpy as np

ROOT = os.path.normpath(os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
GOLDEN = os.path.join(ROOT, "tests", "golden")


def golden(name):
    return np.load(os.path.join(GOLDEN, name))


def run_engine(lib, rate, render, capture, n_streams=1, delay_ms=None, **cfg):
    """Drive n identical legs through wap_process_streams; returns [frames, frame_len] of leg 0
    and checks that all legs agree bit for bit."""
    import wap_b200
    fl = rate // 100
    nf = capture.size // fl
    eng = wap_b200.Engine(n_streams, rate, lib=lib, **cfg)
    out = np.zeros((nf, fl), np.int16)
    for f in range(nf):
        c = np.tile(capture[f * fl:(f + 1) * fl], (n_streams, 1))
        r = None if render is None else np.tile(render[f * fl:(f + 1) * fl], (n_streams, 1))
        if delay_ms is not None:
            eng.set_stream_delay_ms(delay_ms)
        o = eng.process(r, c)
        assert all(np.array_equal(o[0], o[i]) for i in range(1, n_streams))
        out[f] = o[0]
    eng.close()
    return out


def run_legs(lib, rate, legs, stats_every=0, delay_ms=0, pipeline_chunks=None, **cfg):
    """Drive len(legs) different call legs (list of (render, capture) int16 arrays) through ONE
    batched engine, one wap_process_streams call per 10 ms tick.  Returns (out [n, samples],
    stats [n, k, 3] = (erl, erle, delay_ms) sampled every `stats_every` frames)."""
    import wap_b200
    n = len(legs)
    fl = rate // 100
    nf = min(l[1].size for l in legs) // fl
    eng = wap_b200.Engine(n, rate, lib=lib, **cfg)
    if pipeline_chunks is not None:
        eng.set_pipeline_chunks(pipeline_chunks)
    out = np.zeros((n, nf * fl), np.int16)
    stats = []
    R = np.stack([l[0][:nf * fl] for l in legs]).reshape(n, nf, fl)
    Cc = np.stack([l[1][:nf * fl] for l in legs]).reshape(n, nf, fl)
    for f in range(nf):
        if delay_ms is not None:
            eng.set_stream_delay_ms(delay_ms)
        out[:, f * fl:(f + 1) * fl] = eng.process(R[:, f], Cc[:, f])
        if stats_every and (f + 1) % stats_every == 0:
            row = []
            for i in range(n):
                st = eng.stats(i)
                row.append((st.echo_return_loss, st.echo_return_loss_enhancement, st.delay_ms))
            stats.append(row)
    eng.close()
    st = np.array(stats, dtype=np.float64).transpose(1, 0, 2) if stats else np.zeros((n, 0, 3))
    return out, st


# ---- test signals (the bench workload generator is tests/synth.py + tools/wap_synth.c)
def synthetic_leg(i, n_frames, rate=16000):
    """Render/capture int16 for stream i: gated white-noise render, 3-tap echo path with
    per-stream delay, noise floor + double-talk bursts (the shape of SURVEY.md section 8d, drawn
    from numpy's generator; parity-test signal only -- bench.py uses tests/synth.py)."""
    n = n_frames * rate // 100
    rng_r = np.random.default_rng(1000 + 2 * i)
    rng_n = np.random.default_rng(1001 + 2 * i)
    t = np.arange(n) / rate
    x = rng_r.uniform(-8000, 8000, n)
    x *= ((t % 1.0) < 0.9)
    D = 64 * (1 + (i % 48)) + (7 * i) % 64
    y = np.zeros(n)
    for g, d in ((0.5, D), (0.25, D + 37), (0.1, D + 160)):
        y[d:] += g * x[:n - d]
    y += rng_n.uniform(-50, 50, n)
    y += rng_n.uniform(-3000, 3000, n) * ((t % 2.0) > 1.7)
    return (np.clip(np.round(x), -32768, 32767).astype(np.int16),
            np.clip(np.round(y), -32768, 32767).astype(np.int16))


def vanishing_echo_leg(n_frames, rate=16000, seed=11):
    """Loud white render (+-30000) with a single-tap echo path (0.6, 100 samples) that
    disappears after 3 s, over a +-400 noise floor."""
    rng = np.random.default_rng(seed)
    n = n_frames * rate // 100
    x = rng.uniform(-30000, 30000, n)
    y = np.zeros(n)
    y[100:] = 0.6 * x[:-100]
    y[3 * rate:] = 0
    y += rng.uniform(-400, 400, n)
    return (np.round(x).astype(np.int16), np.clip(np.round(y), -32768, 32767).astype(np.int16))


def synthetic_leg_48k(i, n_frames, hf_boost=1.0, rate=48000):
    """48 kHz (or 32 kHz) variant of synthetic_leg: full-band white render (so the upper bands carry
    as much energy as band 0 -- drives SuppressionGain::UpperBandsGain's anti-howling branch),
    3-tap echo path."""
    n = n_frames * rate // 100
    rng_r = np.random.default_rng(5000 + 2 * i)
    rng_n = np.random.default_rng(5001 + 2 * i)
    t = np.arange(n) / rate
    x = rng_r.uniform(-8000, 8000, n) * hf_boost
    x *= ((t % 1.0) < 0.9)
    D = (rate // 16000) * (64 * (1 + (i % 24)) + (7 * i) % 64)
    y = np.zeros(n)
    for g, d in ((0.5, D), (0.25, D + 111), (0.1, D + 480)):
        y[d:] += g * x[:n - d]
    y += rng_n.uniform(-50, 50, n)
    y += rng_n.uniform(-3000, 3000, n) * ((t % 2.0) > 1.7)
    return (np.clip(np.round(x), -32768, 32767).astype(np.int16),
            np.clip(np.round(y), -32768, 32767).astype(np.int16))


def loud_bursty_signal(rate, n_frames, seed=3):
    """Noise bursts + gated tone with peaks near 23 k: pushed through a fixed digital gain this
    drives the AGC2 limiter through its identity, knee, limiter and saturation regions and through
    the first-sub-frame attack interpolation."""
    rng = np.random.default_rng(seed)
    n = n_frames * rate // 100
    t = np.arange(n) / rate
    env = 1500 + 14000 * (np.sin(2 * np.pi * 1.3 * t) > 0.2) * (0.5 + 0.5 * np.sin(2 * np.pi * 0.37 * t) ** 2)
    x = rng.uniform(-1, 1, n) * env + 9000 * np.sin(2 * np.pi * 440 * t) * (t % 0.7 < 0.3)
    return x.clip(-32768, 32767).astype(np.int16)


def float_interface_max_diff(lib, oracle, rate, n_frames, max_rate=32000, leg=3, **kw):
    """Drive one leg through the float interface of both implementations; returns (number of output
    samples whose float32 bits differ, max |delta| in FloatS16 units)."""
    import wap_b200
    fl = rate // 100
    far, near = synthetic_leg_48k(leg, n_frames, 2.0, rate=rate)
    eng = wap_b200.Engine(1, rate, lib=lib, max_rate=max_rate, **kw)
    ref = oracle.RefApm(max_rate=max_rate, **kw)
    differing, worst = 0, 0.0
    for f in range(n_frames):
        c = (near[f * fl:(f + 1) * fl].astype(np.float32) / 32768.0).reshape(1, fl)
        r = (far[f * fl:(f + 1) * fl].astype(np.float32) / 32768.0).reshape(1, fl) if kw.get("aec") else None
        eng.set_stream_delay_ms(0)
        o = eng.process(r, c).reshape(-1)
        ro, err = ref.tick_f32(rate, None if r is None else r.reshape(-1), c.reshape(-1))
        assert err == 0
        differing += int(np.count_nonzero(o.view(np.uint32) != ro.view(np.uint32)))
        worst = max(worst, float(np.abs(o - ro).max()) * 32768.0)
    eng.close()
    return differing, worst


def stereo_leg(rate, n_frames, seed, right_gain=0.6):
    """Interleaved stereo (render, capture) int16: two different synthetic legs mixed per channel.
    right_gain > 2 clips the right capture channel while the left one stays moderate -- AEC3's
    saturation test looks at both channels although only the left one is processed."""
    f1, n1 = synthetic_leg_48k(seed, n_frames, 1.5, rate=rate)
    f2, n2 = synthetic_leg_48k(seed + 40, n_frames, 1.0, rate=rate)
    right = np.clip(n1.astype(np.float64) * right_gain + n2 * 0.4, -32768, 32767).astype(np.int16)
    return np.stack([f1, f2], 1).reshape(-1), np.stack([n1, right], 1).reshape(-1)


def run_stereo_i16(lib, oracle, rate, n_frames, seed, max_rate=32000, right_gain=0.6, **kw):
    """One stereo leg through both implementations (int16 interleaved); returns (ours, reference)."""
    import wap_b200
    far, near = stereo_leg(rate, n_frames, seed, right_gain)
    fl = rate // 100 * 2
    eng = wap_b200.Engine(1, rate, channels=2, lib=lib, max_rate=max_rate, **kw)
    ref_out, _, err = oracle.RefApm(max_rate=max_rate, **kw).run_i16(rate, far, near, render_ch=2, capture_ch=2)
    assert err == 0
    out = np.zeros_like(near)
    for f in range(n_frames):
        eng.set_stream_delay_ms(0)
        out[f * fl:(f + 1) * fl] = eng.process(far[f * fl:(f + 1) * fl].reshape(1, fl),
                                               near[f * fl:(f + 1) * fl].reshape(1, fl)).reshape(-1)
    eng.close()
    return out, ref_out


def run_with_runtime_settings(lib, oracle, rate, n_frames, events, max_rate=32000, leg=7, **kw):
    """One leg through the float interface of both implementations with runtime settings applied in
    front of given frames: events = [(frame, 'pre_gain' | 'post_gain' | 'fixed_post_gain' |
    'playout_volume', value)].  Returns the number of output samples whose float32 bits differ."""
    import wap_b200
    far, near = synthetic_leg(leg, n_frames) if rate == 16000 else synthetic_leg_48k(leg, n_frames, 1.0, rate=rate)
    fl = rate // 100
    eng = wap_b200.Engine(1, rate, lib=lib, max_rate=max_rate, **kw)
    ref = oracle.RefApm(max_rate=max_rate, **kw)
    differing = 0
    for f in range(n_frames):
        for ff, what, val in events:
            if ff == f:
                getattr(eng, "set_" + what)(val)
                getattr(ref, "set_" + what)(val)
        c = (near[f * fl:(f + 1) * fl].astype(np.float32) / 32768.0).reshape(1, fl)
        r = (far[f * fl:(f + 1) * fl].astype(np.float32) / 32768.0).reshape(1, fl) if kw.get("aec") else None
        eng.set_stream_delay_ms(0)
        o = eng.process(r, c).reshape(-1)
        ro, err = ref.tick_f32(rate, None if r is None else r.reshape(-1), c.reshape(-1))
        assert err == 0
        differing += int(np.count_nonzero(o.view(np.uint32) != ro.view(np.uint32)))
    eng.close()
    return differing
