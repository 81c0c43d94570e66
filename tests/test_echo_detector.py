"""The residual echo detector: the reference's injected ResidualEchoDetector
(AudioProcessingBuilder::SetEchoDetector(CreateEchoDetector())) against wap_engine_enable_echo_detector --
WapStats::residual_echo_likelihood / _recent_max equal to the reference's statistics, value for value, on
every frame (or every few frames: the statistics travel through the reference's one-slot queue)."""
import ctypes as C

import numpy as np
import pytest

import wap_b200
from common import synthetic_leg, synthetic_leg_48k
from ref import RefApm


@pytest.fixture(params=["emu", pytest.param("gpu", marks=pytest.mark.gpu)])
def api_lib(request):
    return request.getfixturevalue("emu_lib" if request.param == "emu" else "gpu_lib")


def _residual_echo_leg(nf, seed=3, rate=16000):
    """Loud echo through a path longer than the linear filter: what the suppressor leaves correlates with
    the render power, so the likelihood rises."""
    rng = np.random.default_rng(seed)
    n = nf * rate // 100
    t = np.arange(n) / float(rate)
    x = rng.uniform(-12000, 12000, n) * (0.15 + 0.85 * ((t % 1.1) < 0.6))
    y = np.zeros(n)
    for g, d in ((0.7, rate // 50), (0.4, rate // 9), (0.3, rate // 5)):
        y[d:] += g * x[:n - d]
    y += rng.uniform(-50, 50, n)
    q = lambda v: np.clip(np.round(v), -32768, 32767).astype(np.int16)
    return q(x), q(y)


def _compare(api_lib, rate, legs, every=1, mute=None, aec=True, **cfg):
    n = rate // 100
    nf = legs[0][1].size // n
    kv = {"aec": int(aec), "ns": int(cfg.get("ns", True)), "max_rate": cfg.get("max_rate", 48000), "echo_detector": 1}
    if "ns_level" in cfg:
        kv["ns_level"] = cfg["ns_level"]
    refs = [RefApm(kv=kv) for _ in legs]
    eng = wap_b200.Engine(len(legs), rate, lib=api_lib, aec=aec, echo_detector=True, **cfg)
    seen = 0
    for f in range(nf):
        sl = slice(f * n, (f + 1) * n)
        if mute is not None and f in mute:
            used = mute[f]
            for i, r in enumerate(refs):
                api_lib.wap_set_capture_output_used(eng.handles[i], used)
                r.set_capture_output_used(used)
        eng.set_stream_delay_ms(0)
        out = eng.process(np.stack([l[0][sl] for l in legs]), np.stack([l[1][sl] for l in legs]))
        for i, (far, near) in enumerate(legs):
            ro, _, err = refs[i].run_i16(rate, far[sl], near[sl])
            assert err == 0
            assert np.array_equal(out[i], ro), (f, i)
            if f % every == every - 1:
                ours, theirs = eng.stats(i), refs[i].stats_echo_detector()
                assert bool(ours.has_residual_echo_likelihood) == bool(theirs[0]), (f, i)
                assert bool(ours.has_residual_echo_likelihood_recent_max) == bool(theirs[2]), (f, i)
                if theirs[0]:
                    assert ours.residual_echo_likelihood == theirs[1], (f, i, ours.residual_echo_likelihood, theirs[1])
                    assert ours.residual_echo_likelihood_recent_max == theirs[3], (f, i)
                    seen = max(seen, theirs[3])
    eng.close()
    return seen


def test_echo_likelihood_matches_the_reference_16k(api_lib):
    legs = [_residual_echo_leg(600), synthetic_leg(5, 600)]
    assert _compare(api_lib, 16000, legs, ns=True, ns_level=1) > 0.05   # the statistic is exercised


def test_echo_likelihood_through_the_statistics_slot_48k(api_lib):
    far, near = synthetic_leg_48k(4, 260, 0.5)
    _compare(api_lib, 48000, [(far, near)], every=7, ns=False)


def test_echo_likelihood_fullband_output_and_muted_stretches(api_lib):
    # 48 kHz API rate processed at 32 kHz (the capture_fullband_audio path); the output is muted for a
    # stretch: the detector pauses and the statistics keep their last values
    far, near = synthetic_leg_48k(9, 300, 0.6)
    _compare(api_lib, 48000, [(far, near)], every=3, ns=True, max_rate=32000, mute={90: False, 170: True})


def test_echo_likelihood_at_a_resampled_rate(api_lib):
    far, near = _residual_echo_leg(250, seed=8, rate=24000)
    _compare(api_lib, 24000, [(far, near)], ns=False)


@pytest.mark.parametrize("rate", [16000, 48000])
def test_echo_likelihood_without_echo_canceller(api_lib, rate):
    # NS-only legs: the detector compares the (uncancelled) capture with the render powers
    legs = [_residual_echo_leg(300, seed=5, rate=rate)]
    assert _compare(api_lib, rate, legs, every=2, aec=False, ns=True) > 0.05


def test_echo_detector_refusals_and_state_blob(api_lib):
    L = api_lib
    # multi-channel engines, and engines without AEC3 at a resampled rate: not built
    for rate, kw in ((16000, dict(aec=True, ns=False, mc_render=True, mc_capture=True)), (24000, dict(aec=False, ns=True))):
        ch = 2 if kw.get("mc_render") else 1
        with pytest.raises(RuntimeError):
            wap_b200.Engine(1, rate, channels=ch, lib=L, echo_detector=True, **kw)
    # after the first leg exists: refused
    eng = wap_b200.Engine(1, 16000, lib=L, aec=True, ns=False)
    assert L.wap_engine_enable_echo_detector(eng.h) == 5   # BadStreamParameter
    # the detector's state travels with the leg
    plain_bytes = L.wap_stream_state_bytes(eng.handles[0])
    eng.close()
    far, near = _residual_echo_leg(240, seed=12)
    a = wap_b200.Engine(1, 16000, lib=L, aec=True, ns=False, echo_detector=True)
    b = wap_b200.Engine(1, 16000, lib=L, aec=True, ns=False, echo_detector=True)
    assert L.wap_stream_state_bytes(a.handles[0]) > plain_bytes
    ref = RefApm(kv={"aec": 1, "ns": 0, "max_rate": 48000, "echo_detector": 1})
    for f in range(240):
        sl = slice(f * 160, (f + 1) * 160)
        eng = a if f < 120 else b
        if f == 120:
            b.import_state(a.export_state(0), 0)
        eng.set_stream_delay_ms(0)
        out = eng.process(far[sl].reshape(1, -1), near[sl].reshape(1, -1))
        ro, _, _ = ref.run_i16(16000, far[sl], near[sl])
        assert np.array_equal(out[0], ro)
        if f >= 120:
            ours, theirs = eng.stats(0), ref.stats_echo_detector()
            assert ours.residual_echo_likelihood == theirs[1] and ours.residual_echo_likelihood_recent_max == theirs[3], f
        else:
            eng.stats(0), ref.stats_echo_detector()   # keep the statistics slots in step
    a.close()
    b.close()
