#!/usr/bin/env python3
"""bench.py -- concurrent real-time AEC3+NS streams per GPU (16 kHz, 10 ms frames).

One "step" = one 10 ms tick of the hot path over all S call legs of this GPU
(ProcessReverseStream + set_stream_delay_ms(0) + ProcessStream per leg, i.e. one
wap_process_streams call = k_front + k_delay + k_echo).  metric = legs that can be served
in real time = S * 10 ms / tick time, summed over GPUs (legs shard across GPUs
with no collective: "scaling": "weak").

Workload (tests/synth.py + tools/wap_synth.c, the SAME generator for every arm):
leg i = xorshift64* (webrtc::Random) white-noise render gated 0.9 s on / 0.1 s off, a 3-tap echo
path with a per-leg delay, noise floor and double-talk bursts, one seamless 2 s cycle.

Steady state: whatever --warmup says, every leg is first advanced SETTLE (>= 300) ticks, so the
timed region never sees AEC3's initial state (12 partitions, no delay estimate); the device-resident
number (`value`) and the host-buffer number (`e2e`) are then measured on legs of the same age.

  value : device-resident timing (int16 frames already in HBM, wap_process_streams_device)
  e2e   : same metric through the host-buffer C ABI (wap_process_streams), H2D of the
          render+capture frames and D2H of the output inside the timed region
  roofline : for the dominant kernel of the tick: its algorithmic bytes (SURVEY.md 8(d) byte
             model split per kernel) * S / its launch duration (CUDA events around each kernel
             on the engine's stream) against the measured HBM copy bandwidth
             (MEASURED_PEAKS.json); "whole_tick" gives the same for all three kernels together
  parity_spot_check : after timing, K randomly chosen legs of the batch are compared with the
             compiled reference (oracle/_ref) replaying the whole history of those legs
  other_configs : BASELINE configs 3 (NS-only kHigh 48 kHz) and 5 (AEC3+NS+AGC2) measured the same way
  cpu_baseline : the compiled reference (oracle/_ref) on the host cores, one
             AudioProcessing instance per leg, bounded sample of the same legs

--impl reference times only the reference CPU implementation (oracle/_ref).
"""
import argparse
import ctypes as C
import json
import os
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
for p in (os.path.join(ROOT, "webrtc-audio-processing_b200", "python"), os.path.join(ROOT, "oracle"),
          os.path.join(ROOT, "tests")):
    if p not in sys.path:
        sys.path.insert(0, p)

CYCLE = 200        # frames per leg before the input repeats: 2 s = one double-talk period, two render gates
SETTLE_MIN = 300   # ticks every leg is advanced before anything is timed
MC_SETTLE = 520    # the same for multichannel legs: 201 frames until the stereo detector switches + 3 s


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=300)
    ap.add_argument("--warmup", type=int, default=20)
    ap.add_argument("--settle", type=int, default=400,
                    help="ticks every leg is advanced before the warm-up (>= %d enforced): steady state" % SETTLE_MIN)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--streams", type=int, default=int(os.environ.get("WAP_BENCH_STREAMS", "65536")),
                    help="call legs per GPU")
    ap.add_argument("--aec", type=int, default=1)
    ap.add_argument("--ns", type=int, default=1)
    ap.add_argument("--ns-level", type=int, default=1)
    ap.add_argument("--agc2-gain-db", type=float, default=None,
                    help="add GainController2 (fixed gain + limiter) to the chain: BASELINE config 5")
    ap.add_argument("--max-rate", type=int, default=48000, choices=[32000, 48000],
                    help="pipeline.maximum_internal_processing_rate (32000 = the reference default: 48 kHz legs are resampled)")
    ap.add_argument("--rate", type=int, default=16000, choices=[16000, 32000, 48000],
                    help="native sample rate of the legs (BASELINE config 3: --rate 48000 --aec 0 --ns-level 2)")
    ap.add_argument("--mc", type=int, default=0,
                    help="1: stereo legs with pipeline.multi_channel_render/_capture (multichannel AEC3, BASELINE config 4: "
                         "--mc 1 --rate 48000 --ns 0)")
    ap.add_argument("--external-delay-estimator", type=int, default=0,
                    help="1: EchoCanceller3Config::delay.use_external_delay_estimator (the render buffer follows "
                         "set_stream_delay_ms, no matched filters) -- a side configuration, not the headline")
    ap.add_argument("--cpu-seconds", type=float, default=12.0, help="budget of the cpu_baseline sample")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-other-configs", action="store_true", help="skip the BASELINE config 3 / 5 side lines")
    ap.add_argument("--check-legs", type=int, default=16, help="legs of the parity spot check (0: off)")
    a = ap.parse_args()
    a.settle = max(a.settle, MC_SETTLE if a.mc else SETTLE_MIN)
    return a


def workload_name(a):
    parts = []
    if a.aec:
        parts.append("AEC3(delay.use_external_delay_estimator, else default EchoCanceller3Config)"
                     if getattr(a, "external_delay_estimator", 0) else "AEC3(default EchoCanceller3Config)")
    if a.ns:
        parts.append("NS(%s)" % ["low", "moderate", "high", "veryhigh"][a.ns_level])
    if getattr(a, "agc2_gain_db", None) is not None:
        parts.append("AGC2(fixed %g dB + limiter)" % a.agc2_gain_db)
    if getattr(a, "max_rate", 48000) == 32000 and a.rate == 48000:
        parts.append("processed at 32 kHz (default maximum_internal_processing_rate)")
    if getattr(a, "mc", 0):
        return "%d synthetic stereo %d kHz call legs per GPU (multi_channel_render + _capture: 2 render / 2 capture channels), " \
               "%s, 10 ms frames" % (a.streams, a.rate // 1000, "+".join(parts).replace("default EchoCanceller3Config", "default mono + multichannel EchoCanceller3Config"))
    return "%d synthetic mono %d kHz call legs per GPU, %s, 10 ms frames" % (a.streams, a.rate // 1000, "+".join(parts))


def peaks():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


def synth_kind(a):
    if getattr(a, "mc", 0):
        return 2
    return 0 if a.aec else 1


def channels(a):
    return 2 if getattr(a, "mc", 0) else 1


def ref_kwargs(a):
    if getattr(a, "external_delay_estimator", 0):
        return dict(kv={"aec": int(a.aec), "ns": int(a.ns), "ns_level": a.ns_level, "max_rate": a.max_rate,
                        "ec3.delay.use_external_delay_estimator": 1})
    kw = dict(aec=bool(a.aec), ns=bool(a.ns), ns_level=a.ns_level, max_rate=a.max_rate)
    if getattr(a, "mc", 0):
        kw.update(mc_render=True, mc_capture=True)
    if a.agc2_gain_db is not None:
        kw.update(agc2=True, agc2_fixed_gain_db=a.agc2_gain_db)
    return kw


# ------------------------------------------------------------------ CPU reference leg
def cpu_reference(a, seconds):
    """Times oracle/_ref: threads = host cores, one AudioProcessing instance per leg, legs 0..n-1 of
    the bench workload, warm-up as long as the GPU arm's settle phase (steady state)."""
    import numpy as np
    import ref
    import synth
    cores = os.cpu_count() or 1
    fl = a.rate // 100 * channels(a)
    # multichannel legs: persistent stereo content is detected after 201 frames (Initialize() with two render
    # channels); warm up well past it
    warm, timed = (MC_SETTLE if getattr(a, "mc", 0) else SETTLE_MIN), 300
    per_leg_frames = warm + timed
    # ~65-150 us per leg-frame per core (~400 us for the multichannel 48 kHz shape): size the
    # sample to `seconds`.
    legs_per_thread = max(1, int(seconds / (per_leg_frames * (400e-6 if getattr(a, "mc", 0) else 160e-6))))
    legs = cores * legs_per_thread
    r, c = synth.cycle(synth_kind(a), a.rate, 0, legs, CYCLE)          # [CYCLE][legs][fl]
    idx = np.arange(per_leg_frames) % CYCLE
    r = np.ascontiguousarray(r[idx].transpose(1, 0, 2)).reshape(legs, -1)   # [leg][frame*fl]
    c = np.ascontiguousarray(c[idx].transpose(1, 0, 2)).reshape(legs, -1)
    if getattr(a, "mc", 0):
        kv = dict(aec=int(a.aec), ns=int(a.ns), ns_level=a.ns_level, max_rate=a.max_rate, mc_render=1, mc_capture=1)
        secs = ref.cpu_bench_kv(kv, a.rate, 2, legs, cores, warm, per_leg_frames, r, c, stride=r.shape[1])
    elif getattr(a, "external_delay_estimator", 0):
        secs = ref.cpu_bench_kv(ref_kwargs(a)["kv"], a.rate, 1, legs, cores, warm, per_leg_frames, r, c, stride=r.shape[1])
    else:
        secs = ref.cpu_bench(a.aec, a.ns, a.ns_level, a.rate, legs, cores, warm, per_leg_frames, r, c,
                             stride=r.shape[1])
    frames = legs * timed
    streams_rt = frames / secs / 100.0
    return {"value": streams_rt, "unit": "real-time streams", "cores": cores, "kind": "reference",
            "sample": "legs 0..%d of the bench workload x %d timed frames (after %d warm-up frames: steady state), "
                      "one webrtc::AudioProcessing per leg, %d pinned threads, %.1f s wall"
                      % (legs - 1, timed, warm, cores, secs),
            "us_per_leg_frame_per_core": secs * cores / frames * 1e6}


def run_reference(a):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    t0 = time.time()
    vals = []
    base = None
    for _ in range(max(1, min(a.steps, 3))):
        base = cpu_reference(a, a.cpu_seconds)
        vals.append(base["value"])
    v = max(vals)
    base["value"] = v
    line = {"impl": "reference", "metric": "concurrent real-time AEC3+NS streams (16 kHz, 10 ms)",
            "value": v, "unit": "real-time streams", "n_gpus": a.gpus, "steps": len(vals), "warmup": a.warmup,
            "ms_per_step": (time.time() - t0) * 1e3 / len(vals), "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": workload_name(a), "note": "reference CPU path on all host cores; "
                       "each step is a bounded sample (see cpu_baseline.sample)"},
            "cpu_baseline": base,
            "e2e": {"value": v, "unit": "real-time streams", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    emit(line)


# ------------------------------------------------------------------ GPU leg
class ClockSampler(threading.Thread):
    def __init__(self, index):
        super().__init__(daemon=True)
        self.index, self.stop_flag, self.sm, self.reasons, self.max_mhz = index, False, [], set(), None
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
        except Exception:
            self.nv = None

    def run(self):
        if not self.nv:
            return
        nv = self.nv
        names = {"hw_slowdown": nv.nvmlClocksThrottleReasonHwSlowdown,
                 "hw_thermal_slowdown": nv.nvmlClocksThrottleReasonHwThermalSlowdown,
                 "sw_thermal_slowdown": nv.nvmlClocksThrottleReasonSwThermalSlowdown,
                 "sw_power_cap": nv.nvmlClocksThrottleReasonSwPowerCap}
        while not self.stop_flag:
            try:
                self.sm.append(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM))
                r = nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
                for k, bit in names.items():
                    if r & bit:
                        self.reasons.add(k)
            except Exception:
                pass
            time.sleep(0.05)

    def summary(self):
        sm = sorted(self.sm)
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": self.max_mhz,
                "reasons": sorted(self.reasons)}


def make_inputs(torch, dev, a, first_leg, cycle):
    """Pinned host copies [cycle][S][fl] int16 of this rank's legs (the e2e arm feeds them through the
    host-buffer ABI) and their device copies (the device-resident arm)."""
    import synth
    S, fl = a.streams, a.rate // 100 * channels(a)
    h_r = torch.empty((cycle, S, fl), dtype=torch.int16).pin_memory()
    h_c = torch.empty((cycle, S, fl), dtype=torch.int16).pin_memory()
    synth.cycle(synth_kind(a), a.rate, first_leg, S, cycle, h_r.numpy(), h_c.numpy())
    return h_r, h_c, h_r.to(dev, non_blocking=True), h_c.to(dev, non_blocking=True)


def spot_check(a, L, eng, tick_host, h_o, t_now, h_r, h_c, first_leg, n_legs, n_ticks, cycle, seed):
    """Advance `n_ticks` more ticks through the host ABI, keep the output of `n_legs` randomly chosen
    legs, and compare with the compiled reference replaying those legs' whole history (ticks 0 ..
    t_now + n_ticks - 1; every tick of this run fed frame (t mod cycle) of the leg)."""
    import numpy as np
    try:
        import ref
        ref.lib()
    except Exception as e:
        return {"legs": 0, "unavailable": "oracle/_ref not loadable: %s" % e}, t_now
    S, fl = a.streams, a.rate // 100 * channels(a)
    rng = np.random.default_rng(seed)
    legs = sorted(int(x) for x in rng.choice(S, size=min(n_legs, S), replace=False))
    got = np.zeros((len(legs), n_ticks, fl), np.int16)
    t = t_now
    for k in range(n_ticks):
        tick_host(t)
        L.wap_engine_synchronize(eng.h)
        got[:, k] = h_o[legs]
        t += 1
    idx = np.arange(t) % cycle
    worst, differing = 0, 0
    r_np, c_np = h_r.numpy(), h_c.numpy()
    for j, leg in enumerate(legs):
        cap = np.ascontiguousarray(c_np[idx, leg]).reshape(-1)
        ren = np.ascontiguousarray(r_np[idx, leg]).reshape(-1) if a.aec else None
        o, _, err = ref.RefApm(**ref_kwargs(a)).run_i16(a.rate, ren, cap, render_ch=channels(a), capture_ch=channels(a))
        assert err == 0
        d = np.abs(o.reshape(t, fl)[t_now:].astype(np.int32) - got[j].astype(np.int32))
        worst = max(worst, int(d.max()))
        differing += int(np.count_nonzero(d))
    return {"legs": len(legs), "leg_ids": [first_leg + l for l in legs], "frames_compared_per_leg": n_ticks,
            "history_frames_replayed": t, "max_abs_diff_lsb": worst, "differing_samples": differing,
            "tolerance_lsb": 3, "pass": worst <= 3,
            "oracle": "oracle/_ref (unmodified reference), same int16 frames"}, t


def measure(a, torch, dist, L, wap_b200, dev, local, rank, world, headline):
    """Everything measured for one config class; returns the fields of its (sub-)line."""
    import numpy as np
    S, fl = a.streams, a.rate // 100 * channels(a)
    cycle = CYCLE if a.aec else 100
    extra = {} if a.agc2_gain_db is None else dict(agc2=True, agc2_fixed_gain_db=a.agc2_gain_db)
    if getattr(a, "mc", 0):
        extra.update(mc_render=True, mc_capture=True)
    if getattr(a, "external_delay_estimator", 0):
        extra.update(aec3={"delay.use_external_delay_estimator": 1})
    eng = wap_b200.Engine(S, a.rate, channels=channels(a), lib=L, device=local, aec=bool(a.aec), ns=bool(a.ns),
                          ns_level=a.ns_level, max_rate=a.max_rate, **extra)
    first_leg = rank * S
    h_r, h_c, render, capture = make_inputs(torch, dev, a, first_leg, cycle)
    out = torch.empty((S, fl), dtype=torch.int16, device=dev)
    h_o_t = torch.empty((S, fl), dtype=torch.int16).pin_memory()
    h_o = h_o_t.numpy()
    stream = torch.cuda.ExternalStream(L.wap_engine_cuda_stream(eng.h), device=dev)
    torch.cuda.synchronize()
    use_render = bool(a.aec)

    def tick_device(t):
        k = t % cycle
        L.wap_streams_set_delay_ms(eng.handles, S, 0)   # set_stream_delay_ms(0) per leg, as the reference arm
        err = L.wap_process_streams_device(eng.h, eng.handles, S, render[k].data_ptr() if use_render else None,
                                           capture[k].data_ptr(), out.data_ptr(), 0)
        assert err == 0, err

    r_np, c_np = h_r.numpy(), h_c.numpy()

    def tick_host(t):
        k = t % cycle
        L.wap_streams_set_delay_ms(eng.handles, S, 0)
        err = L.wap_process_streams(eng.handles, S, r_np[k].ctypes.data_as(C.c_void_p) if use_render else None,
                                    c_np[k].ctypes.data_as(C.c_void_p), h_o.ctypes.data_as(C.c_void_p), 0, None)
        assert err == 0, err

    def barrier():
        torch.cuda.synchronize()
        L.wap_engine_synchronize(eng.h)
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ---- settle: leave the initial state behind, independent of --warmup
    t = 0
    for _ in range(a.settle):
        tick_device(t); t += 1
    # ---- device-resident timing
    for _ in range(a.warmup):
        tick_device(t); t += 1
    barrier()
    sampler = ClockSampler(local)
    sampler.start()
    l0 = L.wap_engine_launch_count(eng.h)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(stream)
    for _ in range(a.steps):
        tick_device(t); t += 1
    e1.record(stream)
    barrier()
    ms = e0.elapsed_time(e1)
    launches = L.wap_engine_launch_count(eng.h) - l0
    sampler.stop_flag = True
    sampler.join()

    # ---- per-kernel durations (CUDA events on the engine's stream around each of the three
    # tick kernels, live, separate from the throughput loop above so it is not perturbed)
    kt_ticks = max(10, min(a.steps, 50))
    L.wap_engine_enable_kernel_timing(eng.h, True)
    for _ in range(kt_ticks):
        tick_device(t); t += 1
    barrier()
    kms = (C.c_double * 3)()
    n_timed = L.wap_engine_read_kernel_timing(eng.h, kms)
    kbytes = (C.c_double * 3)()
    L.wap_engine_algorithmic_bytes_per_kernel(eng.h, kbytes)
    L.wap_engine_enable_kernel_timing(eng.h, False)
    kernels = []
    knames = ("k_mc_front", "k_delay", "k_mc_echo(+k_mc_post)") if getattr(a, "mc", 0) else ("k_front", "k_delay", "k_echo")
    for name, m, bts in zip(knames, kms, kbytes):
        per = m / max(1, n_timed)
        kernels.append({"name": name, "ms_per_launch": per, "algorithmic_bytes_per_leg_frame": bts,
                        "achieved_gbs": (bts * S / (per * 1e-3) / 1e9) if per > 0 else 0.0})

    # ---- end to end through the host-buffer ABI (pinned host frames in, result out), same state age
    e2e_steps = max(10, min(a.steps, 100))
    for _ in range(3):
        tick_host(t); t += 1
    barrier()
    trace = [] if os.environ.get("WAP_BENCH_E2E_TRACE") else None   # per-tick wall times to stderr (diagnostics)
    w0 = time.perf_counter()
    for _ in range(e2e_steps):
        t0 = time.perf_counter()
        tick_host(t); t += 1
        if trace is not None:
            trace.append(round((time.perf_counter() - t0) * 1e3, 2))
    barrier()
    e2e_ms = (time.perf_counter() - w0) * 1e3
    if trace is not None and rank == 0:
        sys.stderr.write("e2e per-tick ms: %s\n" % trace)

    # ---- parity spot check against the compiled reference (rank 0's legs)
    check = None
    if a.check_legs > 0 and rank == 0:
        check, t = spot_check(a, L, eng, tick_host, h_o, t, h_r, h_c, first_leg, a.check_legs, 20, cycle, 20261018)

    if world > 1:
        v = torch.tensor([ms, e2e_ms], device=dev, dtype=torch.float64)
        dist.all_reduce(v, op=dist.ReduceOp.MAX)
        ms, e2e_ms = float(v[0]), float(v[1])
    ms_step = ms / a.steps
    total_legs = S * world
    value = total_legs * 10.0 / ms_step
    e2e_value = total_legs * 10.0 / (e2e_ms / e2e_steps)
    alg = L.wap_engine_algorithmic_bytes_per_frame(eng.h)
    peak, peak_src = peaks()
    dom = max(kernels, key=lambda k: k["ms_per_launch"])   # the dominant kernel of a tick
    # DRAM bytes per launch of that kernel: NOT measured in this run -- taken from the committed
    # ncu --set full capture (per leg-frame) and scaled to this run's leg count; null when none.
    traffic, traffic_src = None, None
    for name in ("r02_traffic.json", "r01_traffic.json"):
        try:
            with open(os.path.join(ROOT, "profiles", name)) as f:
                traffic = json.load(f)[dom["name"]]["dram_bytes_per_leg_frame"] * S
            traffic_src = "profile-derived, not measured in this run: profiles/%s (ncu --set full, dram__bytes_read.sum + " \
                          "dram__bytes_write.sum per leg-frame) x %d legs" % (name, S)
            break
        except Exception:
            pass
    achieved = dom["achieved_gbs"]
    whole_tick = alg * S / (ms_step * 1e-3) / 1e9
    state_bytes = L.wap_engine_state_bytes_per_stream(eng.h)
    eng.close()
    del render, capture, out, h_r, h_c, h_o_t
    torch.cuda.empty_cache()
    res = {"value": value, "ms_per_step": ms_step,
           "config": {"workload": workload_name(a), "streams_per_gpu": S, "sample_rate_hz": a.rate,
                      "frame_ms": 10, "state_bytes_per_stream": state_bytes,
                      "settle_ticks": a.settle, "input_cycle_frames": cycle,
                      "generator": "SURVEY 8(d): xorshift64* (webrtc::Random) legs, tools/wap_synth.c; the reference arm uses the same legs",
                      "l2": "per-tick state traffic (S x state) exceeds the 126 MB L2; no flush needed",
                      "parallelism": "legs sharded across GPUs, no collective"},
           "e2e": {"value": e2e_value, "unit": "real-time streams",
                   "h2d_bytes_per_step": (2 if use_render else 1) * S * fl * 2, "d2h_bytes_per_step": S * fl * 2,
                   "steps": e2e_steps, "ms_per_step": e2e_ms / e2e_steps},
           "gpu_launches": int(launches),
           "clocks": sampler.summary(),
           "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s",
                        "frac": achieved / peak, "traffic": traffic, "peak_source": peak_src,
                        "traffic_source": traffic_src,
                        "kernel": dom["name"],
                        "algorithmic_bytes_per_leg_frame": dom["algorithmic_bytes_per_leg_frame"],
                        "ms_per_launch": dom["ms_per_launch"],
                        "whole_tick": {"algorithmic_bytes_per_leg_frame": alg, "achieved": whole_tick,
                                       "frac": whole_tick / peak, "frac_of_nominal_8TBs": whole_tick / 8000.0},
                        "kernels": kernels}}
    if check is not None:
        res["parity_spot_check"] = check
    if not a.aec:
        res["e2e"]["note"] = "host-buffer path is PCIe-bound for this config (H2D+D2H bytes per tick above); value is the device-resident number"
    return res


def run_b200(a):
    import copy
    import torch
    import torch.distributed as dist
    import wap_b200
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device (no CPU fallback); use --impl reference for the CPU arm")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    L = wap_b200.load()
    res = measure(a, torch, dist, L, wap_b200, dev, local, rank, world, True)
    others = []
    default_headline = a.aec and a.ns and a.agc2_gain_db is None and a.rate == 16000
    if not a.no_other_configs and default_headline:
        # BASELINE config 3: NS-only, level high, 48 kHz three-band; config 5: the full chain with AGC2.
        for name, upd in (("BASELINE config 3: NS-only (kHigh) 48 kHz three-band",
                           dict(rate=48000, aec=0, ns=1, ns_level=2, streams=min(a.streams, 16384), settle=SETTLE_MIN)),
                          ("BASELINE config 4: stereo 48 kHz multichannel AEC3 (2 render / 2 capture channels)",
                           dict(rate=48000, aec=1, ns=0, mc=1, streams=min(a.streams, 8192), settle=MC_SETTLE)),
                          ("BASELINE config 5: full chain AEC3+NS+AGC2",
                           dict(agc2_gain_db=6.0)),
                          ("headline workload with EchoCanceller3Config::delay.use_external_delay_estimator "
                           "(hosts that report their audio buffer delay: no matched filters)",
                           dict(external_delay_estimator=1))):
            b = copy.copy(a)
            for k, v in upd.items():
                setattr(b, k, v)
            b.steps, b.warmup, b.check_legs = min(a.steps, 50), min(a.warmup, 10), min(a.check_legs, 4)
            r = measure(b, torch, dist, L, wap_b200, dev, local, rank, world, False)
            others.append({"name": name, "value": r["value"], "unit": "real-time streams", "n_gpus": world,
                           "steps": b.steps, "ms_per_step": r["ms_per_step"], "config": r["config"], "e2e": r["e2e"],
                           "roofline": {k: r["roofline"][k] for k in ("kernel", "achieved", "frac", "whole_tick")},
                           "parity_spot_check": r.get("parity_spot_check")})
    if rank == 0:
        line = {"metric": "concurrent real-time AEC3+NS streams (16 kHz, 10 ms)", "value": res["value"],
                "unit": "real-time streams", "n_gpus": world, "steps": a.steps, "warmup": a.warmup,
                "ms_per_step": res["ms_per_step"], "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
                "dtype": "f32", "data": "synthetic"}
        for k in ("config", "e2e", "gpu_launches", "clocks", "roofline", "parity_spot_check"):
            if k in res:
                line[k] = res[k]
        if others:
            line["other_configs"] = others
        if not a.no_cpu_baseline and world == 1:
            try:
                line["cpu_baseline"] = cpu_reference(a, a.cpu_seconds)
            except Exception as e:  # oracle missing on this box
                line["cpu_baseline"] = {"value": None, "unit": "real-time streams", "cores": os.cpu_count(),
                                        "kind": "reference", "sample": "unavailable: %s" % e}
        emit(line)
    if world > 1:
        dist.destroy_process_group()


def emit(line):
    """The one JSON line goes to the real stdout; everything else written to fd 1 by libraries
    (NCCL's version banner, torchrun chatter) was redirected to stderr at start-up."""
    os.write(_REAL_STDOUT, (json.dumps(line) + "\n").encode())


if __name__ == "__main__":
    sys.stdout.flush()
    _REAL_STDOUT = os.dup(1)
    os.dup2(2, 1)
    args = parse()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_b200(args)
