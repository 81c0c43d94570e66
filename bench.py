#!/usr/bin/env python3
"""bench.py -- concurrent real-time AEC3+NS streams per GPU (16 kHz, 10 ms frames).

One "step" = one 10 ms tick of the hot path over all S call legs of this GPU
(ProcessReverseStream + set_stream_delay_ms(0) + ProcessStream per leg, i.e. one
wap_process_streams call = k_front + k_delay + k_echo).  metric = legs that can be served
in real time = S * 10 ms / tick time, summed over GPUs (legs shard across GPUs
with no collective: "scaling": "weak").

  value : device-resident timing (int16 frames already in HBM, wap_process_streams_device)
  e2e   : same metric through the host-buffer C ABI (wap_process_streams), H2D of the
          render+capture frames and D2H of the output inside the timed region
  roofline : for the dominant kernel of the tick: its algorithmic bytes (SURVEY.md 8(d) byte
             model split per kernel) * S / its launch duration (CUDA events around each kernel
             on the engine's stream) against the measured HBM copy bandwidth
             (MEASURED_PEAKS.json); "whole_tick" gives the same for all three kernels together
  cpu_baseline : the compiled reference (oracle/_ref) on the host cores, one
             AudioProcessing instance per leg, bounded sample

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

RATE = 16000
FL = RATE // 100
CYCLE = 64  # distinct synthetic frames per leg before the input repeats


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=300)
    ap.add_argument("--warmup", type=int, default=100)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--streams", type=int, default=int(os.environ.get("WAP_BENCH_STREAMS", "65536")),
                    help="call legs per GPU")
    ap.add_argument("--aec", type=int, default=1)
    ap.add_argument("--ns", type=int, default=1)
    ap.add_argument("--ns-level", type=int, default=1)
    ap.add_argument("--agc2-gain-db", type=float, default=None,
                    help="add GainController2 (fixed gain + limiter) to the chain: BASELINE config 5 (no CPU arm)")
    ap.add_argument("--max-rate", type=int, default=48000, choices=[32000, 48000],
                    help="pipeline.maximum_internal_processing_rate (32000 = the reference default: 48 kHz legs are resampled)")
    ap.add_argument("--rate", type=int, default=16000, choices=[16000, 32000, 48000],
                    help="native sample rate of the legs (BASELINE config 3: --rate 48000 --aec 0 --ns-level 2)")
    ap.add_argument("--cpu-seconds", type=float, default=12.0, help="budget of the cpu_baseline sample")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    a = ap.parse_args()
    global RATE, FL
    RATE, FL = a.rate, a.rate // 100
    return a


def workload_name(a):
    parts = []
    if a.aec:
        parts.append("AEC3(default EchoCanceller3Config)")
    if a.ns:
        parts.append("NS(%s)" % ["low", "moderate", "high", "veryhigh"][a.ns_level])
    if getattr(a, "agc2_gain_db", None) is not None:
        parts.append("AGC2(fixed %g dB + limiter)" % a.agc2_gain_db)
    if getattr(a, "max_rate", 48000) == 32000 and a.rate == 48000:
        parts.append("processed at 32 kHz (default maximum_internal_processing_rate)")
    return "%d synthetic mono %d kHz call legs per GPU, %s, 10 ms frames" % (a.streams, a.rate // 1000, "+".join(parts))


def peaks():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


# ------------------------------------------------------------------ CPU reference leg
def synthetic_cpu(n_legs, n_frames):
    """int16 render/capture [leg][frame*160] from the SURVEY 8(d) generator."""
    import numpy as np
    from common import synthetic_leg
    r = np.zeros((n_legs, n_frames * FL), np.int16)
    c = np.zeros((n_legs, n_frames * FL), np.int16)
    for i in range(n_legs):
        r[i], c[i] = synthetic_leg(i, n_frames, RATE)
    return r, c


def cpu_reference(a, seconds):
    """Times oracle/_ref: threads = host cores, one AudioProcessing instance per leg."""
    import ref
    cores = os.cpu_count() or 1
    # ~150 us per leg-frame per core: size the sample to `seconds`.
    per_leg_frames = 400
    warm = 100
    legs_per_thread = max(1, int(seconds / (per_leg_frames * 160e-6)))
    legs = cores * legs_per_thread
    r, c = synthetic_cpu(min(legs, 2 * cores), per_leg_frames)
    # legs beyond the generated ones reuse the same audio (stride wraps): use stride 0 groups
    import numpy as np
    reps = (legs + r.shape[0] - 1) // r.shape[0]
    r = np.tile(r, (reps, 1))[:legs].copy()
    c = np.tile(c, (reps, 1))[:legs].copy()
    secs = ref.cpu_bench(a.aec, a.ns, a.ns_level, RATE, legs, cores, warm, per_leg_frames, r, c,
                         stride=r.shape[1])
    frames = legs * (per_leg_frames - warm)
    streams_rt = frames / secs / 100.0
    return {"value": streams_rt, "unit": "real-time streams", "cores": cores, "kind": "reference",
            "sample": "%d legs x %d timed frames (after %d warm-up) of the same synthetic workload, "
                      "one webrtc::AudioProcessing per leg, %d pinned threads, %.1f s wall"
                      % (legs, per_leg_frames - warm, warm, cores, secs),
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


def make_inputs(torch, dev, S, seed):
    """Render/capture int16 [CYCLE][S][160] on the device: white-noise render (amplitude
    8000, so every adaptive filter updates: worst-case work), 3-tap echo path with a
    per-leg delay, noise floor + periodic double talk.  The cycle is
    circular so it can repeat without a discontinuity in the echo path."""
    g = torch.Generator(device=dev)
    g.manual_seed(seed)
    n = CYCLE * FL
    x = (torch.rand((S, n), device=dev, generator=g) * 2 - 1) * 8000.0
    i = torch.arange(S, device=dev)
    D = 64 * (1 + (i % 48)) + (7 * i) % 64
    idx = torch.arange(n, device=dev)[None, :]
    y = torch.zeros_like(x)
    for gain, extra in ((0.5, 0), (0.25, 37), (0.1, 160)):
        y += gain * torch.gather(x, 1, (idx - (D[:, None] + extra)) % n)
    y += (torch.rand((S, n), device=dev, generator=g) * 2 - 1) * 50.0
    burst = ((idx % (32 * FL)) >= 27 * FL).float()
    y += (torch.rand((S, n), device=dev, generator=g) * 2 - 1) * 3000.0 * burst
    to16 = lambda t: t.round().clamp(-32768, 32767).to(torch.int16).view(S, CYCLE, FL).permute(1, 0, 2).contiguous()
    return to16(x), to16(y)


def run_b200(a):
    import numpy as np
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
    S = a.streams
    extra = {} if a.agc2_gain_db is None else dict(agc2=True, agc2_fixed_gain_db=a.agc2_gain_db)
    eng = wap_b200.Engine(S, RATE, lib=L, device=local, aec=bool(a.aec), ns=bool(a.ns), ns_level=a.ns_level, max_rate=a.max_rate, **extra)
    render, capture = make_inputs(torch, dev, S, 1234 + rank)
    out = torch.empty((S, FL), dtype=torch.int16, device=dev)
    stream = torch.cuda.ExternalStream(L.wap_engine_cuda_stream(eng.h), device=dev)
    torch.cuda.synchronize()
    eng.set_stream_delay_ms(0)

    def tick_device(t):
        k = t % CYCLE
        L.wap_streams_set_delay_ms(eng.handles, S, 0)   # set_stream_delay_ms(0) per leg, as the reference arm
        err = L.wap_process_streams_device(eng.h, eng.handles, S, render[k].data_ptr(), capture[k].data_ptr(),
                                           out.data_ptr(), 0)
        assert err == 0, err

    def barrier():
        torch.cuda.synchronize()
        L.wap_engine_synchronize(eng.h)
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ---- device-resident timing
    t = 0
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
    for name, m, bts in zip(("k_front", "k_delay", "k_echo"), kms, kbytes):
        per = m / max(1, n_timed)
        kernels.append({"name": name, "ms_per_launch": per, "algorithmic_bytes_per_leg_frame": bts,
                        "achieved_gbs": (bts * S / (per * 1e-3) / 1e9) if per > 0 else 0.0})

    # ---- end to end through the host-buffer ABI (pinned host frames in, result out)
    h_r = render.cpu().pin_memory().numpy()
    h_c = capture.cpu().pin_memory().numpy()
    h_o = torch.empty((S, FL), dtype=torch.int16).pin_memory().numpy()
    e2e_steps = max(10, min(a.steps, 100))

    def tick_host(tt):
        k = tt % CYCLE
        L.wap_streams_set_delay_ms(eng.handles, S, 0)
        err = L.wap_process_streams(eng.handles, S, h_r[k].ctypes.data_as(C.c_void_p), h_c[k].ctypes.data_as(C.c_void_p),
                                    h_o.ctypes.data_as(C.c_void_p), 0, None)
        assert err == 0, err
    for _ in range(3):
        tick_host(t); t += 1
    barrier()
    w0 = time.perf_counter()
    for _ in range(e2e_steps):
        tick_host(t); t += 1
    barrier()
    e2e_ms = (time.perf_counter() - w0) * 1e3

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
    # DRAM bytes per launch of that kernel from the committed ncu --set full capture (per leg-frame,
    # scaled to this run's leg count); null when no capture is on record.
    traffic = None
    try:
        with open(os.path.join(ROOT, "profiles", "r01_traffic.json")) as f:
            traffic = json.load(f)[dom["name"]]["dram_bytes_per_leg_frame"] * S
    except Exception:
        pass
    achieved = dom["achieved_gbs"]
    whole_tick = alg * S / (ms_step * 1e-3) / 1e9
    state_bytes = L.wap_engine_state_bytes_per_stream(eng.h)
    eng.close()
    if rank == 0:
        line = {"metric": "concurrent real-time AEC3+NS streams (16 kHz, 10 ms)", "value": value,
                "unit": "real-time streams", "n_gpus": world, "steps": a.steps, "warmup": a.warmup,
                "ms_per_step": ms_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
                "dtype": "f32", "data": "synthetic",
                "config": {"workload": workload_name(a), "streams_per_gpu": S, "sample_rate_hz": RATE,
                           "frame_ms": 10, "state_bytes_per_stream": state_bytes,
                           "l2": "per-tick state traffic (S x state) exceeds the 126 MB L2; no flush needed",
                           "parallelism": "legs sharded across GPUs, no collective"},
                "e2e": {"value": e2e_value, "unit": "real-time streams",
                        "h2d_bytes_per_step": 2 * S * FL * 2, "d2h_bytes_per_step": S * FL * 2,
                        "steps": e2e_steps, "ms_per_step": e2e_ms / e2e_steps},
                "gpu_launches": int(launches),
                "clocks": sampler.summary(),
                "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s",
                             "frac": achieved / peak, "traffic": traffic, "peak_source": peak_src,
                             "traffic_source": "profiles/r01_traffic.json (ncu dram__bytes_read.sum + dram__bytes_write.sum)",
                             "kernel": dom["name"],
                             "algorithmic_bytes_per_leg_frame": dom["algorithmic_bytes_per_leg_frame"],
                             "ms_per_launch": dom["ms_per_launch"],
                             "whole_tick": {"algorithmic_bytes_per_leg_frame": alg, "achieved": whole_tick,
                                            "frac": whole_tick / peak, "frac_of_nominal_8TBs": whole_tick / 8000.0},
                             "kernels": kernels}}
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
