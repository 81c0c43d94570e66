// Per-stream frame pipeline: the body of AudioProcessingImpl::
// ProcessRenderStreamLocked / ProcessCaptureStreamLocked for the enabled
// submodules (reference audio_processing_impl.cc:1264-1561,1653-1687): the
// warp-per-leg bodies of k_delay and k_echo (k_front's body is dsp_front.cuh).
#pragma once

#include <stddef.h>

#include "dsp_aec3.cuh"
#include "dsp_agc2.cuh"
#include "dsp_echo_detector.cuh"
#include "dsp_front.cuh"
#include "dsp_filters.cuh"
#include "dsp_ns.cuh"
#include "wap_dev.cuh"
#include "wap_state.h"
#include "wap_tick.h"

namespace wap {

// Shared-memory footprint of one warp, in floats.
constexpr int kScratchFloatsDsp =
    (int)((sizeof(NsScratch) > kAecEchoScratchBytes ? sizeof(NsScratch) : kAecEchoScratchBytes) / sizeof(float)) + 4;
// k_echo: two frame buffers + the DSP scratch; k_delay: the AEC3 delay-stage scratch only.
// (rounded to 16 bytes: the scratch structs hold 128-bit aligned members)
// (a 2-band leg is laid out like a 3-band one: AEC3 writes an unused third band)
inline int echo_scratch_floats(int bands) { return (2 * kFrame * (bands == 2 ? 3 : bands) + kScratchFloatsDsp + 3) & ~3; }
// Engines without AEC3 (k_echo class kEchoNoAec): the DSP scratch only holds the noise suppressor's buffers
// or one output frame at the API rate (AGC2 / level ramps / output resampler).
constexpr int kScratchFloatsDspNoAec =
    (int)((sizeof(NsScratch) / sizeof(float) > (size_t)kRsMaxRequest ? sizeof(NsScratch) / sizeof(float) : (size_t)kRsMaxRequest)) + 4;
inline int echo_scratch_floats_no_aec(int bands) { return (2 * kFrame * (bands == 2 ? 3 : bands) + kScratchFloatsDspNoAec + 3) & ~3; }
inline int delay_scratch_floats() { return ((int)(kAecDelayScratchBytes / sizeof(float)) + 4 + 3) & ~3; }
static_assert(offsetof(StreamState, aec) % 16 == 0 && offsetof(Aec3State, mf_h) % 16 == 0 &&
                  sizeof(StreamState) % 16 == 0 && offsetof(AecScratch, mf) % 16 == 0,
              "128-bit accesses need 16-byte aligned state and scratch members");

// AudioBuffer::CopyTo: the mono frame `src` (FloatS16) to all C output channels of leg `stream`
// (int16 interleaved / float planar).
WAP_DEV void store_frame(void* dst, size_t stream, int len, int fmt, const float* src, int C = 1) {
  const int lane = lane_id();
  if (fmt == 0) {
    int16_t* p = reinterpret_cast<int16_t*>(dst) + stream * len * C;
    for (int i = lane; i < len; i += 32) {  // FloatS16ToS16 (audio_util.h:52-56)
      float v = src[i];
      v = fminr(v, 32767.f);
      v = fmaxr(v, -32768.f);
      const int16_t q = (int16_t)(v + copysignf(0.5f, v));
      for (int c = 0; c < C; ++c) p[i * C + c] = q;
    }
  } else {
    float* p = reinterpret_cast<float*>(dst) + stream * len * C;
    for (int i = lane; i < len; i += 32) {  // FloatS16ToFloat (audio_util.h:71-76)
      float v = src[i];
      v = fminr(v, 32768.f);
      v = fmaxr(v, -32768.f);
      v = v * (1.f / 32768.f);
      for (int c = 0; c < C; ++c) p[(size_t)c * len + i] = v;
    }
  }
}

// Run-time config instances: the engine's Ec3Params into the warp's scratch (kernel parameters live in
// constant memory; the DSP stages read them through `sc.ep`).
WAP_DEV void stage_ec3_params(const TickArgs& a, AecScratch& sc, bool multichannel = false) {
#if WAP_EC3_RUNTIME
  const int* src = reinterpret_cast<const int*>(multichannel ? &a.ep_mc : &a.ep);
  int* dst = reinterpret_cast<int*>(&sc.ep);
  __syncwarp();
  for (int i = lane_id(); i < (int)(sizeof(Ec3Params) / 4); i += 32) dst[i] = src[i];
  __syncwarp();
#else
  (void)a; (void)sc; (void)multichannel;
#endif
}

// k_delay body: AEC3 delay estimation of one leg for the capture blocks of this tick.
WAP_DEV void delay_stream_tick(const TickArgs& a, int idx, float* scratch) {
  const int slot = a.slots ? a.slots[idx] : idx;
  StreamState& st = a.states[slot];
  AecScratch& sc = *reinterpret_cast<AecScratch*>(scratch);
  // multi-channel legs run the multichannel EchoCanceller3Config once stereo content has been detected
  stage_ec3_params(a, sc, a.mc != nullptr && a.mc[slot].det.persistent != 0);
  aec3_delay_frame(st.aec, st.tick, sc);
}

// k_split body (48 kHz AEC3 engines): the three-band analysis of the render and the capture frame,
// which needs the lanes of a warp (in k_front it would be 480 outputs x 10 filters per thread and
// dominate the tick).  Lane 0 runs the serial capture pre-filter first.  `scratch`: 1120 floats.
WAP_DEV void split_tick(const TickArgs& a, int idx, float* scratch) {
  const EngineConfig& cfg = a.cfg;
  const int lane = lane_id();
  const int slot = a.slots ? a.slots[idx] : idx;
  StreamState& st = a.states[slot];
  UpperBandState& up = a.upper[slot];
  const int flen = 3 * kFrame;
  float* full = scratch;
  float* bands = scratch + flen;
  float* sub = scratch + 2 * flen;
  const bool render_live = !(cfg.reinit_on_first_capture && !st.seen_capture);
  if (a.render && render_live) {
    for (int i = lane; i < flen; i += 32) full[i] = front_load_sample(a.render, idx, flen, a.fmt, i, cfg.render_channels, -1);
    __syncwarp();
    three_band_analysis(full, bands, sub, st.render_bands.analysis);
    for (int i = lane; i < flen; i += 32) up.render_frame[i] = bands[i];
    __syncwarp();
  }
  if (a.capture) {
    if (lane == 0) front_capture_prefilter(a, idx);
    __syncwarp();
    for (int i = lane; i < flen; i += 32) full[i] = st.tick.capture_frame[i];
    __syncwarp();
    three_band_analysis(full, bands, sub, st.capture_bands.analysis);
    for (int i = lane; i < flen; i += 32) st.tick.capture_frame[i] = bands[i];
    __syncwarp();
  }
}

// k_resample body (engines whose API rate differs from the processing rate): AudioBuffer::CopyFrom
// with an input resampler for the render and the capture frame of one leg (audio_buffer.cc:116-160,
// 234-300).  `scratch`: 2 * kRsMaxRequest floats.
WAP_DEV void resample_in_tick(const TickArgs& a, int idx, float* scratch) {
  const EngineConfig& cfg = a.cfg;
  const int lane = lane_id();
  const int slot = a.slots ? a.slots[idx] : idx;
  const StreamState& st = a.states[slot];
  const int pf = kFrame * cfg.num_bands;
  float* src = scratch;
  float* dst = scratch + kRsMaxRequest;
  // Render frames in front of the first capture frame are lost to the re-initialisation
  // (EngineConfig::reinit_on_first_capture), and so is the resampler state they left.
  const bool render_live = !(cfg.reinit_on_first_capture && !st.seen_capture);
  const int n_in = (cfg.channels == 2 && cfg.aec_enabled) ? 3 : 2;
  for (int which = 0; which < n_in; ++which) {
    const void* in = which == 0 ? a.render : a.capture;
    if (!in || (which == 0 && (!cfg.aec_enabled || !render_live))) continue;
    // render: its own rate and channel count, averaged to mono; capture: the first buffer channel
    // (downmixed when the input has more channels than the buffer); which == 2: the second channel of
    // a stereo buffer.  A stream that already has the processing rate is only converted to FloatS16.
    const int af = which == 0 ? cfg.render_frame : cfg.api_frame;
    const int C = which == 0 ? cfg.render_channels : cfg.in_channels;
    const int ch = which == 0 ? (C == 1 ? 0 : -1) : (which == 1 ? (C == 1 ? 0 : capture_first_channel(cfg)) : 1);
    const bool resampled = which == 0 ? cfg.resample_render != 0 : cfg.resample != 0;
    float* out = (which == 0 ? a.rs_render : (which == 1 ? a.rs_capture : a.rs_capture1)) + (size_t)idx * pf;
    __syncwarp();
    if (!resampled) {
      for (int i = lane; i < pf; i += 32) out[i] = front_load_sample(in, idx, pf, a.fmt, i, C, ch);
      continue;
    }
    const ResamplerParams p{af, pf, which == 0 ? a.rs_ratio_render : a.rs_ratio_in, which == 0 ? a.rs_kernel_render : a.rs_kernel_in};
    for (int i = lane; i < af; i += 32) src[i] = load_raw_sample(in, idx, af, a.fmt, i, C, ch);
    __syncwarp();
    rs_push(which == 2 ? a.extra[slot].rs : a.rs[slot * kRsPerLeg + which], p, src, dst);
    for (int i = lane; i < pf; i += 32) {
      float v = dst[i];
      if (a.fmt == 1) {  // FloatToFloatS16 after the resampler (audio_buffer.cc:150-155)
        v = fminr(v, 1.f);
        v = fmaxr(v, -1.f);
        v = v * 32768.f;
      }
      out[i] = v;
    }
  }
}

// k_echo body: everything after the front end for one leg (reference
// audio_processing_impl.cc:1359-1448 for the enabled submodules).
// Config classes with their own k_echo instance: the fields below become compile-time constants in
// it, so the band-split / upper-band / resampler / stereo code a class does not use is not even in
// the kernel (instruction footprint is what k_echo is sensitive to).  Class 0 is the generic one.
enum EchoClass {
  kEchoGeneric = 0,
  kEchoMono16k = 1,      // 16 kHz mono, native (the bench workload; with or without AGC2 / levels)
  kEchoMono48kNative = 2,  // 48 kHz mono, three bands
  kEchoMono48kVia32k = 3,  // 48 kHz mono under the default maximum_internal_processing_rate
  kEchoMono32k = 4,      // 32 kHz mono, two bands
  kEchoNoAec = 5,        // mono legs without AEC3 (BASELINE config 3: NS-only): no AEC3 code, registers or scratch
  kEchoClasses = 6
};
inline int echo_class_of(const EngineConfig& c) {
  // (k_echo starts behind the front end: only the output side of the formats matters to it)
  if (c.channels != 1 || c.in_channels != 1) return kEchoGeneric;
  if (!c.aec_enabled) return kEchoNoAec;
  if (c.num_bands == 1 && !c.resample_out) return kEchoMono16k;
  if (!c.split_bands) return kEchoGeneric;
  if (c.num_bands == 3 && !c.resample_out) return kEchoMono48kNative;
  if (c.num_bands == 2 && c.fullband_out) return kEchoMono48kVia32k;
  if (c.num_bands == 2 && !c.resample_out && !c.fullband_out) return kEchoMono32k;
  return kEchoGeneric;
}

template <int kClass>
WAP_DEV void echo_stream_tick(const TickArgs& a, int idx, float* scratch) {
  EngineConfig cfg = a.cfg;
  if (kClass == kEchoNoAec) {
    cfg.channels = 1;
    cfg.in_channels = 1;
    cfg.aec_enabled = 0;
  } else if (kClass != kEchoGeneric) {
    cfg.channels = 1;
    cfg.in_channels = 1;
    cfg.num_bands = kClass == kEchoMono16k ? 1 : (kClass == kEchoMono48kNative ? 3 : 2);
    cfg.split_bands = kClass == kEchoMono16k ? 0 : 1;
    cfg.resample_out = 0;
    cfg.fullband_out = kClass == kEchoMono48kVia32k ? 1 : 0;
    if (kClass == kEchoMono48kVia32k) cfg.out_frame = 480;
  }
  const int B = cfg.num_bands;
  const int flen = kFrame * B;
#if WAP_ECHO_LOCKSTEP
  WAP_PHASE_SYNC();
  if (idx < 0) {  // no leg for this warp in this trip: pass the same phase points
    if (!a.capture) return;
    WAP_PHASE_SYNC();
    if (cfg.aec_enabled) { WAP_PHASE_SYNC(); WAP_PHASE_SYNC(); WAP_PHASE_SYNC(); WAP_PHASE_SYNC(); }
    WAP_PHASE_SYNC();
    return;
  }
#endif
  const int slot = a.slots ? a.slots[idx] : idx;
  StreamState& st = a.states[slot];
  const TickScratch& ts = st.tick;
  // set_output_will_be_muted / kCaptureOutputUsed are per leg.
  const bool output_used = st.capture_output_used != 0;
  const bool output_used_last_frame = st.capture_output_used_last_frame != 0;
  cfg.capture_output_used = output_used ? 1 : 0;
  __syncwarp();
  const int lay = kFrame * (B == 2 ? 3 : B);  // buffer stride: see echo_scratch_floats
  float* full = scratch;
  float* bands = (B == 1 || !cfg.split_bands) ? full : scratch + lay;
  unsigned dsp_off = 2u * (unsigned)lay;
#if !defined(WAP_EMU)
  asm volatile("" : "+r"(dsp_off));  // see k_echo: one register instead of re-reading cfg at every access
#endif
  void* dsp = scratch + dsp_off;
  NsScratch& ns_sc = *reinterpret_cast<NsScratch*>(dsp);
  AecScratch& aec_sc = *reinterpret_cast<AecScratch*>(dsp);

  // Start pulling in what the later phases stream: the noise suppressor's vectors (up to the
  // histograms) and the two adaptive filters.
  if (a.capture) {
    if (cfg.ns_enabled) warp_prefetch_l2(&st.ns, (int)offsetof(NsState, hist_lrt));
    if (cfg.aec_enabled)
      warp_prefetch_l2(st.aec.Hr_re, (int)(reinterpret_cast<const char*>(st.aec.h_time) - reinterpret_cast<const char*>(st.aec.Hr_re)));
  }
  // ---------------- render side: ring / FFT / spectrum writes for the blocks k_front sliced
  UpperBandState* up = (B >= 2 && cfg.aec_enabled) ? &a.upper[slot] : nullptr;
  if (cfg.aec_enabled && ts.n_render_blocks > 0) {
    stage_ec3_params(a, aec_sc);
    aec3_echo_render(st.aec, ts, aec_sc, up);
  }
  if (!a.capture) return;
  WAP_PHASE_SYNC();

  // ---------------- capture side (the frame is already high-pass filtered)
  if (up) {  // k_front already split the capture frame (it needs band 0 for the blocks)
    for (int i = lane_id(); i < flen; i += 32) bands[i] = ts.capture_frame[i];
    __syncwarp();
  } else {
    for (int i = lane_id(); i < flen; i += 32) full[i] = ts.capture_frame[i];
    __syncwarp();
    if (cfg.split_bands) {
      if (B == 3) three_band_analysis(full, bands, reinterpret_cast<float*>(dsp), st.capture_bands.analysis);
      else two_band_analysis(full, bands, reinterpret_cast<float*>(dsp), &st.capture_bands.analysis[0][0]);
    }
  }
  if (cfg.ns_enabled) ns_analyze(st.ns, cfg, bands, ns_sc);
  if (cfg.aec_enabled) {
    WAP_PHASE_SYNC();
    stage_ec3_params(a, aec_sc);   // the noise suppressor's scratch overlays the AEC3 scratch
    aec3_echo_capture(st.aec, cfg, bands, ts, aec_sc, up, &st.ns, (int)offsetof(NsState, hist_lrt));   // three more phase points, one per block slot
  }
  WAP_PHASE_SYNC();
  if (cfg.ns_enabled) ns_process(st.ns, cfg, bands, ns_sc);
  if (cfg.split_bands) {
    if (B == 3) three_band_synthesis(bands, full, reinterpret_cast<float*>(dsp), st.capture_bands.synthesis);
    else two_band_synthesis(bands, full, reinterpret_cast<float*>(dsp), &st.capture_bands.synthesis[0][0]);
  }
  __syncwarp();
#if WAP_L2_POLICY != 0
  // the leg's AEC3 / NS state is dead until its next tick
  if (cfg.ns_enabled) warp_l2_release(&st.ns, (int)offsetof(NsState, hist_lrt));
  if (cfg.aec_enabled)
    warp_l2_release(st.aec.Hr_re, (int)(reinterpret_cast<const char*>(st.aec.h_time) - reinterpret_cast<const char*>(st.aec.Hr_re)));
#endif
  const int slot_rs = slot * kRsPerLeg;
  float* tmp = reinterpret_cast<float*>(dsp);
  int olen = flen;  // samples in `full` from here on
  if (cfg.fullband_out) {
    // capture_fullband_audio (audio_processing_impl.cc:1245-1253,1451-1460): a 48 kHz buffer that is
    // refreshed from the processed frame (resampled, AudioBuffer::CopyTo(AudioBuffer*)) only while
    // the output is used; otherwise it still holds the unprocessed input frame.
    olen = cfg.out_frame;
    if (output_used) {
      const ResamplerParams p{flen, olen, a.rs_ratio_out, a.rs_kernel_out};
      rs_push(a.rs[slot_rs + 2], p, full, tmp);
      for (int i = lane_id(); i < olen; i += 32) full[i] = tmp[i];
    } else {
      // Muted: the (multi-channel) input frame comes back, channel by channel -- downmixed like the
      // capture_fullband_audio buffer's CopyFrom when the output has fewer channels than the input
      // (this path requires input rate == output rate).
#if WAP_EC3_RUNTIME   // engines with the detector run the run-time-parameter kernel instances
      if (a.red) red_capture_tick(a.red[slot], full, olen, false, tmp);
#endif
      if (cfg.in_channels != cfg.channels) {
        for (int i = lane_id(); i < olen; i += 32) {
          float v = load_raw_sample(a.capture, idx, olen, a.fmt, i, cfg.in_channels, capture_first_channel(cfg));
          if (a.fmt == 0) {
            reinterpret_cast<int16_t*>(a.out)[(size_t)idx * olen + i] = (int16_t)v;   // an integer already
          } else {
            v = fmaxr(fminr(v, 1.f), -1.f);
            reinterpret_cast<float*>(a.out)[(size_t)idx * olen + i] = v * 32768.f * (1.f / 32768.f);
          }
        }
        if (lane_id() == 0) st.capture_output_used_last_frame = 0;
        return;
      }
      const int total = olen * cfg.channels;
      for (int i = lane_id(); i < total; i += 32) {
        if (a.fmt == 0) {
          reinterpret_cast<int16_t*>(a.out)[(size_t)idx * total + i] = reinterpret_cast<const int16_t*>(a.capture)[(size_t)idx * total + i];
        } else {
          float v = reinterpret_cast<const float*>(a.capture)[(size_t)idx * total + i];
          v = fmaxr(fminr(v, 1.f), -1.f);
          reinterpret_cast<float*>(a.out)[(size_t)idx * total + i] = v * 32768.f * (1.f / 32768.f);
        }
      }
      if (lane_id() == 0) st.capture_output_used_last_frame = 0;
      return;
    }
    __syncwarp();
  }
  // The residual echo detector looks at the merged frame (audio_processing_impl.cc:1462-1465)
  // (only in the run-time-parameter kernel instances, which engines with the detector are switched to: the
  // hook costs the default-config kernels one more spilled register, 0.3 % of k_echo, measured)
#if WAP_EC3_RUNTIME
  if (a.red) red_capture_tick(a.red[slot], full, olen, output_used, tmp);
#endif
  // GainController2 runs on the merged full-band frame, only while the output is used
  // (audio_processing_impl.cc:1450-1477), before the PostFilter.
  if (cfg.agc2_enabled && output_used) agc2_process(st.agc2, cfg, full, olen, tmp);
  const bool post = up && B == 3;  // 48 kHz AEC3: PostFilter, post gain and output conversion in k_post
  // CaptureLevelsAdjuster::ApplyPostLevelAdjustment (audio_processing_impl.cc:1526-1528), while the
  // output is used: the gain ramp is a serial chain (lane 0), the multiply and clamp are not.
  if (cfg.levels_enabled && output_used && !post) {
    LevelState& lv = st.levels;
    __syncwarp();
    const float prev = lv.post_prev, target = lv.post_target;
    __syncwarp();
    ScalerRun run = scaler_begin(prev, target, olen);
    if (run.mode >= 2) {
      if (lane_id() == 0)
        for (int i = 0; i < olen; ++i) {
          run.gain = run.mode == 2 ? fminr(run.gain + run.increment, run.target) : fmaxr(run.gain + run.increment, run.target);
          tmp[i] = run.gain;
        }
      __syncwarp();
    }
    if (run.mode != 0)
      for (int i = lane_id(); i < olen; i += 32) {
        const float g = run.mode >= 2 ? tmp[i] : prev;
        full[i] = fminr(fmaxr(full[i] * g, -32768.f), 32767.f);
      }
    __syncwarp();
    if (lane_id() == 0) lv.post_prev = target;
  }
  // Output is zeroed for the first frame after un-muting (audio_processing_impl.cc:1540-1552).
  if (!post && !output_used_last_frame && output_used) {
    for (int i = lane_id(); i < olen; i += 32) full[i] = 0.f;
    __syncwarp();
  }
  if (lane_id() == 0) st.capture_output_used_last_frame = output_used ? 1 : 0;
  if (post) {
    // 48 kHz AEC3: PostFilter (a serial IIR) and the output conversion run in k_post.
    for (int i = lane_id(); i < flen; i += 32) st.tick.capture_frame[i] = full[i];
    if (lane_id() == 0) st.tick.pad_[0] = (!output_used_last_frame && output_used) ? 1 : 0;
    return;
  }
  if (cfg.resample_out) {
    // AudioBuffer::CopyTo with an output resampler (audio_buffer.cc:156-176,314-372): the int16
    // interface resamples FloatS16 data and then rounds, the float interface scales first.
    const int alen = cfg.out_frame;
    __syncwarp();
    if (a.fmt == 1) {
      for (int i = lane_id(); i < flen; i += 32) {  // FloatS16ToFloat (audio_util.h:71-76)
        float v = fminr(full[i], 32768.f);
        v = fmaxr(v, -32768.f);
        full[i] = v * (1.f / 32768.f);
      }
    }
    const ResamplerParams p{flen, alen, a.rs_ratio_out, a.rs_kernel_out};
    rs_push(a.rs[slot_rs + 2], p, full, tmp);
    if (a.fmt == 0) {
      store_frame(a.out, idx, alen, 0, tmp, cfg.channels);
    } else {
      float* o = reinterpret_cast<float*>(a.out) + (size_t)idx * alen * cfg.channels;
      for (int i = lane_id(); i < alen; i += 32)
        for (int c = 0; c < cfg.channels; ++c) o[(size_t)c * alen + i] = tmp[i];
    }
    return;
  }
  store_frame(a.out, idx, olen, a.fmt, full, cfg.channels);
}

}  // namespace wap
