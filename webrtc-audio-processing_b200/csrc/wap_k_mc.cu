// Multi-channel legs (stereo frames with pipeline.multi_channel_render / _capture and AEC3: BASELINE
// config 4).  A tick is  k_mc_front -> k_delay_rt -> k_mc_echo [-> k_mc_post at 48 kHz]:
//   k_mc_front  one warp per leg: band split, content detector (may re-initialise the leg), downmix,
//               FrameBlocker, the scalar half of RenderDelayBuffer::Insert, capture high-pass filter,
//               AlignmentMixer + decimators                                           (dsp_mc_front.cuh)
//   k_delay_rt  the mono delay-estimation kernel, unchanged: it sees the mixer outputs  (wap_k_delay.cu)
//   k_mc_echo   one warp per leg: render FFTs per channel, the echo remover with R render and C capture
//               channels, BlockFramer, band merge, output                  (dsp_mc_subtractor / _remover.cuh)
//   k_mc_post   one thread per (leg, channel): PostFilter + output conversion at 48 kHz
// These kernels always read a run-time Ec3Params: a leg switches between the mono and the multichannel
// EchoCanceller3Config when the content detector changes state.
#ifdef WAP_EC3_RUNTIME
#undef WAP_EC3_RUNTIME
#endif
#define WAP_EC3_RUNTIME 1

#include "dsp_agc2.cuh"
#include "dsp_mc_front.cuh"
#include "dsp_mc_remover.cuh"
#include "wap_kernels.h"
#include "wap_launch.h"
#include "wap_pipeline.cuh"

namespace wap {

// Shared-memory footprint of one warp.
constexpr int kMcEchoBase = (int)((kAecEchoScratchBytes + 15) / 16 * 16);
constexpr int kMcEchoBytes = kMcEchoBase + (int)((sizeof(McExtra) + 15) / 16 * 16);
// Noise suppression and the band merge behind the echo remover overlay its scratch, which is dead by then:
// the bands of both channels | one NsScratch per channel, or the full-band frame + filter-bank scratch.
constexpr int kMcNsScratchFloats = (int)((sizeof(NsScratch) + 15) / 16 * 4);
static_assert(sizeof(NsScratch) % 16 == 0, "per-channel NS scratch must stay 16-byte aligned");
constexpr int kMcTailFloats = kMcCh * kFrame * kMaxBands +
                              (kMcCh * kMcNsScratchFloats > 2 * kFrame * kMaxBands ? kMcCh * kMcNsScratchFloats : 2 * kFrame * kMaxBands);
constexpr int kMcEchoFloats = kMcEchoBytes / 4 > kMcTailFloats ? kMcEchoBytes / 4 : kMcTailFloats;
constexpr int kMcFrontFloats = (int)((sizeof(McFrontScratch) + 15) / 16 * 4);

WAP_DEV void mc_stage_scalars(const Aec3Scalars& src_s, Aec3Scalars& dst_s) {
  const int* src = reinterpret_cast<const int*>(&src_s);
  int* dst = reinterpret_cast<int*>(&dst_s);
  for (int i = lane_id(); i < kScalarWords; i += 32) dst[i] = src[i];
}

// Vector half of RenderDelayBufferImpl::InsertBlock for render block r of the tick: block ring (all bands),
// FFT of [previous block | new block] and spectrum per render channel (two channels side by side).
WAP_DEV void mc_render_insert_vector(McState& mc, AecScratch& sc, const RenderInsertRec& rec, int r, int R, int B) {
  const int lane = lane_id();
  McRender& rb = mc.render;
  const McTick& mt = mc.tick;
  const int previous_write = rec.previous_write, bw = rec.blocks_write, sw = rec.spectra_write;
  __syncwarp();
  for (int rc = 0; rc < R; ++rc) {
    float* f = rc == 0 ? sc.fftA : sc.fftB;
    for (int band = 0; band < B; ++band)
      for (int i = lane; i < kBlock; i += 32) rb.blocks[bw][rc][band][i] = mt.render_blocks[r][rc][band][i];
    for (int i = lane; i < kBlock; i += 32) {
      f[i] = rb.blocks[previous_write][rc][0][i];
      f[kBlock + i] = mt.render_blocks[r][rc][0][i];
    }
  }
  fft_pair(sc, false, R > 1);
  for (int rc = 0; rc < R; ++rc) {
    const float* f = rc == 0 ? sc.fftA : sc.fftB;
    for (int k = lane; k < kBins; k += 32) {
      float re, im;
      if (k == 0) { re = f[0]; im = 0.f; }
      else if (k == 64) { re = f[1]; im = 0.f; }
      else { re = f[2 * k]; im = f[2 * k + 1]; }
      rb.fft_re[sw][rc][k] = re;
      rb.fft_im[sw][rc][k] = im;
      rb.spectra[sw][rc][k] = power_bin(re, im, k);
    }
  }
  __syncwarp();
}

// k_mc_echo body.
WAP_DEV void mc_echo_tick(const TickArgs& a, int idx, float* scratch) {
  EngineConfig cfg = a.cfg;
  const int lane = lane_id();
  const int B = cfg.num_bands;
  const int flen = kFrame * B;
  const int C = 2;
  const int slot = a.slots ? a.slots[idx] : idx;
  StreamState& st = a.states[slot];
  McState& mc = a.mc[slot];
  const TickScratch& ts = st.tick;
  McTick& mt = mc.tick;
  AecScratch& sc = *reinterpret_cast<AecScratch*>(scratch);
  McExtra& mx = *reinterpret_cast<McExtra*>(reinterpret_cast<char*>(scratch) + kMcEchoBase);
  const bool output_used = st.capture_output_used != 0;
  const bool output_used_last_frame = st.capture_output_used_last_frame != 0;
  cfg.capture_output_used = output_used ? 1 : 0;
  const bool persistent = mc.det.persistent != 0;
  const int R = persistent ? 2 : 1;
  const bool aec = cfg.aec_enabled != 0;   // stereo legs without AEC3: high-pass filter, NS, AGC2, levels only
  __syncwarp();
  if (aec) stage_ec3_params(a, sc, persistent);

  // ---------------- render side
  if (aec)
    for (int r = 0; r < ts.n_render_blocks; ++r) mc_render_insert_vector(mc, sc, ts.rins[r], r, R, B);
  if (!a.capture) return;

  // ---------------- NoiseSuppressor::Analyze on the capture frame in front of the echo canceller
  // (noise_suppressor.cc:294-386): every channel prepared, one zero-frame test and one frame counter
  // for all channels, then the per-channel analysis.  The NS scratch overlays the echo-remover scratch.
  NsState* ns = cfg.ns_enabled ? a.mc_ns + (size_t)slot * kMcCh : nullptr;
  if (ns) {
    NsScratch& nsc = *reinterpret_cast<NsScratch*>(scratch);
    float* frame = scratch + (sizeof(NsScratch) + 15) / 16 * 4;   // kMcCh x 160 floats behind the scratch
    bool nonzero = false;
    for (int c = 0; c < C; ++c) {
      ns_analyze_prepare(ns[c]);
      for (int i = lane; i < kFrame; i += 32) frame[c * kFrame + i] = mt.capture_frame[c][i];
    }
    __syncwarp();
    for (int c = 0; c < C && !nonzero; ++c) nonzero = ns_frame_nonzero(ns[c], frame + c * kFrame);
    if (nonzero) {
      int naf = ns[0].num_analyzed_frames + 1;
      if (naf < 0) naf = 0;
      __syncwarp();
      if (lane < C) ns[lane].num_analyzed_frames = naf;
      __syncwarp();
      for (int c = 0; c < C; ++c) ns_analyze_channel(ns[c], cfg, frame + c * kFrame, nsc, naf);
    }
    __syncwarp();
    if (aec) stage_ec3_params(a, sc, persistent);   // the NS scratch overlaid them
  }

  // ---------------- capture side: EchoCanceller3::ProcessCapture
  if (aec) {
  aec3_stage_scalars(st.aec, sc);
  for (int c = 0; c < C; ++c) mc_stage_scalars(mc.chan[c].s, mx.cs[c]);
  __syncwarp();
  const int final_blocks_read = sc.s.blocks_read, final_spectra_read = sc.s.spectra_read;
  int framer_len = mc.output_framer_len;
  __syncwarp();
  const int nb = ts.n_capture_blocks == 3 ? 3 : 2;
#pragma unroll 1
  for (int b = 0; b < nb; ++b) {
    __syncwarp();
    for (int c = 0; c < C; ++c)
      for (int i = lane; i < kBlock; i += 32) mx.cv[c].y[i] = mt.capture_blocks[b][c][0][i];
    const CaptureBlockRec& rec = ts.crec[b];
    if (rec.process) {
      if (lane == 0) {
        sc.s.blocks_read = rec.blocks_read;
        sc.s.spectra_read = rec.spectra_read;
      }
      __syncwarp();
      EchoPathVariability v;
      v.gain_change = rec.gain_change;
      v.delay_change = rec.delay_change;
      v.clock_drift = rec.clock_drift;
      mc_echo_remover_process_capture(st.aec, mc, cfg, sc, mx, v, sc.s.saturated_microphone_signal != 0, rec.est_has,
                                      rec.est_delay, b, R, C);
    }
    __syncwarp();
    // BlockFramer per channel and band (all framers fill in lock-step)
    for (int c = 0; c < C; ++c)
      for (int band = 0; band < B; ++band) {
        const float* blk = band == 0 ? mx.cv[c].y : mt.capture_blocks[b][c][band];
        float* fr = mc.cio[c].output_framer[band];
        if (b < 2) {
          framer_insert_and_extract_local(fr, framer_len, blk, mt.capture_frame[c] + band * kFrame + b * kSubFrame);
        } else {
          for (int i = lane; i < kBlock; i += 32) fr[i] = blk[i];
          __syncwarp();
        }
      }
    framer_len = b < 2 ? kBlock - (kSubFrame - framer_len) : kBlock;
  }
  __syncwarp();
  if (lane == 0) {
    Aec3Scalars& s = sc.s;
    mc.output_framer_len = framer_len;
    s.blocks_read = final_blocks_read;
    s.spectra_read = final_spectra_read;
    if (!s.stats_slot_full) {
      s.stats_slot_full = 1;
      s.stats_erl_time_domain = s.erl_time_domain;
      // FullBandErleEstimator::FullbandErleLog2: the minimum over the capture channels
      float e = mx.cs[0].fb_erle_time_domain_log2;
      for (int c = 1; c < C; ++c) e = fminr(e, mx.cs[c].fb_erle_time_domain_log2);
      s.stats_erle_log2 = e;
      s.stats_has_delay = 1;
      s.stats_delay_blocks = rdb_compute_delay(s);
    }
  }
  aec3_unstage_scalars(st.aec, sc);
  for (int c = 0; c < C; ++c) mc_stage_scalars(mx.cs[c], mc.chan[c].s);
  }
  __syncwarp();

  // ---------------- NoiseSuppressor::Process (noise_suppressor.cc:388-559), band merge and output
  // shared memory from here on (the echo-remover scratch is dead): bands of both channels | NS scratch per
  // channel, or full-band frame + filter-bank scratch for the merge
  float* bands_all = scratch;                                  // [C][flen]
  float* work = scratch + C * flen;
  __syncwarp();
  for (int c = 0; c < C; ++c)
    for (int i = lane; i < flen; i += 32) bands_all[c * flen + i] = mt.capture_frame[c][i];
  __syncwarp();
  if (ns) {
    NsScratch* nsc = reinterpret_cast<NsScratch*>(work);      // one per channel
    float energy_before[kMcCh], upper_gain[kMcCh];
    for (int c = 0; c < C; ++c) ns_process_front(ns[c], cfg, bands_all + c * flen, nsc[c], &energy_before[c], &upper_gain[c]);
    if (cfg.capture_output_used) {
      // AggregateWienerFilters: the minimum over the channels, applied to every channel
      __syncwarp();
      for (int i = lane; i < kNsBins; i += 32) {
        float f = nsc[0].prior[i];
        for (int c = 1; c < C; ++c) f = fminr(f, nsc[c].prior[i]);
        for (int c = 0; c < C; ++c) nsc[c].prior[i] = f;
      }
      __syncwarp();
      float gain_adjustment = 0.f, upper = upper_gain[0];
      for (int c = 0; c < C; ++c) {
        const float g = ns_process_filter(ns[c], cfg, nsc[c], energy_before[c]);
        gain_adjustment = c == 0 ? g : fminr(gain_adjustment, g);
        if (c) upper = fminr(upper, upper_gain[c]);
      }
      for (int c = 0; c < C; ++c) ns_process_finish(ns[c], cfg, bands_all + c * flen, nsc[c], gain_adjustment, upper);
    }
    __syncwarp();
  }
  const bool zero_out = !output_used_last_frame && output_used;   // first frame after un-muting
  if (lane == 0) st.capture_output_used_last_frame = output_used ? 1 : 0;
  // Band merge per channel: the merged frame of channel c replaces its bands in `bands_all`.
  if (B == 3) {
    for (int c = 0; c < C; ++c) {
      float* bands = bands_all + c * flen;
      float* full = work;
      __syncwarp();
      three_band_synthesis(bands, full, work + flen, mc.cio[c].bands.synthesis);
      __syncwarp();
      for (int i = lane; i < flen; i += 32) bands[i] = full[i];
    }
    __syncwarp();
  }
  // GainController2 (fixed gain + limiter) on the merged frames of all channels, while the output is used
  // (audio_processing_impl.cc:1450-1477), in front of the PostFilter
  if (cfg.agc2_enabled && output_used) {
    float* frames[kMcCh];
    for (int c = 0; c < C; ++c) frames[c] = bands_all + c * flen;
    agc2_process_channels(st.agc2, cfg, frames, C, flen, work);
  }
  // CaptureLevelsAdjuster::ApplyPostLevelAdjustment (audio_processing_impl.cc:1526-1528) at 16 kHz; at
  // 48 kHz it follows the PostFilter in k_mc_post.  The ramp is a serial chain (lane 0), the same for
  // every channel.
  if (cfg.levels_enabled && output_used && !(B == 3 && aec)) {
    LevelState& lv = st.levels;
    __syncwarp();
    const float prev = lv.post_prev, target = lv.post_target;
    __syncwarp();
    ScalerRun run = scaler_begin(prev, target, flen);
    if (run.mode >= 2) {
      if (lane == 0)
        for (int i = 0; i < flen; ++i) {
          run.gain = run.mode == 2 ? fminr(run.gain + run.increment, run.target) : fmaxr(run.gain + run.increment, run.target);
          work[i] = run.gain;
        }
      __syncwarp();
    }
    if (run.mode != 0)
      for (int c = 0; c < C; ++c)
        for (int i = lane; i < flen; i += 32) {
          const float g = run.mode >= 2 ? work[i] : prev;
          bands_all[c * flen + i] = fminr(fmaxr(bands_all[c * flen + i] * g, -32768.f), 32767.f);
        }
    __syncwarp();
    if (lane == 0) lv.post_prev = target;
  }
  if (B == 3 && aec && lane == 0) {   // the ramp k_mc_post applies behind the PostFilter
    LevelState& lv = st.levels;
    const bool on = cfg.levels_enabled && output_used;
    mt.post_gain_on = on ? 1 : 0;
    mt.post_gain_prev = lv.post_prev;
    mt.post_gain_target = lv.post_target;
    if (on) lv.post_prev = lv.post_target;
  }
  for (int c = 0; c < C; ++c) {
    const float* full = bands_all + c * flen;
    __syncwarp();
    if (B == 3 && aec) {
      // 48 kHz: PostFilter and the output conversion are serial work for k_mc_post
      for (int i = lane; i < flen; i += 32) mt.capture_frame[c][i] = zero_out ? 0.f : full[i];
      if (lane == 0) mt.gain_change = zero_out ? 1 : 0;   // reused as k_mc_post's "zero this frame" flag
    } else {
      for (int i = lane; i < flen; i += 32) {
        float v = zero_out ? 0.f : full[i];
        if (a.fmt == 0) {
          v = fminr(v, 32767.f);
          v = fmaxr(v, -32768.f);
          reinterpret_cast<int16_t*>(a.out)[((size_t)idx * flen + i) * C + c] = (int16_t)(v + copysignf(0.5f, v));
        } else {
          v = fminr(v, 32768.f);
          v = fmaxr(v, -32768.f);
          reinterpret_cast<float*>(a.out)[((size_t)idx * C + c) * flen + i] = v * (1.f / 32768.f);
        }
      }
    }
  }
  __syncwarp();
}

// k_mc_post body (48 kHz): PostFilter::Process (post_filter.cc:64-72) on the merged frame of channel c.
WAP_DEV void mc_post_leg(const TickArgs& a, int idx, int c) {
  const int slot = a.slots ? a.slots[idx] : idx;
  StreamState& st = a.states[slot];
  McState& mc = a.mc[slot];
  const int C = 2;
  const int flen = kFrame * 3;
  Biquad p0 = mc.post_filter[c][0], p1 = mc.post_filter[c][1], p2 = mc.post_filter[c][2], p3 = mc.post_filter[c][3];
  const bool zero = mc.tick.gain_change != 0;
  const bool used = st.capture_output_used != 0;
  // post level adjustment behind the PostFilter: both channel threads of the leg run the ramp k_mc_echo
  // recorded for this frame
  ScalerRun post = scaler_begin(mc.tick.post_gain_prev, mc.tick.post_gain_target, flen);
  if (!mc.tick.post_gain_on) post.mode = 0;
  for (int i = 0; i < flen; ++i) {
    float v = mc.tick.capture_frame[c][i];
    if (used) {
      v = biquad_step(kPostFilter48k[0], p0, v);
      v = biquad_step(kPostFilter48k[1], p1, v);
      v = biquad_step(kPostFilter48k[2], p2, v);
      v = biquad_step(kPostFilter48k[3], p3, v);
    }
    v = scaler_step(post, v);
    if (zero) v = 0.f;
    if (a.fmt == 0) {
      float w = fminr(v, 32767.f);
      w = fmaxr(w, -32768.f);
      reinterpret_cast<int16_t*>(a.out)[((size_t)idx * flen + i) * C + c] = (int16_t)(w + copysignf(0.5f, w));
    } else {
      float w = fminr(v, 32768.f);
      w = fmaxr(w, -32768.f);
      reinterpret_cast<float*>(a.out)[((size_t)idx * C + c) * flen + i] = w * (1.f / 32768.f);
    }
  }
  mc.post_filter[c][0] = p0; mc.post_filter[c][1] = p1; mc.post_filter[c][2] = p2; mc.post_filter[c][3] = p3;
}

__global__ void __launch_bounds__(128) k_mc_front(TickArgs a, int scratch_floats) {
  float* sm = reinterpret_cast<float*>(WAP_DYN_SMEM());
  const int warp = threadIdx.x >> 5;
  const int wpb = blockDim.x >> 5;
  float* scratch = sm + (size_t)warp * scratch_floats;
  for (int idx = blockIdx.x * wpb + warp; idx < a.n; idx += gridDim.x * wpb) {
    mc_front_tick(a, idx, *reinterpret_cast<McFrontScratch*>(scratch));
    __syncwarp();
  }
}

__global__ void __launch_bounds__(128) k_mc_echo(TickArgs a, int scratch_floats) {
  float* sm = reinterpret_cast<float*>(WAP_DYN_SMEM());
  const int warp = threadIdx.x >> 5;
  const int wpb = blockDim.x >> 5;
  float* scratch = sm + (size_t)warp * scratch_floats;
  for (int idx = blockIdx.x * wpb + warp; idx < a.n; idx += gridDim.x * wpb) {
    mc_echo_tick(a, idx, scratch);
    __syncwarp();
  }
}

__global__ void k_mc_post(TickArgs a) {
  const int t = blockIdx.x * blockDim.x + threadIdx.x;
  // (the two channel threads of a leg are neighbouring lanes of one warp: both exist or neither)
  if (t < a.n * 2) mc_post_leg(a, t >> 1, t & 1);
}

int k_mc_front_scratch_floats() { return kMcFrontFloats; }
int k_mc_echo_scratch_floats() { return (kMcEchoFloats + 3) & ~3; }
cudaError_t set_k_mc_smem(int front_bytes, int echo_bytes) {
  cudaError_t e = cudaSuccess;
  if (front_bytes > 48 * 1024) e = cudaFuncSetAttribute(k_mc_front, cudaFuncAttributeMaxDynamicSharedMemorySize, front_bytes);
  if (e == cudaSuccess && echo_bytes > 48 * 1024)
    e = cudaFuncSetAttribute(k_mc_echo, cudaFuncAttributeMaxDynamicSharedMemorySize, echo_bytes);
  return e;
}
cudaError_t launch_k_mc_front(int grid, int block, size_t smem, cudaStream_t stream, const TickArgs& a, int scratch_floats) {
  WAP_LAUNCH(k_mc_front, grid, block, smem, stream, a, scratch_floats);
  return cudaSuccess;
}
cudaError_t launch_k_mc_echo(int grid, int block, size_t smem, cudaStream_t stream, const TickArgs& a, int scratch_floats) {
  WAP_LAUNCH(k_mc_echo, grid, block, smem, stream, a, scratch_floats);
  return cudaSuccess;
}
cudaError_t launch_k_mc_post(cudaStream_t stream, const TickArgs& a) {
  WAP_LAUNCH(k_mc_post, (a.n * 2 + 127) / 128, 128, 0, stream, a);
  return cudaSuccess;
}

}  // namespace wap
