// Front end of a tick, ONE THREAD PER CALL LEG (k_front): everything that is a
// serial recurrence over samples and therefore cannot use the lanes of a warp
// within one leg -- the capture high-pass filter and the two AEC3 decimators --
// plus the frame -> block slicing and the scalar bookkeeping of
// RenderDelayBufferImpl::Insert that goes with the render blocks.  32 legs share a
// warp here, so the IIR chains cost 1/32 of what they cost in a warp-per-leg
// kernel; the per-leg results are handed to k_delay / k_echo through TickScratch.
//   HighPassFilter::Process                 high_pass_filter.cc:90-113
//   CascadedBiQuadFilter::ApplyBiQuad       utility/cascaded_biquad_filter.cc:58-84
//   Decimator::Decimate                     aec3/decimator.cc:75-91
//   FrameBlocker                            aec3/frame_blocker.cc:40-84
//   RenderDelayBufferImpl::Insert           aec3/render_delay_buffer.cc:200-242
//   EchoCanceller3::AnalyzeCapture          aec3/echo_canceller3.cc:862-874
#pragma once

#include "dsp_aec3_render.cuh"
#include "dsp_filters.cuh"
#include "wap_dev.cuh"
#include "wap_state.h"
#include "wap_tick.h"

namespace wap {

// One section, one sample: y = b0*x + b1*x1 + b2*x2 - a0*y1 - a1*y2 (left to right).
WAP_DEV float biquad_step(const BiquadCoef& c, Biquad& m, float x) {
  const float y = c.b0 * x + c.b1 * m.x0 + c.b2 * m.x1 - c.a0 * m.y0 - c.a1 * m.y1;
  m.x1 = m.x0;
  m.x0 = x;
  m.y1 = m.y0;
  m.y0 = y;
  return y;
}

// S16ToFloatS16 / FloatToFloatS16 (audio_util.h:52-69) for sample i of leg `leg`.
// Sample i of leg `leg`'s API frame before any scaling (int16 interleaved / float planar; len =
// samples per channel).  ch >= 0: that channel; ch < 0: AudioBuffer::CopyFrom's downmix by
// averaging (audio_buffer.cc:116-140 float: sum in channel order times 1/C; :256-264 int16:
// int32 sum, integer division).
WAP_DEV float load_raw_sample(const void* src, size_t leg, int len, int fmt, int i, int C, int ch) {
  if (fmt == 0) {
    const int16_t* p = reinterpret_cast<const int16_t*>(src) + leg * (size_t)len * C + (size_t)i * C;
    if (ch >= 0) return (float)p[ch];
    int sum = 0;
    for (int c = 0; c < C; ++c) sum += p[c];
    return (float)(sum / C);
  }
  const float* p = reinterpret_cast<const float*>(src) + leg * (size_t)len * C + i;
  if (ch >= 0) return p[(size_t)ch * len];
  float v = p[0];
  for (int c = 1; c < C; ++c) v += p[(size_t)c * len];
  return v * (1.f / C);
}
// The same as a FloatS16 sample (S16ToFloatS16 / FloatToFloatS16, audio_util.h).
WAP_DEV float front_load_sample(const void* src, size_t leg, int len, int fmt, int i, int C = 1, int ch = 0) {
  if (fmt == 2) return reinterpret_cast<const float*>(src)[leg * len + i];  // already FloatS16 (k_resample)
  float v = load_raw_sample(src, leg, len, fmt, i, C, C == 1 ? 0 : ch);
  if (fmt == 0) return v;
  v = fminr(v, 1.f);
  v = fmaxr(v, -1.f);
  return v * 32768.f;
}

// Channel selection for the first channel of the capture AudioBuffer (audio_buffer.cc:116-140,234-300):
// an input with more channels than the buffer is downmixed on the way in -- the average of all input
// channels, or the first one with capture_downmix_method = UseFirstChannel; otherwise channel 0.
WAP_DEV int capture_first_channel(const EngineConfig& cfg) {
  return (cfg.in_channels > cfg.channels && !cfg.downmix_first) ? -1 : 0;
}

// RenderDelayBufferImpl::Insert for render block `x` (64 samples, thread-local):
// scalar bookkeeping, decimation into the low-rate ring, and the record that tells
// k_echo where the block, its FFT and its spectrum go.
WAP_DEV void front_render_insert(Aec3State& a, TickScratch& ts, int r, const float* x, const Ec3Params& ep) {
  Aec3Scalars& s = a.s;
  float x_energy = 0.f;  // DetectActiveRender: std::inner_product
  for (int i = 0; i < kBlock; ++i) x_energy += x[i] * x[i];
  if (ep.use_external_delay_estimator) ++s.rdb_render_calls;
  if (s.has_delay) {
    if (!s.last_call_was_render) {
      s.last_call_was_render = 1;
      s.num_api_calls_in_a_row = 1;
    } else if (++s.num_api_calls_in_a_row > s.max_observed_jitter) {
      s.max_observed_jitter = s.num_api_calls_in_a_row;
    }
  }
  const int previous_write = s.blocks_write;
  // IncrementWriteIndices (:455-460)
  s.lr_write = ring_off(s.lr_write, -kSubBlock, kLowRateSize);
  s.blocks_write = ring_inc(s.blocks_write, kRingBlocks);
  s.spectra_write = ring_dec(s.spectra_write, kRingBlocks);
  // RenderOverrun (:481-483); BlockProcessorImpl::BufferRender keeps the last event.
  s.render_event = (s.lr_read == s.lr_write || s.blocks_read == s.blocks_write) ? kEventRenderOverrun : kEventNone;
  if (!s.render_activity) {
    s.render_activity_counter += (x_energy > (ep.active_render_limit * ep.active_render_limit) * 64.f) ? 1 : 0;
    s.render_activity = s.render_activity_counter >= 20;
  }
  ts.rins[r].blocks_write = s.blocks_write;
  ts.rins[r].spectra_write = s.spectra_write;
  ts.rins[r].previous_write = previous_write;
  // InsertBlock: decimate and store the sub-block reversed at the low-rate write index.
  Biquad d0 = a.render_decimator[0], d1 = a.render_decimator[1], d2 = a.render_decimator[2], d3 = a.render_decimator[3];
  const int lw = s.lr_write;
  // render_levels.render_power_gain_db (render_delay_buffer.cc:124-125,405-414): the stored block, and
  // everything derived from it, is scaled; the activity detection above saw the block as it came
  const float rgain = ep.render_linear_amplitude_gain;
  for (int i = 0; i < kBlock; ++i) {
    const float xi = rgain != 1.f ? x[i] * rgain : x[i];
    ts.render_blocks[r][i] = xi;
    float v = biquad_step(kDecimator4[0], d0, xi);
    v = biquad_step(kDecimator4[1], d1, v);
    v = biquad_step(kDecimator4[2], d2, v);
    v = biquad_step(kDecimator4[3], d3, v);
    if ((i & (kDownSampling - 1)) == 0) a.low_rate[lw + kSubBlock - 1 - (i >> 2)] = v;
  }
  a.render_decimator[0] = d0; a.render_decimator[1] = d1; a.render_decimator[2] = d2; a.render_decimator[3] = d3;
  if (s.render_event != kEventNone) rdb_reset(s, ep.default_delay);
  s.render_properly_started = 1;
}

// FrameBlocker for one upper band (no decimation, no bookkeeping): the continuous stream
// [blocker buffer | 160 new samples] cut into `nb` blocks; the rest stays buffered.
WAP_DEV void front_slice_band(const float* band, float* blocker, int L, int nb, float (*blocks)[2][kBlock], int bi) {
  for (int b = 0; b < nb; ++b)
    for (int j = 0; j < kBlock; ++j) {
      const int k = kBlock * b + j - L;
      blocks[b][bi][j] = k < 0 ? blocker[L + k] : band[k];
    }
  const int rem = L + kFrame - kBlock * nb;
  for (int j = 0; j < rem; ++j) blocker[j] = band[kFrame - rem + j];
}

// AudioSamplesScaler::Process for one sample: `mode` 0 = gain 1 throughout (nothing, not even the
// clamp), 1 = constant gain, 2 = rising ramp, 3 = falling ramp (audio_samples_scaler.cc:25-90).
struct ScalerRun {
  int mode;
  float gain, increment, target;
};
WAP_DEV ScalerRun scaler_begin(float prev, float target, int samples) {
  ScalerRun r;
  r.gain = prev;
  r.target = target;
  r.increment = 0.f;
  if (target == 1.f && prev == target) r.mode = 0;
  else if (prev == target) r.mode = 1;
  else {
    r.increment = (target - prev) * (1.f / samples);
    r.mode = r.increment > 0.f ? 2 : 3;
  }
  return r;
}
WAP_DEV float scaler_step(ScalerRun& r, float v) {
  if (r.mode == 0) return v;
  if (r.mode == 2) r.gain = fminr(r.gain + r.increment, r.target);
  else if (r.mode == 3) r.gain = fmaxr(r.gain + r.increment, r.target);
  v *= r.gain;
  return fminr(fmaxr(v, -32768.f), 32767.f);  // SafeClamp
}

// Capture side in front of the band split: full-band high-pass filter, pre level adjustment,
// AEC3's saturation test (both channels of a stereo leg), echo-path gain-change flag.  Writes the
// filtered frame to ts.capture_frame.  One thread (k_front, or lane 0 of k_split).
WAP_DEV void front_capture_prefilter(const TickArgs& a, int idx) {
  const EngineConfig& cfg = a.cfg;
  const int flen = kFrame * cfg.num_bands;
  const int slot = a.slots ? a.slots[idx] : idx;
  StreamState& st = a.states[slot];
  TickScratch& ts = st.tick;
  Aec3Scalars& s = st.aec.s;
    const BiquadCoef* hc = cfg.hpf_rate == 48000 ? kHpf48k : (cfg.hpf_rate == 32000 ? kHpf32k : kHpf16k);
    Biquad h0 = st.hpf[0], h1 = st.hpf[1], h2 = st.hpf[2];
    int sat = 0;
    // CaptureLevelsAdjuster::ApplyPreLevelAdjustment follows the full-band high-pass filter
    // (audio_processing_impl.cc:1280-1299).
    LevelState& lv = st.levels;
    ScalerRun pre = scaler_begin(lv.pre_prev, lv.pre_target, flen);
    if (!cfg.levels_enabled) pre.mode = 0;
    for (int i = 0; i < flen; ++i) {
      float v = front_load_sample(a.capture, idx, flen, a.fmt, i, cfg.in_channels, capture_first_channel(cfg));
      if (cfg.hpf_enabled) {
        v = biquad_step(hc[0], h0, v);
        v = biquad_step(hc[1], h1, v);
        v = biquad_step(hc[2], h2, v);
      }
      v = scaler_step(pre, v);
      ts.capture_frame[i] = v;
      sat |= (v >= 32700.0f || v <= -32700.0f) ? 1 : 0;
    }
    st.hpf[0] = h0; st.hpf[1] = h1; st.hpf[2] = h2;
    if (cfg.channels == 2 && cfg.aec_enabled) {
      // The capture AudioBuffer keeps every channel until AEC3 has looked for saturation
      // (audio_processing_impl.cc:1343,1365-1373): the second channel is high-pass filtered with
      // its own state just for that test, then dropped.
      ExtraChannelState& x = a.extra[slot];
      Biquad g0 = x.hpf[0], g1 = x.hpf[1], g2 = x.hpf[2];
      ScalerRun pre1 = scaler_begin(lv.pre_prev, lv.pre_target, flen);
      if (!cfg.levels_enabled) pre1.mode = 0;
      for (int i = 0; i < flen; ++i) {
        float v = a.rs_capture1 ? a.rs_capture1[(size_t)idx * flen + i]
                                : front_load_sample(a.capture, idx, flen, a.fmt, i, cfg.in_channels, 1);
        if (cfg.hpf_enabled) {
          v = biquad_step(hc[0], g0, v);
          v = biquad_step(hc[1], g1, v);
          v = biquad_step(hc[2], g2, v);
        }
        v = scaler_step(pre1, v);
        sat |= (v >= 32700.0f || v <= -32700.0f) ? 1 : 0;
      }
      x.hpf[0] = g0; x.hpf[1] = g1; x.hpf[2] = g2;
    }
    if (cfg.levels_enabled) lv.pre_prev = lv.pre_target;
    if (cfg.aec_enabled) {
      s.saturated_microphone_signal = sat;
      // capture_.echo_path_gain_change (audio_processing_impl.cc:1316-1341): a changed pre-gain or
      // playout volume; EchoCanceller3 hands it to every block of this frame.
      int gc = 0;
      if (cfg.levels_enabled) {
        gc |= (lv.prev_pre_adjustment_gain != lv.pre_target && lv.prev_pre_adjustment_gain >= 0.f) ? 1 : 0;
        lv.prev_pre_adjustment_gain = lv.pre_target;
      }
      gc |= (lv.prev_playout_volume != lv.playout_volume && lv.prev_playout_volume >= 0) ? 1 : 0;
      lv.prev_playout_volume = lv.playout_volume;
      ts.pad_[1] = gc;
    }
  }

// True for legs whose band split runs in k_split (warp per leg) in front of k_front: 48 kHz AEC3.
WAP_DEV bool front_presplit(const EngineConfig& cfg) { return cfg.num_bands == 3 && cfg.aec_enabled && !cfg.pre_stage; }

// The front end of one tick for leg `idx` (thread-private).
WAP_DEV void front_leg(const TickArgs& a, int idx) {
  const EngineConfig& cfg = a.cfg;
  const int B = cfg.num_bands;
  const int flen = kFrame * B;
  const int slot = a.slots ? a.slots[idx] : idx;
  StreamState& st = a.states[slot];
  TickScratch& ts = st.tick;
  Aec3State& aec = st.aec;
  Aec3Scalars& s = aec.s;
  // 32 kHz legs run the upper-band code of the 3-band path with a second upper band that stays
  // all-zero (it never raises the band-energy maximum and its output is dropped).
  UpperBandState* up = (B >= 2 && cfg.aec_enabled) ? &a.upper[slot] : nullptr;
  const bool render_live = !(cfg.reinit_on_first_capture && !st.seen_capture);
  const int delay_ms = a.capture ? (a.delays_ms ? a.delays_ms[idx] : a.uniform_delay_ms) : -1;
  // AudioProcessingImpl forwards set_stream_delay_ms() before EchoCanceller3::ProcessCapture
  // drains the render queue (audio_processing_impl.cc:1409-1415).
  if (cfg.aec_enabled && delay_ms >= 0) rdb_set_audio_buffer_delay(s, delay_ms, a.ep.fixed_capture_delay_samples);
  float frame[kFrame * kMaxBands];  // full-band frame, then its bands [B][160]
  float sub[kFrame];

  // ---------------- render: [band split] -> FrameBlocker -> BlockProcessor::BufferRender
  int nrb = 0;
  if (a.render && cfg.aec_enabled && render_live) {
    const float* band0;
    const float* rbands = frame;  // bands 1.. of the render frame
    if (front_presplit(cfg)) {
      rbands = up->render_frame;  // k_split did the three-band analysis
      band0 = rbands;
    } else if (B >= 2) {
      // AudioBuffer::SplitIntoFrequencyBands on the render side (audio_processing_impl.cc:1660-1664)
      for (int i = 0; i < flen; ++i) sub[i % kFrame] = 0.f, frame[i] = 0.f;
      float* full = ts.capture_frame;  // borrowed as the thread's 320/480-sample input buffer
      for (int i = 0; i < flen; ++i) full[i] = front_load_sample(a.render, idx, flen, a.fmt, i, cfg.render_channels, -1);
      if (B == 3) three_band_analysis_thread(full, frame, sub, st.render_bands.analysis);
      else two_band_analysis_thread(full, frame, &st.render_bands.analysis[0][0]);
      band0 = frame;
    } else {
      for (int i = 0; i < kFrame; ++i) frame[i] = front_load_sample(a.render, idx, flen, a.fmt, i, cfg.render_channels, -1);
      band0 = frame;
    }
    float rb0[kFrame];
    if (a.ep.high_pass_filter_echo_reference) {
      // RenderWriter::Insert (echo_canceller3.cc:733-737): band 0 through HighPassFilter(16000) before
      // the frame is queued for the block processor.
      Biquad h0 = aec.render_hpf[0], h1 = aec.render_hpf[1], h2 = aec.render_hpf[2];
      for (int i = 0; i < kFrame; ++i) {
        float v = biquad_step(kHpf16k[0], h0, band0[i]);
        v = biquad_step(kHpf16k[1], h1, v);
        rb0[i] = biquad_step(kHpf16k[2], h2, v);
      }
      aec.render_hpf[0] = h0; aec.render_hpf[1] = h1; aec.render_hpf[2] = h2;
      band0 = rb0;
    }
    const int L = s.render_blocker_len;
    const int total = L + kFrame;
    nrb = total / kBlock;
    float x[kBlock];
    for (int r = 0; r < nrb; ++r) {
      for (int j = 0; j < kBlock; ++j) {
        const int k = kBlock * r + j - L;
        x[j] = k < 0 ? aec.render_blocker[L + k] : band0[k];
      }
      front_render_insert(aec, ts, r, x, a.ep);
    }
    const int rem = total - kBlock * nrb;
    for (int j = 0; j < rem; ++j) aec.render_blocker[j] = band0[kFrame - rem + j];
    if (up) {
      front_slice_band(rbands + kFrame, up->render_blocker_hi[0], L, nrb, up->render_blocks_hi, 0);
      if (B == 3) front_slice_band(rbands + 2 * kFrame, up->render_blocker_hi[1], L, nrb, up->render_blocks_hi, 1);
      if (a.ep.render_linear_amplitude_gain != 1.f)
        for (int b = 0; b < nrb; ++b)
          for (int bi = 0; bi < B - 1; ++bi)
            for (int j = 0; j < kBlock; ++j) up->render_blocks_hi[b][bi][j] *= a.ep.render_linear_amplitude_gain;
    }
    s.render_blocker_len = rem;
  }
  ts.n_render_blocks = nrb;
  ts.n_capture_blocks = 0;
  if (!a.capture) return;

  // ---------------- capture: high-pass filter, saturation, [band split], FrameBlocker, decimator
  st.seen_capture = 1;
  const bool presplit = front_presplit(cfg);
  if (!presplit) front_capture_prefilter(a, idx);
  if (!cfg.aec_enabled) return;
  const float* cap0 = ts.capture_frame;  // with k_split: already the bands of the filtered frame
  if (B >= 2 && !presplit) {
    // AudioBuffer::SplitIntoFrequencyBands on the capture side (audio_processing_impl.cc:1359-1363);
    // k_echo finds the bands in ts.capture_frame instead of the full-band frame.
    for (int i = 0; i < flen; ++i) frame[i] = 0.f;
    if (B == 3) three_band_analysis_thread(ts.capture_frame, frame, sub, st.capture_bands.analysis);
    else two_band_analysis_thread(ts.capture_frame, frame, &st.capture_bands.analysis[0][0]);
    for (int i = 0; i < flen; ++i) ts.capture_frame[i] = frame[i];
    cap0 = frame;
  }
  if (a.cap_delay) {
    // BlockDelayBuffer::DelaySignal (block_delay_buffer.cc:35-67) on every band of the frame AEC3 is
    // about to process; the noise suppressor's analysis (k_echo, from ts.capture_frame) saw the frame
    // before the delay, like the reference's.
    float* ring = a.cap_delay + (size_t)slot * a.cap_delay_stride;
    const int delay = a.ep.fixed_capture_delay_samples;
    int* last_insert = reinterpret_cast<int*>(ring + (size_t)B * delay);
    const int i_start = *last_insert;
    int i = i_start;
    if (cap0 != frame) {
      for (int k = 0; k < flen; ++k) frame[k] = cap0[k];
      cap0 = frame;
    }
    for (int band = 0; band < B; ++band) {
      float* rb = ring + (size_t)band * delay;
      i = i_start;
      for (int k = 0; k < kFrame; ++k) {
        const float tmp = rb[i];
        rb[i] = frame[band * kFrame + k];
        frame[band * kFrame + k] = tmp;
        i = i < delay - 1 ? i + 1 : 0;
      }
    }
    *last_insert = i;
  }
  {
    const int L = s.capture_blocker_len;
    const int total = L + kFrame;
    const int ncb = total / kBlock;
    const bool decimate = s.render_properly_started != 0;  // blocks are only processed once render has started
    Biquad d0 = aec.capture_decimator[0], d1 = aec.capture_decimator[1], d2 = aec.capture_decimator[2],
           d3 = aec.capture_decimator[3];
    for (int b = 0; b < ncb; ++b) {
      for (int j = 0; j < kBlock; ++j) {
        const int k = kBlock * b + j - L;
        const float y = k < 0 ? aec.capture_blocker[L + k] : cap0[k];
        ts.capture_blocks[b][j] = y;
        if (decimate) {
          float v = biquad_step(kDecimator4[0], d0, y);
          v = biquad_step(kDecimator4[1], d1, v);
          v = biquad_step(kDecimator4[2], d2, v);
          v = biquad_step(kDecimator4[3], d3, v);
          if ((j & (kDownSampling - 1)) == 0) ts.cap_ds[b][j >> 2] = v;
        }
      }
    }
    if (decimate) {
      aec.capture_decimator[0] = d0; aec.capture_decimator[1] = d1; aec.capture_decimator[2] = d2;
      aec.capture_decimator[3] = d3;
    }
    const int rem = total - kBlock * ncb;
    for (int j = 0; j < rem; ++j) aec.capture_blocker[j] = cap0[kFrame - rem + j];
    if (up) {
      front_slice_band(cap0 + kFrame, up->capture_blocker_hi[0], L, ncb, up->capture_blocks_hi, 0);
      if (B == 3) front_slice_band(cap0 + 2 * kFrame, up->capture_blocker_hi[1], L, ncb, up->capture_blocks_hi, 1);
    }
    s.capture_blocker_len = rem;
    ts.n_capture_blocks = ncb;
  }
}

// PostFilter::Process (post_filter.cc:64-72) + output conversion for 48 kHz AEC3 legs: k_echo left
// the merged full-band frame in ts.capture_frame; another serial IIR, hence one thread per leg.
WAP_DEVCONST BiquadCoef kPostFilter48k[4] = {
    {0.56142156f, 1.11499931f, 0.56142156f, 1.57914249f, 0.63379496f},
    {1.00000000f, 1.88944170f, 1.00000000f, 1.55130066f, 0.68708719f},
    {1.00000000f, 1.76057310f, 1.00000000f, 1.53001328f, 0.78591224f},
    {1.00000000f, 1.67448535f, 1.00000000f, 1.56506670f, 0.92096576f}};

WAP_DEV void post_leg(const TickArgs& a, int idx) {
  const int slot = a.slots ? a.slots[idx] : idx;
  StreamState& st = a.states[slot];
  UpperBandState& up = a.upper[slot];
  const int flen = kFrame * 3;
  Biquad p0 = up.post_filter[0], p1 = up.post_filter[1], p2 = up.post_filter[2], p3 = up.post_filter[3];
  const bool zero = st.tick.pad_[0] != 0;  // first frame after un-muting (set by k_echo)
  // CaptureLevelsAdjuster::ApplyPostLevelAdjustment, after the PostFilter, while the output is used
  ScalerRun post = scaler_begin(st.levels.post_prev, st.levels.post_target, flen);
  if (!a.cfg.levels_enabled || !st.capture_output_used) post.mode = 0;
  else st.levels.post_prev = st.levels.post_target;
  for (int i = 0; i < flen; ++i) {
    float v = st.tick.capture_frame[i];
    if (st.capture_output_used) {
      v = biquad_step(kPostFilter48k[0], p0, v);
      v = biquad_step(kPostFilter48k[1], p1, v);
      v = biquad_step(kPostFilter48k[2], p2, v);
      v = biquad_step(kPostFilter48k[3], p3, v);
    }
    v = scaler_step(post, v);
    if (zero) v = 0.f;
    const int C = a.cfg.channels;  // the mono result goes to every output channel
    if (a.fmt == 0) {  // FloatS16ToS16 (audio_util.h:52-56)
      float w = fminr(v, 32767.f);
      w = fmaxr(w, -32768.f);
      const int16_t q = (int16_t)(w + copysignf(0.5f, w));
      for (int c = 0; c < C; ++c) reinterpret_cast<int16_t*>(a.out)[((size_t)idx * flen + i) * C + c] = q;
    } else {           // FloatS16ToFloat (audio_util.h:71-76)
      float w = fminr(v, 32768.f);
      w = fmaxr(w, -32768.f);
      for (int c = 0; c < C; ++c) reinterpret_cast<float*>(a.out)[((size_t)idx * C + c) * flen + i] = w * (1.f / 32768.f);
    }
  }
  up.post_filter[0] = p0; up.post_filter[1] = p1; up.post_filter[2] = p2; up.post_filter[3] = p3;
}

}  // namespace wap
