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
WAP_DEV float front_load_sample(const void* src, size_t leg, int len, int fmt, int i) {
  if (fmt == 0) return (float)(reinterpret_cast<const int16_t*>(src)[leg * len + i]);
  float v = reinterpret_cast<const float*>(src)[leg * len + i];
  v = fminr(v, 1.f);
  v = fmaxr(v, -1.f);
  return v * 32768.f;
}

// RenderDelayBufferImpl::Insert for render block `x` (64 samples, thread-local):
// scalar bookkeeping, decimation into the low-rate ring, and the record that tells
// k_echo where the block, its FFT and its spectrum go.
WAP_DEV void front_render_insert(Aec3State& a, TickScratch& ts, int r, const float* x) {
  Aec3Scalars& s = a.s;
  float x_energy = 0.f;  // DetectActiveRender: std::inner_product
  for (int i = 0; i < kBlock; ++i) x_energy += x[i] * x[i];
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
    s.render_activity_counter += (x_energy > (ec3::kActiveRenderLimit * ec3::kActiveRenderLimit) * 64.f) ? 1 : 0;
    s.render_activity = s.render_activity_counter >= 20;
  }
  ts.rins[r].blocks_write = s.blocks_write;
  ts.rins[r].spectra_write = s.spectra_write;
  ts.rins[r].previous_write = previous_write;
  // InsertBlock: decimate and store the sub-block reversed at the low-rate write index.
  Biquad d0 = a.render_decimator[0], d1 = a.render_decimator[1], d2 = a.render_decimator[2], d3 = a.render_decimator[3];
  const int lw = s.lr_write;
  for (int i = 0; i < kBlock; ++i) {
    const float xi = x[i];
    ts.render_blocks[r][i] = xi;
    float v = biquad_step(kDecimator4[0], d0, xi);
    v = biquad_step(kDecimator4[1], d1, v);
    v = biquad_step(kDecimator4[2], d2, v);
    v = biquad_step(kDecimator4[3], d3, v);
    if ((i & (kDownSampling - 1)) == 0) a.low_rate[lw + kSubBlock - 1 - (i >> 2)] = v;
  }
  a.render_decimator[0] = d0; a.render_decimator[1] = d1; a.render_decimator[2] = d2; a.render_decimator[3] = d3;
  if (s.render_event != kEventNone) rdb_reset(s);
  s.render_properly_started = 1;
}

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
  const bool render_live = !(cfg.reinit_on_first_capture && !st.seen_capture);
  const int delay_ms = a.capture ? (a.delays_ms ? a.delays_ms[idx] : a.uniform_delay_ms) : -1;
  // AudioProcessingImpl forwards set_stream_delay_ms() before EchoCanceller3::ProcessCapture
  // drains the render queue (audio_processing_impl.cc:1409-1415).
  if (cfg.aec_enabled && delay_ms >= 0) rdb_set_audio_buffer_delay(s, delay_ms);

  // ---------------- render: FrameBlocker -> BlockProcessor::BufferRender
  int nrb = 0;
  if (a.render && cfg.aec_enabled && render_live) {
    const int L = s.render_blocker_len;
    const int total = L + kFrame;
    nrb = total / kBlock;
    float x[kBlock];
    for (int r = 0; r < nrb; ++r) {
      for (int j = 0; j < kBlock; ++j) {
        const int k = kBlock * r + j - L;
        x[j] = k < 0 ? aec.render_blocker[L + k] : front_load_sample(a.render, idx, flen, a.fmt, k);
      }
      front_render_insert(aec, ts, r, x);
    }
    const int rem = total - kBlock * nrb;
    for (int j = 0; j < rem; ++j) aec.render_blocker[j] = front_load_sample(a.render, idx, flen, a.fmt, kFrame - rem + j);
    s.render_blocker_len = rem;
  }
  ts.n_render_blocks = nrb;
  ts.n_capture_blocks = 0;
  if (!a.capture) return;

  // ---------------- capture: high-pass filter, saturation, FrameBlocker, decimator
  st.seen_capture = 1;
  {
    const BiquadCoef* hc = (B == 3) ? kHpf48k : kHpf16k;
    Biquad h0 = st.hpf[0], h1 = st.hpf[1], h2 = st.hpf[2];
    int sat = 0;
    for (int i = 0; i < flen; ++i) {
      float v = front_load_sample(a.capture, idx, flen, a.fmt, i);
      if (cfg.hpf_enabled) {
        v = biquad_step(hc[0], h0, v);
        v = biquad_step(hc[1], h1, v);
        v = biquad_step(hc[2], h2, v);
      }
      ts.capture_frame[i] = v;
      sat |= (v >= 32700.0f || v <= -32700.0f) ? 1 : 0;
    }
    st.hpf[0] = h0; st.hpf[1] = h1; st.hpf[2] = h2;
    if (cfg.aec_enabled) s.saturated_microphone_signal = sat;
  }
  if (!cfg.aec_enabled) return;
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
        const float y = k < 0 ? aec.capture_blocker[L + k] : ts.capture_frame[k];
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
    for (int j = 0; j < rem; ++j) aec.capture_blocker[j] = ts.capture_frame[kFrame - rem + j];
    s.capture_blocker_len = rem;
    ts.n_capture_blocks = ncb;
  }
}

}  // namespace wap
