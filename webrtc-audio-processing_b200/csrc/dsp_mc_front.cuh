// Multi-channel call legs (BASELINE config 4), front half of a tick -- one warp per leg:
//   render : frame -> [three-band analysis per channel] -> MultiChannelContentDetector::UpdateDetection
//            (multi_channel_content_detector.cc:113-150) -> [EchoCanceller3::Initialize on a change,
//            echo_canceller3.cc:790-811] -> FillSubFrameView's downmix (:119-166) -> FrameBlocker ->
//            RenderDelayBufferImpl::Insert, scalar half (render_delay_buffer.cc:199-242,387-429: the
//            AlignmentMixer output is decimated into the low-rate ring)
//   capture: per channel high-pass filter -> AnalyzeCapture saturation over all channels
//            (echo_canceller3.cc:825-837) -> [three-band analysis] -> FrameBlocker; per block the capture
//            AlignmentMixer (alignment_mixer.cc:72-167) and the capture decimator
//            (echo_path_delay_estimator.cc:69-76) feed k_delay, which is the mono kernel.
// The serial recurrences (biquads, decimators) run on one lane per channel.
#pragma once

#include "dsp_aec3_common.cuh"
#include "dsp_aec3_render.cuh"
#include "dsp_filters.cuh"
#include "dsp_front.cuh"
#include "wap_mc_state.h"
#include "wap_tick.h"

namespace wap {

// AlignmentMixer::ProduceOutput for a block of `nch` channels x[ch][64] (band 0) -> y[64]; the adaptive
// variant's state is updated by lane 0 (SelectChannel, alignment_mixer.cc:110-167).
// downmix / adaptive / prefer_first_two: the AlignmentMixing config of this side.
WAP_DEV void mc_alignment_mix(McMixer& m, const float* x0, const float* x1, int nch, bool downmix, bool adaptive,
                              bool prefer_first_two, float activity_power_threshold, float* y, float* red) {
  const int lane = lane_id();
  __syncwarp();
  if (nch == 1) {
    for (int i = lane; i < kBlock; i += 32) y[i] = x0[i];
    __syncwarp();
    return;
  }
  if (downmix) {
    const float one_by = 1.f / nch;
    for (int i = lane; i < kBlock; i += 32) {
      float v = x0[i];
      v += x1[i];
      y[i] = v * one_by;
    }
    __syncwarp();
    return;
  }
  int ch = 0;
  if (adaptive) {
    // energies of both channels: two serial chains side by side
    if (lane < 2) {
      const float* p = lane == 0 ? x0 : x1;
      float s = 0.f;
      for (int i = 0; i < kBlock; ++i) s += p[i] * p[i];
      red[lane] = s;
    }
    __syncwarp();
    if (lane == 0) {
      constexpr int kBlocksToChooseLeftOrRight = (int)(0.5f * kNumBlocksPerSecond);
      const bool good = prefer_first_two && (m.strong_block_counters[0] > kBlocksToChooseLeftOrRight ||
                                             m.strong_block_counters[1] > kBlocksToChooseLeftOrRight);
      const int n_analyze = good ? 2 : nch;
      constexpr int kNumBlocksBeforeEnergySmoothing = 60 * kNumBlocksPerSecond;
      ++m.block_counter;
      const float thr = kBlock * activity_power_threshold;
      for (int c = 0; c < n_analyze; ++c) {
        const float x2_sum = red[c];
        if (c < 2 && x2_sum > thr) ++m.strong_block_counters[c];
        if (m.block_counter <= kNumBlocksBeforeEnergySmoothing) {
          m.cumulative_energies[c] += x2_sum;
        } else {
          constexpr float kSmoothing = 1.f / (10 * kNumBlocksPerSecond);
          m.cumulative_energies[c] += kSmoothing * (x2_sum - m.cumulative_energies[c]);
        }
      }
      if (m.block_counter == kNumBlocksBeforeEnergySmoothing) {
        constexpr float kOneBy = 1.f / kNumBlocksBeforeEnergySmoothing;
        for (int c = 0; c < n_analyze; ++c) m.cumulative_energies[c] *= kOneBy;
      }
      int strongest = 0;
      for (int c = 0; c < n_analyze; ++c)
        if (m.cumulative_energies[c] > m.cumulative_energies[strongest]) strongest = c;
      if ((good && m.selected_channel > 1) ||
          m.cumulative_energies[strongest] > 2.f * m.cumulative_energies[m.selected_channel])
        m.selected_channel = strongest;
      red[2] = (float)m.selected_channel;
    }
    __syncwarp();
    ch = (int)red[2];
  }
  const float* src = ch == 0 ? x0 : x1;
  for (int i = lane; i < kBlock; i += 32) y[i] = src[i];
  __syncwarp();
}

// EchoCanceller3::Initialize (echo_canceller3.cc:790-811): a new BlockProcessor (render delay buffer,
// delay controller, echo remover) and render FrameBlocker for the channel count / config the detector now
// asks for.  The capture FrameBlocker, the output BlockFramer, the high-pass filters and the statistics slot
// are not part of it.  `which`: 0 = the mono config's freshly constructed state, 1 = the multichannel one.
WAP_DEV void mc_initialize(StreamState& st, McState& mc, const McTemplates& t, int which) {
  const int lane = lane_id();
  __syncwarp();
  const int keep_sat = st.aec.s.saturated_microphone_signal, keep_full = st.aec.s.stats_slot_full,
            keep_delay = st.aec.s.stats_delay_blocks, keep_has = st.aec.s.stats_has_delay;
  const float keep_erl = st.aec.s.stats_erl_time_domain, keep_erle = st.aec.s.stats_erle_log2;
  __syncwarp();
  {
    // everything of Aec3State behind the (unused) mono render rings
    const int from = (int)(offsetof(Aec3State, low_rate) / 4), n = (int)(sizeof(Aec3State) / 4);
    const int* src = reinterpret_cast<const int*>(&t.aec[which]);
    int* dst = reinterpret_cast<int*>(&st.aec);
    for (int i = from + lane; i < n; i += 32) dst[i] = src[i];
    const int nc = (int)(sizeof(McChan) / 4);
    src = reinterpret_cast<const int*>(&t.chan[which]);
    for (int c = 0; c < kMcCh; ++c) {
      dst = reinterpret_cast<int*>(&mc.chan[c]);
      for (int i = lane; i < nc; i += 32) dst[i] = src[i];
    }
  }
  {
    float* p = reinterpret_cast<float*>(&mc.render);
    for (int i = lane; i < (int)(sizeof(McRender) / 4); i += 32) p[i] = 0.f;
    p = reinterpret_cast<float*>(&mc.filt[0]);
    for (int i = lane; i < (int)(kMcCh * sizeof(McFilters) / 4); i += 32) p[i] = 0.f;
    p = &mc.e_output_old_hi[0][0][0];
    for (int i = lane; i < kMcCh * 2 * kBlock; i += 32) p[i] = 0.f;
    for (int c = 0; c < kMcCh; ++c) {
      p = &mc.rio[c].render_blocker[0][0];
      for (int i = lane; i < kMaxBands * kBlock; i += 32) p[i] = 0.f;
    }
  }
  __syncwarp();
  if (lane == 0) {
    Aec3Scalars& s = st.aec.s;
    // EchoCanceller3 / AudioProcessingImpl members that outlive the block processor
    s.saturated_microphone_signal = keep_sat;
    s.stats_slot_full = keep_full;
    s.stats_erl_time_domain = keep_erl;
    s.stats_erle_log2 = keep_erle;
    s.stats_delay_blocks = keep_delay;
    s.stats_has_delay = keep_has;
    mc.render_blocker_len = 0;
    McMixer z{};
    mc.render_mixer = z;
    mc.capture_mixer = z;
  }
  __syncwarp();
}

// Scalar half of RenderDelayBufferImpl::Insert for one render block, lane 0 (the mono
// front_render_insert with the activity test on channel 0 and the decimator fed by the mixer output).
WAP_DEV void mc_render_insert_scalar(Aec3State& a, TickScratch& ts, int r, const float* x_ch0, const float* x_mixed,
                                     const Ec3Params& ep) {
  Aec3Scalars& s = a.s;
  float x_energy = 0.f;  // DetectActiveRender on band 0 of channel 0
  for (int i = 0; i < kBlock; ++i) x_energy += x_ch0[i] * x_ch0[i];
  if (s.has_delay) {
    if (!s.last_call_was_render) {
      s.last_call_was_render = 1;
      s.num_api_calls_in_a_row = 1;
    } else if (++s.num_api_calls_in_a_row > s.max_observed_jitter) {
      s.max_observed_jitter = s.num_api_calls_in_a_row;
    }
  }
  const int previous_write = s.blocks_write;
  s.lr_write = ring_off(s.lr_write, -kSubBlock, kLowRateSize);
  s.blocks_write = ring_inc(s.blocks_write, kRingBlocks);
  s.spectra_write = ring_dec(s.spectra_write, kRingBlocks);
  s.render_event = (s.lr_read == s.lr_write || s.blocks_read == s.blocks_write) ? kEventRenderOverrun : kEventNone;
  if (!s.render_activity) {
    s.render_activity_counter += (x_energy > (ep.active_render_limit * ep.active_render_limit) * 64.f) ? 1 : 0;
    s.render_activity = s.render_activity_counter >= 20;
  }
  ts.rins[r].blocks_write = s.blocks_write;
  ts.rins[r].spectra_write = s.spectra_write;
  ts.rins[r].previous_write = previous_write;
  Biquad d0 = a.render_decimator[0], d1 = a.render_decimator[1], d2 = a.render_decimator[2], d3 = a.render_decimator[3];
  const int lw = s.lr_write;
  for (int i = 0; i < kBlock; ++i) {
    float v = biquad_step(kDecimator4[0], d0, x_mixed[i]);
    v = biquad_step(kDecimator4[1], d1, v);
    v = biquad_step(kDecimator4[2], d2, v);
    v = biquad_step(kDecimator4[3], d3, v);
    if ((i & (kDownSampling - 1)) == 0) a.low_rate[lw + kSubBlock - 1 - (i >> 2)] = v;
  }
  a.render_decimator[0] = d0; a.render_decimator[1] = d1; a.render_decimator[2] = d2; a.render_decimator[3] = d3;
  if (s.render_event != kEventNone) rdb_reset(s, ep.default_delay);
  s.render_properly_started = 1;
}

// Scratch of one warp: 2 x (full frame | bands | filter-bank scratch) + blocks.
struct McFrontScratch {
  float full[kFrame * kMaxBands];
  float bands[kMcCh][kFrame * kMaxBands];
  float sub[kFrame * 2];
  float blk[kMcCh][kBlock];
  float mixed[kBlock];
  float red[8];
};

// The front half of one tick for a multi-channel leg.  mono_ep / mc_ep: the parameters of the mono and of
// the multichannel EchoCanceller3Config; tmpl[0] / tmpl[1]: freshly constructed Aec3State of each.
WAP_DEV void mc_front_tick(const TickArgs& a, int idx, McFrontScratch& fs) {
  const EngineConfig& cfg = a.cfg;
  const int lane = lane_id();
  const int B = cfg.num_bands;
  const int flen = kFrame * B;
  const int slot = a.slots ? a.slots[idx] : idx;
  StreamState& st = a.states[slot];
  McState& mc = a.mc[slot];
  TickScratch& ts = st.tick;
  McTick& mt = mc.tick;
  const int C = 2, RI = 2;  // capture channels, render input channels of this config class
  const bool render_live = !(cfg.reinit_on_first_capture && !st.seen_capture);
  const int delay_ms = a.capture ? (a.delays_ms ? a.delays_ms[idx] : a.uniform_delay_ms) : -1;
  __syncwarp();
  if (lane == 0 && delay_ms >= 0 && cfg.aec_enabled) rdb_set_audio_buffer_delay(st.aec.s, delay_ms);
  __syncwarp();

  // ---------------- render (stereo legs without AEC3 have no render side)
  int nrb = 0;
  if (a.render && render_live && cfg.aec_enabled) {
    for (int c = 0; c < RI; ++c) {
      for (int i = lane; i < flen; i += 32) fs.full[i] = front_load_sample(a.render, idx, flen, a.fmt, i, RI, c);
      __syncwarp();
      if (B == 3) three_band_analysis(fs.full, fs.bands[c], fs.sub, mc.rio[c].bands.analysis);
      else for (int i = lane; i < flen; i += 32) fs.bands[c][i] = fs.full[i];
      __syncwarp();
    }
    // MultiChannelContentDetector::UpdateDetection (detect_stereo_content = true)
    int stereo_l = 0;
    for (int i = lane; i < flen; i += 32) stereo_l |= fabsf(fs.bands[0][i] - fs.bands[1][i]) > a.mcp[0].detection_threshold;
    const bool stereo_in_frame = a.mcp[0].detect_stereo_content && __any_sync(WAP_FULL, stereo_l);
    McDetector& d = mc.det;
    const int prev_persistent = d.persistent;
    __syncwarp();
    int persistent = prev_persistent;
    {
      const int consecutive = stereo_in_frame ? d.consecutive_frames_with_stereo + 1 : 0;
      const int since = stereo_in_frame ? 0 : d.frames_since_stereo_detected_last + 1;
      if (consecutive > a.mcp[0].hysteresis_frames) persistent = 1;
      if (a.mcp[0].timeout_frames > 0 && since >= a.mcp[0].timeout_frames) persistent = 0;
      if (!a.mcp[0].detect_stereo_content) persistent = prev_persistent;   // UpdateDetection returns early
      __syncwarp();
      if (lane == 0 && a.mcp[0].detect_stereo_content) {
        d.consecutive_frames_with_stereo = consecutive;
        d.frames_since_stereo_detected_last = since;
        d.persistent = persistent;
        d.temporary = persistent ? 0 : (stereo_in_frame ? 1 : 0);
      }
    }
    const bool temporary = !persistent && stereo_in_frame;
    __syncwarp();
    if (persistent != prev_persistent) {
      mc_initialize(st, mc, *a.mc_templates, persistent ? 1 : 0);
      if (lane == 0) d.render_channels_to_aec = persistent ? RI : 1;
      __syncwarp();
    }
    const int R = persistent ? RI : 1;
    // FillSubFrameView: downmix for a mono AEC -- the average of the channels while (temporary) stereo is
    // seen, channel 0 otherwise.  (The reference averages sub-frame by sub-frame, in place: same values.)
    if (R == 1 && temporary) {
      const float one_by = 1.0f / RI;
      for (int i = lane; i < flen; i += 32) {
        float v = fs.bands[0][i];
        v += fs.bands[1][i];
        fs.bands[0][i] = v * one_by;
      }
      __syncwarp();
    }
    // FrameBlocker per channel and band, then BufferRender block by block
    const int L = mc.render_blocker_len;
    const int total = L + kFrame;
    nrb = total / kBlock;
    for (int r = 0; r < nrb; ++r) {
      for (int c = 0; c < R; ++c)
        for (int b = 0; b < B; ++b)
          for (int j = lane; j < kBlock; j += 32) {
            const int k = kBlock * r + j - L;
            const float v = k < 0 ? mc.rio[c].render_blocker[b][L + k] : fs.bands[c][b * kFrame + k];
            mt.render_blocks[r][c][b][j] = v;
            if (b == 0) fs.blk[c][j] = v;
          }
      __syncwarp();
      const Ec3Params& ep = persistent ? a.ep_mc : a.ep;
      const McParams& mp = a.mcp[persistent ? 1 : 0];
      mc_alignment_mix(mc.render_mixer, fs.blk[0], fs.blk[1], R, mp.render_mix_downmix != 0, mp.render_mix_adaptive != 0,
                       mp.render_mix_prefer_first_two != 0, mp.render_mix_threshold, fs.mixed, fs.red);
      if (lane == 0) {
        mc_render_insert_scalar(st.aec, ts, r, fs.blk[0], fs.mixed, ep);
        mt.render_channels[r] = R;
      }
      __syncwarp();
    }
    const int rem = total - kBlock * nrb;
    __syncwarp();
    for (int c = 0; c < R; ++c)
      for (int b = 0; b < B; ++b) {
        // the remainder may overlap what was just read: stage through registers
        float keepv = 0.f;
        const int j = lane;
        if (j < rem) keepv = fs.bands[c][b * kFrame + kFrame - rem + j];
        float keepv2 = 0.f;
        if (j + 32 < rem) keepv2 = fs.bands[c][b * kFrame + kFrame - rem + j + 32];
        __syncwarp();
        if (j < rem) mc.rio[c].render_blocker[b][j] = keepv;
        if (j + 32 < rem) mc.rio[c].render_blocker[b][j + 32] = keepv2;
      }
    __syncwarp();
    if (lane == 0) mc.render_blocker_len = rem;
  }
  __syncwarp();
  if (lane == 0) {
    ts.n_render_blocks = nrb;
    ts.n_capture_blocks = 0;
  }
  __syncwarp();
  if (!a.capture) return;

  // ---------------- capture
  if (lane == 0) st.seen_capture = 1;
  const BiquadCoef* hc = cfg.hpf_rate == 48000 ? kHpf48k : (cfg.hpf_rate == 32000 ? kHpf32k : kHpf16k);
  int sat = 0;
  for (int c = 0; c < C; ++c)
    for (int i = lane; i < flen; i += 32) fs.bands[c][i] = front_load_sample(a.capture, idx, flen, a.fmt, i, C, c);
  __syncwarp();
  // HighPassFilter, then CaptureLevelsAdjuster::ApplyPreLevelAdjustment (audio_processing_impl.cc:1280-1299):
  // serial recurrences, one lane per channel; every channel sees the same gain ramp
  LevelState& lv = st.levels;
  const float pre_prev = lv.pre_prev, pre_target = lv.pre_target;
  __syncwarp();
  if (lane < C && (cfg.hpf_enabled || cfg.levels_enabled)) {
    const int c = lane;
    float* x = fs.bands[c];
    Biquad h0 = mc.cio[c].hpf[0], h1 = mc.cio[c].hpf[1], h2 = mc.cio[c].hpf[2];
    ScalerRun pre = scaler_begin(pre_prev, pre_target, flen);
    if (!cfg.levels_enabled) pre.mode = 0;
    for (int i = 0; i < flen; ++i) {
      float v = x[i];
      if (cfg.hpf_enabled) {
        v = biquad_step(hc[0], h0, v);
        v = biquad_step(hc[1], h1, v);
        v = biquad_step(hc[2], h2, v);
      }
      x[i] = scaler_step(pre, v);
    }
    mc.cio[c].hpf[0] = h0; mc.cio[c].hpf[1] = h1; mc.cio[c].hpf[2] = h2;
  }
  __syncwarp();
  if (lane == 0 && cfg.levels_enabled) lv.pre_prev = pre_target;
  for (int c = 0; c < C; ++c) {
    for (int i = lane; i < flen; i += 32) {
      const float v = fs.bands[c][i];
      sat |= (v >= 32700.0f || v <= -32700.0f) ? 1 : 0;
      fs.full[i] = v;
    }
    __syncwarp();
    if (B == 3) three_band_analysis(fs.full, fs.bands[c], fs.sub, mc.cio[c].bands.analysis);
    __syncwarp();
    // NoiseSuppressor::Analyze (k_mc_echo) looks at band 0 of the capture frame in front of the echo canceller;
    // without AEC3 the bands go to k_mc_echo as they are
    if (!cfg.aec_enabled)
      for (int i = lane; i < flen; i += 32) mt.capture_frame[c][i] = fs.bands[c][i];
    else if (cfg.ns_enabled)
      for (int i = lane; i < kFrame; i += 32) mt.capture_frame[c][i] = fs.bands[c][i];
  }
  if (!cfg.aec_enabled) {
    if (lane == 0 && cfg.levels_enabled) st.levels.prev_pre_adjustment_gain = st.levels.pre_target;
    __syncwarp();
    return;
  }
  const bool saturated = __any_sync(WAP_FULL, sat);
  __syncwarp();
  if (lane == 0) {
    st.aec.s.saturated_microphone_signal = saturated ? 1 : 0;
    // echo_path_gain_change = level_change || aec_reference_is_downmixed_stereo (echo_canceller3.cc:168-170);
    // level_change: a changed pre-gain or playout volume (audio_processing_impl.cc:1316-1341)
    int gc = 0;
    if (cfg.levels_enabled) {
      gc |= (lv.prev_pre_adjustment_gain != lv.pre_target && lv.prev_pre_adjustment_gain >= 0.f) ? 1 : 0;
      lv.prev_pre_adjustment_gain = lv.pre_target;
    }
    gc |= (lv.prev_playout_volume != lv.playout_volume && lv.prev_playout_volume >= 0) ? 1 : 0;
    lv.prev_playout_volume = lv.playout_volume;
    mt.gain_change = (gc || mc.det.temporary) ? 1 : 0;
    ts.pad_[1] = mt.gain_change;
  }
  {
    const int L = mc.capture_blocker_len;
    const int total = L + kFrame;
    const int ncb = total / kBlock;
    const bool decimate = st.aec.s.render_properly_started != 0;
    for (int b = 0; b < ncb; ++b) {
      for (int c = 0; c < C; ++c)
        for (int band = 0; band < B; ++band)
          for (int j = lane; j < kBlock; j += 32) {
            const int k = kBlock * b + j - L;
            const float v = k < 0 ? mc.cio[c].capture_blocker[band][L + k] : fs.bands[c][band * kFrame + k];
            mt.capture_blocks[b][c][band][j] = v;
            if (band == 0) fs.blk[c][j] = v;
          }
      __syncwarp();
      if (decimate) {
        const McParams& mp = a.mcp[mc.det.persistent ? 1 : 0];
        mc_alignment_mix(mc.capture_mixer, fs.blk[0], fs.blk[1], C, mp.capture_mix_downmix != 0, mp.capture_mix_adaptive != 0,
                         mp.capture_mix_prefer_first_two != 0, mp.capture_mix_threshold, fs.mixed, fs.red);
        if (lane == 0) {
          Aec3State& aec = st.aec;
          Biquad d0 = aec.capture_decimator[0], d1 = aec.capture_decimator[1], d2 = aec.capture_decimator[2],
                 d3 = aec.capture_decimator[3];
          for (int j = 0; j < kBlock; ++j) {
            float v = biquad_step(kDecimator4[0], d0, fs.mixed[j]);
            v = biquad_step(kDecimator4[1], d1, v);
            v = biquad_step(kDecimator4[2], d2, v);
            v = biquad_step(kDecimator4[3], d3, v);
            if ((j & (kDownSampling - 1)) == 0) ts.cap_ds[b][j >> 2] = v;
          }
          aec.capture_decimator[0] = d0; aec.capture_decimator[1] = d1; aec.capture_decimator[2] = d2;
          aec.capture_decimator[3] = d3;
        }
        __syncwarp();
      }
    }
    const int rem = total - kBlock * ncb;
    __syncwarp();
    for (int c = 0; c < C; ++c)
      for (int band = 0; band < B; ++band) {
        float k0 = 0.f, k1 = 0.f;
        if (lane < rem) k0 = fs.bands[c][band * kFrame + kFrame - rem + lane];
        if (lane + 32 < rem) k1 = fs.bands[c][band * kFrame + kFrame - rem + lane + 32];
        __syncwarp();
        if (lane < rem) mc.cio[c].capture_blocker[band][lane] = k0;
        if (lane + 32 < rem) mc.cio[c].capture_blocker[band][lane + 32] = k1;
      }
    __syncwarp();
    if (lane == 0) {
      mc.capture_blocker_len = rem;
      ts.n_capture_blocks = ncb;
    }
  }
  __syncwarp();
}

}  // namespace wap
