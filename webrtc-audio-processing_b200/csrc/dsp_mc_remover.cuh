// Echo remover of a multi-channel leg (R render channels, C capture channels): EchoRemoverImpl::ProcessCapture
// with the reference's per-channel loops and cross-channel aggregations.
//   EchoRemoverImpl::ProcessCapture                       aec3/echo_remover.cc:254-460
//   AecState::{HandleEchoPathChange, Update}              aec3/aec_state.cc:152-342  (ComputeAvgRenderReverb :56-108)
//   SubtractorOutputAnalyzer::Update                      aec3/subtractor_output_analyzer.cc:26-66
//   FilterAnalyzer::Update / AnalyzeRegion                aec3/filter_analyzer.cc:80-141
//   SubbandErleEstimator / FullBandErleEstimator          aec3/subband_erle_estimator.cc:72-259, fullband_erle_estimator.cc:62-100
//   ErlEstimator::Update (max over channels)              aec3/erl_estimator.cc:46-149
//   ReverbModelEstimator::Update                          aec3/reverb_model_estimator.cc:39-66
//   ResidualEchoEstimator (render power summed over ch.)  aec3/residual_echo_estimator.cc:133-165,193-425
//   ComfortNoiseGenerator::Compute (one seed, C channels) aec3/comfort_noise_generator.cc:125-192
//   SuppressionGain (min over the capture channels)       aec3/suppression_gain.cc:124-430
//   DominantNearendDetector (per-channel counters)        aec3/dominant_nearend_detector.cc:37-81
//   SuppressionFilter::ApplyGain                          aec3/suppression_filter.cc:88-183
// State: `sh` = StreamState::aec (everything that exists once), `mc.chan[c]` / `mx.cs[c]` per capture channel.
#pragma once

#include "dsp_aec3_remover.cuh"
#include "dsp_mc_subtractor.cuh"

namespace wap {

// SubbandErleEstimator::Reset + FullBandErleEstimator::Reset for one capture channel.
WAP_DEV void mc_erle_reset_channel(McChan& ch, Aec3Scalars& cs, AecScratch& sc) {
  const int lane = lane_id();
  for (int k = lane; k < kBins; k += 32) {
    ch.erle[k] = WAP_EC3(erle_min);
    ch.erle_onset_comp[k] = WAP_EC3(erle_min);
    ch.erle_unbounded[k] = WAP_EC3(erle_min);
    ch.coming_onset[k] = 1;
    ch.erle_hold_counters[k] = 0;
    ch.accum_Y2[k] = 0.f;
    ch.accum_E2[k] = 0.f;
    ch.accum_low_render[k] = 0;
  }
  if (lane == 0) {
    Aec3Scalars& s = cs;
    s.erle_num_points = 0;
    s.fb_has_erle_log2 = 0;
    s.fb_inst_quality = 0.f;
    s.fb_num_points = 0;
    s.fb_E2_acum = 0.f;
    s.fb_Y2_acum = 0.f;
    s.fb_max_erle_log2 = -10.f;
    s.fb_min_erle_log2 = 33.f;
    s.fb_erle_time_domain_log2 = fast_approx_log2f(WAP_EC3(erle_min) + 1e-3f);
    s.fb_hold_counter = 0;
  }
  __syncwarp();
}
WAP_DEV void mc_erle_reset(McState& mc, McExtra& mx, AecScratch& sc, int C, bool delay_change) {
  for (int c = 0; c < C; ++c) mc_erle_reset_channel(mc.chan[c], mx.cs[c], sc);
  if (delay_change && lane_id() == 0) sc.s.erle_blocks_since_reset = 0;
  __syncwarp();
}

// AecState::HandleEchoPathChange
WAP_DEV void mc_aec_state_handle_echo_path_change(McState& mc, McExtra& mx, AecScratch& sc, const EchoPathVariability& v, int C) {
  Aec3Scalars& s = sc.s;
  __syncwarp();
  if (v.delay_change != kDelayAdjNone) {
    if (lane_id() == 0) {
      // FilterAnalyzer::Reset
      s.fa_blocks_since_reset = 0;
      s.fa_region_start = 0;
      s.fa_region_end = 0;
      for (int c = 0; c < C; ++c) {
        Aec3Scalars& cs = mx.cs[c];
        cs.fa_peak_index = 0;
        cs.fa_gain = WAP_EC3(default_gain);
        cs.cfd_significant_peak = 0;
        cs.cfd_floor_accum = 0.f;
        cs.cfd_secondary_peak = 0.f;
        cs.cfd_floor_low_limit = 0;
        cs.cfd_floor_high_limit = 0;
        cs.cfd_consistent_counter = 0;
        cs.cfd_consistent_delay_reference = -10;
        cs.fa_filter_delay_blocks = 0;
      }
      s.capture_signal_saturation = 0;
      s.strong_not_saturated_render_blocks = 0;
      s.blocks_with_active_render = 0;
      s.init_state = 1;
      s.init_strong_blocks = 0;
      s.tm_non_converged_sequence_size = 10000;
      s.tm_diverged_sequence_size = 0;
      s.tm_strong_not_saturated_render_blocks = 0;
      s.erl_blocks_since_reset = 0;
      s.fq_usable = 0;
      s.fq_blocks_since_reset = 0;
    }
    mc_erle_reset(mc, mx, sc, C, true);
  } else if (v.gain_change) {
    mc_erle_reset(mc, mx, sc, C, false);
  }
  if (lane_id() == 0)
    for (int c = 0; c < C; ++c) mx.cs[c].soa_filter_converged = 0;
  __syncwarp();
}

// FilterAnalyzer::AnalyzeRegion for capture channel c (region already set in the shared scalars).
WAP_DEV void mc_filter_analyzer_channel(const McRender& rb, McChan& ch, Aec3Scalars& cs, AecScratch& sc, int R) {
  const int lane = lane_id();
  Aec3Scalars& s = sc.s;
  AecRemoverScratch& r = sc.rm;
  __syncwarp();
  const int size = cs.h_time_size * kBlock;
  const int start = s.fa_region_start, end = s.fa_region_end;
  for (int i = cs.fa_hp_size + lane; i < size; i += 32) ch.h_highpass[i] = 0.f;
  for (int k = start + lane; k <= end; k += 32) {
    float tmp = 0.f;
    if (k >= 2) {
      tmp += ch.h_time[k] * 0.7929742f;
      tmp += ch.h_time[k - 1] * -0.36072128f;
      tmp += ch.h_time[k - 2] * -0.47047766f;
    }
    ch.h_highpass[k] = tmp;
    r.v0[k - start] = tmp;
  }
  __syncwarp();
  const int peak_in = imin(cs.fa_peak_index, size - 1);
  float best = -1.f;
  int bi = 0x7fffffff;
  for (int k = start + lane; k <= end; k += 32) {
    const float v = r.v0[k - start] * r.v0[k - start];
    if (v > best) { best = v; bi = k; }
  }
  for (int m = 16; m; m >>= 1) {
    const float ov = __shfl_xor_sync(WAP_FULL, best, m);
    const int oi = __shfl_xor_sync(WAP_FULL, bi, m);
    if (ov > best || (ov == best && oi < bi)) { best = ov; bi = oi; }
  }
  const float hp_in = ch.h_highpass[peak_in];
  const int peak = (best > hp_in * hp_in) ? bi : peak_in;
  const int delay_blocks = peak >> 6;
  const float h_peak = ch.h_highpass[peak];
  __syncwarp();
  if (lane == 0) {
    cs.fa_hp_size = size;
    cs.fa_peak_index = peak;
    cs.fa_filter_delay_blocks = delay_blocks;
    const bool sufficient_time_to_converge = s.fa_blocks_since_reset > 5 * kNumBlocksPerSecond;
    if (sufficient_time_to_converge && cs.fa_consistent_estimate) {
      cs.fa_gain = fabsf(h_peak);
    } else if (cs.fa_gain) {
      cs.fa_gain = fmaxr(cs.fa_gain, fabsf(h_peak));
    }
    cs.fa_filter_length_blocks = (int)((float)size * (1.f / kBlock));
    if (start == 0) {
      cs.cfd_floor_accum = 0.f;
      cs.cfd_secondary_peak = 0.f;
      cs.cfd_floor_low_limit = peak < 64 ? 0 : peak - 64;
      cs.cfd_floor_high_limit = peak > size - 129 ? 0 : peak + 128;
    }
    float accum = cs.cfd_floor_accum, secondary = cs.cfd_secondary_peak;
    for (int k = start; k < imin(end + 1, cs.cfd_floor_low_limit); ++k) {
      const float abs_h = fabsf(r.v0[k - start]);
      accum += abs_h;
      secondary = fmaxr(secondary, abs_h);
    }
    for (int k = imax(cs.cfd_floor_high_limit, start); k <= end; ++k) {
      const float abs_h = fabsf(r.v0[k - start]);
      accum += abs_h;
      secondary = fmaxr(secondary, abs_h);
    }
    cs.cfd_floor_accum = accum;
    cs.cfd_secondary_peak = secondary;
    if (end == size - 1) {
      const float filter_floor = accum / (float)(cs.cfd_floor_low_limit + size - cs.cfd_floor_high_limit);
      const float abs_peak = fabsf(h_peak);
      cs.cfd_significant_peak = abs_peak > 10.f * filter_floor && abs_peak > 2.f * secondary;
    }
  }
  __syncwarp();
  if (cs.cfd_significant_peak) {
    // active_render_block: any render channel of GetBlock(-delay_blocks) above the activity threshold
    const int row = ring_off(s.blocks_read, -delay_blocks, kRingBlocks);
    if (lane < R) {
      const float* xb = rb.blocks[row][lane][0];
      float acc = 0.f;
      for (int i = 0; i < kBlock; ++i) acc += xb[i] * xb[i];
      sc.red[20 + lane] = acc;
    }
    __syncwarp();
    if (lane == 0) {
      bool active_render_block = false;
      for (int rc = 0; rc < R; ++rc)
        if (sc.red[20 + rc] > kActiveRenderEnergy) { active_render_block = true; break; }
      if (cs.cfd_consistent_delay_reference == delay_blocks) {
        if (active_render_block) ++cs.cfd_consistent_counter;
      } else {
        cs.cfd_consistent_counter = 0;
        cs.cfd_consistent_delay_reference = delay_blocks;
      }
    }
  }
  if (lane == 0) cs.fa_consistent_estimate = (float)cs.cfd_consistent_counter > 1.5f * kNumBlocksPerSecond;
  __syncwarp();
}

// AecState::Update.  In: mx.cv[c].{Y2, E2, metrics}; out: the shared / per-channel state, r.x_aligned unused.
WAP_DEV void mc_aec_state_update(Aec3State& sh, McState& mc, AecScratch& sc, McExtra& mx, int R, int C, int ext_has,
                                 int ext_delay) {
  const int lane = lane_id();
  Aec3Scalars& s = sc.s;
  AecRemoverScratch& r = sc.rm;
  const McRender& rb = mc.render;
  __syncwarp();
  // SubtractorOutputAnalyzer::Update
  bool any_filter_converged = false, all_filters_diverged = true;
  int converged_mask = 0;
  for (int c = 0; c < C; ++c) {
    const float y2 = mx.cv[c].metrics[0], e2_refined = mx.cv[c].metrics[1], e2_coarse = mx.cv[c].metrics[2];
    constexpr float kConvergenceThreshold = 50 * 50 * kBlock;
    const bool refined_filter_converged = e2_refined < 0.5f * y2 && y2 > kConvergenceThreshold;
    const bool coarse_filter_converged_strict = e2_coarse < 0.05f * y2 && y2 > kConvergenceThreshold;
    const float min_e2 = fminr(e2_refined, e2_coarse);
    const bool filter_diverged = min_e2 > 1.5f * y2 && y2 > 30.f * 30.f * kBlock;
    const bool conv = refined_filter_converged || coarse_filter_converged_strict;
    converged_mask |= conv ? (1 << c) : 0;
    any_filter_converged = any_filter_converged || conv;
    all_filters_diverged = all_filters_diverged && filter_diverged;
  }
  __syncwarp();
  if (lane == 0)
    for (int c = 0; c < C; ++c) mx.cs[c].soa_filter_converged = (converged_mask >> c) & 1;

  // FilterAnalyzer::Update
  if (lane == 0) {
    ++s.fa_blocks_since_reset;
    const int size = mx.cs[0].h_time_size * kBlock;
    s.fa_region_start = s.fa_region_end >= size - 1 ? 0 : s.fa_region_end + 1;
    s.fa_region_end = imin(s.fa_region_start + kBlock - 1, size - 1);
  }
  __syncwarp();
  for (int c = 0; c < C; ++c) mc_filter_analyzer_channel(rb, mc.chan[c], mx.cs[c], sc, R);
  bool any_filter_consistent = mx.cs[0].fa_consistent_estimate != 0;
  float max_echo_path_gain = mx.cs[0].fa_gain;
  for (int c = 1; c < C; ++c) {
    any_filter_consistent = any_filter_consistent || mx.cs[c].fa_consistent_estimate != 0;
    max_echo_path_gain = fmaxr(max_echo_path_gain, mx.cs[c].fa_gain);
  }
  __syncwarp();
  // FilterDelay::Update
  if (lane == 0) {
    if (ext_has && (!s.fd_has_external || s.fd_external_delay != ext_delay)) {
      s.fd_has_external = 1;
      s.fd_external_delay = ext_delay;
    }
    const bool may_not_have_converged = s.strong_not_saturated_render_blocks < 2 * kNumBlocksPerSecond;
    int mn = 0x7fffffff;
    for (int c = 0; c < C; ++c) {
      Aec3Scalars& cs = mx.cs[c];
      if (may_not_have_converged && s.fd_has_external) cs.fd_filter_delay = WAP_EC3(delay_headroom_samples) / kBlock;
      else cs.fd_filter_delay = cs.fa_filter_delay_blocks;
      mn = imin(mn, cs.fd_filter_delay);
    }
    s.fd_min_filter_delay = mn;
  }
  __syncwarp();
  const int delay = s.fd_min_filter_delay;

  // aligned render block: activity (any channel) and peak sample (all channels)
  const int row_aligned = ring_off(s.blocks_read, -delay, kRingBlocks);
  if (lane < R) {
    const float* xb = rb.blocks[row_aligned][lane][0];
    float acc = 0.f;
    for (int i = 0; i < kBlock; ++i) acc += xb[i] * xb[i];
    sc.red[20 + lane] = acc;
  }
  float max_sample_l = 0.f;
  for (int rc = 0; rc < R; ++rc)
    for (int i = lane; i < kBlock; i += 32) max_sample_l = fmaxf(max_sample_l, fabsf(rb.blocks[row_aligned][rc][0][i]));
  const float max_sample = warp_max(max_sample_l);
  __syncwarp();
  bool active_render = false;
  for (int rc = 0; rc < R; ++rc)
    if (sc.red[20 + rc] > kActiveRenderEnergy) { active_render = true; break; }
  const bool saturated_capture = s.capture_signal_saturation != 0;
  const bool usable_linear_before = s.fq_usable != 0;
  __syncwarp();
  if (lane == 0) {
    s.blocks_with_active_render += active_render ? 1 : 0;
    s.strong_not_saturated_render_blocks += (active_render && !saturated_capture) ? 1 : 0;
  }

  // ComputeAvgRenderReverb: r.v1 = avg_render_spectrum_with_reverb; r.v2 = max render spectrum at the delay
  // (ErlEstimator).  R > 1: the channel average.
  const int idx_at_delay = ring_off(s.spectra_read, delay, kRingBlocks);
  const int idx_past = ring_inc(idx_at_delay, kRingBlocks);
  for (int k = lane; k < kBins; k += 32) {
    float past, at, mx2;
    if (R > 1) {
      float p = 0.f, q = 0.f;
      mx2 = rb.spectra[idx_at_delay][0][k];
      for (int rc = 0; rc < R; ++rc) {
        p += rb.spectra[idx_past][rc][k];
        q += rb.spectra[idx_at_delay][rc][k];
        if (rc) mx2 = fmaxr(mx2, rb.spectra[idx_at_delay][rc][k]);
      }
      const float normalizer = 1.f / R;
      past = p * normalizer;
      at = q * normalizer;
    } else {
      past = rb.spectra[idx_past][0][k];
      at = rb.spectra[idx_at_delay][0][k];
      mx2 = at;
    }
    const float rev = (sh.avg_render_reverb[k] + past * 1.0f) * WAP_EC3(default_len);
    sh.avg_render_reverb[k] = rev;
    r.v1[k] = at + rev;
    r.v2[k] = mx2;
  }
  __syncwarp();
  if (s.init_transition_triggered) mc_erle_reset(mc, mx, sc, C, false);

  // ErlEstimator's capture spectrum: the maximum over the channels with a converged filter -> r.v3
  {
    int first = 0;
    while (first < C && !((converged_mask >> first) & 1)) ++first;
    if (first < C) {
      for (int k = lane; k < kBins; k += 32) {
        float m = mx.cv[first].Y2[k];
        for (int c = first + 1; c < C; ++c)
          if ((converged_mask >> c) & 1) m = fmaxr(m, mx.cv[c].Y2[k]);
        r.v3[k] = m;
      }
    }
  }
  __syncwarp();
  // 65-term chains side by side: X2_reverb, X2 (max), Y2 (max), and Y2 / E2 of every channel
  if (lane < 3 + 2 * C) {
    const float* p;
    if (lane == 0) p = r.v1;
    else if (lane == 1) p = r.v2;
    else if (lane == 2) p = r.v3;
    else p = ((lane - 3) & 1) ? mx.cv[(lane - 3) >> 1].E2 : mx.cv[(lane - 3) >> 1].Y2;
    sc.red[16 + lane] = chain_sum(p, 0, kBins);
  }
  __syncwarp();
  const float X2rev_sum = sc.red[16], X2max_sum = sc.red[17], Y2max_sum = sc.red[18];

  // ---- ErleEstimator::Update
  if (lane == 0) sc.ired[8] = (++s.erle_blocks_since_reset < 2 * kNumBlocksPerSecond) ? 0 : 1;
  __syncwarp();
  if (sc.ired[8]) {
    for (int c = 0; c < C; ++c) {
      McChan& ch = mc.chan[c];
      Aec3Scalars& cs = mx.cs[c];
      McChanVec& cv = mx.cv[c];
      const bool converged = ((converged_mask >> c) & 1) != 0;
      const float Y2_sum = sc.red[19 + 2 * c], E2_sum = sc.red[20 + 2 * c];
      __syncwarp();
      const bool restart = converged && cs.erle_num_points == 6;
      const int num_points = converged ? (restart ? 1 : cs.erle_num_points + 1) : cs.erle_num_points;
      const bool update_bands = converged && num_points == 6;
      for (int k = lane; k < kBins; k += 32) {
        float accY = ch.accum_Y2[k], accE = ch.accum_E2[k];
        int low = ch.accum_low_render[k];
        if (converged) {
          if (restart) { accY = 0.f; accE = 0.f; low = 0; }
          accY = cv.Y2[k] + accY;
          accE = cv.E2[k] + accE;
          low = low || r.v1[k] < kX2BandEnergyThreshold;
          ch.accum_Y2[k] = accY;
          ch.accum_E2[k] = accE;
          ch.accum_low_render[k] = low;
        }
        if (k >= 1 && k < 64) {
          float erle = ch.erle[k], erle_oc = ch.erle_onset_comp[k], erle_u = ch.erle_unbounded[k];
          int hold = ch.erle_hold_counters[k], onset = ch.coming_onset[k];
          if (update_bands && accE > 0.f) {
            const float new_erle = accY / accE;
            if (!low) {
              if (onset) onset = 0;
              hold = 250;
            }
            const float max_erle = k < 32 ? WAP_EC3(erle_max_l) : WAP_EC3(erle_max_h);
            float alpha = 0.05f;
            if (new_erle < erle) alpha = low ? 0.f : 0.1f;
            erle = clampr(erle + alpha * (new_erle - erle), WAP_EC3(erle_min), max_erle);
            alpha = 0.05f;
            if (new_erle < erle_oc) alpha = low ? 0.f : 0.1f;
            erle_oc = clampr(erle_oc + alpha * (new_erle - erle_oc), WAP_EC3(erle_min), max_erle);
            alpha = 0.05f;
            if (new_erle < erle_u) alpha = low ? 0.f : 0.1f;
            erle_u = clampr(erle_u + alpha * (new_erle - erle_u), WAP_EC3(erle_min), 100000.0f);
          }
          --hold;
          if (hold <= 250 - 100) {
            if (erle_oc > WAP_EC3(erle_min)) erle_oc = fmaxr(WAP_EC3(erle_min), 0.97f * erle_oc);
            if (hold <= 0) { onset = 1; hold = 0; }
          }
          ch.erle[k] = erle;
          ch.erle_onset_comp[k] = erle_oc;
          ch.erle_unbounded[k] = erle_u;
          ch.erle_hold_counters[k] = hold;
          ch.coming_onset[k] = onset;
          if (k == 1) { ch.erle[0] = erle; ch.erle_onset_comp[0] = erle_oc; ch.erle_unbounded[0] = erle_u; }
          if (k == 63) { ch.erle[64] = erle; ch.erle_onset_comp[64] = erle_oc; ch.erle_unbounded[64] = erle_u; }
        }
      }
      __syncwarp();
      if (lane == 0) {
        Aec3Scalars& q = cs;
        q.erle_num_points = num_points;
        if (converged && X2rev_sum > kX2BandEnergyThreshold * (float)kBins) {
          bool update_estimates = false;
          q.fb_E2_acum += E2_sum;
          q.fb_Y2_acum += Y2_sum;
          if (++q.fb_num_points == 6) {
            if (q.fb_E2_acum > 0.f) {
              update_estimates = true;
              q.fb_erle_log2 = fast_approx_log2f(q.fb_Y2_acum / q.fb_E2_acum + 1e-3f);
              q.fb_has_erle_log2 = 1;
            }
            q.fb_num_points = 0;
            q.fb_E2_acum = 0.f;
            q.fb_Y2_acum = 0.f;
          }
          if (update_estimates) {
            q.fb_max_erle_log2 -= 0.0004f;
            q.fb_max_erle_log2 = fmaxr(q.fb_max_erle_log2, q.fb_erle_log2);
            q.fb_min_erle_log2 += 0.0004f;
            q.fb_min_erle_log2 = fminr(q.fb_min_erle_log2, q.fb_erle_log2);
            float quality_estimate = 0.f;
            if (q.fb_max_erle_log2 > q.fb_min_erle_log2)
              quality_estimate = (q.fb_erle_log2 - q.fb_min_erle_log2) / (q.fb_max_erle_log2 - q.fb_min_erle_log2);
            if (quality_estimate > q.fb_inst_quality) q.fb_inst_quality = quality_estimate;
            else q.fb_inst_quality += 0.07f * (quality_estimate - q.fb_inst_quality);
            q.fb_hold_counter = 100;
            q.fb_erle_time_domain_log2 += 0.05f * (q.fb_erle_log2 - q.fb_erle_time_domain_log2);
            q.fb_erle_time_domain_log2 = fmaxr(q.fb_erle_time_domain_log2, fast_approx_log2f(WAP_EC3(erle_min) + 1e-3f));
          }
        }
        --q.fb_hold_counter;
        if (q.fb_hold_counter == 0) {
          q.fb_has_erle_log2 = 0;
          q.fb_inst_quality = 0.f;
          q.fb_num_points = 0;
          q.fb_E2_acum = 0.f;
          q.fb_Y2_acum = 0.f;
        }
      }
      __syncwarp();
    }
  }

  // ---- ErlEstimator::Update
  if (lane == 0) sc.ired[9] = (++s.erl_blocks_since_reset < 2 * kNumBlocksPerSecond || !any_filter_converged) ? 0 : 1;
  __syncwarp();
  if (sc.ired[9]) {
    for (int k = 1 + lane; k < 64; k += 32) {
      float erl = sh.erl[k];
      int hold = sh.erl_hold_counters[k - 1];
      const float X2 = r.v2[k];
      if (X2 > kX2BandEnergyThreshold) {
        const float new_erl = r.v3[k] / X2;
        if (new_erl < erl) {
          hold = 1000;
          erl += 0.1f * (new_erl - erl);
          erl = fmaxr(erl, 0.01f);
        }
      }
      --hold;
      erl = hold > 0 ? erl : fminr(1000.f, 2.f * erl);
      sh.erl[k] = erl;
      sh.erl_hold_counters[k - 1] = hold;
      if (k == 1) sh.erl[0] = erl;
      if (k == 63) sh.erl[64] = erl;
    }
    if (lane == 0) {
      if (X2max_sum > kX2BandEnergyThreshold * (float)kBins) {
        const float new_erl = Y2max_sum / X2max_sum;
        if (new_erl < s.erl_time_domain) {
          s.erl_hold_counter_time_domain = 1000;
          s.erl_time_domain += 0.1f * (new_erl - s.erl_time_domain);
          s.erl_time_domain = fmaxr(s.erl_time_domain, 0.01f);
        }
      }
      --s.erl_hold_counter_time_domain;
      s.erl_time_domain = s.erl_hold_counter_time_domain > 0 ? s.erl_time_domain : fminr(1000.f, 2.f * s.erl_time_domain);
    }
    __syncwarp();
  }

  // ---- scalar state machines, lane 0 (shared)
  if (lane == 0) {
    s.saturated_echo = 0;
    if (saturated_capture) {
      if (usable_linear_before) {
        for (int c = 0; c < C; ++c)
          s.saturated_echo = s.saturated_echo || (mx.cv[c].metrics[5] > 20000.f || mx.cv[c].metrics[6] > 20000.f);
      } else {
        const float peak_echo_amplitude = max_sample * max_echo_path_gain * 10.f;
        s.saturated_echo = peak_echo_amplitude > 32000;
      }
    }
    s.init_strong_blocks += (active_render && !saturated_capture) ? 1 : 0;
    const int prev_initial_state = s.init_state;
    s.init_state = (float)s.init_strong_blocks < WAP_EC3(initial_state_seconds) * kNumBlocksPerSecond;
    s.init_transition_triggered = !s.init_state && prev_initial_state;
    ++s.tm_capture_block_counter;
    s.tm_strong_not_saturated_render_blocks += (active_render && !saturated_capture) ? 1 : 0;
    if (any_filter_consistent && delay < 5) {
      s.tm_sane_filter_observed = 1;
      s.tm_active_blocks_since_sane_filter = 0;
    } else if (active_render) {
      ++s.tm_active_blocks_since_sane_filter;
    }
    bool sane_filter_recently_seen;
    if (!s.tm_sane_filter_observed) sane_filter_recently_seen = s.tm_capture_block_counter <= 5 * kNumBlocksPerSecond;
    else sane_filter_recently_seen = s.tm_active_blocks_since_sane_filter <= 30 * kNumBlocksPerSecond;
    if (any_filter_converged) {
      s.tm_recent_convergence = 1;
      s.tm_active_non_converged_sequence_size = 0;
      s.tm_non_converged_sequence_size = 0;
      ++s.tm_num_converged_blocks;
    } else {
      if (++s.tm_non_converged_sequence_size > 20 * kNumBlocksPerSecond) s.tm_num_converged_blocks = 0;
      if (active_render && ++s.tm_active_non_converged_sequence_size > 60 * kNumBlocksPerSecond)
        s.tm_recent_convergence = 0;
    }
    if (!all_filters_diverged) s.tm_diverged_sequence_size = 0;
    else if (++s.tm_diverged_sequence_size >= 60) s.tm_non_converged_sequence_size = 10000;
    if (s.tm_active_non_converged_sequence_size > 60 * kNumBlocksPerSecond) s.tm_finite_erl_recently_detected = 0;
    if (s.tm_num_converged_blocks > 50) s.tm_finite_erl_recently_detected = 1;
    if (s.tm_finite_erl_recently_detected) s.tm_active = 0;
    else if (sane_filter_recently_seen && s.tm_recent_convergence) s.tm_active = 0;
    else s.tm_active = s.tm_strong_not_saturated_render_blocks > 6 * kNumBlocksPerSecond;
    const bool filter_update = active_render && !saturated_capture;
    s.fq_blocks_since_reset += filter_update ? 1 : 0;
    s.fq_blocks_since_start += filter_update ? 1 : 0;
    s.fq_convergence_seen = s.fq_convergence_seen || any_filter_converged;
    const bool sufficient_at_startup = (float)s.fq_blocks_since_start > kNumBlocksPerSecond * 0.4f;
    const bool sufficient_at_reset = sufficient_at_startup && (float)s.fq_blocks_since_reset > kNumBlocksPerSecond * 0.2f;
    bool usable = sufficient_at_startup && sufficient_at_reset;
    usable = usable && (ext_has || s.fq_convergence_seen);
    usable = usable && !s.tm_active;
    s.fq_usable = usable;
  }
  __syncwarp();

  // ---- ReverbModelEstimator::Update -> ReverbFrequencyResponse::Update, per capture channel
  for (int c = 0; c < C; ++c) {
    McChan& ch = mc.chan[c];
    Aec3Scalars& cs = mx.cs[c];
    __syncwarp();
    if (!cs.fb_has_erle_log2) continue;
    const float quality = fminr(1.f, fmaxr(0.f, cs.fb_inst_quality));
    const float* tail = ch.H2[cs.H2_size - 1];
    const float* direct = ch.H2[cs.fd_filter_delay];
    if (lane < 2) sc.red[16 + lane] = chain_sum(lane == 0 ? direct : tail, 1, kBins);
    __syncwarp();
    const float direct_path_energy = sc.red[16], tail_energy = sc.red[17];
    const float average_decay = direct_path_energy == 0.f ? 0.f : tail_energy / direct_path_energy;
    const float smoothing = 0.2f * quality;
    const float avg = cs.reverb_average_decay + smoothing * (average_decay - cs.reverb_average_decay);
    for (int k = lane; k < kBins; k += 32) r.v3[k] = fmaxr(tail[k], direct[k] * avg);
    __syncwarp();
    if (lane == 0) {
      cs.reverb_average_decay = avg;
      for (int k = 1; k < 64; ++k) {
        const float avg_neighbour = 0.5f * (r.v3[k - 1] + r.v3[k + 1]);
        r.v3[k] = fmaxr(r.v3[k], avg_neighbour);
      }
    }
    __syncwarp();
    for (int k = lane; k < kBins; k += 32) ch.tail_response[k] = r.v3[k];
  }
  __syncwarp();
}

// ComfortNoiseGenerator::Compute for capture channel c: spectra update + GenerateComfortNoise.  The counter /
// initial-estimate flags are those of the block's start (the caller advances them after the last channel);
// the seed advances by 63 draws per channel.
WAP_DEV void mc_cng_channel(McChan& ch, const EngineConfig& cfg, AecScratch& sc, McChanVec& cv, const float* nearend, bool hi) {
  const int lane = lane_id();
  Aec3Scalars& s = sc.s;
  AecRemoverScratch& r = sc.rm;
  const float noise_floor = cfg.cng_noise_floor;
  __syncwarp();
  const bool saturated_capture = s.capture_signal_saturation != 0;
  const int counter = s.cng_N2_counter;
  const bool has_initial = s.cng_has_initial != 0;
  const bool drop_initial = !saturated_capture && has_initial && counter + 1 == 1000;
  const bool use_initial = has_initial && !drop_initial;
  const unsigned seed = s.cng_seed;
  for (int k = lane; k < kBins; k += 32) {
    float N2 = ch.cng_N2[k], N2i = ch.cng_N2_initial[k];
    if (!saturated_capture) {
      float Y2s = ch.cng_Y2_smoothed[k];
      Y2s = Y2s + 0.1f * (nearend[k] - Y2s);
      ch.cng_Y2_smoothed[k] = Y2s;
      if (counter > 50) N2 = Y2s < N2 ? (0.9f * Y2s + 0.1f * N2) * 1.0002f : N2 * 1.0002f;
      if (use_initial) N2i = N2 > N2i ? N2i + 0.001f * (N2 - N2i) : N2;
      N2 = fmaxr(N2, noise_floor);
      if (use_initial) {
        N2i = fmaxr(N2i, noise_floor);
        ch.cng_N2_initial[k] = N2i;
      }
      ch.cng_N2[k] = N2;
    }
    const float N = sqrtf(use_initial ? N2i : N2);
    float re = 0.f, im = 0.f, fx = 0.f, fy = 0.f;
    if (k >= 1 && k < 64) {
      const unsigned seed_k = (kLcgA[k] * seed + kLcgC[k]) & 0x7fffffffu;
      const int i = (int)(seed_k >> 26);
      fx = kSqrt2Sin[i];
      fy = kSqrt2Sin[(i + 8) & 31];
      re = N * fx;
      im = N * fy;
    }
    cv.N_re[k] = re;
    cv.N_im[k] = im;
    if (hi) { r.v1[k] = N; cv.hb_re[k] = fx; cv.hb_im[k] = fy; }
  }
  __syncwarp();
  if (hi) {
    constexpr float kOneByNumBands = 1.f / (kBins / 2 + 1);
    float acc = 0.f;
    for (int k = kBins / 2; k < kBins; ++k) acc += r.v1[k];
    const float lvl = acc * kOneByNumBands;
    for (int k = lane; k < kBins; k += 32) {
      cv.hb_re[k] = lvl * cv.hb_re[k];
      cv.hb_im[k] = lvl * cv.hb_im[k];
    }
  }
  __syncwarp();
  if (lane == 0) s.cng_seed = (kLcgA[63] * seed + kLcgC[63]) & 0x7fffffffu;
  __syncwarp();
}

// ResidualEchoEstimator::Estimate -> mx.cv[c].R2 / R2_unb.
WAP_DEV void mc_residual_echo_estimate(Aec3State& sh, McState& mc, AecScratch& sc, McExtra& mx, int R, int C) {
  const int lane = lane_id();
  Aec3Scalars& s = sc.s;
  const McRender& rb = mc.render;
  __syncwarp();
  const bool dominant_nearend = s.dn_nearend_state != 0;
  const bool usable = s.fq_usable != 0;
  const bool saturated_echo = s.saturated_echo != 0;
  const bool transparent = s.tm_active != 0;
  const float echo_path_gain = transparent ? 0.01f * 0.01f : WAP_EC3(default_gain) * WAP_EC3(default_gain);
  const int delay = s.fd_min_filter_delay;
  const bool add_reverb = usable || !transparent;
  const float reverb_decay = dominant_nearend ? WAP_EC3(nearend_len) : WAP_EC3(default_len);
  const int first_reverb_partition = usable ? mx.cs[0].fa_filter_length_blocks + 1 : delay + 1;
  const int row_reverb = ring_off(s.spectra_read, first_reverb_partition, kRingBlocks);
  const int w0 = ring_off(s.spectra_read, imax(0, delay - 1), kRingBlocks);
  const int wn = delay + 1 - imax(0, delay - 1) + 1;
  for (int k = lane; k < kBins; k += 32) {
    // UpdateRenderNoisePower (render power summed over the channels)
    float floor = sh.X2_noise_floor[k];
    {
      float p;
      if (R > 1) {
        p = 0.f;
        for (int rc = 0; rc < R; ++rc) p += rb.spectra[s.spectra_read][rc][k];
      } else {
        p = rb.spectra[s.spectra_read][0][k];
      }
      int cnt = sh.X2_noise_floor_counter[k];
      if (p < floor) {
        floor = p;
        cnt = 0;
      } else if (cnt >= (int)WAP_EC3(noise_floor_hold)) {
        floor = fmaxr(floor * 1.1f, WAP_EC3(min_noise_floor_power));
      } else {
        ++cnt;
      }
      sh.X2_noise_floor[k] = floor;
      sh.X2_noise_floor_counter[k] = cnt;
    }
    float R2_nl = 0.f;
    if (!usable && !saturated_echo) {
      float X2 = 0.f;
      int idx = w0;
      for (int j = 0; j < wn; ++j) {
        float p;
        if (R > 1) {
          p = 0.f;
          for (int rc = 0; rc < R; ++rc) p += rb.spectra[idx][rc][k];
        } else {
          p = rb.spectra[idx][0][k];
        }
        X2 = fmaxr(X2, p);
        idx = ring_inc(idx, kRingBlocks);
      }
      if (WAP_EC3(noise_gate_power) > X2) X2 = fmaxr(0.f, X2 - WAP_EC3(noise_gate_slope) * (WAP_EC3(noise_gate_power) - X2));
      X2 -= WAP_EC3(stationary_gate_slope) * floor;
      X2 = fmaxr(0.f, X2);
      R2_nl = X2 * echo_path_gain;
    }
    float rev = 0.f;
    if (add_reverb) {
      float p;
      if (R > 1) {
        p = 0.f;
        for (int rc = 0; rc < R; ++rc) p += rb.spectra[row_reverb][rc][k];
      } else {
        p = rb.spectra[row_reverb][0][k];
      }
      const float scaling = usable ? mc.chan[0].tail_response[k] : echo_path_gain;
      rev = (sh.echo_reverb[k] + p * scaling) * reverb_decay;
      sh.echo_reverb[k] = rev;
    }
    for (int c = 0; c < C; ++c) {
      McChanVec& cv = mx.cv[c];
      float R2, R2u;
      if (saturated_echo) {
        R2 = R2u = cv.Y2[k];
      } else if (usable) {
        const float* erle = dominant_nearend ? mc.chan[c].erle : mc.chan[c].erle_onset_comp;
        R2 = cv.S2_lin[k] / erle[k];
        R2u = cv.S2_lin[k] / mc.chan[c].erle_unbounded[k];
      } else {
        R2 = R2u = R2_nl;
      }
      if (add_reverb) {
        R2 += rev;
        R2u += rev;
      }
      cv.R2[k] = R2;
      cv.R2_unb[k] = R2u;
    }
  }
  __syncwarp();
}

// SuppressionGain::GetGain: DominantNearendDetector, LowNoiseRenderDetector, LowerBandGain -> r.gain (amplitude).
// nearend_is_E2: the suppressor input is E2 (else Y2) of each channel.
WAP_DEV void mc_suppression_gain(Aec3State& sh, McState& mc, AecScratch& sc, McExtra& mx, int R, int C, bool nearend_is_E2,
                                 bool clock_drift) {
  const int lane = lane_id();
  Aec3Scalars& s = sc.s;
  AecRemoverScratch& r = sc.rm;
  __syncwarp();
  // chains: per channel nearend / R2_unbounded / N2 (bins 1..15); lane 3C: render block power
  if (lane < 3 * C) {
    const int c = lane / 3, w = lane % 3;
    const float* p = w == 0 ? (nearend_is_E2 ? mx.cv[c].E2 : mx.cv[c].Y2) : w == 1 ? mx.cv[c].R2_unb : mc.chan[c].cng_N2;
    sc.red[8 + lane] = chain_sum(p, 1, 16);
  } else if (lane == 3 * C) {
    float x2_sum = 0.f, x2_max = 0.f;
    for (int rc = 0; rc < R; ++rc) {
      const float* x = rc == 0 ? sc.x : mx.x1;
      for (int i = 0; i < kBlock; ++i) {
        const float x2 = x[i] * x[i];
        x2_sum += x2;
        x2_max = fmaxr(x2_max, x2);
      }
    }
    sc.red[19] = x2_sum / R;
    sc.red[20] = x2_max;
  }
  __syncwarp();
  if (lane == 0) {
    int nearend_state = 0;
    for (int c = 0; c < C; ++c) {
      Aec3Scalars& cs = mx.cs[c];
      const float ne_sum = sc.red[8 + 3 * c], echo_sum = sc.red[9 + 3 * c], noise_sum = sc.red[10 + 3 * c];
      if (echo_sum < WAP_EC3(dn_enr_threshold) * ne_sum && ne_sum > WAP_EC3(dn_snr_threshold) * noise_sum) {
        if (++cs.dn_trigger_counter >= WAP_EC3(dn_trigger_threshold)) {
          cs.dn_hold_counter = WAP_EC3(dn_hold_duration);
          cs.dn_trigger_counter = WAP_EC3(dn_trigger_threshold);
        }
      } else {
        cs.dn_trigger_counter = imax(0, cs.dn_trigger_counter - 1);
      }
      if (echo_sum > WAP_EC3(dn_enr_exit_threshold) * ne_sum && echo_sum > WAP_EC3(dn_snr_threshold) * noise_sum) cs.dn_hold_counter = 0;
      cs.dn_hold_counter = imax(0, cs.dn_hold_counter - 1);
      nearend_state = nearend_state || cs.dn_hold_counter > 0;
    }
    s.dn_nearend_state = nearend_state;
    const float x2_sum = sc.red[19], x2_max = sc.red[20];
    constexpr float kThreshold = 50.f * 50.f * 64.f;
    sc.ired[8] = s.sg_average_power < kThreshold && x2_max < 3 * s.sg_average_power;
    s.sg_average_power = s.sg_average_power * 0.9f + x2_sum * 0.1f;
  }
  __syncwarp();
  const bool low_noise_render = sc.ired[8] != 0;
  const bool nearend_state = s.dn_nearend_state != 0;
  const bool saturated_echo = s.saturated_echo != 0;
  const Ec3Tuning& tun = nearend_state ? WAP_EC3_ARR(nearend_tuning) : WAP_EC3_ARR(normal_tuning);
  const float min_echo_power = low_noise_render ? WAP_EC3(low_render_limit) : WAP_EC3(normal_render_limit);
  for (int k = lane; k < kBins; k += 32) {
    const float last_gain = sh.last_gain[k];
    const float max_gain = fminr(fmaxr(last_gain * tun.max_inc, WAP_EC3(floor_first_increase)), 1.f);
    float enr_t, enr_s, emr_t;
    gain_params(sc, tun, k, &enr_t, &enr_s, &emr_t);
    const float masker = mc.chan[0].cng_N2[k];   // comfort_noise[0] for every channel (suppression_gain.cc:321)
    float gain = 1.f;
    for (int c = 0; c < C; ++c) {
      McChan& ch = mc.chan[c];
      const int mem_index = mx.cs[c].sg_nearend_mem_index;
      const float in = nearend_is_E2 ? mx.cv[c].E2[k] : mx.cv[c].Y2[k];
      float ne = in;
      ne = ch.nearend_mem[0][k] + ne;
      ne = ch.nearend_mem[1][k] + ne;
      ne = ch.nearend_mem[2][k] + ne;
      ne *= 0.25f;
      ch.nearend_mem[mem_index][k] = in;
      const float audibility = k < 3 ? WAP_EC3(audibility_threshold_lf) : (k < 7 ? WAP_EC3(audibility_threshold_mf) : WAP_EC3(audibility_threshold_hf));
      const float threshold = WAP_EC3(floor_power) * audibility;
      const float normalizer = 1.f / (threshold - WAP_EC3(floor_power));
      const float echo = mx.cv[c].R2[k];
      float weighted = echo;
      if (echo < threshold) {
        const float tmp = (threshold - echo) * normalizer;
        weighted = echo * fmaxr(0.f, 1.f - tmp * tmp);
      }
      float min_gain = 0.f;
      if (!saturated_echo) {
        min_gain = weighted > 0.f ? min_echo_power / weighted : 1.f;
        min_gain = fminr(min_gain, 1.f);
        if (k <= WAP_EC3(last_lf_smoothing_band)) {
          if (ch.last_nearend[k] > ch.last_echo[k] || k <= WAP_EC3(last_permanent_lf_smoothing_band)) {
            min_gain = fmaxr(min_gain, last_gain * tun.max_dec_lf);
            min_gain = fminr(min_gain, 1.f);
          }
        }
      }
      const float enr = weighted / (ne + 1.f);
      const float emr = weighted / (masker + 1.f);
      float g = 1.0f;
      if (enr > enr_t && emr > emr_t) {
        g = (enr_s - enr) / (enr_s - enr_t);
        g = fmaxr(g, emr_t / emr);
      }
      g = fmaxr(fminr(g, max_gain), min_gain);
      gain = fminr(gain, g);
      ch.last_nearend[k] = ne;
      ch.last_echo[k] = weighted;
    }
    r.gain[k] = gain;
  }
  __syncwarp();
  {
    const float g12 = fminr(r.gain[1], r.gain[2]);
    const bool limit_hf = !nearend_state || clock_drift;
    float min_upper_gain = 1.f;
    for (int band = WAP_EC3(limiting_gain_band); band < WAP_EC3(limiting_gain_band) + WAP_EC3(bands_in_limiting_gain); ++band)
      min_upper_gain = fminr(min_upper_gain, r.gain[band]);
    const bool limit_bands = WAP_EC3(bands_in_limiting_gain) > 0;
    const float g63 = limit_hf ? (limit_bands ? fminr(r.gain[63], min_upper_gain) : r.gain[63]) : r.gain[63];
    __syncwarp();
    for (int k = lane; k < kBins; k += 32) {
      float g = r.gain[k];
      if (k <= 1) g = g12;
      if (limit_hf) {
        if (limit_bands && k > WAP_EC3(limiting_gain_band)) g = fminr(g, min_upper_gain);
        if (k == 64) g = g63;
      }
      sh.last_gain[k] = g;
      r.gain[k] = sqrtf(g);
    }
  }
  if (lane == 0)
    for (int c = 0; c < C; ++c) mx.cs[c].sg_nearend_mem_index = (mx.cs[c].sg_nearend_mem_index + 1) % 3;
  __syncwarp();
}

// SuppressionGain::UpperBandsGain with R render channels (max_gain_during_echo == 1: no echo bound).
WAP_DEV float mc_upper_bands_gain(const McRender& rb, AecScratch& sc, int R, int B) {
  const int lane = lane_id();
  const Aec3Scalars& s = sc.s;
  AecRemoverScratch& r = sc.rm;
  __syncwarp();
  if (s.rsa_has_narrow_peak && s.rsa_narrow_peak_band > kBins - 10) return 0.001f;
  float g = r.gain[32 + lane];
  if (lane == 0) g = fminr(g, r.gain[64]);
  for (int m = 16; m; m >>= 1) g = fminf(g, __shfl_xor_sync(WAP_FULL, g, m));
  const float gain_below_8_khz = g;
  if (s.saturated_echo) return fminr(0.001f, gain_below_8_khz);
  // lane = band * R + rc: one serial sum of squares each
  if (lane < B * R) {
    const int band = lane / R, rc = lane % R;
    const float* p = rb.blocks[s.blocks_read][rc][band];
    float acc = 0.f;
    for (int i = 0; i < kBlock; ++i) acc = acc + p[i] * p[i];
    sc.red[16 + lane] = acc;
  }
  __syncwarp();
  float low_band_energy = 0.f, high_band_energy = 0.f;
  for (int rc = 0; rc < R; ++rc) low_band_energy = fmaxr(low_band_energy, sc.red[16 + rc]);
  for (int band = 1; band < B; ++band)
    for (int rc = 0; rc < R; ++rc) high_band_energy = fmaxr(high_band_energy, sc.red[16 + band * R + rc]);
  float anti_howling_gain;
  const float activation_threshold = kBlock * WAP_EC3(hb_anti_howling_activation_threshold);
  if (high_band_energy < fmaxr(low_band_energy, activation_threshold)) anti_howling_gain = 1.f;
  else anti_howling_gain = WAP_EC3(hb_anti_howling_gain) * sqrtf(low_band_energy / high_band_energy);
  __syncwarp();
  return fminr(fminr(gain_below_8_khz, anti_howling_gain), 1.f);
}

// EchoRemoverImpl::ProcessCapture for capture block b of the tick: mx.cv[c].y in / out; the upper bands of
// the block (mt.capture_blocks[b][c][1..]) are updated in place.
WAP_DEV void mc_echo_remover_process_capture(Aec3State& sh, McState& mc, const EngineConfig& cfg, AecScratch& sc, McExtra& mx,
                                             EchoPathVariability v, bool capture_signal_saturation, int ext_has,
                                             int ext_delay, int b, int R, int C) {
  const int lane = lane_id();
  Aec3Scalars& s = sc.s;
  AecRemoverScratch& r = sc.rm;
  McRender& rb = mc.render;
  McTick& mt = mc.tick;
  const int B = cfg.num_bands;
  __syncwarp();
  for (int i = lane; i < kBlock; i += 32) {
    sc.x[i] = rb.blocks[s.blocks_read][0][0][i];
    if (R > 1) mx.x1[i] = rb.blocks[s.blocks_read][1][0][i];
  }
  if (lane == 0) s.capture_signal_saturation = capture_signal_saturation;
  if (v.delay_change != kDelayAdjNone || v.gain_change) {
    if (v.gain_change) {
      const bool act = s.er_gain_change_hangover == 0;
      __syncwarp();
      if (act) {
        if (lane == 0) s.er_gain_change_hangover = 3;
      } else {
        v.gain_change = 0;
      }
    }
    __syncwarp();
    for (int c = 0; c < C; ++c) mc_subtractor_handle_echo_path_change(mc.filt[c], mc.chan[c], mx.cs[c], sc, v, R);
    mc_aec_state_handle_echo_path_change(mc, mx, sc, v, C);
    if (v.delay_change != kDelayAdjNone && lane == 0) s.sg_initial_state = 1;
  }
  __syncwarp();
  if (lane == 0) {
    if (s.er_gain_change_hangover > 0) --s.er_gain_change_hangover;
  }
  __syncwarp();

  mc_render_signal_analyzer_update(sh, rb, sc, mx.x1, s.fd_min_filter_delay, R,
                                   B > 1 ? rb.blocks[s.blocks_read][0][1] : nullptr,
                                   (B > 1 && R > 1) ? rb.blocks[s.blocks_read][1][1] : nullptr);

  if (s.init_transition_triggered) {
    if (lane == 0) {
      for (int c = 0; c < C; ++c) mc_subtractor_exit_initial_state(mx.cs[c], sc);
      s.sg_initial_state = 0;
    }
    __syncwarp();
  }

  // Subtractor::Process: the render powers follow channel 0's filter sizes
  mc_spectral_sums(rb, sc, mx.cs[0].fr_current_size, mx.cs[0].fc_current_size, R);
  for (int c = 0; c < C; ++c) {
    McChan& ch = mc.chan[c];
    McChanVec& cv = mx.cv[c];
    mc_subtractor_process_channel(sh, rb, mc.filt[c], ch, mx.cs[c], sc, cv, R, s.capture_signal_saturation != 0);
    // The estimator vectors the stages behind the linear filters read (per channel: H_error .. the
    // time-domain memories; once per leg: StreamState::aec's), asked for now that the filter passes no
    // longer stream through the L1 (same placement as the mono kernel, dsp_aec3_remover.cuh).
    warp_prefetch_l1(ch.H_error, (int)(reinterpret_cast<const char*>(&ch + 1) - reinterpret_cast<const char*>(ch.H_error)));
    if (c == C - 1)
      warp_prefetch_l1(sh.H_error, (int)(reinterpret_cast<const char*>(sh.render_decimator) - reinterpret_cast<const char*>(sh.H_error)));
    // FormLinearFilterOutput (refined_filter_output_last_selected_ is one flag for all channels)
    {
      const float y2 = cv.metrics[0], e2_refined = cv.metrics[1], e2_coarse = cv.metrics[2], s2_refined = cv.metrics[3],
                  s2_coarse = cv.metrics[4];
      bool use_refined_output = true;
      if (e2_coarse < 0.9f * e2_refined && y2 > 30.f * 30.f * kBlock &&
          (s2_refined > 60.f * 60.f * kBlock || s2_coarse > 60.f * 60.f * kBlock)) {
        use_refined_output = false;
      } else if (e2_coarse < e2_refined && y2 < e2_refined) {
        use_refined_output = false;
      }
      const float* from = s.er_refined_last_selected ? r.e_ref : r.e_coa;
      const float* to = use_refined_output ? r.e_ref : r.e_coa;
      __syncwarp();
      for (int i = lane; i < kBlock; i += 32) {
        float o = to[i];
        if (from != to && i < 30) {
          const float aa = (i + 1) * (1.f / 31);
          o = aa * to[i] + (1.f - aa) * from[i];
        }
        cv.e[i] = o;
      }
      __syncwarp();
      if (lane == 0) s.er_refined_last_selected = use_refined_output;
    }
    for (int i = lane; i < kBlock; i += 32) {
      sc.fftA[i] = ch.y_old[i] * kSqrtHanning128[i];
      sc.fftA[kBlock + i] = cv.y[i] * kSqrtHanning128[kBlock + i];
      sc.fftB[i] = ch.e_old[i] * kSqrtHanning128[i];
      sc.fftB[kBlock + i] = cv.e[i] * kSqrtHanning128[kBlock + i];
      ch.y_old[i] = cv.y[i];
      ch.e_old[i] = cv.e[i];
    }
    fft_pair(sc, false, true);
    packed_to_reim(sc.fftA, cv.Y_re, cv.Y_im);
    packed_to_reim(sc.fftB, cv.E_re, cv.E_im);
    __syncwarp();
    for (int k = lane; k < kBins; k += 32) {
      const float dr = cv.Y_re[k] - cv.E_re[k], di = cv.Y_im[k] - cv.E_im[k];
      cv.S2_lin[k] = dr * dr + di * di;
      cv.Y2[k] = power_bin(cv.Y_re[k], cv.Y_im[k], k);
      cv.E2[k] = power_bin(cv.E_re[k], cv.E_im[k], k);
    }
    __syncwarp();
  }
  const bool nearend_is_E2 = s.fq_usable != 0;
  mc_aec_state_update(sh, mc, sc, mx, R, C, ext_has, ext_delay);

  for (int c = 0; c < C; ++c)
    mc_cng_channel(mc.chan[c], cfg, sc, mx.cv[c], nearend_is_E2 ? mx.cv[c].E2 : mx.cv[c].Y2, B > 1);
  if (lane == 0) {
    const bool saturated_capture = s.capture_signal_saturation != 0;
    if (!saturated_capture && s.cng_has_initial) {
      if (++s.cng_N2_counter == 1000) s.cng_has_initial = 0;
    }
  }
  __syncwarp();

  if (cfg.capture_output_used) {
    mc_residual_echo_estimate(sh, mc, sc, mx, R, C);
    const bool usable = s.fq_usable != 0;
    if (usable) {
      for (int c = 0; c < C; ++c)
        for (int k = lane; k < kBins; k += 32) mx.cv[c].E2[k] = fminr(mx.cv[c].E2[k], mx.cv[c].Y2[k]);
      __syncwarp();
    }
    mc_suppression_gain(sh, mc, sc, mx, R, C, nearend_is_E2, v.clock_drift != 0);
    float high_bands_gain = 1.f;
    if (B > 1) high_bands_gain = mc_upper_bands_gain(rb, sc, R, B);
    const float noise_scaling = 0.4f * sqrtf(1.f - high_bands_gain * high_bands_gain);
    constexpr float kIfftNormalization = 2.f / 128;
    for (int c = 0; c < C; ++c) {
      McChan& ch = mc.chan[c];
      McChanVec& cv = mx.cv[c];
      const float* Yf_re = usable ? cv.E_re : cv.Y_re;
      const float* Yf_im = usable ? cv.E_im : cv.Y_im;
      __syncwarp();
      for (int k = lane; k < kBins; k += 32) {
        const float g = r.gain[k];
        const float noise_gain = sqrtf(1.f - g * g);
        const float E_real = Yf_re[k] * g;
        const float E_imag = Yf_im[k] * g;
        const float re = E_real + noise_gain * cv.N_re[k];
        const float im = E_imag + noise_gain * cv.N_im[k];
        if (k == 0) sc.fftA[0] = re;
        else if (k == 64) sc.fftA[1] = re;
        else { sc.fftA[2 * k] = re; sc.fftA[2 * k + 1] = im; }
        if (B > 1) {
          if (k == 0) sc.fftB[0] = 0.f;
          else if (k == 64) sc.fftB[1] = 0.f;
          else { sc.fftB[2 * k] = cv.hb_re[k]; sc.fftB[2 * k + 1] = cv.hb_im[k]; }
        }
      }
      fft_pair(sc, true, B > 1);
      for (int i = lane; i < kBlock; i += 32) {
        float e0 = ch.e_output_old[i] * kSqrtHanning128[kBlock + i];
        e0 += sc.fftA[i] * kSqrtHanning128[i];
        e0 = e0 * kIfftNormalization;
        ch.e_output_old[i] = sc.fftA[kBlock + i];
        cv.y[i] = clampr(e0, -32768.f, 32767.f);
      }
      if (B > 1) {
        const float ngain = noise_scaling * kIfftNormalization;
        __syncwarp();
        for (int band = 1; band < B; ++band)
          for (int i = lane; i < kBlock; i += 32) {
            float e1 = mt.capture_blocks[b][c][band][i] * high_bands_gain;
            if (band == 1) e1 += sc.fftB[i] * ngain;
            const float o1 = mc.e_output_old_hi[c][band - 1][i];
            mc.e_output_old_hi[c][band - 1][i] = e1;
            mt.capture_blocks[b][c][band][i] = clampr(o1, -32768.f, 32767.f);
          }
      }
    }
  }
  __syncwarp();
}

}  // namespace wap
