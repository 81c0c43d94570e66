// AEC3 echo remover for one call leg (mono render / mono capture): everything
// EchoRemoverImpl::ProcessCapture runs after the linear stage.
//   EchoRemoverImpl::{ProcessCapture, FormLinearFilterOutput}     aec3/echo_remover.cc:254-529
//   AecState (+ InitialState, FilterDelay, FilteringQualityAnalyzer,
//             SaturationDetector)                                 aec3/aec_state.cc:115-496
//   SubtractorOutputAnalyzer                                      aec3/subtractor_output_analyzer.cc:26-66
//   FilterAnalyzer (+ ConsistentFilterDetector)                   aec3/filter_analyzer.cc:80-291
//   LegacyTransparentModeImpl                                     aec3/transparent_mode.cc:132-233
//   SubbandErleEstimator / FullBandErleEstimator / ErlEstimator   aec3/subband_erle_estimator.cc:72-259,
//                                                                 aec3/fullband_erle_estimator.cc:50-192,
//                                                                 aec3/erl_estimator.cc:43-149
//   ReverbModel / ReverbFrequencyResponse                         aec3/reverb_model.cc:25-53,
//                                                                 aec3/reverb_frequency_response.cc:61-106
//   ResidualEchoEstimator                                         aec3/residual_echo_estimator.cc:193-425
//   ComfortNoiseGenerator                                         aec3/comfort_noise_generator.cc:60-192
//   SuppressionGain (+ DominantNearendDetector, MovingAverage)    aec3/suppression_gain.cc:124-477,
//                                                                 aec3/dominant_nearend_detector.cc:37-81
//   SuppressionFilter::ApplyGain                                  aec3/suppression_filter.cc:88-183
// Lanes own frequency bins (k = lane, lane + 32, 64).  The reference's
// std::accumulate reductions are evaluated as left-to-right chains, several of
// them side by side on different lanes, and exchanged through shared memory.
#pragma once

#include "dsp_aec3_common.cuh"
#include "dsp_aec3_subtractor.cuh"
#include "wap_libm.cuh"

namespace wap {


constexpr float kX2BandEnergyThreshold = 44015068.0f;  // subband/fullband ERLE, ERL kX2Min
// (active_render_limit^2) * kFftLengthBy2 (aec_state.cc:235-237)
#define kActiveRenderEnergy ((WAP_EC3(active_render_limit) * WAP_EC3(active_render_limit)) * 64.f)

// Left-to-right sum of p[from..to) -- the order std::accumulate uses.
WAP_DEV float chain_sum(const float* p, int from, int to) {
  float s = 0.f;
  for (int i = from; i < to; ++i) s += p[i];
  return s;
}

// ---- ErleEstimator::Reset / FullBandErleEstimator::Reset / SubbandErleEstimator::Reset
WAP_DEV void erle_reset(Aec3State& a, AecScratch& sc, bool delay_change) {
  const int lane = lane_id();
  #pragma unroll
  for (int k = lane; k < kBins; k += 32) {
    a.erle[k] = WAP_EC3(erle_min);
    a.erle_onset_comp[k] = WAP_EC3(erle_min);
    a.erle_unbounded[k] = WAP_EC3(erle_min);
    if (WAP_EC3(erle_num_sections) > 1) { a.sd_erle[k] = WAP_EC3(erle_min); a.sd_erle_onset[k] = WAP_EC3(erle_min); }
    a.coming_onset[k] = 1;
    a.erle_hold_counters[k] = 0;
    a.accum_Y2[k] = 0.f;
    a.accum_E2[k] = 0.f;
    a.accum_low_render[k] = 0;
  }
  if (lane == 0) {
    Aec3Scalars& s = sc.s;
    s.erle_num_points = 0;
    // ErleInstantaneous::Reset
    s.fb_has_erle_log2 = 0;
    s.fb_inst_quality = 0.f;
    s.fb_num_points = 0;
    s.fb_E2_acum = 0.f;
    s.fb_Y2_acum = 0.f;
    s.fb_max_erle_log2 = -10.f;
    s.fb_min_erle_log2 = 33.f;
    s.fb_erle_time_domain_log2 = fast_approx_log2f(WAP_EC3(erle_min) + 1e-3f);
    s.fb_hold_counter = 0;
    if (delay_change) s.erle_blocks_since_reset = 0;
  }
  if (WAP_EC3(erle_num_sections) > 1) {   // SignalDependentErleEstimator::Reset
    for (int i = lane; i < kMaxPartitions * 8; i += 32) {
      a.sd_estimators[i >> 3][i & 7] = WAP_EC3(erle_min);
      a.sd_correction[i >> 3][i & 7] = 1.0f;
    }
    if (lane < 8) { a.sd_erle_ref[lane] = WAP_EC3(erle_min); a.sd_num_updates[lane] = 0; }
  }
  __syncwarp();
}

// ---- SignalDependentErleEstimator::Update (signal_dependent_erle_estimator.cc:177-372), erle.num_sections > 1.
// The refined filter is cut into sections (boundaries computed on the host); per bin, the number of leading
// sections that carry 90 % of the echo estimate selects a correction factor per (section count, subband) that
// scales the subband estimator's ERLE.  In: r.v1 = X2 with reverb, r.Y2, r.E2, the subband estimator's a.erle /
// a.erle_onset_comp of this block.  Scratch: r.R2 (active sections per bin), r.R2_unb (subband sums).
#if WAP_EC3_RUNTIME
WAP_DEV int sd_subband_of_bin(int k) { return k < 8 ? 0 : k < 16 ? 1 : k < 24 ? 2 : k < 32 ? 3 : k < 48 ? 4 : 5; }
WAP_DEV int sd_band_boundary(int i) { return i == 0 ? 1 : i < 5 ? 8 * i : i == 5 ? 48 : kBins; }   // kBandBoundaries
WAP_DEV float sd_safe_clamp(float x, float lo, float hi) { return x <= lo ? lo : x >= hi ? hi : x; }
WAP_DEV void signal_dependent_erle_update(Aec3State& a, AecScratch& sc, bool converged) {
  const int lane = lane_id();
  Aec3Scalars& s = sc.s;
  AecRemoverScratch& r = sc.rm;
  const int num_sections = sc.ep.erle_num_sections;
  const int* bounds = sc.ep.sd_boundaries;
  const int H2_size = s.H2_size;
  __syncwarp();
  // ComputeEchoEstimatePerFilterSection + ComputeActiveFilterSections: the accumulated section estimates of a
  // bin never decrease, so the first section that reaches 0.9 of the total is where the reference's downward
  // scan stops; two passes over the sections (total, then the first hit) with the same operations in the
  // same order
  const int idx_first = ring_off(s.spectra_read, bounds[0], kRingBlocks);
  for (int k = lane; k < kBins; k += 32) {
    float target = 0.f;
    int n_active = 0;
    for (int pass = 0; pass < 2; ++pass) {
      int idx = idx_first;
      float accum = 0.f;
      bool found = false;
      for (int sec = 0; sec < num_sections; ++sec) {
        float X2_section = 0.f, H2_section = 0.f;
        const int limit = imin(bounds[sec + 1], H2_size);
        for (int block = bounds[sec]; block < limit; ++block) {
          X2_section += a.spectra[idx][k] * 1.f;
          H2_section = H2_section + a.H2[block][k];
          idx = ring_inc(idx, kRingBlocks);
        }
        const float S2 = X2_section * H2_section;
        accum = sec == 0 ? S2 : accum + S2;
        if (pass == 1 && !found && accum >= target) { found = true; n_active = sec; }
      }
      target = 0.9f * accum;
    }
    r.R2[k] = (float)n_active;
  }
  __syncwarp();
  // UpdateCorrectionFactors
  if (converged) {
    if (lane < 18) {   // subband powers of X2, E2, Y2: left-to-right sums from 0
      const int sb = lane % 6;
      const float* p = lane < 6 ? r.v1 : lane < 12 ? r.E2 : r.Y2;
      r.R2_unb[lane] = chain_sum(p, sd_band_boundary(sb), sd_band_boundary(sb + 1));
    } else if (lane < 24) {   // the fewest active sections of the subband's bins
      const int sb = lane - 18;
      float m = r.R2[sd_band_boundary(sb)];
      for (int k = sd_band_boundary(sb) + 1; k < sd_band_boundary(sb + 1); ++k) m = fminr(m, r.R2[k]);
      r.R2_unb[lane] = m;
    }
    __syncwarp();
    if (lane < 6) {
      const int sb = lane;
      const float X2_sb = r.R2_unb[sb], E2_sb = r.R2_unb[6 + sb], Y2_sb = r.R2_unb[12 + sb];
      const int idx = (int)r.R2_unb[18 + sb];
      const float max_erle = sb < 4 ? WAP_EC3(erle_max_l) : WAP_EC3(erle_max_h);
      float new_erle = 0.f;
      bool updated = false;
      int num_updates = a.sd_num_updates[sb];
      if (X2_sb > kX2BandEnergyThreshold && E2_sb > 0.f) {
        new_erle = Y2_sb / E2_sb;
        updated = true;
        ++num_updates;
      }
      float est = a.sd_estimators[idx][sb];
      float alpha = new_erle > est ? 0.05f : 0.1f;
      alpha = (updated ? 1.f : 0.f) * alpha;
      est += alpha * (new_erle - est);
      est = sd_safe_clamp(est, WAP_EC3(erle_min), max_erle);
      float ref = a.sd_erle_ref[sb];
      alpha = new_erle > ref ? 0.05f : 0.1f;
      alpha = (updated ? 1.f : 0.f) * alpha;
      ref += alpha * (new_erle - ref);
      ref = sd_safe_clamp(ref, WAP_EC3(erle_min), max_erle);
      if (updated && num_updates > 50) {
        const float new_correction_factor = est / ref;
        const float cf = a.sd_correction[idx][sb];
        a.sd_correction[idx][sb] = cf + 0.1f * (new_correction_factor - cf);
      }
      a.sd_estimators[idx][sb] = est;
      a.sd_erle_ref[sb] = ref;
      a.sd_num_updates[sb] = num_updates;
    }
    __syncwarp();
  }
  for (int k = lane; k < kBlock; k += 32) {   // bins 0..63; bin 64 keeps the value of the last Reset
    const int sb = sd_subband_of_bin(k);
    const float max_erle = sb < 4 ? WAP_EC3(erle_max_l) : WAP_EC3(erle_max_h);
    const float correction_factor = a.sd_correction[(int)r.R2[k]][sb];
    a.sd_erle[k] = sd_safe_clamp(a.erle[k] * correction_factor, WAP_EC3(erle_min), max_erle);
    if (WAP_EC3(erle_onset_detection))
      a.sd_erle_onset[k] = sd_safe_clamp(a.erle_onset_comp[k] * correction_factor, WAP_EC3(erle_min), max_erle);
  }
  __syncwarp();
}
#else
WAP_DEV void signal_dependent_erle_update(Aec3State&, AecScratch&, bool) {}
#endif

// ---- AecState::ReverbDecay(mild) -> ReverbDecayEstimator::Decay (reverb_decay_estimator.h:37-43)
WAP_DEV float aec_reverb_decay(const AecScratch& sc, bool mild) {
  if (WAP_EC3(default_len) < 0.f) return sc.s.rd_decay;   // use_adaptive_echo_decay_
  return mild ? fabsf(WAP_EC3(nearend_len)) : WAP_EC3(default_len);
}

// ---- ReverbDecayEstimator::Update (reverb_decay_estimator.cc:104-143) with the adaptive decay
// (ep_strength.default_len < 0), on a block that is not stationary.  `filter` is the FilterAnalyzer's
// high-passed impulse response, the peak block the direct-path filter delay.
// The EarlyReverbLengthEstimator is not restated: with filter.refined.length_blocks <= 17 it holds fewer
// than the kNumSectionsToAnalyze = 9 sections its Estimate() needs (:374-376), so the early reverb size is
// always 0 and its accumulators are never read.
static_assert(kMaxPartitions - 3 - 6 < 9, "EarlyReverbLengthEstimator::Estimate would become reachable");
WAP_DEV void reverb_decay_update(Aec3State& a, AecScratch& sc, bool has_quality, float quality) {
  const int lane = lane_id();
  Aec3Scalars& s = sc.s;
  AecRemoverScratch& r = sc.rm;
  const int L = WAP_EC3(refined_len);
  const int delay = s.fd_filter_delay;
  const float* h = a.h_highpass;
  __syncwarp();
  const bool feasible = delay <= L - 3 - 1 && s.fa_hp_size == L * kBlock && delay > 0 && s.fq_usable_filter;
  const float smoothing = fmaxr(has_quality ? quality * 0.2f : 0.f, s.rd_smoothing);
  const int b = s.rd_block_to_analyze;
  __syncwarp();
  if (!feasible) {
    if (lane == 0) {   // ResetDecayEstimation (:145-154); LateReverbLinearRegressor::Reset(0)
      s.rd_late_nz = 0.f; s.rd_late_nn = 0.f; s.rd_late_count = 0.f; s.rd_late_N = 0; s.rd_late_n = 0;
      s.rd_block_to_analyze = 0; s.rd_candidate_size = 0; s.rd_region_identified = 0;
      s.rd_smoothing = 0.f; s.rd_late_start = 0; s.rd_late_end = 0;
    }
    __syncwarp();
    return;
  }
  if (lane == 0) s.rd_smoothing = smoothing;
  if (smoothing == 0.f) { __syncwarp(); return; }
  if (b < L) {
    // ---- AnalyzeFilter (:212-247)
    const float* hb = h + b * kBlock;
    float h2_log2[kBlock / 32];
    #pragma unroll
    for (int j = 0; j < kBlock / 32; ++j) {
      const float h2 = hb[lane + 32 * j] * hb[lane + 32 * j];
      r.v3[lane + 32 * j] = h2;
      h2_log2[j] = fast_approx_log2f((float)((double)h2 + 1e-10));
    }
    __syncwarp();
    float gain = 0.f;
    if (lane == 0) gain = fmaxr(chain_sum(r.v3, 0, kBlock) * (1.f / kBlock), 1e-32f);   // AnalyzeBlockGain
    __syncwarp();
    #pragma unroll
    for (int j = 0; j < kBlock / 32; ++j) r.v3[lane + 32 * j] = h2_log2[j];
    __syncwarp();
    if (lane == 0) {
      const float previous = a.rd_previous_gains[b];
      const bool adapting = previous > 1.1f * gain || previous < 0.9f * gain;
      const bool above_noise_floor = gain > s.rd_tail_gain;
      a.rd_previous_gains[b] = gain;
      s.rd_region_identified = s.rd_region_identified || adapting || !above_noise_floor;
      if (!s.rd_region_identified) ++s.rd_candidate_size;
      if (b <= s.rd_late_end && b >= s.rd_late_start) {
        float nz = s.rd_late_nz, count = s.rd_late_count;   // LateReverbLinearRegressor::Accumulate
        for (int i = 0; i < kBlock; ++i) {
          nz += count * r.v3[i];
          count += 1.f;
        }
        s.rd_late_nz = nz; s.rd_late_count = count; s.rd_late_n += kBlock;
      }
      s.rd_block_to_analyze = b + 1;
    }
  } else {
    // ---- EstimateDecay (:156-210)
    const int first_block = imin(delay + 3, L);
    if (lane < 3) {
      const float* blk = h + (lane == 0 ? first_block : lane == 1 ? L - 1 : delay) * kBlock;
      float acc = 0.f;
      if (lane < 2) {   // BlockEnergyAverage
        for (int i = 0; i < kBlock; ++i) acc = acc + blk[i] * blk[i];
        acc *= (1.f / kBlock);
      } else {          // BlockEnergyPeak: the first largest square
        for (int i = 0; i < kBlock; ++i) acc = (acc < blk[i] * blk[i]) ? blk[i] * blk[i] : acc;
      }
      sc.red[16 + lane] = acc;
    }
    __syncwarp();
    if (lane == 0) {
      const float first_reverb_gain = sc.red[16], tail_gain = sc.red[17], peak_energy = sc.red[18];
      s.rd_tail_gain = tail_gain;
      const bool sufficient_reverb_decay = first_reverb_gain > 4.f * tail_gain;
      const bool valid_filter = first_reverb_gain > 2.f * tail_gain && peak_energy < 100.f;
      const int size_late_reverb = imax(s.rd_candidate_size, 0);   // early reverb size 0, see above
      if (size_late_reverb >= 5) {
        if (valid_filter && s.rd_late_n == s.rd_late_N && s.rd_late_N != 0) {
          float decay = libm_pow2f((s.rd_late_nz / s.rd_late_nn) * (float)kBlock);
          decay = fmaxr(.97f * s.rd_decay, decay);
          decay = fminr(decay, 0.95f);
          decay = fmaxr(decay, 0.02f);
          s.rd_decay += smoothing * (decay - s.rd_decay);
        }
        const int N = size_late_reverb * kBlock;   // LateReverbLinearRegressor::Reset(N)
        s.rd_late_nz = 0.f;
        s.rd_late_nn = (float)N * ((float)(N * N) - 1.0f) * (1.f / 12.f);
        s.rd_late_count = -(float)N * 0.5f + 0.5f;
        s.rd_late_N = N; s.rd_late_n = 0;
        s.rd_late_start = delay + 3;
        s.rd_late_end = first_block + s.rd_candidate_size - 1;
      } else {
        s.rd_late_nz = 0.f; s.rd_late_nn = 0.f; s.rd_late_count = 0.f; s.rd_late_N = 0; s.rd_late_n = 0;
        s.rd_late_start = 0; s.rd_late_end = 0;
      }
      s.rd_block_to_analyze = first_block;
      s.rd_region_identified = !(valid_filter && sufficient_reverb_decay);
      s.rd_candidate_size = 0;
      s.rd_smoothing = 0.f;
    }
  }
  __syncwarp();
}

// ---- FilterAnalyzer::Reset (filter_analyzer.cc:71-78), lane 0
WAP_DEV void filter_analyzer_reset(AecScratch& sc) {
  Aec3Scalars& s = sc.s;
  s.fa_blocks_since_reset = 0;
  s.fa_region_start = 0;
  s.fa_region_end = 0;
  s.fa_peak_index = 0;
  s.fa_gain = WAP_EC3(default_gain);
  s.cfd_significant_peak = 0;
  s.cfd_floor_accum = 0.f;
  s.cfd_secondary_peak = 0.f;
  s.cfd_floor_low_limit = 0;
  s.cfd_floor_high_limit = 0;
  s.cfd_consistent_counter = 0;
  s.cfd_consistent_delay_reference = -10;
  s.fa_filter_delay_blocks = 0;
}

// ---- AecState::HandleEchoPathChange (aec_state.cc:152-188)
WAP_DEV void aec_state_handle_echo_path_change(Aec3State& a, AecScratch& sc, const EchoPathVariability& v) {
  Aec3Scalars& s = sc.s;
  __syncwarp();
  if (v.delay_change != kDelayAdjNone) {
    if (lane_id() == 0) {
      filter_analyzer_reset(sc);
      s.capture_signal_saturation = 0;
      s.strong_not_saturated_render_blocks = 0;
      s.blocks_with_active_render = 0;
      s.init_state = 1;  // InitialState::Reset
      s.init_strong_blocks = 0;
      // LegacyTransparentModeImpl::Reset
      s.tm_non_converged_sequence_size = 10000;
      s.tm_diverged_sequence_size = 0;
      s.tm_strong_not_saturated_render_blocks = 0;
      // echo_removal_control.linear_and_stable_echo_path (transparent_mode.cc:142-149)
      if (WAP_EC3(linear_and_stable_echo_path)) s.tm_recent_convergence = 0;
      s.erl_blocks_since_reset = 0;  // ErlEstimator::Reset
      // FilteringQualityAnalyzer::Reset
      s.fq_usable = 0;
      s.fq_blocks_since_reset = 0;
    }
    erle_reset(a, sc, true);
  } else if (v.gain_change) {
    erle_reset(a, sc, false);
  }
  if (lane_id() == 0) s.soa_filter_converged = 0;
  __syncwarp();
}

// ---- FilterAnalyzer::Update for the single capture channel.
// Outputs (lane 0, staged scalars): fa_consistent_estimate, fa_gain,
// fa_filter_delay_blocks.  Uses r.v0 / r.x_aligned as scratch.
WAP_DEV void filter_analyzer_update(Aec3State& a, AecScratch& sc) {
  const int lane = lane_id();
  Aec3Scalars& s = sc.s;
  AecRemoverScratch& r = sc.rm;
  __syncwarp();
  const int size = s.h_time_size * kBlock;
  if (lane == 0) {
    ++s.fa_blocks_since_reset;
    // SetRegionToAnalyze
    s.fa_region_start = s.fa_region_end >= size - 1 ? 0 : s.fa_region_end + 1;
    s.fa_region_end = imin(s.fa_region_start + kBlock - 1, size - 1);
  }
  __syncwarp();
  const int start = s.fa_region_start, end = s.fa_region_end;
  // PreProcessFilters: h_highpass_.resize() exposes zeros; 3-tap high-pass over the region.
  for (int i = s.fa_hp_size + lane; i < size; i += 32) a.h_highpass[i] = 0.f;
  for (int k = start + lane; k <= end; k += 32) {
    float tmp = 0.f;
    if (k >= 2) {
      tmp += a.h_time[k] * 0.7929742f;
      tmp += a.h_time[k - 1] * -0.36072128f;
      tmp += a.h_time[k - 2] * -0.47047766f;
    }
    a.h_highpass[k] = tmp;
    r.v0[k - start] = tmp;
  }
  __syncwarp();
  // FindPeakIndex over the region (first maximum that beats the current peak).
  const int peak_in = imin(s.fa_peak_index, size - 1);
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
  const float hp_in = a.h_highpass[peak_in];
  const int peak = (best > hp_in * hp_in) ? bi : peak_in;
  const int delay_blocks = peak >> 6;
  const float h_peak = a.h_highpass[peak];
  __syncwarp();
  if (lane == 0) {
    s.fa_hp_size = size;
    s.fa_peak_index = peak;
    s.fa_filter_delay_blocks = delay_blocks;
    // UpdateFilterGain
    const bool sufficient_time_to_converge = s.fa_blocks_since_reset > 5 * kNumBlocksPerSecond;
    if (sufficient_time_to_converge && s.fa_consistent_estimate) {
      s.fa_gain = fabsf(h_peak);
    } else if (s.fa_gain) {
      s.fa_gain = fmaxr(s.fa_gain, fabsf(h_peak));
    }
    if (WAP_EC3(bounded_erl) && s.fa_gain) s.fa_gain = fmaxr(s.fa_gain, 0.01f);   // filter_analyzer.cc:156-158
    s.fa_filter_length_blocks = (int)((float)size * (1.f / kBlock));
    // ConsistentFilterDetector::Detect, first part
    if (start == 0) {
      s.cfd_floor_accum = 0.f;
      s.cfd_secondary_peak = 0.f;
      s.cfd_floor_low_limit = peak < 64 ? 0 : peak - 64;
      s.cfd_floor_high_limit = peak > size - 129 ? 0 : peak + 128;
    }
    float accum = s.cfd_floor_accum, secondary = s.cfd_secondary_peak;
    for (int k = start; k < imin(end + 1, s.cfd_floor_low_limit); ++k) {
      const float abs_h = fabsf(r.v0[k - start]);
      accum += abs_h;
      secondary = fmaxr(secondary, abs_h);
    }
    for (int k = imax(s.cfd_floor_high_limit, start); k <= end; ++k) {
      const float abs_h = fabsf(r.v0[k - start]);
      accum += abs_h;
      secondary = fmaxr(secondary, abs_h);
    }
    s.cfd_floor_accum = accum;
    s.cfd_secondary_peak = secondary;
    if (end == size - 1) {
      const float filter_floor = accum / (float)(s.cfd_floor_low_limit + size - s.cfd_floor_high_limit);
      const float abs_peak = fabsf(h_peak);
      s.cfd_significant_peak = abs_peak > 10.f * filter_floor && abs_peak > 2.f * secondary;
    }
  }
  __syncwarp();
  if (s.cfd_significant_peak) {
    const float* xb = a.blocks[ring_off(s.blocks_read, -delay_blocks, kRingBlocks)];
    #pragma unroll
    for (int i = lane; i < kBlock; i += 32) r.x_aligned[i] = xb[i];
    __syncwarp();
    const float x_energy = energy_serial(r.x_aligned, kBlock);
    __syncwarp();
    if (lane == 0) {
      const bool active_render_block = x_energy > kActiveRenderEnergy;
      if (s.cfd_consistent_delay_reference == delay_blocks) {
        if (active_render_block) ++s.cfd_consistent_counter;
      } else {
        s.cfd_consistent_counter = 0;
        s.cfd_consistent_delay_reference = delay_blocks;
      }
    }
  }
  if (lane == 0) s.fa_consistent_estimate = (float)s.cfd_consistent_counter > 1.5f * kNumBlocksPerSecond;
  __syncwarp();
}

// ---- EchoAudibility / StationarityEstimator (echo_audibility.cc:37-119, stationarity_estimator.cc:41-241):
// only with echo_audibility.use_stationarity_properties.  `delay_blocks`: MinDirectPathFilterDelay();
// the render noise estimator walks the spectra written since the previous capture block, the
// stationarity flags sum a window of 13 spectra around the delay (plus the render reverb).
WAP_DEV bool sta_band_stationary(const Aec3State& a, int k) { return a.sta_flags[k] != 0 && a.sta_hangovers[k] == 0; }

WAP_DEV void echo_audibility_update(Aec3State& a, AecScratch& sc, int delay_blocks, bool external_delay_seen) {
  const int lane = lane_id();
  Aec3Scalars& s = sc.s;
  __syncwarp();
  // UpdateRenderNoiseEstimator
  if (!s.ea_has_write_prev) {
    __syncwarp();
    if (lane == 0) {
      s.ea_has_write_prev = 1;
      s.ea_spectrum_write_prev = s.spectra_write;
      s.ea_block_write_prev = s.blocks_write;
    }
    __syncwarp();
  } else {
    const int write_current = s.spectra_write;
    if (!s.ea_non_zero_render_seen && !external_delay_seen) {
      // IsRenderTooLow: every block written since the last look must reach 10 in magnitude
      bool too_low = false;
      const int block_write_current = s.blocks_write;
      if (block_write_current == s.ea_block_write_prev) {
        too_low = true;
      } else {
        for (int idx = s.ea_block_write_prev; idx != block_write_current; idx = ring_inc(idx, kRingBlocks)) {
          float m = 0.f;
          for (int i = lane; i < kBlock; i += 32) m = fmaxf(m, fabsf(a.blocks[idx][i]));
          m = warp_max(m);
          if (m < 10.f) { too_low = true; break; }
        }
      }
      __syncwarp();
      if (lane == 0) {
        s.ea_block_write_prev = block_write_current;
        s.ea_non_zero_render_seen = too_low ? 0 : 1;
      }
      __syncwarp();
    }
    if (s.ea_non_zero_render_seen) {
      for (int idx = s.ea_spectrum_write_prev; idx != write_current; idx = ring_dec(idx, kRingBlocks)) {
        // NoiseSpectrum::Update
        const int block_counter = s.sta_block_counter + 1;
        constexpr float kAlpha = 0.004f, kAlphaInit = 0.04f;
        constexpr float kTiltAlpha = (kAlphaInit - kAlpha) / 500;
        const float alpha = block_counter > 520 ? kAlpha : kAlphaInit - kTiltAlpha * (block_counter - 20);
        #pragma unroll
        for (int k = lane; k < kBins; k += 32) {
          const float power_band = a.spectra[idx][k];
          float noise = a.sta_noise[k];
          if (block_counter <= 20) {
            noise += (1.f / 20) * power_band;
          } else if (noise < power_band) {
            float alpha_inc = alpha * (noise / power_band);
            if (block_counter > 500 && 10.f * noise < power_band) alpha_inc *= 0.1f;
            noise += alpha_inc * (power_band - noise);
          } else {
            noise += alpha * (power_band - noise);
            noise = fmaxr(noise, 10.f);
          }
          a.sta_noise[k] = noise;
        }
        __syncwarp();
        if (lane == 0) s.sta_block_counter = block_counter;
        __syncwarp();
      }
    }
    __syncwarp();
    if (lane == 0) s.ea_spectrum_write_prev = write_current;
    __syncwarp();
  }
  if (!(external_delay_seen || WAP_EC3(use_stationarity_properties_at_init))) return;
  // UpdateRenderStationarityFlags -> StationarityEstimator::UpdateStationarityFlags
  const int idx_at_delay = ring_off(s.spectra_read, delay_blocks, kRingBlocks);
  const int headroom = s.spectra_write < s.spectra_read ? s.spectra_read - s.spectra_write
                                                        : kRingBlocks - s.spectra_write + s.spectra_read;
  const int num_lookahead_bounded = imin(imax(0, headroom - delay_blocks + 1), 13 - 1);
  int idx0 = idx_at_delay;
  if (num_lookahead_bounded < 13 - 1) idx0 = ring_off(idx_at_delay, (13 - 1) - num_lookahead_bounded, kRingBlocks);
  #pragma unroll
  for (int k = lane; k < kBins; k += 32) {
    float acum_power = 0.f;
    int idx = idx0;
    for (int j = 0; j < 13; ++j) {
      acum_power += a.spectra[idx][k] * 1.f;   // one render channel
      idx = ring_dec(idx, kRingBlocks);
    }
    acum_power += a.avg_render_reverb[k];
    const float noise = 13 * a.sta_noise[k];
    a.sta_flags[k] = acum_power < 10.f * noise ? 1 : 0;
  }
  __syncwarp();
  // UpdateHangover, SmoothStationaryPerFreq
  int all_l = 1;
  #pragma unroll
  for (int k = lane; k < kBins; k += 32) all_l &= a.sta_flags[k];
  const bool reduce_hangover = __all_sync(WAP_FULL, all_l);
  int smooth[3] = {0, 0, 0};
  #pragma unroll
  for (int k = lane, j = 0; k < kBins; k += 32, ++j) {
    const int kk = k == 0 ? 1 : (k == 64 ? 63 : k);   // bins 0 and 64 copy their neighbours' result
    smooth[j] = a.sta_flags[kk - 1] && a.sta_flags[kk] && a.sta_flags[kk + 1];
    int h = a.sta_hangovers[k];
    if (!a.sta_flags[k]) h = 12;   // kHangoverBlocks = kNumBlocksPerSecond / 20
    else if (reduce_hangover) h = imax(h - 1, 0);
    a.sta_hangovers[k] = h;
  }
  __syncwarp();
  #pragma unroll
  for (int k = lane, j = 0; k < kBins; k += 32, ++j) a.sta_flags[k] = smooth[j];
  __syncwarp();
}

// ---- AecState::Update (aec_state.cc:190-342).  Inputs in sc.rm: Y2, E2 (spectrum of the
// formed linear output), sc.red[0..6] subtractor metrics.  `ext_has/ext_delay` is
// the block processor's estimated delay.
WAP_DEV void aec_state_update(Aec3State& a, AecScratch& sc, int ext_has, int ext_delay) {
  const int lane = lane_id();
  Aec3Scalars& s = sc.s;
  AecRemoverScratch& r = sc.rm;
  __syncwarp();
  const float y2 = sc.red[0], e2_refined = sc.red[1], e2_coarse = sc.red[2];
  const float s_refined_max_abs = sc.red[5], s_coarse_max_abs = sc.red[6];

  // SubtractorOutputAnalyzer::Update
  constexpr float kConvergenceThreshold = 50 * 50 * kBlock;
  constexpr float kConvergenceThresholdLowLevel = 20 * 20 * kBlock;
  const bool refined_filter_converged = e2_refined < 0.5f * y2 && y2 > kConvergenceThreshold;
  const bool coarse_filter_converged_strict = e2_coarse < 0.05f * y2 && y2 > kConvergenceThreshold;
  const bool any_coarse_filter_converged = e2_coarse < 0.3f * y2 && y2 > kConvergenceThresholdLowLevel;
  const float min_e2 = fminr(e2_refined, e2_coarse);
  const bool all_filters_diverged = min_e2 > 1.5f * y2 && y2 > 30.f * 30.f * kBlock;
  const bool any_filter_converged = refined_filter_converged || coarse_filter_converged_strict;
  __syncwarp();
  if (lane == 0) s.soa_filter_converged = any_filter_converged;

  filter_analyzer_update(a, sc);
  const bool any_filter_consistent = s.fa_consistent_estimate != 0;
  const float max_echo_path_gain = s.fa_gain;

  // FilterDelay::Update
  if (lane == 0) {
    if (ext_has && (!s.fd_has_external || s.fd_external_delay != ext_delay)) {
      s.fd_has_external = 1;
      s.fd_external_delay = ext_delay;
    }
    if (WAP_EC3(use_linear_filter)) {   // aec_state.cc:220-223
      const bool may_not_have_converged = s.strong_not_saturated_render_blocks < 2 * kNumBlocksPerSecond;
      if (may_not_have_converged && s.fd_has_external) s.fd_filter_delay = WAP_EC3(delay_headroom_samples) / kBlock;
      else s.fd_filter_delay = s.fa_filter_delay_blocks;
      s.fd_min_filter_delay = s.fd_filter_delay;
    }
  }
  __syncwarp();
  const int delay = s.fd_min_filter_delay;

  // aligned render block, render counters
  {
    const float* xb = a.blocks[ring_off(s.blocks_read, -delay, kRingBlocks)];
    #pragma unroll
    for (int i = lane; i < kBlock; i += 32) r.x_aligned[i] = xb[i];
  }
  __syncwarp();
  const float render_energy = energy_serial(r.x_aligned, kBlock);
  const bool active_render = render_energy > kActiveRenderEnergy;
  const bool saturated_capture = s.capture_signal_saturation != 0;
  const bool usable_linear_before = s.fq_usable != 0;
  float max_sample_l = 0.f;
  #pragma unroll
  for (int i = lane; i < kBlock; i += 32) max_sample_l = fmaxf(max_sample_l, fabsf(r.x_aligned[i]));
  const float max_sample = warp_max(max_sample_l);
  __syncwarp();
  if (lane == 0) {
    s.blocks_with_active_render += active_render ? 1 : 0;
    s.strong_not_saturated_render_blocks += (active_render && !saturated_capture) ? 1 : 0;
  }

  // ComputeAvgRenderReverb: r.v1 = avg_render_spectrum_with_reverb, r.v2 = X2 at the delay.
  const int idx_at_delay = ring_off(s.spectra_read, delay, kRingBlocks);
  const int idx_past = ring_inc(idx_at_delay, kRingBlocks);
  #pragma unroll
  for (int k = lane; k < kBins; k += 32) {
    const float rev = (a.avg_render_reverb[k] + a.spectra[idx_past][k] * 1.0f) * aec_reverb_decay(sc, false);
    a.avg_render_reverb[k] = rev;
    const float x2 = a.spectra[idx_at_delay][k];
    r.v2[k] = x2;
    r.v1[k] = x2 + rev;
  }
  __syncwarp();
  if (WAP_EC3(use_stationarity_properties)) echo_audibility_update(a, sc, delay, s.fd_has_external != 0);
  if (s.init_transition_triggered) erle_reset(a, sc, false);

  // Four 65-term chains side by side: X2_reverb, Y2, E2, X2.
  if (lane < 4) {
    const float* p = lane == 0 ? r.v1 : lane == 1 ? r.Y2 : lane == 2 ? r.E2 : r.v2;
    sc.red[16 + lane] = chain_sum(p, 0, kBins);
  }
  __syncwarp();
  const float X2rev_sum = sc.red[16], Y2_sum = sc.red[17], E2_sum = sc.red[18], X2_sum = sc.red[19];
  const bool converged = any_filter_converged;

  // ---- ErleEstimator::Update
  if (lane == 0) sc.ired[8] = (++s.erle_blocks_since_reset < 2 * kNumBlocksPerSecond) ? 0 : 1;
  __syncwarp();
  if (sc.ired[8]) {
    // SubbandErleEstimator::Update
    const bool restart = converged && s.erle_num_points == 6;
    const int num_points = converged ? (restart ? 1 : s.erle_num_points + 1) : s.erle_num_points;
    const bool update_bands = converged && num_points == 6;
    #pragma unroll
    for (int k = lane; k < kBins; k += 32) {
      float accY = a.accum_Y2[k], accE = a.accum_E2[k];
      int low = a.accum_low_render[k];
      if (converged) {  // UpdateAccumulatedSpectra
        if (restart) { accY = 0.f; accE = 0.f; low = 0; }
        accY = r.Y2[k] + accY;
        accE = r.E2[k] + accE;
        low = low || r.v1[k] < kX2BandEnergyThreshold;
        a.accum_Y2[k] = accY;
        a.accum_E2[k] = accE;
        a.accum_low_render[k] = low;
      }
      if (k >= 1 && k < 64) {
        float erle = a.erle[k], erle_oc = a.erle_onset_comp[k], erle_u = a.erle_unbounded[k];
        int hold = a.erle_hold_counters[k], onset = a.coming_onset[k];
        if (update_bands && accE > 0.f) {  // UpdateBands
          const float new_erle = accY / accE;
          if (WAP_EC3(erle_onset_detection) && !low) {
            if (onset) onset = 0;
            hold = 250;  // kBlocksForOnsetDetection
          }
          const float max_erle = k < 32 ? WAP_EC3(erle_max_l) : WAP_EC3(erle_max_h);
          float alpha = 0.05f;
          if (new_erle < erle) alpha = low ? 0.f : 0.1f;
          erle = clampr(erle + alpha * (new_erle - erle), WAP_EC3(erle_min), max_erle);
          if (WAP_EC3(erle_onset_detection)) {
            alpha = 0.05f;
            if (new_erle < erle_oc) alpha = low ? 0.f : 0.1f;
            erle_oc = clampr(erle_oc + alpha * (new_erle - erle_oc), WAP_EC3(erle_min), max_erle);
          }
          alpha = 0.05f;
          if (new_erle < erle_u) alpha = low ? 0.f : 0.1f;
          erle_u = clampr(erle_u + alpha * (new_erle - erle_u), WAP_EC3(erle_min), 100000.0f);
        }
        // DecreaseErlePerBandForLowRenderSignals (erle_during_onsets_ stays at min_erle)
        if (WAP_EC3(erle_onset_detection)) {
          --hold;
          if (hold <= 250 - 100) {
            if (erle_oc > WAP_EC3(erle_min)) erle_oc = fmaxr(WAP_EC3(erle_min), 0.97f * erle_oc);
            if (hold <= 0) { onset = 1; hold = 0; }
          }
        }
        a.erle[k] = erle;
        a.erle_onset_comp[k] = erle_oc;
        a.erle_unbounded[k] = erle_u;
        a.erle_hold_counters[k] = hold;
        a.coming_onset[k] = onset;
        if (k == 1) { a.erle[0] = erle; a.erle_onset_comp[0] = erle_oc; a.erle_unbounded[0] = erle_u; }
        if (k == 63) { a.erle[64] = erle; a.erle_onset_comp[64] = erle_oc; a.erle_unbounded[64] = erle_u; }
      }
    }
    __syncwarp();
    if (WAP_EC3(erle_num_sections) > 1) signal_dependent_erle_update(a, sc, converged);
    if (lane == 0) {
      s.erle_num_points = num_points;
      // FullBandErleEstimator::Update
      if (converged && X2rev_sum > kX2BandEnergyThreshold * (float)kBins) {
        // ErleInstantaneous::Update
        bool update_estimates = false;
        s.fb_E2_acum += E2_sum;
        s.fb_Y2_acum += Y2_sum;
        if (++s.fb_num_points == 6) {
          if (s.fb_E2_acum > 0.f) {
            update_estimates = true;
            s.fb_erle_log2 = fast_approx_log2f(s.fb_Y2_acum / s.fb_E2_acum + 1e-3f);
            s.fb_has_erle_log2 = 1;
          }
          s.fb_num_points = 0;
          s.fb_E2_acum = 0.f;
          s.fb_Y2_acum = 0.f;
        }
        if (update_estimates) {
          s.fb_max_erle_log2 -= 0.0004f;
          s.fb_max_erle_log2 = fmaxr(s.fb_max_erle_log2, s.fb_erle_log2);
          s.fb_min_erle_log2 += 0.0004f;
          s.fb_min_erle_log2 = fminr(s.fb_min_erle_log2, s.fb_erle_log2);
          float quality_estimate = 0.f;
          if (s.fb_max_erle_log2 > s.fb_min_erle_log2)
            quality_estimate = (s.fb_erle_log2 - s.fb_min_erle_log2) / (s.fb_max_erle_log2 - s.fb_min_erle_log2);
          if (quality_estimate > s.fb_inst_quality) s.fb_inst_quality = quality_estimate;
          else s.fb_inst_quality += 0.07f * (quality_estimate - s.fb_inst_quality);
          s.fb_hold_counter = 100;  // kBlocksToHoldErle
          s.fb_erle_time_domain_log2 += 0.05f * (s.fb_erle_log2 - s.fb_erle_time_domain_log2);
          s.fb_erle_time_domain_log2 = fmaxr(s.fb_erle_time_domain_log2, fast_approx_log2f(WAP_EC3(erle_min) + 1e-3f));
        }
      }
      --s.fb_hold_counter;
      if (s.fb_hold_counter == 0) {  // ResetAccumulators
        s.fb_has_erle_log2 = 0;
        s.fb_inst_quality = 0.f;
        s.fb_num_points = 0;
        s.fb_E2_acum = 0.f;
        s.fb_Y2_acum = 0.f;
      }
    }
    __syncwarp();
  }

  // ---- ErlEstimator::Update (render spectrum at the delay, capture spectrum)
  if (lane == 0) sc.ired[9] = (++s.erl_blocks_since_reset < 2 * kNumBlocksPerSecond || !converged) ? 0 : 1;
  __syncwarp();
  if (sc.ired[9]) {
    for (int k = 1 + lane; k < 64; k += 32) {
      float erl = a.erl[k];
      int hold = a.erl_hold_counters[k - 1];
      const float X2 = r.v2[k];
      if (X2 > kX2BandEnergyThreshold) {
        const float new_erl = r.Y2[k] / X2;
        if (new_erl < erl) {
          hold = 1000;
          erl += 0.1f * (new_erl - erl);
          erl = fmaxr(erl, 0.01f);
        }
      }
      --hold;
      erl = hold > 0 ? erl : fminr(1000.f, 2.f * erl);
      a.erl[k] = erl;
      a.erl_hold_counters[k - 1] = hold;
      if (k == 1) a.erl[0] = erl;
      if (k == 63) a.erl[64] = erl;
    }
    if (lane == 0) {
      if (X2_sum > kX2BandEnergyThreshold * (float)kBins) {
        const float new_erl = Y2_sum / X2_sum;
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

  // ---- scalar state machines, lane 0
  if (lane == 0) {
    // SaturationDetector::Update
    if (WAP_EC3(echo_can_saturate)) s.saturated_echo = 0;   // aec_state.cc:273-280: else the detector is never updated
    if (WAP_EC3(echo_can_saturate) && saturated_capture) {
      if (usable_linear_before) {
        s.saturated_echo = s_refined_max_abs > 20000.f || s_coarse_max_abs > 20000.f;
      } else {
        const float peak_echo_amplitude = max_sample * max_echo_path_gain * 10.f;
        s.saturated_echo = peak_echo_amplitude > 32000;
      }
    }
    // InitialState::Update
    s.init_strong_blocks += (active_render && !saturated_capture) ? 1 : 0;
    const int prev_initial_state = s.init_state;
    if (WAP_EC3(conservative_initial_phase)) s.init_state = s.init_strong_blocks < 5 * kNumBlocksPerSecond;   // aec_state.cc:360-366
    else s.init_state = (float)s.init_strong_blocks < WAP_EC3(initial_state_seconds) * kNumBlocksPerSecond;
    s.init_transition_triggered = !s.init_state && prev_initial_state;
    // LegacyTransparentModeImpl::Update
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
    // ep_strength.bounded_erl: no TransparentMode object at all (transparent_mode.cc:239-243): never transparent
    if (WAP_EC3(bounded_erl)) s.tm_active = 0;
    // FilteringQualityAnalyzer::Update
    const bool filter_update = active_render && !saturated_capture;
    s.fq_blocks_since_reset += filter_update ? 1 : 0;
    s.fq_blocks_since_start += filter_update ? 1 : 0;
    s.fq_convergence_seen = s.fq_convergence_seen || any_filter_converged;
    const bool sufficient_at_startup = (float)s.fq_blocks_since_start > kNumBlocksPerSecond * 0.4f;
    const bool sufficient_at_reset = sufficient_at_startup && (float)s.fq_blocks_since_reset > kNumBlocksPerSecond * 0.2f;
    bool usable = sufficient_at_startup && sufficient_at_reset;
    usable = usable && (ext_has || s.fq_convergence_seen);
    usable = usable && !s.tm_active;
    s.fq_usable = usable && WAP_EC3(use_linear_filter);   // UsableLinearEstimate() (aec_state.h) / :457-461
    s.fq_usable_filter = usable;                          // UsableLinearFilterOutputs()
  }
  __syncwarp();

  // ---- ReverbModelEstimator::Update -> ReverbFrequencyResponse::Update (not on a stationary render block)
  bool stationary_block = false;
  if (WAP_EC3(use_stationarity_properties)) {
    int cnt = 0;
    #pragma unroll
    for (int k = lane; k < kBins; k += 32) cnt += sta_band_stationary(a, k) ? 1 : 0;
    for (int m = 16; m; m >>= 1) cnt += __shfl_xor_sync(WAP_FULL, cnt, m);
    stationary_block = ((float)cnt * (1.f / kBins)) > 0.75f;   // IsBlockStationary
  }
  float quality = s.fb_inst_quality;   // ErleInstantaneous::GetQualityEstimate
  if (WAP_EC3(clamp_quality_estimate_to_zero)) quality = fmaxr(0.f, quality);
  if (WAP_EC3(clamp_quality_estimate_to_one)) quality = fminr(1.f, quality);
  const bool has_quality = s.fb_has_erle_log2 != 0;
  if (has_quality && !stationary_block) {
    const float* tail = a.H2[s.H2_size - 1];
    const float* direct = a.H2[s.fd_filter_delay];
    if (lane < 2) sc.red[16 + lane] = chain_sum(lane == 0 ? direct : tail, 1, kBins);
    __syncwarp();
    const float direct_path_energy = sc.red[16], tail_energy = sc.red[17];
    const float average_decay = direct_path_energy == 0.f ? 0.f : tail_energy / direct_path_energy;
    const float smoothing = 0.2f * quality;
    const float avg = s.reverb_average_decay + smoothing * (average_decay - s.reverb_average_decay);
    #pragma unroll
    for (int k = lane; k < kBins; k += 32)
      r.v3[k] = WAP_EC3(use_conservative_tail_frequency_response) ? fmaxr(tail[k], direct[k] * avg) : direct[k] * avg;
    __syncwarp();
    if (lane == 0) {
      s.reverb_average_decay = avg;
      for (int k = 1; k < 64; ++k) {  // in-place recursion, ascending k
        const float avg_neighbour = 0.5f * (r.v3[k - 1] + r.v3[k + 1]);
        r.v3[k] = fmaxr(r.v3[k], avg_neighbour);
      }
    }
    __syncwarp();
    #pragma unroll
    for (int k = lane; k < kBins; k += 32) a.tail_response[k] = r.v3[k];
  }
  __syncwarp();
  // ---- ReverbModelEstimator::Update -> ReverbDecayEstimator::Update; without the adaptive decay it only
  // maintains state that nothing reads
  if (WAP_EC3(default_len) < 0.f && !stationary_block) reverb_decay_update(a, sc, has_quality, quality);
}

// ---- ComfortNoiseGenerator::Compute.  `nearend` = capture spectrum chosen by the caller.
// Output: r.N_re / r.N_im (lower-band comfort noise).
// hi: also keep what the upper-band comfort noise needs -- r.v1 = sqrt spectrum, r.v3 / r.v2 = the
// per-bin sqrt(2)*sin / sqrt(2)*cos factors, sc.red[24] = high_band_noise_level.
WAP_DEV void cng_compute(Aec3State& a, const EngineConfig& cfg, AecScratch& sc, const float* nearend, bool hi = false) {
  const int lane = lane_id();
  Aec3Scalars& s = sc.s;
  AecRemoverScratch& r = sc.rm;
  const float noise_floor = cfg.cng_noise_floor;  // GetNoiseFloorFactor(), computed on the host
  __syncwarp();
  const bool saturated_capture = s.capture_signal_saturation != 0;
  const int counter = s.cng_N2_counter;
  const bool has_initial = s.cng_has_initial != 0;
  const bool drop_initial = !saturated_capture && has_initial && counter + 1 == 1000;
  const bool use_initial = has_initial && !drop_initial;
  #pragma unroll
  for (int k = lane; k < kBins; k += 32) {
    float N2 = a.cng_N2[k], N2i = a.cng_N2_initial[k];
    if (!saturated_capture) {
      float Y2s = a.cng_Y2_smoothed[k];
      Y2s = Y2s + 0.1f * (nearend[k] - Y2s);
      a.cng_Y2_smoothed[k] = Y2s;
      if (counter > 50) N2 = Y2s < N2 ? (0.9f * Y2s + 0.1f * N2) * 1.0002f : N2 * 1.0002f;
      if (use_initial) N2i = N2 > N2i ? N2i + 0.001f * (N2 - N2i) : N2;
      N2 = fmaxr(N2, noise_floor);
      if (use_initial) {
        N2i = fmaxr(N2i, noise_floor);
        a.cng_N2_initial[k] = N2i;
      }
      a.cng_N2[k] = N2;
    }
    // GenerateComfortNoise
    const float N = sqrtf(use_initial ? N2i : N2);
    float re = 0.f, im = 0.f, fx = 0.f, fy = 0.f;
    if (k >= 1 && k < 64) {
      const unsigned seed_k = (kLcgA[k] * s.cng_seed + kLcgC[k]) & 0x7fffffffu;
      const int i = (int)(seed_k >> 26);
      fx = kSqrt2Sin[i];
      fy = kSqrt2Sin[(i + 8) & 31];
      re = N * fx;
      im = N * fy;
    }
    r.N_re[k] = re;
    r.N_im[k] = im;
    if (hi) { r.v1[k] = N; r.v3[k] = fx; r.v2[k] = fy; }
  }
  __syncwarp();
  if (hi) {
    // high_band_noise_level = accumulate(N[32..64]) / 33 (comfort_noise_generator.cc:73-78)
    constexpr float kOneByNumBands = 1.f / (kBins / 2 + 1);
    float acc = 0.f;
    for (int k = kBins / 2; k < kBins; ++k) acc += r.v1[k];
    if (lane == 0) sc.red[24] = acc * kOneByNumBands;
  }
  if (lane == 0) {
    if (!saturated_capture && has_initial) {
      ++s.cng_N2_counter;
      if (drop_initial) s.cng_has_initial = 0;
    }
    s.cng_seed = (kLcgA[63] * s.cng_seed + kLcgC[63]) & 0x7fffffffu;
  }
  __syncwarp();
}

// ---- ResidualEchoEstimator::Estimate.  Outputs r.R2, r.R2_unb.
WAP_DEV void residual_echo_estimate(Aec3State& a, AecScratch& sc) {
  const int lane = lane_id();
  Aec3Scalars& s = sc.s;
  AecRemoverScratch& r = sc.rm;
  __syncwarp();
  const bool dominant_nearend = s.dn_nearend_state != 0;
  const bool usable = s.fq_usable != 0;
  const bool saturated_echo = s.saturated_echo != 0;
  const bool transparent = s.tm_active != 0;
  const float* X2_latest = a.spectra[s.spectra_read];
  const float echo_path_gain = transparent ? 0.01f * 0.01f : WAP_EC3(default_gain) * WAP_EC3(default_gain);
  const int delay = s.fd_min_filter_delay;
  const bool add_reverb = usable || (WAP_EC3(model_reverb_in_nonlinear_mode) && !transparent);
  // AecState::ReverbDecay(mild = dominant_nearend) (residual_echo_estimator.cc:384, aec_state.h:127)
  const float reverb_decay = aec_reverb_decay(sc, dominant_nearend);
  const int first_reverb_partition = usable ? s.fa_filter_length_blocks + 1 : delay + 1;
  const float* X2_reverb_src = a.spectra[ring_off(s.spectra_read, first_reverb_partition, kRingBlocks)];
  // LinearEstimate uses Erle(onset_compensated), which only differs from erle_ with onset detection
  // (residual_echo_estimator.cc:249-251, subband_erle_estimator.h:46-50)
  const bool onset_compensated = WAP_EC3(erle_onset_compensation_in_dominant_nearend) || !dominant_nearend;
  // ErleEstimator::Erle / ErleUnbounded (erle_estimator.h:58-74): the signal-dependent estimator's when it exists
  const bool sd_erle = WAP_EC3(erle_num_sections) > 1;
  const float* erle = (onset_compensated && WAP_EC3(erle_onset_detection)) ? (sd_erle ? a.sd_erle_onset : a.erle_onset_comp)
                                                                            : (sd_erle ? a.sd_erle : a.erle);
  const float* erle_unbounded = sd_erle ? a.sd_erle : a.erle_unbounded;
  // GetRenderIndexesToAnalyze (residual_echo_estimator.cc:70-86): echo_model.render_pre / _post_window_size
  const int w_first = imax(0, delay - WAP_EC3(render_pre_window_size));
  const int w0 = ring_off(s.spectra_read, w_first, kRingBlocks);
  const int wn = delay + WAP_EC3(render_post_window_size) - w_first + 1;  // spectra in the window
  #pragma unroll
  for (int k = lane; k < kBins; k += 32) {
    // UpdateRenderNoisePower
    float floor = a.X2_noise_floor[k];
    {
      const float p = X2_latest[k];
      int cnt = a.X2_noise_floor_counter[k];
      if (p < floor) {
        floor = p;
        cnt = 0;
      } else if (cnt >= (int)WAP_EC3(noise_floor_hold)) {
        floor = fmaxr(floor * 1.1f, WAP_EC3(min_noise_floor_power));
      } else {
        ++cnt;
      }
      a.X2_noise_floor[k] = floor;
      a.X2_noise_floor_counter[k] = cnt;
    }
    float R2, R2u;
    if (usable) {
      if (saturated_echo) {
        R2 = R2u = r.Y2[k];
      } else {
        R2 = r.S2_lin[k] / erle[k];
        R2u = r.S2_lin[k] / erle_unbounded[k];
      }
    } else if (saturated_echo) {
      R2 = R2u = r.Y2[k];
    } else {
      // EchoGeneratingPower + ApplyNoiseGate + stationary-noise subtraction
      float X2 = 0.f;
      int idx = w0;
      for (int j = 0; j < wn; ++j) {
        X2 = fmaxr(X2, a.spectra[idx][k]);
        idx = ring_inc(idx, kRingBlocks);
      }
      if (!WAP_EC3(use_stationarity_properties) && WAP_EC3(noise_gate_power) > X2)
        X2 = fmaxr(0.f, X2 - WAP_EC3(noise_gate_slope) * (WAP_EC3(noise_gate_power) - X2));
      X2 -= WAP_EC3(stationary_gate_slope) * floor;
      X2 = fmaxr(0.f, X2);
      R2 = R2u = X2 * echo_path_gain;
    }
    if (add_reverb) {
      // UpdateReverb + AddReverb
      const float scaling = usable ? a.tail_response[k] : echo_path_gain;
      const float rev = (a.echo_reverb[k] + X2_reverb_src[k] * scaling) * reverb_decay;
      a.echo_reverb[k] = rev;
      R2 += rev;
      R2u += rev;
    }
    if (WAP_EC3(use_stationarity_properties)) {
      // AecState::GetResidualEchoScaling (aec_state.cc:115-126, echo_audibility.h:40-51)
      const float converge_blocks = (WAP_EC3(conservative_initial_phase) ? 1.5f : 0.8f) * kNumBlocksPerSecond;
      const bool filter_has_had_time_to_converge = (float)s.strong_not_saturated_render_blocks >= converge_blocks;
      const float scaling = (sta_band_stationary(a, k) &&
                             (filter_has_had_time_to_converge || WAP_EC3(use_stationarity_properties_at_init))) ? 0.f : 1.0f;
      R2 *= scaling;
      R2u *= scaling;
    }
    r.R2[k] = R2;
    r.R2_unb[k] = R2u;
  }
  __syncwarp();
}

// GainParameters thresholds for bin k (suppression_gain.cc:452-477).
WAP_DEV void gain_params(const AecScratch& sc, const Ec3Tuning& t, int k, float* enr_transparent, float* enr_suppress, float* emr_transparent) {
  float aa;
  if (k <= WAP_EC3(last_lf_band)) aa = 0.f;
  else if (k < WAP_EC3(first_hf_band)) aa = fdiv((float)(k - WAP_EC3(last_lf_band)), (float)(WAP_EC3(first_hf_band) - WAP_EC3(last_lf_band)));
  else aa = 1.f;
  *enr_transparent = (1 - aa) * t.lf_t + aa * t.hf_t;
  *enr_suppress = (1 - aa) * t.lf_s + aa * t.hf_s;
  *emr_transparent = (1 - aa) * t.lf_e + aa * t.hf_e;
}

// ---- SuppressionGain::GetGain for one band.  `nearend` / `echo_spectrum` chosen
// by the caller; output r.gain (amplitude domain).  sc.x = render block 0.
WAP_DEV void suppression_gain_get_gain(Aec3State& a, AecScratch& sc, const float* nearend, bool clock_drift) {
  const int lane = lane_id();
  Aec3Scalars& s = sc.s;
  AecRemoverScratch& r = sc.rm;
  __syncwarp();
  // DominantNearendDetector::Update on (nearend, R2_unbounded, N2): three 15-term
  // chains; LowNoiseRenderDetector::Detect: one 64-term chain with running maximum.
  const bool subband_detector = WAP_EC3(use_subband_nearend_detection) != 0;
  if (subband_detector) {
    // SubbandNearendDetector::Update (subband_nearend_detector.cc:37-75): its own MovingAverage of the
    // nearend spectrum (moving_average.cc:36-59; r.gain is free until the gains are formed), then the
    // band powers below as left-to-right chains.
    const int n_mem = WAP_EC3(snd_average_blocks) - 1;
    const float scaling = fdiv(1.f, (float)WAP_EC3(snd_average_blocks));
    const int snd_index = s.snd_mem_index;
    #pragma unroll
    for (int k = lane; k < kBins; k += 32) {
      const float in = nearend[k];
      float o = in;
      for (int j = 0; j < n_mem; ++j) o = a.snd_mem[j][k] + o;
      r.gain[k] = o * scaling;
      if (n_mem > 0) a.snd_mem[snd_index][k] = in;
    }
    __syncwarp();
  }
  if (lane < 3) {
    if (subband_detector) {
      // noise power of region 1, nearend power of region 1, nearend power of region 2
      const float* p = lane == 0 ? a.cng_N2 : r.gain;
      const int lo = lane == 2 ? WAP_EC3(snd_sub2_low) : WAP_EC3(snd_sub1_low);
      const int hi = lane == 2 ? WAP_EC3(snd_sub2_high) : WAP_EC3(snd_sub1_high);
      sc.red[16 + lane] = chain_sum(p, lo, hi + 1);
    } else {
      const float* p = lane == 0 ? nearend : lane == 1 ? (WAP_EC3(dn_use_unbounded_echo_spectrum) ? r.R2_unb : r.R2) : a.cng_N2;
      sc.red[16 + lane] = chain_sum(p, 1, 16);
    }
  } else if (lane == 3) {
    float x2_sum = 0.f, x2_max = 0.f;
    for (int i = 0; i < kBlock; ++i) {
      const float x2 = sc.x[i] * sc.x[i];
      x2_sum += x2;
      x2_max = fmaxr(x2_max, x2);
    }
    sc.red[19] = x2_sum;
    sc.red[20] = x2_max;
  }
  __syncwarp();
  if (lane == 0 && subband_detector) {
    const float one_over1 = fdiv(1.f, (float)(WAP_EC3(snd_sub1_high) - WAP_EC3(snd_sub1_low) + 1));
    const float one_over2 = fdiv(1.f, (float)(WAP_EC3(snd_sub2_high) - WAP_EC3(snd_sub2_low) + 1));
    const float noise_power = sc.red[16] * one_over1;
    const float nearend_power_subband1 = sc.red[17] * one_over1, nearend_power_subband2 = sc.red[18] * one_over2;
    s.dn_nearend_state = (nearend_power_subband1 < WAP_EC3(snd_nearend_threshold) * nearend_power_subband2 &&
                          nearend_power_subband1 > WAP_EC3(snd_snr_threshold) * noise_power) ? 1 : 0;
    const int snd_mem_len = imax(1, WAP_EC3(snd_average_blocks) - 1);
    if (WAP_EC3(snd_average_blocks) > 1) s.snd_mem_index = (s.snd_mem_index + 1) % snd_mem_len;
  }
  if (lane == 0) {
    const float ne_sum = sc.red[16], echo_sum = sc.red[17], noise_sum = sc.red[18];
    if (!subband_detector) {   // (with the subband detector the dominant-nearend detector does not exist)
      if ((!s.sg_initial_state || WAP_EC3(dn_use_during_initial_phase)) &&
          echo_sum < WAP_EC3(dn_enr_threshold) * ne_sum && ne_sum > WAP_EC3(dn_snr_threshold) * noise_sum) {
        if (++s.dn_trigger_counter >= WAP_EC3(dn_trigger_threshold)) {
          s.dn_hold_counter = WAP_EC3(dn_hold_duration);
          s.dn_trigger_counter = WAP_EC3(dn_trigger_threshold);
        }
      } else {
        s.dn_trigger_counter = imax(0, s.dn_trigger_counter - 1);
      }
      if (echo_sum > WAP_EC3(dn_enr_exit_threshold) * ne_sum && echo_sum > WAP_EC3(dn_snr_threshold) * noise_sum) s.dn_hold_counter = 0;
      s.dn_hold_counter = imax(0, s.dn_hold_counter - 1);
      s.dn_nearend_state = s.dn_hold_counter > 0;
    }
    // LowNoiseRenderDetector
    const float x2_sum = sc.red[19] / 1, x2_max = sc.red[20];
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
  const int mem_index = s.sg_nearend_mem_index;
  // LowerBandGain
  #pragma unroll
  for (int k = lane; k < kBins; k += 32) {
    const float last_gain = a.last_gain[k];
    const float max_gain = fminr(fmaxr(last_gain * tun.max_inc, WAP_EC3(floor_first_increase)), 1.f);
    // MovingAverage::Average (mem_len 4 -> 3 stored blocks, scaling 1/4)
    const float in = nearend[k];
    float ne = in;
    constexpr bool kFourBlocks = !WAP_EC3_RUNTIME;   // the default: mem_len 4 -> 3 stored blocks, scaling 1/4
    if (kFourBlocks || WAP_EC3(nearend_average_blocks) == 4) {
      ne = a.nearend_mem[0][k] + ne;
      ne = a.nearend_mem[1][k] + ne;
      ne = a.nearend_mem[2][k] + ne;
      ne *= 0.25f;
      a.nearend_mem[mem_index][k] = in;
    } else {
      const int n_mem = WAP_EC3(nearend_average_blocks) - 1;   // 0..2 stored blocks
      for (int j = 0; j < n_mem; ++j) ne = a.nearend_mem[j][k] + ne;
      ne *= fdiv(1.f, (float)WAP_EC3(nearend_average_blocks));
      if (n_mem > 0) a.nearend_mem[mem_index][k] = in;
    }
    // WeightEchoForAudibility: bins [0,3) / [3,7) / [7,65) (suppression_gain.cc:104-118)
    const float audibility = k < 3 ? WAP_EC3(audibility_threshold_lf) : (k < 7 ? WAP_EC3(audibility_threshold_mf) : WAP_EC3(audibility_threshold_hf));
    const float threshold = WAP_EC3(floor_power) * audibility;
    const float normalizer = 1.f / (threshold - WAP_EC3(floor_power));
    const float echo = r.R2[k];
    float weighted = echo;
    if (echo < threshold) {
      const float tmp = (threshold - echo) * normalizer;
      weighted = echo * fmaxr(0.f, 1.f - tmp * tmp);
    }
    // GetMinGain
    float min_gain = 0.f;
    if (!saturated_echo) {
      min_gain = weighted > 0.f ? min_echo_power / weighted : 1.f;
      min_gain = fminr(min_gain, 1.f);
      if (k <= WAP_EC3(last_lf_smoothing_band) && (!s.sg_initial_state || WAP_EC3(lf_smoothing_during_initial_phase))) {
        if (a.last_nearend[k] > a.last_echo[k] || k <= WAP_EC3(last_permanent_lf_smoothing_band)) {
          min_gain = fmaxr(min_gain, last_gain * tun.max_dec_lf);
          min_gain = fminr(min_gain, 1.f);
        }
      }
    }
    // GainToNoAudibleEcho (masker = comfort noise spectrum N2)
    float enr_t, enr_s, emr_t;
    gain_params(sc, tun, k, &enr_t, &enr_s, &emr_t);
    const float enr = weighted / (ne + 1.f);
    const float emr = weighted / (a.cng_N2[k] + 1.f);
    float g = 1.0f;
    if (enr > enr_t && emr > emr_t) {
      g = (enr_s - enr) / (enr_s - enr_t);
      g = fmaxr(g, emr_t / emr);
    }
    g = fmaxr(fminr(g, max_gain), min_gain);
    r.gain[k] = fminr(1.f, g);
    a.last_nearend[k] = ne;
    a.last_echo[k] = weighted;
  }
  __syncwarp();
  {
    // LimitLowFrequencyGains / LimitHighFrequencyGains
    const float g12 = fminr(r.gain[1], r.gain[2]);
    const bool limit_hf = !nearend_state || clock_drift || WAP_EC3(conservative_hf_suppression);
    // min over bands [limiting_gain_band, + bands_in_limiting_gain) (suppression_gain.cc:44-62)
    float min_upper_gain = 1.f;
    for (int band = WAP_EC3(limiting_gain_band); band < WAP_EC3(limiting_gain_band) + WAP_EC3(bands_in_limiting_gain); ++band)
      min_upper_gain = fminr(min_upper_gain, r.gain[band]);
    const bool limit_bands = WAP_EC3(bands_in_limiting_gain) > 0;
    const float g63 = limit_hf ? (limit_bands ? fminr(r.gain[63], min_upper_gain) : r.gain[63]) : r.gain[63];
    __syncwarp();
    float hf_gain_bound = 0.f;
    const bool hf_bound = limit_hf && WAP_EC3(conservative_hf_suppression);
    if (hf_bound) {
      // LimitHighFrequencyGains, conservative part (suppression_gain.cc:66-83): the mean of the limited
      // gains of bins 20..28 (a left-to-right sum) bounds every bin from 29 up
      #pragma unroll
      for (int k = lane; k < kBins; k += 32) {
        float g = r.gain[k];
        if (limit_bands && k > WAP_EC3(limiting_gain_band)) g = fminr(g, min_upper_gain);
        if (k == 64) g = g63;
        if (k >= 20 && k < 29) r.v0[k] = g;
      }
      __syncwarp();
      constexpr float kOneByBandsInSum = 1 / static_cast<float>(29 - 20);
      hf_gain_bound = chain_sum(r.v0, 20, 29) * kOneByBandsInSum;
      __syncwarp();
    }
    #pragma unroll
    for (int k = lane; k < kBins; k += 32) {
      float g = r.gain[k];
      if (k <= 1) g = g12;
      if (limit_hf) {
        if (limit_bands && k > WAP_EC3(limiting_gain_band)) g = fminr(g, min_upper_gain);
        if (k == 64) g = g63;
        if (hf_bound && k >= 29) g = fminr(g, hf_gain_bound);
      }
      a.last_gain[k] = g;
      r.gain[k] = sqrtf(g);
    }
  }
  if (lane == 0) {
    const int n_mem = imax(1, WAP_EC3(nearend_average_blocks) - 1);
    s.sg_nearend_mem_index = WAP_EC3(nearend_average_blocks) > 1 ? (mem_index + 1) % n_mem : 0;
  }
  __syncwarp();
}

// ---- SuppressionGain::UpperBandsGain (suppression_gain.cc:124-217) for a 3-band leg.
// sc.x = band 0 of the render block GetBlock(0), r.gain = lower-band gain (amplitude domain).
WAP_DEV float upper_bands_gain(const Aec3State& a, AecScratch& sc, const UpperBandState& up, const float* echo_spectrum) {
  const int lane = lane_id();
  const Aec3Scalars& s = sc.s;
  AecRemoverScratch& r = sc.rm;
  __syncwarp();
  if (s.rsa_has_narrow_peak && s.rsa_narrow_peak_band > kBins - 10) return 0.001f;
  // gain_below_8_khz = min over low_band_gain[32..64]
  float g = r.gain[32 + lane];
  if (lane == 0) g = fminr(g, r.gain[64]);
  for (int m = 16; m; m >>= 1) g = fminf(g, __shfl_xor_sync(WAP_FULL, g, m));
  const float gain_below_8_khz = g;
  if (s.saturated_echo) return fminr(0.001f, gain_below_8_khz);
  // band energies: three serial sums of squares (std::accumulate), one per lane
  if (lane < 3) {
    const float* p = lane == 0 ? sc.x : up.blocks_hi[s.blocks_read][lane - 1];
    float acc = 0.f;
    for (int i = 0; i < kBlock; ++i) acc = acc + p[i] * p[i];
    sc.red[16 + lane] = acc;
  }
  __syncwarp();
  const float low_band_energy = sc.red[16];
  const float high_band_energy = fmaxr(fmaxr(0.f, sc.red[17]), sc.red[18]);
  float anti_howling_gain;
  const float activation_threshold = kBlock * WAP_EC3(hb_anti_howling_activation_threshold);
  if (high_band_energy < fmaxr(low_band_energy, activation_threshold)) anti_howling_gain = 1.f;
  else anti_howling_gain = WAP_EC3(hb_anti_howling_gain) * sqrtf(low_band_energy / high_band_energy);
  // Bound the upper gain during significant echo activity (suppression_gain.cc:190-204).
  float gain_bound = 1.f;
  if (!s.dn_nearend_state && WAP_EC3(hb_max_gain_during_echo) != 1.f) {   // (at the default bound of 1 the test is moot)
    if (lane < 2) sc.red[19 + lane] = chain_sum(lane == 0 ? echo_spectrum : a.cng_N2, 1, 16);
    __syncwarp();
    if (sc.red[19] > WAP_EC3(hb_enr_threshold) * sc.red[20]) gain_bound = WAP_EC3(hb_max_gain_during_echo);
  }
  __syncwarp();
  return fminr(fminr(gain_below_8_khz, anti_howling_gain), gain_bound);
}

// ---- EchoRemoverImpl::ProcessCapture.  sc.y = capture block (in/out).
// `up` / `b`: 48 kHz legs only -- bands 1-2 of capture block b (up->capture_blocks_hi[b]) are
// processed too and returned in sc.x (band 1) and sc.rm.x_aligned (band 2).
WAP_DEV void echo_remover_process_capture(Aec3State& a, const EngineConfig& cfg, AecScratch& sc, EchoPathVariability v,
                                          bool capture_signal_saturation, int ext_has, int ext_delay,
                                          UpperBandState* up = nullptr, int b = 0) {
  const int lane = lane_id();
  Aec3Scalars& s = sc.s;
  AecRemoverScratch& r = sc.rm;
  __syncwarp();
  // The estimator vectors and time-domain memories the stages after the linear filter touch
  // (H_error .. output_framer, contiguous in the slab): on their way while the subtractor runs.
  warp_prefetch_l1(a.H_error, (int)(reinterpret_cast<const char*>(a.render_decimator) - reinterpret_cast<const char*>(a.H_error)));
  // x = render_buffer->GetBlock(0)
  {
    const float* xb = a.blocks[s.blocks_read];
    #pragma unroll
    for (int i = lane; i < kBlock; i += 32) sc.x[i] = xb[i];
  }
  if (lane == 0) s.capture_signal_saturation = capture_signal_saturation;
  if (v.delay_change != kDelayAdjNone || v.gain_change) {
    if (v.gain_change) {  // act on a gain change only once per frame
      const bool act = s.er_gain_change_hangover == 0;
      __syncwarp();
      if (act) {
        if (lane == 0) s.er_gain_change_hangover = 3;
      } else {
        v.gain_change = 0;
      }
    }
    __syncwarp();
    subtractor_handle_echo_path_change(a, sc, v);
    aec_state_handle_echo_path_change(a, sc, v);
    if (v.delay_change != kDelayAdjNone && lane == 0) s.sg_initial_state = 1;
  }
  __syncwarp();
  if (lane == 0) {
    if (s.er_gain_change_hangover > 0) --s.er_gain_change_hangover;
  }
  __syncwarp();

  render_signal_analyzer_update(a, sc, s.fd_min_filter_delay, up ? up->blocks_hi[s.blocks_read][0] : nullptr);

  if (s.init_transition_triggered) {
    if (lane == 0) {
      subtractor_exit_initial_state(sc);
      s.sg_initial_state = 0;
    }
    __syncwarp();
  }

  subtractor_process(a, sc, s.capture_signal_saturation != 0);
  // The filter passes have streamed ~28 KB through the L1 since the prefetch at the top of the block
  // (which brought the lines into the L2): ask again now that the stages reading these vectors are next.
  // Measured on B200: k_echo -3.3 %.
  warp_prefetch_l1(a.H_error, (int)(reinterpret_cast<const char*>(a.render_decimator) - reinterpret_cast<const char*>(a.H_error)));

  // FormLinearFilterOutput (echo_remover.cc:497-529)
  {
    const float y2 = sc.red[0], e2_refined = sc.red[1], e2_coarse = sc.red[2], s2_refined = sc.red[3], s2_coarse = sc.red[4];
    bool use_refined_output = true;
    if (!WAP_EC3(enable_coarse_filter_output_usage)) {
      // filter.enable_coarse_filter_output_usage = false: always the refined filter's output
    } else if (e2_coarse < 0.9f * e2_refined && y2 > 30.f * 30.f * kBlock &&
        (s2_refined > 60.f * 60.f * kBlock || s2_coarse > 60.f * 60.f * kBlock)) {
      use_refined_output = false;
    } else if (e2_coarse < e2_refined && y2 < e2_refined) {
      use_refined_output = false;
    }
    const float* from = s.er_refined_last_selected ? r.e_ref : r.e_coa;
    const float* to = use_refined_output ? r.e_ref : r.e_coa;
    #pragma unroll
    for (int i = lane; i < kBlock; i += 32) {
      float o = to[i];
      if (from != to && i < 30) {  // SignalTransition
        const float aa = (i + 1) * (1.f / 31);
        o = aa * to[i] + (1.f - aa) * from[i];
      }
      r.e[i] = o;
    }
    __syncwarp();
    if (lane == 0) s.er_refined_last_selected = use_refined_output;
  }
  // WindowedPaddedFft of y and e (sqrt-Hanning), spectra.
  #pragma unroll
  for (int i = lane; i < kBlock; i += 32) {
    sc.fftA[i] = a.y_old[i] * kSqrtHanning128[i];
    sc.fftA[kBlock + i] = sc.y[i] * kSqrtHanning128[kBlock + i];
    sc.fftB[i] = a.e_old[i] * kSqrtHanning128[i];
    sc.fftB[kBlock + i] = r.e[i] * kSqrtHanning128[kBlock + i];
    a.y_old[i] = sc.y[i];
    a.e_old[i] = r.e[i];
  }
  fft_pair(sc, false, true);
  packed_to_reim(sc.fftA, r.Y_re, r.Y_im);
  packed_to_reim(sc.fftB, r.E_re, r.E_im);
  __syncwarp();
  #pragma unroll
  for (int k = lane; k < kBins; k += 32) {
    const float dr = r.Y_re[k] - r.E_re[k], di = r.Y_im[k] - r.E_im[k];
    r.S2_lin[k] = dr * dr + di * di;  // LinearEchoPower
    r.Y2[k] = power_bin(r.Y_re[k], r.Y_im[k], k);
    r.E2[k] = power_bin(r.E_re[k], r.E_im[k], k);
  }
  __syncwarp();
  // nearend_spectrum is bound before AecState::Update changes UsableLinearEstimate().
  const bool nearend_is_E2 = s.fq_usable != 0;
  aec_state_update(a, sc, ext_has, ext_delay);

  const float* nearend = nearend_is_E2 ? r.E2 : r.Y2;
  cng_compute(a, cfg, sc, nearend, up != nullptr);

  if (cfg.capture_output_used) {
    residual_echo_estimate(a, sc);
    const bool usable = s.fq_usable != 0;
    if (usable) {
      #pragma unroll
      for (int k = lane; k < kBins; k += 32) r.E2[k] = fminr(r.E2[k], r.Y2[k]);
      __syncwarp();
    }
    suppression_gain_get_gain(a, sc, nearend, v.clock_drift != 0 || WAP_EC3(has_clock_drift));
    // SuppressionGain::UpperBandsGain (suppression_gain.cc:124-217); echo_spectrum = S2_linear or R2.
    float high_bands_gain = 1.f;
    if (up) high_bands_gain = upper_bands_gain(a, sc, *up, usable ? r.S2_lin : r.R2);

    // SuppressionFilter::ApplyGain
    const float* Yf_re = usable ? r.E_re : r.Y_re;
    const float* Yf_im = usable ? r.E_im : r.Y_im;
    #pragma unroll
    for (int k = lane; k < kBins; k += 32) {
      const float g = r.gain[k];
      const float noise_gain = sqrtf(1.f - g * g);
      const float E_real = Yf_re[k] * g;
      const float E_imag = Yf_im[k] * g;
      const float re = E_real + noise_gain * r.N_re[k];
      const float im = E_imag + noise_gain * r.N_im[k];
      if (k == 0) sc.fftA[0] = re;
      else if (k == 64) sc.fftA[1] = re;
      else { sc.fftA[2 * k] = re; sc.fftA[2 * k + 1] = im; }
      if (up) {  // comfort_noise_high_band (GenerateComfortNoise): level * (x, y), zero at DC / Nyquist
        const float lvl = sc.red[24];
        if (k == 0) sc.fftB[0] = 0.f;
        else if (k == 64) sc.fftB[1] = 0.f;
        else { sc.fftB[2 * k] = lvl * r.v3[k]; sc.fftB[2 * k + 1] = lvl * r.v2[k]; }
      }
    }
    fft_pair(sc, true, up != nullptr);
    constexpr float kIfftNormalization = 2.f / 128;
    #pragma unroll
    for (int i = lane; i < kBlock; i += 32) {
      float e0 = a.e_output_old[i] * kSqrtHanning128[kBlock + i];
      e0 += sc.fftA[i] * kSqrtHanning128[i];
      e0 = e0 * kIfftNormalization;
      a.e_output_old[i] = sc.fftA[kBlock + i];
      sc.y[i] = clampr(e0, -32768.f, 32767.f);
    }
    if (up) {
      // Upper bands (suppression_filter.cc:153-181): gain, comfort noise on band 1, one block of
      // delay (swap with e_output_old_), clamp.
      const float noise_scaling = 0.4f * sqrtf(1.f - high_bands_gain * high_bands_gain);
      const float ngain = noise_scaling * kIfftNormalization;
      __syncwarp();
      #pragma unroll
      for (int i = lane; i < kBlock; i += 32) {
        float e1 = up->capture_blocks_hi[b][0][i] * high_bands_gain;
        e1 += sc.fftB[i] * ngain;
        const float e2 = up->capture_blocks_hi[b][1][i] * high_bands_gain;
        const float o1 = up->e_output_old_hi[0][i], o2 = up->e_output_old_hi[1][i];
        up->e_output_old_hi[0][i] = e1;
        up->e_output_old_hi[1][i] = e2;
        sc.x[i] = clampr(o1, -32768.f, 32767.f);
        r.x_aligned[i] = clampr(o2, -32768.f, 32767.f);
      }
    }
  } else if (up) {
    __syncwarp();
    #pragma unroll
    for (int i = lane; i < kBlock; i += 32) {
      sc.x[i] = up->capture_blocks_hi[b][0][i];
      r.x_aligned[i] = up->capture_blocks_hi[b][1][i];
    }
  }
  __syncwarp();
}

}  // namespace wap
