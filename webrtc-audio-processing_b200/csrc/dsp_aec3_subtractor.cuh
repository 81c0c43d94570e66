// AEC3 linear echo canceller for one call leg (mono render / mono capture):
//   Subtractor::{Process, HandleEchoPathChange, ExitInitialState}   aec3/subtractor.cc:156-375
//   AdaptiveFirFilter                                               aec3/adaptive_fir_filter.cc:480-747
//   ApplyFilter_Avx2 / AdaptPartitions_Avx2 / ComputeFrequencyResponse_Avx2
//                                                                   aec3/adaptive_fir_filter_avx2.cc:30-194
//   RefinedFilterUpdateGain / CoarseFilterUpdateGain                aec3/refined_filter_update_gain.cc:70-174,
//                                                                   aec3/coarse_filter_update_gain.cc:39-105
//   RenderSignalAnalyzer                                            aec3/render_signal_analyzer.cc:33-159
// Lanes own frequency bins (k = lane, lane+32, 64); every per-bin recursion
// runs over the partitions in the reference's order, so no cross-lane
// reduction is needed for the filters.
#pragma once

#include "dsp_aec3_common.cuh"

namespace wap {

WAP_DEVCONST float kRefinedCfg[5] = WAP_EC3_REFINED;
WAP_DEVCONST float kRefinedInitialCfg[5] = WAP_EC3_REFINED_INITIAL;
WAP_DEVCONST float kCoarseCfg[2] = WAP_EC3_COARSE;
WAP_DEVCONST float kCoarseInitialCfg[2] = WAP_EC3_COARSE_INITIAL;

// ZeroFilter (adaptive_fir_filter.cc:464-476)
WAP_DEV void fir_zero_partitions(float (*H_re)[kBinsPad], float (*H_im)[kBinsPad], int from, int to) {
  for (int p = from; p < to; ++p)
    for (int k = lane_id(); k < kBinsPad; k += 32) { H_re[p][k] = 0.f; H_im[p][k] = 0.f; }
}

// AdaptiveFirFilter::SetSizePartitions(size, immediate_effect = true) scalars; returns the old size.
WAP_DEV int fir_set_size_immediate(int* cur, int* target, int* old_target, int* counter, int* ptc, int size) {
  *target = imin(kMaxPartitions, size);
  const int old = *cur;
  *cur = *old_target = *target;
  *ptc = imin(*ptc, *cur - 1);
  *counter = 0;
  return old;
}

// AdaptiveFirFilter::UpdateSize scalars (:542-565); returns the old size.
WAP_DEV int fir_update_size(int* cur, int* target, int* old_target, int* counter, int* ptc) {
  const int old = *cur;
  if (*counter > 0) {
    --*counter;
    const float one_by = 1.f / ec3::kConfigChangeDuration;
    const float change_factor = *counter * one_by;
    const float v = (float)*old_target * change_factor + (float)*target * (1.f - change_factor);
    *cur = (int)v;
    *ptc = imin(*ptc, *cur - 1);
  } else {
    *cur = *old_target = *target;
  }
  return old;
}

// Subtractor::HandleEchoPathChange (subtractor.cc:156-183)
WAP_DEV void subtractor_handle_echo_path_change(Aec3State& a, AecScratch& sc, const EchoPathVariability& v) {
  const int lane = lane_id();
  Aec3Scalars& s = sc.s;
  if (v.delay_change != kDelayAdjNone) {
    __syncwarp();
    fir_zero_partitions(a.Hr_re, a.Hr_im, s.fr_current_size, kMaxPartitions);
    fir_zero_partitions(a.Hc_re, a.Hc_im, s.fc_current_size, kMaxPartitions);
    for (int k = lane; k < kBins; k += 32) a.H_error[k] = 10000.f;
    __syncwarp();
    if (lane == 0) {
      if (!v.gain_change) {
        s.rg_poor_excitation_counter = 1000;
        s.rg_call_counter = 0;
      }
      s.cg_poor_excitation_counter = 0;
      s.cg_call_counter = 0;
      for (int i = 0; i < 5; ++i) s.rg_cur[i] = s.rg_old[i] = s.rg_tgt[i] = kRefinedInitialCfg[i];
      s.rg_config_change_counter = 0;
      for (int i = 0; i < 2; ++i) s.cg_cur[i] = s.cg_old[i] = s.cg_tgt[i] = kCoarseInitialCfg[i];
      s.cg_config_change_counter = 0;
      sc.ired[0] = fir_set_size_immediate(&s.fr_current_size, &s.fr_target_size, &s.fr_old_target_size,
                                          &s.fr_size_change_counter, &s.fr_partition_to_constrain, kInitPartitions);
      sc.ired[1] = fir_set_size_immediate(&s.fc_current_size, &s.fc_target_size, &s.fc_old_target_size,
                                          &s.fc_size_change_counter, &s.fc_partition_to_constrain, kInitPartitions);
    }
    __syncwarp();
    fir_zero_partitions(a.Hr_re, a.Hr_im, sc.ired[0], s.fr_current_size);
    fir_zero_partitions(a.Hc_re, a.Hc_im, sc.ired[1], s.fc_current_size);
    __syncwarp();
  }
  // gain_change alone only re-runs RefinedFilterUpdateGain::HandleEchoPathChange,
  // whose counters are untouched in that case (refined_filter_update_gain.cc:52-68).
}

// Subtractor::ExitInitialState (subtractor.cc:185-194), lane 0.
WAP_DEV void subtractor_exit_initial_state(Aec3Scalars& s) {
  for (int i = 0; i < 5; ++i) { s.rg_old[i] = s.rg_cur[i]; s.rg_tgt[i] = kRefinedCfg[i]; }
  s.rg_config_change_counter = ec3::kConfigChangeDuration;
  for (int i = 0; i < 2; ++i) { s.cg_old[i] = s.cg_cur[i]; s.cg_tgt[i] = kCoarseCfg[i]; }
  s.cg_config_change_counter = ec3::kConfigChangeDuration;
  s.fr_target_size = imin(kMaxPartitions, kMaxPartitions);
  s.fr_size_change_counter = ec3::kConfigChangeDuration;
  s.fc_target_size = imin(kMaxPartitions, kMaxPartitions);
  s.fc_size_change_counter = ec3::kConfigChangeDuration;
}

// UpdateCurrentConfig of both gain classes, lane 0.
WAP_DEV void gain_update_current_config(float* cur, float* old, const float* tgt, int n, int* counter) {
  if (*counter > 0) {
    if (--*counter > 0) {
      const float one_by = 1.f / ec3::kConfigChangeDuration;
      const float change_factor = *counter * one_by;
      for (int i = 0; i < n; ++i) cur[i] = old[i] * change_factor + tgt[i] * (1.f - change_factor);
    } else {
      for (int i = 0; i < n; ++i) cur[i] = old[i] = tgt[i];
    }
  }
}

// RenderSignalAnalyzer::Update (render_signal_analyzer.cc:131-141) for the
// spectrum at `delay_partitions` and the latest render block in sc.x.
WAP_DEV void render_signal_analyzer_update(Aec3State& a, AecScratch& sc, int delay_partitions) {
  const int lane = lane_id();
  Aec3Scalars& s = sc.s;
  __syncwarp();
  // IdentifySmallNarrowBandRegions
  {
    const float* X2 = a.spectra[ring_off(s.spectra_read, delay_partitions, kRingBlocks)];
    for (int k = 1 + lane; k < 64; k += 32) {
      const bool narrow = X2[k] > 3 * fmaxr(X2[k - 1], X2[k + 1]);
      a.narrow_band_counters[k - 1] = narrow ? a.narrow_band_counters[k - 1] + 1 : 0;
    }
  }
  // IdentifyStrongNarrowBandComponent
  const float* X2_latest = a.spectra[s.spectra_read];
  const int peak_bin = warp_argmax_first(X2_latest, kBins);
  float max_abs_l = 0.f;
  for (int i = lane; i < kBlock; i += 32) max_abs_l = fmaxf(max_abs_l, fabsf(sc.x[i]));
  const float max_abs = warp_max(max_abs_l);
  if (lane == 0) {
    if (s.rsa_has_narrow_peak && ++s.rsa_narrow_peak_counter > kMaxPartitions) s.rsa_has_narrow_peak = 0;
    float non_peak_power = 0.f;
    for (int k = imax(0, peak_bin - 14); k < peak_bin - 4; ++k) non_peak_power = fmaxr(X2_latest[k], non_peak_power);
    for (int k = peak_bin + 5; k < imin(peak_bin + 15, kBins); ++k) non_peak_power = fmaxr(X2_latest[k], non_peak_power);
    const float peak_level = X2_latest[peak_bin];
    if (peak_bin > 0 && max_abs > 100 && peak_level > 100 * non_peak_power) {
      if (peak_level > 0.f) {  // max_peak_level starts at 0 (single render channel)
        s.rsa_has_narrow_peak = 1;
        s.rsa_narrow_peak_band = peak_bin;
        s.rsa_narrow_peak_counter = 0;
      }
    }
  }
  __syncwarp();
}

// Per-bin mask of RenderSignalAnalyzer::MaskRegionsAroundNarrowBands (:143-159):
// mask[k] = 1 when v[k] is zeroed.  Also returns PoorSignalExcitation().
WAP_DEV bool render_signal_analyzer_mask(const Aec3State& a, float* mask) {
  const int lane = lane_id();
  const int* c = a.narrow_band_counters;
  int poor = 0;
  for (int k = lane; k < kBins; k += 32) {
    bool m = false;
    if (k <= 1 && c[0] > 5) m = true;
    if (k >= 63 && c[62] > 5) m = true;
    for (int kk = imax(2, k - 2); kk <= imin(62, k + 2); ++kk)
      if (c[kk - 1] > 5) m = true;
    mask[k] = m ? 1.f : 0.f;
    if (k < 63 && c[k] > 10) poor = 1;
  }
  const bool any_poor = __any_sync(WAP_FULL, poor);
  __syncwarp();
  return any_poor;
}

// One adaptive filter: S = sum_p X_p * H_p (ApplyFilter_Avx2), then the
// time-domain prediction error (PredictionError, subtractor.cc:49-65).
WAP_DEV void fir_filter_and_error(const Aec3State& a, AecScratch& sc, const float (*H_re)[kBinsPad],
                                  const float (*H_im)[kBinsPad], int num_partitions, float* e_out, float* s_out) {
  const int lane = lane_id();
  const int pos = sc.s.spectra_read;
  __syncwarp();
  for (int k = lane; k < kBins; k += 32) {
    float S_re = 0.f, S_im = 0.f;
    int xp = pos;
    for (int p = 0; p < num_partitions; ++p) {
      const float X_re = a.fft_re[xp][k], X_im = a.fft_im[xp][k];
      const float Hre = H_re[p][k], Him = H_im[p][k];
      const float aa = X_re * Hre, bb = X_im * Him, cc = X_re * Him, dd = X_im * Hre;
      S_re = S_re + (aa - bb);
      S_im = S_im + (cc + dd);
      xp = ring_inc(xp, kRingBlocks);
    }
    if (k == 0) sc.fftA[0] = S_re;
    else if (k == 64) sc.fftA[1] = S_re;
    else { sc.fftA[2 * k] = S_re; sc.fftA[2 * k + 1] = S_im; }
  }
  fft_pair(sc, true, false);
  constexpr float kScale = 1.0f / 64;
  for (int i = lane; i < kBlock; i += 32) {
    const float t = sc.fftA[kBlock + i];
    e_out[i] = sc.y[i] - t * kScale;
    s_out[i] = kScale * t;
  }
  __syncwarp();
}

// AdaptPartitions_Avx2: H_p += conj(X_p) * G.
WAP_DEV void fir_adapt_partitions(const Aec3State& a, AecScratch& sc, float (*H_re)[kBinsPad], float (*H_im)[kBinsPad],
                                  int num_partitions, const float* G_re, const float* G_im) {
  const int pos = sc.s.spectra_read;
  for (int k = lane_id(); k < kBins; k += 32) {
    const float Gre = G_re[k], Gim = G_im[k];
    int xp = pos;
    for (int p = 0; p < num_partitions; ++p) {
      const float X_re = a.fft_re[xp][k], X_im = a.fft_im[xp][k];
      const float aa = X_re * Gre, bb = X_im * Gim, cc = X_re * Gim, dd = X_im * Gre;
      H_re[p][k] = H_re[p][k] + (aa + bb);
      H_im[p][k] = H_im[p][k] + (cc - dd);
      xp = ring_inc(xp, kRingBlocks);
    }
  }
  __syncwarp();
}

// AdaptiveFirFilter::Constrain / ConstrainAndUpdateImpulseResponse (:645-706)
// for partition p; `impulse` (may be null) receives the 64 retained taps.
WAP_DEV void fir_constrain(AecScratch& sc, float* H_re_p, float* H_im_p, float* impulse) {
  const int lane = lane_id();
  __syncwarp();
  reim_to_packed(H_re_p, H_im_p, sc.fftA);
  fft_pair(sc, true, false);
  constexpr float kScale = 1.0f / 64;
  for (int i = lane; i < kBlock; i += 32) {
    const float v = sc.fftA[i] * kScale;
    sc.fftA[i] = v;
    sc.fftA[kBlock + i] = 0.f;
    if (impulse) impulse[i] = v;
  }
  fft_pair(sc, false, false);
  packed_to_reim(sc.fftA, H_re_p, H_im_p);
  __syncwarp();
}

// Aec3Fft::ZeroPaddedFft(x, kHanning) into the packed buffer `buf` (not transformed yet).
WAP_DEV void stage_zero_padded_hanning(const float* x, float* buf) {
  for (int i = lane_id(); i < kBlock; i += 32) {
    buf[i] = 0.f;
    buf[kBlock + i] = x[i] * kHanning64[i];
  }
}

// Subtractor::Process (subtractor.cc:196-343).  Inputs: capture block sc.y,
// render rings; outputs in sc.rm (e_ref, e_coa, s_ref, s_coa, E2_ref, E2_coa,
// Er) and sc.red[0..6] = {y2, e2_refined, e2_coarse, s2_refined, s2_coarse,
// s_refined_max_abs, s_coarse_max_abs}.
WAP_DEV void subtractor_process(Aec3State& a, AecScratch& sc, bool saturated_capture) {
  const int lane = lane_id();
  Aec3Scalars& s = sc.s;
  AecRemoverScratch& r = sc.rm;
  __syncwarp();
  const int P_r = s.fr_current_size, P_c = s.fc_current_size;
  // RenderBuffer::SpectralSum(s) (render_buffer.cc:42-83): one running sum per bin.
  {
    const int pmax = imax(P_r, P_c);
    for (int k = lane; k < kBins; k += 32) {
      float x2 = 0.f;
      int pos = s.spectra_read;
      for (int j = 0; j < pmax; ++j) {
        x2 += a.spectra[pos][k];
        pos = ring_inc(pos, kRingBlocks);
        if (j + 1 == P_r) r.X2_ref[k] = x2;
        if (j + 1 == P_c) r.X2_coa[k] = x2;
      }
    }
  }
  fir_filter_and_error(a, sc, a.Hr_re, a.Hr_im, P_r, r.e_ref, r.s_ref);
  fir_filter_and_error(a, sc, a.Hc_re, a.Hc_im, P_c, r.e_coa, r.s_coa);

  // SubtractorOutput::ComputeMetrics (subtractor_output.cc:39-59): five serial
  // sums of squares, one per lane, plus the two peak magnitudes.
  if (lane < 5) {
    const float* p = lane == 0 ? sc.y : lane == 1 ? r.e_ref : lane == 2 ? r.e_coa : lane == 3 ? r.s_ref : r.s_coa;
    float acc = 0.f;
    for (int i = 0; i < kBlock; ++i) acc = acc + p[i] * p[i];
    sc.red[lane] = acc;
  } else if (lane < 7) {
    const float* p = lane == 5 ? r.s_ref : r.s_coa;
    float mx = p[0], mn = p[0];
    for (int i = 1; i < kBlock; ++i) { mx = fmaxr(mx, p[i]); mn = fminr(mn, p[i]); }
    sc.red[lane] = fmaxr(mx, -mn);
  }
  __syncwarp();
  const float y2 = sc.red[0], e2_refined = sc.red[1], e2_coarse = sc.red[2];

  // FilterMisadjustmentEstimator (subtractor.cc:345-375)
  if (lane == 0) {
    s.mis_e2_acum += e2_refined;
    s.mis_y2_acum += y2;
    if (++s.mis_n_blocks_acum == 4) {
      if (s.mis_y2_acum > 4 * 200.f * 200.f * kBlock) {
        const float update = s.mis_e2_acum / s.mis_y2_acum;
        if (s.mis_e2_acum > 4 * 7500.f * 7500.f * kBlock) s.mis_overhang = 4;
        else s.mis_overhang = imax(s.mis_overhang - 1, 0);
        if ((update < s.mis_inv_misadjustment) || (s.mis_overhang > 0))
          s.mis_inv_misadjustment += 0.1f * (update - s.mis_inv_misadjustment);
      }
      s.mis_e2_acum = 0.f;
      s.mis_y2_acum = 0.f;
      s.mis_n_blocks_acum = 0;
    }
    sc.ired[0] = s.mis_inv_misadjustment > 10.f;
    sc.red[8] = sc.ired[0] ? 2.f / sqrtf(s.mis_inv_misadjustment) : 1.f;
    if (sc.ired[0]) {
      s.mis_e2_acum = 0.f; s.mis_y2_acum = 0.f; s.mis_n_blocks_acum = 0;
      s.mis_inv_misadjustment = 0.f; s.mis_overhang = 0;
    }
  }
  __syncwarp();
  const bool refined_filters_adjusted = sc.ired[0] != 0;
  if (refined_filters_adjusted) {
    const float scale = sc.red[8];
    for (int i = lane; i < kMaxPartitions * kBinsPad; i += 32) {
      (&a.Hr_re[0][0])[i] *= scale;
      (&a.Hr_im[0][0])[i] *= scale;
    }
    for (int i = lane; i < s.h_time_size * kBlock; i += 32) a.h_time[i] *= scale;
    for (int i = lane; i < kBlock; i += 32) {  // ScaleFilterOutput
      r.s_ref[i] *= scale;
      r.e_ref[i] = sc.y[i] - r.s_ref[i];
    }
    __syncwarp();
  }

  // FFTs of the two windowed, zero-padded errors; spectra.
  stage_zero_padded_hanning(r.e_ref, sc.fftA);
  stage_zero_padded_hanning(r.e_coa, sc.fftB);
  fft_pair(sc, false, true);
  packed_to_reim(sc.fftA, r.Er_re, r.Er_im);
  packed_to_reim(sc.fftB, r.Ec_re, r.Ec_im);
  __syncwarp();
  power_spectrum(r.Ec_re, r.Ec_im, r.E2_coa);
  power_spectrum(r.Er_re, r.Er_im, r.E2_ref);
  const bool poor_excitation = render_signal_analyzer_mask(a, r.v0);  // r.v0 = narrow-band mask

  // ---- refined filter update
  if (lane == 0) {
    sc.ired[1] = 1;  // G == 0 ?
    if (!refined_filters_adjusted) {
      ++s.rg_call_counter;
      gain_update_current_config(s.rg_cur, s.rg_old, s.rg_tgt, 5, &s.rg_config_change_counter);
      if (poor_excitation) s.rg_poor_excitation_counter = 0;
      const bool zero = (unsigned)(++s.rg_poor_excitation_counter) < (unsigned)P_r || saturated_capture ||
                        (unsigned)s.rg_call_counter <= (unsigned)P_r;
      sc.ired[1] = zero;
    }
    sc.ired[2] = s.coarse_filter_reset_hangover > 0;  // disallow_leakage_diverged
  }
  __syncwarp();
  if (!refined_filters_adjusted) {
    const bool zero_gain = sc.ired[1] != 0;
    const bool disallow_leakage_diverged = sc.ired[2] != 0;
    const float leak_conv = s.rg_cur[0], leak_div = s.rg_cur[1], err_floor = s.rg_cur[2], err_ceil = s.rg_cur[3],
                noise_gate = s.rg_cur[4];
    const int H2_size = s.H2_size;
    for (int k = lane; k < kBins; k += 32) {
      // ComputeErl (adaptive_fir_filter_erl_avx2.cc:27-40)
      float erl = 0.f;
      for (int j = 0; j < H2_size; ++j) erl += a.H2[j][k];
      float H_error = a.H_error[k];
      const float X2 = r.X2_ref[k], E2r = r.E2_ref[k];
      if (zero_gain) {
        r.G_re[k] = 0.f;
        r.G_im[k] = 0.f;
      } else {
        float mu = 0.f;
        if (X2 >= noise_gate) mu = H_error / (0.5f * H_error * X2 + (float)P_r * E2r);
        if (r.v0[k] != 0.f) mu = 0.f;
        H_error -= 0.5f * mu * X2 * H_error;
        r.G_re[k] = mu * r.Er_re[k];
        r.G_im[k] = mu * r.Er_im[k];
      }
      if (E2r <= r.E2_coa[k] || disallow_leakage_diverged) H_error += leak_conv * erl;
      else H_error += leak_div * erl;
      H_error = fmaxr(H_error, err_floor);
      H_error = fminr(H_error, err_ceil);
      a.H_error[k] = H_error;
    }
  } else {
    for (int k = lane; k < kBins; k += 32) { r.G_re[k] = 0.f; r.G_im[k] = 0.f; }
  }
  __syncwarp();
  // AdaptiveFirFilter::Adapt(render_buffer, G, &impulse_response) for the refined filter.
  if (lane == 0) {
    sc.ired[3] = fir_update_size(&s.fr_current_size, &s.fr_target_size, &s.fr_old_target_size,
                                 &s.fr_size_change_counter, &s.fr_partition_to_constrain);
  }
  __syncwarp();
  {
    const int P = s.fr_current_size;
    fir_zero_partitions(a.Hr_re, a.Hr_im, sc.ired[3], P);
    __syncwarp();
    fir_adapt_partitions(a, sc, a.Hr_re, a.Hr_im, P, r.G_re, r.G_im);
    // impulse_response->resize(): newly exposed taps are zero.
    for (int i = s.h_time_size * kBlock + lane; i < P * kBlock; i += 32) a.h_time[i] = 0.f;
    const int p = s.fr_partition_to_constrain;
    fir_constrain(sc, a.Hr_re[p], a.Hr_im[p], a.h_time + p * kBlock);
    if (lane == 0) {
      s.h_time_size = P;
      s.fr_partition_to_constrain = p < (P - 1) ? p + 1 : 0;
      s.H2_size = P;
    }
    // ComputeFrequencyResponse_Avx2 (single render channel: max with 0).
    for (int pp = 0; pp < P; ++pp)
      for (int k = lane; k < kBins; k += 32) {
        const float re = a.Hr_re[pp][k], im = a.Hr_im[pp][k];
        const float v = (k < 64) ? fmaf(im, im, re * re) : re * re + im * im;
        a.H2[pp][k] = fmaxr(0.f, v);
      }
    __syncwarp();
  }

  // ---- coarse filter update
  if (lane == 0) {
    s.poor_coarse_filter_counter = e2_refined < e2_coarse ? s.poor_coarse_filter_counter + 1 : 0;
    sc.ired[4] = s.poor_coarse_filter_counter < 5;
    if (sc.ired[4]) {
      s.coarse_filter_reset_hangover = imax(s.coarse_filter_reset_hangover - 1, 0);
    } else {
      s.poor_coarse_filter_counter = 0;
      s.coarse_filter_reset_hangover = ec3::kCoarseResetHangover;
    }
  }
  __syncwarp();
  const bool coarse_ok = sc.ired[4] != 0;
  if (!coarse_ok) {
    // coarse_filter_->SetFilter(refined size, refined H) (:733-747)
    const int np = imin(P_c, s.fr_current_size);
    for (int i = lane; i < np * kBinsPad; i += 32) {
      (&a.Hc_re[0][0])[i] = (&a.Hr_re[0][0])[i];
      (&a.Hc_im[0][0])[i] = (&a.Hr_im[0][0])[i];
    }
    __syncwarp();
  }
  if (lane == 0) {
    ++s.cg_call_counter;
    gain_update_current_config(s.cg_cur, s.cg_old, s.cg_tgt, 2, &s.cg_config_change_counter);
    if (poor_excitation) s.cg_poor_excitation_counter = 0;
    sc.ired[5] = (unsigned)(++s.cg_poor_excitation_counter) < (unsigned)P_c || saturated_capture ||
                 (unsigned)s.cg_call_counter <= (unsigned)P_c;
  }
  __syncwarp();
  {
    const bool zero_gain = sc.ired[5] != 0;
    const float rate = s.cg_cur[0], noise_gate = s.cg_cur[1];
    const float* E_re = coarse_ok ? r.Ec_re : r.Er_re;
    const float* E_im = coarse_ok ? r.Ec_im : r.Er_im;
    for (int k = lane; k < kBins; k += 32) {
      if (zero_gain) {
        r.G_re[k] = 0.f;
        r.G_im[k] = 0.f;
      } else {
        const float X2 = r.X2_coa[k];
        float mu = 0.f;
        if (X2 > noise_gate) mu = rate / X2;
        if (r.v0[k] != 0.f) mu = 0.f;
        r.G_re[k] = mu * E_re[k];
        r.G_im[k] = mu * E_im[k];
      }
    }
  }
  __syncwarp();
  if (lane == 0) {
    sc.ired[3] = fir_update_size(&s.fc_current_size, &s.fc_target_size, &s.fc_old_target_size,
                                 &s.fc_size_change_counter, &s.fc_partition_to_constrain);
  }
  __syncwarp();
  {
    const int P = s.fc_current_size;
    fir_zero_partitions(a.Hc_re, a.Hc_im, sc.ired[3], P);
    __syncwarp();
    fir_adapt_partitions(a, sc, a.Hc_re, a.Hc_im, P, r.G_re, r.G_im);
    const int p = s.fc_partition_to_constrain;
    fir_constrain(sc, a.Hc_re[p], a.Hc_im[p], nullptr);
    if (lane == 0) s.fc_partition_to_constrain = p < (P - 1) ? p + 1 : 0;
  }
  // e_refined clamp (subtractor.cc:333-334)
  for (int i = lane; i < kBlock; i += 32) r.e_ref[i] = clampr(r.e_ref[i], -32768.f, 32767.f);
  __syncwarp();
}

}  // namespace wap
