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

// ZeroFilter (adaptive_fir_filter.cc:464-476)
WAP_DEV void fir_zero_partitions(float (*H_re)[kBinsPad], float (*H_im)[kBinsPad], int from, int to) {
  for (int p = from; p < to; ++p)
    for (int k = lane_id(); k < kBinsPad; k += 32) { H_re[p][k] = 0.f; H_im[p][k] = 0.f; }
}

// AdaptiveFirFilter::SetSizePartitions(size, immediate_effect = true) scalars; returns the old size.
WAP_DEV int fir_set_size_immediate(int* cur, int* target, int* old_target, int* counter, int* ptc, int size, int max_size) {
  *target = imin(max_size, size);
  const int old = *cur;
  *cur = *old_target = *target;
  *ptc = imin(*ptc, *cur - 1);
  *counter = 0;
  return old;
}

// AdaptiveFirFilter::UpdateSize scalars (:542-565); returns the old size.
WAP_DEV int fir_update_size(const AecScratch& sc, int* cur, int* target, int* old_target, int* counter, int* ptc) {
  const int old = *cur;
  if (*counter > 0) {
    --*counter;
    const float one_by = 1.f / WAP_EC3(config_change_duration_blocks);
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
    fir_zero_partitions(a.Hr_re, a.Hr_im, s.fr_current_size, WAP_EC3(refined_len));
    fir_zero_partitions(a.Hc_re, a.Hc_im, s.fc_current_size, WAP_EC3(coarse_len));
    #pragma unroll
    for (int k = lane; k < kBins; k += 32) a.H_error[k] = 10000.f;
    __syncwarp();
    if (lane == 0) {
      if (!v.gain_change) {
        s.rg_poor_excitation_counter = 1000;
        s.rg_call_counter = 0;
      }
      s.cg_poor_excitation_counter = 0;
      s.cg_call_counter = 0;
      for (int i = 0; i < 5; ++i) s.rg_cur[i] = s.rg_old[i] = s.rg_tgt[i] = WAP_EC3_ARR(refined_initial)[i];
      s.rg_config_change_counter = 0;
      for (int i = 0; i < 2; ++i) s.cg_cur[i] = s.cg_old[i] = s.cg_tgt[i] = WAP_EC3_ARR(coarse_initial)[i];
      s.cg_config_change_counter = 0;
      sc.ired[0] = fir_set_size_immediate(&s.fr_current_size, &s.fr_target_size, &s.fr_old_target_size,
                                          &s.fr_size_change_counter, &s.fr_partition_to_constrain, WAP_EC3(refined_initial_len),
                                          WAP_EC3(refined_len));
      sc.ired[1] = fir_set_size_immediate(&s.fc_current_size, &s.fc_target_size, &s.fc_old_target_size,
                                          &s.fc_size_change_counter, &s.fc_partition_to_constrain, WAP_EC3(coarse_initial_len),
                                          WAP_EC3(coarse_len));
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
WAP_DEV void subtractor_exit_initial_state(AecScratch& sc) {
  Aec3Scalars& s = sc.s;
  for (int i = 0; i < 5; ++i) { s.rg_old[i] = s.rg_cur[i]; s.rg_tgt[i] = WAP_EC3_ARR(refined)[i]; }
  s.rg_config_change_counter = WAP_EC3(config_change_duration_blocks);
  for (int i = 0; i < 2; ++i) { s.cg_old[i] = s.cg_cur[i]; s.cg_tgt[i] = WAP_EC3_ARR(coarse)[i]; }
  s.cg_config_change_counter = WAP_EC3(config_change_duration_blocks);
  s.fr_target_size = WAP_EC3(refined_len);
  s.fr_size_change_counter = WAP_EC3(config_change_duration_blocks);
  s.fc_target_size = WAP_EC3(coarse_len);
  s.fc_size_change_counter = WAP_EC3(config_change_duration_blocks);
}

// UpdateCurrentConfig of both gain classes, lane 0.
WAP_DEV void gain_update_current_config(const AecScratch& sc, float* cur, float* old, const float* tgt, int n, int* counter) {
  if (*counter > 0) {
    if (--*counter > 0) {
      const float one_by = 1.f / WAP_EC3(config_change_duration_blocks);
      const float change_factor = *counter * one_by;
      for (int i = 0; i < n; ++i) cur[i] = old[i] * change_factor + tgt[i] * (1.f - change_factor);
    } else {
      for (int i = 0; i < n; ++i) cur[i] = old[i] = tgt[i];
    }
  }
}

// RenderSignalAnalyzer::Update (render_signal_analyzer.cc:131-141) for the
// spectrum at `delay_partitions` and the latest render block in sc.x.
// x_band1: band 1 of the latest render block for 3-band legs (else nullptr).
WAP_DEV void render_signal_analyzer_update(Aec3State& a, AecScratch& sc, int delay_partitions, const float* x_band1 = nullptr) {
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
  #pragma unroll
  for (int i = lane; i < kBlock; i += 32) max_abs_l = fmaxf(max_abs_l, fabsf(sc.x[i]));
  if (x_band1)
    #pragma unroll
    for (int i = lane; i < kBlock; i += 32) max_abs_l = fmaxf(max_abs_l, fabsf(x_band1[i]));
  const float max_abs = warp_max(max_abs_l);
  if (lane == 0) {
    if (s.rsa_has_narrow_peak && ++s.rsa_narrow_peak_counter > WAP_EC3(refined_len)) s.rsa_has_narrow_peak = 0;
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
  #pragma unroll
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

// Row of the render FFT / spectrum rings that partition p pairs with (adaptive_fir_filter.cc:139-151).
WAP_DEV int ring_row(int pos, int p) {
  const int row = pos + p;
  return row >= kRingBlocks ? row - kRingBlocks : row;
}

// Both adaptive filters at once: S = sum_p X_p * H_p (ApplyFilter_Avx2: per bin, partitions
// in order, separate multiplies and adds) for the refined filter (-> sc.fftA, packed) and the
// coarse filter (-> sc.fftB); the render partitions X_p are loaded once and shared.  All loads
// of a bin are issued before the arithmetic so they overlap in the memory system.
WAP_DEV void fir_filter_both(const Aec3State& a, AecScratch& sc, int P_r, int P_c) {
  const int lane = lane_id();
  const int pos = sc.s.spectra_read;
  const int pmax = imax(P_r, P_c);
  __syncwarp();
  // Bin 64 (real only: FftData::CopyToPackedArray drops im[64]) would cost a third, almost empty
  // pass of the loop below.  Instead lane p forms partition p's term and the terms are added in
  // partition order through shuffles -- the same sequence of additions.
  float t_ref = 0.f, t_coa = 0.f;
  if (lane < pmax) {
    const int row = ring_row(pos, lane);
    const float X_re = a.fft_re[row][64], X_im = a.fft_im[row][64];
    if (lane < P_r) t_ref = X_re * a.Hr_re[lane][64] - X_im * a.Hr_im[lane][64];
    if (lane < P_c) t_coa = X_re * a.Hc_re[lane][64] - X_im * a.Hc_im[lane][64];
  }
#pragma unroll
  for (int k = lane; k < 64; k += 32) {
    float Xr[kMaxPartitions], Xi[kMaxPartitions];
#pragma unroll
    for (int p = 0; p < kMaxPartitions; ++p) {
      const int row = ring_row(pos, p);
      Xr[p] = p < pmax ? a.fft_re[row][k] : 0.f;
      Xi[p] = p < pmax ? a.fft_im[row][k] : 0.f;
    }
    float Hre[kMaxPartitions], Him[kMaxPartitions];
#pragma unroll
    for (int p = 0; p < kMaxPartitions; ++p) {
      Hre[p] = p < P_r ? a.Hr_re[p][k] : 0.f;
      Him[p] = p < P_r ? a.Hr_im[p][k] : 0.f;
    }
    float S_re = 0.f, S_im = 0.f;
#pragma unroll
    for (int p = 0; p < kMaxPartitions; ++p) {
      if (p < P_r) {
        const float aa = Xr[p] * Hre[p], bb = Xi[p] * Him[p], cc = Xr[p] * Him[p], dd = Xi[p] * Hre[p];
        S_re = S_re + (aa - bb);
        S_im = S_im + (cc + dd);
      }
    }
    if (k == 0) sc.fftA[0] = S_re;
    else { sc.fftA[2 * k] = S_re; sc.fftA[2 * k + 1] = S_im; }
#pragma unroll
    for (int p = 0; p < kMaxPartitions; ++p) {
      Hre[p] = p < P_c ? a.Hc_re[p][k] : 0.f;
      Him[p] = p < P_c ? a.Hc_im[p][k] : 0.f;
    }
    S_re = 0.f; S_im = 0.f;
#pragma unroll
    for (int p = 0; p < kMaxPartitions; ++p) {
      if (p < P_c) {
        const float aa = Xr[p] * Hre[p], bb = Xi[p] * Him[p], cc = Xr[p] * Him[p], dd = Xi[p] * Hre[p];
        S_re = S_re + (aa - bb);
        S_im = S_im + (cc + dd);
      }
    }
    if (k == 0) sc.fftB[0] = S_re;
    else { sc.fftB[2 * k] = S_re; sc.fftB[2 * k + 1] = S_im; }
  }
  {
    float S_ref = 0.f, S_coa = 0.f;
#pragma unroll
    for (int p = 0; p < kMaxPartitions; ++p) {
      const float tr = __shfl_sync(WAP_FULL, t_ref, p), tc = __shfl_sync(WAP_FULL, t_coa, p);
      if (p < P_r) S_ref = S_ref + tr;
      if (p < P_c) S_coa = S_coa + tc;
    }
    if (lane == 0) {
      sc.fftA[1] = S_ref;
      sc.fftB[1] = S_coa;
    }
  }
}

// PredictionError (subtractor.cc:49-65) from the inverse transform in `buf`.
WAP_DEV void prediction_error(const float* buf, const float* y, float* e_out, float* s_out) {
  constexpr float kScale = 1.0f / 64;
  #pragma unroll
  for (int i = lane_id(); i < kBlock; i += 32) {
    const float t = buf[kBlock + i];
    e_out[i] = y[i] - t * kScale;
    s_out[i] = kScale * t;
  }
}

// AdaptPartitions_Avx2: H_p += conj(X_p) * G for one filter (used on the rare coarse re-seed path).
WAP_DEV void fir_adapt_partitions(const Aec3State& a, AecScratch& sc, float (*H_re)[kBinsPad], float (*H_im)[kBinsPad],
                                  int num_partitions, const float* G_re, const float* G_im) {
  const int pos = sc.s.spectra_read;
  for (int k = lane_id(); k < kBins; k += 32) {
    const float Gre = G_re[k], Gim = G_im[k];
    for (int p = 0; p < num_partitions; ++p) {
      const int row = ring_row(pos, p);
      const float X_re = a.fft_re[row][k], X_im = a.fft_im[row][k];
      const float aa = X_re * Gre, bb = X_im * Gim, cc = X_re * Gim, dd = X_im * Gre;
      H_re[p][k] = H_re[p][k] + (aa + bb);
      H_im[p][k] = H_im[p][k] + (cc - dd);
    }
  }
  __syncwarp();
}

// ComputeFrequencyResponse_Avx2 for one bin of one partition (single render channel: max with 0).
WAP_DEV float h2_bin(float re, float im, int k) {
  const float v = (k < 64) ? fmaf(im, im, re * re) : re * re + im * im;
  return fmaxr(0.f, v);
}

// Both filters' AdaptPartitions_Avx2 in one pass over the render partitions (X_p loaded once),
// plus ComputeFrequencyResponse of the refined filter from the values just written (the partition
// that Constrain() rewrites afterwards is redone by the caller).  coarse == false: refined only.
WAP_DEV void fir_adapt_both(Aec3State& a, AecScratch& sc, int P_r, int P_c, const float* Gr_re, const float* Gr_im,
                            const float* Gc_re, const float* Gc_im, bool coarse) {
  const int lane = lane_id();
  const int pos = sc.s.spectra_read;
  const int pmax = coarse ? imax(P_r, P_c) : P_r;
  // Bin 64: one partition per lane instead of a third, almost empty pass over the bins.
  if (lane < pmax) {
    const int row = ring_row(pos, lane);
    const float X_re = a.fft_re[row][64], X_im = a.fft_im[row][64];
    if (lane < P_r) {
      const float Gre = Gr_re[64], Gim = Gr_im[64];
      const float aa = X_re * Gre, bb = X_im * Gim, cc = X_re * Gim, dd = X_im * Gre;
      const float re = a.Hr_re[lane][64] + (aa + bb), im = a.Hr_im[lane][64] + (cc - dd);
      a.Hr_re[lane][64] = re;
      a.Hr_im[lane][64] = im;
      a.H2[lane][64] = h2_bin(re, im, 64);
    }
    if (coarse && lane < P_c) {
      const float Gre = Gc_re[64], Gim = Gc_im[64];
      const float aa = X_re * Gre, bb = X_im * Gim, cc = X_re * Gim, dd = X_im * Gre;
      a.Hc_re[lane][64] = a.Hc_re[lane][64] + (aa + bb);
      a.Hc_im[lane][64] = a.Hc_im[lane][64] + (cc - dd);
    }
  }
#pragma unroll
  for (int k = lane; k < 64; k += 32) {
    float Xr[kMaxPartitions], Xi[kMaxPartitions];
#pragma unroll
    for (int p = 0; p < kMaxPartitions; ++p) {
      const int row = ring_row(pos, p);
      Xr[p] = p < pmax ? a.fft_re[row][k] : 0.f;
      Xi[p] = p < pmax ? a.fft_im[row][k] : 0.f;
    }
    float Hre[kMaxPartitions], Him[kMaxPartitions];
#pragma unroll
    for (int p = 0; p < kMaxPartitions; ++p) {
      Hre[p] = p < P_r ? a.Hr_re[p][k] : 0.f;
      Him[p] = p < P_r ? a.Hr_im[p][k] : 0.f;
    }
    {
      const float Gre = Gr_re[k], Gim = Gr_im[k];
#pragma unroll
      for (int p = 0; p < kMaxPartitions; ++p) {
        if (p < P_r) {
          const float aa = Xr[p] * Gre, bb = Xi[p] * Gim, cc = Xr[p] * Gim, dd = Xi[p] * Gre;
          const float re = Hre[p] + (aa + bb), im = Him[p] + (cc - dd);
          a.Hr_re[p][k] = re;
          a.Hr_im[p][k] = im;
          a.H2[p][k] = h2_bin(re, im, k);
        }
      }
    }
    if (coarse) {
#pragma unroll
      for (int p = 0; p < kMaxPartitions; ++p) {
        Hre[p] = p < P_c ? a.Hc_re[p][k] : 0.f;
        Him[p] = p < P_c ? a.Hc_im[p][k] : 0.f;
      }
      const float Gre = Gc_re[k], Gim = Gc_im[k];
#pragma unroll
      for (int p = 0; p < kMaxPartitions; ++p) {
        if (p < P_c) {
          const float aa = Xr[p] * Gre, bb = Xi[p] * Gim, cc = Xr[p] * Gim, dd = Xi[p] * Gre;
          a.Hc_re[p][k] = Hre[p] + (aa + bb);
          a.Hc_im[p][k] = Him[p] + (cc - dd);
        }
      }
    }
  }
  __syncwarp();
}

// AdaptiveFirFilter::Constrain / ConstrainAndUpdateImpulseResponse (:645-706) for one partition
// of each filter side by side (refined on lanes 0-15 / fftA, coarse on lanes 16-31 / fftB);
// `impulse` receives the 64 retained taps of the refined partition.  second == false: refined only.
WAP_DEV void fir_constrain_pair(AecScratch& sc, float* Hr_re_p, float* Hr_im_p, float* impulse, float* Hc_re_p,
                                float* Hc_im_p, bool second) {
  const int lane = lane_id();
  __syncwarp();
  reim_to_packed(Hr_re_p, Hr_im_p, sc.fftA);
  if (second) reim_to_packed(Hc_re_p, Hc_im_p, sc.fftB);
  fft_pair(sc, true, second);
  constexpr float kScale = 1.0f / 64;
  #pragma unroll
  for (int i = lane; i < kBlock; i += 32) {
    const float v = sc.fftA[i] * kScale;
    sc.fftA[i] = v;
    sc.fftA[kBlock + i] = 0.f;
    impulse[i] = v;
    if (second) {
      sc.fftB[i] = sc.fftB[i] * kScale;
      sc.fftB[kBlock + i] = 0.f;
    }
  }
  fft_pair(sc, false, second);
  packed_to_reim(sc.fftA, Hr_re_p, Hr_im_p);
  if (second) packed_to_reim(sc.fftB, Hc_re_p, Hc_im_p);
  __syncwarp();
}
// One partition of one filter (coarse re-seed path).
WAP_DEV void fir_constrain(AecScratch& sc, float* H_re_p, float* H_im_p) {
  const int lane = lane_id();
  __syncwarp();
  reim_to_packed(H_re_p, H_im_p, sc.fftA);
  fft_pair(sc, true, false);
  constexpr float kScale = 1.0f / 64;
  #pragma unroll
  for (int i = lane; i < kBlock; i += 32) {
    sc.fftA[i] = sc.fftA[i] * kScale;
    sc.fftA[kBlock + i] = 0.f;
  }
  fft_pair(sc, false, false);
  packed_to_reim(sc.fftA, H_re_p, H_im_p);
  __syncwarp();
}

// Aec3Fft::ZeroPaddedFft(x, kHanning) into the packed buffer `buf` (not transformed yet).
WAP_DEV void stage_zero_padded_hanning(const float* x, float* buf) {
  #pragma unroll
  for (int i = lane_id(); i < kBlock; i += 32) {
    buf[i] = 0.f;
    buf[kBlock + i] = x[i] * kHanning64[i];
  }
}

// Subtractor::Process (subtractor.cc:196-343).  Inputs: capture block sc.y,
// render rings; outputs in sc.rm (e_ref, e_coa, s_ref, s_coa, E2_ref, E2_coa,
// Er) and sc.red[0..6] = {y2, e2_refined, e2_coarse, s2_refined, s2_coarse,
// s_refined_max_abs, s_coarse_max_abs}.
//
// The refined and the coarse filter are processed together wherever the reference's
// data flow allows it (their outputs, their gains and, unless the coarse filter is
// being re-seeded from the refined one, their adaptation are independent): the
// render partitions are read once per pass instead of once per filter, and the
// 128-point transforms run two at a time on the two half-warps.
WAP_DEV void subtractor_process(Aec3State& a, AecScratch& sc, bool saturated_capture) {
  const int lane = lane_id();
  Aec3Scalars& s = sc.s;
  AecRemoverScratch& r = sc.rm;
  __syncwarp();
  const int P_r = s.fr_current_size, P_c = s.fc_current_size;
  // RenderBuffer::SpectralSum(s) (render_buffer.cc:42-83): one running sum per bin (bin 64: one
  // partition per lane, added in order through shuffles).
  {
    const int pmax = imax(P_r, P_c);
    const int pos = s.spectra_read;
    const float x64 = lane < pmax ? a.spectra[ring_row(pos, lane)][64] : 0.f;
#pragma unroll
    for (int k = lane; k < 64; k += 32) {
      float X2[kMaxPartitions];
#pragma unroll
      for (int j = 0; j < kMaxPartitions; ++j) X2[j] = j < pmax ? a.spectra[ring_row(pos, j)][k] : 0.f;
      float x2 = 0.f;
#pragma unroll
      for (int j = 0; j < kMaxPartitions; ++j) {
        if (j < pmax) x2 += X2[j];
        if (j + 1 == P_r) r.X2_ref[k] = x2;
        if (j + 1 == P_c) r.X2_coa[k] = x2;
      }
    }
    float x2 = 0.f;
#pragma unroll
    for (int j = 0; j < kMaxPartitions; ++j) {
      const float t = __shfl_sync(WAP_FULL, x64, j);
      if (j < pmax) x2 += t;
      if (lane == 0 && j + 1 == P_r) r.X2_ref[64] = x2;
      if (lane == 0 && j + 1 == P_c) r.X2_coa[64] = x2;
    }
  }
  fir_filter_both(a, sc, P_r, P_c);
  fft_pair(sc, true, true);
  prediction_error(sc.fftA, sc.y, r.e_ref, r.s_ref);
  prediction_error(sc.fftB, sc.y, r.e_coa, r.s_coa);
  __syncwarp();

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
    #pragma unroll
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

  // ---- scalar part of both updates (lane 0), in the reference's order: refined gain,
  // refined UpdateSize, coarse-filter bookkeeping, coarse gain, coarse UpdateSize.
  if (lane == 0) {
    sc.ired[1] = 1;  // refined G == 0 ?
    if (!refined_filters_adjusted) {
      ++s.rg_call_counter;
      gain_update_current_config(sc, s.rg_cur, s.rg_old, s.rg_tgt, 5, &s.rg_config_change_counter);
      if (poor_excitation) s.rg_poor_excitation_counter = 0;
      const bool zero = (unsigned)(++s.rg_poor_excitation_counter) < (unsigned)P_r || saturated_capture ||
                        (unsigned)s.rg_call_counter <= (unsigned)P_r;
      sc.ired[1] = zero;
    }
    sc.ired[2] = s.coarse_filter_reset_hangover > 0;  // disallow_leakage_diverged
    sc.ired[3] = fir_update_size(sc, &s.fr_current_size, &s.fr_target_size, &s.fr_old_target_size,
                                 &s.fr_size_change_counter, &s.fr_partition_to_constrain);
    s.poor_coarse_filter_counter = e2_refined < e2_coarse ? s.poor_coarse_filter_counter + 1 : 0;
    sc.ired[4] = s.poor_coarse_filter_counter < 5;
    if (sc.ired[4]) {
      s.coarse_filter_reset_hangover = imax(s.coarse_filter_reset_hangover - 1, 0);
    } else {
      s.poor_coarse_filter_counter = 0;
      s.coarse_filter_reset_hangover = WAP_EC3(coarse_reset_hangover_blocks);
    }
    ++s.cg_call_counter;
    gain_update_current_config(sc, s.cg_cur, s.cg_old, s.cg_tgt, 2, &s.cg_config_change_counter);
    if (poor_excitation) s.cg_poor_excitation_counter = 0;
    sc.ired[5] = (unsigned)(++s.cg_poor_excitation_counter) < (unsigned)P_c || saturated_capture ||
                 (unsigned)s.cg_call_counter <= (unsigned)P_c;
    sc.ired[6] = fir_update_size(sc, &s.fc_current_size, &s.fc_target_size, &s.fc_old_target_size,
                                 &s.fc_size_change_counter, &s.fc_partition_to_constrain);
  }
  __syncwarp();
  const bool coarse_ok = sc.ired[4] != 0;
  const int P_r2 = s.fr_current_size, P_c2 = s.fc_current_size;  // after UpdateSize
  // ---- RefinedFilterUpdateGain::Compute -> r.G ; CoarseFilterUpdateGain::Compute -> r.v1 / r.v2
  {
    const bool zero_ref = sc.ired[1] != 0, zero_coa = sc.ired[5] != 0;
    const bool disallow_leakage_diverged = sc.ired[2] != 0;
    const float leak_conv = s.rg_cur[0], leak_div = s.rg_cur[1], err_floor = s.rg_cur[2], err_ceil = s.rg_cur[3],
                noise_gate = s.rg_cur[4];
    const float rate = s.cg_cur[0], noise_gate_c = s.cg_cur[1];
    const float* Ecx_re = coarse_ok ? r.Ec_re : r.Er_re;
    const float* Ecx_im = coarse_ok ? r.Ec_im : r.Er_im;
    const int H2_size = s.H2_size;
    #pragma unroll
    for (int k = lane; k < kBins; k += 32) {
      const bool masked = r.v0[k] != 0.f;
      if (!refined_filters_adjusted) {
        // ComputeErl (adaptive_fir_filter_erl_avx2.cc:27-40)
        float h2[kMaxPartitions];
#pragma unroll
        for (int j = 0; j < kMaxPartitions; ++j) h2[j] = j < H2_size ? a.H2[j][k] : 0.f;
        float erl = 0.f;
#pragma unroll
        for (int j = 0; j < kMaxPartitions; ++j)
          if (j < H2_size) erl += h2[j];
        float H_error = a.H_error[k];
        const float X2 = r.X2_ref[k], E2r = r.E2_ref[k];
        if (zero_ref) {
          r.G_re[k] = 0.f;
          r.G_im[k] = 0.f;
        } else {
          float mu = 0.f;
          if (X2 >= noise_gate) mu = H_error / (0.5f * H_error * X2 + (float)P_r * E2r);
          if (masked) mu = 0.f;
          H_error -= 0.5f * mu * X2 * H_error;
          r.G_re[k] = mu * r.Er_re[k];
          r.G_im[k] = mu * r.Er_im[k];
        }
        if (E2r <= r.E2_coa[k] || disallow_leakage_diverged) H_error += leak_conv * erl;
        else H_error += leak_div * erl;
        H_error = fmaxr(H_error, err_floor);
        H_error = fminr(H_error, err_ceil);
        a.H_error[k] = H_error;
      } else {
        r.G_re[k] = 0.f;
        r.G_im[k] = 0.f;
      }
      if (zero_coa) {
        r.v1[k] = 0.f;
        r.v2[k] = 0.f;
      } else {
        const float X2 = r.X2_coa[k];
        float mu = 0.f;
        if (X2 > noise_gate_c) mu = rate / X2;
        if (masked) mu = 0.f;
        r.v1[k] = mu * Ecx_re[k];
        r.v2[k] = mu * Ecx_im[k];
      }
    }
  }
  __syncwarp();
  // ---- AdaptiveFirFilter::Adapt for both filters
  fir_zero_partitions(a.Hr_re, a.Hr_im, sc.ired[3], P_r2);   // UpdateSize: newly exposed partitions
  // impulse_response->resize(): newly exposed taps are zero.
  for (int i = s.h_time_size * kBlock + lane; i < P_r2 * kBlock; i += 32) a.h_time[i] = 0.f;
  if (coarse_ok) fir_zero_partitions(a.Hc_re, a.Hc_im, sc.ired[6], P_c2);
  __syncwarp();
  const int pr = s.fr_partition_to_constrain, pc = s.fc_partition_to_constrain;
  if (coarse_ok) {
    fir_adapt_both(a, sc, P_r2, P_c2, r.G_re, r.G_im, r.v1, r.v2, true);
    fir_constrain_pair(sc, a.Hr_re[pr], a.Hr_im[pr], a.h_time + pr * kBlock, a.Hc_re[pc], a.Hc_im[pc], true);
  } else {
    // Coarse filter re-seeded from the adapted refined filter (subtractor.cc:305-316):
    // strictly sequential.
    fir_adapt_both(a, sc, P_r2, P_c2, r.G_re, r.G_im, r.v1, r.v2, false);
    fir_constrain_pair(sc, a.Hr_re[pr], a.Hr_im[pr], a.h_time + pr * kBlock, nullptr, nullptr, false);
    // coarse_filter_->SetFilter(refined size, refined H) (:733-747)
    const int np = imin(P_c, P_r2);
    for (int i = lane; i < np * kBinsPad; i += 32) {
      (&a.Hc_re[0][0])[i] = (&a.Hr_re[0][0])[i];
      (&a.Hc_im[0][0])[i] = (&a.Hr_im[0][0])[i];
    }
    __syncwarp();
    fir_zero_partitions(a.Hc_re, a.Hc_im, sc.ired[6], P_c2);
    __syncwarp();
    fir_adapt_partitions(a, sc, a.Hc_re, a.Hc_im, P_c2, r.v1, r.v2);
    fir_constrain(sc, a.Hc_re[pc], a.Hc_im[pc]);
  }
  // ComputeFrequencyResponse of the partition Constrain() rewrote.
  #pragma unroll
  for (int k = lane; k < kBins; k += 32) a.H2[pr][k] = h2_bin(a.Hr_re[pr][k], a.Hr_im[pr][k], k);
  if (lane == 0) {
    s.h_time_size = P_r2;
    s.H2_size = P_r2;
    s.fr_partition_to_constrain = pr < (P_r2 - 1) ? pr + 1 : 0;
    s.fc_partition_to_constrain = pc < (P_c2 - 1) ? pc + 1 : 0;
  }
  // e_refined clamp (subtractor.cc:333-334)
  #pragma unroll
  for (int i = lane; i < kBlock; i += 32) r.e_ref[i] = clampr(r.e_ref[i], -32768.f, 32767.f);
  __syncwarp();
}

}  // namespace wap
