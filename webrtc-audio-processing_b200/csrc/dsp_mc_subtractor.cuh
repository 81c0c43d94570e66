// Linear echo canceller of a multi-channel leg: one Subtractor channel (refined + coarse AdaptiveFirFilter
// with R render channels each) per capture channel.
//   Subtractor::{Process, HandleEchoPathChange, ExitInitialState}   aec3/subtractor.cc:156-343
//   AdaptiveFirFilter (R render channels)                           aec3/adaptive_fir_filter.cc:464-747
//   ApplyFilter_Avx2 / AdaptPartitions_Avx2 / ComputeFrequencyResponse_Avx2
//                                                                   aec3/adaptive_fir_filter_avx2.cc:30-194
//   RenderBuffer::SpectralSum(s)                                    aec3/render_buffer.cc:42-83
//   RenderSignalAnalyzer (R render channels)                        aec3/render_signal_analyzer.cc:33-141
// A filter's partitions are stored as "virtual partitions" v = p * R + render_channel: the reference's loops
// run `for p { for ch }`, so walking v in order is the reference's order of additions for every bin.
// Scalars: `sc.s` is the staged copy of the shared Aec3Scalars (StreamState::aec.s), `cs` the staged copy
// of the capture channel's (McChan::s).
#pragma once

#include "dsp_aec3_subtractor.cuh"
#include "wap_mc_state.h"

namespace wap {

// What the stages of one capture channel leave for the stages after the channel loop.
struct McChanVec {
  float y[kBlock];           // capture block, band 0 (in / out)
  float e[kBlock];           // formed linear output
  float Y_re[kBinsPad], Y_im[kBinsPad], E_re[kBinsPad], E_im[kBinsPad];
  float Y2[kBinsPad], E2[kBinsPad], S2_lin[kBinsPad], R2[kBinsPad], R2_unb[kBinsPad];
  float N_re[kBinsPad], N_im[kBinsPad];      // comfort noise, lower band
  float hb_re[kBinsPad], hb_im[kBinsPad];    // comfort noise, upper bands
  float metrics[8];          // y2, e2_refined, e2_coarse, s2_refined, s2_coarse, s_refined_max_abs, s_coarse_max_abs
};
struct McExtra {
  Aec3Scalars cs[kMcCh];     // staged per-channel scalars
  McChanVec cv[kMcCh];
  float x1[kBlock];          // band 0 of GetBlock(0), render channel 1 (channel 0: sc.x)
};

WAP_DEV void mc_zero_vparts(float (*H_re)[kBinsPad], float (*H_im)[kBinsPad], int from, int to) {
  for (int i = from * kBinsPad + lane_id(); i < to * kBinsPad; i += 32) {
    (&H_re[0][0])[i] = 0.f;
    (&H_im[0][0])[i] = 0.f;
  }
}

// Subtractor::HandleEchoPathChange for capture channel state (f, ch, cs).
WAP_DEV void mc_subtractor_handle_echo_path_change(McFilters& f, McChan& ch, Aec3Scalars& cs, AecScratch& sc,
                                                   const EchoPathVariability& v, int R) {
  const int lane = lane_id();
  if (v.delay_change == kDelayAdjNone) return;
  __syncwarp();
  mc_zero_vparts(f.Hr_re, f.Hr_im, cs.fr_current_size * R, WAP_EC3(refined_len) * R);
  mc_zero_vparts(f.Hc_re, f.Hc_im, cs.fc_current_size * R, WAP_EC3(coarse_len) * R);
  for (int k = lane; k < kBins; k += 32) ch.H_error[k] = 10000.f;
  __syncwarp();
  if (lane == 0) {
    Aec3Scalars& s = cs;
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
                                        &s.fr_size_change_counter, &s.fr_partition_to_constrain,
                                        WAP_EC3(refined_initial_len), WAP_EC3(refined_len));
    sc.ired[1] = fir_set_size_immediate(&s.fc_current_size, &s.fc_target_size, &s.fc_old_target_size,
                                        &s.fc_size_change_counter, &s.fc_partition_to_constrain,
                                        WAP_EC3(coarse_initial_len), WAP_EC3(coarse_len));
  }
  __syncwarp();
  mc_zero_vparts(f.Hr_re, f.Hr_im, sc.ired[0] * R, cs.fr_current_size * R);
  mc_zero_vparts(f.Hc_re, f.Hc_im, sc.ired[1] * R, cs.fc_current_size * R);
  __syncwarp();
}

// Subtractor::ExitInitialState for one channel's scalars, lane 0.
WAP_DEV void mc_subtractor_exit_initial_state(Aec3Scalars& s, AecScratch& sc) {
  for (int i = 0; i < 5; ++i) { s.rg_old[i] = s.rg_cur[i]; s.rg_tgt[i] = WAP_EC3_ARR(refined)[i]; }
  s.rg_config_change_counter = WAP_EC3(config_change_duration_blocks);
  for (int i = 0; i < 2; ++i) { s.cg_old[i] = s.cg_cur[i]; s.cg_tgt[i] = WAP_EC3_ARR(coarse)[i]; }
  s.cg_config_change_counter = WAP_EC3(config_change_duration_blocks);
  s.fr_target_size = WAP_EC3(refined_len);
  s.fr_size_change_counter = WAP_EC3(config_change_duration_blocks);
  s.fc_target_size = WAP_EC3(coarse_len);
  s.fc_size_change_counter = WAP_EC3(config_change_duration_blocks);
}

// RenderSignalAnalyzer::Update with R render channels.  sc.x / x1: band 0 of the latest render block;
// x_b1[rc]: its band 1 (3-band legs) or nullptr.
WAP_DEV void mc_render_signal_analyzer_update(Aec3State& sh, const McRender& rb, AecScratch& sc, const float* x1,
                                              int delay_partitions, int R, const float* x_b1_0, const float* x_b1_1) {
  const int lane = lane_id();
  Aec3Scalars& s = sc.s;
  __syncwarp();
  {
    const int row = ring_off(s.spectra_read, delay_partitions, kRingBlocks);
    for (int k = 1 + lane; k < 64; k += 32) {
      int cnt = 0;
      for (int rc = 0; rc < R; ++rc) {
        const float* X2 = rb.spectra[row][rc];
        cnt += (X2[k] > 3 * fmaxr(X2[k - 1], X2[k + 1])) ? 1 : 0;
      }
      sh.narrow_band_counters[k - 1] = cnt > 0 ? sh.narrow_band_counters[k - 1] + 1 : 0;
    }
  }
  __syncwarp();
  if (lane == 0 && s.rsa_has_narrow_peak && ++s.rsa_narrow_peak_counter > WAP_EC3(refined_len)) s.rsa_has_narrow_peak = 0;
  float max_peak_level = 0.f;
  for (int rc = 0; rc < R; ++rc) {
    __syncwarp();
    const float* X2_latest = rb.spectra[s.spectra_read][rc];
    const int peak_bin = warp_argmax_first(X2_latest, kBins);
    const float* xb = rc == 0 ? sc.x : x1;
    const float* xb1 = rc == 0 ? x_b1_0 : x_b1_1;
    float max_abs_l = 0.f;
    for (int i = lane; i < kBlock; i += 32) max_abs_l = fmaxf(max_abs_l, fabsf(xb[i]));
    if (xb1)
      for (int i = lane; i < kBlock; i += 32) max_abs_l = fmaxf(max_abs_l, fabsf(xb1[i]));
    const float max_abs = warp_max(max_abs_l);
    float non_peak_power = 0.f;
    for (int k = imax(0, peak_bin - 14); k < peak_bin - 4; ++k) non_peak_power = fmaxr(X2_latest[k], non_peak_power);
    for (int k = peak_bin + 5; k < imin(peak_bin + 15, kBins); ++k) non_peak_power = fmaxr(X2_latest[k], non_peak_power);
    const float peak_level = X2_latest[peak_bin];
    if (peak_bin > 0 && max_abs > 100 && peak_level > 100 * non_peak_power) {
      if (peak_level > max_peak_level) {
        max_peak_level = peak_level;
        __syncwarp();
        if (lane == 0) {
          s.rsa_has_narrow_peak = 1;
          s.rsa_narrow_peak_band = peak_bin;
          s.rsa_narrow_peak_counter = 0;
        }
      }
    }
  }
  __syncwarp();
}

// RenderBuffer::SpectralSum(s): X2 summed over partitions (outer) and render channels (inner), with the
// running sum captured after P_r and after P_c partitions -> r.X2_ref / r.X2_coa.
template <int R>
WAP_DEV void mc_spectral_sums_r(const McRender& rb, AecScratch& sc, int P_r, int P_c) {
  AecRemoverScratch& r = sc.rm;
  const int pos = sc.s.spectra_read;
  const int pmax = imax(P_r, P_c);
  __syncwarp();
  for (int k = lane_id(); k < kBins; k += 32) {
    float x2 = 0.f;
#pragma unroll 1
    for (int j0 = 0; j0 < pmax; j0 += 4) {
      float v[4][R];   // one batch of loads, then the additions in the reference's order
#pragma unroll
      for (int jj = 0; jj < 4; ++jj) {
        const int row = ring_row(pos, imin(j0 + jj, pmax - 1));
#pragma unroll
        for (int rc = 0; rc < R; ++rc) v[jj][rc] = rb.spectra[row][rc][k];
      }
#pragma unroll
      for (int jj = 0; jj < 4; ++jj) {
        const int j = j0 + jj;
        if (j < pmax) {
#pragma unroll
          for (int rc = 0; rc < R; ++rc) x2 += v[jj][rc];
          if (j + 1 == P_r) r.X2_ref[k] = x2;
          if (j + 1 == P_c) r.X2_coa[k] = x2;
        }
      }
    }
  }
  __syncwarp();
}
WAP_DEV void mc_spectral_sums(const McRender& rb, AecScratch& sc, int P_r, int P_c, int R) {
  if (R == 2) mc_spectral_sums_r<2>(rb, sc, P_r, P_c);
  else mc_spectral_sums_r<1>(rb, sc, P_r, P_c);
}

// Both filters of one capture channel: S = sum over virtual partitions of X * H (ApplyFilter_Avx2's
// operations per bin) -> sc.fftA (refined, packed) / sc.fftB (coarse).
template <int R>
WAP_DEV void mc_fir_filter_both_r(const McRender& rb, const McFilters& f, AecScratch& sc, int P_r, int P_c) {
  const int lane = lane_id();
  const int pos = sc.s.spectra_read;
  const int pmax = imax(P_r, P_c);
  __syncwarp();
  // bin 64: lane v forms virtual partition v's term; the terms are added in order through shuffles
  float t_ref = 0.f, t_coa = 0.f;
  {
    const int p = R == 2 ? (lane >> 1) : lane, rc = R == 2 ? (lane & 1) : 0;
    if (p < pmax) {
      const int row = ring_row(pos, p);
      const float X_re = rb.fft_re[row][rc][64], X_im = rb.fft_im[row][rc][64];
      if (p < P_r) t_ref = X_re * f.Hr_re[lane][64] - X_im * f.Hr_im[lane][64];
      if (p < P_c) t_coa = X_re * f.Hc_re[lane][64] - X_im * f.Hc_im[lane][64];
    }
  }
  for (int k = lane; k < 64; k += 32) {
    float Sr_re = 0.f, Sr_im = 0.f, Sc_re = 0.f, Sc_im = 0.f;
#pragma unroll 1
    for (int p0 = 0; p0 < pmax; p0 += 2) {
      // one batch of loads (2 partitions x R channels x {X, Hr, Hc} x {re, im}), then the arithmetic in order
      float Xr[2][R], Xi[2][R], Ar[2][R], Ai[2][R], Cr[2][R], Ci[2][R];
#pragma unroll
      for (int pp = 0; pp < 2; ++pp) {
        const int p = imin(p0 + pp, pmax - 1);
        const int row = ring_row(pos, p);
#pragma unroll
        for (int rc = 0; rc < R; ++rc) {
          const int v = p * R + rc;
          Xr[pp][rc] = rb.fft_re[row][rc][k];
          Xi[pp][rc] = rb.fft_im[row][rc][k];
          Ar[pp][rc] = f.Hr_re[v][k];
          Ai[pp][rc] = f.Hr_im[v][k];
          Cr[pp][rc] = f.Hc_re[v][k];
          Ci[pp][rc] = f.Hc_im[v][k];
        }
      }
#pragma unroll
      for (int pp = 0; pp < 2; ++pp) {
        const int p = p0 + pp;
#pragma unroll
        for (int rc = 0; rc < R; ++rc) {
          if (p < P_r) {
            const float aa = Xr[pp][rc] * Ar[pp][rc], bb = Xi[pp][rc] * Ai[pp][rc], cc = Xr[pp][rc] * Ai[pp][rc],
                        dd = Xi[pp][rc] * Ar[pp][rc];
            Sr_re = Sr_re + (aa - bb);
            Sr_im = Sr_im + (cc + dd);
          }
          if (p < P_c) {
            const float aa = Xr[pp][rc] * Cr[pp][rc], bb = Xi[pp][rc] * Ci[pp][rc], cc = Xr[pp][rc] * Ci[pp][rc],
                        dd = Xi[pp][rc] * Cr[pp][rc];
            Sc_re = Sc_re + (aa - bb);
            Sc_im = Sc_im + (cc + dd);
          }
        }
      }
    }
    if (k == 0) { sc.fftA[0] = Sr_re; sc.fftB[0] = Sc_re; }
    else { sc.fftA[2 * k] = Sr_re; sc.fftA[2 * k + 1] = Sr_im; sc.fftB[2 * k] = Sc_re; sc.fftB[2 * k + 1] = Sc_im; }
  }
  {
    float S_ref = 0.f, S_coa = 0.f;
    const int nv_r = P_r * R, nv_c = P_c * R;
#pragma unroll
    for (int v = 0; v < kMaxPartitions * R; ++v) {
      const float tr = __shfl_sync(WAP_FULL, t_ref, v), tc = __shfl_sync(WAP_FULL, t_coa, v);
      if (v < nv_r) S_ref = S_ref + tr;
      if (v < nv_c) S_coa = S_coa + tc;
    }
    if (lane == 0) {
      sc.fftA[1] = S_ref;
      sc.fftB[1] = S_coa;
    }
  }
}
WAP_DEV void mc_fir_filter_both(const McRender& rb, const McFilters& f, AecScratch& sc, int P_r, int P_c, int R) {
  if (R == 2) mc_fir_filter_both_r<2>(rb, f, sc, P_r, P_c);
  else mc_fir_filter_both_r<1>(rb, f, sc, P_r, P_c);
}

// AdaptPartitions_Avx2 for one filter: H_v += conj(X) * G over P partitions x R channels.  h2 != nullptr:
// also ComputeFrequencyResponse_Avx2 (max over the render channels, from 0) for every partition.
template <int R>
WAP_DEV void mc_fir_adapt_r(const McRender& rb, AecScratch& sc, float (*H_re)[kBinsPad], float (*H_im)[kBinsPad], int P,
                            const float* G_re, const float* G_im, float (*h2)[kBinsPad]) {
  const int pos = sc.s.spectra_read;
  for (int k = lane_id(); k < kBins; k += 32) {
    const float Gre = G_re[k], Gim = G_im[k];
#pragma unroll 1
    for (int p0 = 0; p0 < P; p0 += 2) {
      float Xr[2][R], Xi[2][R], Hr[2][R], Hi[2][R];
#pragma unroll
      for (int pp = 0; pp < 2; ++pp) {
        const int p = imin(p0 + pp, P - 1);
        const int row = ring_row(pos, p);
#pragma unroll
        for (int rc = 0; rc < R; ++rc) {
          Xr[pp][rc] = rb.fft_re[row][rc][k];
          Xi[pp][rc] = rb.fft_im[row][rc][k];
          Hr[pp][rc] = H_re[p * R + rc][k];
          Hi[pp][rc] = H_im[p * R + rc][k];
        }
      }
#pragma unroll
      for (int pp = 0; pp < 2; ++pp) {
        const int p = p0 + pp;
        if (p < P) {
          float m = 0.f;
#pragma unroll
          for (int rc = 0; rc < R; ++rc) {
            const float aa = Xr[pp][rc] * Gre, bb = Xi[pp][rc] * Gim, cc = Xr[pp][rc] * Gim, dd = Xi[pp][rc] * Gre;
            const float re = Hr[pp][rc] + (aa + bb), im = Hi[pp][rc] + (cc - dd);
            H_re[p * R + rc][k] = re;
            H_im[p * R + rc][k] = im;
            const float p2 = (k < 64) ? fmaf(im, im, re * re) : re * re + im * im;
            m = fmaxr(m, p2);
          }
          if (h2) h2[p][k] = m;
        }
      }
    }
  }
  __syncwarp();
}
WAP_DEV void mc_fir_adapt(const McRender& rb, AecScratch& sc, float (*H_re)[kBinsPad], float (*H_im)[kBinsPad], int P,
                          int R, const float* G_re, const float* G_im, float (*h2)[kBinsPad]) {
  if (R == 2) mc_fir_adapt_r<2>(rb, sc, H_re, H_im, P, G_re, G_im, h2);
  else mc_fir_adapt_r<1>(rb, sc, H_re, H_im, P, G_re, G_im, h2);
}

// Subtractor::Process for capture channel c (subtractor.cc:226-342).  In: cv.y, r.X2_ref / r.X2_coa (the
// spectral sums, shared by the channels).  Out: r.e_ref, r.e_coa (formed into cv.e by the caller), cv.metrics.
WAP_DEV void mc_subtractor_process_channel(Aec3State& sh, const McRender& rb, McFilters& f, McChan& ch, Aec3Scalars& cs,
                                           AecScratch& sc, McChanVec& cv, int R, bool saturated_capture) {
  const int lane = lane_id();
  Aec3Scalars& s = cs;
  AecRemoverScratch& r = sc.rm;
  __syncwarp();
  const int P_r = s.fr_current_size, P_c = s.fc_current_size;
  mc_fir_filter_both(rb, f, sc, P_r, P_c, R);
  fft_pair(sc, true, true);
  prediction_error(sc.fftA, cv.y, r.e_ref, r.s_ref);
  prediction_error(sc.fftB, cv.y, r.e_coa, r.s_coa);
  __syncwarp();
  if (lane < 5) {
    const float* p = lane == 0 ? cv.y : lane == 1 ? r.e_ref : lane == 2 ? r.e_coa : lane == 3 ? r.s_ref : r.s_coa;
    float acc = 0.f;
    for (int i = 0; i < kBlock; ++i) acc = acc + p[i] * p[i];
    cv.metrics[lane] = acc;
  } else if (lane < 7) {
    const float* p = lane == 5 ? r.s_ref : r.s_coa;
    float mx = p[0], mn = p[0];
    for (int i = 1; i < kBlock; ++i) { mx = fmaxr(mx, p[i]); mn = fminr(mn, p[i]); }
    cv.metrics[lane] = fmaxr(mx, -mn);
  }
  __syncwarp();
  const float y2 = cv.metrics[0], e2_refined = cv.metrics[1], e2_coarse = cv.metrics[2];

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
    for (int i = lane; i < kMcVParts * kBinsPad; i += 32) {
      (&f.Hr_re[0][0])[i] *= scale;
      (&f.Hr_im[0][0])[i] *= scale;
    }
    for (int i = lane; i < s.h_time_size * kBlock; i += 32) ch.h_time[i] *= scale;
    for (int i = lane; i < kBlock; i += 32) {
      r.s_ref[i] *= scale;
      r.e_ref[i] = cv.y[i] - r.s_ref[i];
    }
    __syncwarp();
  }

  stage_zero_padded_hanning(r.e_ref, sc.fftA);
  stage_zero_padded_hanning(r.e_coa, sc.fftB);
  fft_pair(sc, false, true);
  packed_to_reim(sc.fftA, r.Er_re, r.Er_im);
  packed_to_reim(sc.fftB, r.Ec_re, r.Ec_im);
  __syncwarp();
  power_spectrum(r.Ec_re, r.Ec_im, r.E2_coa);
  power_spectrum(r.Er_re, r.Er_im, r.E2_ref);
  const bool poor_excitation = render_signal_analyzer_mask(sh, r.v0);

  if (lane == 0) {
    sc.ired[1] = 1;
    if (!refined_filters_adjusted) {
      ++s.rg_call_counter;
      gain_update_current_config(sc, s.rg_cur, s.rg_old, s.rg_tgt, 5, &s.rg_config_change_counter);
      if (poor_excitation) s.rg_poor_excitation_counter = 0;
      const bool zero = (unsigned)(++s.rg_poor_excitation_counter) < (unsigned)P_r || saturated_capture ||
                        (unsigned)s.rg_call_counter <= (unsigned)P_r;
      sc.ired[1] = zero;
    }
    sc.ired[2] = s.coarse_filter_reset_hangover > 0;
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
  const int P_r2 = s.fr_current_size, P_c2 = s.fc_current_size;
  {
    const bool zero_ref = sc.ired[1] != 0, zero_coa = sc.ired[5] != 0;
    const bool disallow_leakage_diverged = sc.ired[2] != 0;
    const float leak_conv = s.rg_cur[0], leak_div = s.rg_cur[1], err_floor = s.rg_cur[2], err_ceil = s.rg_cur[3],
                noise_gate = s.rg_cur[4];
    const float rate = s.cg_cur[0], noise_gate_c = s.cg_cur[1];
    const float* Ecx_re = coarse_ok ? r.Ec_re : r.Er_re;
    const float* Ecx_im = coarse_ok ? r.Ec_im : r.Er_im;
    const int H2_size = s.H2_size;
    for (int k = lane; k < kBins; k += 32) {
      const bool masked = r.v0[k] != 0.f;
      if (!refined_filters_adjusted) {
        float erl = 0.f;  // ComputeErl
        for (int j = 0; j < H2_size; ++j) erl += ch.H2[j][k];
        float H_error = ch.H_error[k];
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
        ch.H_error[k] = H_error;
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
  // ---- refined filter: Adapt (UpdateSize, AdaptPartitions, ConstrainAndUpdateImpulseResponse), H2
  mc_zero_vparts(f.Hr_re, f.Hr_im, sc.ired[3] * R, P_r2 * R);
  for (int i = s.h_time_size * kBlock + lane; i < P_r2 * kBlock; i += 32) ch.h_time[i] = 0.f;
  __syncwarp();
  mc_fir_adapt(rb, sc, f.Hr_re, f.Hr_im, P_r2, R, r.G_re, r.G_im, ch.H2);
  const int pr = s.fr_partition_to_constrain, pc = s.fc_partition_to_constrain;
  if (!coarse_ok) {
    // refined constrain first: the coarse filter is re-seeded from the adapted refined filter
    for (int rc = 0; rc < R; ++rc) {
      fir_constrain_pair(sc, f.Hr_re[pr * R + rc], f.Hr_im[pr * R + rc], r.s_ref, nullptr, nullptr, false);
      __syncwarp();
      for (int i = lane; i < kBlock; i += 32) {
        const float h = r.s_ref[i];
        float* dst = ch.h_time + pr * kBlock + i;
        if (rc == 0 || fabsf(*dst) < fabsf(h)) *dst = h;
      }
      __syncwarp();
    }
    // coarse_filter_->SetFilter(refined size, refined H)
    const int np = imin(P_c, P_r2) * R;
    for (int i = lane; i < np * kBinsPad; i += 32) {
      (&f.Hc_re[0][0])[i] = (&f.Hr_re[0][0])[i];
      (&f.Hc_im[0][0])[i] = (&f.Hr_im[0][0])[i];
    }
    __syncwarp();
    mc_zero_vparts(f.Hc_re, f.Hc_im, sc.ired[6] * R, P_c2 * R);
    __syncwarp();
    mc_fir_adapt(rb, sc, f.Hc_re, f.Hc_im, P_c2, R, r.v1, r.v2, nullptr);
    for (int rc = 0; rc < R; ++rc) fir_constrain(sc, f.Hc_re[pc * R + rc], f.Hc_im[pc * R + rc]);
  } else {
    mc_zero_vparts(f.Hc_re, f.Hc_im, sc.ired[6] * R, P_c2 * R);
    __syncwarp();
    mc_fir_adapt(rb, sc, f.Hc_re, f.Hc_im, P_c2, R, r.v1, r.v2, nullptr);
    for (int rc = 0; rc < R; ++rc) {
      fir_constrain_pair(sc, f.Hr_re[pr * R + rc], f.Hr_im[pr * R + rc], r.s_ref, f.Hc_re[pc * R + rc],
                         f.Hc_im[pc * R + rc], true);
      __syncwarp();
      for (int i = lane; i < kBlock; i += 32) {
        const float h = r.s_ref[i];
        float* dst = ch.h_time + pr * kBlock + i;
        if (rc == 0 || fabsf(*dst) < fabsf(h)) *dst = h;
      }
      __syncwarp();
    }
  }
  // ComputeFrequencyResponse of the partition Constrain() rewrote
  for (int k = lane; k < kBins; k += 32) {
    float m = 0.f;
    for (int rc = 0; rc < R; ++rc) {
      const float re = f.Hr_re[pr * R + rc][k], im = f.Hr_im[pr * R + rc][k];
      const float p2 = (k < 64) ? fmaf(im, im, re * re) : re * re + im * im;
      m = fmaxr(m, p2);
    }
    ch.H2[pr][k] = m;
  }
  if (lane == 0) {
    s.h_time_size = P_r2;
    s.H2_size = P_r2;
    s.fr_partition_to_constrain = pr < (P_r2 - 1) ? pr + 1 : 0;
    s.fc_partition_to_constrain = pc < (P_c2 - 1) ? pc + 1 : 0;
  }
  for (int i = lane; i < kBlock; i += 32) r.e_ref[i] = clampr(r.e_ref[i], -32768.f, 32767.f);
  __syncwarp();
}

}  // namespace wap
