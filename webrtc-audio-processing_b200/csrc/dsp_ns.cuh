// Noise suppressor: one warp runs NoiseSuppressor::Analyze and ::Process for
// its stream (reference modules/audio_processing/ns/*).  Mono.
//
// Serial reductions of the reference (energy, spectral sums, variances) are
// kept in reference order: each is one left-to-right chain; independent chains
// are handed to different lanes and broadcast.
#pragma once

#include "dsp_fft.cuh"
#include "wap_dev.cuh"
#include "wap_libm.cuh"
#include "wap_state.h"

namespace wap {

// Per-warp shared-memory scratch for the NS stage.
struct NsScratch {
  float buf[256 + 8];        // extended frame / packed spectrum (+ room for two 129-term chains)
  float spec[kNsBinsPad];    // magnitude spectrum
  float prior[kNsBinsPad];   // prior SNR  (Analyze) / filter (Process)
  float post[kNsBinsPad];    // post SNR
  float tmp[kNsBinsPad];
  float red[32];             // reduction exchange
};

// One shared copy of the forward 256-point transform for its two call sites (Analyze, Process).
WAP_DEV_NOINLINE void ns_fft256_forward(float* buf) { fft256_forward(buf, lane_id()); }

// ---- ns/fast_math.cc:25-84
WAP_DEV float ns_fast_log2(float in) {
  float out = (float)__float_as_uint(in);
  out *= 1.1920929e-7f;
  out -= 126.942695f;
  return out;
}
WAP_DEV float ns_pow2(float p) {
  // reference: powf(2.f, p), the glibc routine restated (wap_libm.cuh)
  return libm_pow2f(p);
}
WAP_DEV float ns_pow_approx(float x, float p) { return ns_pow2(p * ns_fast_log2(x)); }
WAP_DEV float ns_log_approx(float x) { return ns_fast_log2(x) * 0.693147180559945309f; }
WAP_DEV float ns_exp_approx(float x) { return ns_pow_approx(10.f, x * 0.434294481903251828f); }

// `nchains` independent left-to-right sums, chain c evaluated by lane c via
// `term(c, i)`; results land in red[c] (visible to all lanes on return).
#define WAP_NS_CHAINS(red, nchains, n, TERM)                      \
  do {                                                            \
    const int c_ = lane_id();                                     \
    if (c_ < (nchains)) {                                         \
      float s_ = 0.f;                                             \
      for (int i = 0; i < (n); ++i) s_ += TERM(c_, i);            \
      (red)[c_] = s_;                                             \
    }                                                             \
    __syncwarp();                                                 \
  } while (0)

// FormExtendedFrame + ApplyFilterBankWindow (noise_suppressor.cc:78-101).
WAP_DEV void ns_form_windowed_frame(const float* frame, float* mem, float* buf) {
  const int lane = lane_id();
  #pragma unroll
  for (int i = lane; i < 256; i += 32) {
    float v = (i < kNsOverlap) ? mem[i] : frame[i - kNsOverlap];
    buf[i] = v;
  }
  __syncwarp();
  #pragma unroll
  for (int i = lane; i < kNsOverlap; i += 32) mem[i] = buf[160 + i];
  #pragma unroll
  for (int i = lane; i < 256; i += 32) {
    if (i < 96) buf[i] = kNsWindow96[i] * buf[i];
    else if (i >= 161) buf[i] = kNsWindow96[256 - i] * buf[i];
  }
  __syncwarp();
}

// ComputeMagnitudeSpectrum (noise_suppressor.cc:152-164) from the packed FFT.
WAP_DEV void ns_magnitude(const float* a, float* spec) {
  for (int i = lane_id(); i < kNsBins; i += 32) {
    float v;
    if (i == 0) v = fabsf(a[0]) + 1.f;
    else if (i == 128) v = fabsf(a[1]) + 1.f;
    else v = sqrtf(a[2 * i] * a[2 * i] + a[2 * i + 1] * a[2 * i + 1]) + 1.f;
    spec[i] = v;
  }
  __syncwarp();
}

// NoiseSuppressor::Analyze (noise_suppressor.cc:294-386) in the pieces the reference's channel loops cut it
// into: ns_analyze_prepare and ns_frame_nonzero for every channel, then (unless all channels are silent) the
// shared frame counter and ns_analyze_channel for every channel.  `frame`: band-0 samples (160) in shared memory.
// NoiseEstimator::PrepareAnalysis (noise_estimator.cc:66-69)
WAP_DEV void ns_analyze_prepare(NsState& st) {
  #pragma unroll
  for (int i = lane_id(); i < kNsBins; i += 32) st.prev_noise[i] = st.noise[i];
}
// Zero-frame detection: the reference sums v*v over memory+frame and tests
// > 0; a sum of non-negative terms is positive iff one term is.
WAP_DEV bool ns_frame_nonzero(const NsState& st, const float* frame) {
  int nz = 0;
  #pragma unroll
  for (int i = lane_id(); i < 256; i += 32) {
    const float v = (i < kNsOverlap) ? st.analyze_mem[i] : frame[i - kNsOverlap];
    nz |= (v * v > 0.f);
  }
  __syncwarp();
  return __any_sync(WAP_FULL, nz) != 0;
}
WAP_DEV void ns_analyze_channel(NsState& st, const EngineConfig& cfg, const float* frame, NsScratch& sc, int naf);

// One channel (mono legs).
WAP_DEV void ns_analyze(NsState& st, const EngineConfig& cfg, const float* frame, NsScratch& sc) {
  ns_analyze_prepare(st);
  if (!ns_frame_nonzero(st, frame)) return;
  int naf = st.num_analyzed_frames + 1;
  if (naf < 0) naf = 0;
  __syncwarp();
  if (lane_id() == 0) st.num_analyzed_frames = naf;
  ns_analyze_channel(st, cfg, frame, sc, naf);
}

// The per-channel body; naf: num_analyzed_frames_ after this frame's increment.
WAP_DEV void ns_analyze_channel(NsState& st, const EngineConfig& cfg, const float* frame, NsScratch& sc, int naf) {
  const int lane = lane_id();
  ns_form_windowed_frame(frame, st.analyze_mem, sc.buf);
  ns_fft256_forward(sc.buf);
  ns_magnitude(sc.buf, sc.spec);

  // signal_energy, signal_spectral_sum, conservative-noise average, and the
  // flatness log-sum are four independent serial chains (std::accumulate order).  Their
  // terms are formed by all lanes first, so that the four chains then run the SAME
  // instruction stream on four lanes (one pointer each) instead of four divergent loops.
  {
    const float* a = sc.buf;
    #pragma unroll
    for (int i = lane; i < kNsBins; i += 32) {
      const float re = (i == 0) ? a[0] : (i == 128 ? a[1] : a[2 * i]);
      const float im = (i == 0 || i == 128) ? 0.f : a[2 * i + 1];
      sc.prior[i] = re * re + im * im;
      sc.post[i] = st.conservative_noise[i];
      sc.tmp[i] = ns_log_approx(sc.spec[i]);   // also the input of the quantile estimator below
    }
    __syncwarp();
    if (lane < 4) {
      const float* p = lane == 0 ? sc.prior : lane == 1 ? sc.spec : lane == 2 ? sc.post : sc.tmp;
      float s = (lane == 3) ? 0.f : p[0];      // the log-sum starts at bin 1; 0.f + p[0] == p[0]
#pragma unroll 4
      for (int i = 1; i < kNsBins; ++i) s += p[i];
      sc.red[lane] = s;
    }
    __syncwarp();
  }
  float signal_energy = sc.red[0];
  signal_energy = fdiv(signal_energy, (float)kNsBins);
  const float signal_spectral_sum = sc.red[1];
  const float noise_sum = sc.red[2];
  const float flat_log_sum = sc.red[3];
  __syncwarp();

  // ---- NoiseEstimator::PreUpdate -> QuantileNoiseEstimator::Estimate
  //      (quantile_noise_estimator.cc:35-91)
  int qidx = -1;
  {
    int counter[3] = {st.q_counter[0], st.q_counter[1], st.q_counter[2]};
    int num_updates = st.q_num_updates;
    __syncwarp();
    for (int s = 0; s < 3; ++s) {
      const float one_by_counter_plus_1 = 1.f / ((float)counter[s] + 1.f);
      #pragma unroll
      for (int i = lane; i < kNsBins; i += 32) {
        const int j = s * kNsBins + i;
        float dens = st.q_density[j];
        float lq = st.q_log_quantile[j];
        const float delta = dens > 1.f ? 40.f / dens : 40.f;
        const float multiplier = delta * one_by_counter_plus_1;
        if (sc.tmp[i] > lq) lq += 0.25f * multiplier;
        else lq -= 0.75f * multiplier;
        if (fabsf(sc.tmp[i] - lq) < 0.01f) {
          dens = ((float)counter[s] * dens + 50.f) * one_by_counter_plus_1;
          st.q_density[j] = dens;
        }
        st.q_log_quantile[j] = lq;
      }
      if (counter[s] >= 200) {
        counter[s] = 0;
        if (num_updates >= 200) qidx = s * kNsBins;
      }
      ++counter[s];
    }
    if (num_updates < 200) {
      qidx = kNsBins * 2;
      ++num_updates;
    }
    __syncwarp();
    if (lane == 0) {
      st.q_counter[0] = counter[0];
      st.q_counter[1] = counter[1];
      st.q_counter[2] = counter[2];
      st.q_num_updates = num_updates;
    }
    if (qidx >= 0) {
      #pragma unroll
      for (int i = lane; i < kNsBins; i += 32) st.q_quantile[i] = ns_exp_approx(st.q_log_quantile[qidx + i]);
    }
    __syncwarp();
    #pragma unroll
    for (int i = lane; i < kNsBins; i += 32) st.noise[i] = st.q_quantile[i];
    __syncwarp();
  }

  // ---- startup parametric noise model (noise_estimator.cc:77-158)
  if (naf < 50) {
    if (lane < 4) {
      float s = 0.f;
      for (int i = 5; i < kNsBins; ++i) {
        const float log_i = kNsLogTable[i];
        if (lane == 0) s += log_i;
        else if (lane == 1) s += log_i * log_i;
        else {
          const float log_signal = ns_log_approx(sc.spec[i]);
          if (lane == 2) s += log_signal;
          else s += log_i * log_signal;
        }
      }
      sc.red[4 + lane] = s;
    }
    __syncwarp();
    const float sum_log_i = sc.red[4], sum_log_i_square = sc.red[5];
    const float sum_log_magn = sc.red[6], sum_log_i_log_magn = sc.red[7];
    float white = st.white_noise_level, pink_num = st.pink_noise_numerator, pink_exp = st.pink_noise_exp;
    __syncwarp();
    white += signal_spectral_sum * (1.f / kNsBins) * cfg.ns_over_subtraction_factor;
    const float denom = sum_log_i_square * (float)(kNsBins - 5) - sum_log_i * sum_log_i;
    float num = sum_log_i_square * sum_log_magn - sum_log_i * sum_log_i_log_magn;
    float adj = num / denom;
    adj = fmaxr(adj, 0.f);
    pink_num += adj;
    num = sum_log_i * sum_log_magn - (float)(kNsBins - 5) * sum_log_i_log_magn;
    adj = num / denom;
    adj = fmaxr(fminr(adj, 1.f), 0.f);
    pink_exp += adj;
    const float one_by_naf_plus_1 = 1.f / ((float)naf + 1.f);
    float parametric_exp = 0.f, parametric_num = 0.f;
    if (pink_exp > 0.f) {
      parametric_num = ns_exp_approx(pink_num * one_by_naf_plus_1);
      parametric_num *= (float)naf + 1.f;
      parametric_exp = pink_exp * one_by_naf_plus_1;
    }
    if (lane == 0) {
      st.white_noise_level = white;
      st.pink_noise_numerator = pink_num;
      st.pink_noise_exp = pink_exp;
    }
    #pragma unroll
    for (int i = lane; i < kNsBins; i += 32) {
      float pn;
      if (pink_exp == 0.f) {
        pn = white;
      } else {
        const float use_band = (float)(i < 5 ? 5 : i);
        pn = parametric_num / ns_pow_approx(use_band, parametric_exp);
      }
      st.parametric_noise[i] = pn;
      float n = st.noise[i];
      n *= (float)naf;
      const float t = pn * (float)(50 - naf);
      n += t * one_by_naf_plus_1;
      n *= (1.f / 50.f);
      st.noise[i] = n;
    }
    __syncwarp();
  }

  // ---- ComputeSnr (noise_suppressor.cc:167-190)
  #pragma unroll
  for (int i = lane; i < kNsBins; i += 32) {
    const float noise = st.noise[i];
    const float prev_estimate = st.prev_analysis_spectrum[i] / (st.prev_noise[i] + 0.0001f) * st.wiener[i];
    float post = 0.f;
    if (sc.spec[i] > noise) post = sc.spec[i] / (noise + 0.0001f) - 1.f;
    sc.post[i] = post;
    sc.prior[i] = 0.98f * prev_estimate + (1.f - 0.98f) * post;
  }
  __syncwarp();

  // ---- SpeechProbabilityEstimator::Update (speech_probability_estimator.cc:31-107)
  float diff_norm = st.diff_normalization;
  float energy_sum = st.signal_energy_sum;
  int hist_counter = st.histogram_analysis_counter;
  float f_lrt = st.lrt, f_flat = st.spectral_flatness, f_diff = st.spectral_diff;
  __syncwarp();
  if (naf < 200) {  // SignalModelEstimator::AdjustNormalization
    diff_norm *= (float)naf;
    diff_norm += signal_energy;
    diff_norm /= (float)(naf + 1);
  }
  // UpdateSpectralFlatness (signal_model_estimator.cc:64-91)
  {
    int zero = 0;
    for (int i = 1 + lane; i < kNsBins; i += 32) zero |= (sc.spec[i] == 0.f);
    if (__any_sync(WAP_FULL, zero)) {
      f_flat -= 0.3f * f_flat;
    } else {
      float denom = signal_spectral_sum - sc.spec[0];
      denom = denom * (1.f / kNsBins);
      const float numr = flat_log_sum * (1.f / kNsBins);
      const float spectral_tmp = ns_exp_approx(numr) / denom;
      f_flat += 0.3f * (spectral_tmp - f_flat);
    }
  }
  // ComputeSpectralDiff (signal_model_estimator.cc:29-61)
  {
    const float noise_average = noise_sum * (1.f / kNsBins);
    const float signal_average = signal_spectral_sum * (1.f / kNsBins);
    // three chains (covariance, noise variance, signal variance): terms by all lanes, then
    // the same serial loop on three lanes.  sc.tmp and sc.buf (the packed spectrum) are free
    // here; sc.prior / sc.post hold the SNRs the LRT update still needs.
    float* t0 = sc.tmp;
    float* t1 = sc.buf;
    float* t2 = sc.buf + 132;
    __syncwarp();
    #pragma unroll
    for (int i = lane; i < kNsBins; i += 32) {
      const float sd = sc.spec[i] - signal_average;
      const float nd = st.conservative_noise[i] - noise_average;
      t0[i] = sd * nd;
      t1[i] = nd * nd;
      t2[i] = sd * sd;
    }
    __syncwarp();
    if (lane < 3) {
      const float* p = lane == 0 ? t0 : lane == 1 ? t1 : t2;
      float s = 0.f;
#pragma unroll 4
      for (int i = 0; i < kNsBins; ++i) s += p[i];
      sc.red[8 + lane] = s;
    }
    __syncwarp();
    float covariance = sc.red[8], noise_variance = sc.red[9], signal_variance = sc.red[10];
    covariance *= (1.f / kNsBins);
    noise_variance *= (1.f / kNsBins);
    signal_variance *= (1.f / kNsBins);
    const float sdiff = signal_variance - (covariance * covariance) / (noise_variance + 0.0001f);
    const float spectral_diff = sdiff / (diff_norm + 0.0001f);
    f_diff += 0.3f * (spectral_diff - f_diff);
  }
  energy_sum += signal_energy;
  // Histograms (histograms.cc:27-48) / PriorSignalModelEstimator::Update
  float p_lrt = st.prior_lrt, p_flat_thr = st.prior_flatness_threshold, p_diff_thr = st.prior_template_diff_threshold;
  float p_w_lrt = st.prior_lrt_weighting, p_w_flat = st.prior_flatness_weighting, p_w_diff = st.prior_difference_weighting;
  __syncwarp();
  if (--hist_counter > 0) {
    if (lane == 0) {
      if (f_lrt < 1000 * 0.1f && f_lrt >= 0.f) ++st.hist_lrt[(int)(10.f * f_lrt)];
      if (f_flat < 1000 * 0.05f && f_flat >= 0.f) ++st.hist_flatness[(int)(f_flat * 20.f)];
      if (f_diff < 1000 * 0.1f && f_diff >= 0.f) ++st.hist_diff[(int)(f_diff * 10.f)];
    }
  } else {
    // prior_signal_model_estimator.cc:26-170 -- runs once per 500 analysed
    // frames; three serial scans over the 1000-bin histograms, one per lane.
    if (lane == 0) {
      // UpdateLrt
      float average = 0.f, average_compl = 0.f, average_squared = 0.f;
      int count = 0;
      for (int i = 0; i < 10; ++i) {
        const float bin_mid = ((float)i + 0.5f) * 0.1f;
        average += (float)st.hist_lrt[i] * bin_mid;
        count += st.hist_lrt[i];
      }
      if (count > 0) average = average / (float)count;
      for (int i = 0; i < 1000; ++i) {
        const float bin_mid = ((float)i + 0.5f) * 0.1f;
        average_squared += (float)st.hist_lrt[i] * bin_mid * bin_mid;
        average_compl += (float)st.hist_lrt[i] * bin_mid;
      }
      average_squared = average_squared * (1.f / 500.f);
      average_compl = average_compl * (1.f / 500.f);
      const bool low = average_squared - average * average_compl < 0.05f;
      sc.red[12] = low ? 1.f : fminr(1.f, fmaxr(.2f, 1.2f * average));
      sc.red[13] = low ? 1.f : 0.f;
    } else if (lane == 1 || lane == 2) {
      // FindFirstOfTwoLargestPeaks
      const int* h = (lane == 1) ? st.hist_flatness : st.hist_diff;
      const float bin_size = (lane == 1) ? 0.05f : 0.1f;
      int peak_value = 0, secondary_peak_value = 0, peak_weight = 0, secondary_peak_weight = 0;
      float peak_position = 0.f, secondary_peak_position = 0.f;
      for (int i = 0; i < 1000; ++i) {
        const float bin_mid = ((float)i + 0.5f) * bin_size;
        const int v = h[i];
        if (v > peak_value) {
          secondary_peak_value = peak_value;
          secondary_peak_weight = peak_weight;
          secondary_peak_position = peak_position;
          peak_value = v;
          peak_weight = v;
          peak_position = bin_mid;
        } else if (v > secondary_peak_value) {
          secondary_peak_value = v;
          secondary_peak_weight = v;
          secondary_peak_position = bin_mid;
        }
      }
      if ((fabsf(secondary_peak_position - peak_position) < 2 * bin_size) &&
          ((float)secondary_peak_weight > 0.5f * (float)peak_weight)) {
        peak_weight += secondary_peak_weight;
        peak_position = 0.5f * (peak_position + secondary_peak_position);
      }
      sc.red[12 + 2 * lane] = peak_position;
      sc.red[13 + 2 * lane] = (float)peak_weight;
    }
    __syncwarp();
    const bool low_lrt_fluctuations = sc.red[13] != 0.f;
    p_lrt = sc.red[12];
    const float flat_pos = sc.red[14];
    const int flat_w = (int)sc.red[15];
    const float diff_pos = sc.red[16];
    const int diff_w = (int)sc.red[17];
    const int use_spec_flat = ((float)flat_w < 0.3f * 500 || flat_pos < 0.6f) ? 0 : 1;
    const int use_spec_diff = ((float)diff_w < 0.3f * 500 || low_lrt_fluctuations) ? 0 : 1;
    p_diff_thr = 1.2f * diff_pos;
    p_diff_thr = fminr(1.f, fmaxr(0.16f, p_diff_thr));
    const float one_by_feature_sum = 1.f / (1.f + (float)use_spec_flat + (float)use_spec_diff);
    p_w_lrt = one_by_feature_sum;
    if (use_spec_flat == 1) {
      p_flat_thr = 0.9f * flat_pos;
      p_flat_thr = fminr(.95f, fmaxr(0.1f, p_flat_thr));
      p_w_flat = one_by_feature_sum;
    } else {
      p_w_flat = 0.f;
    }
    p_w_diff = (use_spec_diff == 1) ? one_by_feature_sum : 0.f;
    __syncwarp();
    for (int i = lane; i < 1000; i += 32) {
      st.hist_lrt[i] = 0;
      st.hist_flatness[i] = 0;
      st.hist_diff[i] = 0;
    }
    hist_counter = 500;
    energy_sum = fdiv(energy_sum, 500.f);
    diff_norm = 0.5f * (energy_sum + diff_norm);
    energy_sum = 0.f;
    if (lane == 0) {
      st.prior_lrt = p_lrt;
      st.prior_flatness_threshold = p_flat_thr;
      st.prior_template_diff_threshold = p_diff_thr;
      st.prior_lrt_weighting = p_w_lrt;
      st.prior_flatness_weighting = p_w_flat;
      st.prior_difference_weighting = p_w_diff;
    }
  }
  // UpdateSpectralLrt (signal_model_estimator.cc:94-118)
  #pragma unroll
  for (int i = lane; i < kNsBins; i += 32) {
    const float tmp1 = 1.f + 2.f * sc.prior[i];
    const float tmp2 = 2.f * sc.prior[i] / (tmp1 + 0.0001f);
    const float bessel_tmp = (sc.post[i] + 1.f) * tmp2;
    float l = st.avg_log_lrt[i];
    l += .5f * (bessel_tmp - ns_log_approx(tmp1) - l);
    st.avg_log_lrt[i] = l;
    sc.tmp[i] = l;
  }
  __syncwarp();
  f_lrt = serial_sum(sc.tmp, kNsBins) * (1.f / kNsBins);
  // indicator functions
  float prior_prob = st.prior_speech_prob;
  __syncwarp();
  {
    float width_prior = f_lrt < p_lrt ? 8.f : 4.f;
    const float indicator0 = (float)(0.5f * (libm_tanh((double)(width_prior * (f_lrt - p_lrt))) + 1.f));
    width_prior = f_flat > p_flat_thr ? 8.f : 4.f;
    const float indicator1 = (float)(0.5f * (libm_tanh((double)(1.f * width_prior * (p_flat_thr - f_flat))) + 1.f));
    width_prior = f_diff < p_diff_thr ? 8.f : 4.f;
    const float indicator2 = (float)(0.5f * (libm_tanh((double)(width_prior * (f_diff - p_diff_thr))) + 1.f));
    const float ind_prior = p_w_lrt * indicator0 + p_w_flat * indicator1 + p_w_diff * indicator2;
    prior_prob += 0.1f * (ind_prior - prior_prob);
    prior_prob = fmaxr(fminr(prior_prob, 1.f), 0.01f);
  }
  const float gain_prior = (1.f - prior_prob) / (prior_prob + 0.0001f);
  if (lane == 0) {
    st.diff_normalization = diff_norm;
    st.signal_energy_sum = energy_sum;
    st.histogram_analysis_counter = hist_counter;
    st.lrt = f_lrt;
    st.spectral_flatness = f_flat;
    st.spectral_diff = f_diff;
    st.prior_speech_prob = prior_prob;
  }
  #pragma unroll
  for (int i = lane; i < kNsBins; i += 32) {
    const float inv_lrt = ns_exp_approx(-sc.tmp[i]);
    const float p = 1.f / (1.f + gain_prior * inv_lrt);
    st.speech_prob[i] = p;
    sc.post[i] = p;  // reuse as probability
  }
  __syncwarp();

  // ---- NoiseEstimator::PostUpdate (noise_estimator.cc:161-205)
  #pragma unroll
  for (int i = lane; i < kNsBins; i += 32) {
    const float prob_speech = sc.post[i];
    const float prob_non_speech = 1.f - prob_speech;
    const float gamma_in = (i == 0) ? 0.9f : (sc.post[i - 1] > .2f ? .99f : 0.9f);
    const float prev = st.prev_noise[i];
    const float spec = sc.spec[i];
    const float noise_update_tmp = gamma_in * prev + (1.f - gamma_in) * (prob_non_speech * spec + prob_speech * prev);
    const float gamma = prob_speech > .2f ? .99f : 0.9f;
    if (prob_speech < .2f) {
      float c = st.conservative_noise[i];
      c += 0.05f * (spec - c);
      st.conservative_noise[i] = c;
    }
    float n;
    if (gamma == gamma_in) {
      n = noise_update_tmp;
    } else {
      n = gamma * prev + (1.f - gamma) * (prob_non_speech * spec + prob_speech * prev);
      n = fminr(n, noise_update_tmp);
    }
    st.noise[i] = n;
    st.prev_analysis_spectrum[i] = spec;
  }
  __syncwarp();
}

// NoiseSuppressor::Process (noise_suppressor.cc:388-559) in the three pieces the reference's channel loops
// cut it into (the Wiener filters, the gain adjustments and the upper-band gains of the channels are each
// reduced to their minimum in between):
//   ns_process_front   extended frame, FFT, WienerFilter::Update, ComputeUpperBandsGain
//                      -> spectrum in sc.buf, this channel's filter in sc.prior, *energy_before, *upper_band_gain
//   ns_process_filter  filter (sc.prior, possibly replaced by the minimum over the channels) applied, inverse
//                      FFT, synthesis window -> windowed frame in sc.buf; returns ComputeOverallScalingFactor
//   ns_process_finish  gain adjustment, overlap-add, upper bands delayed and scaled, clamp -> bands
// `bands`: [num_bands][160] in shared memory, processed in place.
WAP_DEV void ns_process_front(NsState& st, const EngineConfig& cfg, const float* bands, NsScratch& sc,
                              float* energy_before_out, float* upper_band_gain_out) {
  const int lane = lane_id();
  const int naf = st.num_analyzed_frames;
  ns_form_windowed_frame(bands, st.process_mem, sc.buf);
  // energies_before_filtering: serial sum over the 256 windowed samples.
  const float energy_before = serial_sum_sq_v4(sc.buf, 256);
  __syncwarp();
  ns_fft256_forward(sc.buf);
  ns_magnitude(sc.buf, sc.spec);

  // ---- WienerFilter::Update (wiener_filter.cc:33-84)
  #pragma unroll
  for (int i = lane; i < kNsBins; i += 32) {
    const float spec = sc.spec[i];
    const float noise = st.noise[i];
    float f = st.wiener[i];
    const float prev_tsa = st.spectrum_prev_process[i] / (st.prev_noise[i] + 0.0001f) * f;
    float current_tsa = 0.f;
    if (spec > noise) current_tsa = spec / (noise + 0.0001f) - 1.f;
    const float snr_prior = 0.98f * prev_tsa + (1.f - 0.98f) * current_tsa;
    f = snr_prior / (cfg.ns_over_subtraction_factor + snr_prior);
    f = fmaxr(fminr(f, 1.f), cfg.ns_minimum_attenuating_gain);
    if (naf < 50) {
      float ise = st.initial_spectral_estimate[i];
      ise += spec;
      st.initial_spectral_estimate[i] = ise;
      float filter_initial = ise - cfg.ns_over_subtraction_factor * st.parametric_noise[i];
      filter_initial /= ise + 0.0001f;
      filter_initial = fmaxr(fminr(filter_initial, 1.f), cfg.ns_minimum_attenuating_gain);
      filter_initial *= (float)(50 - naf);
      f *= (float)naf;
      f += filter_initial;
      f *= (1.f / 50.f);
    }
    st.wiener[i] = f;
    st.spectrum_prev_process[i] = spec;
    sc.prior[i] = f;
  }
  __syncwarp();

  // ---- ComputeUpperBandsGain (noise_suppressor.cc:193-242)
  float upper_band_gain = 1.f;
  if (cfg.num_bands > 1) {
    if (lane < 4) {
      float s = 0.f;
      if (lane == 0) for (int i = kNsBins - 33; i < kNsBins - 1; ++i) s += st.speech_prob[i];
      else if (lane == 1) for (int i = kNsBins - 33; i < kNsBins - 1; ++i) s += sc.prior[i];
      else if (lane == 2) for (int i = 0; i < kNsBins; ++i) s += st.prev_analysis_spectrum[i];
      else for (int i = 0; i < kNsBins; ++i) s += sc.spec[i];
      sc.red[lane] = s;
    }
    __syncwarp();
    float avg_prob_speech = sc.red[0] * (1.f / 32.f);
    const float avg_filter_gain = sc.red[1] * (1.f / 32.f);
    avg_prob_speech *= sc.red[3] / sc.red[2];
    // GCC narrows static_cast<float>(tanh(float)) to tanhf (noise_suppressor.cc:231); glibc's
    // routine restated in wap_libm.cuh
    float gain = 0.5f * (1.f + libm_tanhf(2.f * avg_prob_speech - 1.f));
    if (avg_prob_speech >= 0.5f) gain = 0.25f * gain + 0.75f * avg_filter_gain;
    else gain = 0.5f * gain + 0.5f * avg_filter_gain;
    upper_band_gain = fminr(fmaxr(gain, cfg.ns_minimum_attenuating_gain), 1.f);
    __syncwarp();
  }
  *energy_before_out = energy_before;
  *upper_band_gain_out = upper_band_gain;
}

WAP_DEV float ns_process_filter(const NsState& st, const EngineConfig& cfg, NsScratch& sc, float energy_before) {
  const int lane = lane_id();
  const int naf = st.num_analyzed_frames;
  // apply filter to the packed spectrum, inverse FFT, scale 2/256
  #pragma unroll
  for (int i = lane; i < kNsBins; i += 32) {
    const float f = sc.prior[i];
    if (i == 0) sc.buf[0] *= f;
    else if (i == 128) sc.buf[1] *= f;
    else {
      sc.buf[2 * i] *= f;
      sc.buf[2 * i + 1] *= f;
    }
  }
  __syncwarp();
  fft256_inverse(sc.buf, lane);
  #pragma unroll
  for (int i = lane; i < 256; i += 32) sc.buf[i] *= (2.f / 256.f);
  __syncwarp();
  const float energy_after = serial_sum_sq_v4(sc.buf, 256);
  __syncwarp();
  // synthesis window
  #pragma unroll
  for (int i = lane; i < 256; i += 32) {
    if (i < 96) sc.buf[i] = kNsWindow96[i] * sc.buf[i];
    else if (i >= 161) sc.buf[i] = kNsWindow96[256 - i] * sc.buf[i];
  }
  // WienerFilter::ComputeOverallScalingFactor (wiener_filter.cc:86-121)
  float gain_adjustment = 1.f;
  if (cfg.ns_use_attenuation_adjustment && naf > 200) {
    const float prior_speech_probability = st.prior_speech_prob;
    float gain = sqrtf(energy_after / (energy_before + 1.f));
    float scale_factor1 = 1.f;
    if (gain > 0.5f) {
      scale_factor1 = 1.f + 1.3f * (gain - 0.5f);
      if (gain * scale_factor1 > 1.f) scale_factor1 = 1.f / gain;
    }
    float scale_factor2 = 1.f;
    if (gain < 0.5f) {
      gain = fmaxr(gain, cfg.ns_minimum_attenuating_gain);
      scale_factor2 = 1.f - 0.3f * (0.5f - gain);
    }
    gain_adjustment = prior_speech_probability * scale_factor1 + (1.f - prior_speech_probability) * scale_factor2;
  }
  __syncwarp();
  return gain_adjustment;
}

WAP_DEV void ns_process_finish(NsState& st, const EngineConfig& cfg, float* bands, NsScratch& sc, float gain_adjustment,
                               float upper_band_gain) {
  const int lane = lane_id();
  // scale, overlap-add (noise_suppressor.cc:104-116) and clamp
  #pragma unroll
  for (int i = lane; i < 256; i += 32) sc.buf[i] = gain_adjustment * sc.buf[i];
  __syncwarp();
  #pragma unroll
  for (int i = lane; i < kFrame; i += 32) {
    float v = (i < kNsOverlap) ? st.synth_mem[i] + sc.buf[i] : sc.buf[i];
    bands[i] = fminr(fmaxr(v, -32768.f), 32767.f);
  }
  __syncwarp();
  #pragma unroll
  for (int i = lane; i < kNsOverlap; i += 32) st.synth_mem[i] = sc.buf[kFrame + i];
  // upper bands: delay by 96 samples and scale (noise_suppressor.cc:119-131,523-547)
  for (int b = 1; b < cfg.num_bands; ++b) {
    float* y = bands + b * kFrame;
    float* dm = st.delay_mem[b - 1];
    #pragma unroll
    for (int i = lane; i < kFrame; i += 32) sc.buf[i] = (i < kNsOverlap) ? dm[i] : y[i - kNsOverlap];
    __syncwarp();
    #pragma unroll
    for (int i = lane; i < kNsOverlap; i += 32) dm[i] = y[kFrame - kNsOverlap + i];
    __syncwarp();
    #pragma unroll
    for (int i = lane; i < kFrame; i += 32) {
      const float v = upper_band_gain * sc.buf[i];
      y[i] = fminr(fmaxr(v, -32768.f), 32767.f);
    }
    __syncwarp();
  }
}

// One channel (mono legs).
WAP_DEV void ns_process(NsState& st, const EngineConfig& cfg, float* bands, NsScratch& sc) {
  float energy_before, upper_band_gain;
  ns_process_front(st, cfg, bands, sc, &energy_before, &upper_band_gain);
  if (!cfg.capture_output_used) return;
  const float gain_adjustment = ns_process_filter(st, cfg, sc, energy_before);
  ns_process_finish(st, cfg, bands, sc, gain_adjustment, upper_band_gain);
}

}  // namespace wap
