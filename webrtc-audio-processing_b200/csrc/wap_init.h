// Host-side construction of a freshly initialised StreamState (the values the
// reference's constructors / Reset() methods establish; citations per block).
#pragma once
#include <string.h>

#include "wap_ec3_params.h"
#include "wap_mc_state.h"
#include "wap_state.h"

namespace wap {

inline void init_ns_state(NsState& s) {
  memset(&s, 0, sizeof(s));
  for (int i = 0; i < kNsBinsPad; ++i) {
    s.prev_analysis_spectrum[i] = 1.f;  // noise_suppressor.cc:252
    s.wiener[i] = 1.f;                  // wiener_filter.cc:26
    s.avg_log_lrt[i] = 0.5f;            // signal_model.cc:23
  }
  for (int i = 0; i < 3 * kNsBins + 1; ++i) {
    s.q_density[i] = 0.3f;              // quantile_noise_estimator.cc:26-27
    s.q_log_quantile[i] = 8.f;
  }
  // counter_[i] = floor(200 * (i + 1) / 3) (quantile_noise_estimator.cc:29-32)
  s.q_counter[0] = 66; s.q_counter[1] = 133; s.q_counter[2] = 200;
  s.q_num_updates = 1;
  s.num_analyzed_frames = -1;
  s.histogram_analysis_counter = 500;
  s.prior_speech_prob = 0.5f;
  s.lrt = s.spectral_flatness = s.spectral_diff = 0.5f;
  s.prior_lrt = 0.5f;
  s.prior_flatness_threshold = 0.5f;
  s.prior_template_diff_threshold = 0.5f;
  s.prior_lrt_weighting = 1.f;
}

// Initial AEC3 state: what the reference's constructors and the Reset() calls
// they make leave behind (citations per group).
inline void init_aec3_state(Aec3State& a, const Ec3Params& ep) {
  memset(&a, 0, sizeof(a));
  Aec3Scalars& s = a.s;
  // BlockFramer starts with one block of zeros buffered (block_framer.cc:24-33).
  s.output_framer_len = kBlock;
  // RenderDelayBufferImpl ctor -> Reset() without an external delay
  // (render_delay_buffer.cc:118-197): indices 0, low-rate read one sub-block
  // ahead of write, total delay = default_delay (5), delay_ unset.
  s.lr_read = kSubBlock;
  s.blocks_read = kRingBlocks - ep.default_delay;
  s.spectra_read = ep.default_delay;
  s.num_api_calls_in_a_row = 1;
  s.max_observed_jitter = 1;
  // MatchedFilter / lag aggregators (matched_filter.h:156-164, matched_filter_lag_aggregator.h:61-98)
  s.mf_last_detected_best_lag_filter = -1;
  s.agg_candidate = -1;
  for (int n = 0; n < kNumMatchedFilters; ++n)
    for (int k = 0; k < kAccErrLen; ++k) a.mf_acc_err[n][k] = 1.f;
  for (int i = 0; i < 252; ++i) a.pre_hist_data[i] = -1;
  // EchoRemoverImpl (echo_remover.cc:168-176)
  s.er_refined_last_selected = 1;
  // AdaptiveFirFilter x2: initial sizes (echo_canceller3_config.h:76-120)
  s.fr_current_size = s.fr_target_size = s.fr_old_target_size = ep.refined_initial_len;
  s.fc_current_size = s.fc_target_size = s.fc_old_target_size = ep.coarse_initial_len;
  // Subtractor ctor: impulse/frequency responses sized for the longest filter (subtractor.cc:99-112)
  s.h_time_size = ep.refined_len;   // max(refined_initial, refined) = refined (Validate keeps initial <= main)
  s.H2_size = ep.refined_len;
  // RefinedFilterUpdateGain / CoarseFilterUpdateGain with the *_initial configs
  s.rg_poor_excitation_counter = 1000;
  for (int i = 0; i < 5; ++i) s.rg_cur[i] = s.rg_old[i] = s.rg_tgt[i] = ep.refined_initial[i];
  for (int i = 0; i < 2; ++i) s.cg_cur[i] = s.cg_old[i] = s.cg_tgt[i] = ep.coarse_initial[i];
  for (int k = 0; k < kBins; ++k) a.H_error[k] = 10000.f;
  // AecState and members (aec_state.h:170-305)
  s.init_state = 1;
  s.fa_gain = ep.default_gain;           // ep_strength.default_gain
  s.fa_filter_length_blocks = ep.refined_initial_len;
  s.fa_hp_size = ep.refined_len * kBlock;
  s.cfd_consistent_delay_reference = -10;
  s.tm_active_blocks_since_sane_filter = 10000;  // LegacyTransparentModeImpl (transparent_mode.cc:132-140)
  s.tm_non_converged_sequence_size = 10000;
  // ReverbDecayEstimator ctor (reverb_decay_estimator.cc:89-102): kEarlyReverbMinSizeBlocks, |default_len|
  s.rd_late_start = s.rd_late_end = 3;
  s.rd_decay = fabsf(ep.default_len);
  // ErleEstimator::Reset(true) (erle_estimator.cc:46-56, fullband_erle_estimator.cc:50-62,150-155)
  for (int k = 0; k < kBins; ++k) {
    a.erle[k] = a.erle_onset_comp[k] = a.erle_unbounded[k] = ep.erle_min;
    a.coming_onset[k] = 1;
    a.erl[k] = 1000.f;                   // ErlEstimator ctor (erl_estimator.cc:33-39)
    a.X2_noise_floor[k] = ep.min_noise_floor_power;  // ResidualEchoEstimator::Reset (residual_echo_estimator.cc:317-321)
    a.X2_noise_floor_counter[k] = ep.noise_floor_hold;
    a.cng_N2[k] = 1.0e6f;                // ComfortNoiseGenerator ctor (comfort_noise_generator.cc:106-123)
    a.last_gain[k] = 1.f;                // SuppressionGain ctor (suppression_gain.cc:351)
    a.sta_noise[k] = 10.f;               // StationarityEstimator::NoiseSpectrum::Reset (kMinNoisePower)
    a.sd_erle[k] = a.sd_erle_onset[k] = ep.erle_min;   // SignalDependentErleEstimator::Reset
  }
  for (int j = 0; j < kMaxPartitions; ++j)
    for (int b = 0; b < 8; ++b) { a.sd_estimators[j][b] = ep.erle_min; a.sd_correction[j][b] = 1.f; }
  for (int b = 0; b < 8; ++b) a.sd_erle_ref[b] = ep.erle_min;
  {
    // FastApproxLog2f(erle.min + 1e-3) (aec3_common.cc:37-52)
    const float in = ep.erle_min + 1e-3f;
    uint32_t bits;
    memcpy(&bits, &in, 4);
    float out = (float)bits;
    out *= 1.1920929e-7f;
    out -= 126.942695f;
    s.fb_erle_time_domain_log2 = out;
  }
  s.fb_max_erle_log2 = -10.f;
  s.fb_min_erle_log2 = 33.f;
  s.erl_time_domain = 1000.f;
  s.cng_has_initial = 1;
  s.cng_seed = 42;
  s.sg_initial_state = 1;
  s.sg_average_power = 32768.f * 32768.f;  // LowNoiseRenderDetector (suppression_gain.h:106)
}

// The per-capture-channel part of a freshly constructed AEC3 state (wap_mc_state.h).
inline void init_mc_chan(McChan& c, const Aec3State& a) {
  memset(&c, 0, sizeof(c));
  memcpy(c.H_error, a.H_error, sizeof(c.H_error));
  memcpy(c.erle, a.erle, sizeof(c.erle));
  memcpy(c.erle_onset_comp, a.erle_onset_comp, sizeof(c.erle_onset_comp));
  memcpy(c.erle_unbounded, a.erle_unbounded, sizeof(c.erle_unbounded));
  memcpy(c.cng_N2, a.cng_N2, sizeof(c.cng_N2));
  memcpy(c.coming_onset, a.coming_onset, sizeof(c.coming_onset));
  c.s = a.s;
}
inline void init_mc_templates(McTemplates& t, const Ec3Params& ep_mono, const Ec3Params& ep_multichannel) {
  init_aec3_state(t.aec[0], ep_mono);
  init_aec3_state(t.aec[1], ep_multichannel);
  init_mc_chan(t.chan[0], t.aec[0]);
  init_mc_chan(t.chan[1], t.aec[1]);
}

inline void init_stream_state(StreamState& st, const Ec3Params& ep = ec3_default_params()) {
  memset(&st, 0, sizeof(st));
  st.capture_output_used = 1;
  st.capture_output_used_last_frame = 1;
  st.agc2.last_scaling_factor = 1.f;  // limiter.h
  st.levels.pre_prev = st.levels.pre_target = st.levels.post_prev = st.levels.post_target = 1.f;
  st.levels.prev_pre_adjustment_gain = -1.f;
  st.levels.playout_volume = st.levels.prev_playout_volume = -1;
  init_ns_state(st.ns);
  init_aec3_state(st.aec, ep);
}

}  // namespace wap
