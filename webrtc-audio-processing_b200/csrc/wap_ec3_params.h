// Run-time parameters of the AEC3 kernels: the members of webrtc::EchoCanceller3Config
// (reference api/audio/echo_canceller3_config.h:21-275) the hot path reads, in device-friendly form.
//
// Two build modes of the kernel translation units (see build.py):
//  * WAP_EC3_RUNTIME == 0: the default EchoCanceller3Config as compile-time constants
//    (namespace ec3d below) -- the instances every default engine, and the bench, run;
//  * WAP_EC3_RUNTIME == 1: the same kernels reading an Ec3Params copy staged in the warp's shared
//    memory -- selected by the engine when it was created with a non-default config.
// Device code spells a parameter WAP_EC3(name) / WAP_EC3_ARR(name) with `sc` (AecScratch&) in scope.
#pragma once
#include <stdint.h>
#include <string.h>

namespace wap {

struct Ec3Tuning {  // Suppressor::Tuning: mask_lf, mask_hf, max_inc_factor, max_dec_factor_lf
  float lf_t, lf_s, lf_e, hf_t, hf_s, hf_e, max_inc, max_dec_lf;
};

struct Ec3Params {
  // buffering
  int excess_render_detection_interval_blocks, max_allowed_excess_render_blocks;
  // delay
  int default_delay, delay_headroom_samples, hysteresis_limit_blocks, thr_initial, thr_converged;
  float delay_estimate_smoothing, delay_estimate_smoothing_delay_found, delay_candidate_detection_threshold;
  // render_levels
  float active_render_limit, poor_excitation_render_limit;
  // filter: {leakage_converged, leakage_diverged, error_floor, error_ceil, noise_gate} / {rate, noise_gate}
  int refined_len, coarse_len, refined_initial_len, coarse_initial_len;
  float refined[5], refined_initial[5], coarse[2], coarse_initial[2];
  int config_change_duration_blocks, coarse_reset_hangover_blocks;
  float initial_state_seconds;
  // erle
  float erle_min, erle_max_l, erle_max_h;
  // ep_strength
  float default_gain, default_len, nearend_len;
  // echo_audibility
  float low_render_limit, normal_render_limit, floor_power, audibility_threshold_lf, audibility_threshold_mf,
      audibility_threshold_hf;
  // echo_model
  int noise_floor_hold;
  float min_noise_floor_power, stationary_gate_slope, noise_gate_power, noise_gate_slope;
  // suppressor
  Ec3Tuning normal_tuning, nearend_tuning;
  int last_permanent_lf_smoothing_band, last_lf_smoothing_band, last_lf_band, first_hf_band;
  float dn_enr_threshold, dn_enr_exit_threshold, dn_snr_threshold;
  int dn_hold_duration, dn_trigger_threshold;
  float hb_enr_threshold, hb_max_gain_during_echo, hb_anti_howling_activation_threshold, hb_anti_howling_gain;
  int limiting_gain_band, bands_in_limiting_gain;
  float floor_first_increase;
  // filter.high_pass_filter_echo_reference: the render frames' band 0 passes the 16 kHz high-pass filter
  // before it is queued (echo_canceller3.cc:718-720,735-737); read by k_front only
  int high_pass_filter_echo_reference;
  // delay.fixed_capture_delay_samples: BlockDelayBuffer in front of ProcessCapture (echo_canceller3.cc:
  // 796-800,902-905) and its share of the external delay (render_delay_buffer.cc:338-343); k_front only
  int fixed_capture_delay_samples;
  // suppressor.use_subband_nearend_detection + subband_nearend_detection (SubbandNearendDetector in place
  // of DominantNearendDetector, suppression_gain.cc:365-371)
  int use_subband_nearend_detection, snd_average_blocks, snd_sub1_low, snd_sub1_high, snd_sub2_low, snd_sub2_high;
  float snd_nearend_threshold, snd_snr_threshold;
  // boolean switches of the echo remover (reference defaults in ec3d below)
  int echo_can_saturate, bounded_erl, erle_onset_compensation_in_dominant_nearend, use_conservative_tail_frequency_response;
  int erle_onset_detection, clamp_quality_estimate_to_zero, clamp_quality_estimate_to_one;
  int has_clock_drift, linear_and_stable_echo_path;
  int lf_smoothing_during_initial_phase, dn_use_during_initial_phase, dn_use_unbounded_echo_spectrum, conservative_hf_suppression;
  // filter.conservative_initial_phase / enable_coarse_filter_output_usage / use_linear_filter,
  // echo_model.render_pre_window_size / render_post_window_size / model_reverb_in_nonlinear_mode,
  // suppressor.nearend_average_blocks (1..4), render_levels.render_power_gain_db as a linear amplitude gain (k_front)
  int conservative_initial_phase, enable_coarse_filter_output_usage, use_linear_filter;
  int render_pre_window_size, render_post_window_size, model_reverb_in_nonlinear_mode, nearend_average_blocks;
  float render_linear_amplitude_gain;
  // echo_audibility.use_stationarity_properties / use_stationarity_properties_at_init (EchoAudibility)
  int use_stationarity_properties, use_stationarity_properties_at_init;
  // erle.num_sections > 1: SignalDependentErleEstimator; section_boundaries_blocks_ computed on the host
  // (signal_dependent_erle_estimator.cc:46-110), num_sections + 1 entries
  int erle_num_sections, sd_boundaries[14];
  // delay.detect_pre_echo: the accumulated-error side of the matched filters and the PreEchoLagAggregator
  // (matched_filter.cc:686-687,754-770, matched_filter_lag_aggregator.cc:53-56,98-100)
  int detect_pre_echo;
  // delay.use_external_delay_estimator: no RenderDelayController; the render buffer is aligned from the
  // set_stream_delay_ms value on every block (block_processor.cc:162-187, render_delay_buffer.cc:375-384)
  int use_external_delay_estimator;
};

// The default EchoCanceller3Config, member by member (same names as Ec3Params).
namespace ec3d {
constexpr int excess_render_detection_interval_blocks = 250, max_allowed_excess_render_blocks = 8;
constexpr int default_delay = 5, delay_headroom_samples = 32, hysteresis_limit_blocks = 1;
constexpr int thr_initial = 5, thr_converged = 20;
constexpr float delay_estimate_smoothing = 0.7f, delay_estimate_smoothing_delay_found = 0.7f;
constexpr float delay_candidate_detection_threshold = 0.2f;
constexpr float active_render_limit = 100.f, poor_excitation_render_limit = 150.f;
constexpr int refined_len = 13, coarse_len = 13, refined_initial_len = 12, coarse_initial_len = 12;
#define WAP_EC3D_REFINED {0.00005f, 0.05f, 0.001f, 2.f, 20075344.f}
#define WAP_EC3D_REFINED_INITIAL {0.005f, 0.5f, 0.001f, 2.f, 20075344.f}
#define WAP_EC3D_COARSE {0.7f, 20075344.f}
#define WAP_EC3D_COARSE_INITIAL {0.9f, 20075344.f}
constexpr int config_change_duration_blocks = 250, coarse_reset_hangover_blocks = 25;
constexpr float initial_state_seconds = 2.5f;
constexpr float erle_min = 1.f, erle_max_l = 4.f, erle_max_h = 1.5f;
constexpr float default_gain = 1.f, default_len = 0.83f, nearend_len = 0.83f;
constexpr float low_render_limit = 4 * 64.f, normal_render_limit = 64.f, floor_power = 2 * 64.f;
constexpr float audibility_threshold_lf = 10.f, audibility_threshold_mf = 10.f, audibility_threshold_hf = 10.f;
constexpr int noise_floor_hold = 50;
constexpr float min_noise_floor_power = 1638400.f, stationary_gate_slope = 10.f, noise_gate_power = 27509.42f,
                noise_gate_slope = 0.3f;
#define WAP_EC3D_NORMAL_TUNING {.3f, .4f, .3f, .07f, .1f, .3f, 2.0f, 0.25f}
#define WAP_EC3D_NEAREND_TUNING {1.09f, 1.1f, .3f, .1f, .3f, .3f, 2.0f, 0.25f}
constexpr int last_permanent_lf_smoothing_band = 0, last_lf_smoothing_band = 5, last_lf_band = 5, first_hf_band = 8;
constexpr float dn_enr_threshold = .25f, dn_enr_exit_threshold = 10.f, dn_snr_threshold = 30.f;
constexpr int dn_hold_duration = 50, dn_trigger_threshold = 12;
constexpr float hb_enr_threshold = 1.f, hb_max_gain_during_echo = 1.f, hb_anti_howling_activation_threshold = 400.f,
                hb_anti_howling_gain = 1.f;
constexpr int limiting_gain_band = 16, bands_in_limiting_gain = 1;
constexpr float floor_first_increase = 0.00001f;
constexpr int high_pass_filter_echo_reference = 0, fixed_capture_delay_samples = 0;
constexpr int use_subband_nearend_detection = 0, snd_average_blocks = 1, snd_sub1_low = 1, snd_sub1_high = 1, snd_sub2_low = 1,
              snd_sub2_high = 1;
constexpr float snd_nearend_threshold = 1.f, snd_snr_threshold = 1.f;
constexpr int echo_can_saturate = 1, bounded_erl = 0, erle_onset_compensation_in_dominant_nearend = 0,
              use_conservative_tail_frequency_response = 1;
constexpr int erle_onset_detection = 1, clamp_quality_estimate_to_zero = 1, clamp_quality_estimate_to_one = 1;
constexpr int has_clock_drift = 0, linear_and_stable_echo_path = 0;
constexpr int lf_smoothing_during_initial_phase = 1, dn_use_during_initial_phase = 1, dn_use_unbounded_echo_spectrum = 1,
              conservative_hf_suppression = 0;
constexpr int conservative_initial_phase = 0, enable_coarse_filter_output_usage = 1, use_linear_filter = 1;
constexpr int render_pre_window_size = 1, render_post_window_size = 1, model_reverb_in_nonlinear_mode = 1, nearend_average_blocks = 4;
constexpr float render_linear_amplitude_gain = 1.f;
constexpr int use_stationarity_properties = 0, use_stationarity_properties_at_init = 0;
constexpr int erle_num_sections = 1;   // sd_boundaries: all zero unless erle_num_sections > 1
constexpr int detect_pre_echo = 1, use_external_delay_estimator = 0;
}  // namespace ec3d

inline Ec3Params ec3_default_params() {
  Ec3Params p{};
#define WAP_SET(n) p.n = ec3d::n
  WAP_SET(excess_render_detection_interval_blocks); WAP_SET(max_allowed_excess_render_blocks);
  WAP_SET(default_delay); WAP_SET(delay_headroom_samples); WAP_SET(hysteresis_limit_blocks);
  WAP_SET(thr_initial); WAP_SET(thr_converged);
  WAP_SET(delay_estimate_smoothing); WAP_SET(delay_estimate_smoothing_delay_found);
  WAP_SET(delay_candidate_detection_threshold);
  WAP_SET(active_render_limit); WAP_SET(poor_excitation_render_limit);
  WAP_SET(refined_len); WAP_SET(coarse_len); WAP_SET(refined_initial_len); WAP_SET(coarse_initial_len);
  const float r[5] = WAP_EC3D_REFINED, ri[5] = WAP_EC3D_REFINED_INITIAL, c[2] = WAP_EC3D_COARSE, ci[2] = WAP_EC3D_COARSE_INITIAL;
  for (int i = 0; i < 5; ++i) { p.refined[i] = r[i]; p.refined_initial[i] = ri[i]; }
  for (int i = 0; i < 2; ++i) { p.coarse[i] = c[i]; p.coarse_initial[i] = ci[i]; }
  WAP_SET(config_change_duration_blocks); WAP_SET(coarse_reset_hangover_blocks); WAP_SET(initial_state_seconds);
  WAP_SET(erle_min); WAP_SET(erle_max_l); WAP_SET(erle_max_h);
  WAP_SET(default_gain); WAP_SET(default_len); WAP_SET(nearend_len);
  WAP_SET(low_render_limit); WAP_SET(normal_render_limit); WAP_SET(floor_power);
  WAP_SET(audibility_threshold_lf); WAP_SET(audibility_threshold_mf); WAP_SET(audibility_threshold_hf);
  WAP_SET(noise_floor_hold); WAP_SET(min_noise_floor_power); WAP_SET(stationary_gate_slope);
  WAP_SET(noise_gate_power); WAP_SET(noise_gate_slope);
  const Ec3Tuning nt = WAP_EC3D_NORMAL_TUNING, et = WAP_EC3D_NEAREND_TUNING;
  p.normal_tuning = nt;
  p.nearend_tuning = et;
  WAP_SET(last_permanent_lf_smoothing_band); WAP_SET(last_lf_smoothing_band); WAP_SET(last_lf_band); WAP_SET(first_hf_band);
  WAP_SET(dn_enr_threshold); WAP_SET(dn_enr_exit_threshold); WAP_SET(dn_snr_threshold);
  WAP_SET(dn_hold_duration); WAP_SET(dn_trigger_threshold);
  WAP_SET(hb_enr_threshold); WAP_SET(hb_max_gain_during_echo); WAP_SET(hb_anti_howling_activation_threshold);
  WAP_SET(hb_anti_howling_gain);
  WAP_SET(limiting_gain_band); WAP_SET(bands_in_limiting_gain); WAP_SET(floor_first_increase);
  WAP_SET(high_pass_filter_echo_reference); WAP_SET(fixed_capture_delay_samples);
  WAP_SET(use_subband_nearend_detection); WAP_SET(snd_average_blocks); WAP_SET(snd_sub1_low); WAP_SET(snd_sub1_high);
  WAP_SET(snd_sub2_low); WAP_SET(snd_sub2_high); WAP_SET(snd_nearend_threshold); WAP_SET(snd_snr_threshold);
  WAP_SET(echo_can_saturate); WAP_SET(bounded_erl); WAP_SET(erle_onset_compensation_in_dominant_nearend);
  WAP_SET(use_conservative_tail_frequency_response); WAP_SET(erle_onset_detection);
  WAP_SET(clamp_quality_estimate_to_zero); WAP_SET(clamp_quality_estimate_to_one); WAP_SET(has_clock_drift);
  WAP_SET(linear_and_stable_echo_path); WAP_SET(lf_smoothing_during_initial_phase); WAP_SET(dn_use_during_initial_phase);
  WAP_SET(dn_use_unbounded_echo_spectrum); WAP_SET(conservative_hf_suppression);
  WAP_SET(conservative_initial_phase); WAP_SET(enable_coarse_filter_output_usage); WAP_SET(use_linear_filter);
  WAP_SET(render_pre_window_size); WAP_SET(render_post_window_size); WAP_SET(model_reverb_in_nonlinear_mode);
  WAP_SET(nearend_average_blocks); WAP_SET(render_linear_amplitude_gain);
  WAP_SET(use_stationarity_properties); WAP_SET(use_stationarity_properties_at_init);
  WAP_SET(erle_num_sections); WAP_SET(detect_pre_echo); WAP_SET(use_external_delay_estimator);
#undef WAP_SET
  return p;
}

// Every member is a 4-byte scalar: equality of the object representations is equality of the configs.
static_assert(sizeof(Ec3Params) % 4 == 0 && alignof(Ec3Params) == 4, "Ec3Params: 4-byte members only");
inline bool same_ec3_params(const Ec3Params& a, const Ec3Params& b) {
  // memcmp, not a walk over uint32_t words: reading the float members through another type is undefined
  // behaviour and g++ -O2 did reorder it (the default config then compared unequal to itself and every
  // default engine ran the run-time-parameter kernels).
  return memcmp(&a, &b, sizeof(Ec3Params)) == 0;
}

}  // namespace wap

#ifndef WAP_EC3_RUNTIME
#define WAP_EC3_RUNTIME 0
#endif
#if WAP_EC3_RUNTIME
#define WAP_EC3(name) (sc.ep.name)
#else
#define WAP_EC3(name) (::wap::ec3d::name)
#endif
