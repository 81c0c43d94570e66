// Per-stream persistent state of the batched APM engine, laid out for HBM.
//
// One StreamState slab per call leg, slabs contiguous in one device arena
// (wap_engine).  Inside a slab every vector of the algorithm is its own
// contiguous, 16-byte aligned array (spectra as separate re[] / im[] arrays,
// filter partitions as [partition][bin]) so that the warp that owns the stream
// reads and writes each of them as consecutive 128-byte lines: structure of
// arrays within the stream, one warp per stream across the arena.
//
// The field inventory follows SURVEY.md appendix A; each block cites the
// reference class whose members it restates.  Plain C++ POD: shared by host
// (initialisation, statistics read-back) and device code.
#pragma once

#include <stddef.h>
#include <stdint.h>
#include <string.h>

namespace wap {

constexpr int kMaxBands = 3;
constexpr int kFrame = 160;          // samples per band per 10 ms (aec3_common.h:44)
constexpr int kBlock = 64;           // AEC3 block (aec3_common.h:47)
constexpr int kBins = 65;            // kFftLengthBy2Plus1
constexpr int kBinsPad = 68;         // padded to a float4 multiple
constexpr int kNsBins = 129;         // ns_common.h:19
constexpr int kNsBinsPad = 132;
constexpr int kNsOverlap = 96;       // kFftSize - kNsFrameSize

// --- AEC3 default geometry (echo_canceller3_config.h:38-120, aec3_common.h:70-83)
constexpr int kMaxPartitions = 13;   // filter.refined/coarse.length_blocks
constexpr int kInitPartitions = 12;  // *_initial.length_blocks
constexpr int kNumMatchedFilters = 5;
constexpr int kDownSampling = 4;
constexpr int kSubBlock = kBlock / kDownSampling;            // 16
constexpr int kMfWindowSubBlocks = 32;                       // aec3_common.h:52
constexpr int kMfShiftSubBlocks = kMfWindowSubBlocks * 3 / 4;  // 24
constexpr int kMfLen = kMfWindowSubBlocks * kSubBlock;       // 512 taps
constexpr int kMfShift = kMfShiftSubBlocks * kSubBlock;      // 384
constexpr int kLowRateSize =                                 // GetDownSampledBufferSize: 2448
    kSubBlock * (kMfShiftSubBlocks * kNumMatchedFilters + kMfWindowSubBlocks + 1);
constexpr int kRingBlocks = kLowRateSize / kSubBlock + kMaxPartitions + 1;  // 167
constexpr int kMaxFilterLag = kNumMatchedFilters * kMfShift + kMfLen;       // GetMaxFilterLag: 2432
constexpr int kLagHistSize = kMaxFilterLag + 1;                             // 2433
constexpr int kPreEchoHistSize = (kLagHistSize * kDownSampling) >> 6;       // 152
constexpr int kAccErrLen = kMfLen / 4;                                      // 128

struct Biquad {
  float x0, x1, y0, y1;  // CascadedBiQuadFilter::BiQuad::{x,y} (cascaded_biquad_filter.h)
};

// ThreeBandFilterBank::{state_analysis_, state_synthesis_} (three_band_filter_bank.h:67-70)
struct ThreeBandState {
  float analysis[10][16];
  float synthesis[10][16];
};

// NoiseSuppressor::ChannelState and everything it owns (ns/noise_suppressor.h:62-76).
struct alignas(16) NsState {
  // NoiseSuppressor::ChannelState
  float analyze_mem[kNsOverlap];
  float process_mem[kNsOverlap];
  float synth_mem[kNsOverlap];
  float delay_mem[kMaxBands - 1][kNsOverlap];
  float prev_analysis_spectrum[kNsBinsPad];  // init 1
  // NoiseEstimator (ns/noise_estimator.h)
  float noise[kNsBinsPad];
  float prev_noise[kNsBinsPad];
  float conservative_noise[kNsBinsPad];
  float parametric_noise[kNsBinsPad];
  // QuantileNoiseEstimator (ns/quantile_noise_estimator.h:36-41)
  float q_density[3 * kNsBins + 1];   // init 0.3
  float q_log_quantile[3 * kNsBins + 1];  // init 8
  float q_quantile[kNsBinsPad];       // init 0
  // WienerFilter (ns/wiener_filter.h)
  float wiener[kNsBinsPad];           // init 1
  float initial_spectral_estimate[kNsBinsPad];
  float spectrum_prev_process[kNsBinsPad];
  // SpeechProbabilityEstimator / SignalModelEstimator
  float speech_prob[kNsBinsPad];
  float avg_log_lrt[kNsBinsPad];      // init 0.5
  int hist_lrt[1000];
  int hist_flatness[1000];
  int hist_diff[1000];
  // scalars
  int q_counter[3];                   // {66,133,200}
  int q_num_updates;                  // 1
  int num_analyzed_frames;            // -1 (NoiseSuppressor::num_analyzed_frames_)
  int histogram_analysis_counter;     // 500
  float white_noise_level, pink_noise_numerator, pink_noise_exp;
  float prior_speech_prob;            // 0.5
  float lrt, spectral_flatness, spectral_diff;  // SignalModel, init 0.5
  float prior_lrt;                    // PriorSignalModel, init 0.5
  float prior_flatness_threshold;     // 0.5
  float prior_template_diff_threshold;  // 0.5
  float prior_lrt_weighting;          // 1
  float prior_flatness_weighting;     // 0
  float prior_difference_weighting;   // 0
  float diff_normalization, signal_energy_sum;
  float pad_[2];
};

// ---------------------------------------------------------------- AEC3
// Scalar state of every AEC3 class (one group, staged in shared memory for the
// duration of a tick and mutated by lane 0; see dsp_aec3.cuh).  std::optional
// members are split into has_x / x.
struct Aec3Scalars {
  // FrameBlocker / BlockFramer fill levels (frame_blocker.cc, block_framer.cc)
  int render_blocker_len, capture_blocker_len, output_framer_len;
  // EchoCanceller3::saturated_microphone_signal_
  int saturated_microphone_signal;
  // BlockProcessorImpl (block_processor.cc:67-78)
  int capture_properly_started, render_properly_started, render_event;
  int bp_has_estimated_delay, bp_est_quality, bp_est_delay;  // estimated_delay_
  // RenderDelayBufferImpl (render_delay_buffer.cc:72-101)
  int blocks_write, blocks_read, spectra_write, spectra_read;  // fft ring shares the spectra indices
  int lr_write, lr_read;
  int has_delay, delay;
  int last_call_was_render, num_api_calls_in_a_row, max_observed_jitter;
  int render_activity, render_activity_counter, rb_render_activity;  // rb_: RenderBuffer::render_activity_
  int has_external_delay, external_delay, external_delay_verified;
  int min_latency_blocks, excess_render_detection_counter;
  // RenderDelayControllerImpl (render_delay_controller.cc:55-62)
  int ctl_has_delay, ctl_delay, ctl_delay_quality;
  int ctl_has_delay_samples, ctl_delay_samples, ctl_delay_samples_quality;
  int ctl_delay_change_counter, ctl_last_quality;
  // EchoPathDelayEstimator
  int est_has_old_lag, est_old_lag, est_consistent_counter;
  // MatchedFilter (matched_filter.h:156-164)
  int mf_last_detected_best_lag_filter;   // -1
  int mf_number_pre_echo_updates;
  // MatchedFilterLagAggregator
  int agg_significant_candidate_found;
  int agg_hist_data_index, agg_candidate;                     // HighestPeakAggregator (candidate -1)
  int agg_candidate_valid;                                    // agg_candidate is the argmax of the current histogram
  int pre_hist_data_index, pre_candidate, pre_number_updates; // PreEchoLagAggregator
  // ClockdriftDetector
  int cd_history[3], cd_level, cd_stability_counter;
  // EchoRemoverImpl
  int er_gain_change_hangover, er_refined_last_selected;  // init true
  // AdaptiveFirFilter x2 (refined, coarse)
  int fr_current_size, fr_target_size, fr_old_target_size, fr_size_change_counter, fr_partition_to_constrain;
  int fc_current_size, fc_target_size, fc_old_target_size, fc_size_change_counter, fc_partition_to_constrain;
  int h_time_size;                        // refined_impulse_responses_.size() / 64
  int H2_size;                            // refined_frequency_responses_.size()
  // RefinedFilterUpdateGain / CoarseFilterUpdateGain
  int rg_poor_excitation_counter, rg_call_counter, rg_config_change_counter;
  float rg_cur[5], rg_old[5], rg_tgt[5];  // leakage_converged, leakage_diverged, error_floor, error_ceil, noise_gate
  int cg_poor_excitation_counter, cg_call_counter, cg_config_change_counter;
  float cg_cur[2], cg_old[2], cg_tgt[2];  // rate, noise_gate
  // Subtractor::FilterMisadjustmentEstimator + coarse re-seed logic
  int mis_n_blocks_acum, mis_overhang;
  float mis_e2_acum, mis_y2_acum, mis_inv_misadjustment;
  int poor_coarse_filter_counter, coarse_filter_reset_hangover;
  // RenderSignalAnalyzer
  int rsa_has_narrow_peak, rsa_narrow_peak_band, rsa_narrow_peak_counter;
  // AecState
  int capture_signal_saturation;
  int strong_not_saturated_render_blocks, blocks_with_active_render;
  int init_state, init_transition_triggered, init_strong_blocks;   // InitialState
  int fd_filter_delay, fd_min_filter_delay, fd_has_external, fd_external_delay;  // FilterDelay
  int fq_usable, fq_blocks_since_reset, fq_blocks_since_start, fq_convergence_seen;  // FilteringQualityAnalyzer
  int saturated_echo;
  int soa_filter_converged;               // SubtractorOutputAnalyzer::filters_converged_[0]
  // FilterAnalyzer (+ ConsistentFilterDetector)
  int fa_blocks_since_reset, fa_region_start, fa_region_end, fa_peak_index, fa_filter_length_blocks;
  int fa_consistent_estimate, fa_filter_delay_blocks, fa_min_filter_delay_blocks;
  int fa_hp_size;                         // h_highpass_.size() in samples
  float fa_gain;
  int cfd_significant_peak, cfd_floor_low_limit, cfd_floor_high_limit;
  int cfd_consistent_counter, cfd_consistent_delay_reference;
  float cfd_floor_accum, cfd_secondary_peak;
  // LegacyTransparentModeImpl (transparent_mode.cc:222-233)
  int tm_capture_block_counter, tm_active, tm_active_blocks_since_sane_filter, tm_sane_filter_observed;
  int tm_finite_erl_recently_detected, tm_non_converged_sequence_size, tm_diverged_sequence_size;
  int tm_active_non_converged_sequence_size, tm_num_converged_blocks, tm_recent_convergence;
  int tm_strong_not_saturated_render_blocks;
  // ErleEstimator / SubbandErleEstimator / FullBandErleEstimator
  int erle_blocks_since_reset, erle_num_points;
  int fb_hold_counter, fb_has_erle_log2, fb_num_points;
  float fb_erle_time_domain_log2, fb_erle_log2, fb_inst_quality, fb_max_erle_log2, fb_min_erle_log2;
  float fb_Y2_acum, fb_E2_acum;
  int fb_has_quality; float fb_quality;   // linear_filters_qualities_[0]
  // ErlEstimator
  int erl_blocks_since_reset, erl_hold_counter_time_domain;
  float erl_time_domain;
  // ReverbFrequencyResponse
  float reverb_average_decay;
  // ComfortNoiseGenerator
  int cng_N2_counter, cng_has_initial;
  unsigned cng_seed;                      // 42
  // SuppressionGain (+ LowNoiseRenderDetector, DominantNearendDetector)
  int sg_initial_state, sg_nearend_mem_index;
  float sg_average_power;
  int dn_nearend_state, dn_trigger_counter, dn_hold_counter;
  int snd_mem_index;                      // SubbandNearendDetector's MovingAverage::mem_index_
  // EchoAudibility (echo_audibility.h:70-75) + StationarityEstimator::NoiseSpectrum::block_counter_
  int ea_has_write_prev, ea_spectrum_write_prev, ea_block_write_prev, ea_non_zero_render_seen, sta_block_counter;
  // ReverbDecayEstimator with the adaptive decay (ep_strength.default_len < 0; reverb_decay_estimator.h:95-108)
  // and its LateReverbLinearRegressor (:56-61).  fq_usable_filter: FilteringQualityAnalyzer::LinearFilterUsable()
  // before the filter.use_linear_filter gate of fq_usable.
  int fq_usable_filter;
  // RenderDelayBufferImpl::render_call_counter_ / capture_call_counter_ (render_delay_buffer.cc:94-95), kept
  // with delay.use_external_delay_estimator only (AlignFromExternalDelay reads their difference)
  int rdb_render_calls, rdb_capture_calls;
  int rd_late_start, rd_late_end, rd_block_to_analyze, rd_candidate_size, rd_region_identified;
  int rd_late_N, rd_late_n;
  float rd_late_nz, rd_late_nn, rd_late_count;
  float rd_decay, rd_tail_gain, rd_smoothing;
  // ApmStatsReporter one-slot queue (audio_processing_impl.cc:2312-2327)
  int stats_slot_full;
  float stats_erl_time_domain, stats_erle_log2;
  int stats_delay_blocks, stats_has_delay;
};

struct alignas(16) Aec3State {
  // ---- RenderDelayBuffer rings (render_delay_buffer.cc:72-101)
  float blocks[kRingBlocks][kBlock];        // BlockBuffer, band 0 / channel 0
  float fft_re[kRingBlocks][kBinsPad];      // FftBuffer
  float fft_im[kRingBlocks][kBinsPad];
  float spectra[kRingBlocks][kBinsPad];     // SpectrumBuffer
  float low_rate[kLowRateSize];             // DownsampledRenderBuffer
  // ---- MatchedFilter (matched_filter.h:156-164)
  float mf_h[kNumMatchedFilters][kMfLen];
  float mf_acc_err[kNumMatchedFilters][kAccErrLen];  // init 1
  // ---- lag aggregator histograms (matched_filter_lag_aggregator.h:61-98)
  int lag_hist[kLagHistSize + 3];
  int lag_hist_data[250 + 2];
  int pre_hist[kPreEchoHistSize];
  int pre_hist_data[250 + 2];               // init -1
  // ---- Subtractor / AdaptiveFirFilter (subtractor.h, adaptive_fir_filter.h)
  float Hr_re[kMaxPartitions][kBinsPad];    // refined filter
  float Hr_im[kMaxPartitions][kBinsPad];
  float Hc_re[kMaxPartitions][kBinsPad];    // coarse filter
  float Hc_im[kMaxPartitions][kBinsPad];
  float H2[kMaxPartitions][kBinsPad];       // refined_frequency_responses_
  float h_time[kMaxPartitions * kBlock];    // refined_impulse_responses_
  float h_highpass[kMaxPartitions * kBlock];  // FilterAnalyzer::h_highpass_
  float H_error[kBinsPad];                  // RefinedFilterUpdateGain::H_error_, init 10000
  // ---- 65-bin estimator vectors
  float erle[kBinsPad], erle_onset_comp[kBinsPad], erle_unbounded[kBinsPad];  // init 1 (min_erle)
  float accum_Y2[kBinsPad], accum_E2[kBinsPad];
  float erl[kBinsPad];                      // init 1000
  float avg_render_reverb[kBinsPad];        // AecState::avg_render_reverb_
  float echo_reverb[kBinsPad];              // ResidualEchoEstimator::echo_reverb_
  float tail_response[kBinsPad];            // ReverbFrequencyResponse
  float X2_noise_floor[kBinsPad];           // init 1638400
  float cng_Y2_smoothed[kBinsPad], cng_N2[kBinsPad], cng_N2_initial[kBinsPad];
  float last_gain[kBinsPad], last_nearend[kBinsPad], last_echo[kBinsPad];
  float nearend_mem[3][kBinsPad];           // aec3::MovingAverage memory (mem_len 4 -> 3 slots)
  float snd_mem[3][kBinsPad];               // SubbandNearendDetector::nearend_smoothers_ (nearend_average_blocks <= 4)
  // SignalDependentErleEstimator (erle.num_sections > 1; signal_dependent_erle_estimator.h:84-97), six subbands
  float sd_erle[kBinsPad], sd_erle_onset[kBinsPad];
  float sd_estimators[kMaxPartitions][8], sd_correction[kMaxPartitions][8], sd_erle_ref[8];
  int sd_num_updates[8];
  float rd_previous_gains[kMaxPartitions + 3];   // ReverbDecayEstimator::previous_gains_
  float sta_noise[kBinsPad];                // StationarityEstimator::NoiseSpectrum::noise_spectrum_, init 10 (kMinNoisePower)
  int sta_flags[kBinsPad], sta_hangovers[kBinsPad];   // stationarity_flags_, hangovers_
  int narrow_band_counters[kBinsPad];       // RenderSignalAnalyzer (63 used, index k-1)
  int erle_hold_counters[kBinsPad];
  int erl_hold_counters[kBinsPad];          // 63 used, index k-1
  int X2_noise_floor_counter[kBinsPad];     // init 50
  int accum_low_render[kBinsPad];           // SubbandErleEstimator accum_spectra_.low_render_energy
  int coming_onset[kBinsPad];               // init true
  // ---- time-domain memories
  float e_old[kBlock], y_old[kBlock], e_output_old[kBlock];
  float render_blocker[kBlock], capture_blocker[kBlock], output_framer[kBlock];
  Biquad render_decimator[4], capture_decimator[4];
  Biquad render_hpf[3];                     // RenderWriter::high_pass_filter_ (filter.high_pass_filter_echo_reference)
  int pad_hpf_[4];
  Aec3Scalars s;
};

// Hand-over between the three kernels of one tick (k_front -> k_delay -> k_echo).
// Lives in the leg's slab; only meaningful inside a tick.
struct RenderInsertRec {   // RenderDelayBufferImpl::Insert: where k_echo writes block r
  int blocks_write, spectra_write, previous_write, pad_;
};
struct CaptureBlockRec {   // BlockProcessorImpl::ProcessCapture: what EchoRemover sees for block b
  int process;             // 0 => no render data yet, block passes through
  int blocks_read, spectra_read;            // render-buffer read indices after AlignFromDelay
  int gain_change, delay_change, clock_drift;  // EchoPathVariability
  int est_has, est_delay;                   // estimated_delay_
};
struct alignas(16) TickScratch {
  int n_render_blocks, n_capture_blocks;
  int pad_[2];
  RenderInsertRec rins[3];
  CaptureBlockRec crec[3];
  float render_blocks[3][kBlock];
  float capture_blocks[3][kBlock];          // after the high-pass filter
  float cap_ds[3][kSubBlock];               // decimated capture blocks
  float capture_frame[kFrame * kMaxBands];  // full-band capture frame after the high-pass filter
};

// Extra state of a 48 kHz (three-band) AEC3 leg; lives in a second arena that only
// engines of that config class allocate.  Bands 1 and 2 only ever see gains, comfort
// noise and delays (reference aec3/suppression_filter.cc:153-183), but the render ring
// keeps them because UpperBandsGain / RenderSignalAnalyzer read GetBlock(0).
struct alignas(16) UpperBandState {
  float blocks_hi[kRingBlocks][2][kBlock];      // BlockBuffer, bands 1-2 (render_delay_buffer.cc:387-400)
  float e_output_old_hi[2][kBlock];             // SuppressionFilter::e_output_old_[1..2]
  float render_blocker_hi[2][kBlock], capture_blocker_hi[2][kBlock], output_framer_hi[2][kBlock];
  Biquad post_filter[4];                        // PostFilter (post_filter.cc:27-72), 48 kHz only
  // tick scratch (k_front -> k_echo -> k_post)
  float render_frame[3 * 160];                  // bands of this tick's render frame (k_split -> k_front)
  float render_blocks_hi[3][2][kBlock];
  float capture_blocks_hi[3][2][kBlock];
};

// Limiter of GainController2 (agc2/limiter.h:56-63, fixed_digital_level_estimator.h).
struct Agc2State {
  float filter_state_level;    // FixedDigitalLevelEstimator::filter_state_level_, init 0
  float last_scaling_factor;   // Limiter::last_scaling_factor_, init 1
  // fixed GainApplier::{last_gain_factor_, current_gain_factor_} (init: the configured gain) and a
  // pending Limiter::Reset() from GainController2::SetFixedGainDb (gain_controller2.cc:160-168)
  float gain_last, gain_current;
  int reset_limiter;
  int pad_;
};

// CaptureLevelsAdjuster (capture_levels_adjuster/{capture_levels_adjuster,audio_samples_scaler}.cc)
// and the echo-path gain-change detection around it (audio_processing_impl.cc:1316-1341).
struct LevelState {
  float pre_prev, pre_target;      // pre AudioSamplesScaler::{previous_gain_, target_gain_}
  float post_prev, post_target;    // post scaler
  float prev_pre_adjustment_gain;  // capture_.prev_pre_adjustment_gain (-1)
  int playout_volume;              // capture_.playout_volume (-1)
  int prev_playout_volume;         // capture_.prev_playout_volume (-1)
  int pad_;
};

// One call leg.
// ResidualEchoDetector (residual_echo_detector.h:60-90), the optional echo-likelihood statistic; a per-engine
// arena of its own (wap_engine_enable_echo_detector).  All-zero is the state after construction.
constexpr int kRedLookback = 650, kRedRenderBuffer = 30, kRedAggregation = 10 * 100;
struct EchoDetectorState {
  float render_buffer[kRedRenderBuffer];   // CircularBuffer
  int rb_next, rb_count;
  int frames_since_zero_buffer_size, seen_capture;   // seen_capture = !first_process_call_
  float render_power[kRedLookback], render_power_mean[kRedLookback], render_power_std_dev[kRedLookback];
  float covariance[kRedLookback];          // NormalizedCovarianceEstimator::covariance_ per lag
  int next_insertion_index;
  float render_mean, render_variance, capture_mean, capture_variance;   // MeanVarianceEstimator x 2
  float reliability, echo_likelihood;
  float mm_max; int mm_counter;            // MovingMax recent_likelihood_max_
  // capture_.stats.residual_echo_likelihood(_recent_max) and their copy in the ApmStatsReporter slot
  int stats_valid; float stats_likelihood, stats_recent_max;
  int slot_full, slot_valid; float slot_likelihood, slot_recent_max;
};

struct alignas(16) StreamState {
  Biquad hpf[3];                // HighPassFilter (capture, channel 0)
  // 1 once a capture frame has been processed.  Until then the reference may
  // still re-initialise on its first ProcessStream call (EngineConfig::
  // reinit_on_first_capture) and render audio received earlier is lost.
  int seen_capture;
  // AudioProcessingImpl::ApmCaptureState::{capture_output_used, capture_output_used_last_frame}
  // (audio_processing_impl.cc:818-838,1540-1552); both start true.
  int capture_output_used;
  int capture_output_used_last_frame;
  int pad_[1];
  Agc2State agc2;
  int pad2_[2];
  LevelState levels;
  ThreeBandState capture_bands; // AudioBuffer's SplittingFilter (48 kHz only)
  ThreeBandState render_bands;
  NsState ns;
  Aec3State aec;
  TickScratch tick;
};

// Engine-wide (config class) constants uploaded once.
struct EngineConfig {
  int sample_rate_hz;   // processing rate: 16000, 32000 or 48000
  int num_bands;        // 1 (16 kHz), 2 (32 kHz) or 3 (48 kHz)
  int aec_enabled;
  int ns_enabled;
  int hpf_enabled;
  // SuppressionParams (ns/suppression_params.cc:18-48)
  float ns_over_subtraction_factor;
  float ns_minimum_attenuating_gain;
  int ns_use_attenuation_adjustment;
  int capture_output_used;
  // ComfortNoiseGenerator::noise_floor_ = GetNoiseFloorFactor(-96.03406 dBFS)
  // (comfort_noise_generator.cc:41-45), evaluated on the host.
  float cng_noise_floor;
  // AudioProcessingImpl re-runs InitializeLocked() on the first ProcessStream
  // call when the capture format differs from the constructor's default
  // (16 kHz mono, audio_processing_impl.h:415-418) or when a submodule that
  // SubmoduleStates tracks by pointer (noise suppressor, AGC2) was created by the
  // constructor's own InitializeLocked() after UpdateActiveSubmoduleStates() ran
  // (audio_processing_impl.cc:558-559,894-925,1874-1881).  That rebuilds
  // EchoCanceller3, so render frames queued before the first capture frame are
  // dropped.  Reproduced because it changes the output of the first frames.
  int reinit_on_first_capture;
  // GainController2, default sub-configuration: fixed digital gain (DbToRatio on the host) + limiter.
  int agc2_enabled;
  float agc2_fixed_gain;
  // SubmoduleStates::CaptureMultiBandProcessingPresent (audio_processing_impl.cc:399-412): the
  // 48 kHz frame is only split into bands when a multi-band submodule is active.
  int split_bands;
  // API rate != processing rate (audio_processing_impl.cc:632-692, audio_buffer.cc:79-94): the
  // frames are resampled on the way in and out.
  int api_frame;        // samples per 10 ms at the API rate
  int resample;         // 1: API rate differs from the processing rate
  int fullband_out;     // 1: processing rate < 48 kHz output: capture_fullband_audio path (:598-611,1451-1460)
  int hpf_rate;         // rate whose high-pass coefficients are used (proc_fullband_sample_rate_hz, :1892)
  // Stereo API frames with the default pipeline (multi_channel_render / _capture off) and AEC3:
  // render is averaged to mono (render AudioBuffer has one channel), capture keeps both channels
  // until AEC3's saturation test and then continues with the first one only
  // (audio_processing_impl.cc:585-594,1365-1373); the mono result goes to both output channels.
  int channels;         // channels of the capture AudioBuffer = of the output stream (1 or 2)
  int levels_enabled;   // pre_amplifier.enabled || capture_level_adjustment.enabled
  int post_gain_enabled;  // capture_level_adjustment.enabled: kCapturePostGain is honoured
  // Stereo frames with pipeline.multi_channel_render and _capture on and AEC3: every channel is processed
  // (EchoCanceller3 with 2 capture and 1 or 2 render channels, wap_mc_state.h) -- BASELINE config 4.
  int mc;
  // Formats that differ between the three streams of a leg (audio_processing_impl.cc:527-612,632-692;
  // SURVEY 8(f)-2).  api_frame / resample describe the capture input, `channels` the capture buffer
  // (= output) channels; the capture AudioBuffer downmixes an input with more channels on the way in
  // (audio_buffer.cc:116-140,234-300: average, or the first channel with
  // pipeline.capture_downmix_method = UseFirstChannel).
  int in_channels;      // API channels of the capture input
  int out_frame;        // samples per channel and 10 ms at the output rate
  int resample_out;     // output rate != processing rate (not the capture_fullband_audio path)
  int render_frame;     // samples per channel and 10 ms of the render (reverse) stream
  int render_channels;  // API channels of the render stream (averaged to mono unless `mc`)
  int resample_render;  // render rate != processing rate
  int pre_stage;        // resample || resample_render: k_resample hands k_front processing-rate frames
  int downmix_first;    // capture downmix takes the first channel instead of the average
};

// Every member of EngineConfig is a 4-byte scalar, so the struct has no padding bytes and two
// configs are equal exactly when their object representations are (export / import of leg state).
static_assert(alignof(EngineConfig) == 4 && sizeof(EngineConfig) % 4 == 0, "EngineConfig: 4-byte members only");
inline bool same_engine_config(const EngineConfig& a, const EngineConfig& b) {
  return memcmp(&a, &b, sizeof(EngineConfig)) == 0;  // see same_ec3_params: no type-punned reads
}

}  // namespace wap
