/* wap_audio_processing.h -- C ABI of the B200 batched AudioProcessing engine.
 *
 * This is the header the reference's alternate-backend seam includes and does
 * not ship (reference modules/audio_processing/rust_audio_processing.cc:7,
 * types declared at rust_audio_processing.h:10-13).  Every entry point below
 * is bound by that seam; the citation after each one is the reference call
 * site that fixes its signature.  A build of the reference with
 * -Drust-backend=true (meson.build:201-248) links this library in place of the
 * sonora-ffi crate and webrtc::BuiltinAudioProcessingBuilder::Build returns a
 * RustAudioProcessing that forwards here
 * (api/audio/builtin_audio_processing_builder.cc:32-43).
 *
 * Extensions for the many-stream engine (not in the seam today) are marked
 * EXT: WapEngine, wap_engine_*, wap_process_streams*.
 *
 * Audio conventions are the reference's (api/audio/audio_processing.h:66-77):
 * 10 ms frames; int16 data interleaved; float data planar in [-1, 1].
 * There is no CPU fallback: every call needs a CUDA device.
 *
 * Threading contract (webrtc::AudioProcessing's render / capture split,
 * audio_processing_impl.h:199-215):
 *  - a handle made by wap_create*() may be driven by ONE render thread
 *    (wap_process_reverse_stream_*) and ONE capture thread (everything else)
 *    at the same time: the reverse call only validates and enqueues, the
 *    capture side owns the private engine; capture-side calls on one handle
 *    serialise on a per-handle lock;
 *  - every wap_engine_* / wap_process_streams* call and every per-leg setter on
 *    a leg of an engine takes that engine's lock, so calls on one engine from
 *    several threads are safe but serialise; different engines are independent.
 */
#ifndef WAP_AUDIO_PROCESSING_H_
#define WAP_AUDIO_PROCESSING_H_

#include <stdbool.h>
#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct WapAudioProcessing WapAudioProcessing; /* rust_audio_processing.h:11 */
typedef struct WapEngine WapEngine;                   /* EXT */

/* rust_audio_processing.cc:13-18 */
typedef struct WapStreamConfig {
  int sample_rate_hz;
  int32_t num_channels;
} WapStreamConfig;

/* rust_audio_processing.cc:21-40; values map onto AudioProcessing::Error
 * (api/audio/audio_processing.h:663-683).  The seam is C++ and spells the
 * enumerators WapError::None, WapDownmixMethod::UseFirstChannel,
 * WapNoiseSuppressionLevel::Low ... (rust_audio_processing.cc:23-35,376,410),
 * so C++ sees scoped enums; C sees prefixed constants of the same int32 ABI. */
#ifdef __cplusplus
enum class WapError : int32_t {
  None = 0,
  NullPointer = 1,
  Internal = 2,
  BadSampleRate = 3,
  BadNumberChannels = 4,
  BadStreamParameter = 5,
  BadDataLength = 6,
  UnsupportedConfig = 7 /* EXT: config class outside SURVEY section 8 scope */
};
enum class WapDownmixMethod : int32_t { AverageChannels = 0, UseFirstChannel = 1 };
enum class WapNoiseSuppressionLevel : int32_t { Low = 0, Moderate = 1, High = 2, VeryHigh = 3 };
enum class WapSampleFormat : int32_t { I16 = 0, F32 = 1 }; /* EXT */
#else
typedef int32_t WapError;
enum {
  WapErrorNone = 0,
  WapErrorNullPointer = 1,
  WapErrorInternal = 2,
  WapErrorBadSampleRate = 3,
  WapErrorBadNumberChannels = 4,
  WapErrorBadStreamParameter = 5,
  WapErrorBadDataLength = 6,
  WapErrorUnsupportedConfig = 7
};
typedef int32_t WapDownmixMethod;
enum { WapDownmixAverageChannels = 0, WapDownmixUseFirstChannel = 1 };
typedef int32_t WapNoiseSuppressionLevel;
enum { WapNsLow = 0, WapNsModerate = 1, WapNsHigh = 2, WapNsVeryHigh = 3 };
typedef int32_t WapSampleFormat;
enum { WapSampleI16 = 0, WapSampleF32 = 1 };
#endif

/* Flat mirror of AudioProcessing::Config (api/audio/audio_processing.h:137-376)
 * with the field set rust_audio_processing.cc:367-442 maps. */
typedef struct WapConfig {
  int pipeline_maximum_internal_processing_rate;
  bool pipeline_multi_channel_render;
  bool pipeline_multi_channel_capture;
  WapDownmixMethod pipeline_capture_downmix_method;
  bool pre_amplifier_enabled;
  float pre_amplifier_fixed_gain_factor;
  bool capture_level_adjustment_enabled;
  float capture_level_adjustment_pre_gain_factor;
  float capture_level_adjustment_post_gain_factor;
  bool analog_mic_gain_emulation_enabled;
  int analog_mic_gain_emulation_initial_level;
  bool high_pass_filter_enabled;
  bool high_pass_filter_apply_in_full_band;
  bool echo_canceller_enabled;
  bool echo_canceller_enforce_high_pass_filtering;
  bool noise_suppression_enabled;
  WapNoiseSuppressionLevel noise_suppression_level;
  bool noise_suppression_analyze_linear_aec_output_when_available;
  bool gain_controller2_enabled;
  float gain_controller2_fixed_digital_gain_db;
  bool gain_controller2_adaptive_digital_enabled;
  float gain_controller2_adaptive_digital_headroom_db;
  float gain_controller2_adaptive_digital_max_gain_db;
  float gain_controller2_adaptive_digital_initial_gain_db;
  float gain_controller2_adaptive_digital_max_gain_change_db_per_second;
  float gain_controller2_adaptive_digital_max_output_noise_level_dbfs;
  bool gain_controller2_input_volume_controller_enabled;
} WapConfig;

/* rust_audio_processing.cc:323-347 */
typedef struct WapStats {
  bool has_echo_return_loss;
  double echo_return_loss;
  bool has_echo_return_loss_enhancement;
  double echo_return_loss_enhancement;
  bool has_divergent_filter_fraction;
  double divergent_filter_fraction;
  bool has_delay_median_ms;
  int32_t delay_median_ms;
  bool has_delay_standard_deviation_ms;
  int32_t delay_standard_deviation_ms;
  bool has_residual_echo_likelihood;
  double residual_echo_likelihood;
  bool has_residual_echo_likelihood_recent_max;
  double residual_echo_likelihood_recent_max;
  bool has_delay_ms;
  int32_t delay_ms;
} WapStats;

/* ---- lifetime ---------------------------------------------------------- */
WapAudioProcessing* wap_create(void);                         /* rust_audio_processing.cc:46 */
WapAudioProcessing* wap_create_with_config(WapConfig config); /* :51 */
void wap_destroy(WapAudioProcessing* apm);                    /* :55 */
WapConfig wap_config_default(void);                           /* :365 */
WapError wap_get_config(const WapAudioProcessing* apm, WapConfig* out); /* :357 */
WapError wap_apply_config(WapAudioProcessing* apm, WapConfig config);   /* :83 */
WapError wap_initialize(WapAudioProcessing* apm, WapStreamConfig input, WapStreamConfig output,
                        WapStreamConfig reverse_input, WapStreamConfig reverse_output); /* :76-77 */

/* ---- runtime settings -------------------------------------------------- */
void wap_set_capture_output_used(WapAudioProcessing* apm, bool used);      /* :117,157 */
void wap_set_capture_pre_gain(WapAudioProcessing* apm, float gain);        /* :127 */
void wap_set_capture_post_gain(WapAudioProcessing* apm, float gain);       /* :133 */
void wap_set_capture_fixed_post_gain(WapAudioProcessing* apm, float gain_db); /* :139 */
void wap_set_playout_volume(WapAudioProcessing* apm, int volume);          /* :145 */
void wap_set_playout_audio_device(WapAudioProcessing* apm, int id, int max_volume); /* :151 */
void wap_set_stream_analog_level(WapAudioProcessing* apm, int level);      /* :276 */
int wap_recommended_stream_analog_level(const WapAudioProcessing* apm);    /* :280 */
WapError wap_set_stream_delay_ms(WapAudioProcessing* apm, int delay_ms);   /* :286 */
int wap_stream_delay_ms(const WapAudioProcessing* apm);                    /* :291 */

/* ---- per-frame processing (one call leg) -------------------------------- */
/* src_len / dest_len = frames * channels (rust_audio_processing.cc:182-191). */
WapError wap_process_stream_i16(WapAudioProcessing* apm, const int16_t* src, int32_t src_len,
                                WapStreamConfig input, WapStreamConfig output, int16_t* dest,
                                int32_t dest_len);                          /* :189-191 */
WapError wap_process_stream_f32(WapAudioProcessing* apm, const float* const* src, WapStreamConfig input,
                                WapStreamConfig output, float* const* dest); /* :203-205 */
WapError wap_process_reverse_stream_i16(WapAudioProcessing* apm, const int16_t* src, int32_t src_len,
                                        WapStreamConfig input, WapStreamConfig output, int16_t* dest,
                                        int32_t dest_len);                  /* :223-225 */
WapError wap_process_reverse_stream_f32(WapAudioProcessing* apm, const float* const* src,
                                        WapStreamConfig input, WapStreamConfig output,
                                        float* const* dest);                /* :236-238 */
WapError wap_get_statistics(const WapAudioProcessing* apm, WapStats* out);  /* :325 */

/* ---- EXT: the batched many-stream engine -------------------------------- */
/* One engine = one GPU + one config class (sample rate, channel layout,
 * enabled submodules).  Streams are slots of its HBM state arena. */
WapEngine* wap_engine_create(int cuda_device, int32_t max_streams, WapConfig config,
                             WapStreamConfig stream_format);
void wap_engine_destroy(WapEngine* engine);
/* Creates `n` call legs in the engine; handles are owned by the caller and
 * released with wap_destroy. */
WapError wap_engine_create_streams(WapEngine* engine, int32_t n, WapAudioProcessing** out_handles);
size_t wap_engine_state_bytes_per_stream(const WapEngine* engine);
/* Algorithmic HBM bytes one stream touches per 10 ms frame for this engine's
 * config class (SURVEY.md section 8(d) model; used for roofline reporting). */
double wap_engine_algorithmic_bytes_per_frame(const WapEngine* engine);

/* One 10 ms tick for `n` call legs: for leg i, ProcessReverseStream(render_i)
 * (skipped when render_frames is NULL), set_stream_delay_ms(0) semantics as
 * configured on the handle, then ProcessStream(capture_i) -> out_i.
 * Frames are packed [leg][channel-interleaved samples] for I16 and
 * [leg][channel][sample] for F32, HOST memory; copies are part of the call.
 * per_stream_err may be NULL. */
WapError wap_process_streams(WapAudioProcessing* const* handles, int32_t n, const void* render_frames,
                             const void* capture_frames, void* out_frames, WapSampleFormat fmt,
                             WapError* per_stream_err);
/* set_stream_delay_ms(delay_ms) on `n` legs at once (same clamp / warning code as the
 * single-leg call); the per-tick companion of wap_process_streams. */
WapError wap_streams_set_delay_ms(WapAudioProcessing* const* handles, int32_t n, int delay_ms);
/* Same, but the three buffers are DEVICE pointers on the engine's GPU and the
 * call only enqueues the tick on the engine's CUDA stream. */
WapError wap_process_streams_device(WapEngine* engine, WapAudioProcessing* const* handles, int32_t n,
                                    const void* d_render, const void* d_capture, void* d_out,
                                    WapSampleFormat fmt);
/* One tick is three kernels for the 16 kHz classes (k_front, k_delay, k_echo; k_resample / k_split in
 * front and k_post behind for the resampled and 48 kHz classes, counted with k_front / k_echo here).
 * While kernel timing is enabled
 * every tick records CUDA events around them on the engine's stream (and waits), so a
 * bench can attribute time and algorithmic bytes per kernel; out arrays have 3 entries. */
WapError wap_engine_enable_kernel_timing(WapEngine* engine, bool on);
int64_t wap_engine_read_kernel_timing(const WapEngine* engine, double* out_ms);
void wap_engine_algorithmic_bytes_per_kernel(const WapEngine* engine, double* out_bytes);
WapError wap_engine_synchronize(WapEngine* engine);
/* wap_process_streams on a large batch cuts the legs into ranges so the PCIe copies of one range
 * overlap the kernels of another. chunks = 0: automatic (4 ranges of whole occupancy waves from
 * 4 waves up), 1: off, 2..8: that many ranges. Results do not depend on it. */
WapError wap_engine_set_pipeline_chunks(WapEngine* engine, int32_t chunks);
/* CUDA stream the engine launches on (cudaStream_t), for event timing. */
void* wap_engine_cuda_stream(WapEngine* engine);
/* Number of kernel launches issued by the engine so far. */
int64_t wap_engine_launch_count(const WapEngine* engine);
/* 1 when the engine runs the kernel instances that read the EchoCanceller3Config parameters at run
 * time (a non-default config, or multi-channel legs), 0 for the default-config instances. */
int32_t wap_engine_uses_runtime_aec3_parameters(const WapEngine* engine);

/* Stream lifecycle: the complete state of a leg (device slabs + host-side settings) as an opaque
 * blob, e.g. to move a live call to another engine or GPU.  The importing leg must belong to an
 * engine of the same config class (else UnsupportedConfig); processing continues bit-identically. */
size_t wap_stream_state_bytes(const WapAudioProcessing* handle);
WapError wap_stream_export_state(WapAudioProcessing* handle, void* blob, size_t blob_bytes);
WapError wap_stream_import_state(WapAudioProcessing* handle, const void* blob, size_t blob_bytes);
/* Live migration: moves the leg to a free slot of `destination` -- an engine of the same config class
 * on the same or another GPU of the box -- with device-to-device copies of its state slabs
 * (cudaMemcpyPeer: NVLink between GPUs); the handle stays valid and belongs to `destination`
 * afterwards.  Neither engine may be inside a tick.  BadStreamParameter: the destination is full or
 * the handle owns a private engine; UnsupportedConfig: another config class. */
WapError wap_stream_migrate(WapAudioProcessing* handle, WapEngine* destination);

/* EXT: the residual echo detector.  The reference injects it when the instance is built
 * (AudioProcessingBuilder::SetEchoDetector(CreateEchoDetector()), api/audio/echo_detector_creator.h:21,
 * modules/audio_processing/residual_echo_detector.cc); the seam has no entry point for it.  Call on a
 * new engine before its first leg is created: every leg then reports
 * WapStats::residual_echo_likelihood / _recent_max (audio_processing_impl.cc:1499-1505).
 * UnsupportedConfig: multi-channel engines, and engines without AEC3 whose streams are resampled;
 * BadStreamParameter: legs exist. */
WapError wap_engine_enable_echo_detector(WapEngine* engine);

/* Stage taps: internal signals of one leg as of the last processed 64-sample block / 10 ms frame,
 * named after the reference's ApmDataDumper taps (modules/audio_processing/logging/
 * apm_data_dumper.h; e.g. "aec3_erle" in aec3/subband_erle_estimator.cc).  Read-only snapshot of
 * the leg's state slab; costs one small device->host copy. */
typedef struct WapStageTaps {
  float aec3_erle[65];                    /* "aec3_erle" */
  float aec3_erle_onset_compensated[65];  /* "aec3_erle_onset_compensated" */
  float aec3_erl[65];                     /* "aec3_erl" */
  float aec3_erl_time_domain;             /* "aec3_erl_time_domain" */
  float aec3_fullband_erle_log2;          /* "aec3_fullband_erle_log2" */
  float aec3_suppressor_gain[65];         /* "aec3_suppressor_gain" */
  float aec3_N2[65];                      /* "aec3_N2": comfort-noise spectrum */
  float aec3_refined_gain_H_error[65];    /* "aec3_refined_gain_H_error" */
  int32_t aec3_filter_delay;              /* "aec3_filter_delay": FilterAnalyzer::MinFilterDelayBlocks */
  int32_t aec3_min_direct_path_filter_delay;  /* AecState::MinDirectPathFilterDelay, blocks */
  int32_t aec3_render_delay_controller_buffer_delay; /* "aec3_render_delay_controller_buffer_delay", blocks (0: none) */
  int32_t aec3_usable_linear_estimate, aec3_transparent_mode, aec3_initial_state;
  int32_t aec3_echo_saturation, aec3_capture_saturation, aec3_dominant_nearend;
  float ns_noise_spectrum[129];           /* NoiseEstimator::noise_spectrum_ */
  float ns_filter[129];                   /* WienerFilter::filter_ */
  float ns_speech_probability[129];       /* SpeechProbabilityEstimator::speech_probability_ */
  float ns_prior_speech_probability;
} WapStageTaps;
WapError wap_stream_read_taps(WapAudioProcessing* handle, WapStageTaps* out);

/* ---- EXT: EchoCanceller3Config through the boundary -----------------------------------------
 * The reference injects a non-default AEC3 configuration when the instance is built
 * (BuiltinAudioProcessingBuilder::SetEchoCancellerConfig, api/audio/
 * builtin_audio_processing_builder.h:51-58); the seam has no entry point for it, so these are
 * extensions.  WapEchoCanceller3Config mirrors webrtc::EchoCanceller3Config member for member
 * (api/audio/echo_canceller3_config.h:21-275; size_t members are int32_t here).
 * wap_echo_canceller3_config_validate restates EchoCanceller3Config::Validate
 * (echo_canceller3_config.cc:101-286): it clamps in place and returns true when nothing changed.
 * Engines run the default config on compile-time constants; any other config selects kernel
 * instances that read the parameters at run time.  Members that would change the structure of
 * the engine (see wap_echo_canceller3_config_supported) must keep their default value. */
typedef struct WapEc3MaskingThresholds {
  float enr_transparent, enr_suppress, emr_transparent;
} WapEc3MaskingThresholds;
typedef struct WapEc3Tuning {
  WapEc3MaskingThresholds mask_lf, mask_hf;
  float max_inc_factor, max_dec_factor_lf;
} WapEc3Tuning;
typedef struct WapEc3AlignmentMixing {
  bool downmix, adaptive_selection;
  float activity_power_threshold;
  bool prefer_first_two_channels;
} WapEc3AlignmentMixing;
typedef struct WapEc3RefinedConfiguration {
  int32_t length_blocks;
  float leakage_converged, leakage_diverged, error_floor, error_ceil, noise_gate;
} WapEc3RefinedConfiguration;
typedef struct WapEc3CoarseConfiguration {
  int32_t length_blocks;
  float rate, noise_gate;
} WapEc3CoarseConfiguration;
typedef struct WapEc3SubbandRegion {
  int32_t low, high;
} WapEc3SubbandRegion;
typedef struct WapEchoCanceller3Config {
  struct {
    int32_t excess_render_detection_interval_blocks, max_allowed_excess_render_blocks;
  } buffering;
  struct {
    int32_t default_delay, down_sampling_factor, num_filters, delay_headroom_samples, hysteresis_limit_blocks,
        fixed_capture_delay_samples;
    float delay_estimate_smoothing, delay_estimate_smoothing_delay_found, delay_candidate_detection_threshold;
    struct {
      int32_t initial, converged;
    } delay_selection_thresholds;
    bool use_external_delay_estimator, log_warning_on_delay_changes;
    WapEc3AlignmentMixing render_alignment_mixing, capture_alignment_mixing;
    bool detect_pre_echo;
  } delay;
  struct {
    WapEc3RefinedConfiguration refined;
    WapEc3CoarseConfiguration coarse;
    WapEc3RefinedConfiguration refined_initial;
    WapEc3CoarseConfiguration coarse_initial;
    int32_t config_change_duration_blocks;
    float initial_state_seconds;
    int32_t coarse_reset_hangover_blocks;
    bool conservative_initial_phase, enable_coarse_filter_output_usage, use_linear_filter,
        high_pass_filter_echo_reference, export_linear_aec_output;
  } filter;
  struct {
    float min, max_l, max_h;
    bool onset_detection;
    int32_t num_sections;
    bool clamp_quality_estimate_to_zero, clamp_quality_estimate_to_one;
  } erle;
  struct {
    float default_gain, default_len, nearend_len;
    bool echo_can_saturate, bounded_erl, erle_onset_compensation_in_dominant_nearend,
        use_conservative_tail_frequency_response;
  } ep_strength;
  struct {
    float low_render_limit, normal_render_limit, floor_power, audibility_threshold_lf, audibility_threshold_mf,
        audibility_threshold_hf;
    bool use_stationarity_properties, use_stationarity_properties_at_init;
  } echo_audibility;
  struct {
    float active_render_limit, poor_excitation_render_limit, poor_excitation_render_limit_ds8, render_power_gain_db;
  } render_levels;
  struct {
    bool has_clock_drift, linear_and_stable_echo_path;
  } echo_removal_control;
  struct {
    int32_t noise_floor_hold;
    float min_noise_floor_power, stationary_gate_slope, noise_gate_power, noise_gate_slope;
    int32_t render_pre_window_size, render_post_window_size;
    bool model_reverb_in_nonlinear_mode;
  } echo_model;
  struct {
    float noise_floor_dbfs;
  } comfort_noise;
  struct {
    int32_t nearend_average_blocks;
    WapEc3Tuning normal_tuning, nearend_tuning;
    bool lf_smoothing_during_initial_phase;
    int32_t last_permanent_lf_smoothing_band, last_lf_smoothing_band, last_lf_band, first_hf_band;
    struct {
      float enr_threshold, enr_exit_threshold, snr_threshold;
      int32_t hold_duration, trigger_threshold;
      bool use_during_initial_phase, use_unbounded_echo_spectrum;
    } dominant_nearend_detection;
    struct {
      int32_t nearend_average_blocks;
      WapEc3SubbandRegion subband1, subband2;
      float nearend_threshold, snr_threshold;
    } subband_nearend_detection;
    bool use_subband_nearend_detection;
    struct {
      float enr_threshold, max_gain_during_echo, anti_howling_activation_threshold, anti_howling_gain;
    } high_bands_suppression;
    struct {
      int32_t limiting_gain_band, bands_in_limiting_gain;
    } high_frequency_suppression;
    float floor_first_increase;
    bool conservative_hf_suppression;
  } suppressor;
  struct {
    bool detect_stereo_content;
    float stereo_detection_threshold;
    int32_t stereo_detection_timeout_threshold_seconds;
    float stereo_detection_hysteresis_seconds;
  } multi_channel;
} WapEchoCanceller3Config;

/* EchoCanceller3Config() / ::CreateDefaultMultichannelConfig() (echo_canceller3_config.cc:288-301) */
WapEchoCanceller3Config wap_echo_canceller3_config_default(void);
WapEchoCanceller3Config wap_echo_canceller3_config_default_multichannel(void);
size_t wap_echo_canceller3_config_sizeof(void);
/* EchoCanceller3Config::Validate: clamps *config, returns true iff it was valid as given. */
bool wap_echo_canceller3_config_validate(WapEchoCanceller3Config* config);
/* None when this library can run `config` (after Validate); UnsupportedConfig when a member that
 * fixes the engine's structure differs from what is built: delay.down_sampling_factor (4),
 * delay.num_filters (5), delay.fixed_capture_delay_samples (0..5000), filter lengths (1..13 blocks, the initial ones not above the
 * final ones), filter.export_linear_aec_output (false), erle.num_sections (1 .. refined filter blocks
 * behind the delay headroom),
 * ep_strength.default_len < 0 (adaptive reverb decay) with fewer than 10 refined filter blocks,
 * echo_model.render_pre/post_window_size (0..100), suppressor.nearend_average_blocks (1..4) and the
 * same bound for the subband nearend detector.  Every other member, the boolean switches of the
 * echo remover included, is a run-time parameter.  Multi-channel engines additionally need the
 * defaults of the boolean switches, the render high-pass filter, the fixed capture delay, the subband
 * nearend detector, non-negative ep_strength lengths and erle.num_sections 1 (wap_engine_create*
 * reports it), delay.detect_pre_echo true and delay.use_external_delay_estimator false. */
WapError wap_echo_canceller3_config_supported(const WapEchoCanceller3Config* config);
/* wap_create_with_config / wap_engine_create with an injected AEC3 config
 * (BuiltinAudioProcessingBuilder::SetEchoCancellerConfig(config, multichannel_config)).
 * multichannel_config may be NULL: `config` then serves both (audio_processing_impl.cc:1928-1943; with no
 * injected config at all the multichannel one is CreateDefaultMultichannelConfig()).  It is what legs
 * with stereo frames and pipeline_multi_channel_render + _capture run once MultiChannelContentDetector
 * has seen persistent stereo content (echo_canceller3.cc:790-811, config_selector.cc:51-72).  The two
 * configs must agree in multi_channel.detect_stereo_content / stereo_detection_timeout_threshold_seconds
 * (ConfigSelector) and in comfort_noise.noise_floor_dbfs. */
WapAudioProcessing* wap_create_with_aec3_config(WapConfig config, const WapEchoCanceller3Config* aec3_config,
                                                const WapEchoCanceller3Config* aec3_multichannel_config);
WapEngine* wap_engine_create_with_aec3_config(int cuda_device, int32_t max_streams, WapConfig config,
                                              WapStreamConfig stream_format,
                                              const WapEchoCanceller3Config* aec3_config,
                                              const WapEchoCanceller3Config* aec3_multichannel_config);

/* ---- EXT: legs whose three streams have different formats ---------------------
 * The reference negotiates one format per stream (ProcessStream(src, input_config, output_config, dest),
 * ProcessReverseStream(src, input_config, ...): audio_processing_impl.cc:527-612,632-692,894-940): the
 * processing rate follows the lower of the capture input and output rates, every stream gets its own
 * PushSincResampler, a capture input with more channels than the output is downmixed on the way in
 * (average, or first channel with pipeline_capture_downmix_method).  The single-leg entry points take
 * the formats per call like the reference; a batched engine fixes them when it is created.
 * aec3_config / aec3_multichannel_config may be NULL (defaults).  Rates: multiples of 100 Hz up to
 * 96 kHz.  Refused (UnsupportedConfig), never approximated: an output of 48 kHz above the processing
 * rate whose input has another rate (capture_fullband_audio), 48 kHz AEC3 with another output rate,
 * multi-channel processing with differing formats, a rate conversion of the render pass-through output. */
WapEngine* wap_engine_create_with_formats(int cuda_device, int32_t max_streams, WapConfig config,
                                          WapStreamConfig input, WapStreamConfig output,
                                          WapStreamConfig reverse_input,
                                          const WapEchoCanceller3Config* aec3_config,
                                          const WapEchoCanceller3Config* aec3_multichannel_config);

const char* wap_version(void);

#ifdef __cplusplus
}
#endif
#endif /* WAP_AUDIO_PROCESSING_H_ */
