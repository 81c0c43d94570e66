// Test-harness code (NOT product code): a flat C ABI over the UNMODIFIED
// reference (built by oracle/build_ref.py into oracle/_ref/libwap_ref.so) so
// that pytest / bench.py can drive webrtc::AudioProcessing and a few of the
// reference's internal DSP classes through ctypes.  Only tests/, smoke() and
// bench.py's cpu_baseline / --impl reference legs may load it.
//
// Mirrors the loop of examples/run-offline.cpp:45-63 (reference), with the
// rate / submodule selection BASELINE.json's configs name.
#include <pthread.h>
#include <sched.h>

#include <atomic>
#include <chrono>
#include <cstdint>
#include <cstring>
#include <memory>
#include <optional>
#include <string>
#include <thread>
#include <vector>

#include "api/audio/audio_processing.h"
#include "api/audio/builtin_audio_processing_builder.h"
#include "api/make_ref_counted.h"
#include "modules/audio_processing/residual_echo_detector.h"
#include "api/audio/echo_canceller3_config.h"
#include "api/environment/environment_factory.h"
#include "api/scoped_refptr.h"
#include "common_audio/third_party/ooura/fft_size_128/ooura_fft.h"
#include "common_audio/third_party/ooura/fft_size_256/fft4g.h"
#include "modules/audio_processing/aec3/aec3_common.h"
#include "modules/audio_processing/high_pass_filter.h"
#include "modules/audio_processing/logging/apm_data_dumper.h"
#include "modules/audio_processing/ns/noise_suppressor.h"
#include "modules/audio_processing/ns/ns_config.h"
#include "modules/audio_processing/three_band_filter_bank.h"
#include "modules/audio_processing/audio_buffer.h"
#include "rtc_base/cpu_info.h"
#include "rtc_base/system/arch.h"
#include "system_wrappers/include/denormal_disabler.h"

using webrtc::AudioProcessing;

namespace {

struct RefApm {
  webrtc::scoped_refptr<AudioProcessing> apm;
  int stream_delay_ms = 0;   // passed to set_stream_delay_ms before every ProcessStream (ref_apm_set_stream_delay_ms)
};

AudioProcessing::Config MakeConfig(int aec, int ns, int ns_level, int max_rate,
                                   int hpf, int mc_render, int mc_capture) {
  AudioProcessing::Config c;
  c.echo_canceller.enabled = aec != 0;
  c.noise_suppression.enabled = ns != 0;
  c.noise_suppression.level =
      static_cast<AudioProcessing::Config::NoiseSuppression::Level>(ns_level);
  c.pipeline.maximum_internal_processing_rate = max_rate;
  c.pipeline.multi_channel_render = mc_render != 0;
  c.pipeline.multi_channel_capture = mc_capture != 0;
  c.high_pass_filter.enabled = hpf != 0;
  return c;
}

}  // namespace

extern "C" {

// ---------------------------------------------------------------- full APM
void* ref_apm_create(int aec, int ns, int ns_level, int max_rate, int hpf,
                     int mc_render, int mc_capture) {
  auto* h = new RefApm;
  webrtc::Environment env = webrtc::CreateEnvironment();
  h->apm = webrtc::BuiltinAudioProcessingBuilder(
               MakeConfig(aec, ns, ns_level, max_rate, hpf, mc_render, mc_capture))
               .Build(env);
  return h;
}

// Same with GainController2 in its default sub-configuration (fixed digital gain + limiter).
void* ref_apm_create_agc2(int aec, int ns, int ns_level, int max_rate, int agc2, float fixed_gain_db) {
  auto* h = new RefApm;
  webrtc::Environment env = webrtc::CreateEnvironment();
  AudioProcessing::Config c = MakeConfig(aec, ns, ns_level, max_rate, 0, 0, 0);
  c.gain_controller2.enabled = agc2 != 0;
  c.gain_controller2.fixed_digital.gain_db = fixed_gain_db;
  h->apm = webrtc::BuiltinAudioProcessingBuilder(c).Build(env);
  return h;
}

// Everything above plus pre-amplifier / capture level adjustment, configured at construction.
void* ref_apm_create_levels(int aec, int ns, int ns_level, int max_rate, int hpf, int agc2, float fixed_gain_db,
                            int pre_amp_enabled, float pre_amp_gain, int cla_enabled, float cla_pre, float cla_post) {
  auto* h = new RefApm;
  webrtc::Environment env = webrtc::CreateEnvironment();
  AudioProcessing::Config c = MakeConfig(aec, ns, ns_level, max_rate, hpf, 0, 0);
  c.gain_controller2.enabled = agc2 != 0;
  c.gain_controller2.fixed_digital.gain_db = fixed_gain_db;
  c.pre_amplifier.enabled = pre_amp_enabled != 0;
  c.pre_amplifier.fixed_gain_factor = pre_amp_gain;
  c.capture_level_adjustment.enabled = cla_enabled != 0;
  c.capture_level_adjustment.pre_gain_factor = cla_pre;
  c.capture_level_adjustment.post_gain_factor = cla_post;
  h->apm = webrtc::BuiltinAudioProcessingBuilder(c).Build(env);
  return h;
}

void ref_apm_destroy(void* p) { delete static_cast<RefApm*>(p); }

// ---- generic "key=value;key=value" configuration (tests of ApplyConfig, EchoCanceller3Config
// injection through BuiltinAudioProcessingBuilder::SetEchoCancellerConfig and the multi-channel
// pipeline).  Unknown keys return null / -1 so that a typo in a test cannot pass silently.
}  // extern "C"
namespace {
bool SetEc3(webrtc::EchoCanceller3Config& c, const std::string& k, double v) {
#define F(path) if (k == #path) { c.path = static_cast<decltype(c.path)>(v); return true; }
  F(buffering.excess_render_detection_interval_blocks) F(buffering.max_allowed_excess_render_blocks)
  F(delay.default_delay) F(delay.down_sampling_factor) F(delay.num_filters) F(delay.delay_headroom_samples)
  F(delay.hysteresis_limit_blocks) F(delay.fixed_capture_delay_samples) F(delay.delay_estimate_smoothing)
  F(delay.delay_estimate_smoothing_delay_found) F(delay.delay_candidate_detection_threshold)
  F(delay.delay_selection_thresholds.initial) F(delay.delay_selection_thresholds.converged)
  F(delay.use_external_delay_estimator) F(delay.log_warning_on_delay_changes)
  F(delay.render_alignment_mixing.downmix) F(delay.render_alignment_mixing.adaptive_selection)
  F(delay.render_alignment_mixing.activity_power_threshold) F(delay.render_alignment_mixing.prefer_first_two_channels)
  F(delay.capture_alignment_mixing.downmix) F(delay.capture_alignment_mixing.adaptive_selection)
  F(delay.capture_alignment_mixing.activity_power_threshold) F(delay.capture_alignment_mixing.prefer_first_two_channels)
  F(delay.detect_pre_echo)
  F(filter.refined.length_blocks) F(filter.refined.leakage_converged) F(filter.refined.leakage_diverged)
  F(filter.refined.error_floor) F(filter.refined.error_ceil) F(filter.refined.noise_gate)
  F(filter.coarse.length_blocks) F(filter.coarse.rate) F(filter.coarse.noise_gate)
  F(filter.refined_initial.length_blocks) F(filter.refined_initial.leakage_converged) F(filter.refined_initial.leakage_diverged)
  F(filter.refined_initial.error_floor) F(filter.refined_initial.error_ceil) F(filter.refined_initial.noise_gate)
  F(filter.coarse_initial.length_blocks) F(filter.coarse_initial.rate) F(filter.coarse_initial.noise_gate)
  F(filter.config_change_duration_blocks) F(filter.initial_state_seconds) F(filter.coarse_reset_hangover_blocks)
  F(filter.conservative_initial_phase) F(filter.enable_coarse_filter_output_usage) F(filter.use_linear_filter)
  F(filter.high_pass_filter_echo_reference) F(filter.export_linear_aec_output)
  F(erle.min) F(erle.max_l) F(erle.max_h) F(erle.onset_detection) F(erle.num_sections)
  F(erle.clamp_quality_estimate_to_zero) F(erle.clamp_quality_estimate_to_one)
  F(ep_strength.default_gain) F(ep_strength.default_len) F(ep_strength.nearend_len) F(ep_strength.echo_can_saturate)
  F(ep_strength.bounded_erl) F(ep_strength.erle_onset_compensation_in_dominant_nearend)
  F(ep_strength.use_conservative_tail_frequency_response)
  F(echo_audibility.low_render_limit) F(echo_audibility.normal_render_limit) F(echo_audibility.floor_power)
  F(echo_audibility.audibility_threshold_lf) F(echo_audibility.audibility_threshold_mf) F(echo_audibility.audibility_threshold_hf)
  F(echo_audibility.use_stationarity_properties) F(echo_audibility.use_stationarity_properties_at_init)
  F(render_levels.active_render_limit) F(render_levels.poor_excitation_render_limit)
  F(render_levels.poor_excitation_render_limit_ds8) F(render_levels.render_power_gain_db)
  F(echo_removal_control.has_clock_drift) F(echo_removal_control.linear_and_stable_echo_path)
  F(echo_model.noise_floor_hold) F(echo_model.min_noise_floor_power) F(echo_model.stationary_gate_slope)
  F(echo_model.noise_gate_power) F(echo_model.noise_gate_slope) F(echo_model.render_pre_window_size)
  F(echo_model.render_post_window_size) F(echo_model.model_reverb_in_nonlinear_mode)
  F(comfort_noise.noise_floor_dbfs)
  F(suppressor.nearend_average_blocks)
  F(suppressor.normal_tuning.mask_lf.enr_transparent) F(suppressor.normal_tuning.mask_lf.enr_suppress) F(suppressor.normal_tuning.mask_lf.emr_transparent)
  F(suppressor.normal_tuning.mask_hf.enr_transparent) F(suppressor.normal_tuning.mask_hf.enr_suppress) F(suppressor.normal_tuning.mask_hf.emr_transparent)
  F(suppressor.normal_tuning.max_inc_factor) F(suppressor.normal_tuning.max_dec_factor_lf)
  F(suppressor.nearend_tuning.mask_lf.enr_transparent) F(suppressor.nearend_tuning.mask_lf.enr_suppress) F(suppressor.nearend_tuning.mask_lf.emr_transparent)
  F(suppressor.nearend_tuning.mask_hf.enr_transparent) F(suppressor.nearend_tuning.mask_hf.enr_suppress) F(suppressor.nearend_tuning.mask_hf.emr_transparent)
  F(suppressor.nearend_tuning.max_inc_factor) F(suppressor.nearend_tuning.max_dec_factor_lf)
  F(suppressor.lf_smoothing_during_initial_phase) F(suppressor.last_permanent_lf_smoothing_band)
  F(suppressor.last_lf_smoothing_band) F(suppressor.last_lf_band) F(suppressor.first_hf_band)
  F(suppressor.dominant_nearend_detection.enr_threshold) F(suppressor.dominant_nearend_detection.enr_exit_threshold)
  F(suppressor.dominant_nearend_detection.snr_threshold) F(suppressor.dominant_nearend_detection.hold_duration)
  F(suppressor.dominant_nearend_detection.trigger_threshold) F(suppressor.dominant_nearend_detection.use_during_initial_phase)
  F(suppressor.dominant_nearend_detection.use_unbounded_echo_spectrum)
  F(suppressor.subband_nearend_detection.nearend_average_blocks)
  F(suppressor.subband_nearend_detection.subband1.low) F(suppressor.subband_nearend_detection.subband1.high)
  F(suppressor.subband_nearend_detection.subband2.low) F(suppressor.subband_nearend_detection.subband2.high)
  F(suppressor.subband_nearend_detection.nearend_threshold) F(suppressor.subband_nearend_detection.snr_threshold)
  F(suppressor.use_subband_nearend_detection)
  F(suppressor.high_bands_suppression.enr_threshold) F(suppressor.high_bands_suppression.max_gain_during_echo)
  F(suppressor.high_bands_suppression.anti_howling_activation_threshold) F(suppressor.high_bands_suppression.anti_howling_gain)
  F(suppressor.high_frequency_suppression.limiting_gain_band) F(suppressor.high_frequency_suppression.bands_in_limiting_gain)
  F(suppressor.floor_first_increase) F(suppressor.conservative_hf_suppression)
  F(multi_channel.detect_stereo_content) F(multi_channel.stereo_detection_threshold)
  F(multi_channel.stereo_detection_timeout_threshold_seconds) F(multi_channel.stereo_detection_hysteresis_seconds)
#undef F
  return false;
}

struct KvConfig {
  AudioProcessing::Config apm;
  webrtc::EchoCanceller3Config ec3;
  std::optional<webrtc::EchoCanceller3Config> ec3mc;
  bool has_ec3 = false;
  bool echo_detector = false;   // AudioProcessingBuilder::SetEchoDetector(CreateEchoDetector())
};

bool ParseKv(const char* text, KvConfig* out) {
  AudioProcessing::Config& c = out->apm;
  std::string s(text ? text : "");
  size_t pos = 0;
  while (pos < s.size()) {
    size_t end = s.find(';', pos);
    if (end == std::string::npos) end = s.size();
    const std::string item = s.substr(pos, end - pos);
    pos = end + 1;
    if (item.empty()) continue;
    const size_t eq = item.find('=');
    if (eq == std::string::npos) return false;
    const std::string k = item.substr(0, eq);
    const double v = atof(item.c_str() + eq + 1);
    if (k == "echo_detector") out->echo_detector = v != 0;
    else if (k == "aec") c.echo_canceller.enabled = v != 0;
    else if (k == "aec_enforce_hpf") c.echo_canceller.enforce_high_pass_filtering = v != 0;
    else if (k == "ns") c.noise_suppression.enabled = v != 0;
    else if (k == "ns_level") c.noise_suppression.level = static_cast<AudioProcessing::Config::NoiseSuppression::Level>((int)v);
    else if (k == "max_rate") c.pipeline.maximum_internal_processing_rate = (int)v;
    else if (k == "mc_render") c.pipeline.multi_channel_render = v != 0;
    else if (k == "mc_capture") c.pipeline.multi_channel_capture = v != 0;
    else if (k == "downmix") c.pipeline.capture_downmix_method = static_cast<AudioProcessing::Config::Pipeline::DownmixMethod>((int)v);
    else if (k == "hpf") c.high_pass_filter.enabled = v != 0;
    else if (k == "hpf_full_band") c.high_pass_filter.apply_in_full_band = v != 0;
    else if (k == "agc2") c.gain_controller2.enabled = v != 0;
    else if (k == "agc2_gain_db") c.gain_controller2.fixed_digital.gain_db = (float)v;
    else if (k == "agc2_adaptive") c.gain_controller2.adaptive_digital.enabled = v != 0;
    else if (k == "pre_amp") c.pre_amplifier.enabled = v != 0;
    else if (k == "pre_amp_gain") c.pre_amplifier.fixed_gain_factor = (float)v;
    else if (k == "cla") c.capture_level_adjustment.enabled = v != 0;
    else if (k == "cla_pre") c.capture_level_adjustment.pre_gain_factor = (float)v;
    else if (k == "cla_post") c.capture_level_adjustment.post_gain_factor = (float)v;
    else if (k == "ec3mc_default") { if (v != 0) out->ec3mc = webrtc::EchoCanceller3Config::CreateDefaultMultichannelConfig(); out->has_ec3 = true; }
    else if (k.rfind("ec3mc.", 0) == 0) {
      if (!out->ec3mc) out->ec3mc = webrtc::EchoCanceller3Config::CreateDefaultMultichannelConfig();
      out->has_ec3 = true;
      if (!SetEc3(*out->ec3mc, k.substr(6), v)) return false;
    } else if (k.rfind("ec3.", 0) == 0) {
      out->has_ec3 = true;
      if (!SetEc3(out->ec3, k.substr(4), v)) return false;
    } else return false;
  }
  return true;
}
}  // namespace
extern "C" {

void* ref_apm_create_kv(const char* text) {
  KvConfig kv;
  kv.apm = MakeConfig(1, 1, 1, 48000, 0, 0, 0);
  if (!ParseKv(text, &kv)) return nullptr;
  auto* h = new RefApm;
  webrtc::Environment env = webrtc::CreateEnvironment();
  webrtc::BuiltinAudioProcessingBuilder b(kv.apm);
  if (kv.has_ec3) b.SetEchoCancellerConfig(kv.ec3, kv.ec3mc);
  if (kv.echo_detector) b.SetEchoDetector(webrtc::make_ref_counted<webrtc::ResidualEchoDetector>());   // = CreateEchoDetector(), api/audio/echo_detector_creator.cc:19-21
  h->apm = b.Build(env);
  return h;
}

// AudioProcessing::ApplyConfig with the current config updated by the given keys (APM keys only).
int ref_apm_apply_kv(void* p, const char* text) {
  auto* h = static_cast<RefApm*>(p);
  KvConfig kv;
  kv.apm = h->apm->GetConfig();
  if (!ParseKv(text, &kv) || kv.has_ec3 || kv.echo_detector) return -1;
  h->apm->ApplyConfig(kv.apm);
  return 0;
}

// EchoCanceller3Config::Validate on the config the keys describe; writes the (possibly clamped)
// value of `probe` back.  Returns 1 when the config was valid as given, 0 when it was changed.
int ref_ec3_validate_kv(const char* text, const char* probe, double* probe_out) {
  KvConfig kv;
  if (!ParseKv(text, &kv)) return -1;
  const bool ok = webrtc::EchoCanceller3Config::Validate(&kv.ec3);
  if (probe && probe_out) {
    // read back through the setter table: find the value v with SetEc3(copy, probe, v) leaving it unchanged
    // is not possible generically, so expose a few representative fields
    const std::string k(probe);
    if (k == "filter.refined.length_blocks") *probe_out = kv.ec3.filter.refined.length_blocks;
    else if (k == "delay.down_sampling_factor") *probe_out = kv.ec3.delay.down_sampling_factor;
    else if (k == "erle.min") *probe_out = kv.ec3.erle.min;
    else if (k == "erle.max_l") *probe_out = kv.ec3.erle.max_l;
    else if (k == "suppressor.normal_tuning.max_inc_factor") *probe_out = kv.ec3.suppressor.normal_tuning.max_inc_factor;
    else if (k == "filter.coarse.rate") *probe_out = kv.ec3.filter.coarse.rate;
    else if (k == "delay.default_delay") *probe_out = kv.ec3.delay.default_delay;
    else if (k == "filter.initial_state_seconds") *probe_out = kv.ec3.filter.initial_state_seconds;
    else return -2;
  }
  return ok ? 1 : 0;
}


// One 10 ms tick on interleaved int16 frames: render then capture, exactly as
// examples/run-offline.cpp:58-59 (+ set_stream_delay_ms(0)).
int ref_apm_tick_i16(void* p, int rate, int render_ch, int capture_ch,
                     const int16_t* render, const int16_t* capture, int16_t* out,
                     int16_t* render_out) {
  auto* h = static_cast<RefApm*>(p);
  webrtc::StreamConfig rc(rate, render_ch), cc(rate, capture_ch);
  std::vector<int16_t> scratch;
  if (!render_out) {
    scratch.resize(rc.num_frames() * render_ch);
    render_out = scratch.data();
  }
  int e1 = 0;
  if (render) e1 = h->apm->ProcessReverseStream(render, rc, rc, render_out);
  h->apm->set_stream_delay_ms(h->stream_delay_ms);
  int e2 = h->apm->ProcessStream(capture, cc, cc, out);
  return e1 ? e1 : e2;
}

// Same with planar float [-1,1] (channel-major: [ch][frame]).
int ref_apm_tick_f32(void* p, int rate, int render_ch, int capture_ch,
                     const float* render, const float* capture, float* out) {
  auto* h = static_cast<RefApm*>(p);
  webrtc::StreamConfig rc(rate, render_ch), cc(rate, capture_ch);
  const int n = rate / 100;
  std::vector<const float*> rp(render_ch), cp(capture_ch);
  std::vector<float*> op(capture_ch), rop(render_ch);
  std::vector<float> rscratch(n * render_ch);
  for (int i = 0; i < render_ch; ++i) {
    rp[i] = render ? render + i * n : nullptr;
    rop[i] = rscratch.data() + i * n;
  }
  for (int i = 0; i < capture_ch; ++i) {
    cp[i] = capture + i * n;
    op[i] = out + i * n;
  }
  int e1 = 0;
  if (render) e1 = h->apm->ProcessReverseStream(rp.data(), rc, rc, rop.data());
  h->apm->set_stream_delay_ms(h->stream_delay_ms);
  int e2 = h->apm->ProcessStream(cp.data(), cc, cc, op.data());
  return e1 ? e1 : e2;
}

// One tick with a format of its own per stream: render (rr Hz, rc channels), capture
// input (ir, ic), capture output (orate, oc).  Interleaved int16.
int ref_apm_tick_fmt_i16(void* p, int rr, int rc, int ir, int ic, int orate, int oc, const int16_t* render,
                         const int16_t* capture, int16_t* out) {
  auto* h = static_cast<RefApm*>(p);
  webrtc::StreamConfig rcfg(rr, rc), icfg(ir, ic), ocfg(orate, oc);
  std::vector<int16_t> scratch(rcfg.num_frames() * rc);
  int e1 = 0;
  if (render) e1 = h->apm->ProcessReverseStream(render, rcfg, rcfg, scratch.data());
  h->apm->set_stream_delay_ms(h->stream_delay_ms);
  int e2 = h->apm->ProcessStream(capture, icfg, ocfg, out);
  return e1 ? e1 : e2;
}
// Same with planar float [-1,1] ([ch][frame] per stream).
int ref_apm_tick_fmt_f32(void* p, int rr, int rc, int ir, int ic, int orate, int oc, const float* render,
                         const float* capture, float* out) {
  auto* h = static_cast<RefApm*>(p);
  webrtc::StreamConfig rcfg(rr, rc), icfg(ir, ic), ocfg(orate, oc);
  const int nr = rr / 100, ni = ir / 100, no = orate / 100;
  std::vector<const float*> rp(rc), cp(ic);
  std::vector<float*> op(oc), rop(rc);
  std::vector<float> rscratch((size_t)nr * rc);
  for (int i = 0; i < rc; ++i) {
    rp[i] = render ? render + (size_t)i * nr : nullptr;
    rop[i] = rscratch.data() + (size_t)i * nr;
  }
  for (int i = 0; i < ic; ++i) cp[i] = capture + (size_t)i * ni;
  for (int i = 0; i < oc; ++i) op[i] = out + (size_t)i * no;
  int e1 = 0;
  if (render) e1 = h->apm->ProcessReverseStream(rp.data(), rcfg, rcfg, rop.data());
  h->apm->set_stream_delay_ms(h->stream_delay_ms);
  int e2 = h->apm->ProcessStream(cp.data(), icfg, ocfg, op.data());
  return e1 ? e1 : e2;
}
// ProcessReverseStream alone, float, with an output format of its own: the render pass-through output.
int ref_apm_reverse_f32(void* p, int rr, int rc, int orr, int orc, const float* render, float* render_out) {
  auto* h = static_cast<RefApm*>(p);
  webrtc::StreamConfig rcfg(rr, rc), ocfg(orr, orc);
  const int nr = rr / 100, no = orr / 100;
  std::vector<const float*> rp(rc);
  std::vector<float*> rop(orc);
  for (int i = 0; i < rc; ++i) rp[i] = render + (size_t)i * nr;
  for (int i = 0; i < orc; ++i) rop[i] = render_out + (size_t)i * no;
  return h->apm->ProcessReverseStream(rp.data(), rcfg, ocfg, rop.data());
}

// The runtime settings behind the wap_set_capture_*_gain / wap_set_playout_volume entry points.
void ref_apm_set_pre_gain(void* p, float g) {
  static_cast<RefApm*>(p)->apm->SetRuntimeSetting(AudioProcessing::RuntimeSetting::CreateCapturePreGain(g));
}
void ref_apm_set_post_gain(void* p, float g) {
  static_cast<RefApm*>(p)->apm->SetRuntimeSetting(AudioProcessing::RuntimeSetting::CreateCapturePostGain(g));
}
void ref_apm_set_fixed_post_gain(void* p, float db) {
  static_cast<RefApm*>(p)->apm->SetRuntimeSetting(AudioProcessing::RuntimeSetting::CreateCaptureFixedPostGain(db));
}
void ref_apm_set_playout_volume(void* p, int v) {
  static_cast<RefApm*>(p)->apm->SetRuntimeSetting(AudioProcessing::RuntimeSetting::CreatePlayoutVolumeChange(v));
}

void ref_apm_set_capture_output_used(void* p, int used) {
  static_cast<RefApm*>(p)->apm->set_output_will_be_muted(!used);
}

// stats: [has_erl, erl, has_erle, erle, has_delay, delay_ms]
void ref_apm_stats(void* p, float* out6) {
  auto* h = static_cast<RefApm*>(p);
  webrtc::AudioProcessingStats s = h->apm->GetStatistics();
  out6[0] = s.echo_return_loss.has_value();
  out6[1] = s.echo_return_loss.value_or(0.0);
  out6[2] = s.echo_return_loss_enhancement.has_value();
  out6[3] = s.echo_return_loss_enhancement.value_or(0.0);
  out6[4] = s.delay_ms.has_value();
  out6[5] = s.delay_ms.value_or(0);
}

void ref_apm_set_stream_delay_ms(void* p, int ms) { static_cast<RefApm*>(p)->stream_delay_ms = ms; }

// residual echo detector statistics: [has_likelihood, likelihood, has_recent_max, recent_max]
void ref_apm_stats_echo_detector(void* p, double* out4) {
  auto* h = static_cast<RefApm*>(p);
  webrtc::AudioProcessingStats s = h->apm->GetStatistics();
  out4[0] = s.residual_echo_likelihood.has_value();
  out4[1] = s.residual_echo_likelihood.value_or(0.0);
  out4[2] = s.residual_echo_likelihood_recent_max.has_value();
  out4[3] = s.residual_echo_likelihood_recent_max.value_or(0.0);
}

// Run nframes ticks of int16 audio laid out [frame][sample*ch]; stats every
// `stats_every` frames into stats_out (6 floats each) when non-null.
int ref_apm_run_i16(void* p, int rate, int render_ch, int capture_ch, int nframes,
                    const int16_t* render, const int16_t* capture, int16_t* out,
                    int stats_every, float* stats_out) {
  const int n = rate / 100;
  int err = 0, k = 0;
  for (int f = 0; f < nframes; ++f) {
    int e = ref_apm_tick_i16(p, rate, render_ch, capture_ch,
                             render ? render + (size_t)f * n * render_ch : nullptr,
                             capture + (size_t)f * n * capture_ch,
                             out + (size_t)f * n * capture_ch, nullptr);
    if (e && !err) err = e;
    if (stats_out && stats_every > 0 && (f + 1) % stats_every == 0)
      ref_apm_stats(p, stats_out + 6 * k++);
  }
  return err;
}

// CPU baseline: `threads` pinned workers, each owning
// streams t, t+T, ...; every stream is its own AudioProcessing instance fed
// nframes ticks (the first `warm` untimed).  The audio for stream s is
// render/capture + s*stride (int16, [frame][sample]) so callers can share one
// buffer (stride 0) or give each stream its own.  Returns timed wall seconds.
double ref_apm_bench(int aec, int ns, int ns_level, int rate, int streams,
                     int threads, int warm, int nframes, const int16_t* render,
                     const int16_t* capture, size_t stride) {
  const int n = rate / 100;
  std::vector<void*> h(streams);
  for (auto& x : h) x = ref_apm_create(aec, ns, ns_level, 48000, 0, 0, 0);
  std::atomic<int> ready{0};
  std::atomic<bool> go{false};
  std::vector<double> secs(threads, 0.0);
  std::vector<std::thread> th;
  for (int t = 0; t < threads; ++t) {
    th.emplace_back([&, t] {
      cpu_set_t set;
      CPU_ZERO(&set);
      CPU_SET(t % std::thread::hardware_concurrency(), &set);
      pthread_setaffinity_np(pthread_self(), sizeof(set), &set);
      std::vector<int16_t> out(n), ro(n);
      for (int s = t; s < streams; s += threads)
        for (int f = 0; f < warm; ++f)
          ref_apm_tick_i16(h[s], rate, 1, 1, render + s * stride + (size_t)f * n,
                           capture + s * stride + (size_t)f * n, out.data(), ro.data());
      ready++;
      while (!go.load()) std::this_thread::yield();
      auto t0 = std::chrono::steady_clock::now();
      for (int s = t; s < streams; s += threads)
        for (int f = warm; f < nframes; ++f)
          ref_apm_tick_i16(h[s], rate, 1, 1, render + s * stride + (size_t)f * n,
                           capture + s * stride + (size_t)f * n, out.data(), ro.data());
      secs[t] = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
    });
  }
  while (ready.load() < threads) std::this_thread::yield();
  go = true;
  double mx = 0;
  for (int t = 0; t < threads; ++t) {
    th[t].join();
    if (secs[t] > mx) mx = secs[t];
  }
  for (auto x : h) ref_apm_destroy(x);
  return mx;
}

// `kv`: configuration text as for ref_apm_create_kv (e.g. "aec=1;ns=0;mc_render=1;mc_capture=1"); `ch`: channels
// of the interleaved render / capture frames.
double ref_apm_bench_kv(const char* kv, int rate, int ch, int streams,
                        int threads, int warm, int nframes, const int16_t* render,
                        const int16_t* capture, size_t stride) {
  const int n = rate / 100 * ch;
  std::vector<void*> h(streams);
  for (auto& x : h) {
    x = ref_apm_create_kv(kv);
    if (!x) return -1.0;
  }
  std::atomic<int> ready{0};
  std::atomic<bool> go{false};
  std::vector<double> secs(threads, 0.0);
  std::vector<std::thread> th;
  for (int t = 0; t < threads; ++t) {
    th.emplace_back([&, t] {
      cpu_set_t set;
      CPU_ZERO(&set);
      CPU_SET(t % std::thread::hardware_concurrency(), &set);
      pthread_setaffinity_np(pthread_self(), sizeof(set), &set);
      std::vector<int16_t> out(n), ro(n);
      for (int s = t; s < streams; s += threads)
        for (int f = 0; f < warm; ++f)
          ref_apm_tick_i16(h[s], rate, ch, ch, render + s * stride + (size_t)f * n,
                           capture + s * stride + (size_t)f * n, out.data(), ro.data());
      ready++;
      while (!go.load()) std::this_thread::yield();
      auto t0 = std::chrono::steady_clock::now();
      for (int s = t; s < streams; s += threads)
        for (int f = warm; f < nframes; ++f)
          ref_apm_tick_i16(h[s], rate, ch, ch, render + s * stride + (size_t)f * n,
                           capture + s * stride + (size_t)f * n, out.data(), ro.data());
      secs[t] = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
    });
  }
  while (ready.load() < threads) std::this_thread::yield();
  go = true;
  double mx = 0;
  for (int t = 0; t < threads; ++t) {
    th[t].join();
    if (secs[t] > mx) mx = secs[t];
  }
  for (auto x : h) ref_apm_destroy(x);
  return mx;
}

// ------------------------------------------------------------ stage level
// Ooura 128-point real FFT as AEC3 uses it (reference
// common_audio/third_party/ooura/fft_size_128/ooura_fft.cc:334-349); the
// SSE2 bodies are selected when the stubbed cpu_info reports SSE2.
void ref_fft128(float* a, int inverse) {
  static webrtc::OouraFft sse2(true), plain(false);
  webrtc::OouraFft& f =
      webrtc::cpu_info::Supports(webrtc::cpu_info::ISA::kSSE2) ? sse2 : plain;
  if (inverse) f.InverseFft(a); else f.Fft(a);
}

// fft4g 256-point real FFT as NS uses it (reference ns/ns_fft.cc:22-67).
void ref_rdft256(float* a, int isgn) {
  static size_t ip[2 + 16] = {0};  // sqrt(128)+2
  static float w[128];
  static bool init = false;
  if (!init) {
    ip[0] = 0;
    float tmp[256] = {0};
    webrtc::WebRtc_rdft(256, 1, tmp, ip, w);
    init = true;
  }
  webrtc::WebRtc_rdft(256, isgn, a, ip, w);
}

void* ref_hpf_create(int rate, int channels) {
  return new webrtc::HighPassFilter(rate, channels);
}
void ref_hpf_destroy(void* p) { delete static_cast<webrtc::HighPassFilter*>(p); }
// data: [ch][n] planar, in place (reference high_pass_filter.cc:98-106).
void ref_hpf_process(void* p, float* data, int channels, int n) {
  std::vector<std::vector<float>> v(channels);
  for (int c = 0; c < channels; ++c) v[c].assign(data + c * n, data + (c + 1) * n);
  static_cast<webrtc::HighPassFilter*>(p)->Process(&v);
  for (int c = 0; c < channels; ++c) std::memcpy(data + c * n, v[c].data(), n * sizeof(float));
}

void* ref_3band_create() { return new webrtc::ThreeBandFilterBank(); }
void ref_3band_destroy(void* p) { delete static_cast<webrtc::ThreeBandFilterBank*>(p); }
// in[480] -> out[3][160]  (reference three_band_filter_bank.cc:178-225)
void ref_3band_analysis(void* p, const float* in, float* out) {
  webrtc::ArrayView<const float, 480> iv(in, 480);
  std::array<webrtc::ArrayView<float>, 3> ov = {
      webrtc::ArrayView<float>(out, 160), webrtc::ArrayView<float>(out + 160, 160),
      webrtc::ArrayView<float>(out + 320, 160)};
  static_cast<webrtc::ThreeBandFilterBank*>(p)->Analysis(iv, ov);
}
// in[3][160] -> out[480]  (reference three_band_filter_bank.cc:233-278)
void ref_3band_synthesis(void* p, const float* in, float* out) {
  float* m = const_cast<float*>(in);
  std::array<webrtc::ArrayView<float>, 3> iv = {
      webrtc::ArrayView<float>(m, 160), webrtc::ArrayView<float>(m + 160, 160),
      webrtc::ArrayView<float>(m + 320, 160)};
  webrtc::ArrayView<float, 480> ov(out, 480);
  static_cast<webrtc::ThreeBandFilterBank*>(p)->Synthesis(iv, ov);
}

// Stage taps: with the -DWEBRTC_APM_DEBUG_DUMP=1 variant (libwap_ref_dump.so) every
// ApmDataDumper::DumpRaw call of the reference writes <dir>/<name>_<inst>-<reinit>.dat.
int ref_dump_activate(const char* dir) {
  webrtc::ApmDataDumper::SetOutputDirectory(dir);
  webrtc::ApmDataDumper::SetActivated(true);
  return webrtc::ApmDataDumper::IsAvailable() ? 1 : 0;
}

int ref_isa_level() {
  using webrtc::cpu_info::ISA;
  return webrtc::cpu_info::Supports(ISA::kAVX2) ? 2
         : webrtc::cpu_info::Supports(ISA::kSSE2) ? 1 : 0;
}

}  // extern "C"
