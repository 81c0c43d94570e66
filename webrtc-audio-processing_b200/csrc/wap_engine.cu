// Host engine + C ABI (include/wap_audio_processing.h) + the tick kernel.
//
// Host responsibilities restated from AudioProcessingImpl: format validation
// and error codes (audio_processing_impl.cc:163-323), processing-rate
// selection (:92-107), submodule on/off matrix (:362-446), stream-delay clamp
// (:1689-1707), statistics (:1509-1518).  The DSP itself runs only on the GPU:
// there is no CPU fallback and creation fails loudly without a CUDA device.
#include <math.h>
#include <cmath>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <algorithm>
#include <atomic>
#include <deque>
#include <mutex>
#include <new>
#include <functional>
#include <vector>

#include "wap_audio_processing.h"
#include "wap_init.h"
#include "wap_launch.h"
#include "wap_kernels.h"
#include "wap_pipeline.cuh"

namespace wap {

// One 10 ms tick = k_front -> (k_delay) -> k_echo on the engine's stream.
__global__ void __launch_bounds__(128) k_front(TickArgs a) {
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;
  if (idx < a.n) front_leg(a, idx);
}

// Engines with the residual echo detector: the power of the leg's render frame, one thread per leg.
__global__ void __launch_bounds__(128) k_red_render(TickArgs a) {
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;
  if (idx < a.n) red_analyze_render(a, idx);
}

// 48 kHz AEC3 engines: band split of render and capture in front of k_front, one warp per leg.
__global__ void __launch_bounds__(128) k_split(TickArgs a) {
  float* sm = reinterpret_cast<float*>(WAP_DYN_SMEM());  // 4 warps x 1120 floats
  const int warp = threadIdx.x >> 5;
  const int idx = blockIdx.x * 4 + warp;
  if (idx < a.n) split_tick(a, idx, sm + warp * 1120);
}

// Resampled engines: API-rate frames -> processing-rate FloatS16 frames, one warp per leg.
__global__ void __launch_bounds__(128) k_resample(TickArgs a) {
  float* sm = reinterpret_cast<float*>(WAP_DYN_SMEM());  // 4 warps x 2 * kRsMaxRequest floats
  const int warp = threadIdx.x >> 5;
  const int idx = blockIdx.x * 4 + warp;
  if (idx < a.n) resample_in_tick(a, idx, sm + warp * 2 * kRsMaxRequest);
}

// 48 kHz AEC3 legs only: PostFilter + output conversion, one thread per leg.
__global__ void __launch_bounds__(128) k_post(TickArgs a) {
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;
  if (idx < a.n) post_leg(a, idx);
}

// Broadcasts the initial-state template into `n` arena slots.
__global__ void k_init_slots(StreamState* states, const StreamState* tmpl, const int* slots, int n) {
  const size_t words = sizeof(StreamState) / 4;
  const uint32_t* src = reinterpret_cast<const uint32_t*>(tmpl);
  for (int i = blockIdx.y; i < n; i += gridDim.y) {
    uint32_t* dst = reinterpret_cast<uint32_t*>(&states[slots[i]]);
    for (size_t w = (size_t)blockIdx.x * blockDim.x + threadIdx.x; w < words; w += (size_t)gridDim.x * blockDim.x)
      dst[w] = src[w];
  }
}

// Multi-channel engines: a fresh McState (zeros, the channel templates, BlockFramer holding one block of
// zeros, the content detector's start state) into `n` arena slots.  which: 1 when the legs start with
// persistent multichannel content (detect_stereo_content off), else 0.
__global__ void k_mc_init_slots(McState* mcs, const McTemplates* t, const int* slots, int n, int which) {
  const size_t words = sizeof(McState) / 4;
  const size_t chan0 = offsetof(McState, chan) / 4, chan_words = sizeof(McChan) / 4;
  const uint32_t* tc = reinterpret_cast<const uint32_t*>(&t->chan[which]);
  for (int i = blockIdx.y; i < n; i += gridDim.y) {
    uint32_t* dst = reinterpret_cast<uint32_t*>(&mcs[slots[i]]);
    for (size_t w = (size_t)blockIdx.x * blockDim.x + threadIdx.x; w < words; w += (size_t)gridDim.x * blockDim.x) {
      uint32_t v = 0;
      if (w >= chan0 && w < chan0 + kMcCh * chan_words) v = tc[(w - chan0) % chan_words];
      else if (w == offsetof(McState, output_framer_len) / 4) v = kBlock;
      else if (w == (offsetof(McState, det) + offsetof(McDetector, persistent)) / 4) v = (uint32_t)which;
      else if (w == (offsetof(McState, det) + offsetof(McDetector, render_channels_to_aec)) / 4) v = which ? 2u : 1u;
      dst[w] = v;
    }
  }
}

}  // namespace wap

using wap::EngineConfig;
using wap::StreamState;

#define WAP_CUDA(x)                                                                      \
  do {                                                                                   \
    cudaError_t e_ = (x);                                                                \
    if (e_ != cudaSuccess) {                                                             \
      fprintf(stderr, "[wap_b200] CUDA error %s at %s:%d: %s\n", #x, __FILE__, __LINE__, \
              cudaGetErrorString(e_));                                                   \
      return WapError::Internal;                                                         \
    }                                                                                    \
  } while (0)

constexpr int kMaxChunks = 8;

// The three API formats of a leg: capture input, capture output, render (reverse) input.
struct WapFormats {
  WapStreamConfig in, out, render;
};
inline bool same_format(const WapStreamConfig& a, const WapStreamConfig& b) {
  return a.sample_rate_hz == b.sample_rate_hz && a.num_channels == b.num_channels;
}
inline WapFormats uniform_formats(const WapStreamConfig& f) { return WapFormats{f, f, f}; }

struct WapEngine {
  int device = 0;
  int capacity = 0;
  WapConfig config{};
  WapStreamConfig format{};   // capture input format
  WapFormats formats{};       // capture input, capture output, render input
  EngineConfig cfg{};
  StreamState* d_states = nullptr;
  wap::UpperBandState* d_upper = nullptr;  // 32 / 48 kHz AEC3 engines only
  StreamState* d_template = nullptr;
  cudaStream_t stream = nullptr;
  std::vector<int> free_slots;
  // Guards every member below and the engine's stream: tick, create / destroy and the per-leg
  // setters that touch engine state take it (recursive: the host-buffer tick calls the device tick).
  std::recursive_mutex mu;
  // staging for the host-buffer entry points
  void* d_render = nullptr;
  void* d_capture = nullptr;
  void* d_out = nullptr;
  int* d_slots = nullptr;
  int* d_delays = nullptr;
  void* h_pinned = nullptr;  // render | capture | out
  size_t staged_streams = 0;
  std::vector<int> last_slots;
  std::vector<WapAudioProcessing*> last_handles;  // the handle array last_slots was built from
  // per-leg host state indexed by slot, so the per-tick bookkeeping walks contiguous memory
  std::vector<int> leg_delay_ms;            // last set_stream_delay_ms value
  std::vector<unsigned char> leg_delay_set; // was_stream_delay_set (cleared by every capture frame)
  std::atomic<int> dirty_legs{0};           // handles whose capture_output_used is not yet in the slab
  int64_t launches = 0;
  int frame_len = 0;   // samples per capture input frame (all channels)
  int out_len = 0;     // ... per output frame
  int render_len = 0;  // ... per render frame
  int echo_scratch_floats = 0;
  int echo_smem_pad = 0;   // WAP_ECHO_SMEM_PAD_KB (occupancy experiments)
  // optional per-kernel timing (bench roofline): events around the three tick kernels
  bool timing = false;
  cudaEvent_t ev[4] = {nullptr, nullptr, nullptr, nullptr};
  double kernel_ms[3] = {0, 0, 0};
  int64_t timed_ticks = 0;
  int delay_scratch_floats = 0;
  bool is_default = false;
  int sm_count = 148;
  int echo_class = 0;  // wap::EchoClass: which k_echo instance serves this engine
  // EchoCanceller3Config: the default config runs on compile-time constants, anything else on the
  // run-time-parameter kernel instances (wap_ec3_params.h)
  wap::Ec3Params ep = wap::ec3_default_params();
  bool ec3_runtime = false;
  WapEchoCanceller3Config aec3_config{};
  // resampled engines (API rate != processing rate)
  wap::ResamplerState* d_rs = nullptr;   // [capacity][kRsPerLeg]
  float* d_rs_kernels = nullptr;         // in | out | render tables
  float* d_rs_render = nullptr;          // [staged][proc frame]
  float* d_rs_capture = nullptr;
  float* d_rs_capture1 = nullptr;        // stereo: second capture channel
  wap::ExtraChannelState* d_extra = nullptr;  // stereo engines: [capacity]
  float* d_cap_delay = nullptr;               // delay.fixed_capture_delay_samples > 0: [capacity][cap_delay_stride]
  int cap_delay_stride = 0;                   // floats per leg: bands * delay + 4 (insert position, padding)
  wap::EchoDetectorState* d_red = nullptr;    // wap_engine_enable_echo_detector: [capacity]
  // multi-channel engines (EngineConfig::mc)
  wap::McState* d_mc = nullptr;             // [capacity]
  wap::McTemplates* d_mc_templates = nullptr;
  wap::NsState* d_mc_ns = nullptr;          // [capacity][kMcCh] when NS is enabled
  wap::NsState* d_mc_ns_template = nullptr;
  wap::Ec3Params ep_mc = wap::ec3_default_params();
  wap::McParams mcp[2] = {};
  int mc_front_floats = 0, mc_echo_floats = 0, mc_echo_wpb = 4;
  double rs_ratio_in = 1.0, rs_ratio_out = 1.0, rs_ratio_render = 1.0;
  int forced_chunks = 0;  // wap_engine_set_pipeline_chunks; 0 = automatic
  // host-buffer entry point, large batches: copies of one half overlap the kernels of the other
  cudaStream_t copy_in = nullptr, copy_out = nullptr;
  cudaEvent_t ev_in[kMaxChunks] = {}, ev_done[kMaxChunks] = {}, ev_start = nullptr;
};

struct WapAudioProcessing {
  WapEngine* engine = nullptr;
  int slot = -1;
  WapConfig config{};
  int stream_delay_ms = 0;
  bool was_stream_delay_set = false;
  bool capture_output_used = true;
  bool capture_output_used_dirty = false;  // not yet written to the leg's state slab
  // runtime settings waiting for the next capture frame (HandleCaptureRuntimeSettings)
  bool pre_gain_dirty = false, post_gain_dirty = false, playout_volume_dirty = false, agc2_gain_dirty = false;
  float agc2_gain_factor = -1.f;  // < 0: the engine's configured gain
  bool agc2_reset_limiter = false;
  float pre_gain_target = 1.f, post_gain_target = 1.f;
  int playout_volume = -1;
  int analog_level = 0;
  std::deque<std::vector<unsigned char>> render_queue;  // SwapQueue stand-in (aec3_common.h:41)
  WapSampleFormat render_fmt = WapSampleFormat::I16;
  WapStreamConfig render_format{0, 0};  // format of the queued render frames
  std::mutex render_mu;
  // Capture-side lock of a single-leg handle (webrtc::AudioProcessing's mutex_capture_): engine
  // creation, ApplyConfig, Initialize, the setters and ProcessStream serialise on it; the render
  // thread only ever takes render_mu (ProcessReverseStream enqueues and returns).
  std::recursive_mutex mu;
  bool skip_first_reinit = false;  // the next private engine starts "already initialised" (ApplyConfig / Initialize ran InitializeLocked)
  bool owns_engine = false;
  bool has_aec3_config = false;            // wap_create_with_aec3_config: injected EchoCanceller3Config
  WapEchoCanceller3Config aec3_config{};
  bool has_aec3_mc_config = false;         // ... and multichannel EchoCanceller3Config
  WapEchoCanceller3Config aec3_mc_config{};
  WapStats cached_stats{};  // ApmStatsReporter::cached_stats_
};

namespace {

bool g_warned_no_device = false;

WapError check_device() {
  int n = 0;
  cudaError_t e = cudaGetDeviceCount(&n);
  if (e != cudaSuccess || n <= 0) {
    if (!g_warned_no_device) {
      fprintf(stderr,
              "[wap_b200] FATAL: no CUDA device available (%s). This library has no CPU fallback.\n",
              e != cudaSuccess ? cudaGetErrorString(e) : "device count 0");
      g_warned_no_device = true;
    }
    return WapError::Internal;
  }
  return WapError::None;
}

// API rates the resamplers are sized for (kRsMaxRequest samples per 10 ms).
constexpr int kMaxApiRate = wap::kRsMaxRequest * 100;

// SuitableProcessRate / InitializeLocked (audio_processing_impl.cc:92-107,527-692) for the config
// classes of SURVEY.md section 8.
WapError resolve_config(const WapConfig& c, const WapFormats& fm, EngineConfig* out) {
  const WapStreamConfig& f = fm.in;
  for (const WapStreamConfig* s : {&fm.in, &fm.out, &fm.render}) {
    if (s->sample_rate_hz < 8000 || s->sample_rate_hz > 384000) return WapError::BadSampleRate;
    if (s->num_channels <= 0) return WapError::BadNumberChannels;
  }
  if (fm.out.num_channels != 1 && fm.out.num_channels != fm.in.num_channels) return WapError::BadNumberChannels;
  EngineConfig e{};
  // InitializeLocked (audio_processing_impl.cc:632-692): the processing rate is the lowest native
  // rate that covers the lower of the capture input and output rates, capped by
  // maximum_internal_processing_rate when a multi-band submodule is active; with AEC3 the render
  // stream is processed at the same rate.
  for (const WapStreamConfig* s : {&fm.in, &fm.out, &fm.render})
    if (s->sample_rate_hz % 100 != 0 || s->sample_rate_hz > kMaxApiRate) return WapError::UnsupportedConfig;
  if (c.pipeline_maximum_internal_processing_rate != 32000 && c.pipeline_maximum_internal_processing_rate != 48000)
    return WapError::UnsupportedConfig;
  const bool multi_band = c.high_pass_filter_enabled || c.noise_suppression_enabled || c.echo_canceller_enabled;
  const int uppermost = multi_band ? c.pipeline_maximum_internal_processing_rate : 48000;
  const int min_rate = std::min(fm.in.sample_rate_hz, fm.out.sample_rate_hz);
  int proc = uppermost;
  for (int rate : {16000, 32000, 48000}) {
    if (rate >= uppermost) { proc = uppermost; break; }
    if (rate >= min_rate) { proc = rate; break; }
  }
  e.num_bands = proc / 16000;
  e.api_frame = fm.in.sample_rate_hz / 100;
  e.out_frame = fm.out.sample_rate_hz / 100;
  e.render_frame = fm.render.sample_rate_hz / 100;
  e.resample = proc != fm.in.sample_rate_hz ? 1 : 0;
  e.resample_render = (c.echo_canceller_enabled && proc != fm.render.sample_rate_hz) ? 1 : 0;
  e.pre_stage = (e.resample || e.resample_render) ? 1 : 0;
  e.fullband_out = (proc < fm.out.sample_rate_hz && fm.out.sample_rate_hz == 48000) ? 1 : 0;
  e.resample_out = (proc != fm.out.sample_rate_hz && !e.fullband_out) ? 1 : 0;
  e.hpf_rate = e.fullband_out ? 48000 : proc;
  e.in_channels = fm.in.num_channels;
  e.render_channels = fm.render.num_channels;
  e.downmix_first = c.pipeline_capture_downmix_method == WapDownmixMethod::UseFirstChannel ? 1 : 0;
  // The capture_fullband_audio buffer is filled from the capture input (resampled to 48 kHz when the
  // input has another rate) and comes back unprocessed while the output is muted: only built for an
  // input that already has the output's rate (a channel downmix on the way in is).
  if (e.fullband_out && (fm.in.sample_rate_hz != fm.out.sample_rate_hz ||
                         (fm.in.num_channels != fm.out.num_channels && fm.out.num_channels != 1)))
    return WapError::UnsupportedConfig;
  // 48 kHz AEC3 runs the PostFilter and the output conversion in k_post: the output must be 48 kHz too.
  if (e.num_bands == 3 && c.echo_canceller_enabled && e.resample_out) return WapError::UnsupportedConfig;
  // The capture AudioBuffer has the OUTPUT's channel count (an input with more channels is downmixed on
  // the way in).  Two-channel buffers are processed as the default pipeline does -- first channel
  // only, mono result copied to both output channels -- or, with pipeline.multi_channel_render /
  // _capture, truly multi-channel.  Stereo buffers need AEC3 (without it the reference
  // runs NS / AGC2 on both channels).  The flags only matter for frames with more than one channel
  // (mono legs run the mono EchoCanceller3Config whatever they say: config_selector.cc:44-58).
  const int buf_channels = fm.out.num_channels;
  if (buf_channels > 2 || fm.in.num_channels > 8 || fm.render.num_channels > 8) return WapError::UnsupportedConfig;
  if (buf_channels == 2 && !c.echo_canceller_enabled) {
    // Stereo capture without an echo controller: both channels run through the high-pass filter, the noise
    // suppressor (minima over the channels) and AGC2 whatever the pipeline flags say (the reduction to one
    // channel only happens next to an echo controller, audio_processing_impl.cc:1365-1373).  Served by the
    // multi-channel kernels with their AEC3 half switched off; native 16 / 48 kHz, stereo input.
    // (48 kHz frames are only split into bands next to a multi-band submodule: AGC2 alone is not one)
    if (e.pre_stage || e.resample_out || e.fullband_out || e.num_bands == 2 || fm.in.num_channels != 2 ||
        (e.num_bands == 3 && !multi_band))
      return WapError::UnsupportedConfig;
    e.mc = 1;
  }
  if (buf_channels == 2 && fm.render.num_channels != 2 && c.echo_canceller_enabled &&
      (c.pipeline_multi_channel_render || c.pipeline_multi_channel_capture))
    return WapError::UnsupportedConfig;
  if (buf_channels == 2 && fm.render.num_channels == 2 && (c.pipeline_multi_channel_render || c.pipeline_multi_channel_capture)) {
    // True multi-channel processing (BASELINE config 4): both flags, AEC3 (+ its high-pass filter) with or
    // without the noise suppressor, at a native rate of 16 or 48 kHz, one format for all three streams.
    if (!(c.pipeline_multi_channel_render && c.pipeline_multi_channel_capture) || e.pre_stage ||
        e.resample_out || e.num_bands == 2 || !same_format(fm.in, fm.out) || !same_format(fm.in, fm.render))
      return WapError::UnsupportedConfig;
    e.mc = 1;
  }
  e.channels = buf_channels;
  e.levels_enabled = (c.pre_amplifier_enabled || c.capture_level_adjustment_enabled) ? 1 : 0;
  e.post_gain_enabled = c.capture_level_adjustment_enabled ? 1 : 0;
  // capture level adjustment: pre / post gains; the analog mic gain emulation needs an input volume
  // controller.
  if (c.capture_level_adjustment_enabled && c.analog_mic_gain_emulation_enabled) return WapError::UnsupportedConfig;
  // AGC2: fixed digital gain + limiter (the default sub-configuration); the adaptive digital
  // controller and the input volume controller are SURVEY 8(f)-1.
  if (c.gain_controller2_enabled &&
      (c.gain_controller2_adaptive_digital_enabled || c.gain_controller2_input_volume_controller_enabled))
    return WapError::UnsupportedConfig;
  e.sample_rate_hz = proc;
  // high_pass_filter.apply_in_full_band = false moves the filter onto split band 0 with the 16 kHz
  // coefficients (audio_processing_impl.cc:1283,1376,1890): not built -- refused rather than approximated.
  if (!c.high_pass_filter_apply_in_full_band && e.num_bands > 1 &&
      (c.high_pass_filter_enabled || c.noise_suppression_enabled ||
       (c.echo_canceller_enabled && c.echo_canceller_enforce_high_pass_filtering)))
    return WapError::UnsupportedConfig;
  e.aec_enabled = c.echo_canceller_enabled;
  e.ns_enabled = c.noise_suppression_enabled;
  // InitializeHighPassFilter (audio_processing_impl.cc:1883-1907)
  e.hpf_enabled = c.high_pass_filter_enabled || c.noise_suppression_enabled ||
                  (c.echo_canceller_enabled && c.echo_canceller_enforce_high_pass_filtering);
  switch (c.noise_suppression_level) {  // suppression_params.cc:18-48
    case WapNoiseSuppressionLevel::Low:
      e.ns_over_subtraction_factor = 1.f; e.ns_minimum_attenuating_gain = 0.5f; e.ns_use_attenuation_adjustment = 0; break;
    case WapNoiseSuppressionLevel::Moderate:
      e.ns_over_subtraction_factor = 1.f; e.ns_minimum_attenuating_gain = 0.25f; e.ns_use_attenuation_adjustment = 1; break;
    case WapNoiseSuppressionLevel::High:
      e.ns_over_subtraction_factor = 1.1f; e.ns_minimum_attenuating_gain = 0.125f; e.ns_use_attenuation_adjustment = 1; break;
    default:
      e.ns_over_subtraction_factor = 1.25f; e.ns_minimum_attenuating_gain = 0.09f; e.ns_use_attenuation_adjustment = 1; break;
  }
  e.capture_output_used = 1;
  e.agc2_enabled = c.gain_controller2_enabled ? 1 : 0;
  e.split_bands = (e.num_bands >= 2 && (c.high_pass_filter_enabled || e.ns_enabled || e.aec_enabled)) ? 1 : 0;
  e.agc2_fixed_gain = powf(10.0f, c.gain_controller2_fixed_digital_gain_db / 20.0f);  // DbToRatio (audio_util.h:85-87)
  // The constructor's api_format is 16 kHz mono for every stream.  MaybeInitializeCapture re-initialises
  // when the capture input / output formats differ from it, or when UpdateActiveSubmoduleStates() sees
  // the submodules the constructor created for the first time (audio_processing_impl.cc:894-925) --
  // unless a render call with another format came first: MaybeInitializeRender's InitializeLocked() has
  // then already taken note of them (:543-556,632-633), and render audio queued before the first
  // capture call survives.
  const bool capture_default = fm.in.sample_rate_hz == 16000 && fm.in.num_channels == 1 &&
                               fm.out.sample_rate_hz == 16000 && fm.out.num_channels == 1;
  const bool render_default = fm.render.sample_rate_hz == 16000 && fm.render.num_channels == 1;
  e.reinit_on_first_capture =
      (!capture_default || ((c.noise_suppression_enabled || c.gain_controller2_enabled) && render_default)) ? 1 : 0;
  e.cng_noise_floor = 64.f * powf(10.f, (90.30899869919436f + -96.03406f) * 0.1f);
  *out = e;
  return WapError::None;
}

WapError resolve_config(const WapConfig& c, const WapStreamConfig& f, EngineConfig* out) {
  return resolve_config(c, uniform_formats(f), out);
}

WapError ensure_staging(WapEngine* e, size_t n) {
  if (n <= e->staged_streams) return WapError::None;
  size_t cap = std::max<size_t>(n, std::min<size_t>((size_t)e->capacity, std::max<size_t>(64, 2 * e->staged_streams)));
  const size_t fb = (size_t)e->frame_len * sizeof(float);     // capture input
  const size_t rfb = (size_t)e->render_len * sizeof(float);   // render
  const size_t ofb = (size_t)e->out_len * sizeof(float);      // output
  const size_t pb = (size_t)wap::kFrame * e->cfg.num_bands * sizeof(float);
  // Allocate the new buffers first and swap them in only when every allocation succeeded: a failure
  // leaves the engine with its old (smaller) staging area instead of dangling pointers.
  void *n_render = nullptr, *n_capture = nullptr, *n_out = nullptr, *n_pinned = nullptr;
  int *n_slots = nullptr, *n_delays = nullptr;
  float *n_rs_render = nullptr, *n_rs_capture = nullptr, *n_rs_capture1 = nullptr;
  bool ok = cudaMalloc(&n_render, cap * rfb) == cudaSuccess && cudaMalloc(&n_capture, cap * fb) == cudaSuccess &&
            cudaMalloc(&n_out, cap * ofb) == cudaSuccess && cudaMalloc((void**)&n_slots, cap * sizeof(int)) == cudaSuccess &&
            cudaMalloc((void**)&n_delays, cap * sizeof(int)) == cudaSuccess &&
            cudaMallocHost(&n_pinned, cap * (rfb + fb + ofb + 2 * sizeof(int))) == cudaSuccess;
  if (ok && e->cfg.pre_stage) {
    ok = cudaMalloc((void**)&n_rs_render, cap * pb) == cudaSuccess && cudaMalloc((void**)&n_rs_capture, cap * pb) == cudaSuccess;
    if (ok && e->cfg.channels == 2) ok = cudaMalloc((void**)&n_rs_capture1, cap * pb) == cudaSuccess;
  }
  if (!ok) {
    fprintf(stderr, "[wap_b200] staging allocation for %zu legs failed: %s\n", cap, cudaGetErrorString(cudaGetLastError()));
    cudaFree(n_render); cudaFree(n_capture); cudaFree(n_out); cudaFree(n_slots); cudaFree(n_delays);
    if (n_pinned) cudaFreeHost(n_pinned);
    cudaFree(n_rs_render); cudaFree(n_rs_capture); cudaFree(n_rs_capture1);
    return WapError::Internal;
  }
  cudaStreamSynchronize(e->stream);
  cudaFree(e->d_render); cudaFree(e->d_capture); cudaFree(e->d_out); cudaFree(e->d_slots); cudaFree(e->d_delays);
  if (e->h_pinned) cudaFreeHost(e->h_pinned);
  e->d_render = n_render; e->d_capture = n_capture; e->d_out = n_out; e->d_slots = n_slots; e->d_delays = n_delays;
  e->h_pinned = n_pinned;
  if (e->cfg.pre_stage) {
    cudaFree(e->d_rs_render); cudaFree(e->d_rs_capture); cudaFree(e->d_rs_capture1);
    e->d_rs_render = n_rs_render; e->d_rs_capture = n_rs_capture; e->d_rs_capture1 = n_rs_capture1;
  }
  e->staged_streams = cap;
  e->last_slots.clear();
  e->last_handles.clear();
  return WapError::None;
}

bool is_pinned_host(const void* p) {
#if defined(WAP_EMU)
  (void)p;
  return false;
#else
  cudaPointerAttributes at;
  if (cudaPointerGetAttributes(&at, p) != cudaSuccess) {
    cudaGetLastError();  // unregistered host memory on older runtimes: clear the error
    return false;
  }
  return at.type == cudaMemoryTypeHost;
#endif
}

// SincResampler::InitializeKernel (sinc_resampler.cc:212-246): 33 sub-sample offsets of a 32-tap
// Blackman-windowed sinc, evaluated like the reference (float index arithmetic, double
// trigonometry, float storage) so that the table is the reference's bit for bit.
void build_sinc_kernel(double io_ratio, float* table) {
  const double alpha = 0.16;
  const double a0 = 0.5 * (1.0 - alpha), a1 = 0.5, a2 = 0.5 * alpha;
  const double pi = 3.14159265358979323846;
  double scale = io_ratio > 1.0 ? 1.0 / io_ratio : 1.0;  // SincScaleFactor: low-pass when downsampling
  scale *= 0.9;                                          // ... with some roll-off margin
  const int taps = wap::kRsKernelSize;
  for (int o = 0; o <= wap::kRsOffsetCount; ++o) {
    const float frac = static_cast<float>(o) / wap::kRsOffsetCount;
    for (int i = 0; i < taps; ++i) {
      const float arg = static_cast<float>(pi * (static_cast<float>(i - taps / 2) - frac));
      const float x = (static_cast<float>(i) - frac) / static_cast<float>(taps);
      const float window = static_cast<float>(a0 - a1 * cos(2.0 * pi * x) + a2 * cos(4.0 * pi * x));
      const double sinc = arg == 0 ? scale : sin(scale * arg) / arg;
      table[o * taps + i] = static_cast<float>(window * sinc);
    }
  }
}

// ---- EchoCanceller3Config (api/audio/echo_canceller3_config.h:21-275) ------------------------
WapEchoCanceller3Config ec3_config_default() {
  WapEchoCanceller3Config c{};
  c.buffering = {250, 8};
  c.delay.default_delay = 5; c.delay.down_sampling_factor = 4; c.delay.num_filters = 5;
  c.delay.delay_headroom_samples = 32; c.delay.hysteresis_limit_blocks = 1; c.delay.fixed_capture_delay_samples = 0;
  c.delay.delay_estimate_smoothing = 0.7f; c.delay.delay_estimate_smoothing_delay_found = 0.7f;
  c.delay.delay_candidate_detection_threshold = 0.2f;
  c.delay.delay_selection_thresholds = {5, 20};
  c.delay.use_external_delay_estimator = false; c.delay.log_warning_on_delay_changes = false;
  c.delay.render_alignment_mixing = {false, true, 10000.f, true};
  c.delay.capture_alignment_mixing = {false, true, 10000.f, false};
  c.delay.detect_pre_echo = true;
  c.filter.refined = {13, 0.00005f, 0.05f, 0.001f, 2.f, 20075344.f};
  c.filter.coarse = {13, 0.7f, 20075344.f};
  c.filter.refined_initial = {12, 0.005f, 0.5f, 0.001f, 2.f, 20075344.f};
  c.filter.coarse_initial = {12, 0.9f, 20075344.f};
  c.filter.config_change_duration_blocks = 250; c.filter.initial_state_seconds = 2.5f;
  c.filter.coarse_reset_hangover_blocks = 25;
  c.filter.conservative_initial_phase = false; c.filter.enable_coarse_filter_output_usage = true;
  c.filter.use_linear_filter = true; c.filter.high_pass_filter_echo_reference = false;
  c.filter.export_linear_aec_output = false;
  c.erle = {1.f, 4.f, 1.5f, true, 1, true, true};
  c.ep_strength = {1.f, 0.83f, 0.83f, true, false, false, true};
  c.echo_audibility = {4 * 64.f, 64.f, 2 * 64.f, 10.f, 10.f, 10.f, false, false};
  c.render_levels = {100.f, 150.f, 20.f, 0.f};
  c.echo_removal_control = {false, false};
  c.echo_model = {50, 1638400.f, 10.f, 27509.42f, 0.3f, 1, 1, true};
  c.comfort_noise.noise_floor_dbfs = -96.03406f;
  c.suppressor.nearend_average_blocks = 4;
  c.suppressor.normal_tuning = {{.3f, .4f, .3f}, {.07f, .1f, .3f}, 2.0f, 0.25f};
  c.suppressor.nearend_tuning = {{1.09f, 1.1f, .3f}, {.1f, .3f, .3f}, 2.0f, 0.25f};
  c.suppressor.lf_smoothing_during_initial_phase = true;
  c.suppressor.last_permanent_lf_smoothing_band = 0; c.suppressor.last_lf_smoothing_band = 5;
  c.suppressor.last_lf_band = 5; c.suppressor.first_hf_band = 8;
  c.suppressor.dominant_nearend_detection = {.25f, 10.f, 30.f, 50, 12, true, true};
  c.suppressor.subband_nearend_detection = {1, {1, 1}, {1, 1}, 1.f, 1.f};
  c.suppressor.use_subband_nearend_detection = false;
  c.suppressor.high_bands_suppression = {1.f, 1.f, 400.f, 1.f};
  c.suppressor.high_frequency_suppression = {16, 1};
  c.suppressor.floor_first_increase = 0.00001f;
  c.suppressor.conservative_hf_suppression = false;
  c.multi_channel = {true, 0.0f, 300, 2.0f};
  return c;
}

// Members that fix the structure of the engine must keep their default value (see the header).
WapError ec3_config_supported(const WapEchoCanceller3Config& c) {
  const WapEchoCanceller3Config d = ec3_config_default();
  const bool ok =
      c.delay.down_sampling_factor == 4 && c.delay.num_filters == 5 && c.delay.fixed_capture_delay_samples >= 0 &&
      c.delay.fixed_capture_delay_samples <= 5000 &&
      c.delay.default_delay >= 0 &&
      c.delay.default_delay <= wap::kMaxRingDelay &&
      c.filter.refined.length_blocks >= 1 && c.filter.refined.length_blocks <= wap::kMaxPartitions &&
      c.filter.coarse.length_blocks >= 1 && c.filter.coarse.length_blocks <= wap::kMaxPartitions &&
      c.filter.refined_initial.length_blocks >= 1 && c.filter.refined_initial.length_blocks <= c.filter.refined.length_blocks &&
      c.filter.coarse_initial.length_blocks >= 1 && c.filter.coarse_initial.length_blocks <= c.filter.coarse.length_blocks &&
      c.filter.config_change_duration_blocks >= 1 &&
      c.filter.export_linear_aec_output == d.filter.export_linear_aec_output &&
      // SignalDependentErleEstimator: sections of the refined filter behind the delay headroom
      c.erle.num_sections >= 1 &&
      (c.erle.num_sections == 1 ||
       (c.delay.delay_headroom_samples >= 0 &&
        c.erle.num_sections <= c.filter.refined.length_blocks - c.delay.delay_headroom_samples / (int)wap::kBlock)) &&
      // default_len < 0: the adaptive reverb decay (ReverbDecayEstimator); its EarlyReverbLengthEstimator holds
      // length_blocks - 9 sections, so the reference itself needs 10 blocks or more
      (c.ep_strength.default_len >= 0.f || c.filter.refined.length_blocks >= 10) &&


      c.echo_model.render_pre_window_size >= 0 && c.echo_model.render_pre_window_size <= 100 &&
      c.echo_model.render_post_window_size >= 0 && c.echo_model.render_post_window_size <= 100 &&
      c.suppressor.nearend_average_blocks >= 1 && c.suppressor.nearend_average_blocks <= 4 &&
      // SubbandNearendDetector: its smoother holds at most three past blocks here
      (!c.suppressor.use_subband_nearend_detection ||
       (c.suppressor.subband_nearend_detection.nearend_average_blocks >= 1 &&
        c.suppressor.subband_nearend_detection.nearend_average_blocks <= 4 &&
        c.suppressor.subband_nearend_detection.subband1.low >= 0 && c.suppressor.subband_nearend_detection.subband1.high < wap::kBins &&
        c.suppressor.subband_nearend_detection.subband1.low <= c.suppressor.subband_nearend_detection.subband1.high &&
        c.suppressor.subband_nearend_detection.subband2.low >= 0 && c.suppressor.subband_nearend_detection.subband2.high < wap::kBins &&
        c.suppressor.subband_nearend_detection.subband2.low <= c.suppressor.subband_nearend_detection.subband2.high)) &&

      c.suppressor.high_frequency_suppression.limiting_gain_band >= 0 &&
      c.suppressor.high_frequency_suppression.bands_in_limiting_gain >= 0 &&
      c.suppressor.high_frequency_suppression.limiting_gain_band +
              c.suppressor.high_frequency_suppression.bands_in_limiting_gain <= wap::kBins &&
      c.suppressor.last_lf_band >= 0 && c.suppressor.first_hf_band > c.suppressor.last_lf_band;
  return ok ? WapError::None : WapError::UnsupportedConfig;
}

wap::Ec3Params ec3_params_from_config(const WapEchoCanceller3Config& c) {
  wap::Ec3Params p{};
  p.excess_render_detection_interval_blocks = c.buffering.excess_render_detection_interval_blocks;
  p.max_allowed_excess_render_blocks = c.buffering.max_allowed_excess_render_blocks;
  p.default_delay = c.delay.default_delay;
  p.delay_headroom_samples = c.delay.delay_headroom_samples;
  p.hysteresis_limit_blocks = c.delay.hysteresis_limit_blocks;
  p.thr_initial = c.delay.delay_selection_thresholds.initial;
  p.thr_converged = c.delay.delay_selection_thresholds.converged;
  p.delay_estimate_smoothing = c.delay.delay_estimate_smoothing;
  p.delay_estimate_smoothing_delay_found = c.delay.delay_estimate_smoothing_delay_found;
  p.delay_candidate_detection_threshold = c.delay.delay_candidate_detection_threshold;
  p.active_render_limit = c.render_levels.active_render_limit;
  p.poor_excitation_render_limit = c.render_levels.poor_excitation_render_limit;
  p.refined_len = c.filter.refined.length_blocks;
  p.coarse_len = c.filter.coarse.length_blocks;
  p.refined_initial_len = c.filter.refined_initial.length_blocks;
  p.coarse_initial_len = c.filter.coarse_initial.length_blocks;
  const WapEc3RefinedConfiguration* rc[2] = {&c.filter.refined, &c.filter.refined_initial};
  float* rd[2] = {p.refined, p.refined_initial};
  for (int i = 0; i < 2; ++i) {
    rd[i][0] = rc[i]->leakage_converged; rd[i][1] = rc[i]->leakage_diverged; rd[i][2] = rc[i]->error_floor;
    rd[i][3] = rc[i]->error_ceil; rd[i][4] = rc[i]->noise_gate;
  }
  p.coarse[0] = c.filter.coarse.rate; p.coarse[1] = c.filter.coarse.noise_gate;
  p.coarse_initial[0] = c.filter.coarse_initial.rate; p.coarse_initial[1] = c.filter.coarse_initial.noise_gate;
  p.config_change_duration_blocks = c.filter.config_change_duration_blocks;
  p.coarse_reset_hangover_blocks = c.filter.coarse_reset_hangover_blocks;
  p.initial_state_seconds = c.filter.initial_state_seconds;
  p.erle_min = c.erle.min; p.erle_max_l = c.erle.max_l; p.erle_max_h = c.erle.max_h;
  p.default_gain = c.ep_strength.default_gain; p.default_len = c.ep_strength.default_len;
  p.nearend_len = c.ep_strength.nearend_len;
  p.low_render_limit = c.echo_audibility.low_render_limit; p.normal_render_limit = c.echo_audibility.normal_render_limit;
  p.floor_power = c.echo_audibility.floor_power;
  p.audibility_threshold_lf = c.echo_audibility.audibility_threshold_lf;
  p.audibility_threshold_mf = c.echo_audibility.audibility_threshold_mf;
  p.audibility_threshold_hf = c.echo_audibility.audibility_threshold_hf;
  p.noise_floor_hold = c.echo_model.noise_floor_hold; p.min_noise_floor_power = c.echo_model.min_noise_floor_power;
  p.stationary_gate_slope = c.echo_model.stationary_gate_slope; p.noise_gate_power = c.echo_model.noise_gate_power;
  p.noise_gate_slope = c.echo_model.noise_gate_slope;
  const WapEc3Tuning* tc[2] = {&c.suppressor.normal_tuning, &c.suppressor.nearend_tuning};
  wap::Ec3Tuning* td[2] = {&p.normal_tuning, &p.nearend_tuning};
  for (int i = 0; i < 2; ++i)
    *td[i] = {tc[i]->mask_lf.enr_transparent, tc[i]->mask_lf.enr_suppress, tc[i]->mask_lf.emr_transparent,
              tc[i]->mask_hf.enr_transparent, tc[i]->mask_hf.enr_suppress, tc[i]->mask_hf.emr_transparent,
              tc[i]->max_inc_factor, tc[i]->max_dec_factor_lf};
  p.last_permanent_lf_smoothing_band = c.suppressor.last_permanent_lf_smoothing_band;
  p.last_lf_smoothing_band = c.suppressor.last_lf_smoothing_band;
  p.last_lf_band = c.suppressor.last_lf_band; p.first_hf_band = c.suppressor.first_hf_band;
  p.dn_enr_threshold = c.suppressor.dominant_nearend_detection.enr_threshold;
  p.dn_enr_exit_threshold = c.suppressor.dominant_nearend_detection.enr_exit_threshold;
  p.dn_snr_threshold = c.suppressor.dominant_nearend_detection.snr_threshold;
  p.dn_hold_duration = c.suppressor.dominant_nearend_detection.hold_duration;
  p.dn_trigger_threshold = c.suppressor.dominant_nearend_detection.trigger_threshold;
  p.hb_enr_threshold = c.suppressor.high_bands_suppression.enr_threshold;
  p.hb_max_gain_during_echo = c.suppressor.high_bands_suppression.max_gain_during_echo;
  p.hb_anti_howling_activation_threshold = c.suppressor.high_bands_suppression.anti_howling_activation_threshold;
  p.hb_anti_howling_gain = c.suppressor.high_bands_suppression.anti_howling_gain;
  p.limiting_gain_band = c.suppressor.high_frequency_suppression.limiting_gain_band;
  p.bands_in_limiting_gain = c.suppressor.high_frequency_suppression.bands_in_limiting_gain;
  p.floor_first_increase = c.suppressor.floor_first_increase;
  p.high_pass_filter_echo_reference = c.filter.high_pass_filter_echo_reference ? 1 : 0;
  p.fixed_capture_delay_samples = c.delay.fixed_capture_delay_samples;
  p.use_subband_nearend_detection = c.suppressor.use_subband_nearend_detection ? 1 : 0;
  p.snd_average_blocks = c.suppressor.subband_nearend_detection.nearend_average_blocks;
  p.snd_sub1_low = c.suppressor.subband_nearend_detection.subband1.low;
  p.snd_sub1_high = c.suppressor.subband_nearend_detection.subband1.high;
  p.snd_sub2_low = c.suppressor.subband_nearend_detection.subband2.low;
  p.snd_sub2_high = c.suppressor.subband_nearend_detection.subband2.high;
  p.snd_nearend_threshold = c.suppressor.subband_nearend_detection.nearend_threshold;
  p.snd_snr_threshold = c.suppressor.subband_nearend_detection.snr_threshold;
  p.echo_can_saturate = c.ep_strength.echo_can_saturate;
  p.bounded_erl = c.ep_strength.bounded_erl;
  p.erle_onset_compensation_in_dominant_nearend = c.ep_strength.erle_onset_compensation_in_dominant_nearend;
  p.use_conservative_tail_frequency_response = c.ep_strength.use_conservative_tail_frequency_response;
  p.erle_onset_detection = c.erle.onset_detection;
  p.clamp_quality_estimate_to_zero = c.erle.clamp_quality_estimate_to_zero;
  p.clamp_quality_estimate_to_one = c.erle.clamp_quality_estimate_to_one;
  p.has_clock_drift = c.echo_removal_control.has_clock_drift;
  p.linear_and_stable_echo_path = c.echo_removal_control.linear_and_stable_echo_path;
  p.lf_smoothing_during_initial_phase = c.suppressor.lf_smoothing_during_initial_phase;
  p.dn_use_during_initial_phase = c.suppressor.dominant_nearend_detection.use_during_initial_phase;
  p.dn_use_unbounded_echo_spectrum = c.suppressor.dominant_nearend_detection.use_unbounded_echo_spectrum;
  p.conservative_hf_suppression = c.suppressor.conservative_hf_suppression;
  p.conservative_initial_phase = c.filter.conservative_initial_phase;
  p.enable_coarse_filter_output_usage = c.filter.enable_coarse_filter_output_usage;
  p.use_linear_filter = c.filter.use_linear_filter;
  p.render_pre_window_size = c.echo_model.render_pre_window_size;
  p.render_post_window_size = c.echo_model.render_post_window_size;
  p.model_reverb_in_nonlinear_mode = c.echo_model.model_reverb_in_nonlinear_mode;
  p.nearend_average_blocks = c.suppressor.nearend_average_blocks;
  // RenderDelayBufferImpl: std::pow(10.0f, render_power_gain_db / 20.f) (render_delay_buffer.cc:124-125)
  p.render_linear_amplitude_gain = powf(10.0f, c.render_levels.render_power_gain_db / 20.f);
  p.detect_pre_echo = c.delay.detect_pre_echo;
  p.use_external_delay_estimator = c.delay.use_external_delay_estimator;
  p.erle_num_sections = c.erle.num_sections;
  if (c.erle.num_sections > 1) {
    // SetSectionsBoundaries / DefineFilterSectionSizes (signal_dependent_erle_estimator.cc:46-110): sections
    // of 2, 4, 8 ... blocks while more than that many blocks per remaining section are left, the rest split
    // evenly, the remainder to the last one; the first section starts behind the delay headroom
    const int num_sections = c.erle.num_sections, num_blocks = c.filter.refined.length_blocks;
    const int headroom = c.delay.delay_headroom_samples / (int)wap::kBlock;
    int sizes[wap::kMaxPartitions] = {};
    int remaining_blocks = num_blocks - headroom, remaining_sections = num_sections, estimator_size = 2, idx = 0;
    while (remaining_sections > 1 && remaining_blocks > estimator_size * remaining_sections) {
      sizes[idx++] = estimator_size;
      remaining_blocks -= estimator_size;
      --remaining_sections;
      estimator_size *= 2;
    }
    const int last_groups_size = remaining_blocks / remaining_sections;
    for (; idx < num_sections; ++idx) sizes[idx] = last_groups_size;
    sizes[num_sections - 1] += remaining_blocks - last_groups_size * remaining_sections;
    p.sd_boundaries[0] = headroom;
    int section = 0, current_size_block = 0;
    for (int k = headroom; k < num_blocks; ++k) {
      if (++current_size_block >= sizes[section]) {
        if (++section == num_sections) break;
        p.sd_boundaries[section] = k + 1;
        current_size_block = 0;
      }
    }
    p.sd_boundaries[num_sections] = num_blocks;
  }
  p.use_stationarity_properties = c.echo_audibility.use_stationarity_properties;
  p.use_stationarity_properties_at_init = c.echo_audibility.use_stationarity_properties_at_init;
  return p;
}

// EchoCanceller3Config::Validate (echo_canceller3_config.cc:101-286).
template <class T>
bool ec3_limit(T* v, T lo, T hi) {
  T c = *v < lo ? lo : (*v > hi ? hi : *v);
  bool res = *v == c;
  *v = c;
  return res;
}
bool ec3_limit(float* v, float lo, float hi) {
  float c = *v < lo ? lo : (*v > hi ? hi : *v);   // SafeClamp; NaN compares false and falls through
  if (!std::isfinite(c)) c = lo;
  bool res = *v == c;
  *v = c;
  return res;
}
bool ec3_floor(int32_t* v, int32_t lo) {
  bool res = *v >= lo;
  if (!res) *v = lo;
  return res;
}
bool ec3_validate(WapEchoCanceller3Config* c) {
  bool res = true;
  if (c->delay.down_sampling_factor != 4 && c->delay.down_sampling_factor != 8) { c->delay.down_sampling_factor = 4; res = false; }
  res &= ec3_limit(&c->delay.default_delay, 0, 5000);
  res &= ec3_limit(&c->delay.num_filters, 0, 5000);
  res &= ec3_limit(&c->delay.delay_headroom_samples, 0, 5000);
  res &= ec3_limit(&c->delay.hysteresis_limit_blocks, 0, 5000);
  res &= ec3_limit(&c->delay.fixed_capture_delay_samples, 0, 5000);
  res &= ec3_limit(&c->delay.delay_estimate_smoothing, 0.f, 1.f);
  res &= ec3_limit(&c->delay.delay_candidate_detection_threshold, 0.f, 1.f);
  res &= ec3_limit(&c->delay.delay_selection_thresholds.initial, 1, 250);
  res &= ec3_limit(&c->delay.delay_selection_thresholds.converged, 1, 250);
  WapEc3RefinedConfiguration* rc[2] = {&c->filter.refined, &c->filter.refined_initial};
  for (int i = 0; i < 2; ++i) {
    res &= ec3_floor(&rc[i]->length_blocks, 1);
    res &= ec3_limit(&rc[i]->leakage_converged, 0.f, 1000.f);
    res &= ec3_limit(&rc[i]->leakage_diverged, 0.f, 1000.f);
    res &= ec3_limit(&rc[i]->error_floor, 0.f, 1000.f);
    res &= ec3_limit(&rc[i]->error_ceil, 0.f, 100000000.f);
    res &= ec3_limit(&rc[i]->noise_gate, 0.f, 100000000.f);
  }
  if (c->filter.refined.length_blocks < c->filter.refined_initial.length_blocks) {
    c->filter.refined_initial.length_blocks = c->filter.refined.length_blocks;
    res = false;
  }
  WapEc3CoarseConfiguration* cc[2] = {&c->filter.coarse, &c->filter.coarse_initial};
  for (int i = 0; i < 2; ++i) {
    res &= ec3_floor(&cc[i]->length_blocks, 1);
    res &= ec3_limit(&cc[i]->rate, 0.f, 1.f);
    res &= ec3_limit(&cc[i]->noise_gate, 0.f, 100000000.f);
  }
  if (c->filter.coarse.length_blocks < c->filter.coarse_initial.length_blocks) {
    c->filter.coarse_initial.length_blocks = c->filter.coarse.length_blocks;
    res = false;
  }
  res &= ec3_limit(&c->filter.config_change_duration_blocks, 0, 100000);
  res &= ec3_limit(&c->filter.initial_state_seconds, 0.f, 100.f);
  res &= ec3_limit(&c->filter.coarse_reset_hangover_blocks, 0, 250000);
  res &= ec3_limit(&c->erle.min, 1.f, 100000.f);
  res &= ec3_limit(&c->erle.max_l, 1.f, 100000.f);
  res &= ec3_limit(&c->erle.max_h, 1.f, 100000.f);
  if (c->erle.min > c->erle.max_l || c->erle.min > c->erle.max_h) {
    c->erle.min = std::min(c->erle.max_l, c->erle.max_h);
    res = false;
  }
  res &= ec3_limit(&c->erle.num_sections, 1, c->filter.refined.length_blocks);
  res &= ec3_limit(&c->ep_strength.default_gain, 0.f, 1000000.f);
  res &= ec3_limit(&c->ep_strength.default_len, -1.f, 1.f);
  res &= ec3_limit(&c->ep_strength.nearend_len, -1.0f, 1.0f);
  const float kMaxPower = 32768.f * 32768.f;
  res &= ec3_limit(&c->echo_audibility.low_render_limit, 0.f, kMaxPower);
  res &= ec3_limit(&c->echo_audibility.normal_render_limit, 0.f, kMaxPower);
  res &= ec3_limit(&c->echo_audibility.floor_power, 0.f, kMaxPower);
  res &= ec3_limit(&c->echo_audibility.audibility_threshold_lf, 0.f, kMaxPower);
  res &= ec3_limit(&c->echo_audibility.audibility_threshold_mf, 0.f, kMaxPower);
  res &= ec3_limit(&c->echo_audibility.audibility_threshold_hf, 0.f, kMaxPower);
  res &= ec3_limit(&c->render_levels.active_render_limit, 0.f, kMaxPower);
  res &= ec3_limit(&c->render_levels.poor_excitation_render_limit, 0.f, kMaxPower);
  res &= ec3_limit(&c->render_levels.poor_excitation_render_limit_ds8, 0.f, kMaxPower);
  res &= ec3_limit(&c->echo_model.noise_floor_hold, 0, 1000);
  res &= ec3_limit(&c->echo_model.min_noise_floor_power, 0.f, 2000000.f);
  res &= ec3_limit(&c->echo_model.stationary_gate_slope, 0.f, 1000000.f);
  res &= ec3_limit(&c->echo_model.noise_gate_power, 0.f, 1000000.f);
  res &= ec3_limit(&c->echo_model.noise_gate_slope, 0.f, 1000000.f);
  res &= ec3_limit(&c->echo_model.render_pre_window_size, 0, 100);
  res &= ec3_limit(&c->echo_model.render_post_window_size, 0, 100);
  res &= ec3_limit(&c->comfort_noise.noise_floor_dbfs, -200.f, 0.f);
  res &= ec3_limit(&c->suppressor.nearend_average_blocks, 1, 5000);
  WapEc3Tuning* tc[2] = {&c->suppressor.normal_tuning, &c->suppressor.nearend_tuning};
  for (int i = 0; i < 2; ++i) {
    res &= ec3_limit(&tc[i]->mask_lf.enr_transparent, 0.f, 100.f);
    res &= ec3_limit(&tc[i]->mask_lf.enr_suppress, 0.f, 100.f);
    res &= ec3_limit(&tc[i]->mask_lf.emr_transparent, 0.f, 100.f);
    res &= ec3_limit(&tc[i]->mask_hf.enr_transparent, 0.f, 100.f);
    res &= ec3_limit(&tc[i]->mask_hf.enr_suppress, 0.f, 100.f);
    res &= ec3_limit(&tc[i]->mask_hf.emr_transparent, 0.f, 100.f);
    res &= ec3_limit(&tc[i]->max_inc_factor, 0.f, 100.f);
    res &= ec3_limit(&tc[i]->max_dec_factor_lf, 0.f, 100.f);
  }
  res &= ec3_limit(&c->suppressor.last_permanent_lf_smoothing_band, 0, 64);
  res &= ec3_limit(&c->suppressor.last_lf_smoothing_band, 0, 64);
  res &= ec3_limit(&c->suppressor.last_lf_band, 0, 63);
  res &= ec3_limit(&c->suppressor.first_hf_band, c->suppressor.last_lf_band + 1, 64);
  res &= ec3_limit(&c->suppressor.dominant_nearend_detection.enr_threshold, 0.f, 1000000.f);
  res &= ec3_limit(&c->suppressor.dominant_nearend_detection.snr_threshold, 0.f, 1000000.f);
  res &= ec3_limit(&c->suppressor.dominant_nearend_detection.hold_duration, 0, 10000);
  res &= ec3_limit(&c->suppressor.dominant_nearend_detection.trigger_threshold, 0, 10000);
  res &= ec3_limit(&c->suppressor.subband_nearend_detection.nearend_average_blocks, 1, 1024);
  res &= ec3_limit(&c->suppressor.subband_nearend_detection.subband1.low, 0, 65);
  res &= ec3_limit(&c->suppressor.subband_nearend_detection.subband1.high, c->suppressor.subband_nearend_detection.subband1.low, 65);
  res &= ec3_limit(&c->suppressor.subband_nearend_detection.subband2.low, 0, 65);
  res &= ec3_limit(&c->suppressor.subband_nearend_detection.subband2.high, c->suppressor.subband_nearend_detection.subband2.low, 65);
  res &= ec3_limit(&c->suppressor.subband_nearend_detection.nearend_threshold, 0.f, 1.e24f);
  res &= ec3_limit(&c->suppressor.subband_nearend_detection.snr_threshold, 0.f, 1.e24f);
  res &= ec3_limit(&c->suppressor.high_bands_suppression.enr_threshold, 0.f, 1000000.f);
  res &= ec3_limit(&c->suppressor.high_bands_suppression.max_gain_during_echo, 0.f, 1.f);
  res &= ec3_limit(&c->suppressor.high_bands_suppression.anti_howling_activation_threshold, 0.f, kMaxPower);
  res &= ec3_limit(&c->suppressor.high_bands_suppression.anti_howling_gain, 0.f, 1.f);
  res &= ec3_limit(&c->suppressor.high_frequency_suppression.limiting_gain_band, 1, 64);
  res &= ec3_limit(&c->suppressor.high_frequency_suppression.bands_in_limiting_gain, 0,
                   64 - c->suppressor.high_frequency_suppression.limiting_gain_band);
  res &= ec3_limit(&c->suppressor.floor_first_increase, 0.f, 1000000.f);
  return res;
}

int grid_for(int n_streams) {
  const int wpb = 4;
  int blocks = (n_streams + wpb - 1) / wpb;
  return std::max(1, blocks);
}

WapError launch_tick(WapEngine* e, const int* d_slots, const int* d_delays, int uniform_delay, int n,
                     const void* d_render, const void* d_capture, void* d_out, WapSampleFormat fmt) {
  wap::TickArgs a{};
  a.states = e->d_states;
  a.upper = e->d_upper;
  a.slots = d_slots;
  a.delays_ms = d_delays;
  a.uniform_delay_ms = uniform_delay;
  a.n = n;
  a.render = d_render;
  a.capture = d_capture;
  a.out = d_out;
  a.fmt = (int)fmt;
  a.cfg = e->cfg;
  a.ep = e->ep;
  const int wpb = 4;
  const bool timing = e->timing;
  if (timing) cudaEventRecord(e->ev[0], e->stream);
  a.extra = e->d_extra;
  a.cap_delay = e->d_cap_delay;
  a.cap_delay_stride = e->cap_delay_stride;
  a.red = e->d_red;
  a.rs_capture1 = nullptr;
  a.mc = e->d_mc;
  a.mc_templates = e->d_mc_templates;
  a.mc_ns = e->d_mc_ns;
  a.ep_mc = e->ep_mc;
  a.mcp[0] = e->mcp[0];
  a.mcp[1] = e->mcp[1];
  if (e->cfg.mc) {
    // multi-channel legs: k_mc_front -> k_delay_rt -> k_mc_echo [-> k_mc_post]
    wap::launch_k_mc_front(grid_for(n), wpb * 32, (size_t)wpb * e->mc_front_floats * sizeof(float), e->stream, a, e->mc_front_floats);
    e->launches++;
    if (timing) cudaEventRecord(e->ev[1], e->stream);
    if (d_capture && e->cfg.aec_enabled) {
      wap::launch_k_delay_rt(grid_for(n), wpb * 32, (size_t)wpb * e->delay_scratch_floats * sizeof(float), e->stream, a,
                             e->delay_scratch_floats);
      e->launches++;
    }
    if (timing) cudaEventRecord(e->ev[2], e->stream);
    // k_mc_echo: ~18 KB of scratch per warp (12 warps per SM whatever the CTA size; 4-warp CTAs measured best)
    const int ew = e->mc_echo_wpb;
    wap::launch_k_mc_echo((n + ew - 1) / ew, ew * 32, (size_t)ew * e->mc_echo_floats * sizeof(float), e->stream, a, e->mc_echo_floats);
    e->launches++;
    if (e->cfg.num_bands == 3 && d_capture && e->cfg.aec_enabled) {
      wap::launch_k_mc_post(e->stream, a);
      e->launches++;
    }
    if (timing) {
      cudaEventRecord(e->ev[3], e->stream);
      WAP_CUDA(cudaEventSynchronize(e->ev[3]));
      for (int k = 0; k < 3; ++k) {
        float ms = 0.f;
        cudaEventElapsedTime(&ms, e->ev[k], e->ev[k + 1]);
        e->kernel_ms[k] += ms;
      }
      e->timed_ticks++;
    }
    WAP_CUDA(cudaGetLastError());
    return WapError::None;
  }
  if (e->d_rs) {   // some stream of the leg is resampled
    a.rs = e->d_rs;
    a.rs_kernel_in = e->d_rs_kernels;
    a.rs_kernel_out = e->d_rs_kernels + wap::kRsTableFloats;
    a.rs_kernel_render = e->d_rs_kernels + 2 * wap::kRsTableFloats;
    a.rs_ratio_in = e->rs_ratio_in;
    a.rs_ratio_out = e->rs_ratio_out;
    a.rs_ratio_render = e->rs_ratio_render;
  }
  if (e->cfg.pre_stage) {
    // API-rate frames -> processing-rate frames; k_front then reads those (already FloatS16).
    a.rs_capture1 = e->d_rs_capture1;
    a.rs_render = e->d_rs_render;
    a.rs_capture = e->d_rs_capture;
    WAP_LAUNCH(wap::k_resample, grid_for(n), wpb * 32, (size_t)wpb * 2 * wap::kRsMaxRequest * sizeof(float), e->stream, a);
    e->launches++;
    wap::TickArgs af = a;
    af.render = d_render ? e->d_rs_render : nullptr;
    af.capture = d_capture ? e->d_rs_capture : nullptr;
    af.fmt = 2;
    if (e->d_red && d_render) {   // before k_front: it latches "a capture frame has been seen"
      WAP_LAUNCH(wap::k_red_render, (n + 127) / 128, 128, 0, e->stream, af);
      e->launches++;
    }
    WAP_LAUNCH(wap::k_front, (n + 127) / 128, 128, 0, e->stream, af);
  } else {
    if (e->d_upper && e->cfg.num_bands == 3) {
      WAP_LAUNCH(wap::k_split, grid_for(n), wpb * 32, (size_t)wpb * 1120 * sizeof(float), e->stream, a);
      e->launches++;
    }
    if (e->d_red && d_render) {
      WAP_LAUNCH(wap::k_red_render, (n + 127) / 128, 128, 0, e->stream, a);
      e->launches++;
    }
    WAP_LAUNCH(wap::k_front, (n + 127) / 128, 128, 0, e->stream, a);
  }
  e->launches++;
  if (timing) cudaEventRecord(e->ev[1], e->stream);
  if (e->cfg.aec_enabled && d_capture) {
    const size_t smem_d = (size_t)wpb * e->delay_scratch_floats * sizeof(float);
    if (e->ec3_runtime) wap::launch_k_delay_rt(grid_for(n), wpb * 32, smem_d, e->stream, a, e->delay_scratch_floats);
    else wap::launch_k_delay(grid_for(n), wpb * 32, smem_d, e->stream, a, e->delay_scratch_floats);
    e->launches++;
  }
  if (timing) cudaEventRecord(e->ev[2], e->stream);
  const size_t smem_e = (size_t)wpb * e->echo_scratch_floats * sizeof(float) + (size_t)e->echo_smem_pad;
  if (e->ec3_runtime) wap::launch_k_echo_rt(e->echo_class, grid_for(n), wpb * 32, smem_e, e->stream, a, e->echo_scratch_floats);
  else wap::launch_k_echo(e->echo_class, grid_for(n), wpb * 32, smem_e, e->stream, a, e->echo_scratch_floats);
  e->launches++;
  if (e->d_upper && e->cfg.num_bands == 3 && d_capture) {  // PostFilter: 48 kHz only (post_filter.cc:44-52)
    WAP_LAUNCH(wap::k_post, (n + 127) / 128, 128, 0, e->stream, a);
    e->launches++;
  }
  if (timing) {
    cudaEventRecord(e->ev[3], e->stream);
    WAP_CUDA(cudaEventSynchronize(e->ev[3]));
    for (int k = 0; k < 3; ++k) {
      float ms = 0.f;
      cudaEventElapsedTime(&ms, e->ev[k], e->ev[k + 1]);
      e->kernel_ms[k] += ms;
    }
    e->timed_ticks++;
  }
  WAP_CUDA(cudaGetLastError());
  return WapError::None;
}

// Number of leg ranges the host-buffer tick is pipelined over (1 = plain copy/compute/copy).
// Legs of one full-occupancy wave of the tick's heavy kernel: 20 warps per SM for the AEC3 kernels (k_delay),
// 24 for the kernel class without AEC3 (6 CTAs of 4 warps), whose ticks are short enough for the PCIe copies
// to dominate the host-buffer path from much smaller batches on.
int wave_legs(const WapEngine* e) { return e->sm_count * (e->cfg.aec_enabled ? 80 : 24); }

int pipeline_chunks(WapEngine* e, int n) {
  int chunks = e->forced_chunks ? e->forced_chunks : (n >= 4 * wave_legs(e) ? 4 : 1);

  if (e->timing || n < 2 * chunks) chunks = 1;
  if (chunks > 1 && !e->copy_in) {
    bool ok = cudaStreamCreateWithFlags(&e->copy_in, cudaStreamNonBlocking) == cudaSuccess &&
              cudaStreamCreateWithFlags(&e->copy_out, cudaStreamNonBlocking) == cudaSuccess &&
              cudaEventCreateWithFlags(&e->ev_start, cudaEventDisableTiming) == cudaSuccess;
    for (int k = 0; ok && k < kMaxChunks; ++k)
      ok = cudaEventCreateWithFlags(&e->ev_in[k], cudaEventDisableTiming) == cudaSuccess &&
           cudaEventCreateWithFlags(&e->ev_done[k], cudaEventDisableTiming) == cudaSuccess;
    if (!ok) return 1;
  }
  return chunks;
}

std::mutex g_default_mu;
WapEngine* g_default_engines[8] = {nullptr};

}  // namespace

extern "C" {

const char* wap_version(void) {
  static char v[160];
  snprintf(v, sizeof(v), "wap_b200 0.2 (sm_100a; per-warp smem: k_delay %d B, k_echo %d B at 16 kHz; k_echo min blocks/SM %d)",
           wap::delay_scratch_floats() * 4, wap::echo_scratch_floats(1) * 4, wap::k_echo_min_blocks());
  return v;
}

WapConfig wap_config_default(void) {
  // Defaults of AudioProcessing::Config (api/audio/audio_processing.h:137-376).
  WapConfig c{};
  c.pipeline_maximum_internal_processing_rate = 32000;
  c.pipeline_capture_downmix_method = WapDownmixMethod::AverageChannels;
  c.pre_amplifier_fixed_gain_factor = 1.f;
  c.capture_level_adjustment_pre_gain_factor = 1.f;
  c.capture_level_adjustment_post_gain_factor = 1.f;
  c.analog_mic_gain_emulation_initial_level = 255;
  c.high_pass_filter_apply_in_full_band = true;
  c.echo_canceller_enforce_high_pass_filtering = true;
  c.noise_suppression_level = WapNoiseSuppressionLevel::Moderate;
  c.gain_controller2_adaptive_digital_headroom_db = 5.f;
  c.gain_controller2_adaptive_digital_max_gain_db = 50.f;
  c.gain_controller2_adaptive_digital_initial_gain_db = 15.f;
  c.gain_controller2_adaptive_digital_max_gain_change_db_per_second = 6.f;
  c.gain_controller2_adaptive_digital_max_output_noise_level_dbfs = -50.f;
  return c;
}

WapEchoCanceller3Config wap_echo_canceller3_config_default(void) { return ec3_config_default(); }
WapEchoCanceller3Config wap_echo_canceller3_config_default_multichannel(void) {
  // EchoCanceller3Config::CreateDefaultMultichannelConfig (echo_canceller3_config.cc:288-301)
  WapEchoCanceller3Config c = ec3_config_default();
  c.filter.coarse.length_blocks = 11;
  c.filter.coarse.rate = 0.95f;
  c.filter.coarse_initial.length_blocks = 11;
  c.filter.coarse_initial.rate = 0.95f;
  c.suppressor.normal_tuning.max_dec_factor_lf = 0.35f;
  c.suppressor.normal_tuning.max_inc_factor = 1.5f;
  return c;
}
size_t wap_echo_canceller3_config_sizeof(void) { return sizeof(WapEchoCanceller3Config); }
bool wap_echo_canceller3_config_validate(WapEchoCanceller3Config* config) { return config ? ec3_validate(config) : false; }
WapError wap_echo_canceller3_config_supported(const WapEchoCanceller3Config* config) {
  return config ? ec3_config_supported(*config) : WapError::NullPointer;
}

WapEngine* wap_engine_create(int cuda_device, int32_t max_streams, WapConfig config, WapStreamConfig fmt) {
  return wap_engine_create_with_aec3_config(cuda_device, max_streams, config, fmt, nullptr, nullptr);
}

WapEngine* wap_engine_create_with_aec3_config(int cuda_device, int32_t max_streams, WapConfig config, WapStreamConfig fmt,
                                              const WapEchoCanceller3Config* aec3_config,
                                              const WapEchoCanceller3Config* aec3_multichannel_config) {
  return wap_engine_create_with_formats(cuda_device, max_streams, config, fmt, fmt, fmt, aec3_config,
                                        aec3_multichannel_config);
}

WapEngine* wap_engine_create_with_formats(int cuda_device, int32_t max_streams, WapConfig config, WapStreamConfig input,
                                          WapStreamConfig output, WapStreamConfig reverse_input,
                                          const WapEchoCanceller3Config* aec3_config,
                                          const WapEchoCanceller3Config* aec3_multichannel_config) {
  if (check_device() != WapError::None) return nullptr;
  EngineConfig cfg{};
  const WapFormats formats{input, output, reverse_input};
  const WapStreamConfig fmt = input;
  WapError err = resolve_config(config, formats, &cfg);
  const WapEchoCanceller3Config aec3 = aec3_config ? *aec3_config : ec3_config_default();
  // AudioProcessingImpl::InitializeEchoController (audio_processing_impl.cc:1928-1943): the default
  // multichannel config only when the user set neither; a mono config alone serves both.
  const WapEchoCanceller3Config aec3_mc = aec3_multichannel_config ? *aec3_multichannel_config
                                          : (aec3_config ? *aec3_config : wap_echo_canceller3_config_default_multichannel());
  if (err == WapError::None && config.echo_canceller_enabled) err = ec3_config_supported(aec3);
  if (err == WapError::None && cfg.mc) {
    err = ec3_config_supported(aec3_mc);
    // the render high-pass filter and the fixed capture delay are built for the mono kernels only
    // (the boolean switches of the echo remover below are built for the mono kernels only)
    for (const WapEchoCanceller3Config* c : {&aec3, &aec3_mc}) {
      const WapEchoCanceller3Config d = ec3_config_default();
      if (c->ep_strength.echo_can_saturate != d.ep_strength.echo_can_saturate || c->ep_strength.bounded_erl ||
          c->ep_strength.erle_onset_compensation_in_dominant_nearend || !c->ep_strength.use_conservative_tail_frequency_response ||
          !c->erle.onset_detection || !c->erle.clamp_quality_estimate_to_zero || !c->erle.clamp_quality_estimate_to_one ||
          c->echo_removal_control.has_clock_drift || c->echo_removal_control.linear_and_stable_echo_path ||
          !c->suppressor.lf_smoothing_during_initial_phase || !c->suppressor.dominant_nearend_detection.use_during_initial_phase ||
          !c->suppressor.dominant_nearend_detection.use_unbounded_echo_spectrum || c->suppressor.conservative_hf_suppression ||
          c->suppressor.high_bands_suppression.max_gain_during_echo != 1.f ||
          c->filter.conservative_initial_phase || !c->filter.enable_coarse_filter_output_usage || !c->filter.use_linear_filter ||
          c->echo_model.render_pre_window_size != 1 || c->echo_model.render_post_window_size != 1 ||
          !c->echo_model.model_reverb_in_nonlinear_mode || c->suppressor.nearend_average_blocks != 4 ||
          c->render_levels.render_power_gain_db != 0.f || c->echo_audibility.use_stationarity_properties ||
          c->echo_audibility.use_stationarity_properties_at_init ||
          c->ep_strength.default_len < 0.f || c->ep_strength.nearend_len < 0.f || c->erle.num_sections != 1 ||
          !c->delay.detect_pre_echo || c->delay.use_external_delay_estimator)
        err = WapError::UnsupportedConfig;
    }
    if (aec3.suppressor.use_subband_nearend_detection || aec3_mc.suppressor.use_subband_nearend_detection ||
        aec3.filter.high_pass_filter_echo_reference || aec3_mc.filter.high_pass_filter_echo_reference ||
        aec3.delay.fixed_capture_delay_samples || aec3_mc.delay.fixed_capture_delay_samples)
      err = WapError::UnsupportedConfig;
    // ConfigSelector's CompatibleConfigs (config_selector.cc:23-48), plus what the engine fixes per engine
    // rather than per leg state: the delay-estimation and buffering parameters and the comfort-noise floor.
    if (err == WapError::None &&
        (aec3.multi_channel.detect_stereo_content != aec3_mc.multi_channel.detect_stereo_content ||
         aec3.multi_channel.stereo_detection_timeout_threshold_seconds !=
             aec3_mc.multi_channel.stereo_detection_timeout_threshold_seconds ||
         aec3.comfort_noise.noise_floor_dbfs != aec3_mc.comfort_noise.noise_floor_dbfs ||
         (aec3.delay.render_alignment_mixing.downmix && aec3.delay.render_alignment_mixing.adaptive_selection) ||
         (aec3_mc.delay.render_alignment_mixing.downmix && aec3_mc.delay.render_alignment_mixing.adaptive_selection) ||
         (aec3.delay.capture_alignment_mixing.downmix && aec3.delay.capture_alignment_mixing.adaptive_selection) ||
         (aec3_mc.delay.capture_alignment_mixing.downmix && aec3_mc.delay.capture_alignment_mixing.adaptive_selection)))
      err = WapError::UnsupportedConfig;
  }
  if (err != WapError::None) {
    fprintf(stderr, "[wap_b200] unsupported engine config (error %d)\n", (int)err);
    return nullptr;
  }
  // ComfortNoiseGenerator: GetNoiseFloorFactor(comfort_noise.noise_floor_dbfs) (comfort_noise_generator.cc:41-45)
  cfg.cng_noise_floor = 64.f * powf(10.f, (90.30899869919436f + aec3.comfort_noise.noise_floor_dbfs) * 0.1f);
  if (max_streams <= 0) return nullptr;
  WapEngine* e = new (std::nothrow) WapEngine;
  if (!e) return nullptr;
  e->device = cuda_device;
  e->capacity = max_streams;
  e->config = config;
  e->format = fmt;
  e->formats = formats;
  e->cfg = cfg;
  e->frame_len = input.sample_rate_hz / 100 * input.num_channels;
  e->out_len = output.sample_rate_hz / 100 * output.num_channels;
  e->render_len = reverse_input.sample_rate_hz / 100 * reverse_input.num_channels;
  e->aec3_config = aec3;
  e->ep = ec3_params_from_config(aec3);
  e->ec3_runtime = config.echo_canceller_enabled && (cfg.mc || !wap::same_ec3_params(e->ep, wap::ec3_default_params()));
  if (cfg.mc) {
    e->ep_mc = ec3_params_from_config(aec3_mc);
    const WapEchoCanceller3Config* cc[2] = {&aec3, &aec3_mc};
    // The detector is built from the config that is active at construction (echo_canceller3.cc:765-775).
    const WapEchoCanceller3Config& dc = aec3.multi_channel.detect_stereo_content ? aec3 : aec3_mc;
    for (int i = 0; i < 2; ++i) {
      wap::McParams& m = e->mcp[i];
      m.detect_stereo_content = dc.multi_channel.detect_stereo_content ? 1 : 0;
      m.detection_threshold = dc.multi_channel.stereo_detection_threshold;
      m.timeout_frames = dc.multi_channel.stereo_detection_timeout_threshold_seconds > 0
                             ? dc.multi_channel.stereo_detection_timeout_threshold_seconds * 100 : 0;
      m.hysteresis_frames = (int)(dc.multi_channel.stereo_detection_hysteresis_seconds * 100);
      const WapEc3AlignmentMixing& rm = cc[i]->delay.render_alignment_mixing;
      const WapEc3AlignmentMixing& cm = cc[i]->delay.capture_alignment_mixing;
      m.render_mix_downmix = rm.downmix; m.render_mix_adaptive = rm.adaptive_selection;
      m.render_mix_prefer_first_two = rm.prefer_first_two_channels; m.render_mix_threshold = rm.activity_power_threshold;
      m.capture_mix_downmix = cm.downmix; m.capture_mix_adaptive = cm.adaptive_selection;
      m.capture_mix_prefer_first_two = cm.prefer_first_two_channels; m.capture_mix_threshold = cm.activity_power_threshold;
    }
    e->mc_front_floats = wap::k_mc_front_scratch_floats();
    e->mc_echo_floats = wap::k_mc_echo_scratch_floats();
    if (const char* w = getenv("WAP_MC_ECHO_WPB")) e->mc_echo_wpb = std::max(1, std::min(4, atoi(w)));   // tuning knob
  }
  e->echo_class = wap::echo_class_of(cfg);
  e->echo_scratch_floats = e->ec3_runtime ? wap::k_echo_scratch_floats_rt(cfg.num_bands, e->echo_class)
                                          : wap::k_echo_scratch_floats(cfg.num_bands, e->echo_class);
  e->delay_scratch_floats = e->ec3_runtime ? wap::k_delay_scratch_floats_rt() : wap::k_delay_scratch_floats();
  bool ok = cudaSetDevice(cuda_device) == cudaSuccess &&
            cudaDeviceGetAttribute(&e->sm_count, cudaDevAttrMultiProcessorCount, cuda_device) == cudaSuccess &&
            cudaStreamCreateWithFlags(&e->stream, cudaStreamNonBlocking) == cudaSuccess &&
            cudaMalloc((void**)&e->d_states, (size_t)max_streams * sizeof(StreamState)) == cudaSuccess &&
            cudaMalloc((void**)&e->d_template, sizeof(StreamState)) == cudaSuccess;
  if (ok && cfg.aec_enabled && cfg.num_bands >= 2) {
    const size_t ub = (size_t)max_streams * sizeof(wap::UpperBandState);
    ok = cudaMalloc((void**)&e->d_upper, ub) == cudaSuccess && cudaMemset(e->d_upper, 0, ub) == cudaSuccess;
  }
  if (ok && (cfg.pre_stage || cfg.resample_out || cfg.fullband_out)) {
    const size_t rb = (size_t)max_streams * wap::kRsPerLeg * sizeof(wap::ResamplerState);
    const int pf = wap::kFrame * cfg.num_bands;
    e->rs_ratio_in = (double)cfg.api_frame * 1.0 / pf;  // source_frames * 1.0 / destination_frames
    e->rs_ratio_out = (double)pf * 1.0 / cfg.out_frame;
    e->rs_ratio_render = (double)cfg.render_frame * 1.0 / pf;
    std::vector<float> tables(3 * wap::kRsTableFloats);
    build_sinc_kernel(e->rs_ratio_in, tables.data());
    build_sinc_kernel(e->rs_ratio_out, tables.data() + wap::kRsTableFloats);
    build_sinc_kernel(e->rs_ratio_render, tables.data() + 2 * wap::kRsTableFloats);
    ok = cudaMalloc((void**)&e->d_rs, rb) == cudaSuccess && cudaMemset(e->d_rs, 0, rb) == cudaSuccess &&
         cudaMalloc((void**)&e->d_rs_kernels, tables.size() * sizeof(float)) == cudaSuccess &&
         cudaMemcpy(e->d_rs_kernels, tables.data(), tables.size() * sizeof(float), cudaMemcpyHostToDevice) == cudaSuccess;
  }
  if (ok && cfg.mc) {
    wap::McTemplates* t = new wap::McTemplates;
    wap::init_mc_templates(*t, e->ep, e->ep_mc);
    ok = cudaMalloc((void**)&e->d_mc, (size_t)max_streams * sizeof(wap::McState)) == cudaSuccess &&
         cudaMalloc((void**)&e->d_mc_templates, sizeof(wap::McTemplates)) == cudaSuccess &&
         cudaMemcpy(e->d_mc_templates, t, sizeof(wap::McTemplates), cudaMemcpyHostToDevice) == cudaSuccess &&
         wap::set_k_mc_smem(4 * e->mc_front_floats * (int)sizeof(float), 4 * e->mc_echo_floats * (int)sizeof(float)) == cudaSuccess;
    delete t;
    if (ok && cfg.ns_enabled) {
      wap::NsState* nt = new wap::NsState;
      wap::init_ns_state(*nt);
      ok = cudaMalloc((void**)&e->d_mc_ns, (size_t)max_streams * wap::kMcCh * sizeof(wap::NsState)) == cudaSuccess &&
           cudaMalloc((void**)&e->d_mc_ns_template, sizeof(wap::NsState)) == cudaSuccess &&
           cudaMemcpy(e->d_mc_ns_template, nt, sizeof(wap::NsState), cudaMemcpyHostToDevice) == cudaSuccess;
      delete nt;
    }
  }
  if (ok && cfg.aec_enabled && e->ep.fixed_capture_delay_samples > 0) {
    // BlockDelayBuffer (block_delay_buffer.cc:22-31): num_bands rings of `delay` samples, zero
    e->cap_delay_stride = cfg.num_bands * e->ep.fixed_capture_delay_samples + 4;
    const size_t cb = (size_t)max_streams * e->cap_delay_stride * sizeof(float);
    ok = cudaMalloc((void**)&e->d_cap_delay, cb) == cudaSuccess && cudaMemset(e->d_cap_delay, 0, cb) == cudaSuccess;
  }
  if (ok && cfg.channels == 2 && !cfg.mc) {
    const size_t xb = (size_t)max_streams * sizeof(wap::ExtraChannelState);
    ok = cudaMalloc((void**)&e->d_extra, xb) == cudaSuccess && cudaMemset(e->d_extra, 0, xb) == cudaSuccess;
  }
  if (ok) {
    StreamState* tmpl = new StreamState;
    // legs of a multi-channel engine without stereo detection start on the multichannel config
    wap::init_stream_state(*tmpl, (cfg.mc && !e->mcp[0].detect_stereo_content) ? e->ep_mc : e->ep);
    tmpl->agc2.gain_last = tmpl->agc2.gain_current = cfg.agc2_fixed_gain;
    if (cfg.levels_enabled) {  // InitializeCaptureLevelsAdjuster (audio_processing_impl.cc:2108-2130)
      float pre_gain = 1.f;
      if (config.pre_amplifier_enabled) pre_gain *= config.pre_amplifier_fixed_gain_factor;
      if (config.capture_level_adjustment_enabled) pre_gain *= config.capture_level_adjustment_pre_gain_factor;
      tmpl->levels.pre_prev = tmpl->levels.pre_target = pre_gain;
      tmpl->levels.post_prev = tmpl->levels.post_target = config.capture_level_adjustment_post_gain_factor;
    }
    ok = cudaMemcpy(e->d_template, tmpl, sizeof(StreamState), cudaMemcpyHostToDevice) == cudaSuccess;
    delete tmpl;
  }
  // tuning knob: extra dynamic shared memory per k_echo CTA = fewer CTAs (legs in flight) per SM
  if (const char* pad = getenv("WAP_ECHO_SMEM_PAD_KB")) e->echo_smem_pad = std::max(0, std::min(160, atoi(pad))) * 1024;
  const size_t smem_e = (size_t)4 * e->echo_scratch_floats * sizeof(float) + (size_t)e->echo_smem_pad;
  const size_t smem_d = (size_t)4 * e->delay_scratch_floats * sizeof(float);
  if (ok && smem_e > 48 * 1024) ok = (e->ec3_runtime ? wap::set_k_echo_smem_rt((int)smem_e) : wap::set_k_echo_smem((int)smem_e)) == cudaSuccess;
  if (ok && smem_d > 48 * 1024) ok = (e->ec3_runtime ? wap::set_k_delay_smem_rt((int)smem_d) : wap::set_k_delay_smem((int)smem_d)) == cudaSuccess;
  if (!ok) {
    fprintf(stderr, "[wap_b200] engine allocation failed: %s\n", cudaGetErrorString(cudaGetLastError()));
    wap_engine_destroy(e);
    return nullptr;
  }
  e->leg_delay_ms.assign(max_streams, 0);
  e->leg_delay_set.assign(max_streams, 0);
  e->free_slots.reserve(max_streams);
  for (int i = max_streams - 1; i >= 0; --i) e->free_slots.push_back(i);
  return e;
}

void wap_engine_destroy(WapEngine* e) {
  if (!e) return;
  cudaSetDevice(e->device);
  if (e->stream) cudaStreamSynchronize(e->stream);
  cudaFree(e->d_states);
  cudaFree(e->d_upper);
  cudaFree(e->d_rs);
  cudaFree(e->d_rs_kernels);
  cudaFree(e->d_rs_render);
  cudaFree(e->d_rs_capture);
  cudaFree(e->d_rs_capture1);
  cudaFree(e->d_extra);
  cudaFree(e->d_cap_delay);
  cudaFree(e->d_red);
  cudaFree(e->d_mc);
  cudaFree(e->d_mc_templates);
  cudaFree(e->d_mc_ns);
  cudaFree(e->d_mc_ns_template);
  cudaFree(e->d_template);
  cudaFree(e->d_render);
  cudaFree(e->d_capture);
  cudaFree(e->d_out);
  cudaFree(e->d_slots);
  cudaFree(e->d_delays);
  if (e->h_pinned) cudaFreeHost(e->h_pinned);
  for (int k = 0; k < 4; ++k) if (e->ev[k]) cudaEventDestroy(e->ev[k]);
  for (int k = 0; k < kMaxChunks; ++k) {
    if (e->ev_in[k]) cudaEventDestroy(e->ev_in[k]);
    if (e->ev_done[k]) cudaEventDestroy(e->ev_done[k]);
  }
  if (e->ev_start) cudaEventDestroy(e->ev_start);
  if (e->copy_in) cudaStreamDestroy(e->copy_in);
  if (e->copy_out) cudaStreamDestroy(e->copy_out);
  if (e->stream) cudaStreamDestroy(e->stream);
  delete e;
}

// EXT: AudioProcessingBuilder::SetEchoDetector(CreateEchoDetector()) for every leg of the engine.
WapError wap_engine_enable_echo_detector(WapEngine* e) {
  if (!e) return WapError::NullPointer;
  std::lock_guard<std::recursive_mutex> lk(e->mu);
  if (e->d_red) return WapError::None;
  // built for the mono kernel classes (with or without AEC3; without it the render stream is not
  // resampled, so only at a native rate); before the first leg joins
  if (e->cfg.mc || (!e->cfg.aec_enabled && e->cfg.pre_stage)) return WapError::UnsupportedConfig;
  if (e->free_slots.size() != (size_t)e->capacity) return WapError::BadStreamParameter;
  WAP_CUDA(cudaSetDevice(e->device));
  const size_t bytes = (size_t)e->capacity * sizeof(wap::EchoDetectorState);
  // The capture side of the detector is compiled into the run-time-parameter kernel instances only (the
  // default-config instances stay as they are): the engine switches to those.
  if (!e->ec3_runtime) {
    e->ec3_runtime = true;
    e->echo_scratch_floats = wap::k_echo_scratch_floats_rt(e->cfg.num_bands, e->echo_class);
    e->delay_scratch_floats = wap::k_delay_scratch_floats_rt();
    const size_t smem_e = (size_t)4 * e->echo_scratch_floats * sizeof(float) + (size_t)e->echo_smem_pad;
    const size_t smem_d = (size_t)4 * e->delay_scratch_floats * sizeof(float);
    if (smem_e > 48 * 1024) WAP_CUDA(wap::set_k_echo_smem_rt((int)smem_e));
    if (smem_d > 48 * 1024) WAP_CUDA(wap::set_k_delay_smem_rt((int)smem_d));
  }
  WAP_CUDA(cudaMalloc((void**)&e->d_red, bytes));
  WAP_CUDA(cudaMemset(e->d_red, 0, bytes));
  return WapError::None;
}

WapError wap_engine_create_streams(WapEngine* e, int32_t n, WapAudioProcessing** out) {
  if (!e || !out) return WapError::NullPointer;
  std::lock_guard<std::recursive_mutex> lk(e->mu);
  if (n <= 0 || (size_t)n > e->free_slots.size()) return WapError::BadStreamParameter;
  WAP_CUDA(cudaSetDevice(e->device));
  std::vector<int> slots(n);
  for (int i = 0; i < n; ++i) {
    slots[i] = e->free_slots.back();
    e->free_slots.pop_back();
  }
  int* d_slots = nullptr;
  // Any CUDA failure below hands the slots back and frees the scratch list.
  auto init = [&]() -> WapError {
    WAP_CUDA(cudaMalloc((void**)&d_slots, (size_t)n * sizeof(int)));
    WAP_CUDA(cudaMemcpy(d_slots, slots.data(), (size_t)n * sizeof(int), cudaMemcpyHostToDevice));
    WAP_LAUNCH(wap::k_init_slots, dim3(16, std::min(n, 4096)), 256, 0, e->stream, e->d_states,
               (const StreamState*)e->d_template, (const int*)d_slots, (int)n);
    if (e->d_upper)
      for (int i = 0; i < n; ++i) WAP_CUDA(cudaMemsetAsync(&e->d_upper[slots[i]], 0, sizeof(wap::UpperBandState), e->stream));
    if (e->d_extra)
      for (int i = 0; i < n; ++i) WAP_CUDA(cudaMemsetAsync(&e->d_extra[slots[i]], 0, sizeof(wap::ExtraChannelState), e->stream));
    if (e->d_cap_delay)
      for (int i = 0; i < n; ++i)
        WAP_CUDA(cudaMemsetAsync(e->d_cap_delay + (size_t)slots[i] * e->cap_delay_stride, 0, e->cap_delay_stride * sizeof(float), e->stream));
    if (e->d_red)
      for (int i = 0; i < n; ++i) WAP_CUDA(cudaMemsetAsync(&e->d_red[slots[i]], 0, sizeof(wap::EchoDetectorState), e->stream));
    if (e->d_mc) {
      WAP_LAUNCH(wap::k_mc_init_slots, dim3(32, std::min(n, 2048)), 256, 0, e->stream, e->d_mc,
                 (const wap::McTemplates*)e->d_mc_templates, (const int*)d_slots, (int)n,
                 e->mcp[0].detect_stereo_content ? 0 : 1);
      e->launches++;
      if (e->d_mc_ns)
        for (int i = 0; i < n * wap::kMcCh; ++i)
          WAP_CUDA(cudaMemcpyAsync(&e->d_mc_ns[(size_t)slots[i / wap::kMcCh] * wap::kMcCh + i % wap::kMcCh], e->d_mc_ns_template,
                                   sizeof(wap::NsState), cudaMemcpyDeviceToDevice, e->stream));
    }
    if (e->d_rs)
      for (int i = 0; i < n; ++i)
        WAP_CUDA(cudaMemsetAsync(&e->d_rs[(size_t)slots[i] * wap::kRsPerLeg], 0, wap::kRsPerLeg * sizeof(wap::ResamplerState), e->stream));
    e->launches++;
    WAP_CUDA(cudaStreamSynchronize(e->stream));
    return WapError::None;
  };
  const WapError init_err = init();
  cudaFree(d_slots);
  if (init_err != WapError::None) {
    for (int i = n - 1; i >= 0; --i) e->free_slots.push_back(slots[i]);
    return init_err;
  }
  e->last_handles.clear();
  for (int i = 0; i < n; ++i) {
    e->leg_delay_ms[slots[i]] = 0;
    e->leg_delay_set[slots[i]] = 0;
    WapAudioProcessing* h = new WapAudioProcessing;
    h->engine = e;
    h->slot = slots[i];
    h->config = e->config;
    out[i] = h;
  }
  return WapError::None;
}

size_t wap_engine_state_bytes_per_stream(const WapEngine* e) {
  return sizeof(StreamState) + ((e && e->d_mc) ? sizeof(wap::McState) : 0) +
         ((e && e->d_mc_ns) ? wap::kMcCh * sizeof(wap::NsState) : 0);
}

WapError wap_engine_synchronize(WapEngine* e) {
  if (!e) return WapError::NullPointer;
  WAP_CUDA(cudaStreamSynchronize(e->stream));
  return WapError::None;
}
WapError wap_engine_set_pipeline_chunks(WapEngine* e, int32_t chunks) {
  if (!e) return WapError::NullPointer;
  if (chunks < 0 || chunks > kMaxChunks) return WapError::BadStreamParameter;
  e->forced_chunks = chunks;
  return WapError::None;
}
void* wap_engine_cuda_stream(WapEngine* e) { return e ? (void*)e->stream : nullptr; }
int64_t wap_engine_launch_count(const WapEngine* e) { return e ? e->launches : 0; }
int32_t wap_engine_uses_runtime_aec3_parameters(const WapEngine* e) { return e && e->ec3_runtime ? 1 : 0; }

// Per-tick host bookkeeping in front of the launches: slot list, un/mute flags, stream delays.
static WapError prepare_tick(WapEngine* e, WapAudioProcessing* const* handles, int32_t n, const int** d_delays_out,
                             int* uniform_delay_out) {
  if (!e || !handles) return WapError::NullPointer;
  if (n <= 0) return WapError::BadStreamParameter;
  WAP_CUDA(cudaSetDevice(e->device));
  WapError err = ensure_staging(e, n);
  if (err != WapError::None) return err;
  // slot list (cached across ticks while the caller passes the same handle array)
  const bool same = e->last_handles.size() == (size_t)n &&
                    memcmp(handles, e->last_handles.data(), (size_t)n * sizeof(handles[0])) == 0;
  if (!same) {
    for (int i = 0; i < n; ++i)
      if (!handles[i] || handles[i]->engine != e) return WapError::BadStreamParameter;
    // Two entries for one leg would let two warps update the same state slab.
    std::vector<unsigned char> seen(e->capacity, 0);
    for (int i = 0; i < n; ++i) {
      const int sl = handles[i]->slot;
      if (sl < 0 || sl >= e->capacity || seen[sl]) return WapError::BadStreamParameter;
      seen[sl] = 1;
    }
    e->last_slots.resize(n);
    for (int i = 0; i < n; ++i) e->last_slots[i] = handles[i]->slot;
    e->last_handles.assign(handles, handles + n);
    WAP_CUDA(cudaMemcpyAsync(e->d_slots, e->last_slots.data(), (size_t)n * sizeof(int), cudaMemcpyHostToDevice, e->stream));
    WAP_CUDA(cudaStreamSynchronize(e->stream));
  }
  const int* slots = e->last_slots.data();
  int uniform_delay = e->leg_delay_set[slots[0]] ? e->leg_delay_ms[slots[0]] : -1;
  bool uniform = true;
  for (int i = 1; i < n; ++i) {
    const int d = e->leg_delay_set[slots[i]] ? e->leg_delay_ms[slots[i]] : -1;
    uniform = uniform && d == uniform_delay;
  }
  if (e->dirty_legs > 0) {  // rare: un/mute events
    for (int i = 0; i < n; ++i) {
      WapAudioProcessing* h = handles[i];
      StreamState* slab = &e->d_states[h->slot];
      if (h->pre_gain_dirty) WAP_CUDA(cudaMemcpyAsync(&slab->levels.pre_target, &h->pre_gain_target, sizeof(float), cudaMemcpyHostToDevice, e->stream));
      if (h->post_gain_dirty) WAP_CUDA(cudaMemcpyAsync(&slab->levels.post_target, &h->post_gain_target, sizeof(float), cudaMemcpyHostToDevice, e->stream));
      if (h->playout_volume_dirty) WAP_CUDA(cudaMemcpyAsync(&slab->levels.playout_volume, &h->playout_volume, sizeof(int), cudaMemcpyHostToDevice, e->stream));
      if (h->agc2_gain_dirty) {
        WAP_CUDA(cudaMemcpyAsync(&slab->agc2.gain_current, &h->agc2_gain_factor, sizeof(float), cudaMemcpyHostToDevice, e->stream));
        if (h->agc2_reset_limiter) {
          static const int one = 1;
          WAP_CUDA(cudaMemcpyAsync(&slab->agc2.reset_limiter, &one, sizeof(int), cudaMemcpyHostToDevice, e->stream));
          h->agc2_reset_limiter = false;
        }
      }
      const int v = h->capture_output_used ? 1 : 0;
      if (h->capture_output_used_dirty) WAP_CUDA(cudaMemcpyAsync(&slab->capture_output_used, &v, sizeof(int), cudaMemcpyHostToDevice, e->stream));
      const int n_dirty = (int)h->pre_gain_dirty + (int)h->post_gain_dirty + (int)h->playout_volume_dirty + (int)h->capture_output_used_dirty +
                          (int)h->agc2_gain_dirty;
      if (!n_dirty) continue;
      WAP_CUDA(cudaStreamSynchronize(e->stream));
      h->pre_gain_dirty = h->post_gain_dirty = h->playout_volume_dirty = h->capture_output_used_dirty = h->agc2_gain_dirty = false;
      e->dirty_legs -= n_dirty;
    }
  }
  const int* d_delays = nullptr;
  if (!uniform) {
    std::vector<int> dl(n);
    for (int i = 0; i < n; ++i) dl[i] = e->leg_delay_set[slots[i]] ? e->leg_delay_ms[slots[i]] : -1;
    WAP_CUDA(cudaMemcpyAsync(e->d_delays, dl.data(), (size_t)n * sizeof(int), cudaMemcpyHostToDevice, e->stream));
    WAP_CUDA(cudaStreamSynchronize(e->stream));
    d_delays = e->d_delays;
  }
  *d_delays_out = d_delays;
  *uniform_delay_out = uniform_delay;
  return WapError::None;
}

WapError wap_process_streams_device(WapEngine* e, WapAudioProcessing* const* handles, int32_t n,
                                    const void* d_render, const void* d_capture, void* d_out,
                                    WapSampleFormat fmt) {
  if (!e) return WapError::NullPointer;
  std::lock_guard<std::recursive_mutex> lk(e->mu);
  const int* d_delays = nullptr;
  int uniform_delay = -1;
  WapError err = prepare_tick(e, handles, n, &d_delays, &uniform_delay);
  if (err != WapError::None) return err;
  err = launch_tick(e, e->d_slots, d_delays, uniform_delay, n, d_render, d_capture, d_out, fmt);
  // was_stream_delay_set is one-shot: every capture frame clears it (audio_processing_impl.cc:1556),
  // through this entry point as through the host-buffer one.
  if (d_capture)
    for (int i = 0; i < n; ++i) e->leg_delay_set[e->last_slots[i]] = 0;
  return err;
}

WapError wap_process_streams(WapAudioProcessing* const* handles, int32_t n, const void* render,
                             const void* capture, void* out, WapSampleFormat fmt, WapError* per_stream_err) {
  if (!handles || !capture || !out) return WapError::NullPointer;
  if (n <= 0 || !handles[0]) return WapError::BadStreamParameter;
  WapEngine* e = handles[0]->engine;
  if (!e) return WapError::BadStreamParameter;  // a wap_create() handle that has not processed yet has no engine
  std::lock_guard<std::recursive_mutex> lk(e->mu);
  WAP_CUDA(cudaSetDevice(e->device));
  WapError err = ensure_staging(e, n);
  if (err != WapError::None) return err;
  const size_t esz = fmt == WapSampleFormat::I16 ? sizeof(int16_t) : sizeof(float);
  // the three streams of a leg may have different formats (wap_engine_create_with_formats)
  const size_t bytes = (size_t)n * e->frame_len * esz;        // capture input
  const size_t rbytes = (size_t)n * e->render_len * esz;      // render
  const size_t obytes = (size_t)n * e->out_len * esz;         // output
  unsigned char* hp = static_cast<unsigned char*>(e->h_pinned);
  unsigned char* hp_capture = hp + e->staged_streams * e->render_len * sizeof(float);
  unsigned char* hp_out = hp_capture + e->staged_streams * e->frame_len * sizeof(float);
  // Page-locked caller buffers are copied from / to directly; pageable ones go through the
  // engine's pinned staging area.
  const void* src_r = render;
  const void* src_c = capture;
  if (render && !is_pinned_host(render)) {
    memcpy(hp, render, rbytes);
    src_r = hp;
  }
  if (!is_pinned_host(capture)) {
    memcpy(hp_capture, capture, bytes);
    src_c = hp_capture;
  }
  const bool out_pinned = is_pinned_host(out);
  void* dst_o = out_pinned ? out : (void*)hp_out;
  const int chunks = pipeline_chunks(e, n);
  if (chunks <= 1) {
    if (render) WAP_CUDA(cudaMemcpyAsync(e->d_render, src_r, rbytes, cudaMemcpyHostToDevice, e->stream));
    WAP_CUDA(cudaMemcpyAsync(e->d_capture, src_c, bytes, cudaMemcpyHostToDevice, e->stream));
    err = wap_process_streams_device(e, handles, n, render ? e->d_render : nullptr, e->d_capture, e->d_out, fmt);
    if (err != WapError::None) return err;
    WAP_CUDA(cudaMemcpyAsync(dst_o, e->d_out, obytes, cudaMemcpyDeviceToHost, e->stream));
    WAP_CUDA(cudaStreamSynchronize(e->stream));
  } else {
    // Large batch: the legs are cut into `chunks` ranges; host->device copies run on one copy
    // stream, the tick kernels of each range on the engine stream, device->host copies on a
    // second copy stream, so the PCIe traffic of one range hides behind the kernels of another.
    const size_t leg_bytes = (size_t)e->frame_len * esz, rleg_bytes = (size_t)e->render_len * esz,
                 oleg_bytes = (size_t)e->out_len * esz;
    WAP_CUDA(cudaEventRecord(e->ev_start, e->stream));
    WAP_CUDA(cudaStreamWaitEvent(e->copy_in, e->ev_start, 0));
    // Range boundaries.  Only two copies are ever exposed: the host->device copy of the FIRST range
    // (nothing to compute yet) and the device->host copy of the LAST one (nothing left to compute), so
    // those two ranges are one full-occupancy wave of the warp-per-leg kernels (4 and 5 CTAs of 4 legs
    // per SM) and the middle takes the rest in at most `chunks - 3` (>= 1) ranges of whole waves: every extra
    // range costs the tail of two kernels, so there are as few as the overlap needs.
    // (engines without AEC3 are copy-bound: first / last range of a fifth of the batch when that is more than
    // one wave -- measured on B200 at 16,384 and 65,536 NS-only 48 kHz legs against equal ranges and against
    // one-wave edges: 175 k / 210 k legs through host buffers instead of 137 k / 156 k)
    int wave = e->forced_chunks ? 4 : wave_legs(e);
    if (!e->forced_chunks && !e->cfg.aec_enabled) wave = std::max(wave, n / 5 / 128 * 128);
    int bounds[kMaxChunks + 1];
    int nr = 0;
    bounds[0] = 0;
    if (e->forced_chunks || n < 3 * wave) {
      const int per = ((n + chunks - 1) / chunks + wave - 1) / wave * wave;
      for (int off = per; off < n && nr < kMaxChunks - 1; off += per) bounds[++nr] = off;
    } else {
      bounds[++nr] = wave;
      const int mid = n - 2 * wave;
      const int parts = std::max(1, std::min(chunks - 3, mid / wave));
      const int per = ((mid + parts - 1) / parts + wave - 1) / wave * wave;
      for (int off = wave + per; off < n - wave && nr < kMaxChunks - 2; off += per) bounds[++nr] = off;
      bounds[++nr] = n - wave;
    }
    bounds[++nr] = n;
    for (int c = 0; c < nr; ++c) {
      const int off = bounds[c], cnt = bounds[c + 1] - off;
      const size_t bo = (size_t)off * leg_bytes, bc = (size_t)cnt * leg_bytes;
      if (render)
        WAP_CUDA(cudaMemcpyAsync((char*)e->d_render + (size_t)off * rleg_bytes, (const char*)src_r + (size_t)off * rleg_bytes,
                                 (size_t)cnt * rleg_bytes, cudaMemcpyHostToDevice, e->copy_in));
      WAP_CUDA(cudaMemcpyAsync((char*)e->d_capture + bo, (const char*)src_c + bo, bc, cudaMemcpyHostToDevice, e->copy_in));
      WAP_CUDA(cudaEventRecord(e->ev_in[c], e->copy_in));
    }
    // host bookkeeping runs while the first copies are in flight
    const int* d_delays = nullptr;
    int uniform_delay = -1;
    err = prepare_tick(e, handles, n, &d_delays, &uniform_delay);
    if (err != WapError::None) return err;
    for (int c = 0; c < nr; ++c) {
      const int off = bounds[c], cnt = bounds[c + 1] - off;
      const size_t bo = (size_t)off * leg_bytes, ro = (size_t)off * rleg_bytes, oo = (size_t)off * oleg_bytes;
      WAP_CUDA(cudaStreamWaitEvent(e->stream, e->ev_in[c], 0));
      err = launch_tick(e, e->d_slots + off, d_delays ? d_delays + off : nullptr, uniform_delay, cnt,
                        render ? (const void*)((const char*)e->d_render + ro) : nullptr, (const char*)e->d_capture + bo,
                        (char*)e->d_out + oo, fmt);
      if (err != WapError::None) return err;
      WAP_CUDA(cudaEventRecord(e->ev_done[c], e->stream));
      WAP_CUDA(cudaStreamWaitEvent(e->copy_out, e->ev_done[c], 0));
      WAP_CUDA(cudaMemcpyAsync((char*)dst_o + oo, (const char*)e->d_out + oo, (size_t)cnt * oleg_bytes, cudaMemcpyDeviceToHost,
                               e->copy_out));
    }
    for (int i = 0; i < n; ++i) e->leg_delay_set[e->last_slots[i]] = 0;  // audio_processing_impl.cc:1556
    WAP_CUDA(cudaStreamSynchronize(e->copy_out));
    WAP_CUDA(cudaStreamSynchronize(e->stream));
  }
  if (!out_pinned) memcpy(out, hp_out, obytes);
  if (per_stream_err) for (int i = 0; i < n; ++i) per_stream_err[i] = WapError::None;
  return WapError::None;
}

}  // extern "C"

extern "C" {

// AudioProcessingImpl::GetStatistics (audio_processing_impl.cc:1509-1518,2312-2320):
// the statistics of the capture frame that filled ApmStatsReporter's one-slot
// queue (EchoRemoverImpl::GetMetrics echo_remover.cc:247-252 and
// BlockProcessorImpl::GetMetrics block_processor.cc:222-227, evaluated on the
// scalars the tick kernel latched), else the previously returned ones.
WapError wap_get_statistics(const WapAudioProcessing* hc, WapStats* out) {
  if (!hc || !out) return WapError::NullPointer;
  WapAudioProcessing* h = const_cast<WapAudioProcessing*>(hc);
  if (h->engine && h->slot >= 0 && (h->engine->cfg.aec_enabled || h->engine->d_red)) {
    WapEngine* e = h->engine;
    std::lock_guard<std::recursive_mutex> lk(e->mu);
    WAP_CUDA(cudaSetDevice(e->device));
    WAP_CUDA(cudaStreamSynchronize(e->stream));
    WapStats st;
    memset(&st, 0, sizeof(st));
    bool slot_full = false;
    const int zero = 0;
    if (e->cfg.aec_enabled) {
      wap::Aec3Scalars s;
      WAP_CUDA(cudaMemcpy(&s, &e->d_states[h->slot].aec.s, sizeof(s), cudaMemcpyDeviceToHost));
      if (s.stats_slot_full) {
        slot_full = true;
        st.has_echo_return_loss = true;
        st.echo_return_loss = -10.0 * log10((double)s.stats_erl_time_domain);
        st.has_echo_return_loss_enhancement = true;
        // Log2TodB (aec3_common.cc:54-56): double product rounded to float.
        st.echo_return_loss_enhancement = (double)(float)(3.0102999566398121 * (double)s.stats_erle_log2);
        st.has_delay_ms = true;
        st.delay_ms = s.stats_delay_blocks * 4;  // block_size_ms = 4
        WAP_CUDA(cudaMemcpy(&e->d_states[h->slot].aec.s.stats_slot_full, &zero, sizeof(int), cudaMemcpyHostToDevice));
      }
    }
    if (e->d_red) {
      // the residual echo detector's metrics travel in the same slot (audio_processing_impl.cc:1499-1505);
      // with AEC3 the two slots fill and empty together
      wap::EchoDetectorState* dr = &e->d_red[h->slot];
      struct { int full, valid; float likelihood, recent_max; } rs;
      static_assert(offsetof(wap::EchoDetectorState, slot_recent_max) - offsetof(wap::EchoDetectorState, slot_full) == 12,
                    "the slot members are read as one block");
      WAP_CUDA(cudaMemcpy(&rs, &dr->slot_full, sizeof(rs), cudaMemcpyDeviceToHost));
      if (rs.full) {
        slot_full = true;
        if (rs.valid) {
          st.has_residual_echo_likelihood = true;
          st.residual_echo_likelihood = (double)rs.likelihood;
          st.has_residual_echo_likelihood_recent_max = true;
          st.residual_echo_likelihood_recent_max = (double)rs.recent_max;
        }
        WAP_CUDA(cudaMemcpy(&dr->slot_full, &zero, sizeof(int), cudaMemcpyHostToDevice));
      }
    }
    if (slot_full) h->cached_stats = st;
  }
  *out = h->cached_stats;
  return WapError::None;
}

// ---- stream lifecycle: a leg's complete state as an opaque blob (move a call between engines / GPUs)
namespace {
struct BlobHeader {
  uint32_t magic, version;
  uint64_t total_bytes;
  wap::EngineConfig cfg;  // the blob only fits an engine of the same config class
  wap::Ec3Params ep;      // ... and the same EchoCanceller3Config
  wap::Ec3Params ep_mc;   // ... and multichannel EchoCanceller3Config
  // host-side per-leg state
  WapConfig config;
  int32_t delay_ms, delay_set, capture_output_used, analog_level, playout_volume;
  float pre_gain_target, post_gain_target, agc2_gain_factor;
  uint8_t dirty[8];  // capture_output_used, pre, post, playout, agc2 gain, agc2 limiter reset
  WapStats cached_stats;
};
constexpr uint32_t kBlobMagic = 0x57415042u;  // "WAPB"
constexpr uint32_t kBlobVersion = 5;
size_t blob_bytes(const WapEngine* e) {
  size_t n = sizeof(BlobHeader) + sizeof(StreamState);
  if (e->d_upper) n += sizeof(wap::UpperBandState);
  if (e->d_rs) n += wap::kRsPerLeg * sizeof(wap::ResamplerState);
  if (e->d_extra) n += sizeof(wap::ExtraChannelState);
  if (e->d_cap_delay) n += (size_t)e->cap_delay_stride * sizeof(float);
  if (e->d_red) n += sizeof(wap::EchoDetectorState);
  if (e->d_mc) n += sizeof(wap::McState);
  if (e->d_mc_ns) n += wap::kMcCh * sizeof(wap::NsState);
  return n;
}
}  // namespace

size_t wap_stream_state_bytes(const WapAudioProcessing* h) {
  return (h && h->engine && h->slot >= 0) ? blob_bytes(h->engine) : 0;
}

WapError wap_stream_export_state(WapAudioProcessing* h, void* blob, size_t bytes) {
  if (!h || !blob) return WapError::NullPointer;
  if (!h->engine || h->slot < 0 || bytes < blob_bytes(h->engine)) return WapError::BadStreamParameter;
  WapEngine* e = h->engine;
  std::lock_guard<std::recursive_mutex> lk(e->mu);
  WAP_CUDA(cudaSetDevice(e->device));
  WAP_CUDA(cudaStreamSynchronize(e->stream));
  BlobHeader hd;
  memset(&hd, 0, sizeof(hd));
  hd.magic = kBlobMagic;
  hd.version = kBlobVersion;
  hd.total_bytes = blob_bytes(e);
  hd.cfg = e->cfg;
  hd.ep = e->ep;
  hd.ep_mc = e->ep_mc;
  hd.config = h->config;
  hd.delay_ms = e->leg_delay_ms[h->slot];
  hd.delay_set = e->leg_delay_set[h->slot];
  hd.capture_output_used = h->capture_output_used;
  hd.analog_level = h->analog_level;
  hd.playout_volume = h->playout_volume;
  hd.pre_gain_target = h->pre_gain_target;
  hd.post_gain_target = h->post_gain_target;
  hd.agc2_gain_factor = h->agc2_gain_factor;
  hd.dirty[0] = h->capture_output_used_dirty; hd.dirty[1] = h->pre_gain_dirty; hd.dirty[2] = h->post_gain_dirty;
  hd.dirty[3] = h->playout_volume_dirty; hd.dirty[4] = h->agc2_gain_dirty; hd.dirty[5] = h->agc2_reset_limiter;
  hd.cached_stats = h->cached_stats;
  unsigned char* p = static_cast<unsigned char*>(blob);
  memcpy(p, &hd, sizeof(hd));
  p += sizeof(hd);
  WAP_CUDA(cudaMemcpy(p, &e->d_states[h->slot], sizeof(StreamState), cudaMemcpyDeviceToHost));
  p += sizeof(StreamState);
  if (e->d_upper) {
    WAP_CUDA(cudaMemcpy(p, &e->d_upper[h->slot], sizeof(wap::UpperBandState), cudaMemcpyDeviceToHost));
    p += sizeof(wap::UpperBandState);
  }
  if (e->d_rs) {
    WAP_CUDA(cudaMemcpy(p, &e->d_rs[(size_t)h->slot * wap::kRsPerLeg], wap::kRsPerLeg * sizeof(wap::ResamplerState), cudaMemcpyDeviceToHost));
    p += wap::kRsPerLeg * sizeof(wap::ResamplerState);
  }
  if (e->d_extra) {
    WAP_CUDA(cudaMemcpy(p, &e->d_extra[h->slot], sizeof(wap::ExtraChannelState), cudaMemcpyDeviceToHost));
    p += sizeof(wap::ExtraChannelState);
  }
  if (e->d_cap_delay) {
    WAP_CUDA(cudaMemcpy(p, e->d_cap_delay + (size_t)h->slot * e->cap_delay_stride, (size_t)e->cap_delay_stride * sizeof(float), cudaMemcpyDeviceToHost));
    p += (size_t)e->cap_delay_stride * sizeof(float);
  }
  if (e->d_red) {
    WAP_CUDA(cudaMemcpy(p, &e->d_red[h->slot], sizeof(wap::EchoDetectorState), cudaMemcpyDeviceToHost));
    p += sizeof(wap::EchoDetectorState);
  }
  if (e->d_mc) {
    WAP_CUDA(cudaMemcpy(p, &e->d_mc[h->slot], sizeof(wap::McState), cudaMemcpyDeviceToHost));
    p += sizeof(wap::McState);
  }
  if (e->d_mc_ns)
    WAP_CUDA(cudaMemcpy(p, &e->d_mc_ns[(size_t)h->slot * wap::kMcCh], wap::kMcCh * sizeof(wap::NsState), cudaMemcpyDeviceToHost));
  return WapError::None;
}

namespace {
// Every device slab that belongs to slot `slot` of engine `e` (pointer, bytes), in blob order.
void for_each_slab(WapEngine* e, int slot, const std::function<void(void*, size_t)>& fn) {
  fn(&e->d_states[slot], sizeof(StreamState));
  if (e->d_upper) fn(&e->d_upper[slot], sizeof(wap::UpperBandState));
  if (e->d_rs) fn(&e->d_rs[(size_t)slot * wap::kRsPerLeg], wap::kRsPerLeg * sizeof(wap::ResamplerState));
  if (e->d_extra) fn(&e->d_extra[slot], sizeof(wap::ExtraChannelState));
  if (e->d_cap_delay) fn(e->d_cap_delay + (size_t)slot * e->cap_delay_stride, (size_t)e->cap_delay_stride * sizeof(float));
  if (e->d_red) fn(&e->d_red[slot], sizeof(wap::EchoDetectorState));
  if (e->d_mc) fn(&e->d_mc[slot], sizeof(wap::McState));
  if (e->d_mc_ns) fn(&e->d_mc_ns[(size_t)slot * wap::kMcCh], wap::kMcCh * sizeof(wap::NsState));
}
}  // namespace

// Live migration of a leg between two engines of the same config class (the same or another GPU): the
// state slabs go device to device (cudaMemcpyPeer: over NVLink between GPUs of one box), the handle stays
// valid and belongs to `dst` afterwards; the call continues bit-identically.
WapError wap_stream_migrate(WapAudioProcessing* h, WapEngine* dst) {
  if (!h || !dst) return WapError::NullPointer;
  WapEngine* src = h->engine;
  if (!src || h->slot < 0 || h->owns_engine) return WapError::BadStreamParameter;
  if (src == dst) return WapError::None;
  // both engines, always in the same order
  std::unique_lock<std::recursive_mutex> l1(src < dst ? src->mu : dst->mu);
  std::unique_lock<std::recursive_mutex> l2(src < dst ? dst->mu : src->mu);
  if (!wap::same_engine_config(src->cfg, dst->cfg) || !wap::same_ec3_params(src->ep, dst->ep) ||
      !wap::same_ec3_params(src->ep_mc, dst->ep_mc) || blob_bytes(src) != blob_bytes(dst))
    return WapError::UnsupportedConfig;
  if (dst->free_slots.empty()) return WapError::BadStreamParameter;  // destination engine is full
  WAP_CUDA(cudaSetDevice(src->device));
  WAP_CUDA(cudaStreamSynchronize(src->stream));
  WAP_CUDA(cudaSetDevice(dst->device));
  WAP_CUDA(cudaStreamSynchronize(dst->stream));
  const int to = dst->free_slots.back();
  std::vector<std::pair<void*, size_t>> from_slabs, to_slabs;
  for_each_slab(src, h->slot, [&](void* p, size_t n) { from_slabs.emplace_back(p, n); });
  for_each_slab(dst, to, [&](void* p, size_t n) { to_slabs.emplace_back(p, n); });
  if (from_slabs.size() != to_slabs.size()) return WapError::UnsupportedConfig;
  for (size_t i = 0; i < from_slabs.size(); ++i) {
    if (from_slabs[i].second != to_slabs[i].second) return WapError::UnsupportedConfig;
    WAP_CUDA(cudaMemcpyPeer(to_slabs[i].first, dst->device, from_slabs[i].first, src->device, from_slabs[i].second));
  }
  dst->free_slots.pop_back();
  const int dirty = (int)h->capture_output_used_dirty + (int)h->pre_gain_dirty + (int)h->post_gain_dirty +
                    (int)h->playout_volume_dirty + (int)h->agc2_gain_dirty;
  dst->leg_delay_ms[to] = src->leg_delay_ms[h->slot];
  dst->leg_delay_set[to] = src->leg_delay_set[h->slot];
  src->dirty_legs -= dirty;
  dst->dirty_legs += dirty;
  src->free_slots.push_back(h->slot);
  src->last_slots.clear(); src->last_handles.clear();
  dst->last_slots.clear(); dst->last_handles.clear();
  h->engine = dst;
  h->slot = to;
  return WapError::None;
}

WapError wap_stream_import_state(WapAudioProcessing* h, const void* blob, size_t bytes) {
  if (!h || !blob) return WapError::NullPointer;
  if (!h->engine || h->slot < 0 || bytes < sizeof(BlobHeader)) return WapError::BadStreamParameter;
  WapEngine* e = h->engine;
  std::lock_guard<std::recursive_mutex> lk(e->mu);
  BlobHeader hd;
  memcpy(&hd, blob, sizeof(hd));
  if (hd.magic != kBlobMagic || hd.version != kBlobVersion || hd.total_bytes != blob_bytes(e) || bytes < hd.total_bytes ||
      !wap::same_engine_config(hd.cfg, e->cfg) || !wap::same_ec3_params(hd.ep, e->ep) ||
      !wap::same_ec3_params(hd.ep_mc, e->ep_mc))
    return WapError::UnsupportedConfig;  // another config class (or library version)
  WAP_CUDA(cudaSetDevice(e->device));
  WAP_CUDA(cudaStreamSynchronize(e->stream));
  const unsigned char* p = static_cast<const unsigned char*>(blob) + sizeof(hd);
  WAP_CUDA(cudaMemcpy(&e->d_states[h->slot], p, sizeof(StreamState), cudaMemcpyHostToDevice));
  p += sizeof(StreamState);
  if (e->d_upper) {
    WAP_CUDA(cudaMemcpy(&e->d_upper[h->slot], p, sizeof(wap::UpperBandState), cudaMemcpyHostToDevice));
    p += sizeof(wap::UpperBandState);
  }
  if (e->d_rs) {
    WAP_CUDA(cudaMemcpy(&e->d_rs[(size_t)h->slot * wap::kRsPerLeg], p, wap::kRsPerLeg * sizeof(wap::ResamplerState), cudaMemcpyHostToDevice));
    p += wap::kRsPerLeg * sizeof(wap::ResamplerState);
  }
  if (e->d_extra) {
    WAP_CUDA(cudaMemcpy(&e->d_extra[h->slot], p, sizeof(wap::ExtraChannelState), cudaMemcpyHostToDevice));
    p += sizeof(wap::ExtraChannelState);
  }
  if (e->d_cap_delay) {
    WAP_CUDA(cudaMemcpy(e->d_cap_delay + (size_t)h->slot * e->cap_delay_stride, p, (size_t)e->cap_delay_stride * sizeof(float), cudaMemcpyHostToDevice));
    p += (size_t)e->cap_delay_stride * sizeof(float);
  }
  if (e->d_red) {
    WAP_CUDA(cudaMemcpy(&e->d_red[h->slot], p, sizeof(wap::EchoDetectorState), cudaMemcpyHostToDevice));
    p += sizeof(wap::EchoDetectorState);
  }
  if (e->d_mc) {
    WAP_CUDA(cudaMemcpy(&e->d_mc[h->slot], p, sizeof(wap::McState), cudaMemcpyHostToDevice));
    p += sizeof(wap::McState);
  }
  if (e->d_mc_ns)
    WAP_CUDA(cudaMemcpy(&e->d_mc_ns[(size_t)h->slot * wap::kMcCh], p, wap::kMcCh * sizeof(wap::NsState), cudaMemcpyHostToDevice));
  e->dirty_legs -= (int)h->capture_output_used_dirty + (int)h->pre_gain_dirty + (int)h->post_gain_dirty +
                   (int)h->playout_volume_dirty + (int)h->agc2_gain_dirty;
  h->config = hd.config;
  e->leg_delay_ms[h->slot] = hd.delay_ms;
  e->leg_delay_set[h->slot] = (unsigned char)hd.delay_set;
  h->stream_delay_ms = hd.delay_ms;
  h->capture_output_used = hd.capture_output_used != 0;
  h->analog_level = hd.analog_level;
  h->playout_volume = hd.playout_volume;
  h->pre_gain_target = hd.pre_gain_target;
  h->post_gain_target = hd.post_gain_target;
  h->agc2_gain_factor = hd.agc2_gain_factor;
  h->capture_output_used_dirty = hd.dirty[0]; h->pre_gain_dirty = hd.dirty[1]; h->post_gain_dirty = hd.dirty[2];
  h->playout_volume_dirty = hd.dirty[3]; h->agc2_gain_dirty = hd.dirty[4]; h->agc2_reset_limiter = hd.dirty[5];
  e->dirty_legs += (int)h->capture_output_used_dirty + (int)h->pre_gain_dirty + (int)h->post_gain_dirty +
                   (int)h->playout_volume_dirty + (int)h->agc2_gain_dirty;
  h->cached_stats = hd.cached_stats;
  return WapError::None;
}

WapError wap_stream_read_taps(WapAudioProcessing* h, WapStageTaps* out) {
  if (!h || !out) return WapError::NullPointer;
  if (!h->engine || h->slot < 0) return WapError::BadStreamParameter;
  WapEngine* e = h->engine;
  if (e->cfg.mc) return WapError::UnsupportedConfig;  // the stage taps describe a single capture channel
  std::lock_guard<std::recursive_mutex> lk(e->mu);
  WAP_CUDA(cudaSetDevice(e->device));
  WAP_CUDA(cudaStreamSynchronize(e->stream));
  memset(out, 0, sizeof(*out));
  const StreamState* slab = &e->d_states[h->slot];
  auto fetch = [&](void* dst, const void* src, size_t n) {
    return cudaMemcpy(dst, src, n, cudaMemcpyDeviceToHost) == cudaSuccess;
  };
  bool ok = true;
  if (e->cfg.aec_enabled) {
    wap::Aec3Scalars s;
    ok = ok && fetch(&s, &slab->aec.s, sizeof(s)) && fetch(out->aec3_erle, slab->aec.erle, sizeof(out->aec3_erle)) &&
         fetch(out->aec3_erle_onset_compensated, slab->aec.erle_onset_comp, sizeof(out->aec3_erle)) &&
         fetch(out->aec3_erl, slab->aec.erl, sizeof(out->aec3_erl)) &&
         fetch(out->aec3_suppressor_gain, slab->aec.last_gain, sizeof(out->aec3_suppressor_gain)) &&  // G^2, see below
         fetch(out->aec3_N2, slab->aec.cng_N2, sizeof(out->aec3_N2)) &&
         fetch(out->aec3_refined_gain_H_error, slab->aec.H_error, sizeof(out->aec3_refined_gain_H_error));
    for (float& g : out->aec3_suppressor_gain) g = sqrtf(g);  // the reference dumps the amplitude gain
    out->aec3_erl_time_domain = s.erl_time_domain;
    out->aec3_fullband_erle_log2 = s.fb_erle_time_domain_log2;
    out->aec3_filter_delay = s.fa_filter_delay_blocks;  // one capture channel: the minimum is the value
    out->aec3_min_direct_path_filter_delay = s.fd_min_filter_delay;
    out->aec3_render_delay_controller_buffer_delay = s.ctl_has_delay ? s.ctl_delay : 0;
    out->aec3_usable_linear_estimate = s.fq_usable;
    out->aec3_transparent_mode = s.tm_active;
    out->aec3_initial_state = s.init_state;
    out->aec3_echo_saturation = s.saturated_echo;
    out->aec3_capture_saturation = s.capture_signal_saturation;
    out->aec3_dominant_nearend = s.dn_nearend_state;
  }
  if (e->cfg.ns_enabled) {
    ok = ok && fetch(out->ns_noise_spectrum, slab->ns.noise, sizeof(out->ns_noise_spectrum)) &&
         fetch(out->ns_filter, slab->ns.wiener, sizeof(out->ns_filter)) &&
         fetch(out->ns_speech_probability, slab->ns.speech_prob, sizeof(out->ns_speech_probability)) &&
         fetch(&out->ns_prior_speech_probability, &slab->ns.prior_speech_prob, sizeof(float));
  }
  return ok ? WapError::None : WapError::Internal;
}

// Test / tooling hook: raw copy of one leg's state slab (wap_state.h layout).
int wapdbg_read_state(const WapAudioProcessing* h, void* out, size_t bytes) {
  if (!h || !h->engine || h->slot < 0 || bytes > sizeof(StreamState)) return -1;
  WapEngine* e = h->engine;
  if (cudaSetDevice(e->device) != cudaSuccess || cudaStreamSynchronize(e->stream) != cudaSuccess) return -1;
  return cudaMemcpy(out, &e->d_states[h->slot], bytes, cudaMemcpyDeviceToHost) == cudaSuccess ? 0 : -1;
}
size_t wapdbg_state_size(void) { return sizeof(StreamState); }

// Per-kernel timing for the roofline report: while enabled every tick records CUDA
// events around k_front / k_delay / k_echo on the engine's stream and waits for them.
WapError wap_engine_enable_kernel_timing(WapEngine* e, bool on) {
  if (!e) return WapError::NullPointer;
  WAP_CUDA(cudaSetDevice(e->device));
  if (on && !e->ev[0])
    for (int k = 0; k < 4; ++k) WAP_CUDA(cudaEventCreate(&e->ev[k]));
  e->timing = on;
  for (int k = 0; k < 3; ++k) e->kernel_ms[k] = 0;
  e->timed_ticks = 0;
  return WapError::None;
}
// out_ms[3] = accumulated milliseconds of {k_front, k_delay, k_echo}; returns the number of timed ticks.
int64_t wap_engine_read_kernel_timing(const WapEngine* e, double* out_ms) {
  if (!e || !out_ms) return 0;
  for (int k = 0; k < 3; ++k) out_ms[k] = e->kernel_ms[k];
  return e->timed_ticks;
}
static const double kBlockBytesHi = 2 * 64;  // floats of bands 1-2 per block
// Algorithmic HBM bytes per leg-frame attributed to each tick kernel (SURVEY.md 8(d) groups):
// k_front: audio in + HPF state; k_delay: the "delay estimation" group; k_echo: the rest.
void wap_engine_algorithmic_bytes_per_kernel(const WapEngine* e, double* out_bytes) {
  if (!e || !out_bytes) return;
  const double total = wap_engine_algorithmic_bytes_per_frame(e);
  const double front = e->frame_len * 4.0 * (e->cfg.aec_enabled ? 2 : 1) + (e->cfg.hpf_enabled ? 96.0 : 0.0);
  const double delay = e->cfg.aec_enabled ? 2.5 * 40256.0 : 0.0;
  out_bytes[0] = front;
  out_bytes[1] = delay;
  out_bytes[2] = total - front - delay;
}

double wap_engine_algorithmic_bytes_per_frame(const WapEngine* e) {
  if (!e) return 0.0;
  // SURVEY.md section 8(d) byte model.
  const int B = e->cfg.num_bands;
  double bytes = 0.0;
  if (e->cfg.mc) {
    // Multi-channel legs after stereo detection (R = 2 render, C = 2 capture channels, multichannel config:
    // refined 13 / coarse 11 partitions), per 64-sample block:
    //   delay estimation (on the mixed channels)                          40256   (as mono)
    //   adaptive filters  C x (13 + 11) x R x 65 bins x re,im x 4 B x r/w  99840
    //   render FFT partitions  13 x R x 65 x re,im x 4 B x {filter, adapt}  27040   (shared by the capture channels)
    //   render spectra  13 x R x 65 x 4 B                                    6760
    //   H2  C x 13 x 65 x 4 B x r/w                                          13520
    //   estimators, comfort noise, suppressor state  C x 16504               33008   (mono: 16504)
    bytes += 2.5 * (40256.0 + 99840.0 + 27040.0 + 6760.0 + 13520.0 + 33008.0);
    bytes += 2 * 96.0;                                   // high-pass filter state, both channels
    if (e->cfg.ns_enabled) bytes += 2 * (2223.0 * 8.0 + 24.0);   // noise suppressor state, both channels
    if (B == 3) {
      bytes += (2 * 2 + 2) * 150 * 4.0;                  // capture analysis + synthesis and render analysis state, x2 channels
      bytes += 2 * (2.5 * (4.0 * kBlockBytesHi * 5) + 128.0);  // upper-band ring / delay / framers, PostFilter state
    }
    bytes += e->frame_len * 4.0 * 3;                     // render + capture in, output
    return bytes;
  }
  if (e->cfg.aec_enabled) bytes += 2.5 * 107460.0;
  if (e->cfg.ns_enabled) bytes += 2223.0 * 8.0 + 24.0;
  if (e->cfg.hpf_enabled) bytes += 96.0;
  if (B == 3) bytes += 2 * 2 * 150 * 4.0;
  if (B == 2) bytes += 2 * 2 * 8 * 4.0;
  if (B >= 2 && e->cfg.aec_enabled)  // render split state, upper-band ring / delay / framers, PostFilter state
    bytes += 2 * 150 * 4.0 + 2.5 * (4.0 * kBlockBytesHi * 5) + 128.0;
  bytes += e->frame_len * 4.0 * (e->cfg.aec_enabled ? 3 : 2);
  return bytes;
}

}  // extern "C"
#include "wap_single.inc"
