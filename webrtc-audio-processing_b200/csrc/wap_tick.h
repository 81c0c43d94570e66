// Arguments of one 10 ms tick, shared by the three tick kernels.
#pragma once
#include "dsp_resampler.cuh"
#include "wap_ec3_params.h"
#include "wap_mc_state.h"
#include "wap_state.h"

namespace wap {

// Second capture channel of a stereo leg: it only feeds AEC3's saturation test.
struct ExtraChannelState {
  Biquad hpf[3];
  int pad_[4];
  ResamplerState rs;
};

struct TickArgs {
  StreamState* states;
  UpperBandState* upper;  // [arena slot] for 48 kHz AEC3 engines, else nullptr
  const int* slots;       // [n] arena slot of each stream, or nullptr => slot i
  const int* delays_ms;   // [n] per-stream set_stream_delay_ms value (-1 unset) or nullptr
  int uniform_delay_ms;   // used when delays_ms == nullptr (-1 unset)
  int n;
  EchoDetectorState* red; // [arena slot] residual echo detector, or nullptr (not enabled)
  const void* render;     // [n][frame] or nullptr
  const void* capture;    // [n][frame] or nullptr
  void* out;              // [n][frame]
  int fmt;                // 0 = int16, 1 = float [-1,1]  (k_front of a resampled engine: 2 = FloatS16 floats)
  EngineConfig cfg;
  Ec3Params ep;           // the engine's EchoCanceller3Config parameters (the default config for default engines)
  // Resampled engines only (cfg.pre_stage / resample_out / fullband_out): per-leg resampler states [slot][kRsPerLeg], the
  // processing-rate frames k_resample leaves for k_front (FloatS16 floats), kernels and ratios.
  ResamplerState* rs;
  float* rs_render;       // [n][proc frame]
  float* rs_capture;      // [n][proc frame]
  const float* rs_kernel_in;      // capture input -> processing rate
  const float* rs_kernel_out;     // processing rate -> output
  const float* rs_kernel_render;  // render input -> processing rate
  double rs_ratio_in, rs_ratio_out, rs_ratio_render;
  // Engines with delay.fixed_capture_delay_samples > 0 only: BlockDelayBuffer rings, per leg
  // [band][delay] floats followed by the insert position (one int), `cap_delay_stride` floats per leg.
  float* cap_delay;
  int cap_delay_stride;
  // Stereo engines only: second capture channel (high-pass state, input resampler, resampled frame).
  ExtraChannelState* extra;  // [slot]
  float* rs_capture1;        // [n][proc frame] or nullptr
  // Multi-channel engines only (stereo frames with pipeline.multi_channel_render / _capture, BASELINE config 4):
  // the per-leg multi-channel arena, the freshly constructed state EchoCanceller3::Initialize copies in, the
  // multichannel EchoCanceller3Config (`ep` is the mono one) and the parameters outside Ec3Params ([0]: of the
  // mono config, [1]: of the multichannel config; the detector reads [0]).
  McState* mc;               // [slot] or nullptr
  NsState* mc_ns;            // [slot][kMcCh]: per-channel NoiseSuppressor::ChannelState when NS is enabled, else nullptr
  const McTemplates* mc_templates;
  Ec3Params ep_mc;
  McParams mcp[2];
};

}  // namespace wap
