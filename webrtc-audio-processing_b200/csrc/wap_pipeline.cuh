// Per-stream frame pipeline: the body of AudioProcessingImpl::
// ProcessRenderStreamLocked / ProcessCaptureStreamLocked for the enabled
// submodules (reference audio_processing_impl.cc:1264-1561,1653-1687), run by
// one warp per stream.
#pragma once

#include "dsp_aec3.cuh"
#include "dsp_filters.cuh"
#include "dsp_ns.cuh"
#include "wap_dev.cuh"
#include "wap_state.h"

namespace wap {

struct TickArgs {
  StreamState* states;
  const int* slots;       // [n] arena slot of each stream, or nullptr => slot i
  const int* delays_ms;   // [n] per-stream set_stream_delay_ms value (-1 unset) or nullptr
  int uniform_delay_ms;   // used when delays_ms == nullptr (-1 unset)
  int n;
  const void* render;     // [n][frame] or nullptr
  const void* capture;    // [n][frame] or nullptr
  void* out;              // [n][frame]
  int fmt;                // 0 = int16, 1 = float [-1,1]
  EngineConfig cfg;
};

// Shared-memory footprint of one warp, in floats.
WAP_DEV constexpr int scratch_floats_frame(int bands) { return 2 * kFrame * bands; }
constexpr int kScratchFloatsDsp =
    (int)((sizeof(NsScratch) > sizeof(AecScratch) ? sizeof(NsScratch) : sizeof(AecScratch)) / sizeof(float)) + 4;
inline int warp_scratch_floats(int bands) { return 2 * kFrame * bands + kScratchFloatsDsp; }

WAP_DEV void load_frame(const void* src, size_t stream, int len, int fmt, float* dst) {
  const int lane = lane_id();
  if (fmt == 0) {
    const int16_t* p = reinterpret_cast<const int16_t*>(src) + stream * len;
    for (int i = lane; i < len; i += 32) dst[i] = (float)p[i];  // S16ToFloatS16
  } else {
    const float* p = reinterpret_cast<const float*>(src) + stream * len;
    for (int i = lane; i < len; i += 32) {  // FloatToFloatS16 (audio_util.h:65-69)
      float v = p[i];
      v = fminr(v, 1.f);
      v = fmaxr(v, -1.f);
      dst[i] = v * 32768.f;
    }
  }
  __syncwarp();
}

WAP_DEV void store_frame(void* dst, size_t stream, int len, int fmt, const float* src) {
  const int lane = lane_id();
  if (fmt == 0) {
    int16_t* p = reinterpret_cast<int16_t*>(dst) + stream * len;
    for (int i = lane; i < len; i += 32) {  // FloatS16ToS16 (audio_util.h:52-56)
      float v = src[i];
      v = fminr(v, 32767.f);
      v = fmaxr(v, -32768.f);
      p[i] = (int16_t)(v + copysignf(0.5f, v));
    }
  } else {
    float* p = reinterpret_cast<float*>(dst) + stream * len;
    for (int i = lane; i < len; i += 32) {  // FloatS16ToFloat (audio_util.h:71-76)
      float v = src[i];
      v = fminr(v, 32768.f);
      v = fmaxr(v, -32768.f);
      p[i] = v * (1.f / 32768.f);
    }
  }
}

// One 10 ms tick of one stream.
WAP_DEV void process_stream_tick(const TickArgs& a, int idx, float* scratch) {
  const EngineConfig& cfg = a.cfg;
  const int B = cfg.num_bands;
  const int flen = kFrame * B;
  const int slot = a.slots ? a.slots[idx] : idx;
  StreamState& st = a.states[slot];
  float* full = scratch;
  float* bands = (B == 1) ? full : scratch + flen;
  void* dsp = scratch + 2 * flen;
  NsScratch& ns_sc = *reinterpret_cast<NsScratch*>(dsp);
  AecScratch& aec_sc = *reinterpret_cast<AecScratch*>(dsp);

  // ---------------- render side (ProcessReverseStream)
  const bool render_live = !(cfg.reinit_on_first_capture && !st.seen_capture);
  __syncwarp();
  if (a.render && cfg.aec_enabled && render_live) {
    load_frame(a.render, idx, flen, a.fmt, full);
    if (B == 3) three_band_analysis(full, bands, reinterpret_cast<float*>(dsp), st.render_bands.analysis);
    aec3_buffer_render_frame(st.aec, cfg, bands, aec_sc);
  }
  if (!a.capture) return;

  // ---------------- capture side (ProcessStream)
  if (lane_id() == 0) st.seen_capture = 1;
  load_frame(a.capture, idx, flen, a.fmt, full);
  if (cfg.hpf_enabled) {
    biquad_cascade<3>(full, flen, B == 3 ? kHpf48k : kHpf16k, st.hpf);
  }
  if (cfg.aec_enabled) {
    aec3_analyze_capture(st.aec, full, flen);
  }
  if (B == 3) three_band_analysis(full, bands, reinterpret_cast<float*>(dsp), st.capture_bands.analysis);
  if (cfg.ns_enabled) ns_analyze(st.ns, cfg, bands, ns_sc);
  if (cfg.aec_enabled) {
    const int d = a.delays_ms ? a.delays_ms[idx] : a.uniform_delay_ms;
    aec3_process_capture_frame(st.aec, cfg, bands, d, aec_sc);
  }
  if (cfg.ns_enabled) ns_process(st.ns, cfg, bands, ns_sc);
  if (B == 3) three_band_synthesis(bands, full, reinterpret_cast<float*>(dsp), st.capture_bands.synthesis);
  __syncwarp();
  store_frame(a.out, idx, flen, a.fmt, full);
}

}  // namespace wap
