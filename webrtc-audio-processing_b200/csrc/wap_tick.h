// Arguments of one 10 ms tick, shared by the three tick kernels.
#pragma once
#include "wap_state.h"

namespace wap {

struct TickArgs {
  StreamState* states;
  UpperBandState* upper;  // [arena slot] for 48 kHz AEC3 engines, else nullptr
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

}  // namespace wap
