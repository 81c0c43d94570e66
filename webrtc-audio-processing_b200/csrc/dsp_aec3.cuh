// AEC3 (placeholder being filled in): see dsp_aec3_*.cuh
#pragma once
#include "wap_dev.cuh"
#include "wap_state.h"
namespace wap {
struct AecScratch { float tmp[16]; };
WAP_DEV void aec3_buffer_render_frame(Aec3State&, const EngineConfig&, const float*, AecScratch&) {}
WAP_DEV void aec3_analyze_capture(Aec3State&, const float*, int) {}
WAP_DEV void aec3_process_capture_frame(Aec3State&, const EngineConfig&, float*, int, AecScratch&) {}
}  // namespace wap
