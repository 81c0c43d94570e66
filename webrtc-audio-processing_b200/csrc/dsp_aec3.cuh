// AEC3 frame driver for one call leg: the parts of EchoCanceller3 and
// BlockProcessorImpl that turn 10 ms frames into 64-sample blocks and sequence
// the per-block stages.
//   EchoCanceller3::{AnalyzeCapture, ProcessCapture, EmptyRenderQueue}  aec3/echo_canceller3.cc:862-1002
//   FrameBlocker / BlockFramer                                          aec3/frame_blocker.cc:40-84,
//                                                                       aec3/block_framer.cc:41-85
//   BlockProcessorImpl::{ProcessCapture, BufferRender}                  aec3/block_processor.cc:104-216
// The stages themselves live in dsp_aec3_{render,delay,subtractor,remover}.cuh.
//
// The reference queues render frames in AnalyzeRender and buffers them at the
// start of the next ProcessCapture; with the batched tick order "render frame,
// then capture frame" (and for the single-leg entry points, whose queued render
// frames are drained in front of the capture frame) that is the same sequence of
// BufferRender / ProcessCapture block calls.
#pragma once

#include "dsp_aec3_common.cuh"
#include "dsp_aec3_delay.cuh"
#include "dsp_aec3_remover.cuh"
#include "dsp_aec3_render.cuh"
#include "dsp_aec3_subtractor.cuh"

namespace wap {

constexpr int kSubFrame = 80;  // kSubFrameLength (aec3_common.h:45)
constexpr int kScalarWords = (int)(sizeof(Aec3Scalars) / sizeof(int));

WAP_DEV void aec3_stage_scalars(const Aec3State& a, AecScratch& sc) {
  const int* src = reinterpret_cast<const int*>(&a.s);
  int* dst = reinterpret_cast<int*>(&sc.s);
  for (int i = lane_id(); i < kScalarWords; i += 32) dst[i] = src[i];
  __syncwarp();
}
WAP_DEV void aec3_unstage_scalars(Aec3State& a, const AecScratch& sc) {
  __syncwarp();
  const int* src = reinterpret_cast<const int*>(&sc.s);
  int* dst = reinterpret_cast<int*>(&a.s);
  for (int i = lane_id(); i < kScalarWords; i += 32) dst[i] = src[i];
  __syncwarp();
}

// FrameBlocker::InsertSubFrameAndExtractBlock: block = buffered samples + the
// head of the sub-frame; the tail of the sub-frame becomes the new buffer.
WAP_DEV void blocker_insert_and_extract(float* buffer, int* len, const float* sub_frame, float* block) {
  const int lane = lane_id();
  const int n = *len;
  const int samples_to_block = kBlock - n;
  __syncwarp();
  for (int i = lane; i < kBlock; i += 32) block[i] = i < n ? buffer[i] : sub_frame[i - n];
  __syncwarp();
  for (int i = lane; i < kSubFrame - samples_to_block; i += 32) buffer[i] = sub_frame[samples_to_block + i];
  __syncwarp();
  if (lane == 0) *len = kSubFrame - samples_to_block;
  __syncwarp();
}
// FrameBlocker::ExtractBlock
WAP_DEV void blocker_extract(const float* buffer, int* len, float* block) {
  for (int i = lane_id(); i < kBlock; i += 32) block[i] = buffer[i];
  __syncwarp();
  if (lane_id() == 0) *len = 0;
  __syncwarp();
}
// BlockFramer::InsertBlockAndExtractSubFrame
WAP_DEV void framer_insert_and_extract(float* buffer, int* len, const float* block, float* sub_frame) {
  const int lane = lane_id();
  const int n = *len;
  const int samples_to_frame = kSubFrame - n;
  __syncwarp();
  for (int i = lane; i < kSubFrame; i += 32) sub_frame[i] = i < n ? buffer[i] : block[i - n];
  __syncwarp();
  for (int i = lane; i < kBlock - samples_to_frame; i += 32) buffer[i] = block[samples_to_frame + i];
  __syncwarp();
  if (lane == 0) *len = kBlock - samples_to_frame;
  __syncwarp();
}
// BlockFramer::InsertBlock
WAP_DEV void framer_insert(float* buffer, int* len, const float* block) {
  for (int i = lane_id(); i < kBlock; i += 32) buffer[i] = block[i];
  __syncwarp();
  if (lane_id() == 0) *len = kBlock;
  __syncwarp();
}

// EmptyRenderQueue for one queued render frame (band 0, 160 samples).
WAP_DEV void aec3_buffer_render_frame(Aec3State& a, const EngineConfig& cfg, const float* band0, AecScratch& sc) {
  aec3_stage_scalars(a, sc);
  for (int sub = 0; sub < 2; ++sub) {
    blocker_insert_and_extract(a.render_blocker, &sc.s.render_blocker_len, band0 + sub * kSubFrame, sc.x);
    aec3_buffer_render_block(a, sc);
  }
  if (sc.s.render_blocker_len == kBlock) {
    blocker_extract(a.render_blocker, &sc.s.render_blocker_len, sc.x);
    aec3_buffer_render_block(a, sc);
  }
  aec3_unstage_scalars(a, sc);
}

// EchoCanceller3::AnalyzeCapture: microphone saturation over the full-band frame.
WAP_DEV void aec3_analyze_capture(Aec3State& a, const float* full, int n) {
  int sat = 0;
  for (int i = lane_id(); i < n; i += 32) sat |= (full[i] >= 32700.0f || full[i] <= -32700.0f) ? 1 : 0;
  sat = warp_or(sat);
  if (lane_id() == 0) a.s.saturated_microphone_signal = sat ? 1 : 0;
  __syncwarp();
}

// BlockProcessorImpl::ProcessCapture for the capture block in sc.y (in place).
WAP_DEV void aec3_process_capture_block(Aec3State& a, const EngineConfig& cfg, AecScratch& sc, bool echo_path_gain_change) {
  const int lane = lane_id();
  Aec3Scalars& s = sc.s;
  __syncwarp();
  if (!s.render_properly_started) return;  // no render data yet: capture passes through
  const bool first_capture = !s.capture_properly_started;
  const bool render_overrun = s.render_event == kEventRenderOverrun;
  __syncwarp();
  if (first_capture) {
    if (lane == 0) {
      s.capture_properly_started = 1;
      rdb_reset(s);
    }
    __syncwarp();
    delay_controller_reset(a, sc, true);
  }
  EchoPathVariability v;
  v.gain_change = echo_path_gain_change ? 1 : 0;
  v.delay_change = kDelayAdjNone;
  v.clock_drift = 0;
  if (render_overrun) {
    v.delay_change = kDelayAdjBufferFlush;
    delay_controller_reset(a, sc, true);
  }
  __syncwarp();
  if (lane == 0) {
    s.render_event = kEventNone;
    sc.ired[0] = rdb_prepare_capture_processing(s);
  }
  __syncwarp();
  if (sc.ired[0] == kEventRenderUnderrun) delay_controller_reset(a, sc, false);

  aec3_get_delay(a, sc);
  if (lane == 0) {
    s.bp_has_estimated_delay = s.ctl_has_delay;
    s.bp_est_delay = s.ctl_delay;
    s.bp_est_quality = s.ctl_delay_quality;
    sc.ired[0] = 0;
    if (s.ctl_has_delay) sc.ired[0] = rdb_align_from_delay(s, s.ctl_delay) ? 1 : 0;
  }
  __syncwarp();
  if (sc.ired[0]) v.delay_change = kDelayAdjNewDetectedDelay;
  v.clock_drift = s.cd_level != 0;
  __syncwarp();
  echo_remover_process_capture(a, cfg, sc, v, s.saturated_microphone_signal != 0, s.bp_has_estimated_delay,
                               s.bp_est_delay);
}

// EchoCanceller3::ProcessCapture for one capture frame (band 0 in place).
// `delay_ms` >= 0 : AudioProcessingImpl forwarded set_stream_delay_ms() through
// SetAudioBufferDelay (audio_processing_impl.cc:1409-1411).
WAP_DEV void aec3_process_capture_frame(Aec3State& a, const EngineConfig& cfg, float* band0, int delay_ms,
                                        AecScratch& sc) {
  aec3_stage_scalars(a, sc);
  if (delay_ms >= 0) {
    if (lane_id() == 0) rdb_set_audio_buffer_delay(sc.s, delay_ms);
    __syncwarp();
  }
  for (int sub = 0; sub < 2; ++sub) {
    float* sub_frame = band0 + sub * kSubFrame;
    blocker_insert_and_extract(a.capture_blocker, &sc.s.capture_blocker_len, sub_frame, sc.y);
    aec3_process_capture_block(a, cfg, sc, false);
    framer_insert_and_extract(a.output_framer, &sc.s.output_framer_len, sc.y, sub_frame);
  }
  if (sc.s.capture_blocker_len == kBlock) {
    blocker_extract(a.capture_blocker, &sc.s.capture_blocker_len, sc.y);
    aec3_process_capture_block(a, cfg, sc, false);
    framer_insert(a.output_framer, &sc.s.output_framer_len, sc.y);
  }
  // ApmStatsReporter::UpdateStatistics (audio_processing_impl.cc:2322-2328): a
  // one-slot queue -- while the slot is full (nobody called GetStatistics) the
  // newer statistics are discarded.
  if (lane_id() == 0 && !sc.s.stats_slot_full) {
    Aec3Scalars& s = sc.s;
    s.stats_slot_full = 1;
    s.stats_erl_time_domain = s.erl_time_domain;
    s.stats_erle_log2 = s.fb_erle_time_domain_log2;
    // BlockProcessorImpl::GetMetrics reports RenderDelayBuffer::Delay() == ComputeDelay()
    // (render_delay_buffer.cc:57), not the aligned delay_.
    s.stats_has_delay = 1;
    s.stats_delay_blocks = rdb_compute_delay(s);
  }
  aec3_unstage_scalars(a, sc);
}

}  // namespace wap
