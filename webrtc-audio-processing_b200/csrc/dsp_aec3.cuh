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
//
// A tick runs as three kernels: k_front (one thread per leg: frame -> block
// slicing, the serial IIRs, RenderDelayBuffer::Insert bookkeeping), k_delay (one
// warp per leg: delay estimation for all capture blocks of the tick -- nothing in
// it depends on the echo remover) and k_echo (one warp per leg: render FFTs, NS,
// echo remover, output).  TickScratch carries the per-block hand-over.
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
  #pragma unroll
  for (int i = lane_id(); i < kScalarWords; i += 32) dst[i] = src[i];
  __syncwarp();
}
WAP_DEV void aec3_unstage_scalars(Aec3State& a, const AecScratch& sc) {
  __syncwarp();
  const int* src = reinterpret_cast<const int*>(&sc.s);
  int* dst = reinterpret_cast<int*>(&a.s);
  #pragma unroll
  for (int i = lane_id(); i < kScalarWords; i += 32) dst[i] = src[i];
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
// Same for an upper band, with the fill level passed by value.
WAP_DEV void framer_insert_and_extract_local(float* buffer, int n, const float* block, float* sub_frame) {
  const int lane = lane_id();
  const int samples_to_frame = kSubFrame - n;
  __syncwarp();
  for (int i = lane; i < kSubFrame; i += 32) sub_frame[i] = i < n ? buffer[i] : block[i - n];
  __syncwarp();
  for (int i = lane; i < kBlock - samples_to_frame; i += 32) buffer[i] = block[samples_to_frame + i];
  __syncwarp();
}
// BlockFramer::InsertBlock
WAP_DEV void framer_insert(float* buffer, int* len, const float* block) {
  #pragma unroll
  for (int i = lane_id(); i < kBlock; i += 32) buffer[i] = block[i];
  __syncwarp();
  if (lane_id() == 0) *len = kBlock;
  __syncwarp();
}

// ---------------------------------------------------------------- k_delay
// BlockProcessorImpl::ProcessCapture up to (not including) the echo remover, for
// capture block b of this tick: start-up gating, over/under-run handling, delay
// estimation and render-buffer alignment.  Writes the hand-over record for k_echo.
WAP_DEV void aec3_delay_block(Aec3State& a, AecScratch& sc, TickScratch& ts, int b) {
  const int lane = lane_id();
  Aec3Scalars& s = sc.s;
  CaptureBlockRec& rec = ts.crec[b];
  __syncwarp();
  const bool external = WAP_EC3(use_external_delay_estimator) != 0;   // no delay controller at all
  if (!s.render_properly_started) {  // no render data yet: capture passes through
    if (lane == 0) {
      rec.process = 0;
      if (external) ++s.rdb_capture_calls;   // HandleSkippedCaptureProcessing
    }
    return;
  }
  const bool first_capture = !s.capture_properly_started;
  const bool render_overrun = s.render_event == kEventRenderOverrun;
  if (lane < kSubBlock) sc.ds[lane] = ts.cap_ds[b][lane];
  __syncwarp();
  if (first_capture) {
    if (lane == 0) {
      s.capture_properly_started = 1;
      rdb_reset(s, WAP_EC3(default_delay));
    }
    __syncwarp();
    if (!external) delay_controller_reset(a, sc, true);
  }
  int delay_change = kDelayAdjNone;
  if (render_overrun) {
    delay_change = kDelayAdjBufferFlush;
    if (!external) delay_controller_reset(a, sc, true);
  }
  __syncwarp();
  if (lane == 0) {
    s.render_event = kEventNone;
    if (external) ++s.rdb_capture_calls;
    sc.ired[0] = rdb_prepare_capture_processing(s, WAP_EC3(default_delay), WAP_EC3(excess_render_detection_interval_blocks),
                                                WAP_EC3(max_allowed_excess_render_blocks));
  }
  __syncwarp();
  if (external) {
    // RenderDelayBufferImpl::AlignFromExternalDelay (render_delay_buffer.cc:375-384); the echo remover only
    // runs once a buffer delay has been received (block_processor.cc:190-194); estimated_delay_ stays empty
    if (lane == 0) {
      if (s.has_external_delay) {
        const int delay = s.rdb_render_calls - s.rdb_capture_calls + s.external_delay;
        rdb_apply_total_delay(s, delay - WAP_EC3(delay_headroom_samples) / kBlock);
      }
      s.bp_has_estimated_delay = 0;
      rec.process = s.has_external_delay ? 1 : 0;
      rec.blocks_read = s.blocks_read;
      rec.spectra_read = s.spectra_read;
      rec.gain_change = ts.pad_[1];
      rec.delay_change = delay_change;
      rec.clock_drift = 0;
      rec.est_has = 0;
      rec.est_delay = 0;
    }
    __syncwarp();
    return;
  }
  if (sc.ired[0] == kEventRenderUnderrun) delay_controller_reset(a, sc, false);

  aec3_get_delay(a, sc);
  if (lane == 0) {
    s.bp_has_estimated_delay = s.ctl_has_delay;
    s.bp_est_delay = s.ctl_delay;
    s.bp_est_quality = s.ctl_delay_quality;
    if (s.ctl_has_delay && rdb_align_from_delay(s, s.ctl_delay)) delay_change = kDelayAdjNewDetectedDelay;
    rec.process = 1;
    rec.blocks_read = s.blocks_read;
    rec.spectra_read = s.spectra_read;
    rec.gain_change = ts.pad_[1];  // capture_.echo_path_gain_change, set by k_front
    rec.delay_change = delay_change;
    rec.clock_drift = s.cd_level != 0;
    rec.est_has = s.bp_has_estimated_delay;
    rec.est_delay = s.bp_est_delay;
  }
  __syncwarp();
}

// All capture blocks of this tick.
WAP_DEV void aec3_delay_frame(Aec3State& a, TickScratch& ts, AecScratch& sc) {
  const int nb = ts.n_capture_blocks;
  if (nb == 0) return;
  aec3_stage_scalars(a, sc);
  for (int b = 0; b < nb; ++b) aec3_delay_block(a, sc, ts, b);
  aec3_unstage_scalars(a, sc);
}

// ---------------------------------------------------------------- k_echo
// Render blocks of this tick: vector half of the insert.
WAP_DEV void aec3_echo_render(Aec3State& a, const TickScratch& ts, AecScratch& sc, UpperBandState* up) {
  const int lane = lane_id();
  for (int r = 0; r < ts.n_render_blocks; ++r) {
    __syncwarp();
    #pragma unroll
    for (int i = lane; i < kBlock; i += 32) sc.x[i] = ts.render_blocks[r][i];
    aec3_render_insert_vector(a, sc, ts.rins[r]);
    if (up) {  // bands 1-2 of the block ring
      const int bw = ts.rins[r].blocks_write;
      #pragma unroll
      for (int i = lane; i < 2 * kBlock; i += 32) (&up->blocks_hi[bw][0][0])[i] = (&up->render_blocks_hi[r][0][0])[i];
    }
  }
}

// EchoRemover for capture block b (sc.y in/out) against the render-buffer view k_delay recorded.
// 3-band legs: bands 1-2 of the block come back in sc.x / sc.rm.x_aligned.
WAP_DEV void aec3_echo_block(Aec3State& a, const EngineConfig& cfg, AecScratch& sc, const TickScratch& ts, int b,
                             UpperBandState* up) {
  const int lane = lane_id();
  __syncwarp();
  #pragma unroll
  for (int i = lane; i < kBlock; i += 32) sc.y[i] = ts.capture_blocks[b][i];
  const CaptureBlockRec& rec = ts.crec[b];
  if (!rec.process) {
    if (up)
      #pragma unroll
      for (int i = lane; i < kBlock; i += 32) {
        sc.x[i] = up->capture_blocks_hi[b][0][i];
        sc.rm.x_aligned[i] = up->capture_blocks_hi[b][1][i];
      }
    __syncwarp();
    return;
  }
  if (lane == 0) {
    sc.s.blocks_read = rec.blocks_read;
    sc.s.spectra_read = rec.spectra_read;
  }
  __syncwarp();
  EchoPathVariability v;
  v.gain_change = rec.gain_change;
  v.delay_change = rec.delay_change;
  v.clock_drift = rec.clock_drift;
  echo_remover_process_capture(a, cfg, sc, v, sc.s.saturated_microphone_signal != 0, rec.est_has, rec.est_delay, up, b);
}

// EchoCanceller3::ProcessCapture for one capture frame (band 0 in place): the echo
// remover on the 2-3 blocks k_front sliced, BlockFramer back into the frame.
WAP_DEV void aec3_echo_capture(Aec3State& a, const EngineConfig& cfg, float* band0, const TickScratch& ts,
                               AecScratch& sc, UpperBandState* up, const void* ns_prefetch = nullptr,
                               int ns_prefetch_bytes = 0) {
  aec3_stage_scalars(a, sc);
  // The staged read indices are the ones after the last block; blocks see their own.
  const int final_blocks_read = sc.s.blocks_read, final_spectra_read = sc.s.spectra_read;
  __syncwarp();
  // BlockFramer per band: the three framers fill in lock-step, so the upper bands work on a
  // copy of the band-0 fill level.  Blocks 0 and 1 each complete a sub-frame
  // (InsertBlockAndExtractSubFrame), a third block is only buffered (InsertBlock).  One real loop:
  // a single copy of the block code in the kernel instead of three.
  const int nb = ts.n_capture_blocks == 3 ? 3 : 2;
#pragma unroll 1
  for (int b = 0; b < nb; ++b) {
    WAP_PHASE_SYNC();
    aec3_echo_block(a, cfg, sc, ts, b, up);
    // the noise suppressor's Process half follows the last block
    if (b == nb - 1 && cfg.ns_enabled && ns_prefetch) warp_prefetch_l1(ns_prefetch, ns_prefetch_bytes);
    if (b < 2) {
      const int len = sc.s.output_framer_len;
      __syncwarp();
      if (up) {
        framer_insert_and_extract_local(up->output_framer_hi[0], len, sc.x, band0 + kFrame + b * kSubFrame);
        framer_insert_and_extract_local(up->output_framer_hi[1], len, sc.rm.x_aligned, band0 + 2 * kFrame + b * kSubFrame);
      }
      framer_insert_and_extract(a.output_framer, &sc.s.output_framer_len, sc.y, band0 + b * kSubFrame);
    } else {
      if (up) {
        #pragma unroll
        for (int i = lane_id(); i < kBlock; i += 32) {
          up->output_framer_hi[0][i] = sc.x[i];
          up->output_framer_hi[1][i] = sc.rm.x_aligned[i];
        }
      }
      framer_insert(a.output_framer, &sc.s.output_framer_len, sc.y);
    }
  }
  if (nb == 2) WAP_PHASE_SYNC();   // the third block slot of the tick
  __syncwarp();
  // ApmStatsReporter::UpdateStatistics (audio_processing_impl.cc:2322-2328): a
  // one-slot queue -- while the slot is full (nobody called GetStatistics) the
  // newer statistics are discarded.
  if (lane_id() == 0) {
    Aec3Scalars& s = sc.s;
    s.blocks_read = final_blocks_read;
    s.spectra_read = final_spectra_read;
    if (!s.stats_slot_full) {
      s.stats_slot_full = 1;
      s.stats_erl_time_domain = s.erl_time_domain;
      s.stats_erle_log2 = s.fb_erle_time_domain_log2;
      // BlockProcessorImpl::GetMetrics reports RenderDelayBuffer::Delay() == ComputeDelay()
      // (render_delay_buffer.cc:57), not the aligned delay_.
      s.stats_has_delay = 1;
      s.stats_delay_blocks = rdb_compute_delay(s);
    }
  }
  aec3_unstage_scalars(a, sc);
}

}  // namespace wap
