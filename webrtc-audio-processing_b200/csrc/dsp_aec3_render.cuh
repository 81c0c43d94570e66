// RenderDelayBufferImpl (reference aec3/render_delay_buffer.cc:161-500) and the
// AEC3 decimator (aec3/decimator.cc:75-91): the render rings of one call leg.
#pragma once

#include "dsp_aec3_common.cuh"

namespace wap {

// ---- scalar helpers (lane 0, on the staged scalars) -----------------------
// BufferLatency (render_delay_buffer.cc:447-452)
WAP_DEV int rdb_buffer_latency(const Aec3Scalars& s) {
  const int latency_samples = (kLowRateSize + s.lr_read - s.lr_write) % kLowRateSize;
  return latency_samples / kSubBlock;
}
// ApplyTotalDelay (:379-385)
WAP_DEV void rdb_apply_total_delay(Aec3Scalars& s, int delay) {
  s.blocks_read = ring_off(s.blocks_write, -delay, kRingBlocks);
  s.spectra_read = ring_off(s.spectra_write, delay, kRingBlocks);
}
// ComputeDelay (:368-376)
WAP_DEV int rdb_compute_delay(const Aec3Scalars& s) {
  const int latency_blocks = rdb_buffer_latency(s);
  const int internal_delay = s.spectra_read >= s.spectra_write ? s.spectra_read - s.spectra_write
                                                               : kRingBlocks + s.spectra_read - s.spectra_write;
  return internal_delay - latency_blocks;
}
// Reset (:161-197)
WAP_DEV void rdb_reset(Aec3Scalars& s, int default_delay) {
  s.last_call_was_render = 0;
  s.num_api_calls_in_a_row = 1;
  s.min_latency_blocks = 0;
  s.excess_render_detection_counter = 0;
  s.lr_read = ring_off(s.lr_write, kSubBlock, kLowRateSize);
  if (s.has_external_delay) {
    const int headroom = 2;
    int to_set = (s.external_delay <= headroom) ? 1 : s.external_delay - headroom;
    to_set = imin(to_set, kMaxRingDelay);
    rdb_apply_total_delay(s, to_set);
    s.delay = rdb_compute_delay(s);
    s.has_delay = 1;
    s.external_delay_verified = 0;
  } else {
    rdb_apply_total_delay(s, default_delay);
    s.has_delay = 0;
  }
}
// IncrementReadIndices (:472-478)
WAP_DEV void rdb_increment_read_indices(Aec3Scalars& s) {
  if (s.blocks_read != s.blocks_write) {
    s.blocks_read = ring_inc(s.blocks_read, kRingBlocks);
    s.spectra_read = ring_dec(s.spectra_read, kRingBlocks);
  }
}
// AlignFromDelay (:304-328); returns whether the delay changed.
WAP_DEV bool rdb_align_from_delay(Aec3Scalars& s, int delay) {
  if (!s.external_delay_verified && s.has_external_delay && s.has_delay) s.external_delay_verified = 1;
  if (s.has_delay && s.delay == delay) return false;
  s.delay = delay;
  s.has_delay = 1;
  int total = rdb_buffer_latency(s) + delay;
  total = imin(kMaxRingDelay, imax(total, 0));
  rdb_apply_total_delay(s, total);
  return true;
}
// SetAudioBufferDelay (:330-344): ms -> blocks (rounded down).
WAP_DEV void rdb_set_audio_buffer_delay(Aec3Scalars& s, int delay_ms, int fixed_capture_delay_samples = 0) {
  s.external_delay = (delay_ms * 16 + fixed_capture_delay_samples) / (4 * 16);
  s.has_external_delay = 1;
}
// PrepareCaptureProcessing (:249-301)
// default_delay / interval / max_excess: delay.default_delay, buffering.excess_render_detection_interval_blocks,
// buffering.max_allowed_excess_render_blocks
WAP_DEV int rdb_prepare_capture_processing(Aec3Scalars& s, int default_delay, int interval, int max_excess) {
  int event = kEventNone;
  if (s.has_delay) {
    if (s.last_call_was_render) {
      s.last_call_was_render = 0;
      s.num_api_calls_in_a_row = 1;
    } else if (++s.num_api_calls_in_a_row > s.max_observed_jitter) {
      s.max_observed_jitter = s.num_api_calls_in_a_row;
    }
  }
  // DetectExcessRenderBlocks (:420-444)
  bool excess = false;
  const int latency_blocks = rdb_buffer_latency(s);
  s.min_latency_blocks = imin(s.min_latency_blocks, latency_blocks);
  if (++s.excess_render_detection_counter >= interval) {
    excess = s.min_latency_blocks > max_excess;
    s.min_latency_blocks = latency_blocks;
    s.excess_render_detection_counter = 0;
  }
  if (excess) {
    rdb_reset(s, default_delay);
    event = kEventRenderOverrun;
  } else if (s.lr_read == s.lr_write) {  // RenderUnderrun
    rdb_increment_read_indices(s);
    if (s.has_delay && (unsigned)s.delay > 0u) s.delay = s.delay - 1;
    event = kEventRenderUnderrun;
  } else {
    s.lr_read = ring_off(s.lr_read, -kSubBlock, kLowRateSize);
    rdb_increment_read_indices(s);
  }
  s.rb_render_activity = s.render_activity;
  if (s.render_activity) {
    s.render_activity_counter = 0;
    s.render_activity = 0;
  }
  return event;
}

// Vector half of RenderDelayBufferImpl::InsertBlock (:387-429) for the block in sc.x:
// block ring, FFT of [previous block | new block], power spectrum.  The scalar half
// (indices, events, decimation into the low-rate ring) already ran in k_front, which
// left the ring positions in `rec`.
WAP_DEV void aec3_render_insert_vector(Aec3State& a, AecScratch& sc, const RenderInsertRec& rec) {
  const int lane = lane_id();
  const int previous_write = rec.previous_write, bw = rec.blocks_write, sw = rec.spectra_write;
  __syncwarp();
  #pragma unroll
  for (int i = lane; i < kBlock; i += 32) {
    a.blocks[bw][i] = sc.x[i];
    sc.fftA[i] = a.blocks[previous_write][i];
    sc.fftA[kBlock + i] = sc.x[i];
  }
  fft_pair(sc, false, false);
  #pragma unroll
  for (int k = lane; k < kBins; k += 32) {
    float re, im;
    if (k == 0) { re = sc.fftA[0]; im = 0.f; }
    else if (k == 64) { re = sc.fftA[1]; im = 0.f; }
    else { re = sc.fftA[2 * k]; im = sc.fftA[2 * k + 1]; }
    a.fft_re[sw][k] = re;
    a.fft_im[sw][k] = im;
    a.spectra[sw][k] = power_bin(re, im, k);
  }
  __syncwarp();
}

}  // namespace wap
