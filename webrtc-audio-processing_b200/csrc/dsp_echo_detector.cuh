// ResidualEchoDetector (modules/audio_processing/residual_echo_detector.cc:43-209 and
// echo_detector/{circular_buffer,mean_variance_estimator,moving_max,normalized_covariance_estimator}.cc):
// the optional echo-likelihood statistic of the reference, an injected component
// (AudioProcessingBuilder::SetEchoDetector(CreateEchoDetector())).  Per 10 ms frame it compares the power of
// the processed capture frame with the powers of the last 650 render frames: one normalised covariance
// estimator per lag, the largest normalised cross-correlation is the likelihood.
//
// Render side: one thread per leg (the frame power is a serial sum).  Capture side: the leg's warp, inside
// k_echo behind the band merge (audio_processing_impl.cc:1462-1465): lane 0 runs the scalar estimators, the
// 650 lags are spread over the lanes.
#pragma once

#include "dsp_front.cuh"
#include "wap_dev.cuh"
#include "wap_state.h"
#include "wap_tick.h"

namespace wap {

constexpr float kRedAlpha = 0.001f;

// MeanVarianceEstimator::Update (mean_variance_estimator.cc:23-30)
WAP_DEV void red_mean_variance_update(float& mean, float& variance, float value) {
  mean = (1.f - kRedAlpha) * mean + kRedAlpha * value;
  variance = (1.f - kRedAlpha) * variance + kRedAlpha * (value - mean) * (value - mean);
}

// ---- ResidualEchoDetector::AnalyzeRenderAudio (residual_echo_detector.cc:56-70) on the render AudioBuffer's
// first channel (PackRenderAudioBufferForEchoDetector, audio_processing_impl.cc:137-142).  Render frames in
// front of the first capture frame never reach the capture side (the first AnalyzeCaptureAudio clears the
// buffer, :76-79, and an empty buffer restarts frames_since_zero_buffer_size_): they are skipped.
WAP_DEV void red_analyze_render(const TickArgs& a, int idx) {
  const EngineConfig& cfg = a.cfg;
  const int slot = a.slots ? a.slots[idx] : idx;
  EchoDetectorState& d = a.red[slot];
  if (!d.seen_capture) return;
  const int len = kFrame * cfg.num_bands;
  float acc = 0.f;   // Power(): std::inner_product from 0.f, then / size
  for (int i = 0; i < len; ++i) {
    const float v = front_load_sample(a.render, idx, len, a.fmt, i, cfg.render_channels, -1);
    acc = acc + v * v;
  }
  const float power = fdiv(acc, (float)len);   // a true division also where len is a compile-time constant
  if (d.rb_count == 0) {
    d.frames_since_zero_buffer_size = 0;
  } else if (d.frames_since_zero_buffer_size >= kRedRenderBuffer) {
    --d.rb_count;   // Pop
    d.frames_since_zero_buffer_size = 0;
  }
  ++d.frames_since_zero_buffer_size;
  d.render_buffer[d.rb_next] = power;   // CircularBuffer::Push
  d.rb_next = (d.rb_next + 1) % kRedRenderBuffer;
  d.rb_count = d.rb_count + 1 < kRedRenderBuffer ? d.rb_count + 1 : kRedRenderBuffer;
}

// ---- the capture side of one tick: AnalyzeCaptureAudio (:72-160) and the statistics
// (audio_processing_impl.cc:1499-1505) while the output is used; then the ApmStatsReporter slot
// (:2312-2327) like the echo remover's statistics.  `x` = first channel of the merged capture frame.
WAP_DEV void red_capture_tick(EchoDetectorState& d, const float* x, int len, bool output_used, float* exch) {
  const int lane = lane_id();
  __syncwarp();
  if (output_used) {
    const bool first = d.seen_capture == 0;
    const int rb_count = first ? 0 : d.rb_count;   // first_process_call_: render_buffer_.Clear()
    const int rb_next = first ? 0 : d.rb_next;
    const int insert = d.next_insertion_index;
    __syncwarp();
    if (lane == 0) {
      d.seen_capture = 1;
      d.rb_next = rb_next;
      d.rb_count = rb_count > 0 ? rb_count - 1 : 0;
    }
    if (rb_count > 0) {
      if (lane == 0) {
        const float render_power = d.render_buffer[(kRedRenderBuffer + rb_next - rb_count) % kRedRenderBuffer];   // Pop
        float mean = d.render_mean, variance = d.render_variance;
        red_mean_variance_update(mean, variance, render_power);
        d.render_mean = mean; d.render_variance = variance;
        d.render_power[insert] = render_power;
        d.render_power_mean[insert] = mean;
        d.render_power_std_dev[insert] = sqrtf(variance);
        float acc = 0.f;
        for (int i = 0; i < len; ++i) acc = acc + x[i] * x[i];
        const float capture_power = fdiv(acc, (float)len);
        mean = d.capture_mean; variance = d.capture_variance;
        red_mean_variance_update(mean, variance, capture_power);
        d.capture_mean = mean; d.capture_variance = variance;
        exch[0] = capture_power; exch[1] = mean; exch[2] = sqrtf(variance);
      }
      __syncwarp();
      const float capture_power = exch[0], capture_mean = exch[1], capture_std_deviation = exch[2];
      // NormalizedCovarianceEstimator::Update per lag (normalized_covariance_estimator.cc:23-34); the
      // likelihood is the largest normalised cross-correlation above 0
      float best = 0.f;
      for (int delay = lane; delay < kRedLookback; delay += 32) {
        const int ri = insert - delay < 0 ? insert - delay + kRedLookback : insert - delay;
        const float covariance = (1.f - kRedAlpha) * d.covariance[delay] +
                                 kRedAlpha * (capture_power - capture_mean) * (d.render_power[ri] - d.render_power_mean[ri]);
        d.covariance[delay] = covariance;
        const float ncc = covariance / (capture_std_deviation * d.render_power_std_dev[ri] + .0001f);
        if (ncc > best) best = ncc;
      }
      for (int m = 16; m; m >>= 1) {
        const float o = __shfl_xor_sync(WAP_FULL, best, m);
        if (o > best) best = o;
      }
      if (lane == 0) {
        const float reliability = (1.0f - kRedAlpha) * d.reliability + kRedAlpha * 1.0f;
        d.reliability = reliability;
        float likelihood = best * reliability;
        likelihood = (1.0f < likelihood) ? 1.0f : likelihood;
        d.echo_likelihood = likelihood;
        // MovingMax::Update, window 10 * 100 (moving_max.cc:27-38)
        if (d.mm_counter >= kRedAggregation - 1) d.mm_max *= 0.99f;
        else ++d.mm_counter;
        if (likelihood > d.mm_max) { d.mm_max = likelihood; d.mm_counter = 0; }
        d.next_insertion_index = insert < kRedLookback - 1 ? insert + 1 : 0;
      }
    }
    __syncwarp();
    if (lane == 0) {
      d.stats_valid = 1;
      d.stats_likelihood = d.echo_likelihood;
      d.stats_recent_max = d.mm_max;
    }
  }
  __syncwarp();
  if (lane == 0 && !d.slot_full) {
    d.slot_full = 1;
    d.slot_valid = d.stats_valid;
    d.slot_likelihood = d.stats_likelihood;
    d.slot_recent_max = d.stats_recent_max;
  }
  __syncwarp();
}

}  // namespace wap
