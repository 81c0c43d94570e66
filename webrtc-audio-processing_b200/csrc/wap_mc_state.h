// State of a multi-channel call leg (BASELINE config 4: stereo frames with
// pipeline.multi_channel_render / _capture on and AEC3 enabled).
//
// The reference's EchoCanceller3 then runs with C capture channels and R render channels (R = 1 with the
// render frame downmixed while no stereo content has been detected, R = the frame's channel count with
// the multichannel EchoCanceller3Config afterwards; echo_canceller3.cc:790-811,969-1002).  What is
// shared between the capture channels and what exists once per channel follows the reference's classes:
//
//   shared      : RenderDelayBuffer (rings with a render-channel dimension), delay estimation (on the
//                 AlignmentMixer outputs: the mono structures of wap_state.h, used by k_delay unchanged),
//                 RenderSignalAnalyzer, AecState's state machines, ErlEstimator, the reverb models of the
//                 render power, render noise floor, SuppressionGain::last_gain_, the comfort-noise seed
//   per channel : Subtractor (both filters with R render channels each, update gains, misadjustment),
//                 FilterAnalyzer state, ERLE estimators, ReverbFrequencyResponse, comfort noise spectra,
//                 nearend / echo smoothing, DominantNearendDetector counters, the time-domain memories
//
// Layout: the leg keeps its StreamState (wap_state.h): `aec` there holds everything shared (and the
// delay-estimation state k_delay works on; its mono render rings stay unused), `tick` the k_front ->
// k_delay -> k_echo hand-over.  McState adds the render rings with their channel dimension and, per capture
// channel, a McChan (the per-channel vectors and an Aec3Scalars of which only the per-channel members are
// used) plus the adaptive filters, whose partitions are indexed v = p * R + render_channel ("virtual
// partitions": the reference's loops are `for p { for ch }`, adaptive_fir_filter_avx2.cc:60-196).
#pragma once

#include "wap_state.h"

namespace wap {

constexpr int kMcCh = 2;                              // maximum render / capture channels
constexpr int kMcVParts = kMaxPartitions * kMcCh;     // virtual partitions of one adaptive filter

struct alignas(16) McRender {   // RenderDelayBuffer rings (render_delay_buffer.cc:72-101) with a channel dimension
  float blocks[kRingBlocks][kMcCh][kMaxBands][kBlock];   // BlockBuffer (all bands)
  float fft_re[kRingBlocks][kMcCh][kBinsPad];            // FftBuffer (band 0)
  float fft_im[kRingBlocks][kMcCh][kBinsPad];
  float spectra[kRingBlocks][kMcCh][kBinsPad];           // SpectrumBuffer
};

struct alignas(16) McFilters {  // one capture channel's Subtractor filters (adaptive_fir_filter.h)
  float Hr_re[kMcVParts][kBinsPad];
  float Hr_im[kMcVParts][kBinsPad];
  float Hc_re[kMcVParts][kBinsPad];
  float Hc_im[kMcVParts][kBinsPad];
};

// Per-capture-channel estimator state: the members of Aec3State (wap_state.h, same names) that exist once
// per capture channel in the reference.
struct alignas(16) McChan {
  float H2[kMaxPartitions][kBinsPad];       // refined_frequency_responses_[ch] (max over the render channels)
  float h_time[kMaxPartitions * kBlock];    // refined_impulse_responses_[ch]
  float h_highpass[kMaxPartitions * kBlock];  // FilterAnalyzer::h_highpass_[ch]
  float H_error[kBinsPad];
  float erle[kBinsPad], erle_onset_comp[kBinsPad], erle_unbounded[kBinsPad];
  float accum_Y2[kBinsPad], accum_E2[kBinsPad];
  float tail_response[kBinsPad];
  float cng_Y2_smoothed[kBinsPad], cng_N2[kBinsPad], cng_N2_initial[kBinsPad];
  float last_nearend[kBinsPad], last_echo[kBinsPad];
  float nearend_mem[3][kBinsPad];
  int erle_hold_counters[kBinsPad];
  int accum_low_render[kBinsPad];
  int coming_onset[kBinsPad];
  float e_old[kBlock], y_old[kBlock], e_output_old[kBlock];
  Aec3Scalars s;
};

// MultiChannelContentDetector (multi_channel_content_detector.h:70-87) and what EchoCanceller3 derives from it.
struct McDetector {
  int persistent;            // persistent_multichannel_content_detected_
  int temporary;             // temporary_multichannel_content_detected_
  int consecutive_frames_with_stereo;
  int frames_since_stereo_detected_last;
  int render_channels_to_aec;  // num_render_channels_to_aec_: 1 or the frame's channel count
  int pad_[3];
};

// AlignmentMixer (alignment_mixer.h:47-57), adaptive variant.
struct McMixer {
  float cumulative_energies[kMcCh];
  int strong_block_counters[2];
  int block_counter;
  int selected_channel;
  int pad_[2];
};

struct alignas(16) McChanIo {   // per capture channel: front end and framing state
  Biquad hpf[3];                           // HighPassFilter, this channel
  float capture_blocker[kMaxBands][kBlock];  // FrameBlocker (all bands fill in lock-step: one length)
  float output_framer[kMaxBands][kBlock];    // BlockFramer
  ThreeBandState bands;                      // SplittingFilter of the capture AudioBuffer, this channel
};
struct alignas(16) McRenderIo { // per render channel
  float render_blocker[kMaxBands][kBlock];
  ThreeBandState bands;
};

// k_front -> k_echo hand-over of a multi-channel leg (the mono TickScratch keeps the mixed, decimated
// capture blocks and the block records for k_delay).
struct alignas(16) McTick {
  float render_blocks[3][kMcCh][kMaxBands][kBlock];
  float capture_blocks[3][kMcCh][kMaxBands][kBlock];
  float capture_frame[kMcCh][kFrame * kMaxBands];   // k_mc_echo: the processed bands; then the merged full-band frame
  int render_channels[3];    // channels of render block r (1: downmixed, 2: as is)
  int gain_change;           // echo_path_gain_change of this tick's capture blocks
  // k_mc_echo -> k_mc_post (48 kHz): the post level adjustment's gain ramp of this frame (mode 0: none)
  float post_gain_prev, post_gain_target;
  int post_gain_on, pad_post_;
};

struct alignas(16) McState {
  McRender render;
  McFilters filt[kMcCh];
  McChan chan[kMcCh];
  McChanIo cio[kMcCh];
  McRenderIo rio[kMcCh];
  float e_output_old_hi[kMcCh][2][kBlock];   // SuppressionFilter::e_output_old_[1..2][ch]
  Biquad post_filter[kMcCh][4];              // PostFilter, 48 kHz only
  McDetector det;
  McMixer render_mixer, capture_mixer;
  int capture_blocker_len, output_framer_len, render_blocker_len, pad_;
  McTick tick;
};

// Freshly constructed state of the two EchoCanceller3Config variants (index 0: the mono config, 1: the
// multichannel config), copied into a leg by EchoCanceller3::Initialize (mc_initialize).
struct alignas(16) McTemplates {
  Aec3State aec[2];
  McChan chan[2];
};

// Multi-channel parameters that are not part of Ec3Params (EchoCanceller3Config::{multi_channel,
// delay.render_alignment_mixing, delay.capture_alignment_mixing}).
struct McParams {
  int detect_stereo_content;
  float detection_threshold;
  int timeout_frames;        // 0: no timeout
  int hysteresis_frames;
  int render_mix_downmix, render_mix_adaptive, render_mix_prefer_first_two;
  float render_mix_threshold;
  int capture_mix_downmix, capture_mix_adaptive, capture_mix_prefer_first_two;
  float capture_mix_threshold;
};

}  // namespace wap
