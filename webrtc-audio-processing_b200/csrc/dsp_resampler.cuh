// Windowed-sinc resampler between the API rate and the processing rate, one warp
// per call leg:
//   PushSincResampler::Resample / Run     common_audio/resampler/push_sinc_resampler.cc:52-101
//   SincResampler::{Resample, UpdateRegions, Flush, InitializeKernel}
//                                         common_audio/resampler/sinc_resampler.cc:176-316
//   SincResampler::Convolve_AVX2          common_audio/resampler/sinc_resampler_avx2.cc:21-64
// (AudioBuffer owns one per direction: audio_buffer.cc:79-94.)
//
// The reference walks a double-precision read position through a buffer of
// request_frames + 32 samples and evaluates, per output sample, two 32-tap kernels
// (neighbouring sub-sample offsets) that it blends linearly.  The position
// recursion is a serial chain of double additions, but it is cheap and identical
// for every lane: all lanes run it redundantly (no exchange), lane j keeps the
// position of output j of the current group of 32, and the 32 convolutions of a
// group run in parallel.  Each convolution reproduces the AVX2 body: eight
// accumulators of four fused multiply-adds, folded 8 -> 4, blended, folded 4 -> 1.
#pragma once

#include "wap_dev.cuh"

namespace wap {

constexpr int kRsKernelSize = 32;    // SincResampler::kKernelSize
constexpr int kRsOffsetCount = 32;   // SincResampler::kKernelOffsetCount
constexpr int kRsTableFloats = kRsKernelSize * (kRsOffsetCount + 1);
constexpr int kRsMaxRequest = 960;   // API rates up to 96 kHz (processing rates up to 48 kHz)

// One PushSincResampler (+ its SincResampler).
struct alignas(16) ResamplerState {
  float buf[kRsMaxRequest + kRsKernelSize];  // input_buffer_ (zero)
  double vsi;                                // virtual_source_idx_ (0)
  int primed;                                // buffer_primed_ (0)
  int second_load;                           // UpdateRegions(true) has happened (0)
  int started;                               // !PushSincResampler::first_pass_ (0): all-zero init
  int pad_;
};
constexpr int kRsPerLeg = 3;  // render in, capture in, capture out

struct ResamplerParams {
  int request;          // source frames per push (= SincResampler::request_frames_)
  int dst_frames;       // destination frames per push
  double ratio;         // io_sample_rate_ratio_ = request / dst_frames
  const float* kernel;  // kRsTableFloats, built on the host (InitializeKernel)
};

// Convolve_AVX2 for one output sample.
WAP_DEV float rs_convolve(const float* in, const float* k1, const float* k2, double factor) {
  float s1[8], s2[8];
#pragma unroll
  for (int l = 0; l < 8; ++l) s1[l] = 0.f, s2[l] = 0.f;
#pragma unroll
  for (int m = 0; m < kRsKernelSize; m += 8) {
#pragma unroll
    for (int l = 0; l < 8; ++l) {
      const float x = in[m + l];
      s1[l] = fmaf(x, k1[m + l], s1[l]);
      s2[l] = fmaf(x, k2[m + l], s2[l]);
    }
  }
  const float w1 = (float)(1.0 - factor), w2 = (float)factor;
  float c[4];
#pragma unroll
  for (int l = 0; l < 4; ++l) {
    const float q1 = s1[l] + s1[l + 4], q2 = s2[l] + s2[l + 4];
    const float a = q1 * w1, b = q2 * w2;
    c[l] = a + b;
  }
  const float e0 = c[2] + c[0], e1 = c[3] + c[1];
  return e0 + e1;
}

// PushSincResampler::Run: the read callback.  `src` may be nullptr only while first_pass.
WAP_DEV void rs_read(ResamplerState& st, const ResamplerParams& p, const float* src, int r0, int& first_pass) {
  const int lane = lane_id();
  __syncwarp();
  if (first_pass) {
    for (int i = lane; i < p.request; i += 32) st.buf[r0 + i] = 0.f;
    first_pass = 0;
  } else {
    for (int i = lane; i < p.request; i += 32) st.buf[r0 + i] = src[i];
  }
  __syncwarp();
}

// SincResampler::Resample(frames, dst).  State scalars are carried in registers (warp-uniform).
WAP_DEV void rs_resample(ResamplerState& st, const ResamplerParams& p, const float* src, int frames, float* dst,
                         double& vsi, int& primed, int& second_load, int& first_pass) {
  const int lane = lane_id();
  int remaining = frames;
  int out = 0;
  if (!primed && remaining) {
    rs_read(st, p, src, second_load ? kRsKernelSize : kRsKernelSize / 2, first_pass);
    primed = 1;
  }
  while (remaining) {
    // r2 = 16, r4 = r0 + request - 16  =>  block_size = r4 - r2
    const int r0 = second_load ? kRsKernelSize : kRsKernelSize / 2;
    const int block_size = r0 + p.request - kRsKernelSize;
    int count = (int)ceil(((double)block_size - vsi) / p.ratio);
    if (count < 0) count = 0;
    const int n_seg = count < remaining ? count : remaining;
    for (int base = 0; base < n_seg; base += 32) {
      const int m = n_seg - base < 32 ? n_seg - base : 32;
      double mine = 0.0;
      for (int t = 0; t < m; ++t) {  // virtual_source_idx_ += io_ratio, one output at a time
        if (t == lane) mine = vsi;
        vsi += p.ratio;
      }
      if (lane < m) {
        const int source_idx = (int)mine;
        const double subsample_remainder = mine - source_idx;
        const double virtual_offset_idx = subsample_remainder * kRsOffsetCount;
        const int offset_idx = (int)virtual_offset_idx;
        const float* k1 = p.kernel + offset_idx * kRsKernelSize;
        dst[out + base + lane] =
            rs_convolve(st.buf + source_idx, k1, k1 + kRsKernelSize, virtual_offset_idx - offset_idx);
      }
    }
    out += n_seg;
    remaining -= n_seg;
    if (!remaining) break;  // the reference returns from inside the loop: no wrap yet
    // Wrap: step back one block, keep the last kernel's worth of input, read the next block.
    vsi -= block_size;
    __syncwarp();
    const float keep = st.buf[r0 + p.request - kRsKernelSize + lane];  // r3[lane]
    __syncwarp();
    st.buf[lane] = keep;                                             // r1[lane]
    second_load = 1;
    rs_read(st, p, src, kRsKernelSize, first_pass);
  }
  __syncwarp();
}

// PushSincResampler::Resample(src[request]) -> dst[dst_frames]; src / dst in shared or global memory.
WAP_DEV void rs_push(ResamplerState& st, const ResamplerParams& p, const float* src, float* dst) {
  __syncwarp();
  double vsi = st.vsi;
  int primed = st.primed, second_load = st.second_load, first_pass = !st.started;
  __syncwarp();
  if (first_pass) {
    // Prime with one chunk of silence; its output is overwritten by the real call below.
    const int block_size = kRsKernelSize / 2 + p.request - kRsKernelSize;
    const int chunk = (int)((double)block_size / p.ratio);  // SincResampler::ChunkSize
    rs_resample(st, p, src, chunk, dst, vsi, primed, second_load, first_pass);
  }
  rs_resample(st, p, src, p.dst_frames, dst, vsi, primed, second_load, first_pass);
  if (lane_id() == 0) {
    st.vsi = vsi;
    st.primed = primed;
    st.second_load = second_load;
    st.started = !first_pass;
  }
  __syncwarp();
}

}  // namespace wap
