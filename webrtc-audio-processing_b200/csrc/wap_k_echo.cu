// k_echo<class>: everything after the front end for one leg (wap_pipeline.cuh: echo_stream_tick), one
// warp per call leg.  This file is compiled once per config class (-DWAP_ECHO_CLASS=0..4, see
// build.py): each instance has the band-split / upper-band / resampler / stereo code it does not
// need compiled out, and the instances build in parallel.
#include "wap_kernels.h"
#include "wap_launch.h"
#include "wap_pipeline.cuh"

#ifndef WAP_ECHO_CLASS
#error "compile with -DWAP_ECHO_CLASS=<0..5>"
#endif
#ifndef WAP_ECHO_MINBLOCKS
#if WAP_ECHO_CLASS == 5
#define WAP_ECHO_MINBLOCKS 6   // no AEC3: the noise suppressor and the band filters fit 80 registers
#else
#define WAP_ECHO_MINBLOCKS 4
#endif
#endif
#if WAP_EC3_RUNTIME
#define WAP_KSUF(x) x##_rt
#else
#define WAP_KSUF(x) x
#endif
#define WAP_CAT2(a, b) a##b
#define WAP_CAT(a, b) WAP_CAT2(a, b)

namespace wap {

template <int kClass>
__global__ void __launch_bounds__(128, WAP_ECHO_MINBLOCKS) WAP_KSUF(k_echo)(TickArgs a, int scratch_floats) {
  float* sm = reinterpret_cast<float*>(WAP_DYN_SMEM());
  const int warp = threadIdx.x >> 5;
  const int wpb = blockDim.x >> 5;
  unsigned scratch_off = (unsigned)warp * (unsigned)scratch_floats;
#if !defined(WAP_EMU)
  // Opaque to the optimiser: keeps the per-warp offset in one register instead of
  // re-deriving it from tid / the kernel parameter at every shared-memory access.
  asm volatile("" : "+r"(scratch_off));
#endif
  float* scratch = sm + scratch_off;
#if WAP_ECHO_LOCKSTEP
  // every warp of the CTA makes the same number of trips; a warp without a leg only keeps the phase points
  for (int base = blockIdx.x * wpb; base < a.n; base += gridDim.x * wpb) {
    const int idx = base + warp;
    echo_stream_tick<kClass>(a, idx < a.n ? idx : -1, scratch);
    __syncwarp();
  }
#else
  for (int idx = blockIdx.x * wpb + warp; idx < a.n; idx += gridDim.x * wpb) {
    echo_stream_tick<kClass>(a, idx, scratch);
    __syncwarp();
  }
#endif
}

cudaError_t WAP_CAT(WAP_KSUF(launch_k_echo), WAP_CAT(_, WAP_ECHO_CLASS))(int grid, int block, size_t smem, cudaStream_t stream, const TickArgs& a,
                                                     int scratch_floats) {
  WAP_LAUNCH(WAP_KSUF(k_echo)<WAP_ECHO_CLASS>, grid, block, smem, stream, a, scratch_floats);
  return cudaSuccess;
}
cudaError_t WAP_CAT(WAP_KSUF(set_k_echo_smem), WAP_CAT(_, WAP_ECHO_CLASS))(int bytes) {
  return cudaFuncSetAttribute(WAP_KSUF(k_echo)<WAP_ECHO_CLASS>, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes);
}

#if WAP_ECHO_CLASS == 0
#define WAP_L(N) WAP_CAT(WAP_KSUF(launch_k_echo), _##N)
#define WAP_S(N) WAP_CAT(WAP_KSUF(set_k_echo_smem), _##N)
cudaError_t WAP_KSUF(launch_k_echo)(int cls, int grid, int block, size_t smem, cudaStream_t stream, const TickArgs& a, int scratch_floats) {
  switch (cls) {
    case kEchoMono16k: return WAP_L(1)(grid, block, smem, stream, a, scratch_floats);
    case kEchoMono48kNative: return WAP_L(2)(grid, block, smem, stream, a, scratch_floats);
    case kEchoMono48kVia32k: return WAP_L(3)(grid, block, smem, stream, a, scratch_floats);
    case kEchoMono32k: return WAP_L(4)(grid, block, smem, stream, a, scratch_floats);
    case kEchoNoAec: return WAP_L(5)(grid, block, smem, stream, a, scratch_floats);
    default: return WAP_L(0)(grid, block, smem, stream, a, scratch_floats);
  }
}
cudaError_t WAP_KSUF(set_k_echo_smem)(int bytes) {
  cudaError_t e = WAP_S(0)(bytes);
  if (e == cudaSuccess) e = WAP_S(1)(bytes);
  if (e == cudaSuccess) e = WAP_S(2)(bytes);
  if (e == cudaSuccess) e = WAP_S(3)(bytes);
  if (e == cudaSuccess) e = WAP_S(4)(bytes);
  if (e == cudaSuccess) e = WAP_S(5)(bytes);
  return e;
}
int WAP_KSUF(k_echo_scratch_floats)(int bands, int cls) {
  return cls == kEchoNoAec ? echo_scratch_floats_no_aec(bands) : echo_scratch_floats(bands);
}
#if !WAP_EC3_RUNTIME
int k_echo_min_blocks() { return WAP_ECHO_MINBLOCKS; }
#endif
#endif

}  // namespace wap
