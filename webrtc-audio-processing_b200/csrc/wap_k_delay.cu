// k_delay: AEC3 delay estimation (matched filters, lag aggregation, delay controller) for the
// capture blocks of one tick, one warp per call leg.  See wap_pipeline.cuh: delay_stream_tick.
#include "wap_kernels.h"
#include "wap_launch.h"
#include "wap_pipeline.cuh"

#if WAP_EC3_RUNTIME
#define WAP_KSUF(x) x##_rt
#else
#define WAP_KSUF(x) x
#endif

namespace wap {

#ifndef WAP_DELAY_MINBLOCKS
#define WAP_DELAY_MINBLOCKS 5
#endif
__global__ void __launch_bounds__(128, WAP_DELAY_MINBLOCKS) WAP_KSUF(k_delay)(TickArgs a, int scratch_floats) {
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
  for (int idx = blockIdx.x * wpb + warp; idx < a.n; idx += gridDim.x * wpb) {
    delay_stream_tick(a, idx, scratch);
    __syncwarp();
  }
}

cudaError_t WAP_KSUF(launch_k_delay)(int grid, int block, size_t smem, cudaStream_t stream, const TickArgs& a, int scratch_floats) {
  WAP_LAUNCH(WAP_KSUF(k_delay), grid, block, smem, stream, a, scratch_floats);
  return cudaSuccess;
}
cudaError_t WAP_KSUF(set_k_delay_smem)(int bytes) {
  return cudaFuncSetAttribute(WAP_KSUF(k_delay), cudaFuncAttributeMaxDynamicSharedMemorySize, bytes);
}

int WAP_KSUF(k_delay_scratch_floats)() { return delay_scratch_floats(); }

}  // namespace wap
