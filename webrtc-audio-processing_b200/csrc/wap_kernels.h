// The two heavy tick kernels live in their own translation units (wap_k_delay.cu; wap_k_echo.cu,
// compiled once per config class with -DWAP_ECHO_CLASS=N) so that the library builds in parallel;
// the engine reaches them through these launchers.
#pragma once
#include "wap_tick.h"
#if !defined(WAP_EMU)
#include <cuda_runtime.h>
#endif

namespace wap {

cudaError_t launch_k_delay(int grid, int block, size_t smem, cudaStream_t stream, const TickArgs& a, int scratch_floats);
cudaError_t set_k_delay_smem(int bytes);
// cls: wap::EchoClass (wap_pipeline.cuh)
cudaError_t launch_k_echo(int cls, int grid, int block, size_t smem, cudaStream_t stream, const TickArgs& a, int scratch_floats);
cudaError_t set_k_echo_smem(int bytes);
int k_echo_min_blocks();

#define WAP_DECLARE_ECHO_CLASS(N)                                                                            \
  cudaError_t launch_k_echo_##N(int grid, int block, size_t smem, cudaStream_t stream, const TickArgs& a, \
                                int scratch_floats);                                                         \
  cudaError_t set_k_echo_smem_##N(int bytes);
WAP_DECLARE_ECHO_CLASS(0)
WAP_DECLARE_ECHO_CLASS(1)
WAP_DECLARE_ECHO_CLASS(2)
WAP_DECLARE_ECHO_CLASS(3)
WAP_DECLARE_ECHO_CLASS(4)

}  // namespace wap
