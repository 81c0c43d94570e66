// The two heavy tick kernels live in their own translation units (wap_k_delay.cu; wap_k_echo.cu,
// compiled once per config class with -DWAP_ECHO_CLASS=N) so that the library builds in parallel;
// the engine reaches them through these launchers.
#pragma once
#include "wap_tick.h"
#if !defined(WAP_EMU)
#include <cuda_runtime.h>
#endif

namespace wap {

// Two instances of every heavy kernel: the default EchoCanceller3Config on compile-time constants
// (no suffix) and the run-time-parameter build (_rt, -DWAP_EC3_RUNTIME=1) for engines created with
// another config.  cls: wap::EchoClass (wap_pipeline.cuh).
#define WAP_DECLARE_ECHO_CLASS(S, N)                                                                              \
  cudaError_t launch_k_echo##S##_##N(int grid, int block, size_t smem, cudaStream_t stream, const TickArgs& a, \
                                     int scratch_floats);                                                         \
  cudaError_t set_k_echo_smem##S##_##N(int bytes);
#define WAP_DECLARE_KERNELS(S)                                                                                          \
  cudaError_t launch_k_delay##S(int grid, int block, size_t smem, cudaStream_t stream, const TickArgs& a,             \
                                int scratch_floats);                                                                    \
  cudaError_t set_k_delay_smem##S(int bytes);                                                                           \
  int k_delay_scratch_floats##S();                                                                                      \
  cudaError_t launch_k_echo##S(int cls, int grid, int block, size_t smem, cudaStream_t stream, const TickArgs& a,     \
                               int scratch_floats);                                                                     \
  cudaError_t set_k_echo_smem##S(int bytes);                                                                            \
  int k_echo_scratch_floats##S(int bands, int cls);                                                                              \
  WAP_DECLARE_ECHO_CLASS(S, 0)                                                                                          \
  WAP_DECLARE_ECHO_CLASS(S, 1)                                                                                          \
  WAP_DECLARE_ECHO_CLASS(S, 2)                                                                                          \
  WAP_DECLARE_ECHO_CLASS(S, 3)                                                                                          \
  WAP_DECLARE_ECHO_CLASS(S, 4)                                                                                          \
  WAP_DECLARE_ECHO_CLASS(S, 5)
WAP_DECLARE_KERNELS()
WAP_DECLARE_KERNELS(_rt)
int k_echo_min_blocks();

// Multi-channel legs (wap_k_mc.cu)
int k_mc_front_scratch_floats();
int k_mc_echo_scratch_floats();
cudaError_t set_k_mc_smem(int front_bytes, int echo_bytes);
cudaError_t launch_k_mc_front(int grid, int block, size_t smem, cudaStream_t stream, const TickArgs& a, int scratch_floats);
cudaError_t launch_k_mc_echo(int grid, int block, size_t smem, cudaStream_t stream, const TickArgs& a, int scratch_floats);
cudaError_t launch_k_mc_post(cudaStream_t stream, const TickArgs& a);

}  // namespace wap
