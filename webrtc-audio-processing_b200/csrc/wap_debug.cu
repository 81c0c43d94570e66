// Stage-level debug entry points (C ABI, declared in include/wap_debug.h):
// run ONE DSP primitive on the device for unit parity tests against the
// reference.  Not part of the hot path.
#include <stdio.h>
#include <vector>

#include "dsp_fft.cuh"
#include "wap_libm.cuh"
#include "wap_launch.h"

namespace wap {

// data: [count][128]; two transforms per warp.
__global__ void k_dbg_fft128(float* data, int count, int inverse) {
  float* sm = reinterpret_cast<float*>(WAP_DYN_SMEM());
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int half = lane >> 4, t = lane & 15;
  const int idx = (blockIdx.x * (blockDim.x >> 5) + warp) * 2 + half;
  float* a = sm + (warp * 2 + half) * 128;
  const bool on = idx < count;
  if (on) for (int i = t; i < 128; i += 16) a[i] = data[(size_t)idx * 128 + i];
  __syncwarp();
  if (inverse) fft128_inverse(a, t, on); else fft128_forward(a, t, on);
  if (on) for (int i = t; i < 128; i += 16) data[(size_t)idx * 128 + i] = a[i];
}

__global__ void k_dbg_fft256(float* data, int count, int inverse) {
  float* sm = reinterpret_cast<float*>(WAP_DYN_SMEM());
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int idx = blockIdx.x * (blockDim.x >> 5) + warp;
  if (idx >= count) return;
  float* a = sm + warp * 256;
  for (int i = lane; i < 256; i += 32) a[i] = data[(size_t)idx * 256 + i];
  __syncwarp();
  if (inverse) fft256_inverse(a, lane); else fft256_forward(a, lane);
  for (int i = lane; i < 256; i += 32) data[(size_t)idx * 256 + i] = a[i];
}

// which: 0 = powf(2, x), 1 = tanhf(x), 2 = (float)(0.5 * (tanh(x) + 1)) with the double tanh; in place.
__global__ void k_dbg_libm(float* data, int count, int which) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= count) return;
  if (which == 3) data[i] = fdiv(data[i], 500.f);
  else if (which == 2) data[i] = (float)(0.5f * (libm_tanh((double)data[i]) + 1.f));  // the NS prior-model indicators
  else data[i] = which == 0 ? libm_pow2f(data[i]) : libm_tanhf(data[i]);
}

}  // namespace wap

#define WAPDBG_CHECK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { \
  fprintf(stderr, "wapdbg: %s failed: %s\n", #x, cudaGetErrorString(e_)); return -1; } } while (0)

extern "C" {

int wapdbg_fft128(float* host, int count, int inverse) {
  float* d = nullptr;
  size_t bytes = (size_t)count * 128 * sizeof(float);
  WAPDBG_CHECK(cudaMalloc((void**)&d, bytes));
  WAPDBG_CHECK(cudaMemcpy(d, host, bytes, cudaMemcpyHostToDevice));
  int warps = (count + 1) / 2, wpb = 4, blocks = (warps + wpb - 1) / wpb;
  WAP_LAUNCH(wap::k_dbg_fft128, blocks, wpb * 32, wpb * 2 * 128 * sizeof(float), 0, d, count, inverse);
  WAPDBG_CHECK(cudaDeviceSynchronize());
  WAPDBG_CHECK(cudaGetLastError());
  WAPDBG_CHECK(cudaMemcpy(host, d, bytes, cudaMemcpyDeviceToHost));
  cudaFree(d);
  return 0;
}

int wapdbg_fft256(float* host, int count, int inverse) {
  float* d = nullptr;
  size_t bytes = (size_t)count * 256 * sizeof(float);
  WAPDBG_CHECK(cudaMalloc((void**)&d, bytes));
  WAPDBG_CHECK(cudaMemcpy(d, host, bytes, cudaMemcpyHostToDevice));
  int wpb = 4, blocks = (count + wpb - 1) / wpb;
  WAP_LAUNCH(wap::k_dbg_fft256, blocks, wpb * 32, wpb * 256 * sizeof(float), 0, d, count, inverse);
  WAPDBG_CHECK(cudaDeviceSynchronize());
  WAPDBG_CHECK(cudaGetLastError());
  WAPDBG_CHECK(cudaMemcpy(host, d, bytes, cudaMemcpyDeviceToHost));
  cudaFree(d);
  return 0;
}

int wapdbg_libm(float* host, int count, int which) {
  float* d = nullptr;
  size_t bytes = (size_t)count * sizeof(float);
  WAPDBG_CHECK(cudaMalloc((void**)&d, bytes));
  WAPDBG_CHECK(cudaMemcpy(d, host, bytes, cudaMemcpyHostToDevice));
  WAP_LAUNCH(wap::k_dbg_libm, (count + 127) / 128, 128, 0, 0, d, count, which);
  WAPDBG_CHECK(cudaDeviceSynchronize());
  WAPDBG_CHECK(cudaGetLastError());
  WAPDBG_CHECK(cudaMemcpy(host, d, bytes, cudaMemcpyDeviceToHost));
  cudaFree(d);
  return 0;
}

}  // extern "C"
