// Device-side basics shared by every DSP stage of the batched APM engine.
//
// Execution model: ONE WARP ADVANCES ONE STREAM (call leg).  Every branch
// that depends on stream state is therefore warp-uniform by construction; the
// 32 lanes split the 64/65/128/129-bin vectors of the algorithm.  Per-warp
// scratch lives in dynamic shared memory, persistent state in the stream's
// slab of the HBM arena (wap_state.h).
//
// Numerics contract: FP32, flush-to-zero, no implicit
// FMA contraction (nvcc -fmad=false); fmaf() appears only where the
// reference's AVX2 path has an explicit fused multiply-add.  IEEE division
// and square root (-prec-div=true -prec-sqrt=true).
#pragma once

#include <stdint.h>

#if defined(WAP_EMU)
// g++ build for the test-only emulator (tests/emu/cuda_emu.h is force-included).
#define WAP_DEVCONST static const
#define WAP_DEV static inline
#define WAP_DEV_NOINLINE static __attribute__((noinline))
#define WAP_DYN_SMEM() (emu::smem_ptr())
#else
#include <cuda_runtime.h>
#define WAP_DEVCONST static __device__ const
#define WAP_DEV static __device__ __forceinline__
// one shared copy of a large leaf routine instead of one per call site (instruction-cache footprint)
#define WAP_DEV_NOINLINE static __device__ __noinline__
#define WAP_DYN_SMEM() (wap_dyn_smem_raw)
extern __shared__ __align__(16) unsigned char wap_dyn_smem_raw[];
#endif

#define WAP_FULL 0xffffffffu

// Experiment (-DWAP_ECHO_LOCKSTEP=1): the warps of a k_echo CTA meet at a fixed sequence of points of
// a tick, so that they walk the same stretch of the kernel's (~375 KB) code at about the same time and
// share instruction-cache lines.  Every warp of the CTA passes every point exactly once per tick.
#ifndef WAP_ECHO_LOCKSTEP
#define WAP_ECHO_LOCKSTEP 0
#endif
#if WAP_ECHO_LOCKSTEP
#define WAP_PHASE_SYNC() __syncthreads()
#else
#define WAP_PHASE_SYNC() ((void)0)
#endif

namespace wap {

// IEEE division for a divisor that is (or may become, after inlining) a compile-time constant:
// with -ftz=true nvcc rewrites `x / c` into `x * (1 / c)` even under -prec-div=true, which is an
// ulp off for most x.  __fdiv_rn is not rewritten.  (`tools/div_sites.sh` lists the sites: it
// compares the number of div.rn in the PTX with and without -ftz.)
WAP_DEV float fdiv(float a, float b) {
#if defined(WAP_EMU)
  return a / b;
#else
  return __fdiv_rn(a, b);
#endif
}

// std::min / std::max semantics of the reference ((b < a) ? b : a etc.).
WAP_DEV float fminr(float a, float b) { return (b < a) ? b : a; }
WAP_DEV float fmaxr(float a, float b) { return (a < b) ? b : a; }
WAP_DEV int imin(int a, int b) { return (b < a) ? b : a; }
WAP_DEV int imax(int a, int b) { return (a < b) ? b : a; }
WAP_DEV float clampr(float x, float lo, float hi) { return fminr(fmaxr(x, lo), hi); }

WAP_DEV int lane_id() { return (int)(threadIdx.x & 31u); }

template <class T>
WAP_DEV T bcast(T v, int src) { return __shfl_sync(WAP_FULL, v, src); }

// Sum over the warp in xor-butterfly order (NOT used where the reference's
// summation order matters).
WAP_DEV float warp_sum_any_order(float v) {
  for (int m = 16; m; m >>= 1) v += __shfl_xor_sync(WAP_FULL, v, m);
  return v;
}
WAP_DEV float warp_max(float v) {
  for (int m = 16; m; m >>= 1) v = fmaxf(v, __shfl_xor_sync(WAP_FULL, v, m));
  return v;
}
WAP_DEV float warp_min(float v) {
  for (int m = 16; m; m >>= 1) v = fminf(v, __shfl_xor_sync(WAP_FULL, v, m));
  return v;
}
WAP_DEV int warp_or(int v) { return __any_sync(WAP_FULL, v); }

// Left-to-right (reference order) sum of p[0..n) -- every lane evaluates the
// same serial chain from shared memory, so the result is warp-uniform without
// a broadcast.  Used wherever the reference runs std::accumulate and the value
// feeds a threshold decision.
WAP_DEV float serial_sum(const float* p, int n) {
  float s = 0.f;
#pragma unroll 4
  for (int i = 0; i < n; ++i) s += p[i];
  return s;
}
WAP_DEV float serial_sum_sq(const float* p, int n) {
  float s = 0.f;
#pragma unroll 4
  for (int i = 0; i < n; ++i) s += p[i] * p[i];
  return s;
}
// Same chains for 16-byte aligned shared-memory arrays with n % 4 == 0: 128-bit loads.
WAP_DEV float serial_sum_sq_v4(const float* p, int n) {
  float s = 0.f;
  const float4* q = reinterpret_cast<const float4*>(p);
#pragma unroll 4
  for (int i = 0; i < n / 4; ++i) {
    const float4 v = q[i];
    s += v.x * v.x;
    s += v.y * v.y;
    s += v.z * v.z;
    s += v.w * v.w;
  }
  return s;
}

// First index of the maximum (std::max_element semantics) over p[0..n).
WAP_DEV int warp_argmax_first(const float* p, int n) {
  const int lane = lane_id();
  float best = -3.4e38f;
  int bi = 0x7fffffff;
  for (int i = lane; i < n; i += 32) {
    float v = p[i];
    if (v > best) { best = v; bi = i; }
  }
  for (int m = 16; m; m >>= 1) {
    float ov = __shfl_xor_sync(WAP_FULL, best, m);
    int oi = __shfl_xor_sync(WAP_FULL, bi, m);
    if (ov > best || (ov == best && oi < bi)) { best = ov; bi = oi; }
  }
  return bi;
}

// Non-blocking prefetch of [p, p + bytes) into L1 / L2 by the whole warp (one 128-byte line per
// lane and step).  Phases that follow find their state vectors on chip instead of paying one
// exposed DRAM/L2 round trip per small loop.
WAP_DEV void warp_prefetch_l1(const void* p, int bytes) {
#if !defined(WAP_EMU)
  const char* c = reinterpret_cast<const char*>(p);
  for (int off = lane_id() * 128; off < bytes; off += 32 * 128) asm volatile("prefetch.global.L1 [%0];" ::"l"(c + off));
#endif
}
// WAP_L2_POLICY (experiment knob, default 0): 1 = the tick-start prefetches are tagged evict_last and released with
// applypriority when the leg's tick is over; 2 = the leg's state is re-touched with an evict_first policy when the
// tick is over (dead until the next tick, 65 k legs later).
#ifndef WAP_L2_POLICY
#define WAP_L2_POLICY 0
#endif
WAP_DEV void warp_prefetch_l2(const void* p, int bytes) {
#if !defined(WAP_EMU)
  const char* c = reinterpret_cast<const char*>(p);
#if WAP_L2_POLICY == 1
  for (int off = lane_id() * 128; off < bytes; off += 32 * 128) asm volatile("prefetch.global.L2::evict_last [%0];" ::"l"(c + off));
#else
  for (int off = lane_id() * 128; off < bytes; off += 32 * 128) asm volatile("prefetch.global.L2 [%0];" ::"l"(c + off));
#endif
#endif
}
WAP_DEV void warp_l2_release(const void* p, int bytes) {
#if !defined(WAP_EMU) && WAP_L2_POLICY != 0
  const unsigned long long a0 = reinterpret_cast<unsigned long long>(p) & ~127ull;
  const unsigned long long a1 = reinterpret_cast<unsigned long long>(p) + (unsigned long long)bytes;
#if WAP_L2_POLICY == 1
  for (unsigned long long a = a0 + lane_id() * 128ull; a < a1; a += 32 * 128ull)
    asm volatile("applypriority.global.L2::evict_normal [%0], 128;" ::"l"(a));
#else
  unsigned long long pol;
  asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(pol));
  for (unsigned long long a = a0 + lane_id() * 128ull; a + 128 <= a1; a += 32 * 128ull) {
    unsigned v;
    asm volatile("ld.global.L2::cache_hint.u32 %0, [%1], %2;" : "=r"(v) : "l"(a), "l"(pol));
  }
#endif
#endif
}

WAP_DEV void warp_copy(float* dst, const float* src, int n) {
  for (int i = lane_id(); i < n; i += 32) dst[i] = src[i];
}
WAP_DEV void warp_fill(float* dst, float v, int n) {
  for (int i = lane_id(); i < n; i += 32) dst[i] = v;
}
WAP_DEV void warp_fill_i(int* dst, int v, int n) {
  for (int i = lane_id(); i < n; i += 32) dst[i] = v;
}

}  // namespace wap
