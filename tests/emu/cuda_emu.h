// Test-only CUDA execution emulator (NOT product code, never linked into
// libwap_b200.so).
//
// The development container has nvcc but no GPU.  To debug the hand-written
// kernels against the compiled reference (oracle/_ref) before spending GPU
// minutes, the kernel translation unit is ALSO compiled by g++ with this header
// force-included.  Every CUDA thread of a block becomes a coroutine; warp
// primitives (__syncwarp, __shfl_*_sync, __ballot_sync ...) are real
// rendezvous points, so a lane only ever sees another lane's writes after a
// barrier both took part in.  Between barriers lanes run one after the other,
// which is *stricter* than hardware lock-step: a missing __syncwarp() shows up
// here as a wrong answer instead of working by accident.
//
// Floating point: the emu build uses -ffp-contract=off and MXCSR FTZ|DAZ, the
// device build -fmad=false -ftz=true -prec-div=true -prec-sqrt=true, so both
// evaluate the same IEEE single-precision expression trees.
#pragma once
#ifndef WAP_EMU
#define WAP_EMU 1
#endif

#include <sys/mman.h>
#include <xmmintrin.h>

#include <algorithm>
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <functional>
#include <vector>

#define __host__
#define __device__
#define __global__
#define __forceinline__ inline __attribute__((always_inline))
#define __noinline__ __attribute__((noinline))
#define __launch_bounds__(...)
#define __constant__
#define __align__(n) __attribute__((aligned(n)))

struct uint3 {
  unsigned x, y, z;
};
struct dim3 {
  unsigned x, y, z;
  dim3(unsigned x_ = 1, unsigned y_ = 1, unsigned z_ = 1) : x(x_), y(y_), z(z_) {}
};
struct float2 {
  float x, y;
};
struct float4 {
  float x, y, z, w;
};
struct uchar4 {
  unsigned char x, y, z, w;
};
static inline float2 make_float2(float x, float y) { return float2{x, y}; }
static inline float4 make_float4(float x, float y, float z, float w) {
  return float4{x, y, z, w};
}

namespace emu {

struct Warp {
  int arrived = 0;
  unsigned gen = 0;
  uint64_t xchg[32];
  void* first_trace[8] = {};  // call stack of the first lane at the current barrier (divergence check)
  int first_n = 0;
};

struct Thread {
  void* sp = nullptr;
  uint3 tid{0, 0, 0};
  Warp* warp = nullptr;
  bool done = false;
  void* wait_pc = nullptr;  // call site of the barrier this lane is parked at (deadlock reports)
};

struct Block {
  std::vector<Thread> threads;
  std::vector<Warp> warps;
  int cur = 0;
  int alive = 0;
  int barrier_arrived = 0;
  unsigned barrier_gen = 0;
  void* main_sp = nullptr;
  std::function<void()> body;
  unsigned char* smem = nullptr;
};

extern Block* g_block;
extern Thread* g_cur;
extern uint3 g_blockIdx;
extern dim3 g_blockDim;
extern dim3 g_gridDim;

extern "C" void emu_switch(void** save_sp, void* new_sp);

void yield_next();
void warp_barrier();
void block_barrier();
void run_block(const std::function<void()>& body, dim3 grid, dim3 block, size_t smem_bytes);

template <class K, class... Args>
void launch(K kernel, dim3 grid, dim3 block, size_t smem_bytes, Args... args) {
  std::function<void()> body = [=]() { kernel(args...); };
  run_block(body, grid, block, smem_bytes);
}

inline unsigned char* smem_ptr() { return g_block->smem; }

template <class T>
inline uint64_t to_bits(T v) {
  static_assert(sizeof(T) <= 8, "shuffle payload too wide");
  uint64_t b = 0;
  std::memcpy(&b, &v, sizeof(T));
  return b;
}
template <class T>
inline T from_bits(uint64_t b) {
  T v;
  std::memcpy(&v, &b, sizeof(T));
  return v;
}

template <class T>
inline T shfl_idx(T v, int src) {
  Warp& w = *g_cur->warp;
  int lane = g_cur->tid.x & 31;
  w.xchg[lane] = to_bits(v);
  warp_barrier();
  T r = from_bits<T>(w.xchg[src & 31]);
  warp_barrier();
  return r;
}

}  // namespace emu

#define threadIdx (emu::g_cur->tid)
#define blockIdx (emu::g_blockIdx)
#define blockDim (emu::g_blockDim)
#define gridDim (emu::g_gridDim)

static inline void __syncwarp(unsigned = 0xffffffffu) { emu::warp_barrier(); }
static inline void __syncthreads() { emu::block_barrier(); }
static inline unsigned __activemask() { return 0xffffffffu; }

template <class T>
static inline T __shfl_sync(unsigned, T v, int src, int width = 32) {
  int lane = threadIdx.x & 31;
  int base = lane & ~(width - 1);
  return emu::shfl_idx(v, base + (src & (width - 1)));
}
template <class T>
static inline T __shfl_xor_sync(unsigned, T v, int m, int width = 32) {
  int lane = threadIdx.x & 31;
  (void)width;
  return emu::shfl_idx(v, lane ^ m);
}
template <class T>
static inline T __shfl_down_sync(unsigned, T v, unsigned d, int width = 32) {
  int lane = threadIdx.x & 31;
  int src = lane + (int)d;
  if ((src & ~(width - 1)) != (lane & ~(width - 1))) src = lane;
  return emu::shfl_idx(v, src);
}
template <class T>
static inline T __shfl_up_sync(unsigned, T v, unsigned d, int width = 32) {
  int lane = threadIdx.x & 31;
  int src = lane - (int)d;
  if (src < (lane & ~(width - 1))) src = lane;
  return emu::shfl_idx(v, src);
}
static inline unsigned __ballot_sync(unsigned, int pred) {
  emu::Warp& w = *emu::g_cur->warp;
  int lane = threadIdx.x & 31;
  w.xchg[lane] = pred ? 1u : 0u;
  emu::warp_barrier();
  unsigned r = 0;
  for (int i = 0; i < 32; ++i) r |= (unsigned)(w.xchg[i] & 1u) << i;
  emu::warp_barrier();
  return r;
}
static inline int __any_sync(unsigned m, int pred) { return __ballot_sync(m, pred) != 0; }
static inline int __all_sync(unsigned m, int pred) { return __ballot_sync(m, pred) == 0xffffffffu; }

static inline unsigned __float_as_uint(float f) { return emu::from_bits<unsigned>(emu::to_bits(f)); }
static inline int __float_as_int(float f) { return emu::from_bits<int>(emu::to_bits(f)); }
static inline long long __double_as_longlong(double d) { return emu::from_bits<long long>(emu::to_bits(d)); }
static inline double __longlong_as_double(long long v) { return emu::from_bits<double>(emu::to_bits(v)); }
static inline float __uint_as_float(unsigned u) { return emu::from_bits<float>(emu::to_bits(u)); }
static inline float __int_as_float(int u) { return emu::from_bits<float>(emu::to_bits(u)); }
template <class T>
static inline T __ldg(const T* p) { return *p; }
static inline int __popc(unsigned v) { return __builtin_popcount(v); }
static inline int __ffs(int v) { return __builtin_ffs(v); }
static inline int __clz(int v) { return v ? __builtin_clz((unsigned)v) : 32; }
static inline float __fmaf_rn(float a, float b, float c) { return std::fmaf(a, b, c); }
static inline float __fadd_rn(float a, float b) { return a + b; }
static inline float __fsub_rn(float a, float b) { return a - b; }
static inline float __fmul_rn(float a, float b) { return a * b; }
static inline float __fdiv_rn(float a, float b) { return a / b; }
static inline float __fsqrt_rn(float a) { return std::sqrt(a); }
static inline int __float2int_rz(float a) { return (int)a; }
static inline float __int2float_rn(int a) { return (float)a; }
static inline float __uint2float_rn(unsigned a) { return (float)a; }

// ---- minimal runtime API used by the host engine --------------------------
typedef int cudaError_t;
typedef void* cudaStream_t;
typedef void* cudaEvent_t;
enum { cudaSuccess = 0 };
enum cudaMemcpyKind { cudaMemcpyHostToDevice, cudaMemcpyDeviceToHost, cudaMemcpyDeviceToDevice, cudaMemcpyDefault };
static inline cudaError_t cudaMalloc(void** p, size_t n) { *p = std::malloc(n ? n : 1); return *p ? 0 : 2; }
static inline cudaError_t cudaFree(void* p) { std::free(p); return 0; }
static inline cudaError_t cudaMallocHost(void** p, size_t n) { *p = std::malloc(n ? n : 1); return *p ? 0 : 2; }
static inline cudaError_t cudaFreeHost(void* p) { std::free(p); return 0; }
static inline cudaError_t cudaMemset(void* p, int v, size_t n) { std::memset(p, v, n); return 0; }
static inline cudaError_t cudaMemsetAsync(void* p, int v, size_t n, cudaStream_t) { std::memset(p, v, n); return 0; }
static inline cudaError_t cudaMemcpy(void* d, const void* s, size_t n, cudaMemcpyKind) { std::memmove(d, s, n); return 0; }
static inline cudaError_t cudaMemcpyPeer(void* d, int, const void* s, int, size_t n) { std::memmove(d, s, n); return 0; }
static inline cudaError_t cudaMemcpyAsync(void* d, const void* s, size_t n, cudaMemcpyKind, cudaStream_t) { std::memmove(d, s, n); return 0; }
static inline cudaError_t cudaStreamCreate(cudaStream_t* s) { *s = nullptr; return 0; }
static inline cudaError_t cudaStreamCreateWithFlags(cudaStream_t* s, unsigned) { *s = nullptr; return 0; }
static inline cudaError_t cudaStreamDestroy(cudaStream_t) { return 0; }
static inline cudaError_t cudaStreamSynchronize(cudaStream_t) { return 0; }
static inline cudaError_t cudaDeviceSynchronize() { return 0; }
static inline cudaError_t cudaSetDevice(int) { return 0; }
static inline cudaError_t cudaGetDevice(int* d) { *d = 0; return 0; }
static inline cudaError_t cudaGetDeviceCount(int* n) { *n = 1; return 0; }
static inline cudaError_t cudaGetLastError() { return 0; }
static inline cudaError_t cudaPeekAtLastError() { return 0; }
static inline const char* cudaGetErrorString(cudaError_t) { return "emu"; }
static inline cudaError_t cudaEventCreate(cudaEvent_t* e) { *e = (void*)1; return 0; }
static inline cudaError_t cudaEventCreateWithFlags(cudaEvent_t* e, unsigned) { *e = (void*)1; return 0; }
static inline cudaError_t cudaStreamWaitEvent(cudaStream_t, cudaEvent_t, unsigned) { return 0; }
#define cudaEventDisableTiming 2
#define cudaDevAttrMultiProcessorCount 16
static inline cudaError_t cudaDeviceGetAttribute(int* v, int, int) { *v = 1; return 0; }
static inline cudaError_t cudaEventDestroy(cudaEvent_t) { return 0; }
static inline cudaError_t cudaEventRecord(cudaEvent_t, cudaStream_t) { return 0; }
static inline cudaError_t cudaEventSynchronize(cudaEvent_t) { return 0; }
static inline cudaError_t cudaEventElapsedTime(float* ms, cudaEvent_t, cudaEvent_t) { *ms = 0.f; return 0; }
#define cudaStreamNonBlocking 1
template <class F>
static inline cudaError_t cudaFuncSetAttribute(F, int, int) { return 0; }
#define cudaFuncAttributeMaxDynamicSharedMemorySize 8
