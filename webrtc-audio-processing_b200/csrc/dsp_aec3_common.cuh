// AEC3 on one warp per call leg: constants of the default EchoCanceller3Config,
// the per-warp shared-memory scratch and small helpers shared by the stages.
//
// Execution pattern (see wap_dev.cuh): vector phases run on all 32 lanes over
// the 64-sample / 65-bin arrays; the scalar control logic of the reference's
// classes (counters, optionals, state machines) runs on lane 0 against the copy
// of Aec3Scalars staged in shared memory, separated from the vector phases by
// __syncwarp().  All per-stream branches are warp-uniform.
#pragma once

#include <stddef.h>

#include "dsp_fft.cuh"
#include "dsp_filters.cuh"
#include "wap_dev.cuh"
#include "wap_ec3_params.h"
#include "wap_state.h"

namespace wap {

// EchoCanceller3Config parameters: wap_ec3_params.h (compile-time constants for the default config,
// a run-time Ec3Params copy in `sc.ep` for the instances that serve other configs).
#if WAP_EC3_RUNTIME
#define WAP_EC3_ARR(name) (sc.ep.name)
#else
WAP_DEVCONST float ec3d_refined[5] = WAP_EC3D_REFINED;
WAP_DEVCONST float ec3d_refined_initial[5] = WAP_EC3D_REFINED_INITIAL;
WAP_DEVCONST float ec3d_coarse[2] = WAP_EC3D_COARSE;
WAP_DEVCONST float ec3d_coarse_initial[2] = WAP_EC3D_COARSE_INITIAL;
WAP_DEVCONST Ec3Tuning ec3d_normal_tuning = WAP_EC3D_NORMAL_TUNING;
WAP_DEVCONST Ec3Tuning ec3d_nearend_tuning = WAP_EC3D_NEAREND_TUNING;
#define WAP_EC3_ARR(name) (::wap::ec3d_##name)
#endif

constexpr int kNumBlocksPerSecond = 250;
constexpr int kMaxRingDelay = kRingBlocks - 1 - kMaxPartitions;  // RenderDelayBufferImpl::MaxDelay(): 153

// RenderDelayBuffer::BufferingEvent
enum { kEventNone = 0, kEventRenderUnderrun = 1, kEventRenderOverrun = 2 };
// EchoPathVariability::DelayAdjustment
enum { kDelayAdjNone = 0, kDelayAdjBufferFlush = 1, kDelayAdjNewDetectedDelay = 2 };
// DelayEstimate::Quality
enum { kQualityCoarse = 0, kQualityRefined = 1 };

struct EchoPathVariability {  // echo_path_variability.h:16-29
  int gain_change, delay_change, clock_drift;
};

// Shared-memory scratch of one warp for the AEC3 stages.  The matched-filter
// buffers (k_delay) and the echo-remover vectors (k_echo) are never live at the
// same time; k_delay only allocates up to the end of `mf`.
constexpr int kMfWin = kMfLen + kSubBlock - 1;   // 527 low-rate samples one filter sees in a block
constexpr int kMfWinPad = 544;                   // second window starts 16 banks after the first
constexpr int kMfShiftCopy = 528;                // one shifted window copy of the accumulated-error path
static_assert(4 * kMfShiftCopy >= 2 * kMfWinPad + 32, "pair path windows must fit");
struct AecMfScratch {
  // Linearised low-rate windows: xp[w] = low_rate[(read + n*384 + w) % size], w < 527, so that
  // tap t of capture sample i is xp[15 - i + t].  The pair path keeps two filters' windows
  // (the second at +560 floats: other half of the banks); the one-filter paths use the first.
  // The accumulated-error path keeps FOUR copies of its window, copy sh shifted by sh samples
  // (xs(sh)[w] = window[w + sh], 528 floats each), so that the 4 consecutive taps a lane owns
  // can be fetched with one aligned 128-bit load whatever the sample's offset is.
  // The two paths never overlap in time.
  alignas(16) float xp[4 * kMfShiftCopy];
  float inst_err[kAccErrLen];  // MatchedFilter::instantaneous_accumulated_error_
  alignas(16) float q[kAccErrLen];     // per-4-tap partial sums / prefix sums of the accumulated-error core
  float x2chain[2][32];    // pair path: the 31 distinct x*x chains of a block, per filter
  float x2sum[2][kSubBlock];   // pair path: x2_sum of every capture sample, per filter
  float err_sum[kNumMatchedFilters];
  int updated[kNumMatchedFilters];
  int peak[kNumMatchedFilters];
};
struct AecRemoverScratch {
  float e_ref[kBlock], e_coa[kBlock], e[kBlock];
  float v0[kBinsPad], v1[kBinsPad], v2[kBinsPad], v3[kBinsPad];
  float x_aligned[kBlock];                   // render block at -MinDirectPathFilterDelay
  // The linear stage (Subtractor::Process) and the stages after it never need their
  // vectors at the same time; only e_ref / e_coa and the scalar metrics cross over.
  union {
    struct {  // Subtractor::Process
      float s_ref[kBlock], s_coa[kBlock];
      float Er_re[kBinsPad], Er_im[kBinsPad], Ec_re[kBinsPad], Ec_im[kBinsPad];
      float E2_ref[kBinsPad], E2_coa[kBinsPad];
      float X2_ref[kBinsPad], X2_coa[kBinsPad];
      float G_re[kBinsPad], G_im[kBinsPad];      // refined update gain
    };
    struct {  // AecState / CNG / residual echo / suppression gain / suppression filter
      float Y_re[kBinsPad], Y_im[kBinsPad], E_re[kBinsPad], E_im[kBinsPad];
      float Y2[kBinsPad], E2[kBinsPad], S2_lin[kBinsPad], R2[kBinsPad], R2_unb[kBinsPad];
      float N_re[kBinsPad], N_im[kBinsPad];      // comfort noise
      float gain[kBinsPad];
    };
  };
};
struct AecEchoScratch {     // k_echo only
  float fftA[128];         // lanes 0-15
  float fftB[128];         // lanes 16-31
  float x[kBlock];         // render block being inserted / GetBlock(0)
  float y[kBlock];         // capture block (in / out)
  AecRemoverScratch rm;
};
struct AecScratch {
  Aec3Scalars s;           // staged copy of Aec3State::s
  float ds[kSubBlock];     // decimated capture sub-block
  float red[32];           // reduction / broadcast exchange
  int ired[32];
#if WAP_EC3_RUNTIME
  Ec3Params ep;            // this engine's EchoCanceller3Config parameters (copied from TickArgs per tick)
#endif
  union {
    AecMfScratch mf;       // k_delay
    struct {               // k_echo (same members as AecEchoScratch)
      float fftA[128];
      float fftB[128];
      float x[kBlock];
      float y[kBlock];
      AecRemoverScratch rm;
    };
  };
};
static_assert(offsetof(AecScratch, fftA) % 8 == 0 && offsetof(AecScratch, fftB) % 8 == 0,
              "the transforms access (re, im) pairs with 64-bit loads / stores");
// Each kernel allocates the common head plus its own member of the union.
constexpr size_t kAecScratchHead = offsetof(AecScratch, mf);
constexpr size_t kAecDelayScratchBytes = kAecScratchHead + sizeof(AecMfScratch);
constexpr size_t kAecEchoScratchBytes = kAecScratchHead + sizeof(AecEchoScratch);

WAP_DEV int ring_inc(int i, int size) { return i < size - 1 ? i + 1 : 0; }
WAP_DEV int ring_dec(int i, int size) { return i > 0 ? i - 1 : size - 1; }
WAP_DEV int ring_off(int i, int off, int size) { return (size + i + off) % size; }

// FftData::CopyFromPackedArray / CopyToPackedArray (fft_data.h:77-98)
WAP_DEV void packed_to_reim(const float* a, float* re, float* im) {
  #pragma unroll
  for (int k = lane_id(); k < kBins; k += 32) {
    if (k == 0) { re[0] = a[0]; im[0] = 0.f; }
    else if (k == 64) { re[64] = a[1]; im[64] = 0.f; }
    else { re[k] = a[2 * k]; im[k] = a[2 * k + 1]; }
  }
}
WAP_DEV void reim_to_packed(const float* re, const float* im, float* a) {
  #pragma unroll
  for (int k = lane_id(); k < kBins; k += 32) {
    if (k == 0) a[0] = re[0];
    else if (k == 64) a[1] = re[64];
    else { a[2 * k] = re[k]; a[2 * k + 1] = im[k]; }
  }
}
// FftData::SpectrumAVX2 (fft_data_avx2.cc:21-33): fused for bins 0..63, plain for bin 64.
WAP_DEV float power_bin(float re, float im, int k) {
  return (k < 64) ? fmaf(re, re, im * im) : re * re + im * im;
}
WAP_DEV void power_spectrum(const float* re, const float* im, float* out) {
  #pragma unroll
  for (int k = lane_id(); k < kBins; k += 32) out[k] = power_bin(re[k], im[k], k);
}

// Two 128-point transforms side by side: lanes 0-15 on fftA, lanes 16-31 on fftB.
// The transforms are leaf routines of ~650 instructions called from a dozen places per block: one
// shared copy of each direction keeps the kernel's instruction footprint small.
WAP_DEV_NOINLINE void fft_pair_forward(float* fftA, float* fftB, bool second_on) {
  const int lane = lane_id();
  float* a = (lane < 16) ? fftA : fftB;
  __syncwarp();
  fft128_forward(a, lane & 15, (lane < 16) || second_on);
}
WAP_DEV_NOINLINE void fft_pair_inverse(float* fftA, float* fftB, bool second_on) {
  const int lane = lane_id();
  float* a = (lane < 16) ? fftA : fftB;
  __syncwarp();
  fft128_inverse(a, lane & 15, (lane < 16) || second_on);
}
WAP_DEV void fft_pair(AecScratch& sc, bool inverse, bool second_on) {
  if (inverse) fft_pair_inverse(sc.fftA, sc.fftB, second_on);
  else fft_pair_forward(sc.fftA, sc.fftB, second_on);
}

// FastApproxLog2f (aec3_common.cc:37-52)
WAP_DEV float fast_approx_log2f(float in) {
  float out = __uint2float_rn(__float_as_uint(in));
  out *= 1.1920929e-7f;
  out -= 126.942695f;
  return out;
}

// std::inner_product(x, x + n, x, 0.f): every lane evaluates the same chain.
WAP_DEV float energy_serial(const float* p, int n) { return serial_sum_sq(p, n); }

// First index of the maximum of an int array (std::max_element), warp-wide.
WAP_DEV int warp_argmax_first_int(const int* p, int n) {
  const int lane = lane_id();
  int best = -2147483647 - 1, bi = 0x7fffffff;
  for (int i = lane; i < n; i += 32) {
    const int v = p[i];
    if (v > best) { best = v; bi = i; }
  }
  for (int m = 16; m; m >>= 1) {
    const int ov = __shfl_xor_sync(WAP_FULL, best, m);
    const int oi = __shfl_xor_sync(WAP_FULL, bi, m);
    if (ov > best || (ov == best && oi < bi)) { best = ov; bi = oi; }
  }
  return bi;
}

}  // namespace wap
