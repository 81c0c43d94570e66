// AEC3 delay estimation for one call leg: capture decimator, the five matched
// filters (NLMS cross-correlators on the decimated render ring), the lag
// aggregators, clock-drift detector and RenderDelayController.
//
// Arithmetic follows the reference's AVX2 path, which is part of the parity
// contract:
//   MatchedFilterCore_AVX2 / _AccumulatedError_AVX2  aec3/matched_filter_avx2.cc:45-270
//   MatchedFilter::Update                            aec3/matched_filter.cc:657-778
//   MatchedFilterLagAggregator                       aec3/matched_filter_lag_aggregator.cc:73-189
//   EchoPathDelayEstimator::EstimateDelay            aec3/echo_path_delay_estimator.cc:66-131
//   RenderDelayControllerImpl::GetDelay              aec3/render_delay_controller.cc:113-168
//
// Lane mapping of one dot product: the AVX2 core keeps 16 h*x accumulators and
// 16 x*x accumulators, each a chain of fused multiply-adds over every 16th tap.
// Lanes 0-15 own the h*x chains, lanes 16-31 the x*x chains, so a warp
// evaluates exactly the reference's 32 chains and then its fixed combine tree.
#pragma once

#include "dsp_aec3_common.cuh"
#include "dsp_aec3_render.cuh"

namespace wap {

// x2_sum_threshold = filters_[0].size() * excitation_limit^2 (matched_filter.cc:667-668)
#define kMfX2SumThreshold (512.f * WAP_EC3(poor_excitation_render_limit) * WAP_EC3(poor_excitation_render_limit))

// MatchedFilter::Update(use_slow_smoothing = MatchedFilterLagAggregator::ReliableDelayFound()):
// delay_estimate_smoothing until a significant candidate was found, then ..._delay_found
// (echo_path_delay_estimator.cc:49-50,85-86; matched_filter.cc:662-663).
WAP_DEV float mf_smoothing(const AecScratch& sc) {
#if WAP_EC3_RUNTIME
  return sc.s.agg_significant_candidate_found ? sc.ep.delay_estimate_smoothing_delay_found : sc.ep.delay_estimate_smoothing;
#else
  static_assert(ec3d::delay_estimate_smoothing == ec3d::delay_estimate_smoothing_delay_found, "default config: one smoothing");
  return ec3d::delay_estimate_smoothing;
#endif
}

// hsum over the 8 "c" values of one accumulator group in the order of hsum_ab
// (matched_filter_avx2.cc:35-43): ((c0+c1)+(c2+c3)) + ((c4+c5)+(c6+c7)),
// c_j = chain_j + chain_{j+8}.  Works on both half-warps at once.
WAP_DEV float mf_hsum16(float acc) {
  acc += __shfl_xor_sync(WAP_FULL, acc, 8);
  acc += __shfl_xor_sync(WAP_FULL, acc, 1);
  acc += __shfl_xor_sync(WAP_FULL, acc, 2);
  acc += __shfl_xor_sync(WAP_FULL, acc, 4);
  return acc;
}

// h[k], h[k+1] += alpha * x[k], x[k+1] as ONE packed fused multiply-add (sm_100 FFMA2): the
// reference's NLMS update is an explicit _mm256_fmadd_ps (matched_filter_avx2.cc:251), and each
// half of fma.rn.f32x2 is the same correctly rounded single-precision FMA.
WAP_DEV void mf_fma2(float& h0, float& h1, float x0, float x1, float alpha) {
#if defined(WAP_EMU)
  h0 = fmaf(x0, alpha, h0);
  h1 = fmaf(x1, alpha, h1);
#else
  unsigned long long hh, xx, aa;
  asm("mov.b64 %0, {%1, %2};" : "=l"(hh) : "f"(h0), "f"(h1));
  asm("mov.b64 %0, {%1, %2};" : "=l"(xx) : "f"(x0), "f"(x1));
  asm("mov.b64 %0, {%1, %1};" : "=l"(aa) : "f"(alpha));
  asm("fma.rn.ftz.f32x2 %0, %1, %2, %0;" : "+l"(hh) : "l"(xx), "l"(aa));
  asm("mov.b64 {%0, %1}, %2;" : "=f"(h0), "=f"(h1) : "l"(hh));
#endif
}

// The filter that won the previous block, with the accumulated-error side output
// (MatchedFilterCore_AccumulatedError_AVX2, matched_filter_avx2.cc:45-149).  The
// reference linearises the window first, so there is no ring-wrap chunking here:
//  * h*x: plain products, summed four taps at a time as (p0+p1)+(p2+p3), then a STRICTLY
//    serial running sum over the 128 groups (s_acum) whose every prefix feeds the
//    instantaneous accumulated error.  Lane l owns taps 4l+b+128m (b,m < 4), i.e. groups
//    l+32m, with h, the x values and the error accumulators in registers; the prefix
//    sum itself is one dependent chain, run by lane 0 on the group sums in shared memory;
//  * x*x: 16 fused chains combined as ((d0+d1)+d2)+d3, d_m = c_m + c_{m+4},
//    c_j = chain_j + chain_{j+8}; the chains come from the per-block table (see mf_pair_fast);
//  * update: fused over all taps.
// Results go to sc.mf.{err_sum,updated,peak}[n] and sc.mf.inst_err.
WAP_DEV void mf_acc_filter(Aec3State& a, AecScratch& sc, int n, const float* y) {
  const int lane = lane_id();
  __syncwarp();
  // ---- stage the window, four shifted copies
  {
    int start = sc.s.lr_read + n * kMfShift;
    if (start >= kLowRateSize) start -= kLowRateSize;
    #pragma unroll
    for (int w = lane; w < kMfWin; w += 32) {
      int r = start + w;
      if (r >= kLowRateSize) r -= kLowRateSize;
      const float v = a.low_rate[r];
#pragma unroll
      for (int sh = 0; sh < 4; ++sh)
        if (w >= sh) sc.mf.xp[sh * kMfShiftCopy + w - sh] = v;
    }
  }
  float h[16];
#pragma unroll
  for (int m = 0; m < 4; ++m) {
    const float4 v = *reinterpret_cast<const float4*>(&a.mf_h[n][4 * lane + 128 * m]);
    h[4 * m] = v.x; h[4 * m + 1] = v.y; h[4 * m + 2] = v.z; h[4 * m + 3] = v.w;
  }
  __syncwarp();
  // ---- x*x chains of the block (copy 0 is the unshifted window)
  {
    const int hw = lane >> 4, L = lane & 15;
    const float* x = hw ? sc.mf.xp + (kSubBlock - 1) - L : sc.mf.xp + (kSubBlock - 1) + L;
    float c = 0.f;
#pragma unroll
    for (int k = 0; k < 32; ++k) {
      const float u = x[16 * k];
      c = fmaf(u, u, c);
    }
    if (hw == 0) sc.mf.x2chain[0][15 - L] = c;       // chain L of sample 0
    else if (L > 0) sc.mf.x2chain[0][15 + L] = c;    // chain 0 of sample L
  }
  __syncwarp();
  if (lane < kSubBlock) {
    const float* E = sc.mf.x2chain[0] + 15 + lane;
    const float d0 = (E[0] + E[-8]) + (E[-4] + E[-12]), d1 = (E[-1] + E[-9]) + (E[-5] + E[-13]);
    const float d2 = (E[-2] + E[-10]) + (E[-6] + E[-14]), d3 = (E[-3] + E[-11]) + (E[-7] + E[-15]);
    sc.mf.x2sum[0][lane] = ((d0 + d1) + d2) + d3;
  }
  __syncwarp();
  float inst[4] = {0.f, 0.f, 0.f, 0.f};
  float error_sum = 0.f;
  int updated = 0;
#pragma unroll 1
  for (int i = 0; i < kSubBlock; ++i) {
    const int o = kSubBlock - 1 - i;
    const float* x = sc.mf.xp + (o & 3) * kMfShiftCopy + (o & ~3) + 4 * lane;
    float xv[16];
#pragma unroll
    for (int m = 0; m < 4; ++m) {
      const float4 v = *reinterpret_cast<const float4*>(x + 128 * m);
      xv[4 * m] = v.x; xv[4 * m + 1] = v.y; xv[4 * m + 2] = v.z; xv[4 * m + 3] = v.w;
      const float p0 = h[4 * m] * v.x, p1 = h[4 * m + 1] * v.y, p2 = h[4 * m + 2] * v.z, p3 = h[4 * m + 3] * v.w;
      sc.mf.q[lane + 32 * m] = (p0 + p1) + (p2 + p3);
    }
    __syncwarp();
    if (lane == 0) {  // the one dependent chain of the block: 128 additions in order
      float s_acum = 0.f;
#pragma unroll
      for (int g = 0; g < kAccErrLen; g += 4) {
        float4 v = *reinterpret_cast<const float4*>(&sc.mf.q[g]);
        s_acum += v.x; v.x = s_acum;
        s_acum += v.y; v.y = s_acum;
        s_acum += v.z; v.z = s_acum;
        s_acum += v.w; v.w = s_acum;
        *reinterpret_cast<float4*>(&sc.mf.q[g]) = v;
      }
    }
    __syncwarp();
    const float yi = y[i];
#pragma unroll
    for (int m = 0; m < 4; ++m) {
      const float eg = sc.mf.q[lane + 32 * m] - yi;
      inst[m] = fmaf(eg, eg, inst[m]);
    }
    const float s_acum = sc.mf.q[kAccErrLen - 1];
    const float x2_sum = sc.mf.x2sum[0][i];
    const float e = yi - s_acum;
    const bool saturation = yi >= 32000.f || yi <= -32000.f;
    error_sum += e * e;
    __syncwarp();
    if (x2_sum > kMfX2SumThreshold && !saturation) {
      const float alpha = mf_smoothing(sc) * e / x2_sum;
#pragma unroll
      for (int j = 0; j < 16; j += 2) mf_fma2(h[j], h[j + 1], xv[j], xv[j + 1], alpha);
      updated = 1;
    }
  }
  // ---- write back; MaxSquarePeakIndex (first maximum among even taps, among odd taps)
  float best_e = -1.f, best_o = -1.f;
  int bi_e = 0, bi_o = 0;
#pragma unroll
  for (int m = 0; m < 4; ++m) {
    *reinterpret_cast<float4*>(&a.mf_h[n][4 * lane + 128 * m]) = make_float4(h[4 * m], h[4 * m + 1], h[4 * m + 2], h[4 * m + 3]);
    sc.mf.inst_err[lane + 32 * m] = inst[m];
#pragma unroll
    for (int b = 0; b < 4; ++b) {
      const float v = h[4 * m + b] * h[4 * m + b];
      const int t = 4 * lane + 128 * m + b;
      if (b & 1) { if (v > best_o || (v == best_o && t < bi_o)) { best_o = v; bi_o = t; } }
      else { if (v > best_e || (v == best_e && t < bi_e)) { best_e = v; bi_e = t; } }
    }
  }
  for (int msk = 16; msk; msk >>= 1) {
    const float ove = __shfl_xor_sync(WAP_FULL, best_e, msk), ovo = __shfl_xor_sync(WAP_FULL, best_o, msk);
    const int oie = __shfl_xor_sync(WAP_FULL, bi_e, msk), oio = __shfl_xor_sync(WAP_FULL, bi_o, msk);
    if (ove > best_e || (ove == best_e && oie < bi_e)) { best_e = ove; bi_e = oie; }
    if (ovo > best_o || (ovo == best_o && oio < bi_o)) { best_o = ovo; bi_o = oio; }
  }
  if (lane == 0) {
    sc.mf.err_sum[n] = error_sum;
    sc.mf.updated[n] = updated;
    sc.mf.peak[n] = (best_o > best_e) ? bi_o : bi_e;
  }
  __syncwarp();
}

// ---- window table of the pair path -------------------------------------------------------
// One filter sees 527 consecutive low-rate samples w[0..526] during a block (tap t of capture
// sample i is w[15 - i + t]).  Lane L of a half-warp owns taps t = L + 16k, k = 0..31, so for
// sample i it needs w[m + 16k] with m = 15 - i + L in 0..30.  The table stores the window
// "transposed": T[g][m][j] = w[m + 16 (4g + j)], g < 8, m < 31, j < 4 -- a lane fetches its 32 taps of
// one sample with 8 aligned 128-bit loads (consecutive lanes read consecutive 16-byte chunks, so
// the loads are bank-conflict free) instead of 32 scalar ones.  Most samples are stored twice.
constexpr int kMfTabGroup = 31 * 4;          // floats per k-group
constexpr int kMfTab = 8 * kMfTabGroup;      // 992 floats per filter
constexpr int kMfTabStride = 1024;           // second filter's table
static_assert(kMfTabStride + kMfTab <= 4 * kMfShiftCopy, "pair-path tables must fit the matched-filter scratch");
WAP_DEV int mf_tab_index(int m, int k) { return ((k >> 2) * 31 + m) * 4 + (k & 3); }

WAP_DEV void mf_stage_table(const Aec3State& a, const AecScratch& sc, int n, float* T) {
  int start = sc.s.lr_read + n * kMfShift;
  if (start >= kLowRateSize) start -= kLowRateSize;
#pragma unroll
  for (int w = lane_id(); w < kMfWin; w += 32) {  // 17 independent loads in flight
    int r = start + w;
    if (r >= kLowRateSize) r -= kLowRateSize;
    const float v = a.low_rate[r];
    const int m = w & 15, k = w >> 4;
    if (k < 32) T[mf_tab_index(m, k)] = v;
    if (m < 15 && k >= 1) T[mf_tab_index(m + 16, k - 1)] = v;
  }
}

// The 32 taps lane L sees for the sample whose table row is m (= 15 - i + L).
WAP_DEV void mf_load_row(const float* T, int m, float (&xv)[32]) {
  const float4* row = reinterpret_cast<const float4*>(T + 4 * m);
#pragma unroll
  for (int g = 0; g < 8; ++g) {
    const float4 v = row[g * 31];
    xv[4 * g] = v.x; xv[4 * g + 1] = v.y; xv[4 * g + 2] = v.z; xv[4 * g + 3] = v.w;
  }
}

// a == b, opaque to the optimiser.  Inside `if (k == v1w)` with k unrolled the compiler would
// otherwise rewrite h[k] / xv[k] as h[v1w] / xv[v1w] -- a dynamically indexed register array, i.e.
// both arrays would move to local memory (measured: 3x the kernel's global traffic in spills).
WAP_DEV bool mf_opaque_eq(int a, int b) {
#if defined(WAP_EMU)
  return a == b;
#else
  int r;
  asm("{\n\t.reg .pred p;\n\tsetp.eq.s32 p, %1, %2;\n\tselp.s32 %0, 1, 0, p;\n\t}" : "=r"(r) : "r"(a), "r"(b));
  return r != 0;
#endif
}

// Two matched filters side by side, one per half-warp: the non-accumulating core of
// matched_filter_avx2.cc:151-270 with
//  * the 32 taps of chain L (t = L + 16k) and their h in REGISTERS of lane L, so the
//    NLMS update reuses the x values the dot product loaded and h never touches
//    shared memory;
//  * kWrap == false (neither window crosses the end of the reference's ring during the
//    block): the x*x chains computed once per block instead of once per sample: chain j of
//    sample i runs over x[s - i + j + 16k], k = 0..31, i.e. it is chain 0 of "sample
//    i - j".  The block therefore has only 31 distinct chains (E[m], m = -15..15, each
//    the same fused 32-term recursion the reference evaluates), and x2_sum of every sample is
//    the reference's hsum tree over E[i - j];
//  * kWrap == true (filter nA's window crosses the ring end; nB's does not): for a sample whose
//    512 taps split into chunk1 | chunk2 at a position that is not a multiple of 16 the reference
//    runs its 16 fused chains over the first 16*floor(chunk1/16) taps, adds the r1 = chunk1 % 16
//    left-over taps unfused to the scalar sums, CONTINUES the same 16 accumulators over chunk 2
//    (chain = tap offset within the chunk mod 16) and adds the last r2 = 16 - r1 taps unfused.
//    With v1 = floor(chunk1/16): lane L's tap L + 16k belongs to chain L for k < v1, is a
//    chunk-1 left-over for k == v1 && L < r1, belongs to chain (L - r1) mod 16 for the taps
//    after that, and is a chunk-2 left-over for k == 31 && L >= r1.  So the accumulators are
//    rotated by r1 lanes once (one shuffle) at k == v1, lanes skip the one FMA their left-over
//    tap would be, and the 16 left-over products are added in tap order -- which is lane
//    order.  The NLMS update is fused except for the < 8 taps at the end of each chunk
//    (k == v1 and k == 31 again).  x*x runs per sample through the same structure.
// nA / nB: filter indices (nB < 0: second half idle).
template <bool kWrap>
WAP_DEV void mf_pair(Aec3State& a, AecScratch& sc, int nA, int nB, const float* y) {
  const int lane = lane_id();
  const int hw = lane >> 4, L = lane & 15;
  const int n = hw ? nB : nA;
  const bool on = n >= 0;
  const int nn = on ? n : nA;
  const float* T = sc.mf.xp + hw * kMfTabStride;
  __syncwarp();
  mf_stage_table(a, sc, nA, sc.mf.xp);
  if (nB >= 0) mf_stage_table(a, sc, nB, sc.mf.xp + kMfTabStride);
  float h[32];
#pragma unroll
  for (int k = 0; k < 32; ++k) h[k] = a.mf_h[nn][L + 16 * k];
  __syncwarp();
  if (!kWrap) {
    // ---- the 31 x*x chains of this block: E[15 - L] = chain L of sample 0, E[15 + L] = chain 0 of sample L
    {
      float u[32], v[32];
      mf_load_row(T, 15 + L, u);   // tap L of sample 0 onwards
      mf_load_row(T, 15 - L, v);   // tap 0 of sample L onwards
      float c_init = 0.f, c_new = 0.f;
#pragma unroll
      for (int k = 0; k < 32; ++k) {
        c_init = fmaf(u[k], u[k], c_init);
        c_new = fmaf(v[k], v[k], c_new);
      }
      sc.mf.x2chain[hw][15 - L] = c_init;
      if (L > 0) sc.mf.x2chain[hw][15 + L] = c_new;
    }
    __syncwarp();
    {
      // x2_sum of sample i = L: c_j = chain_j + chain_{j+8}; ((c0+c1)+(c2+c3)) + ((c4+c5)+(c6+c7))  (hsum_ab)
      const float* E = sc.mf.x2chain[hw] + 15 + L;  // E[-j] = chain j of sample L
      const float c0 = E[0] + E[-8], c1 = E[-1] + E[-9], c2 = E[-2] + E[-10], c3 = E[-3] + E[-11];
      const float c4 = E[-4] + E[-12], c5 = E[-5] + E[-13], c6 = E[-6] + E[-14], c7 = E[-7] + E[-15];
      sc.mf.x2sum[hw][L] = ((c0 + c1) + (c2 + c3)) + ((c4 + c5) + (c6 + c7));
    }
    __syncwarp();
  }
  // position of tap 0 of sample 0 of filter nA in the reference's ring (kWrap only)
  const int x_start0 = sc.s.lr_read + nA * kMfShift + kSubBlock - 1;
  float error_sum = 0.f;
  int updated = 0;
#pragma unroll 1
  for (int i = 0; i < kSubBlock; ++i) {
    float xv[32];
    mf_load_row(T, kSubBlock - 1 - i + L, xv);
    float acc = 0.f, accx = 0.f;
    float s, x2_sum;
    // kWrap: where filter nA wraps for this sample (warp-uniform; r1 == 0: the chains are unaffected)
    int v1w = 99, r1 = 0, chunk1 = kMfLen;
    if (kWrap) {
      int xs = x_start0 - i;
      if (xs >= kLowRateSize) xs -= kLowRateSize;
      chunk1 = imin(kMfLen, kLowRateSize - xs);
      r1 = chunk1 & 15;
      v1w = r1 ? (chunk1 >> 4) : 99;
    }
    const bool wrapped_half = kWrap && hw == 0;
    const bool tail1 = wrapped_half && L < r1;             // my tap at k == v1w is a chunk-1 left-over
    const bool tail2 = wrapped_half && r1 > 0 && L >= r1;  // my tap at k == 31 is a chunk-2 left-over
    float p_h = 0.f, p_x = 0.f;
#pragma unroll
    for (int g = 0; g < 4; ++g) {
      if (kWrap && mf_opaque_eq(v1w >> 3, g)) {   // warp-uniform: the wrap step is among these eight k
#pragma unroll
        for (int kk = 0; kk < 8; ++kk) {
          const int k = 8 * g + kk;
          const bool at_wrap = mf_opaque_eq(k, v1w);
          if (at_wrap) {
            // the accumulators move to the lane that owns their chain's taps from here on (one
            // convergent shuffle: the half-warp whose filter does not wrap reads itself)
            const int src = wrapped_half ? ((L - r1) & 15) : lane;
            acc = __shfl_sync(WAP_FULL, acc, src);
            accx = __shfl_sync(WAP_FULL, accx, src);
          }
          if (at_wrap || k == 31) {   // warp-uniform
            const bool left_over = (at_wrap && tail1) || (k == 31 && tail2);
            const float ph = h[k] * xv[k], px = xv[k] * xv[k];
            const float a1 = fmaf(h[k], xv[k], acc), a2 = fmaf(xv[k], xv[k], accx);
            acc = left_over ? acc : a1;
            accx = left_over ? accx : a2;
            p_h = left_over ? ph : p_h;
            p_x = left_over ? px : p_x;
          } else {
            acc = fmaf(h[k], xv[k], acc);
            accx = fmaf(xv[k], xv[k], accx);
          }
        }
      } else {
#pragma unroll
        for (int kk = 0; kk < 8; ++kk) {
          const int k = 8 * g + kk;
          if (kWrap && k == 31) {
            const float a1 = fmaf(h[k], xv[k], acc), a2 = fmaf(xv[k], xv[k], accx);
            p_h = tail2 ? h[k] * xv[k] : p_h;
            p_x = tail2 ? xv[k] * xv[k] : p_x;
            acc = tail2 ? acc : a1;
            accx = tail2 ? accx : a2;
          } else {
            acc = fmaf(h[k], xv[k], acc);
            if (kWrap) accx = fmaf(xv[k], xv[k], accx);
          }
        }
      }
    }
    if (kWrap) {
      float t_h = 0.f, t_x = 0.f;
      if (r1 > 0) {   // warp-uniform
        // back to chain order for the combine tree
        const int src = wrapped_half ? ((L + r1) & 15) : lane;
        acc = __shfl_sync(WAP_FULL, acc, src);
        accx = __shfl_sync(WAP_FULL, accx, src);
        // the 16 left-over products in tap (= lane) order: lane 0 adds the h*x ones, lane 1 the x*x ones
        __syncwarp();
        if (wrapped_half) { sc.mf.q[L] = p_h; sc.mf.q[16 + L] = p_x; }
        __syncwarp();
        const float4* q = reinterpret_cast<const float4*>(sc.mf.q + 16 * (lane & 1));
        float t = 0.f;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const float4 v = q[j];
          t += v.x; t += v.y; t += v.z; t += v.w;
        }
        t_h = __shfl_sync(WAP_FULL, t, 0);
        t_x = __shfl_sync(WAP_FULL, t, 1);
        if (!wrapped_half) { t_h = 0.f; t_x = 0.f; }
      }
      s = t_h + mf_hsum16(acc);        // reference: s (left-overs) += hsum(vector accumulators)
      x2_sum = t_x + mf_hsum16(accx);
    } else {
      s = mf_hsum16(acc);     // uniform within the half-warp; reference: s = 0 + hsum
      x2_sum = sc.mf.x2sum[hw][i];
    }
    const float yi = y[i];
    const float e = yi - s;
    const bool saturation = yi >= 32000.f || yi <= -32000.f;
    error_sum += e * e;
    if (on && x2_sum > kMfX2SumThreshold && !saturation) {
      const float alpha = mf_smoothing(sc) * e / x2_sum;
      if (kWrap) {
        // groups of 8 of each chunk are fused; the < 8 taps at the end of each chunk are not
        const int c1 = wrapped_half ? chunk1 : kMfLen;
        const int f1 = c1 & ~7, f2 = c1 + ((kMfLen - c1) & ~7);
#pragma unroll
        for (int g = 0; g < 4; ++g) {
          if (mf_opaque_eq(v1w >> 3, g) || g == 3) {
#pragma unroll
            for (int kk = 0; kk < 8; ++kk) {
              const int k = 8 * g + kk;
              if (mf_opaque_eq(k, v1w) || k == 31) {   // warp-uniform: the only steps that can hold unfused taps
                const int t = L + 16 * k;
                const bool unfused = (t >= f1 && t < c1) || t >= f2;
                const float hf = fmaf(xv[k], alpha, h[k]), hu = h[k] + alpha * xv[k];
                h[k] = unfused ? hu : hf;
              } else {
                h[k] = fmaf(xv[k], alpha, h[k]);
              }
            }
          } else {
#pragma unroll
            for (int kk = 0; kk < 8; kk += 2) mf_fma2(h[8 * g + kk], h[8 * g + kk + 1], xv[8 * g + kk], xv[8 * g + kk + 1], alpha);
          }
        }
      } else {
#pragma unroll
        for (int k = 0; k < 32; k += 2) mf_fma2(h[k], h[k + 1], xv[k], xv[k + 1], alpha);
      }
      updated = 1;
    }
  }
  // ---- write back, MaxSquarePeakIndex: tap parity == lane parity (t = L + 16k)
  float best = -1.f;
  int bi = 0;
#pragma unroll
  for (int k = 0; k < 32; ++k) {
    if (on) a.mf_h[nn][L + 16 * k] = h[k];
    const float v = h[k] * h[k];
    if (v > best) { best = v; bi = L + 16 * k; }
  }
  for (int m = 2; m < 16; m <<= 1) {
    const float ov = __shfl_xor_sync(WAP_FULL, best, m);
    const int oi = __shfl_xor_sync(WAP_FULL, bi, m);
    if (ov > best || (ov == best && oi < bi)) { best = ov; bi = oi; }
  }
  const float odd_v = __shfl_xor_sync(WAP_FULL, best, 1);
  const int odd_i = __shfl_xor_sync(WAP_FULL, bi, 1);
  if (on && L == 0) {  // lane with the even-tap maximum; odd wins only when strictly larger
    sc.mf.err_sum[n] = error_sum;
    sc.mf.updated[n] = updated;
    sc.mf.peak[n] = (odd_v > best) ? odd_i : bi;
  }
  __syncwarp();
}

// MatchedFilter::Reset (matched_filter.cc:641-655)
WAP_DEV void mf_reset(Aec3State& a, AecScratch& sc, bool full_reset) {
  const int lane = lane_id();
  for (int i = lane; i < kNumMatchedFilters * kMfLen; i += 32) (&a.mf_h[0][0])[i] = 0.f;
  if (full_reset) {
    for (int i = lane; i < kNumMatchedFilters * kAccErrLen; i += 32) (&a.mf_acc_err[0][0])[i] = 1.f;
    if (lane == 0) sc.s.mf_number_pre_echo_updates = 0;
  }
  __syncwarp();
}

// MatchedFilterLagAggregator::Reset (matched_filter_lag_aggregator.cc:62-70)
WAP_DEV void lag_aggregator_reset(Aec3State& a, AecScratch& sc, bool hard_reset) {
  const int lane = lane_id();
  for (int i = lane; i < kLagHistSize; i += 32) a.lag_hist[i] = 0;
  for (int i = lane; i < 250; i += 32) { a.lag_hist_data[i] = 0; a.pre_hist_data[i] = -1; }
  for (int i = lane; i < kPreEchoHistSize; i += 32) a.pre_hist[i] = 0;
  if (lane == 0) {
    sc.s.agg_hist_data_index = 0;
    sc.s.agg_candidate_valid = 0;  // candidate_ itself survives a Reset(); it is recomputed by the next Aggregate
    sc.s.pre_hist_data_index = 0;
    sc.s.pre_candidate = 0;
    if (hard_reset) sc.s.agg_significant_candidate_found = 0;
  }
  __syncwarp();
}

// EchoPathDelayEstimator::Reset(reset_lag_aggregator, reset_delay_confidence) (:123-131)
WAP_DEV void delay_estimator_reset(Aec3State& a, AecScratch& sc, bool reset_lag_aggregator, bool reset_delay_confidence) {
  if (reset_lag_aggregator) lag_aggregator_reset(a, sc, reset_delay_confidence);
  mf_reset(a, sc, reset_lag_aggregator);
  if (lane_id() == 0) {
    sc.s.est_has_old_lag = 0;
    sc.s.est_consistent_counter = 0;
  }
  __syncwarp();
}

// RenderDelayControllerImpl::Reset (render_delay_controller.cc:103-111)
WAP_DEV void delay_controller_reset(Aec3State& a, AecScratch& sc, bool reset_delay_confidence) {
  if (lane_id() == 0) {
    sc.s.ctl_has_delay = 0;
    sc.s.ctl_has_delay_samples = 0;
    sc.s.ctl_delay_change_counter = 0;
    if (reset_delay_confidence) sc.s.ctl_last_quality = kQualityCoarse;
  }
  delay_estimator_reset(a, sc, true, reset_delay_confidence);
}

// ClockdriftDetector::Update (clockdrift_detector.cc:21-60), lane 0.
WAP_DEV void clockdrift_update(Aec3Scalars& s, int delay_estimate) {
  if (delay_estimate == s.cd_history[0]) {
    if (++s.cd_stability_counter > 7500) s.cd_level = 0;
    return;
  }
  s.cd_stability_counter = 0;
  const int d1 = s.cd_history[0] - delay_estimate;
  const int d2 = s.cd_history[1] - delay_estimate;
  const int d3 = s.cd_history[2] - delay_estimate;
  const bool probable_up = (d1 == -1 && d2 == -2) || (d1 == -2 && d2 == -1);
  const bool drift_up = probable_up && d3 == -3;
  const bool probable_down = (d1 == 1 && d2 == 2) || (d1 == 2 && d2 == 1);
  const bool drift_down = probable_down && d3 == 3;
  if (drift_up || drift_down) s.cd_level = 2;                                   // kVerified
  else if ((probable_up || probable_down) && s.cd_level == 0) s.cd_level = 1;  // kProbable
  s.cd_history[2] = s.cd_history[1];
  s.cd_history[1] = s.cd_history[0];
  s.cd_history[0] = delay_estimate;
}

// RenderDelayControllerImpl::GetDelay for the decimated capture block in sc.ds
// (EchoPathDelayEstimator::EstimateDelay's capture decimation ran in k_front).
// Leaves the controller's delay_ in sc.s.ctl_{has_delay,delay,delay_quality}.
WAP_DEV void aec3_get_delay(Aec3State& a, AecScratch& sc) {
  const int lane = lane_id();
  Aec3Scalars& s = sc.s;
  __syncwarp();
  const float* y = sc.ds;
  // ---- MatchedFilter::Update
  float error_sum_anchor = 0.f;
  for (int k = 0; k < kSubBlock; ++k) error_sum_anchor += y[k] * y[k];
  // delay.detect_pre_echo = false: no filter takes the accumulated-error path (matched_filter.cc:686-687)
  const int last_best = WAP_EC3(detect_pre_echo) ? s.mf_last_detected_best_lag_filter : -1;
  // The filter that won the previous block needs the accumulated-error side output and goes through
  // its own path; the others are processed two at a time, one per half-warp.  A filter whose window
  // crosses the end of the reference's ring during this block is paired with one that does not
  // (adjacent filters can wrap together: they go into different passes).
  unsigned plain = 0, wrapped = 0;   // bit n: filter n (no arrays: the kernel keeps no stack frame)
  for (int n = 0; n < kNumMatchedFilters; ++n) {
    if (n == last_best) continue;
    int x_start0 = s.lr_read + n * kMfShift + kSubBlock - 1;
    if (x_start0 >= kLowRateSize) x_start0 -= kLowRateSize;
    const bool no_wrap = x_start0 >= kSubBlock - 1 && x_start0 + kMfLen <= kLowRateSize;
    if (no_wrap) plain |= 1u << n;
    else wrapped |= 1u << n;
  }
  // One call site per variant inside one loop: each is inlined exactly once (instruction-cache
  // footprint), and nothing is passed through local memory.
#pragma unroll 1
  while (plain | wrapped) {
    const bool wrap = wrapped != 0;
    int nA, nB = -1;
    if (wrap) { nA = __ffs((int)wrapped) - 1; wrapped &= wrapped - 1; }
    else { nA = __ffs((int)plain) - 1; plain &= plain - 1; }
    if (plain) { nB = __ffs((int)plain) - 1; plain &= plain - 1; }
    if (wrap) mf_pair<true>(a, sc, nA, nB, y);
    else mf_pair<false>(a, sc, nA, nB, y);
  }
  if (last_best >= 0) mf_acc_filter(a, sc, last_best, y);
  // winner selection (matched_filter.cc:729-776), lane 0
  if (lane == 0) {
    float winner_error_sum = error_sum_anchor;
    int has_winner_lag = 0, winner_lag = 0, winner_index = -1;
    int has_prev = 0, prev_lag = 0, alignment_shift = 0;
    for (int n = 0; n < kNumMatchedFilters; ++n) {
      const int lag_estimate = sc.mf.peak[n];
      const float error_sum = sc.mf.err_sum[n];
      const bool reliable = lag_estimate > 2 && lag_estimate < (kMfLen - 10) &&
                            error_sum < WAP_EC3(delay_candidate_detection_threshold) * error_sum_anchor;
      const int lag = lag_estimate + alignment_shift;
      if (sc.mf.updated[n] && reliable && error_sum < winner_error_sum) {
        winner_error_sum = error_sum;
        winner_index = n;
        if (has_prev && prev_lag == lag) {
          winner_lag = prev_lag;
          winner_index = n - 1;
        } else {
          winner_lag = lag;
        }
        has_winner_lag = 1;
      }
      has_prev = 1;
      prev_lag = lag;
      alignment_shift += kMfShift;
    }
    sc.ired[0] = winner_index;
    sc.ired[1] = winner_lag;
    sc.ired[2] = winner_lag;  // pre_echo_lag
    sc.ired[3] = 0;           // update accumulated error?
    if (winner_index != -1 && last_best == winner_index) {
      if (error_sum_anchor > 1.0f) {
        sc.ired[3] = 1;
        s.mf_number_pre_echo_updates++;
      }
    }
    (void)has_winner_lag;
  }
  __syncwarp();
  const int winner_index = sc.ired[0];
  const int winner_lag = sc.ired[1];
  if (winner_index != -1) {
    if (sc.ired[3]) {
      // UpdateAccumulatedError (matched_filter.cc:43-58)
      const float one_over_anchor = 1.0f / error_sum_anchor;
      #pragma unroll
      for (int k = lane; k < kAccErrLen; k += 32) {
        const float error_norm = sc.mf.inst_err[k] * one_over_anchor;
        float acc = a.mf_acc_err[winner_index][k];
        if (error_norm < acc) acc = error_norm;
        else acc += 0.015f * (error_norm - acc);
        a.mf_acc_err[winner_index][k] = acc;
      }
      __syncwarp();
    }
    if (lane == 0) {
      if (last_best == winner_index && s.mf_number_pre_echo_updates >= 50) {
        // ComputePreEchoLag (matched_filter.cc:60-76)
        const int shift_winner = winner_index * kMfShift;
        int pre = winner_lag - shift_winner;
        const int maximum_pre_echo_lag = imin(pre / 4, kAccErrLen);
        for (int k = maximum_pre_echo_lag - 1; k >= 0; --k) {
          if (a.mf_acc_err[winner_index][k] > 0.5f) break;
          pre = (k + 1) * 4 - 1;
        }
        sc.ired[2] = pre + shift_winner;
      }
      s.mf_last_detected_best_lag_filter = winner_index;
    }
    __syncwarp();
  }
  const int pre_echo_lag = sc.ired[2];

  // ---- MatchedFilterLagAggregator::Aggregate
  int has_agg = 0, agg_quality = 0, agg_delay = 0;
  if (winner_index != -1) {
    const int headroom = WAP_EC3(delay_headroom_samples) / kDownSampling;
    // PreEchoLagAggregator::Aggregate (:139-183); it only exists with delay.detect_pre_echo (:53-56)
    if (WAP_EC3(detect_pre_echo)) {
      int blk = imax(0, pre_echo_lag - headroom) >> 4;
      blk = imin(imax(blk, 0), kPreEchoHistSize - 1);
      if (lane == 0) {
        const int old = a.pre_hist_data[s.pre_hist_data_index];
        if (old != -1) --a.pre_hist[old];
        a.pre_hist_data[s.pre_hist_data_index] = blk;
        ++a.pre_hist[blk];
        s.pre_hist_data_index = (s.pre_hist_data_index + 1) % 250;
      }
      __syncwarp();
      int cand = 0;
      if (s.pre_number_updates < kNumBlocksPerSecond * 2) {
        float penalization = 1.0f, max_value = -1.0f;
        for (int w = 0; w + kMfWindowSubBlocks <= kPreEchoHistSize; w += kMfWindowSubBlocks) {
          const int v = a.pre_hist[w + lane];
          // first maximum inside the 32-bin window
          int best = v, bi = lane;
          for (int m = 16; m; m >>= 1) {
            const int ov = __shfl_xor_sync(WAP_FULL, best, m);
            const int oi = __shfl_xor_sync(WAP_FULL, bi, m);
            if (ov > best || (ov == best && oi < bi)) { best = ov; bi = oi; }
          }
          const float weighted = (float)best * penalization;
          if (weighted > max_value) {
            max_value = weighted;
            cand = w + bi;
          }
          penalization *= 0.7f;
        }
        __syncwarp();
        if (lane == 0) s.pre_number_updates++;
      } else {
        cand = warp_argmax_first_int(a.pre_hist, kPreEchoHistSize);
      }
      if (lane == 0) s.pre_candidate = cand << 4;
      __syncwarp();
    }
    // HighestPeakAggregator::Aggregate (:115-127).  candidate_ is the FIRST index of the
    // histogram maximum (std::max_element).  One bin loses a count and one gains one per call,
    // so the argmax is maintained incrementally and the 2433-bin scan only runs when the bin
    // that lost a count was the candidate itself (or after a reset).
    {
      const int lag = imax(0, winner_lag - headroom);
      const int idx = s.agg_hist_data_index;
      const int old_lag = a.lag_hist_data[idx];
      const int prev_cand = s.agg_candidate;
      const bool valid = s.agg_candidate_valid != 0;
      __syncwarp();
      if (lane == 0) {
        --a.lag_hist[old_lag];
        a.lag_hist_data[idx] = lag;
        ++a.lag_hist[lag];
        s.agg_hist_data_index = (idx + 1) % 250;
      }
      __syncwarp();
      int cand;
      if (!valid || (old_lag == prev_cand && old_lag != lag)) {
        cand = warp_argmax_first_int(a.lag_hist, kLagHistSize);
      } else if (old_lag == lag) {
        cand = prev_cand;  // histogram unchanged
      } else {
        const int vmax = a.lag_hist[prev_cand], v = a.lag_hist[lag];
        cand = (v > vmax || (v == vmax && lag < prev_cand)) ? lag : prev_cand;
      }
      const int count = a.lag_hist[cand];
      const int sig = s.agg_significant_candidate_found || count > WAP_EC3(thr_converged);
      if (count > WAP_EC3(thr_converged) || (count > WAP_EC3(thr_initial) && !sig)) {
        has_agg = 1;
        agg_quality = sig ? kQualityRefined : kQualityCoarse;
        agg_delay = WAP_EC3(detect_pre_echo) ? s.pre_candidate : cand;   // :98-100
      }
      __syncwarp();
      if (lane == 0) {
        s.agg_candidate = cand;
        s.agg_candidate_valid = 1;
        s.agg_significant_candidate_found = sig;
      }
      __syncwarp();
    }
  }

  // ---- rest of EstimateDelay + RenderDelayControllerImpl::GetDelay, lane 0
  if (lane == 0) {
    if (has_agg && agg_quality == kQualityRefined) clockdrift_update(s, s.agg_candidate);
    if (has_agg) agg_delay *= kDownSampling;
    if (s.est_has_old_lag && has_agg && s.est_old_lag == agg_delay) ++s.est_consistent_counter;
    else s.est_consistent_counter = 0;
    s.est_has_old_lag = has_agg;
    s.est_old_lag = agg_delay;
    sc.ired[4] = s.est_consistent_counter > kNumBlocksPerSecond / 2;
  }
  __syncwarp();
  if (sc.ired[4]) delay_estimator_reset(a, sc, false, false);
  if (lane == 0) {
    if (has_agg) {
      if (!s.ctl_has_delay_samples || agg_delay != s.ctl_delay_samples) s.ctl_delay_change_counter = 0;
      s.ctl_has_delay_samples = 1;
      s.ctl_delay_samples = agg_delay;
      s.ctl_delay_samples_quality = agg_quality;
    }
    if (s.ctl_delay_change_counter < 2 * kNumBlocksPerSecond) ++s.ctl_delay_change_counter;
    if (s.ctl_has_delay_samples) {
      const bool use_hysteresis =
          s.ctl_last_quality == kQualityRefined && s.ctl_delay_samples_quality == kQualityRefined;
      // ComputeBufferDelay (render_delay_controller.cc:65-82)
      const int hysteresis = use_hysteresis ? WAP_EC3(hysteresis_limit_blocks) : 0;
      int new_delay_blocks = s.ctl_delay_samples >> 6;
      if (s.ctl_has_delay) {
        const int current = s.ctl_delay;
        if (new_delay_blocks > current && new_delay_blocks <= current + hysteresis) new_delay_blocks = current;
      }
      s.ctl_has_delay = 1;
      s.ctl_delay = new_delay_blocks;
      s.ctl_delay_quality = s.ctl_delay_samples_quality;
      s.ctl_last_quality = s.ctl_delay_samples_quality;
    }
  }
  __syncwarp();
}

}  // namespace wap
