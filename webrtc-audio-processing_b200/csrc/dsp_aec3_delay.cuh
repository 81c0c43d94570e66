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

constexpr float kMfX2SumThreshold = 512.f * ec3::kMfExcitationLimit * ec3::kMfExcitationLimit;

// Copies the part of the low-rate ring filter n sees during this block into shared
// memory, linearised: dst[w] = low_rate[(read + n*shift + w) % size], w < 527.
WAP_DEV void mf_stage_window(const Aec3State& a, const AecScratch& sc, int n, float* dst) {
  int start = sc.s.lr_read + n * kMfShift;
  if (start >= kLowRateSize) start -= kLowRateSize;
#pragma unroll
  for (int w = lane_id(); w < kMfWin; w += 32) {  // 17 independent loads in flight
    int r = start + w;
    if (r >= kLowRateSize) r -= kLowRateSize;
    dst[w] = a.low_rate[r];
  }
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

// One matched filter, 16 decimated capture samples, non-accumulating core
// (matched_filter_avx2.cc:151-270), general version for blocks in which the window crosses
// the end of the reference's ring: per sample the 512 taps split at the wrap into two
// chunks; in each chunk the first floor(len/16)*16 taps feed the 16 fused chains (lane =
// tap index mod 16 WITHIN the chunk), the remaining len%16 taps are added to the scalar
// sums with separate multiply and add, in order.  Lanes 0-15 run the h*x chains, lanes
// 16-31 the x*x chains -- the same code on a per-lane operand pointer (h or x).
// h lives in (sc.mf.xp + kMfHOffset), the window in sc.mf.xp.
WAP_DEV void mf_core(AecScratch& sc, int n, const float* y, float* error_sum_out, int* updated_out) {
  const int lane = lane_id();
  const int half = lane >> 4, L = lane & 15;
  float* h = (sc.mf.xp + kMfHOffset);
  float error_sum = 0.f;
  int updated = 0;
#pragma unroll 1
  for (int i = 0; i < kSubBlock; ++i) {
    int x_start = sc.s.lr_read + n * kMfShift + kSubBlock - 1 - i;  // position in the reference's ring
    if (x_start >= kLowRateSize) x_start -= kLowRateSize;
    const float* x = sc.mf.xp + (kSubBlock - 1 - i);                // tap 0 of sample i
    const float* pa = half ? x : h;                                  // first operand of this lane's chain
    const int chunk1 = imin(kMfLen, kLowRateSize - x_start);
    const int chunk2 = kMfLen - chunk1;
    const int v1 = chunk1 >> 4, v2 = chunk2 >> 4;
    float acc = 0.f;
    {
      const float* qa = pa + L;
      const float* qx = x + L;
      int k = 0;
      for (; k + 4 <= v1; k += 4, qa += 64, qx += 64) {
        acc = fmaf(qa[0], qx[0], acc);
        acc = fmaf(qa[16], qx[16], acc);
        acc = fmaf(qa[32], qx[32], acc);
        acc = fmaf(qa[48], qx[48], acc);
      }
      for (; k < v1; ++k, qa += 16, qx += 16) acc = fmaf(qa[0], qx[0], acc);
      qa = pa + chunk1 + L;
      qx = x + chunk1 + L;
      for (k = 0; k + 4 <= v2; k += 4, qa += 64, qx += 64) {
        acc = fmaf(qa[0], qx[0], acc);
        acc = fmaf(qa[16], qx[16], acc);
        acc = fmaf(qa[32], qx[32], acc);
        acc = fmaf(qa[48], qx[48], acc);
      }
      for (; k < v2; ++k, qa += 16, qx += 16) acc = fmaf(qa[0], qx[0], acc);
    }
    acc = mf_hsum16(acc);   // uniform within each half: h*x on lanes 0-15, x*x on lanes 16-31
    // Scalar tails: r1 taps after chunk 1's groups, then r2 after chunk 2's (r1 + r2 is 0 or 16).
    // One product per lane, then the additions in tap order through shuffles.
    const int r1 = chunk1 & 15, r2 = chunk2 & 15;
    float tail = 0.f;
    if (r1 + r2) {
      const int t = L < r1 ? 16 * v1 + L : chunk1 + 16 * v2 + (L - r1);
      const float p = pa[t] * x[t];
#pragma unroll
      for (int j = 0; j < 16; ++j) tail += __shfl_sync(WAP_FULL, p, (lane & 16) | j);
    }
    const float mine = tail + acc;                      // reference: s += vec (x2_sum += vec)
    const float s = __shfl_sync(WAP_FULL, mine, 0);
    const float x2_sum = __shfl_sync(WAP_FULL, mine, 16);
    const float yi = y[i];
    const float e = yi - s;
    const bool saturation = yi >= 32000.f || yi <= -32000.f;
    error_sum += e * e;
    __syncwarp();
    if (x2_sum > kMfX2SumThreshold && !saturation) {
      const float alpha = ec3::kMfSmoothing * e / x2_sum;
      // The groups of 8 of each chunk are fused, the (< 8 tap) ends are not: those few taps are
      // computed unfused from the old h first and written over the fused result afterwards.
      const int f1 = chunk1 & ~7, n1 = chunk1 - f1;
      const int f2 = chunk1 + (chunk2 & ~7), n2 = kMfLen - f2;
      int tn = -1;
      float hn = 0.f;
      if (lane < n1) tn = f1 + lane;
      else if (lane - n1 < n2) tn = f2 + (lane - n1);
      if (tn >= 0) hn = h[tn] + alpha * x[tn];
      __syncwarp();
#pragma unroll
      for (int k = 0; k < kMfLen / 32; ++k) {
        const int t = lane + 32 * k;
        h[t] = fmaf(x[t], alpha, h[t]);
      }
      __syncwarp();
      if (tn >= 0) h[tn] = hn;
      updated = 1;
    }
    __syncwarp();
  }
  *error_sum_out = error_sum;
  *updated_out = updated;
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
#pragma unroll 4
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
      const float alpha = ec3::kMfSmoothing * e / x2_sum;
#pragma unroll
      for (int j = 0; j < 16; ++j) h[j] = fmaf(xv[j], alpha, h[j]);
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

// Two matched filters side by side, one per half-warp, for blocks in which neither
// window crosses the end of the reference's ring (chunk1 == 512 for all 16 capture
// samples, no scalar tails): the non-accumulating core of matched_filter_avx2.cc:
// 151-270 with
//  * the 32 taps of chain L (t = L + 16k) and their h in REGISTERS of lane L, so the
//    NLMS update reuses the x values the dot product loaded and h never touches
//    shared memory;
//  * the x*x chains computed once per block instead of once per sample: chain j of
//    sample i runs over x[s - i + j + 16k], k = 0..31, i.e. it is chain 0 of "sample
//    i - j".  The block therefore has only 31 distinct chains (E[m], m = -15..15, each
//    the same fused 32-term recursion the reference evaluates), and x2_sum of sample i
//    is the reference's hsum tree over E[i - j].
// nA / nB: filter indices (nB < 0: second half idle).
WAP_DEV void mf_pair_fast(Aec3State& a, AecScratch& sc, int nA, int nB, const float* y) {
  const int lane = lane_id();
  const int hw = lane >> 4, L = lane & 15;
  const int n = hw ? nB : nA;
  const bool on = n >= 0;
  const int nn = on ? n : nA;
  float* xp = sc.mf.xp + hw * (kMfWinPad + 16);
  __syncwarp();
  mf_stage_window(a, sc, nA, sc.mf.xp);
  if (nB >= 0) mf_stage_window(a, sc, nB, sc.mf.xp + kMfWinPad + 16);
  float h[32];
#pragma unroll
  for (int k = 0; k < 32; ++k) h[k] = a.mf_h[nn][L + 16 * k];
  __syncwarp();
  // ---- the 31 x*x chains of this block: E[15 - L] = chain L of sample 0, E[15 + L] = chain 0 of sample L
  {
    float c_init = 0.f, c_new = 0.f;
    const float* x0 = xp + (kSubBlock - 1) + L;  // tap L of sample 0
    const float* x1 = xp + (kSubBlock - 1) - L;  // tap 0 of sample L
#pragma unroll
    for (int k = 0; k < 32; ++k) {
      const float u = x0[16 * k], v = x1[16 * k];
      c_init = fmaf(u, u, c_init);
      c_new = fmaf(v, v, c_new);
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
  float error_sum = 0.f;
  int updated = 0;
#pragma unroll 1
  for (int i = 0; i < kSubBlock; ++i) {
    const float* x = xp + (kSubBlock - 1 - i) + L;
    float xv[32];
    float acc = 0.f;
#pragma unroll
    for (int k = 0; k < 32; ++k) {
      xv[k] = x[16 * k];
      acc = fmaf(h[k], xv[k], acc);
    }
    const float s = mf_hsum16(acc);     // uniform within the half-warp; reference: s = 0 + hsum
    const float x2_sum = sc.mf.x2sum[hw][i];
    const float yi = y[i];
    const float e = yi - s;
    const bool saturation = yi >= 32000.f || yi <= -32000.f;
    error_sum += e * e;
    if (on && x2_sum > kMfX2SumThreshold && !saturation) {
      const float alpha = ec3::kMfSmoothing * e / x2_sum;
#pragma unroll
      for (int k = 0; k < 32; ++k) h[k] = fmaf(xv[k], alpha, h[k]);
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

// aec3::MaxSquarePeakIndex (matched_filter.cc:558-591) for a 512-tap filter:
// first maximum among even taps, first maximum among odd taps, odd wins only
// when strictly larger.
WAP_DEV int mf_max_square_peak_index(const float* h) {
  const int lane = lane_id();
  float best = -1.f;
  int bi = 0;
  #pragma unroll
  for (int t = lane; t < kMfLen; t += 32) {  // lane parity == tap parity
    const float v = h[t] * h[t];
    if (v > best) { best = v; bi = t; }
  }
  for (int m = 2; m < 32; m <<= 1) {
    const float ov = __shfl_xor_sync(WAP_FULL, best, m);
    const int oi = __shfl_xor_sync(WAP_FULL, bi, m);
    if (ov > best || (ov == best && oi < bi)) { best = ov; bi = oi; }
  }
  const float even_v = __shfl_sync(WAP_FULL, best, 0), odd_v = __shfl_sync(WAP_FULL, best, 1);
  const int even_i = __shfl_sync(WAP_FULL, bi, 0), odd_i = __shfl_sync(WAP_FULL, bi, 1);
  return (odd_v > even_v) ? odd_i : even_i;
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
  const int last_best = s.mf_last_detected_best_lag_filter;
  // Filters whose window does not wrap in the reference's ring during this block (and
  // that do not need the accumulated-error side output) go through the pair path.
  int fast[kNumMatchedFilters];
  int nfast = 0;
  unsigned slow_mask = 0;
  for (int n = 0; n < kNumMatchedFilters; ++n) {
    int x_start0 = s.lr_read + n * kMfShift + kSubBlock - 1;
    if (x_start0 >= kLowRateSize) x_start0 -= kLowRateSize;
    const bool no_wrap = x_start0 >= kSubBlock - 1 && x_start0 + kMfLen <= kLowRateSize;
    if (no_wrap && n != last_best) fast[nfast++] = n;
    else slow_mask |= 1u << n;
  }
  for (int p = 0; p < nfast; p += 2) mf_pair_fast(a, sc, fast[p], p + 1 < nfast ? fast[p + 1] : -1, y);
  for (int n = 0; n < kNumMatchedFilters; ++n) {
    if (!((slow_mask >> n) & 1u)) continue;
    if (n == last_best) {
      mf_acc_filter(a, sc, n, y);
      continue;
    }
    __syncwarp();
    mf_stage_window(a, sc, n, sc.mf.xp);
    #pragma unroll
    for (int t = lane; t < kMfLen; t += 32) (sc.mf.xp + kMfHOffset)[t] = a.mf_h[n][t];
    __syncwarp();
    float error_sum;
    int updated;
    mf_core(sc, n, y, &error_sum, &updated);
    const int peak = mf_max_square_peak_index((sc.mf.xp + kMfHOffset));
    #pragma unroll
    for (int t = lane; t < kMfLen; t += 32) a.mf_h[n][t] = (sc.mf.xp + kMfHOffset)[t];
    if (lane == 0) {
      sc.mf.err_sum[n] = error_sum;
      sc.mf.updated[n] = updated;
      sc.mf.peak[n] = peak;
    }
    __syncwarp();
  }
  // winner selection (matched_filter.cc:729-776), lane 0
  if (lane == 0) {
    float winner_error_sum = error_sum_anchor;
    int has_winner_lag = 0, winner_lag = 0, winner_index = -1;
    int has_prev = 0, prev_lag = 0, alignment_shift = 0;
    for (int n = 0; n < kNumMatchedFilters; ++n) {
      const int lag_estimate = sc.mf.peak[n];
      const float error_sum = sc.mf.err_sum[n];
      const bool reliable = lag_estimate > 2 && lag_estimate < (kMfLen - 10) &&
                            error_sum < ec3::kMfThreshold * error_sum_anchor;
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
    const int headroom = ec3::kHeadroomSamples / kDownSampling;
    // PreEchoLagAggregator::Aggregate (:139-183)
    {
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
      const int sig = s.agg_significant_candidate_found || count > ec3::kThrConverged;
      if (count > ec3::kThrConverged || (count > ec3::kThrInitial && !sig)) {
        has_agg = 1;
        agg_quality = sig ? kQualityRefined : kQualityCoarse;
        agg_delay = s.pre_candidate;
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
      const int hysteresis = use_hysteresis ? ec3::kHysteresisLimitBlocks : 0;
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
