// Warp-cooperative real FFTs on shared memory, bit-compatible with the two
// Ooura variants the reference uses:
//
//  * 128-point (AEC3): reference OouraFft::Fft / InverseFft
//    (common_audio/third_party/ooura/fft_size_128/ooura_fft.cc:334-349) in the
//    operation order of its SSE2 bodies (ooura_fft_sse2.cc:22-425), which is
//    what both x86 SIMD builds of the reference run (aec3_fft.cc:78-88).
//  * 256-point (NS): reference WebRtc_rdft (fft_size_256/fft4g.cc:828-866),
//    scalar C order.
//
// Both are a bit-reversal, three radix-4 decimation stages (plus a radix-2
// stage for 256) and a real-FFT post/pre pass.  A stage has N/8 independent
// radix-4 butterflies on disjoint elements, so the butterflies of a stage are
// spread over lanes and stages are separated by __syncwarp().  Each
// butterfly evaluates exactly the expression tree of the reference, hence the
// result is bit-identical regardless of the lane mapping.
//
// 128-point transforms use 16 lanes; a warp can run two of them side by side
// (lanes 0-15 -> transform A, lanes 16-31 -> transform B).
#pragma once

#include "wap_dev.cuh"
#include "wap_tables.inc"

namespace wap {

struct Cx {
  float r, i;
};

// One complex element as a single 64-bit access.  Every transform buffer is 8-byte aligned and the
// butterflies only touch (re, im) pairs at even float offsets.
WAP_DEV float2 ldc(const float* p) { return *reinterpret_cast<const float2*>(p); }
WAP_DEV void stc(float* p, float r, float i) { *reinterpret_cast<float2*>(p) = make_float2(r, i); }

// out = (c*x.r - s*x.i, c*x.i + s*x.r), separate multiplies and add/sub.
WAP_DEV void cmul_store(float* p, float c, float s, float xr, float xi) {
  const float re = c * xr - s * xi;
  const float im = c * xi + s * xr;
  stc(p, re, im);
}

// One radix-4 butterfly of the stage with element stride l (in floats).
// Block q = 0 / 1 are the twiddle-free and the cos(pi/4) special cases of
// Ooura's cft1st/cftmdl unless `table_for_all` (SSE2 cft1st_128 treats every
// block with table twiddles, ooura_fft_sse2.cc:22-83).
WAP_DEV void r4_butterfly_values(float* p0, float* p1, float* p2, float* p3, float2 a0, float2 a1, float2 a2, float2 a3, int q,
                                 const float* tw, bool table_for_all) {
  const float x0r = a0.x + a1.x, x0i = a0.y + a1.y;
  const float x1r = a0.x - a1.x, x1i = a0.y - a1.y;
  const float x2r = a2.x + a3.x, x2i = a2.y + a3.y;
  const float x3r = a2.x - a3.x, x3i = a2.y - a3.y;
  stc(p0, x0r + x2r, x0i + x2i);
  if (q == 0 && !table_for_all) {
    stc(p2, x0r - x2r, x0i - x2i);
    stc(p1, x1r - x3i, x1i + x3r);
    stc(p3, x1r + x3i, x1i - x3r);
  } else if (q == 1 && !table_for_all) {
    const float w = tw[6];
    stc(p2, x2i - x0i, x0r - x2r);
    float yr = x1r - x3i, yi = x1i + x3r;
    stc(p1, w * (yr - yi), w * (yr + yi));
    yr = x3i + x1r;
    yi = x3r - x1i;
    stc(p3, w * (yi - yr), w * (yi + yr));
  } else {
    // the six twiddles of block q, three 64-bit loads (the tables are 16-byte aligned, 6 * q is even)
    const float* w = tw + 6 * q;
    const float2 w1 = ldc(w), w2 = ldc(w + 2), w3 = ldc(w + 4);
    cmul_store(p2, w2.x, w2.y, x0r - x2r, x0i - x2i);
    cmul_store(p1, w1.x, w1.y, x1r - x3i, x1i + x3r);
    cmul_store(p3, w3.x, w3.y, x1r + x3i, x1i - x3r);
  }
}
WAP_DEV void r4_butterfly(float* a, int l, int t, const float* tw, bool table_for_all) {
  const int per = l >> 1;
  const int q = t / per;
  const int j = t - q * per;
  float* p0 = a + q * 4 * l + 2 * j;
  float* p1 = p0 + l;
  float* p2 = p1 + l;
  float* p3 = p2 + l;
  r4_butterfly_values(p0, p1, p2, p3, ldc(p0), ldc(p1), ldc(p2), ldc(p3), q, tw, table_for_all);
}

// Last radix-4 stage of cftfsub / cftbsub (a single block, no twiddles).
template <bool kInverse>
WAP_DEV void r4_final(float* a, int l, int t) {
  float* p0 = a + 2 * t;
  float* p1 = p0 + l;
  float* p2 = p1 + l;
  float* p3 = p2 + l;
  const float2 a0 = ldc(p0), a1 = ldc(p1), a2 = ldc(p2), a3 = ldc(p3);
  if (!kInverse) {
    const float x0r = a0.x + a1.x, x0i = a0.y + a1.y;
    const float x1r = a0.x - a1.x, x1i = a0.y - a1.y;
    const float x2r = a2.x + a3.x, x2i = a2.y + a3.y;
    const float x3r = a2.x - a3.x, x3i = a2.y - a3.y;
    stc(p0, x0r + x2r, x0i + x2i);
    stc(p2, x0r - x2r, x0i - x2i);
    stc(p1, x1r - x3i, x1i + x3r);
    stc(p3, x1r + x3i, x1i - x3r);
  } else {
    const float x0r = a0.x + a1.x, x0i = -a0.y - a1.y;
    const float x1r = a0.x - a1.x, x1i = -a0.y + a1.y;
    const float x2r = a2.x + a3.x, x2i = a2.y + a3.y;
    const float x3r = a2.x - a3.x, x3i = a2.y - a3.y;
    stc(p0, x0r + x2r, x0i - x2i);
    stc(p2, x0r - x2r, x0i + x2i);
    stc(p1, x1r - x3i, x1i - x3r);
    stc(p3, x1r + x3i, x1i + x3r);
  }
}

// Real-FFT split pass for one (j, n-j) bin pair; `c` is the half-cosine table.
template <bool kInverse>
WAP_DEV void rft_pair(float* a, int n, int j2, float wkr_src, float wki) {
  const int k2 = n - j2;
  const float wkr = 0.5f - wkr_src;
  const float2 aj = ldc(a + j2), ak = ldc(a + k2);
  const float xr = aj.x - ak.x;
  const float xi = aj.y + ak.y;
  if (!kInverse) {
    const float yr = wkr * xr - wki * xi;
    const float yi = wkr * xi + wki * xr;
    stc(a + j2, aj.x - yr, aj.y - yi);
    stc(a + k2, ak.x + yr, ak.y - yi);
  } else {
    const float yr = wkr * xr + wki * xi;
    const float yi = wkr * xi - wki * xr;
    stc(a + j2, aj.x - yr, yi - aj.y);
    stc(a + k2, yr + ak.x, yi - ak.y);
  }
}

// ------------------------------------------------------------ 128-point
// `a` points at this half-warp's 128-float packed array, t = lane & 15,
// `on` = this half-warp has a transform to do (all lanes must call).
// bitrv2 followed by the first radix-4 stage (cft1st): lane t's butterfly works on the complex
// elements 4t..4t+3 of the permuted array, i.e. on a[kBitrv128[4t + m]]: they are fetched from
// their source positions and the permuted array is never written.
WAP_DEV void bitrev_first_stage128(float* a, int t, bool on) {
  float2 v0, v1, v2, v3;
  if (on) {
    const uchar4 s = *reinterpret_cast<const uchar4*>(kBitrv128 + 4 * t);  // one 32-bit table load
    v0 = ldc(a + 2 * s.x);
    v1 = ldc(a + 2 * s.y);
    v2 = ldc(a + 2 * s.z);
    v3 = ldc(a + 2 * s.w);
  }
  __syncwarp();
  if (on) {
    float* p0 = a + 8 * t;
    r4_butterfly_values(p0, p0 + 2, p0 + 4, p0 + 6, v0, v1, v2, v3, t, kTw128, true);
  }
  __syncwarp();
}

// In-place forward transform: time samples a[0..127] -> Ooura packed spectrum
// (a[0]=Re0, a[1]=Re64, a[2k],a[2k+1] = Re k, Im k).  Unscaled.
WAP_DEV void fft128_forward(float* a, int t, bool on) {
  bitrev_first_stage128(a, t, on);
  if (on) r4_butterfly(a, 8, t, kTw128, false);
  __syncwarp();
  if (on) r4_final<false>(a, 32, t);
  __syncwarp();
  if (on) {
    rft_pair<false>(a, 128, 2 * (t + 1), kRc128[32 - (t + 1)], kRc128[t + 1]);
    if (t + 17 < 32) rft_pair<false>(a, 128, 2 * (t + 17), kRc128[32 - (t + 17)], kRc128[t + 17]);
    if (t == 0) {
      // a[0], a[1] are not touched by the pair pass (j2 >= 2, k2 <= 126).
      const float2 dc = ldc(a);
      stc(a, dc.x + dc.y, dc.x - dc.y);
    }
  }
  __syncwarp();
}

// In-place inverse transform of an Ooura packed spectrum; the caller applies
// the 1/64 (or 2/128) scaling like the reference's callers do.
WAP_DEV void fft128_inverse(float* a, int t, bool on) {
  if (on) {
    rft_pair<true>(a, 128, 2 * (t + 1), kRc128[32 - (t + 1)], kRc128[t + 1]);
    if (t + 17 < 32) rft_pair<true>(a, 128, 2 * (t + 17), kRc128[32 - (t + 17)], kRc128[t + 17]);
    if (t == 0) {
      a[1] = 0.5f * (a[0] - a[1]);
      a[0] -= a[1];
      a[1] = -a[1];
      a[65] = -a[65];
    }
  }
  __syncwarp();
  bitrev_first_stage128(a, t, on);
  if (on) r4_butterfly(a, 8, t, kTw128, false);
  __syncwarp();
  if (on) r4_final<true>(a, 32, t);
  __syncwarp();
}

// ------------------------------------------------------------ 256-point
// bitrv2 + first radix-4 stage of the 256-point transform (one butterfly per lane), see
// bitrev_first_stage128.
WAP_DEV void bitrev_first_stage256(float* a, int lane) {
  const uchar4 s = *reinterpret_cast<const uchar4*>(kBitrv256 + 4 * lane);
  const float2 v0 = ldc(a + 2 * s.x);
  const float2 v1 = ldc(a + 2 * s.y);
  const float2 v2 = ldc(a + 2 * s.z);
  const float2 v3 = ldc(a + 2 * s.w);
  __syncwarp();
  float* p0 = a + 8 * lane;
  r4_butterfly_values(p0, p0 + 2, p0 + 4, p0 + 6, v0, v1, v2, v3, lane, kTw256, false);
  __syncwarp();
}

// Forward WebRtc_rdft(256, +1): a[0]=Re0, a[1]=Re128, then (Re k, Im k).
WAP_DEV void fft256_forward(float* a, int lane) {
  bitrev_first_stage256(a, lane);
  r4_butterfly(a, 8, lane, kTw256, false);
  __syncwarp();
  r4_butterfly(a, 32, lane, kTw256, false);
  __syncwarp();
#pragma unroll
  for (int k = 0; k < 2; ++k) {  // radix-2 tail of cftfsub, l = 128
    float* p = a + 2 * (lane + 32 * k);
    float* q = p + 128;
    const float2 vp = ldc(p), vq = ldc(q);
    stc(p, vp.x + vq.x, vp.y + vq.y);
    stc(q, vp.x - vq.x, vp.y - vq.y);
  }
  __syncwarp();
#pragma unroll
  for (int k = 0; k < 2; ++k) {
    const int kk = lane + 32 * k + 1;  // 1..64, pairs exist for kk <= 63
    if (kk < 64) rft_pair<false>(a, 256, 2 * kk, kRc256[64 - kk], kRc256[kk]);
  }
  if (lane == 0) {
    const float2 dc = ldc(a);
    stc(a, dc.x + dc.y, dc.x - dc.y);
  }
  __syncwarp();
}

// Inverse WebRtc_rdft(256, -1); caller scales by 2/256 (ns_fft.cc:62-66).
WAP_DEV void fft256_inverse(float* a, int lane) {
#pragma unroll
  for (int k = 0; k < 2; ++k) {
    const int kk = lane + 32 * k + 1;
    if (kk < 64) rft_pair<true>(a, 256, 2 * kk, kRc256[64 - kk], kRc256[kk]);
  }
  if (lane == 0) {
    a[1] = 0.5f * (a[0] - a[1]);
    a[0] -= a[1];
    a[1] = -a[1];
    a[129] = -a[129];
  }
  __syncwarp();
  bitrev_first_stage256(a, lane);
  r4_butterfly(a, 8, lane, kTw256, false);
  __syncwarp();
  r4_butterfly(a, 32, lane, kTw256, false);
  __syncwarp();
#pragma unroll
  for (int k = 0; k < 2; ++k) {  // radix-2 tail of cftbsub
    float* p = a + 2 * (lane + 32 * k);
    float* q = p + 128;
    const float2 vp = ldc(p), vq = ldc(q);
    stc(p, vp.x + vq.x, -vp.y - vq.y);
    stc(q, vp.x - vq.x, -vp.y + vq.y);
  }
  __syncwarp();
}

}  // namespace wap
