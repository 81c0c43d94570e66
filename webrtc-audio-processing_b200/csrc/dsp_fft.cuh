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

// out = (c*x.r - s*x.i, c*x.i + s*x.r), separate multiplies and add/sub.
WAP_DEV void cmul_store(float* p, float c, float s, float xr, float xi) {
  p[0] = c * xr - s * xi;
  p[1] = c * xi + s * xr;
}

// One radix-4 butterfly of the stage with element stride l (in floats).
// Block q = 0 / 1 are the twiddle-free and the cos(pi/4) special cases of
// Ooura's cft1st/cftmdl unless `table_for_all` (SSE2 cft1st_128 treats every
// block with table twiddles, ooura_fft_sse2.cc:22-83).
WAP_DEV void r4_butterfly(float* a, int l, int t, const float* tw, bool table_for_all) {
  const int per = l >> 1;
  const int q = t / per;
  const int j = t - q * per;
  float* p0 = a + q * 4 * l + 2 * j;
  float* p1 = p0 + l;
  float* p2 = p1 + l;
  float* p3 = p2 + l;
  const float x0r = p0[0] + p1[0], x0i = p0[1] + p1[1];
  const float x1r = p0[0] - p1[0], x1i = p0[1] - p1[1];
  const float x2r = p2[0] + p3[0], x2i = p2[1] + p3[1];
  const float x3r = p2[0] - p3[0], x3i = p2[1] - p3[1];
  p0[0] = x0r + x2r;
  p0[1] = x0i + x2i;
  if (q == 0 && !table_for_all) {
    p2[0] = x0r - x2r;
    p2[1] = x0i - x2i;
    p1[0] = x1r - x3i;
    p1[1] = x1i + x3r;
    p3[0] = x1r + x3i;
    p3[1] = x1i - x3r;
  } else if (q == 1 && !table_for_all) {
    const float w = tw[6];
    p2[0] = x2i - x0i;
    p2[1] = x0r - x2r;
    float yr = x1r - x3i, yi = x1i + x3r;
    p1[0] = w * (yr - yi);
    p1[1] = w * (yr + yi);
    yr = x3i + x1r;
    yi = x3r - x1i;
    p3[0] = w * (yi - yr);
    p3[1] = w * (yi + yr);
  } else {
    const float* w = tw + 6 * q;
    cmul_store(p2, w[2], w[3], x0r - x2r, x0i - x2i);
    cmul_store(p1, w[0], w[1], x1r - x3i, x1i + x3r);
    cmul_store(p3, w[4], w[5], x1r + x3i, x1i - x3r);
  }
}

// Last radix-4 stage of cftfsub / cftbsub (a single block, no twiddles).
template <bool kInverse>
WAP_DEV void r4_final(float* a, int l, int t) {
  float* p0 = a + 2 * t;
  float* p1 = p0 + l;
  float* p2 = p1 + l;
  float* p3 = p2 + l;
  if (!kInverse) {
    const float x0r = p0[0] + p1[0], x0i = p0[1] + p1[1];
    const float x1r = p0[0] - p1[0], x1i = p0[1] - p1[1];
    const float x2r = p2[0] + p3[0], x2i = p2[1] + p3[1];
    const float x3r = p2[0] - p3[0], x3i = p2[1] - p3[1];
    p0[0] = x0r + x2r; p0[1] = x0i + x2i;
    p2[0] = x0r - x2r; p2[1] = x0i - x2i;
    p1[0] = x1r - x3i; p1[1] = x1i + x3r;
    p3[0] = x1r + x3i; p3[1] = x1i - x3r;
  } else {
    const float x0r = p0[0] + p1[0], x0i = -p0[1] - p1[1];
    const float x1r = p0[0] - p1[0], x1i = -p0[1] + p1[1];
    const float x2r = p2[0] + p3[0], x2i = p2[1] + p3[1];
    const float x3r = p2[0] - p3[0], x3i = p2[1] - p3[1];
    p0[0] = x0r + x2r; p0[1] = x0i - x2i;
    p2[0] = x0r - x2r; p2[1] = x0i + x2i;
    p1[0] = x1r - x3i; p1[1] = x1i - x3r;
    p3[0] = x1r + x3i; p3[1] = x1i + x3r;
  }
}

// Real-FFT split pass for one (j, n-j) bin pair; `c` is the half-cosine table.
template <bool kInverse>
WAP_DEV void rft_pair(float* a, int n, int j2, float wkr_src, float wki) {
  const int k2 = n - j2;
  const float wkr = 0.5f - wkr_src;
  const float xr = a[j2] - a[k2];
  const float xi = a[j2 + 1] + a[k2 + 1];
  if (!kInverse) {
    const float yr = wkr * xr - wki * xi;
    const float yi = wkr * xi + wki * xr;
    a[j2] -= yr;
    a[j2 + 1] -= yi;
    a[k2] += yr;
    a[k2 + 1] -= yi;
  } else {
    const float yr = wkr * xr + wki * xi;
    const float yi = wkr * xi - wki * xr;
    a[j2] = a[j2] - yr;
    a[j2 + 1] = yi - a[j2 + 1];
    a[k2] = yr + a[k2];
    a[k2 + 1] = yi - a[k2 + 1];
  }
}

// ------------------------------------------------------------ 128-point
// `a` points at this half-warp's 128-float packed array, t = lane & 15,
// `on` = this half-warp has a transform to do (all lanes must call).
WAP_DEV void bitrev128(float* a, int t, bool on) {
  float v[8];
  if (on) {
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const int s = kBitrv128[t + 16 * k];
      v[2 * k] = a[2 * s];
      v[2 * k + 1] = a[2 * s + 1];
    }
  }
  __syncwarp();
  if (on) {
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      a[2 * (t + 16 * k)] = v[2 * k];
      a[2 * (t + 16 * k) + 1] = v[2 * k + 1];
    }
  }
  __syncwarp();
}

// In-place forward transform: time samples a[0..127] -> Ooura packed spectrum
// (a[0]=Re0, a[1]=Re64, a[2k],a[2k+1] = Re k, Im k).  Unscaled.
WAP_DEV void fft128_forward(float* a, int t, bool on) {
  bitrev128(a, t, on);
  if (on) r4_butterfly(a, 2, t, kTw128, true);
  __syncwarp();
  if (on) r4_butterfly(a, 8, t, kTw128, false);
  __syncwarp();
  if (on) r4_final<false>(a, 32, t);
  __syncwarp();
  if (on) {
    rft_pair<false>(a, 128, 2 * (t + 1), kRc128[32 - (t + 1)], kRc128[t + 1]);
    if (t + 17 < 32) rft_pair<false>(a, 128, 2 * (t + 17), kRc128[32 - (t + 17)], kRc128[t + 17]);
    if (t == 0) {
      // a[0], a[1] are not touched by the pair pass (j2 >= 2, k2 <= 126).
      const float xi = a[0] - a[1];
      a[0] += a[1];
      a[1] = xi;
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
  bitrev128(a, t, on);
  if (on) r4_butterfly(a, 2, t, kTw128, true);
  __syncwarp();
  if (on) r4_butterfly(a, 8, t, kTw128, false);
  __syncwarp();
  if (on) r4_final<true>(a, 32, t);
  __syncwarp();
}

// ------------------------------------------------------------ 256-point
WAP_DEV void bitrev256(float* a, int lane) {
  float v[8];
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    const int s = kBitrv256[lane + 32 * k];
    v[2 * k] = a[2 * s];
    v[2 * k + 1] = a[2 * s + 1];
  }
  __syncwarp();
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    a[2 * (lane + 32 * k)] = v[2 * k];
    a[2 * (lane + 32 * k) + 1] = v[2 * k + 1];
  }
  __syncwarp();
}

// Forward WebRtc_rdft(256, +1): a[0]=Re0, a[1]=Re128, then (Re k, Im k).
WAP_DEV void fft256_forward(float* a, int lane) {
  bitrev256(a, lane);
  r4_butterfly(a, 2, lane, kTw256, false);
  __syncwarp();
  r4_butterfly(a, 8, lane, kTw256, false);
  __syncwarp();
  r4_butterfly(a, 32, lane, kTw256, false);
  __syncwarp();
#pragma unroll
  for (int k = 0; k < 2; ++k) {  // radix-2 tail of cftfsub, l = 128
    float* p = a + 2 * (lane + 32 * k);
    float* q = p + 128;
    const float x0r = p[0] - q[0], x0i = p[1] - q[1];
    p[0] += q[0];
    p[1] += q[1];
    q[0] = x0r;
    q[1] = x0i;
  }
  __syncwarp();
#pragma unroll
  for (int k = 0; k < 2; ++k) {
    const int kk = lane + 32 * k + 1;  // 1..64, pairs exist for kk <= 63
    if (kk < 64) rft_pair<false>(a, 256, 2 * kk, kRc256[64 - kk], kRc256[kk]);
  }
  if (lane == 0) {
    const float xi = a[0] - a[1];
    a[0] += a[1];
    a[1] = xi;
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
  bitrev256(a, lane);
  r4_butterfly(a, 2, lane, kTw256, false);
  __syncwarp();
  r4_butterfly(a, 8, lane, kTw256, false);
  __syncwarp();
  r4_butterfly(a, 32, lane, kTw256, false);
  __syncwarp();
#pragma unroll
  for (int k = 0; k < 2; ++k) {  // radix-2 tail of cftbsub
    float* p = a + 2 * (lane + 32 * k);
    float* q = p + 128;
    const float x0r = p[0] - q[0], x0i = -p[1] + q[1];
    p[0] += q[0];
    p[1] = -p[1] - q[1];
    q[0] = x0r;
    q[1] = x0i;
  }
  __syncwarp();
}

}  // namespace wap
