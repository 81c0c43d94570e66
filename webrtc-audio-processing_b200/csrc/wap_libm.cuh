// Single-precision libm calls of the reference, restated operation by operation.
//
// The reference's NS calls powf(2.f, p) (ns/fast_math.cc:51) and, after GCC narrows
// static_cast<float>(tanh(float)), tanhf (ns/noise_suppressor.cc:231).  Neither is
// correctly rounded in the glibc the oracle links (2.39): powf differs from the
// rounded exact value for about 1 argument in 1600, tanhf for almost every second
// one, and a 1-ulp difference in a noise estimate lingers for a frame or two.  To be
// identical at float level the two are evaluated here exactly like glibc does:
//   powf(2, p)  sysdeps/ieee754/flt-32/e_powf.c: log2(2) is exactly 1 in its table,
//               so the call reduces to exp2_inline(p): k/32 + r split through the
//               1.5*2^47 shift, 32-entry table of 2^(i/32), cubic in r, in double;
//   tanhf       sysdeps/ieee754/flt-32/s_tanhf.c on top of s_expm1f.c (the fdlibm
//               algorithms), in float.
//   tanh        (double; the NS prior-model indicators call it on floats, speech_probability_
//               estimator.cc:62-84, and add 1 to results near -1, which magnifies a last-bit
//               difference 10^5-fold) sysdeps/ieee754/dbl-64/s_tanh.c on top of s_expm1.c.  On
//               CPUs with FMA glibc 2.39 dispatches to a build of s_expm1.c with contraction
//               enabled (sysdeps/x86_64/fpu/multiarch/s_expm1-fma.c); the fused operations below
//               are the ones that build contracts.  Without them 3 results in 10^4 differ.
// All restatements were checked against libm on several 100 k random arguments
// (tools/check_libm_restatement.py) with zero mismatches, on the host and on the device.
#pragma once

#include "wap_dev.cuh"

namespace wap {

// asuint64(2^(i/32)) - (i << 47), i = 0..31 (each 2^(i/32) correctly rounded to double)
WAP_DEVCONST unsigned long long kExp2fTab[32] = {
    0x3ff0000000000000ull, 0x3fefd9b0d3158574ull, 0x3fefb5586cf9890full, 0x3fef9301d0125b51ull,
    0x3fef72b83c7d517bull, 0x3fef54873168b9aaull, 0x3fef387a6e756238ull, 0x3fef1e9df51fdee1ull,
    0x3fef06fe0a31b715ull, 0x3feef1a7373aa9cbull, 0x3feedea64c123422ull, 0x3feece086061892dull,
    0x3feebfdad5362a27ull, 0x3feeb42b569d4f82ull, 0x3feeab07dd485429ull, 0x3feea47eb03a5585ull,
    0x3feea09e667f3bcdull, 0x3fee9f75e8ec5f74ull, 0x3feea11473eb0187ull, 0x3feea589994cce13ull,
    0x3feeace5422aa0dbull, 0x3feeb737b0cdc5e5ull, 0x3feec49182a3f090ull, 0x3feed503b23e255dull,
    0x3feee89f995ad3adull, 0x3feeff76f2fb5e47ull, 0x3fef199bdd85529cull, 0x3fef3720dcef9069ull,
    0x3fef5818dcfba487ull, 0x3fef7c97337b9b5full, 0x3fefa4afa2a490daull, 0x3fefd0765b6e4540ull,
};

WAP_DEV float libm_pow2f(float p) {
  if (!(fabsf(p) < 126.f)) return (float)exp2((double)p);  // overflow / subnormal results: not reached by NS
  const double c0 = 0.05550361559341535, c1 = 0.2402284522445722, c2 = 0.6931471806916203;
  const double shift = 211106232532992.0;  // 0x1.8p+52 / 32
  const double xd = (double)p;
  double kd = xd + shift;
  const unsigned long long ki = (unsigned long long)__double_as_longlong(kd);
  kd -= shift;  // k / 32
  const double r = xd - kd;
  unsigned long long t = kExp2fTab[ki & 31];
  t += ki << 47;
  const double s = __longlong_as_double((long long)t);
  const double z = c0 * r + c1;
  const double r2 = r * r;
  double y = c2 * r + 1.0;
  y = z * r2 + y;
  y = y * s;
  return (float)y;
}

// expm1f for |x| < 88 (s_expm1f.c).
WAP_DEV float libm_expm1f(float x) {
  const float one = 1.0f;
  const float ln2_hi = __uint_as_float(0x3f317180u), ln2_lo = __uint_as_float(0x3717f7d1u);
  const float invln2 = __uint_as_float(0x3fb8aa3bu);
  const float Q1 = __uint_as_float(0xbd088889u), Q2 = __uint_as_float(0x3ad00d01u), Q3 = __uint_as_float(0xb8a670cdu),
              Q4 = __uint_as_float(0x36867e54u), Q5 = __uint_as_float(0xb457edbbu);
  unsigned hx = __float_as_uint(x);
  const unsigned xsb = hx & 0x80000000u;
  hx &= 0x7fffffffu;
  if (hx >= 0x4195b844u && xsb) return 1.0e-30f - one;  // x <= -27 ln2
  float c = 0.f;
  int k = 0;
  if (hx > 0x3eb17218u) {       // |x| > 0.5 ln2
    float hi, lo;
    if (hx < 0x3F851592u) {     // |x| < 1.5 ln2
      if (!xsb) { hi = x - ln2_hi; lo = ln2_lo; k = 1; }
      else { hi = x + ln2_hi; lo = -ln2_lo; k = -1; }
    } else {
      k = (int)(invln2 * x + (xsb ? -0.5f : 0.5f));
      const float t = (float)k;
      hi = x - t * ln2_hi;
      lo = t * ln2_lo;
    }
    x = hi - lo;
    c = (hi - x) - lo;
  } else if (hx < 0x33000000u) {  // |x| < 2^-25
    return x;
  }
  const float hfx = 0.5f * x;
  const float hxs = x * hfx;
  const float r1 = one + hxs * (Q1 + hxs * (Q2 + hxs * (Q3 + hxs * (Q4 + hxs * Q5))));
  float t = 3.0f - r1 * hfx;
  float e = hxs * ((r1 - t) / (6.0f - x * t));
  if (k == 0) return x - (x * e - hxs);
  e = (x * (e - c) - c);
  e -= hxs;
  if (k == -1) return 0.5f * (x - e) - 0.5f;
  if (k == 1) {
    if (x < -0.25f) return -2.0f * (e - (x + 0.5f));
    return one + 2.0f * (x - e);
  }
  float y;
  if (k <= -2 || k > 56) {
    y = one - (e - x);
    y = __uint_as_float(__float_as_uint(y) + ((unsigned)k << 23));
    return y - one;
  }
  if (k < 23) {
    t = __uint_as_float(0x3f800000u - (0x1000000u >> k));  // 1 - 2^-k
    y = t - (e - x);
    y = __uint_as_float(__float_as_uint(y) + ((unsigned)k << 23));
  } else {
    t = __uint_as_float((unsigned)(0x7f - k) << 23);  // 2^-k
    y = x - (e + t);
    y += one;
    y = __uint_as_float(__float_as_uint(y) + ((unsigned)k << 23));
  }
  return y;
}

// tanhf (s_tanhf.c), finite arguments.
WAP_DEV float libm_tanhf(float x) {
  const unsigned jx = __float_as_uint(x);
  const unsigned ix = jx & 0x7fffffffu;
  float z;
  if (ix < 0x41b00000u) {  // |x| < 22
    if (ix == 0) return x;
    if (ix < 0x24000000u) return x * (1.0f + x);  // |x| < 2^-55
    const float ax = __uint_as_float(ix);
    if (ix >= 0x3f800000u) {  // |x| >= 1
      const float t = libm_expm1f(2.0f * ax);
      z = 1.0f - 2.0f / (t + 2.0f);
    } else {
      const float t = libm_expm1f(-2.0f * ax);
      z = -t / (t + 2.0f);
    }
  } else {
    z = 1.0f - 1.0e-30f;
  }
  return (jx & 0x80000000u) ? -z : z;
}

// expm1 (double) as the FMA build of glibc's s_expm1.c evaluates it; |x| < 700.
WAP_DEV double libm_expm1(double x) {
  const double ln2_hi = 6.93147180369123816490e-01, ln2_lo = 1.90821492927058770002e-10,
               invln2 = 1.44269504088896338700e+00;
  const double Q1 = -3.33333333333331316428e-02, Q2 = 1.58730158725481460165e-03, Q3 = -7.93650757867487942473e-05,
               Q4 = 4.00821782732936239552e-06, Q5 = -2.01099218183624371326e-07;
  unsigned hx = (unsigned)((unsigned long long)__double_as_longlong(x) >> 32);
  const unsigned xsb = hx & 0x80000000u;
  hx &= 0x7fffffffu;
  if (hx >= 0x4043687Au && xsb) return 1.0e-300 - 1.0;  // x <= -56 ln2
  double c = 0.0;
  int k = 0;
  if (hx > 0x3fd62e42u) {        // |x| > 0.5 ln2
    double hi, lo;
    if (hx < 0x3FF0A2B2u) {      // |x| < 1.5 ln2
      if (!xsb) { hi = x - ln2_hi; lo = ln2_lo; k = 1; }
      else { hi = x + ln2_hi; lo = -ln2_lo; k = -1; }
    } else {
      k = (int)fma(invln2, x, xsb ? -0.5 : 0.5);
      const double t = (double)k;
      hi = fma(-t, ln2_hi, x);
      lo = t * ln2_lo;
    }
    x = hi - lo;
    c = (hi - x) - lo;
  } else if (hx < 0x3c900000u) {  // |x| < 2^-54
    return x;
  }
  const double hfx = 0.5 * x;
  const double hxs = x * hfx;
  const double R1 = fma(hxs, Q1, 1.0), h2 = hxs * hxs, R2 = fma(hxs, Q3, Q2), h4 = h2 * h2, R3 = fma(hxs, Q5, Q4);
  const double r1 = fma(h4, R3, fma(h2, R2, R1));
  double t = fma(-r1, hfx, 3.0);
  double e = hxs * ((r1 - t) / fma(-x, t, 6.0));
  if (k == 0) return x - fma(x, e, -hxs);
  e = fma(x, e - c, -c);
  e -= hxs;
  if (k == -1) return fma(0.5, x - e, -0.5);
  if (k == 1) {
    if (x < -0.25) return -2.0 * (e - (x + 0.5));
    return fma(2.0, x - e, 1.0);
  }
  double y;
  if (k <= -2 || k > 56) {
    y = 1.0 - (e - x);
    y = __longlong_as_double(__double_as_longlong(y) + ((long long)k << 52));
    return y - 1.0;
  }
  if (k < 20) {
    t = __longlong_as_double((long long)(0x3ff00000u - (0x200000u >> k)) << 32);  // 1 - 2^-k
    y = t - (e - x);
    y = __longlong_as_double(__double_as_longlong(y) + ((long long)k << 52));
  } else {
    t = __longlong_as_double((long long)(0x3ff - k) << 52);  // 2^-k
    y = x - (e + t);
    y += 1.0;
    y = __longlong_as_double(__double_as_longlong(y) + ((long long)k << 52));
  }
  return y;
}

// tanh (double, s_tanh.c), finite arguments.
WAP_DEV double libm_tanh(double x) {
  const unsigned long long bits = (unsigned long long)__double_as_longlong(x);
  const unsigned jx = (unsigned)(bits >> 32), lx = (unsigned)bits;
  const unsigned ix = jx & 0x7fffffffu;
  double z;
  if (ix < 0x40360000u) {  // |x| < 22
    if ((ix | lx) == 0) return x;
    if (ix < 0x3c800000u) return x * (1.0 + x);  // |x| < 2^-55
    const double ax = fabs(x);
    if (ix >= 0x3ff00000u) {  // |x| >= 1
      const double t = libm_expm1(2.0 * ax);
      z = 1.0 - 2.0 / (t + 2.0);
    } else {
      const double t = libm_expm1(-2.0 * ax);
      z = -t / (t + 2.0);
    }
  } else {
    z = 1.0 - 1.0e-300;
  }
  return (jx & 0x80000000u) ? -z : z;
}

}  // namespace wap
