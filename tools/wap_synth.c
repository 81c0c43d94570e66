// Synthetic call-leg generator of SURVEY.md section 8(d) -- measurement / test infrastructure, not
// product code.  One generator for BOTH bench arms (the B200 engine and the reference CPU arm) and
// for the parity spot check, so that every arm sees the same int16 frames.
//
// Restates webrtc::Random (reference rtc_base/random.h:71-77, random.cc:52-56: xorshift64* with
// shifts 12/25/27 and multiplier 2685821657736338717; Rand<float>() = float((NextOutput()-1) /
// (2^64-1))) and RandomizeSampleVector (reference tests/test_utils/echo_canceller_test_tools.cc:
// 28-35: v = 2*amplitude*Rand<float>() - amplitude in float).
//
// Leg i (seed_r = 1000+2i, seed_n = 1001+2i), one CYCLE of n = frames*rate/100 samples that
// repeats seamlessly (the echo path wraps around the cycle):
//   kind 0 (BASELINE configs 1/2/5): render = white noise, amplitude 8000, gated 0.9 s on / 0.1 s
//     off; capture = 0.5 x[n-D] + 0.25 x[n-D-37] + 0.1 x[n-D-160], D = (rate/16000) * (64 (1 + i mod
//     48) + 7i mod 64), + noise floor (amplitude 50) + double-talk bursts (amplitude 3000, the last
//     0.3 s of every 2 s).
//   kind 1 (BASELINE config 3, near end only): noise amplitude 300 + 1 kHz and 2.3 kHz tones of
//     amplitude 4000 gated 0.5 s on / 0.5 s off; render = 0.
//   kind 2 (BASELINE config 4, stereo render and capture): render L / R = independent white noise
//     (seeds seed_r, seed_r + 7919), amplitude 8000, gated like kind 0; capture channel c =
//     0.5 L[n-Dc] + 0.25 L[n-Dc-37] + 0.1 L[n-Dc-160] + 0.35 R[n-Dc-11] + 0.15 R[n-Dc-53], D0 = D,
//     D1 = D + 3, + an independent noise floor per channel (seeds seed_n, seed_n + 7919, amplitude 50)
//     + the same double-talk bursts on both channels.  Frames are interleaved stereo.
// Output layout: [frame][leg][samples per frame (x 2 interleaved channels for kind 2)] int16 (what
// wap_process_streams takes per tick).
//
// build: gcc -O2 -shared -fPIC -fopenmp tools/wap_synth.c -o tools/_build/libwap_synth.so -lm
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

typedef struct { uint64_t s; } Rng;
static inline uint64_t next_output(Rng* r) {
  uint64_t s = r->s;
  s ^= s >> 12;
  s ^= s << 25;
  s ^= s >> 27;
  r->s = s;
  return s * 2685821657736338717ull;
}
static inline float rand_float(Rng* r) {
  double v = (double)(next_output(r) - 1);
  v = v / (double)0xFFFFFFFFFFFFFFFFull;
  return (float)v;
}
static inline float sample(Rng* r, float amplitude) { return 2 * amplitude * rand_float(r) - amplitude; }
static inline int16_t q16(double v) {
  v = rint(v);
  if (v > 32767.0) v = 32767.0;
  if (v < -32768.0) v = -32768.0;
  return (int16_t)v;
}

// render / capture: [frames][legs][rate/100] int16; legs first_leg .. first_leg+legs-1.
int wap_synth_cycle(int kind, int rate, int first_leg, int legs, int frames, int16_t* render, int16_t* capture) {
  const int fl = rate / 100;
  const int n = frames * fl;
  int ok = 1;
#pragma omp parallel
  {
    float* x = (float*)malloc(sizeof(float) * (size_t)n);
    if (!x) ok = 0;
#pragma omp for schedule(static)
    for (int l = 0; l < legs; ++l) {
      if (!x) continue;
      const int i = first_leg + l;
      Rng rr = {(uint64_t)(1000 + 2 * (int64_t)i)}, rn = {(uint64_t)(1001 + 2 * (int64_t)i)};
      if (kind == 0) {
        for (int k = 0; k < n; ++k) {
          const float v = sample(&rr, 8000.f);
          x[k] = ((k % rate) < (rate / 10) * 9) ? v : 0.f;
        }
        const int D = (rate / 16000) * (64 * (1 + (i % 48)) + (7 * i) % 64);
        for (int k = 0; k < n; ++k) {
          const float floor_ = sample(&rn, 50.f), burst = sample(&rn, 3000.f);
          double y = 0.5 * x[((k - D) % n + n) % n] + 0.25 * x[((k - D - 37) % n + n) % n] +
                     0.1 * x[((k - D - 160) % n + n) % n];
          y += floor_;
          if ((k % (2 * rate)) >= (rate / 10) * 17) y += burst;
          const size_t o = ((size_t)(k / fl) * legs + l) * fl + (k % fl);
          render[o] = q16(x[k]);
          capture[o] = q16(y);
        }
      } else if (kind == 2) {
        float* xr = (float*)malloc(sizeof(float) * (size_t)n);
        if (!xr) { ok = 0; continue; }
        Rng rr2 = {(uint64_t)(1000 + 2 * (int64_t)i + 7919)}, rn2 = {(uint64_t)(1001 + 2 * (int64_t)i + 7919)};
        for (int k = 0; k < n; ++k) {
          const float v = sample(&rr, 8000.f), w = sample(&rr2, 8000.f);
          const int on = (k % rate) < (rate / 10) * 9;
          x[k] = on ? v : 0.f;
          xr[k] = on ? w : 0.f;
        }
        const int D = (rate / 16000) * (64 * (1 + (i % 48)) + (7 * i) % 64);
#define WAP_AT(a, d) (a)[(((k - (d)) % n) + n) % n]
        for (int k = 0; k < n; ++k) {
          const float floor0 = sample(&rn, 50.f), burst = sample(&rn, 3000.f), floor1 = sample(&rn2, 50.f);
          const size_t o = (((size_t)(k / fl) * legs + l) * fl + (k % fl)) * 2;
          for (int c = 0; c < 2; ++c) {
            const int Dc = D + 3 * c;
            double y = 0.5 * WAP_AT(x, Dc) + 0.25 * WAP_AT(x, Dc + 37) + 0.1 * WAP_AT(x, Dc + 160) +
                       0.35 * WAP_AT(xr, Dc + 11) + 0.15 * WAP_AT(xr, Dc + 53);
            y += c ? floor1 : floor0;
            if ((k % (2 * rate)) >= (rate / 10) * 17) y += burst;
            capture[o + c] = q16(y);
          }
          render[o] = q16(x[k]);
          render[o + 1] = q16(xr[k]);
        }
#undef WAP_AT
        free(xr);
      } else {
        for (int k = 0; k < n; ++k) {
          double y = sample(&rn, 300.f);
          if ((k % rate) < rate / 2) {
            const double t = (double)k / rate;
            y += 4000.0 * sin(2 * M_PI * 1000.0 * t) + 4000.0 * sin(2 * M_PI * 2300.0 * t);
          }
          const size_t o = ((size_t)(k / fl) * legs + l) * fl + (k % fl);
          if (render) render[o] = 0;
          capture[o] = q16(y);
        }
      }
    }
    free(x);
  }
  return ok ? 0 : -1;
}
