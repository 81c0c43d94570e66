// Time-domain filters of the capture/render front end: cascaded biquads
// (high-pass filter, AEC3 decimator) and the three-band analysis / synthesis
// filter bank.
#pragma once

#include "wap_dev.cuh"
#include "wap_state.h"

namespace wap {

struct BiquadCoef {
  float b0, b1, b2, a0, a1;
};

// Coefficients as published in the reference (signal.* designs quoted there):
// high_pass_filter.cc:25-55 and aec3/decimator.cc:23-52.
WAP_DEVCONST BiquadCoef kHpf16k[3] = {
    {0.8773539420715290582f, -1.754683920749088077f, 0.8773539420715289472f, -1.881687317862849707f, 0.8880584644559580410f},
    {1.0f, -1.999810143464515022f, 1.0f, -1.976035417167170793f, 0.9779708644868606582f},
    {1.0f, -1.999669231394235469f, 1.0f, -1.994265767864654482f, 0.9954861594635392441f}};
WAP_DEVCONST BiquadCoef kHpf32k[3] = {  // high_pass_filter.cc:37-46
    {0.9102055685511306615f, -1.820404922871161624f, 0.9102055685511306615f, -1.940710875829138482f, 0.9423512845457852061f},
    {1.0f, -1.999952541587768806f, 1.0f, -1.988434609801665420f, 0.9889212529819323416f},
    {1.0f, -1.999917315632020021f, 1.0f, -1.997434723613889629f, 0.9977401885079651978f}};
WAP_DEVCONST BiquadCoef kHpf48k[3] = {
    {0.9213790163564168f, -1.8427552370064049f, 0.9213790163564168f, -1.9604500061078971f, 0.9611862979079667f},
    {1.0f, -1.9999789078432082f, 1.0f, -1.9923834169149972f, 0.9926001112941157f},
    {1.0f, -1.9999632520325810f, 1.0f, -1.9983570340145236f, 0.9984928491805198f}};
// Decimator for down_sampling_factor 4: 3-section elliptic low-pass then a
// 1-section high-pass (noise reduction).
WAP_DEVCONST BiquadCoef kDecimator4[4] = {
    {0.0180919877f, 0.00320961363f, 0.0180919877f, -1.5183195f, 0.633165865f},
    {1.0f, -1.24550459f, 1.0f, -1.49784254f, 0.853586692f},
    {1.0f, -1.4221681f, 1.0f, -1.49791282f, 0.969572384f},
    {0.757076375f, -1.51415275f, 0.757076375f, -1.45424359f, 0.574061915f}};

// Cascade of kSections biquads over n samples, in place on `buf` (shared
// memory).  Reference recurrence and evaluation order:
// CascadedBiQuadFilter::ApplyBiQuad (utility/cascaded_biquad_filter.cc:58-84)
//   y = b0*x + b1*x1 + b2*x2 - a0*y1 - a1*y2   (left to right).
// The sections are software-pipelined over lanes: lane s runs section s one
// sample behind lane s-1 and receives its input by shuffle, so the serial
// dependency chain costs n + kSections - 1 steps instead of n * kSections.
// Each section still evaluates the reference expression for every sample, so
// the output is bit-identical to the serial cascade.
template <int kSections>
WAP_DEV void biquad_cascade(float* buf, int n, const BiquadCoef* coef, Biquad* state) {
  const int lane = lane_id();
  const bool on = lane < kSections;
  BiquadCoef c = coef[on ? lane : 0];
  Biquad m = state[on ? lane : 0];
  float y = 0.f;
  for (int step = 0; step < n + kSections - 1; ++step) {
    // Output of the previous section at its previous step == my input now.
    const float from_prev = __shfl_up_sync(WAP_FULL, y, 1);
    const int k = step - lane;
    if (on && k >= 0 && k < n) {
      const float x = (lane == 0) ? buf[k] : from_prev;
      y = c.b0 * x + c.b1 * m.x0 + c.b2 * m.x1 - c.a0 * m.y0 - c.a1 * m.y1;
      m.x1 = m.x0;
      m.x0 = x;
      m.y1 = m.y0;
      m.y0 = y;
      if (lane == kSections - 1) buf[k] = y;
    }
  }
  if (on) state[lane] = m;
  __syncwarp();
}

// ------------------------------------------------------------ three-band bank
// Prototype filter taps and DCT modulation of the reference's 3-band bank
// (three_band_filter_bank.cc:78-108): 12 polyphase components of which two
// are identically zero; the ten non-zero 4-tap sparse filters in order.
WAP_DEVCONST float kBandFilter[10][4] = {
    {-0.00047749f, -0.00496888f, +0.16547118f, +0.00425496f},
    {-0.00173287f, -0.01585778f, +0.14989004f, +0.00994113f},
    {-0.00304815f, -0.02536082f, +0.12154542f, +0.01157993f},
    {-0.00346946f, -0.02587886f, +0.04760441f, +0.00607594f},
    {-0.00154717f, -0.01136076f, +0.01387458f, +0.00186353f},
    {+0.00186353f, +0.01387458f, -0.01136076f, -0.00154717f},
    {+0.00607594f, +0.04760441f, -0.02587886f, -0.00346946f},
    {+0.00983212f, +0.08543175f, -0.02982767f, -0.00383509f},
    {+0.00994113f, +0.14989004f, -0.01585778f, -0.00173287f},
    {+0.00425496f, +0.16547118f, -0.00496888f, -0.00047749f}};
#define WAP_SQRT3 1.73205080756887729f
WAP_DEVCONST float kBandDct[10][3] = {
    {2.f, 2.f, 2.f},        {WAP_SQRT3, 0.f, -WAP_SQRT3}, {1.f, -2.f, 1.f},  {-1.f, 2.f, -1.f},
    {-WAP_SQRT3, 0.f, WAP_SQRT3}, {-2.f, -2.f, -2.f},     {-WAP_SQRT3, 0.f, WAP_SQRT3},
    {-1.f, 2.f, -1.f},      {1.f, -2.f, 1.f},             {WAP_SQRT3, 0.f, -WAP_SQRT3}};

// FilterCore (three_band_filter_bank.cc:112-151) for one output sample k:
//   out = ((((0 + z0*f0) + z1*f1) + z2*f2) + z3*f3),  z_i = in[k - shift - 4i]
// with negative indices read from the 15-sample state.
WAP_DEV float band_filter_sample(const float* in, const float* st, const float* f, int shift, int k) {
  float out = 0.f;
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int j = k - shift - 4 * i;
    const float z = (j >= 0) ? in[j] : st[15 + j];
    out += z * f[i];
  }
  return out;
}

// Analysis: full[480] -> bands[3][160].  Reference ThreeBandFilterBank::Analysis
// (three_band_filter_bank.cc:178-225).  `sub` is a 160-float scratch.  Output
// sample n of band b accumulates the ten modulated filter outputs in filter
// order, exactly like the reference's += over (downsampling_index, in_shift).
WAP_DEV void three_band_analysis(const float* full, float* bands, float* sub, float (*state)[16]) {
  const int lane = lane_id();
  float acc[3][5];
#pragma unroll
  for (int b = 0; b < 3; ++b)
#pragma unroll
    for (int r = 0; r < 5; ++r) acc[b][r] = 0.f;
  for (int ds = 0; ds < 3; ++ds) {
    for (int k = lane; k < 160; k += 32) sub[k] = full[2 - ds + 3 * k];
    __syncwarp();
    for (int shift = 0; shift < 4; ++shift) {
      const int index = ds + shift * 3;
      if (index == 3 || index == 9) continue;
      const int fi = index < 3 ? index : (index < 9 ? index - 1 : index - 2);
#pragma unroll
      for (int r = 0; r < 5; ++r) {
        const int k = lane + 32 * r;
        const float o = band_filter_sample(sub, state[fi], kBandFilter[fi], shift, k);
#pragma unroll
        for (int b = 0; b < 3; ++b) acc[b][r] += kBandDct[fi][b] * o;
      }
    }
    __syncwarp();
    // State update for the filters fed by this downsampled signal happens
    // after all their outputs are formed (each FilterCore call copies the
    // last 15 input samples into its own state).
    for (int shift = 0; shift < 4; ++shift) {
      const int index = ds + shift * 3;
      if (index == 3 || index == 9) continue;
      const int fi = index < 3 ? index : (index < 9 ? index - 1 : index - 2);
      if (lane < 15) state[fi][lane] = sub[160 - 15 + lane];
    }
    __syncwarp();
  }
#pragma unroll
  for (int b = 0; b < 3; ++b)
#pragma unroll
    for (int r = 0; r < 5; ++r) bands[b * 160 + lane + 32 * r] = acc[b][r];
  __syncwarp();
}

// The same analysis run by ONE thread (k_front, one thread per leg): identical expression
// trees, the loops over the 160 output samples are simply serial.  `bands` must be zeroed by
// the caller; `sub` is a 160-float thread-local scratch.
WAP_DEV void three_band_analysis_thread(const float* full, float* bands, float* sub, float (*state)[16]) {
  for (int ds = 0; ds < 3; ++ds) {
    for (int k = 0; k < 160; ++k) sub[k] = full[2 - ds + 3 * k];
    for (int shift = 0; shift < 4; ++shift) {
      const int index = ds + shift * 3;
      if (index == 3 || index == 9) continue;
      const int fi = index < 3 ? index : (index < 9 ? index - 1 : index - 2);
      for (int k = 0; k < 160; ++k) {
        const float o = band_filter_sample(sub, state[fi], kBandFilter[fi], shift, k);
#pragma unroll
        for (int b = 0; b < 3; ++b) bands[b * 160 + k] += kBandDct[fi][b] * o;
      }
    }
    for (int shift = 0; shift < 4; ++shift) {
      const int index = ds + shift * 3;
      if (index == 3 || index == 9) continue;
      const int fi = index < 3 ? index : (index < 9 ? index - 1 : index - 2);
      for (int j = 0; j < 15; ++j) state[fi][j] = sub[160 - 15 + j];
    }
  }
}

// Synthesis: bands[3][160] -> full[480].  Reference ThreeBandFilterBank::Synthesis
// (three_band_filter_bank.cc:233-278).
WAP_DEV void three_band_synthesis(const float* bands, float* full, float* sub, float (*state)[16]) {
  const int lane = lane_id();
  for (int us = 0; us < 3; ++us) {
    float acc[5];
#pragma unroll
    for (int r = 0; r < 5; ++r) acc[r] = 0.f;
    for (int shift = 0; shift < 4; ++shift) {
      const int index = us + shift * 3;
      if (index == 3 || index == 9) continue;
      const int fi = index < 3 ? index : (index < 9 ? index - 1 : index - 2);
      for (int k = lane; k < 160; k += 32) {
        float v = 0.f;
#pragma unroll
        for (int b = 0; b < 3; ++b) v += kBandDct[fi][b] * bands[b * 160 + k];
        sub[k] = v;
      }
      __syncwarp();
#pragma unroll
      for (int r = 0; r < 5; ++r) {
        const int k = lane + 32 * r;
        acc[r] += 3.f * band_filter_sample(sub, state[fi], kBandFilter[fi], shift, k);
      }
      __syncwarp();
      if (lane < 15) state[fi][lane] = sub[160 - 15 + lane];
      __syncwarp();
    }
#pragma unroll
    for (int r = 0; r < 5; ++r) full[us + 3 * (lane + 32 * r)] = acc[r];
  }
  __syncwarp();
}

// ------------------------------------------------------------ two-band QMF (32 kHz)
// SplittingFilter::TwoBands{Analysis,Synthesis} (splitting_filter.cc:68-101) ->
// WebRtcSpl_{Analysis,Synthesis}QMF (common_audio/signal_processing/splitting_filter.c:
// 136-204): the even and odd samples each pass three cascaded first-order all-pass
// sections (WebRtcSpl_AllPassQMF, :31-134), the band signals are their half sum / half
// difference.  The reference keeps six state floats per branch of which two are duplicates;
// four are stored here: {x[-1], y1[-1], y2[-1], y3[-1]}.
WAP_DEVCONST float kQmfAllPass1[3] = {0.0979309082f, 0.5643005371f, 0.8737335205f};
WAP_DEVCONST float kQmfAllPass2[3] = {0.32551574707f, 0.74862670898f, 0.96145629882f};
constexpr int kQmfStateFloats = 4;

// One all-pass branch, serial: y_i[n] = y_{i-1}[n-1] + a_i * (y_{i-1}[n] - y_i[n-1]).
WAP_DEV void qmf_allpass_branch(const float* in, int stride, int n, const float* c, float* st, float* out) {
  float xp = st[0], y1p = st[1], y2p = st[2], y3p = st[3];
  for (int k = 0; k < n; ++k) {
    const float x = in[k * stride];
    const float y1 = xp + c[0] * (x - y1p);
    const float y2 = y1p + c[1] * (y1 - y2p);
    const float y3 = y2p + c[2] * (y2 - y3p);
    out[k] = y3;
    xp = x; y1p = y1; y2p = y2; y3p = y3;
  }
  st[0] = xp; st[1] = y1p; st[2] = y2p; st[3] = y3p;
}

// Analysis by ONE thread (k_front): full[320] -> bands[2][160]; state = 2 x kQmfStateFloats.
WAP_DEV void two_band_analysis_thread(const float* full, float* bands, float* state) {
  qmf_allpass_branch(full + 1, 2, kFrame, kQmfAllPass1, state, bands);                              // odd samples
  qmf_allpass_branch(full, 2, kFrame, kQmfAllPass2, state + kQmfStateFloats, bands + kFrame);       // even samples
  for (int i = 0; i < kFrame; ++i) {
    const float f1 = bands[i], f2 = bands[kFrame + i];
    bands[i] = (f1 + f2) * 0.5f;
    bands[kFrame + i] = (f1 - f2) * 0.5f;
  }
}

// Analysis by a warp: the two serial branches run on lanes 0 and 1.  `tmp`: 320-float scratch.
WAP_DEV void two_band_analysis(const float* full, float* bands, float* tmp, float* state) {
  const int lane = lane_id();
  __syncwarp();
  if (lane == 0) qmf_allpass_branch(full + 1, 2, kFrame, kQmfAllPass1, state, tmp);
  if (lane == 1) qmf_allpass_branch(full, 2, kFrame, kQmfAllPass2, state + kQmfStateFloats, tmp + kFrame);
  __syncwarp();
  #pragma unroll
  for (int i = lane; i < kFrame; i += 32) {
    const float f1 = tmp[i], f2 = tmp[kFrame + i];
    bands[i] = (f1 + f2) * 0.5f;
    bands[kFrame + i] = (f1 - f2) * 0.5f;
  }
  __syncwarp();
}

// Synthesis by a warp: bands[2][160] -> full[320], saturated to the int16 range like the reference.
WAP_DEV void two_band_synthesis(const float* bands, float* full, float* tmp, float* state) {
  const int lane = lane_id();
  __syncwarp();
  #pragma unroll
  for (int i = lane; i < kFrame; i += 32) {
    const float lo = bands[i], hi = bands[kFrame + i];
    tmp[i] = lo + hi;
    tmp[kFrame + i] = lo - hi;
  }
  __syncwarp();
  if (lane == 0) qmf_allpass_branch(tmp, 1, kFrame, kQmfAllPass2, state, tmp);
  if (lane == 1) qmf_allpass_branch(tmp + kFrame, 1, kFrame, kQmfAllPass1, state + kQmfStateFloats, tmp + kFrame);
  __syncwarp();
  #pragma unroll
  for (int i = lane; i < kFrame; i += 32) {
    const float f1 = tmp[i], f2 = tmp[kFrame + i];
    full[2 * i] = f2 > -32768.0f ? (f2 < 32767.0f ? f2 : 32767.0f) : -32768.0f;
    full[2 * i + 1] = f1 > -32768.0f ? (f1 < 32767.0f ? f1 : 32767.0f) : -32768.0f;
  }
  __syncwarp();
}

}  // namespace wap
