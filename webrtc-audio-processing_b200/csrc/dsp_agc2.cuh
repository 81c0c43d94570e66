// AGC2 in its default sub-configuration (gain_controller2.enabled, adaptive_digital
// disabled): fixed digital gain + limiter on the full-band capture frame, one warp per leg.
//   GainController2::Process                 gain_controller2.cc:183-260
//   GainApplier::ApplyGain                   agc2/gain_applier.cc:39-89
//   Limiter::Process                         agc2/limiter.cc:40-132
//   FixedDigitalLevelEstimator::ComputeLevel agc2/fixed_digital_level_estimator.cc:57-112
//   InterpolatedGainCurve::LookUpGainToApply agc2/interpolated_gain_curve.cc:162-197,
//                                            tables agc2/interpolated_gain_curve.h:104-146
// The adaptive digital controller (RNN VAD, speech / noise level estimators, saturation
// protector) is SURVEY.md 8(f)-1 and not built: resolve_config rejects it.
#pragma once

#include "wap_dev.cuh"
#include "wap_state.h"

namespace wap {

constexpr int kAgc2SubFrames = 20;                               // kSubFramesInFrame
constexpr float kAgc2MaxInputLevelLinear = 36766.300710566735f;  // kMaxInputLevelLinear

// Under-approximating piece-wise linear gain curve: knots x, slopes m, offsets q.
WAP_DEVCONST float kAgc2CurveX[32] = {
    30057.296875f,    30148.986328125f, 30240.67578125f,  30424.052734375f, 30607.4296875f,   30790.806640625f,
    30974.18359375f,  31157.560546875f, 31340.939453125f, 31524.31640625f,  31707.693359375f, 31891.0703125f,
    32074.447265625f, 32257.82421875f,  32441.201171875f, 32624.580078125f, 32807.95703125f,  32991.33203125f,
    33174.7109375f,   33358.08984375f,  33541.46484375f,  33724.84375f,     33819.53515625f,  34009.5390625f,
    34200.05859375f,  34389.81640625f,  34674.48828125f,  35054.375f,       35434.86328125f,  35814.81640625f,
    36195.16796875f,  36575.03125f};
WAP_DEVCONST float kAgc2CurveM[32] = {
    -3.515235675877192989e-07f, -1.050251626111275982e-06f, -2.085213736791047268e-06f, -3.443004743530764244e-06f,
    -4.773849468620028347e-06f, -6.077375928725814447e-06f, -7.353257842623861507e-06f, -8.601219633419532329e-06f,
    -9.821013009059242904e-06f, -1.101243378798244521e-05f, -1.217532644659513608e-05f, -1.330956911260727793e-05f,
    -1.441507538402220234e-05f, -1.549179251014720649e-05f, -1.653970684856176376e-05f, -1.755882840370759368e-05f,
    -1.854918446042574942e-05f, -1.951086778717581183e-05f, -2.044398024736437947e-05f, -2.1348627342376858e-05f,
    -2.222496914328075945e-05f, -2.265374678245279938e-05f, -2.242570917587727308e-05f, -2.220122041762806475e-05f,
    -2.19802095671184361e-05f,  -2.176260204578284174e-05f, -2.133731686626560986e-05f, -2.092481918225530535e-05f,
    -2.052459603874012828e-05f, -2.013615448959171772e-05f, -1.975903069251216948e-05f, -1.939277899509761482e-05f};
WAP_DEVCONST float kAgc2CurveQ[32] = {
    1.010565876960754395f, 1.031631827354431152f, 1.062929749488830566f, 1.104239225387573242f,
    1.144973039627075195f, 1.185109615325927734f, 1.224629044532775879f, 1.263512492179870605f,
    1.301741957664489746f, 1.339300632476806641f, 1.376173257827758789f, 1.412345528602600098f,
    1.447803974151611328f, 1.482536554336547852f, 1.516532182693481445f, 1.549780607223510742f,
    1.582272171974182129f, 1.613999366760253906f, 1.644955039024353027f, 1.675132393836975098f,
    1.704526185989379883f, 1.718986630439758301f, 1.711274504661560059f, 1.703639745712280273f,
    1.696081161499023438f, 1.688597679138183594f, 1.673851132392883301f, 1.659391283988952637f,
    1.645209431648254395f, 1.631297469139099121f, 1.617647409439086914f, 1.604251742362976074f};

WAP_DEV float agc2_lookup_gain(float input_level) {
  if (input_level <= kAgc2CurveX[0]) return 1.0f;                              // identity region
  if (input_level >= kAgc2MaxInputLevelLinear) return 32768.f / input_level;   // saturation region
  int lb = 0;  // std::lower_bound: first knot that is not less than the level
  while (lb < 32 && kAgc2CurveX[lb] < input_level) ++lb;
  const int index = lb - 1;
  return kAgc2CurveM[index] * input_level + kAgc2CurveQ[index];
}

// One frame (flen = 160 or 480 samples in shared memory, in place).  `fac`: flen-float scratch.
WAP_DEV void agc2_process(Agc2State& st, const EngineConfig& cfg, float* frame, int flen, float* fac) {
  const int lane = lane_id();
  const int sub = flen / kAgc2SubFrames;
  __syncwarp();
  // GainApplier::ApplyGain -> ApplyGainWithRamping (gain_applier.cc:41-74,88-97): untouched when the
  // gain is constant and so close to one that int16 samples cannot change, a plain multiply when
  // constant, else a linear ramp from the last gain (a serial chain of additions: lane 0).
  const float g_last = st.gain_last, g = st.gain_current;
  if (st.reset_limiter) {  // SetFixedGainDb changed the gain: Limiter::Reset()
    __syncwarp();
    if (lane == 0) { st.filter_state_level = 0.f; st.reset_limiter = 0; }
    __syncwarp();
  }
  if (g_last != g) {
    const float increment = (g - g_last) * (1.f / flen);
    if (lane == 0) {
      float gain = g_last;
      for (int i = 0; i < flen; ++i) { fac[i] = gain; gain += increment; }
    }
    __syncwarp();
    for (int i = lane; i < flen; i += 32) frame[i] *= fac[i];
    __syncwarp();
    if (lane == 0) st.gain_last = g;
    __syncwarp();
  } else if (!(1.f - 1.f / 32767.f <= g && g <= 1.f + 1.f / 32767.f)) {
    for (int i = lane; i < flen; i += 32) frame[i] *= g;
    __syncwarp();
  }
  // FixedDigitalLevelEstimator::ComputeLevel: max envelope per sub-frame, increases moved one
  // sub-frame earlier, instant attack / slow decay smoothing (a 20-step recurrence).
  float env = 0.f;
  if (lane < kAgc2SubFrames)
    for (int j = 0; j < sub; ++j) env = fmaxr(env, fabsf(frame[lane * sub + j]));
  const float next = __shfl_down_sync(WAP_FULL, env, 1);
  if (lane < kAgc2SubFrames - 1 && env < next) env = next;
  float level = st.filter_state_level;
  float mine = 0.f;
  for (int sf = 0; sf < kAgc2SubFrames; ++sf) {
    const float v = __shfl_sync(WAP_FULL, env, sf);
    float out;
    if (v > level) out = v * (1 - 0.0f) + level * 0.0f;                    // kAttackFilterConstant = 0
    else out = v * (1 - 0.9971259f) + level * 0.9971259f;                  // kDecayFilterConstant
    level = out;
    if (lane == sf + 1) mine = out;   // lane l keeps scaling-factor input l-1
  }
  // scaling_factors_[0] = last frame's final factor, [1..20] = curve look-ups
  float factor = (lane == 0) ? st.last_scaling_factor : agc2_lookup_gain(mine);
  if (lane > kAgc2SubFrames) factor = 0.f;
  const float f_next = __shfl_down_sync(WAP_FULL, factor, 1);
  const float f0 = __shfl_sync(WAP_FULL, factor, 0), f1 = __shfl_sync(WAP_FULL, factor, 1);
  const bool is_attack = f0 > f1;
  // ComputePerSampleSubframeFactors: lane l < 20 fills sub-frame l
  if (lane < kAgc2SubFrames) {
    if (lane == 0 && is_attack) {
      for (int i = 0; i < sub; ++i) {
        const float t = (float)i / sub;
        // std::pow(1.f - t, 8.f): the power in double, rounded once -- for the arguments 1 - i / sub
        // of every sub-frame length up to 48 it is glibc's powf bit for bit (checked; CUDA's powf is
        // several ulp off)
        const double b = (double)(1.f - t), b2 = b * b, b4 = b2 * b2;
        fac[i] = (float)(b4 * b4) * (factor - f_next) + f_next;
      }
    } else {
      const float diff = (f_next - factor) / sub;
      for (int j = 0; j < sub; ++j) fac[lane * sub + j] = factor + diff * j;
    }
  }
  __syncwarp();
  // ScaleSamples
  for (int i = lane; i < flen; i += 32) frame[i] = clampr(frame[i] * fac[i], -32768.f, 32767.f);
  const float last = __shfl_sync(WAP_FULL, factor, kAgc2SubFrames);   // scaling_factors_.back()
  if (lane == 0) {
    st.filter_state_level = level;
    st.last_scaling_factor = last;
  }
  __syncwarp();
}

// The same for the C channels of a multi-channel leg (frames[c]: flen samples each): one gain applier
// ramp, one level estimate -- the envelope is the maximum over the channels
// (fixed_digital_level_estimator.cc:62-75) -- and one set of per-sample factors for every channel.
WAP_DEV void agc2_process_channels(Agc2State& st, const EngineConfig& cfg, float* const* frames, int C, int flen, float* fac) {
  const int lane = lane_id();
  const int sub = flen / kAgc2SubFrames;
  __syncwarp();
  const float g_last = st.gain_last, g = st.gain_current;
  if (st.reset_limiter) {
    __syncwarp();
    if (lane == 0) { st.filter_state_level = 0.f; st.reset_limiter = 0; }
    __syncwarp();
  }
  if (g_last != g) {
    const float increment = (g - g_last) * (1.f / flen);
    if (lane == 0) {
      float gain = g_last;
      for (int i = 0; i < flen; ++i) { fac[i] = gain; gain += increment; }
    }
    __syncwarp();
    for (int c = 0; c < C; ++c)
      for (int i = lane; i < flen; i += 32) frames[c][i] *= fac[i];
    __syncwarp();
    if (lane == 0) st.gain_last = g;
    __syncwarp();
  } else if (!(1.f - 1.f / 32767.f <= g && g <= 1.f + 1.f / 32767.f)) {
    for (int c = 0; c < C; ++c)
      for (int i = lane; i < flen; i += 32) frames[c][i] *= g;
    __syncwarp();
  }
  float env = 0.f;
  if (lane < kAgc2SubFrames)
    for (int c = 0; c < C; ++c)
      for (int j = 0; j < sub; ++j) env = fmaxr(env, fabsf(frames[c][lane * sub + j]));
  const float next = __shfl_down_sync(WAP_FULL, env, 1);
  if (lane < kAgc2SubFrames - 1 && env < next) env = next;
  float level = st.filter_state_level;
  float mine = 0.f;
  for (int sf = 0; sf < kAgc2SubFrames; ++sf) {
    const float v = __shfl_sync(WAP_FULL, env, sf);
    float out;
    if (v > level) out = v * (1 - 0.0f) + level * 0.0f;
    else out = v * (1 - 0.9971259f) + level * 0.9971259f;
    level = out;
    if (lane == sf + 1) mine = out;
  }
  float factor = (lane == 0) ? st.last_scaling_factor : agc2_lookup_gain(mine);
  if (lane > kAgc2SubFrames) factor = 0.f;
  const float f_next = __shfl_down_sync(WAP_FULL, factor, 1);
  const float f0 = __shfl_sync(WAP_FULL, factor, 0), f1 = __shfl_sync(WAP_FULL, factor, 1);
  const bool is_attack = f0 > f1;
  if (lane < kAgc2SubFrames) {
    if (lane == 0 && is_attack) {
      for (int i = 0; i < sub; ++i) {
        const float t = (float)i / sub;
        const double b = (double)(1.f - t), b2 = b * b, b4 = b2 * b2;
        fac[i] = (float)(b4 * b4) * (factor - f_next) + f_next;
      }
    } else {
      const float diff = (f_next - factor) / sub;
      for (int j = 0; j < sub; ++j) fac[lane * sub + j] = factor + diff * j;
    }
  }
  __syncwarp();
  for (int c = 0; c < C; ++c)
    for (int i = lane; i < flen; i += 32) frames[c][i] = clampr(frames[c][i] * fac[i], -32768.f, 32767.f);
  const float last = __shfl_sync(WAP_FULL, factor, kAgc2SubFrames);
  if (lane == 0) {
    st.filter_state_level = level;
    st.last_scaling_factor = last;
  }
  __syncwarp();
}

}  // namespace wap
