// Host-side construction of a freshly initialised StreamState (the values the
// reference's constructors / Reset() methods establish; citations per block).
#pragma once
#include <string.h>

#include "wap_state.h"

namespace wap {

inline void init_ns_state(NsState& s) {
  memset(&s, 0, sizeof(s));
  for (int i = 0; i < kNsBinsPad; ++i) {
    s.prev_analysis_spectrum[i] = 1.f;  // noise_suppressor.cc:252
    s.wiener[i] = 1.f;                  // wiener_filter.cc:26
    s.avg_log_lrt[i] = 0.5f;            // signal_model.cc:23
  }
  for (int i = 0; i < 3 * kNsBins + 1; ++i) {
    s.q_density[i] = 0.3f;              // quantile_noise_estimator.cc:26-27
    s.q_log_quantile[i] = 8.f;
  }
  // counter_[i] = floor(200 * (i + 1) / 3) (quantile_noise_estimator.cc:29-32)
  s.q_counter[0] = 66; s.q_counter[1] = 133; s.q_counter[2] = 200;
  s.q_num_updates = 1;
  s.num_analyzed_frames = -1;
  s.histogram_analysis_counter = 500;
  s.prior_speech_prob = 0.5f;
  s.lrt = s.spectral_flatness = s.spectral_diff = 0.5f;
  s.prior_lrt = 0.5f;
  s.prior_flatness_threshold = 0.5f;
  s.prior_template_diff_threshold = 0.5f;
  s.prior_lrt_weighting = 1.f;
}

void init_aec3_state(Aec3State& a);  // dsp_aec3 host part (wap_engine.cu)

inline void init_stream_state(StreamState& st) {
  memset(&st, 0, sizeof(st));
  init_ns_state(st.ns);
  init_aec3_state(st.aec);
}

}  // namespace wap
