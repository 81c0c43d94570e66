// Test-harness code (NOT product code): drives the reference's OWN alternate-backend class,
// webrtc::RustAudioProcessing (reference modules/audio_processing/rust_audio_processing.{h,cc},
// compiled unmodified from /root/reference against include/wap_audio_processing.h), through the
// loop of the reference's examples/run-offline.cpp:45-63.  The wap_* symbols it calls are left
// undefined in oracle/_ref/libwap_seam.so and are resolved at load time by whichever
// implementation the test loaded first with RTLD_GLOBAL: libwap_b200.so on the GPU box, the
// emulator build of the same sources on a CPU box.  This is the drop-in proof: the reference's seam
// class, linked against this repo's C ABI, produces the reference's output.
#include <cstdint>
#include <vector>

#include "api/audio/audio_processing.h"
#include "api/make_ref_counted.h"
#include "api/scoped_refptr.h"
#include "modules/audio_processing/rust_audio_processing.h"

using webrtc::AudioProcessing;

extern "C" {

// nframes 10 ms ticks of interleaved int16 audio: ProcessReverseStream, set_stream_delay_ms(0),
// ProcessStream -- run-offline.cpp's loop.  stats_out (optional): [erl, erle, delay_ms] every
// `stats_every` frames.  Returns the first non-zero AudioProcessing error code.
int seam_run_offline_i16(int aec, int ns, int ns_level, int max_rate, int agc2, float agc2_gain_db, int rate,
                         int channels, int nframes, const int16_t* render, const int16_t* capture, int16_t* out,
                         int stats_every, double* stats_out) {
  AudioProcessing::Config c;
  c.echo_canceller.enabled = aec != 0;
  c.noise_suppression.enabled = ns != 0;
  c.noise_suppression.level = static_cast<AudioProcessing::Config::NoiseSuppression::Level>(ns_level);
  c.pipeline.maximum_internal_processing_rate = max_rate;
  c.gain_controller2.enabled = agc2 != 0;
  c.gain_controller2.fixed_digital.gain_db = agc2_gain_db;
  webrtc::scoped_refptr<AudioProcessing> apm = webrtc::make_ref_counted<webrtc::RustAudioProcessing>(c);
  const webrtc::StreamConfig sc(rate, channels);
  const size_t n = (size_t)rate / 100 * channels;
  std::vector<int16_t> render_out(n);
  int err = 0, k = 0;
  for (int f = 0; f < nframes; ++f) {
    int e = 0;
    if (render) e = apm->ProcessReverseStream(render + f * n, sc, sc, render_out.data());
    if (e && !err) err = e;
    apm->set_stream_delay_ms(0);
    e = apm->ProcessStream(capture + f * n, sc, sc, out + f * n);
    if (e && !err) err = e;
    if (stats_out && stats_every > 0 && (f + 1) % stats_every == 0) {
      webrtc::AudioProcessingStats s = apm->GetStatistics();
      stats_out[3 * k + 0] = s.echo_return_loss.value_or(0.0);
      stats_out[3 * k + 1] = s.echo_return_loss_enhancement.value_or(0.0);
      stats_out[3 * k + 2] = s.delay_ms.value_or(-1);
      ++k;
    }
  }
  // the seam's config round trip (RustAudioProcessing::GetConfig -> wap_get_config)
  const AudioProcessing::Config back = apm->GetConfig();
  if (back.noise_suppression.enabled != c.noise_suppression.enabled ||
      back.echo_canceller.enabled != c.echo_canceller.enabled ||
      static_cast<int>(back.noise_suppression.level) != ns_level)
    return -100;
  return err;
}

}  // extern "C"
