"""ctypes binding of libwap_b200.so -- the C ABI in include/wap_audio_processing.h.

Python is plumbing only (tests, bench): the product is the shared library.
`load()` binds the CUDA build; it raises if the library is missing -- there is
no CPU fallback.  (The test suite may pass the path of the emulator build of
the same sources, tests/emu/_build/libwap_emu.so, to exercise the kernel
source on a GPU-less box; that library is test infrastructure.)
"""
import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
DEFAULT_LIB = os.path.normpath(os.path.join(_HERE, "..", "libwap_b200.so"))

NS_LOW, NS_MODERATE, NS_HIGH, NS_VERY_HIGH = 0, 1, 2, 3
ERR_NONE = 0
ERRORS = {0: "None", 1: "NullPointer", 2: "Internal", 3: "BadSampleRate", 4: "BadNumberChannels",
          5: "BadStreamParameter", 6: "BadDataLength", 7: "UnsupportedConfig"}


class WapStreamConfig(C.Structure):
    _fields_ = [("sample_rate_hz", C.c_int), ("num_channels", C.c_int32)]


class WapConfig(C.Structure):
    _fields_ = [
        ("pipeline_maximum_internal_processing_rate", C.c_int),
        ("pipeline_multi_channel_render", C.c_bool),
        ("pipeline_multi_channel_capture", C.c_bool),
        ("pipeline_capture_downmix_method", C.c_int32),
        ("pre_amplifier_enabled", C.c_bool),
        ("pre_amplifier_fixed_gain_factor", C.c_float),
        ("capture_level_adjustment_enabled", C.c_bool),
        ("capture_level_adjustment_pre_gain_factor", C.c_float),
        ("capture_level_adjustment_post_gain_factor", C.c_float),
        ("analog_mic_gain_emulation_enabled", C.c_bool),
        ("analog_mic_gain_emulation_initial_level", C.c_int),
        ("high_pass_filter_enabled", C.c_bool),
        ("high_pass_filter_apply_in_full_band", C.c_bool),
        ("echo_canceller_enabled", C.c_bool),
        ("echo_canceller_enforce_high_pass_filtering", C.c_bool),
        ("noise_suppression_enabled", C.c_bool),
        ("noise_suppression_level", C.c_int32),
        ("noise_suppression_analyze_linear_aec_output_when_available", C.c_bool),
        ("gain_controller2_enabled", C.c_bool),
        ("gain_controller2_fixed_digital_gain_db", C.c_float),
        ("gain_controller2_adaptive_digital_enabled", C.c_bool),
        ("gain_controller2_adaptive_digital_headroom_db", C.c_float),
        ("gain_controller2_adaptive_digital_max_gain_db", C.c_float),
        ("gain_controller2_adaptive_digital_initial_gain_db", C.c_float),
        ("gain_controller2_adaptive_digital_max_gain_change_db_per_second", C.c_float),
        ("gain_controller2_adaptive_digital_max_output_noise_level_dbfs", C.c_float),
        ("gain_controller2_input_volume_controller_enabled", C.c_bool),
    ]


class WapStats(C.Structure):
    _fields_ = [
        ("has_echo_return_loss", C.c_bool), ("echo_return_loss", C.c_double),
        ("has_echo_return_loss_enhancement", C.c_bool), ("echo_return_loss_enhancement", C.c_double),
        ("has_divergent_filter_fraction", C.c_bool), ("divergent_filter_fraction", C.c_double),
        ("has_delay_median_ms", C.c_bool), ("delay_median_ms", C.c_int32),
        ("has_delay_standard_deviation_ms", C.c_bool), ("delay_standard_deviation_ms", C.c_int32),
        ("has_residual_echo_likelihood", C.c_bool), ("residual_echo_likelihood", C.c_double),
        ("has_residual_echo_likelihood_recent_max", C.c_bool), ("residual_echo_likelihood_recent_max", C.c_double),
        ("has_delay_ms", C.c_bool), ("delay_ms", C.c_int32),
    ]


def _struct(name, fields):
    return type(name, (C.Structure,), {"_fields_": fields})


_f, _i, _b = C.c_float, C.c_int32, C.c_bool
WapEc3MaskingThresholds = _struct("WapEc3MaskingThresholds", [("enr_transparent", _f), ("enr_suppress", _f), ("emr_transparent", _f)])
WapEc3Tuning = _struct("WapEc3Tuning", [("mask_lf", WapEc3MaskingThresholds), ("mask_hf", WapEc3MaskingThresholds),
                                        ("max_inc_factor", _f), ("max_dec_factor_lf", _f)])
WapEc3AlignmentMixing = _struct("WapEc3AlignmentMixing", [("downmix", _b), ("adaptive_selection", _b),
                                                          ("activity_power_threshold", _f), ("prefer_first_two_channels", _b)])
WapEc3RefinedConfiguration = _struct("WapEc3RefinedConfiguration", [
    ("length_blocks", _i), ("leakage_converged", _f), ("leakage_diverged", _f), ("error_floor", _f), ("error_ceil", _f),
    ("noise_gate", _f)])
WapEc3CoarseConfiguration = _struct("WapEc3CoarseConfiguration", [("length_blocks", _i), ("rate", _f), ("noise_gate", _f)])
WapEc3SubbandRegion = _struct("WapEc3SubbandRegion", [("low", _i), ("high", _i)])


class WapEchoCanceller3Config(C.Structure):
    """Mirror of include/wap_audio_processing.h: WapEchoCanceller3Config (= webrtc::EchoCanceller3Config)."""
    _fields_ = [
        ("buffering", _struct("Buffering", [("excess_render_detection_interval_blocks", _i), ("max_allowed_excess_render_blocks", _i)])),
        ("delay", _struct("Delay", [
            ("default_delay", _i), ("down_sampling_factor", _i), ("num_filters", _i), ("delay_headroom_samples", _i),
            ("hysteresis_limit_blocks", _i), ("fixed_capture_delay_samples", _i), ("delay_estimate_smoothing", _f),
            ("delay_estimate_smoothing_delay_found", _f), ("delay_candidate_detection_threshold", _f),
            ("delay_selection_thresholds", _struct("DelaySelectionThresholds", [("initial", _i), ("converged", _i)])),
            ("use_external_delay_estimator", _b), ("log_warning_on_delay_changes", _b),
            ("render_alignment_mixing", WapEc3AlignmentMixing), ("capture_alignment_mixing", WapEc3AlignmentMixing),
            ("detect_pre_echo", _b)])),
        ("filter", _struct("Filter", [
            ("refined", WapEc3RefinedConfiguration), ("coarse", WapEc3CoarseConfiguration),
            ("refined_initial", WapEc3RefinedConfiguration), ("coarse_initial", WapEc3CoarseConfiguration),
            ("config_change_duration_blocks", _i), ("initial_state_seconds", _f), ("coarse_reset_hangover_blocks", _i),
            ("conservative_initial_phase", _b), ("enable_coarse_filter_output_usage", _b), ("use_linear_filter", _b),
            ("high_pass_filter_echo_reference", _b), ("export_linear_aec_output", _b)])),
        ("erle", _struct("Erle", [("min", _f), ("max_l", _f), ("max_h", _f), ("onset_detection", _b), ("num_sections", _i),
                                  ("clamp_quality_estimate_to_zero", _b), ("clamp_quality_estimate_to_one", _b)])),
        ("ep_strength", _struct("EpStrength", [
            ("default_gain", _f), ("default_len", _f), ("nearend_len", _f), ("echo_can_saturate", _b), ("bounded_erl", _b),
            ("erle_onset_compensation_in_dominant_nearend", _b), ("use_conservative_tail_frequency_response", _b)])),
        ("echo_audibility", _struct("EchoAudibility", [
            ("low_render_limit", _f), ("normal_render_limit", _f), ("floor_power", _f), ("audibility_threshold_lf", _f),
            ("audibility_threshold_mf", _f), ("audibility_threshold_hf", _f), ("use_stationarity_properties", _b),
            ("use_stationarity_properties_at_init", _b)])),
        ("render_levels", _struct("RenderLevels", [("active_render_limit", _f), ("poor_excitation_render_limit", _f),
                                                   ("poor_excitation_render_limit_ds8", _f), ("render_power_gain_db", _f)])),
        ("echo_removal_control", _struct("EchoRemovalControl", [("has_clock_drift", _b), ("linear_and_stable_echo_path", _b)])),
        ("echo_model", _struct("EchoModel", [
            ("noise_floor_hold", _i), ("min_noise_floor_power", _f), ("stationary_gate_slope", _f), ("noise_gate_power", _f),
            ("noise_gate_slope", _f), ("render_pre_window_size", _i), ("render_post_window_size", _i),
            ("model_reverb_in_nonlinear_mode", _b)])),
        ("comfort_noise", _struct("ComfortNoise", [("noise_floor_dbfs", _f)])),
        ("suppressor", _struct("Suppressor", [
            ("nearend_average_blocks", _i), ("normal_tuning", WapEc3Tuning), ("nearend_tuning", WapEc3Tuning),
            ("lf_smoothing_during_initial_phase", _b), ("last_permanent_lf_smoothing_band", _i), ("last_lf_smoothing_band", _i),
            ("last_lf_band", _i), ("first_hf_band", _i),
            ("dominant_nearend_detection", _struct("DominantNearendDetection", [
                ("enr_threshold", _f), ("enr_exit_threshold", _f), ("snr_threshold", _f), ("hold_duration", _i),
                ("trigger_threshold", _i), ("use_during_initial_phase", _b), ("use_unbounded_echo_spectrum", _b)])),
            ("subband_nearend_detection", _struct("SubbandNearendDetection", [
                ("nearend_average_blocks", _i), ("subband1", WapEc3SubbandRegion), ("subband2", WapEc3SubbandRegion),
                ("nearend_threshold", _f), ("snr_threshold", _f)])),
            ("use_subband_nearend_detection", _b),
            ("high_bands_suppression", _struct("HighBandsSuppression", [
                ("enr_threshold", _f), ("max_gain_during_echo", _f), ("anti_howling_activation_threshold", _f),
                ("anti_howling_gain", _f)])),
            ("high_frequency_suppression", _struct("HighFrequencySuppression", [("limiting_gain_band", _i), ("bands_in_limiting_gain", _i)])),
            ("floor_first_increase", _f), ("conservative_hf_suppression", _b)])),
        ("multi_channel", _struct("MultiChannel", [
            ("detect_stereo_content", _b), ("stereo_detection_threshold", _f),
            ("stereo_detection_timeout_threshold_seconds", _i), ("stereo_detection_hysteresis_seconds", _f)])),
    ]


def ec3_set(cfg, path, value):
    """cfg.<a.b.c> = value with the reference's member names, e.g. "filter.refined.length_blocks"."""
    obj = cfg
    parts = path.split(".")
    for p in parts[:-1]:
        obj = getattr(obj, p)
    cur = getattr(obj, parts[-1])
    setattr(obj, parts[-1], type(cur)(value) if not isinstance(cur, float) else float(value))


def ec3_get(cfg, path):
    obj = cfg
    for p in path.split("."):
        obj = getattr(obj, p)
    return obj


def make_aec3_config(lib, overrides=None, multichannel=False):
    c = lib.wap_echo_canceller3_config_default_multichannel() if multichannel else lib.wap_echo_canceller3_config_default()
    for k, v in (overrides or {}).items():
        ec3_set(c, k, v)
    return c


# Every symbol include/wap_audio_processing.h declares.
EXPORTS = [
    "wap_create", "wap_create_with_config", "wap_destroy", "wap_config_default", "wap_get_config",
    "wap_apply_config", "wap_initialize", "wap_set_capture_output_used", "wap_set_capture_pre_gain",
    "wap_set_capture_post_gain", "wap_set_capture_fixed_post_gain", "wap_set_playout_volume",
    "wap_set_playout_audio_device", "wap_set_stream_analog_level", "wap_recommended_stream_analog_level",
    "wap_set_stream_delay_ms", "wap_stream_delay_ms", "wap_process_stream_i16", "wap_process_stream_f32",
    "wap_process_reverse_stream_i16", "wap_process_reverse_stream_f32", "wap_get_statistics",
    "wap_engine_create", "wap_engine_destroy", "wap_engine_create_streams", "wap_engine_state_bytes_per_stream",
    "wap_engine_algorithmic_bytes_per_frame", "wap_process_streams", "wap_process_streams_device",
    "wap_engine_synchronize", "wap_engine_cuda_stream", "wap_engine_launch_count", "wap_engine_uses_runtime_aec3_parameters", "wap_version",
    "wap_streams_set_delay_ms", "wap_engine_enable_kernel_timing", "wap_engine_read_kernel_timing", "wap_engine_algorithmic_bytes_per_kernel",
    "wap_engine_set_pipeline_chunks", "wap_stream_state_bytes", "wap_stream_export_state", "wap_stream_import_state", "wap_stream_read_taps",
    "wap_echo_canceller3_config_default", "wap_echo_canceller3_config_default_multichannel", "wap_echo_canceller3_config_sizeof",
    "wap_echo_canceller3_config_validate", "wap_echo_canceller3_config_supported", "wap_create_with_aec3_config",
    "wap_engine_create_with_aec3_config", "wap_engine_create_with_formats", "wap_stream_migrate",
    "wap_engine_enable_echo_detector",
]

_libs = {}


def load(path=None):
    path = path or os.environ.get("WAP_B200_LIB") or DEFAULT_LIB
    if path in _libs:
        return _libs[path]
    if not os.path.exists(path):
        raise RuntimeError("%s not found: build it with webrtc-audio-processing_b200/build.py "
                           "(CUDA extension required; there is no CPU fallback)" % path)
    L = C.CDLL(path)
    vp, i32, sc, cfg = C.c_void_p, C.c_int32, WapStreamConfig, WapConfig
    L.wap_config_default.restype = cfg
    L.wap_create.restype = vp
    L.wap_create_with_config.restype = vp
    L.wap_create_with_config.argtypes = [cfg]
    L.wap_destroy.argtypes = [vp]
    L.wap_get_config.argtypes = [vp, C.POINTER(cfg)]
    L.wap_apply_config.argtypes = [vp, cfg]
    L.wap_initialize.argtypes = [vp, sc, sc, sc, sc]
    L.wap_set_stream_delay_ms.argtypes = [vp, C.c_int]
    L.wap_stream_delay_ms.argtypes = [vp]
    L.wap_set_capture_output_used.argtypes = [vp, C.c_bool]
    L.wap_process_stream_i16.argtypes = [vp, vp, i32, sc, sc, vp, i32]
    L.wap_process_reverse_stream_i16.argtypes = [vp, vp, i32, sc, sc, vp, i32]
    L.wap_process_stream_f32.argtypes = [vp, vp, sc, sc, vp]
    L.wap_process_reverse_stream_f32.argtypes = [vp, vp, sc, sc, vp]
    L.wap_get_statistics.argtypes = [vp, C.POINTER(WapStats)]
    L.wap_engine_create.restype = vp
    L.wap_engine_create.argtypes = [C.c_int, i32, cfg, sc]
    L.wap_engine_destroy.argtypes = [vp]
    L.wap_engine_create_streams.argtypes = [vp, i32, vp]
    L.wap_engine_state_bytes_per_stream.restype = C.c_size_t
    L.wap_engine_state_bytes_per_stream.argtypes = [vp]
    L.wap_engine_algorithmic_bytes_per_frame.restype = C.c_double
    L.wap_engine_algorithmic_bytes_per_frame.argtypes = [vp]
    L.wap_process_streams.argtypes = [vp, i32, vp, vp, vp, i32, vp]
    L.wap_process_streams_device.argtypes = [vp, vp, i32, vp, vp, vp, i32]
    L.wap_engine_synchronize.argtypes = [vp]
    L.wap_engine_cuda_stream.restype = vp
    L.wap_engine_cuda_stream.argtypes = [vp]
    L.wap_engine_launch_count.restype = C.c_int64
    L.wap_engine_launch_count.argtypes = [vp]
    L.wap_engine_uses_runtime_aec3_parameters.restype = C.c_int32
    L.wap_engine_uses_runtime_aec3_parameters.argtypes = [vp]
    L.wap_streams_set_delay_ms.argtypes = [vp, i32, C.c_int]
    L.wap_engine_enable_kernel_timing.argtypes = [vp, C.c_bool]
    L.wap_engine_set_pipeline_chunks.argtypes = [vp, i32]
    L.wap_set_capture_pre_gain.argtypes = [vp, C.c_float]
    L.wap_stream_state_bytes.restype = C.c_size_t
    L.wap_stream_state_bytes.argtypes = [vp]
    L.wap_stream_export_state.argtypes = [vp, vp, C.c_size_t]
    L.wap_stream_import_state.argtypes = [vp, vp, C.c_size_t]
    L.wap_set_capture_post_gain.argtypes = [vp, C.c_float]
    L.wap_set_playout_volume.argtypes = [vp, C.c_int]
    L.wap_set_capture_fixed_post_gain.argtypes = [vp, C.c_float]
    L.wap_engine_read_kernel_timing.restype = C.c_int64
    L.wap_engine_read_kernel_timing.argtypes = [vp, C.POINTER(C.c_double)]
    L.wap_engine_algorithmic_bytes_per_kernel.argtypes = [vp, C.POINTER(C.c_double)]
    ec3 = WapEchoCanceller3Config
    L.wap_echo_canceller3_config_default.restype = ec3
    L.wap_echo_canceller3_config_default_multichannel.restype = ec3
    L.wap_echo_canceller3_config_sizeof.restype = C.c_size_t
    L.wap_echo_canceller3_config_validate.restype = C.c_bool
    L.wap_echo_canceller3_config_validate.argtypes = [C.POINTER(ec3)]
    L.wap_echo_canceller3_config_supported.argtypes = [C.POINTER(ec3)]
    L.wap_create_with_aec3_config.restype = vp
    L.wap_create_with_aec3_config.argtypes = [cfg, C.POINTER(ec3), C.POINTER(ec3)]
    L.wap_engine_create_with_aec3_config.restype = vp
    L.wap_engine_create_with_aec3_config.argtypes = [C.c_int, i32, cfg, sc, C.POINTER(ec3), C.POINTER(ec3)]
    L.wap_stream_migrate.restype = C.c_int
    L.wap_stream_migrate.argtypes = [vp, vp]
    L.wap_engine_enable_echo_detector.argtypes = [vp]
    L.wap_engine_create_with_formats.restype = vp
    L.wap_engine_create_with_formats.argtypes = [C.c_int, i32, cfg, sc, sc, sc, C.POINTER(ec3), C.POINTER(ec3)]
    L.wap_version.restype = C.c_char_p
    L.wap_version.argtypes = []
    L.wap_config_default.argtypes = []
    L.wap_create.argtypes = []
    L.wap_echo_canceller3_config_default.argtypes = []
    L.wap_echo_canceller3_config_default_multichannel.argtypes = []
    L.wap_echo_canceller3_config_sizeof.argtypes = []
    L.wap_set_playout_audio_device.argtypes = [vp, C.c_int, C.c_int]
    L.wap_set_stream_analog_level.argtypes = [vp, C.c_int]
    L.wap_recommended_stream_analog_level.argtypes = [vp]
    L.wap_stream_read_taps.argtypes = [vp, C.POINTER(WapStageTaps)]
    # a handle passed without argtypes would be truncated to a C int: every export must be declared
    missing = [n for n in EXPORTS if getattr(L, n).argtypes is None]
    if missing:
        raise RuntimeError("wap_b200: no argtypes for " + ", ".join(missing))
    _libs[path] = L
    return L


def make_config(lib, aec=True, ns=True, ns_level=NS_MODERATE, max_rate=48000, hpf=False, agc2=False,
                agc2_fixed_gain_db=0.0, pre_amp=None, pre_gain=None, post_gain=None, mc_render=False, mc_capture=False,
                downmix=0):
    c = lib.wap_config_default()
    c.pipeline_capture_downmix_method = int(downmix)   # 0: AverageChannels, 1: UseFirstChannel
    c.pipeline_multi_channel_render = bool(mc_render)
    c.pipeline_multi_channel_capture = bool(mc_capture)
    if pre_amp is not None:
        c.pre_amplifier_enabled = True
        c.pre_amplifier_fixed_gain_factor = float(pre_amp)
    if pre_gain is not None or post_gain is not None:
        c.capture_level_adjustment_enabled = True
        c.capture_level_adjustment_pre_gain_factor = float(1.0 if pre_gain is None else pre_gain)
        c.capture_level_adjustment_post_gain_factor = float(1.0 if post_gain is None else post_gain)
    c.gain_controller2_enabled = bool(agc2)
    c.gain_controller2_fixed_digital_gain_db = float(agc2_fixed_gain_db)
    c.echo_canceller_enabled = bool(aec)
    c.noise_suppression_enabled = bool(ns)
    c.noise_suppression_level = int(ns_level)
    c.pipeline_maximum_internal_processing_rate = int(max_rate)
    c.high_pass_filter_enabled = bool(hpf)
    return c


def _ptr(a):
    return a.ctypes.data_as(C.c_void_p) if a is not None else None


class WapStageTaps(C.Structure):
    _fields_ = [("aec3_erle", C.c_float * 65), ("aec3_erle_onset_compensated", C.c_float * 65),
                ("aec3_erl", C.c_float * 65), ("aec3_erl_time_domain", C.c_float),
                ("aec3_fullband_erle_log2", C.c_float), ("aec3_suppressor_gain", C.c_float * 65),
                ("aec3_N2", C.c_float * 65), ("aec3_refined_gain_H_error", C.c_float * 65),
                ("aec3_filter_delay", C.c_int32), ("aec3_min_direct_path_filter_delay", C.c_int32),
                ("aec3_render_delay_controller_buffer_delay", C.c_int32),
                ("aec3_usable_linear_estimate", C.c_int32), ("aec3_transparent_mode", C.c_int32),
                ("aec3_initial_state", C.c_int32), ("aec3_echo_saturation", C.c_int32),
                ("aec3_capture_saturation", C.c_int32), ("aec3_dominant_nearend", C.c_int32),
                ("ns_noise_spectrum", C.c_float * 129), ("ns_filter", C.c_float * 129),
                ("ns_speech_probability", C.c_float * 129), ("ns_prior_speech_probability", C.c_float)]


class Engine:
    """Batched engine: `n` call legs of one config class on one GPU."""

    def __init__(self, n_streams, rate=16000, channels=1, lib=None, device=0, capacity=None, aec3=None,
                 aec3_multichannel=None, out_format=None, render_format=None, echo_detector=False, **cfg):
        """aec3: None (default EchoCanceller3Config), a dict of overrides keyed by the reference's member
        paths ("filter.refined.length_blocks": 10, ...) or a WapEchoCanceller3Config.  aec3_multichannel: the
        same for the multichannel config (overrides apply to CreateDefaultMultichannelConfig)."""
        self.lib = lib or load()
        self.rate, self.channels, self.n = rate, channels, n_streams
        self.frame = rate // 100 * channels
        self.config = make_config(self.lib, **cfg)
        # out_format / render_format: (rate, channels) of the output / render stream when they differ from
        # the capture input's (wap_engine_create_with_formats)
        self.out_rate, self.out_channels = out_format or (rate, channels)
        self.render_rate, self.render_channels = render_format or (rate, channels)
        self.out_frame = self.out_rate // 100 * self.out_channels
        if out_format is not None or render_format is not None:
            def conv(c, mc):
                if c is None:
                    return None
                return c if isinstance(c, WapEchoCanceller3Config) else make_aec3_config(self.lib, c, multichannel=mc)
            self.aec3, self.aec3_mc = conv(aec3, False), conv(aec3_multichannel, True)
            self.h = self.lib.wap_engine_create_with_formats(
                device, capacity or n_streams, self.config, WapStreamConfig(rate, channels),
                WapStreamConfig(self.out_rate, self.out_channels), WapStreamConfig(self.render_rate, self.render_channels),
                C.byref(self.aec3) if self.aec3 is not None else None,
                C.byref(self.aec3_mc) if self.aec3_mc is not None else None)
        elif aec3 is None and aec3_multichannel is None:
            self.h = self.lib.wap_engine_create(device, capacity or n_streams, self.config,
                                                WapStreamConfig(rate, channels))
        else:
            def conv(c, mc):
                if c is None:
                    return None
                return c if isinstance(c, WapEchoCanceller3Config) else make_aec3_config(self.lib, c, multichannel=mc)
            self.aec3, self.aec3_mc = conv(aec3, False), conv(aec3_multichannel, True)
            if self.aec3 is None:   # BuiltinAudioProcessingBuilder::SetEchoCancellerConfig needs the mono config
                self.aec3 = make_aec3_config(self.lib, {})
            self.h = self.lib.wap_engine_create_with_aec3_config(
                device, capacity or n_streams, self.config, WapStreamConfig(rate, channels), C.byref(self.aec3),
                C.byref(self.aec3_mc) if self.aec3_mc is not None else None)
        if not self.h:
            raise RuntimeError("wap_engine_create failed (no CUDA device or unsupported config)")
        if echo_detector:   # AudioProcessingBuilder::SetEchoDetector(CreateEchoDetector())
            err = self.lib.wap_engine_enable_echo_detector(self.h)
            if err:
                self.lib.wap_engine_destroy(self.h)
                self.h = None
                raise RuntimeError("wap_engine_enable_echo_detector: " + ERRORS.get(err, str(err)))
        self.handles = (C.c_void_p * n_streams)()
        err = self.lib.wap_engine_create_streams(self.h, n_streams, self.handles)
        if err:
            raise RuntimeError("wap_engine_create_streams: " + ERRORS.get(err, str(err)))

    def set_stream_delay_ms(self, ms):
        self.lib.wap_streams_set_delay_ms(self.handles, self.n, ms)

    def set_pipeline_chunks(self, chunks):
        err = self.lib.wap_engine_set_pipeline_chunks(self.h, int(chunks))
        if err:
            raise RuntimeError("wap_engine_set_pipeline_chunks -> WapError %d" % err)

    def taps(self, i=0):
        t = WapStageTaps()
        err = self.lib.wap_stream_read_taps(self.handles[i], C.byref(t))
        if err:
            raise RuntimeError("wap_stream_read_taps: " + ERRORS.get(err, str(err)))
        return t

    def export_state(self, i=0):
        n = self.lib.wap_stream_state_bytes(self.handles[i])
        blob = np.zeros(n, np.uint8)
        err = self.lib.wap_stream_export_state(self.handles[i], blob.ctypes.data_as(C.c_void_p), n)
        if err:
            raise RuntimeError("wap_stream_export_state: " + ERRORS.get(err, str(err)))
        return blob

    def import_state(self, blob, i=0):
        err = self.lib.wap_stream_import_state(self.handles[i], blob.ctypes.data_as(C.c_void_p), blob.size)
        if err:
            raise RuntimeError("wap_stream_import_state: " + ERRORS.get(err, str(err)))

    def set_pre_gain(self, gain, legs=None):
        for i in (range(self.n) if legs is None else legs):
            self.lib.wap_set_capture_pre_gain(self.handles[i], float(gain))

    def set_post_gain(self, gain, legs=None):
        for i in (range(self.n) if legs is None else legs):
            self.lib.wap_set_capture_post_gain(self.handles[i], float(gain))

    def set_fixed_post_gain(self, gain_db, legs=None):
        for i in (range(self.n) if legs is None else legs):
            self.lib.wap_set_capture_fixed_post_gain(self.handles[i], float(gain_db))

    def set_playout_volume(self, volume, legs=None):
        for i in (range(self.n) if legs is None else legs):
            self.lib.wap_set_playout_volume(self.handles[i], int(volume))

    def set_capture_output_used(self, used, legs=None):
        for i in (range(self.n) if legs is None else legs):
            self.lib.wap_set_capture_output_used(self.handles[i], bool(used))

    def process(self, render, capture):
        """render/capture: [n, frame] int16 or float32 (render may be None). Returns out [n, frame]."""
        capture = np.ascontiguousarray(capture)
        fmt = 0 if capture.dtype == np.int16 else 1
        if render is not None:
            render = np.ascontiguousarray(render, dtype=capture.dtype)
        out = np.empty_like(capture) if self.out_frame == self.frame else np.empty((self.n, self.out_frame), capture.dtype)
        err = self.lib.wap_process_streams(self.handles, self.n, _ptr(render), _ptr(capture), _ptr(out), fmt, None)
        if err:
            raise RuntimeError("wap_process_streams: " + ERRORS.get(err, str(err)))
        return out

    def stats(self, i=0):
        s = WapStats()
        self.lib.wap_get_statistics(self.handles[i], C.byref(s))
        return s

    def close(self):
        if self.h:
            for h in self.handles:
                self.lib.wap_destroy(h)
            self.lib.wap_engine_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass
