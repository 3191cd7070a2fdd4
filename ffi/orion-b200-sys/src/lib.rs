//! Raw declarations of the C ABI in `include/orion_b200.h` (ABI version 1).
//! NOT COMPILED in the build environment of this repository (no Rust toolchain there);
//! kept in lock-step with the header by `tests/test_host_logic.py::test_rust_ffi_declares_every_symbol`.
#![allow(non_camel_case_types)]
use core::ffi::{c_char, c_int, c_void};

pub const ORION_B200_ABI_VERSION: c_int = 1;
pub const ORION_B200_OK: c_int = 0;
pub const ORION_B200_ERR_INVALID: c_int = 1;
pub const ORION_B200_ERR_NO_DEVICE: c_int = 2;
pub const ORION_B200_ERR_CUDA: c_int = 3;
pub const ORION_B200_ERR_ALLOC: c_int = 4;
pub const ORION_B200_ERR_UNSUPPORTED: c_int = 5;
pub const ORION_B200_ERR_INTERNAL: c_int = 6;

pub const ORION_B200_MIX_NONE: i32 = 0;
pub const ORION_B200_MIX_ROTATE: i32 = 1;
pub const ORION_B200_MIX_NCO: i32 = 2;
pub const ORION_B200_FIR_NONE: i32 = 0;
pub const ORION_B200_FIR_DECIM: i32 = 1;
pub const ORION_B200_FIR_IQ: i32 = 2;
pub const ORION_B200_DEMOD_NONE: i32 = 0;
pub const ORION_B200_DEMOD_FM: i32 = 1;
pub const ORION_B200_DEMOD_PM: i32 = 2;
pub const ORION_B200_DEMOD_AM: i32 = 3;
pub const ORION_B200_DEMOD_AM_ABS: i32 = 4;
pub const ORION_B200_DEMOD_SSB: i32 = 5;
pub const ORION_B200_DEMOD_CW: i32 = 6;
pub const ORION_B200_DEMOD_USB: i32 = 7;

#[repr(C)]
pub struct orion_b200_block {
    _private: [u8; 0],
}
#[repr(C)]
pub struct orion_b200_bank {
    _private: [u8; 0],
}

/// `num_complex::Complex32` is `#[repr(C)] { re: f32, im: f32 }`, so `&[Complex32]` can be passed as
/// `*const orion_b200_c32` without a copy.
#[repr(C)]
#[derive(Clone, Copy, Debug, Default)]
pub struct orion_b200_c32 {
    pub re: f32,
    pub im: f32,
}

#[repr(C)]
#[derive(Clone, Copy, Debug, Default)]
pub struct orion_b200_work_report {
    pub in_read: usize,
    pub out_written: usize,
}

#[repr(C)]
pub struct orion_b200_chain_spec {
    pub struct_size: u32,
    pub mix: i32,
    pub mix_freq_hz: f32,
    pub mix_fs: f32,
    pub fir: i32,
    pub taps: *const f32,
    pub ntaps: usize,
    pub decim: usize,
    pub demod: i32,
    pub fs_demod: f32,
    pub p0: f32,
    pub p1: f32,
    pub audio_bw_hz: f32,
    pub translate: i32,
    pub translate_hz: f32,
    pub post_sos: *const f32,
    pub n_post: usize,
}

extern "C" {
    pub fn orion_b200_abi_version() -> c_int;
    pub fn orion_b200_build_info() -> *const c_char;
    pub fn orion_b200_device_count() -> c_int;
    pub fn orion_b200_set_device(ordinal: c_int) -> c_int;
    pub fn orion_b200_status_string(status: c_int) -> *const c_char;
    pub fn orion_b200_host_alloc(ptr: *mut *mut c_void, bytes: usize) -> c_int;
    pub fn orion_b200_host_free(ptr: *mut c_void);

    pub fn orion_b200_fir_lowpass_design(fs: f32, pass_hz: f32, trans_hz: f32, taps: *mut f32, cap: usize) -> usize;
    pub fn orion_b200_kaiser_lowpass_taps(num_taps: usize, cutoff_norm: f32, stopband_db: f32, taps: *mut f32, cap: usize) -> usize;
    pub fn orion_b200_kaiser_transition_norm(num_taps: usize, stopband_db: f32) -> f32;
    pub fn orion_b200_kaiser_num_taps(transition_norm: f32, stopband_db: f32) -> usize;
    pub fn orion_b200_lp_biquad_design(fs: f32, fc: f32, coeffs: *mut f32);
    pub fn orion_b200_dc_pole(fs: f32, cut_hz: f32) -> f32;
    pub fn orion_b200_cw_alpha(fs: f32, env_bw_hz: f32) -> f32;

    pub fn orion_b200_fir_decimator_create(fs: f32, m: usize, cutoff_hz: f32, trans_hz: f32, out: *mut *mut orion_b200_block) -> c_int;
    pub fn orion_b200_fir_decimator_create_taps(taps: *const f32, ntaps: usize, m: usize, out: *mut *mut orion_b200_block) -> c_int;
    pub fn orion_b200_fir_lowpass_iq_create(num_taps: usize, cutoff_norm: f32, stopband_db: f32, out: *mut *mut orion_b200_block) -> c_int;
    pub fn orion_b200_fir_lowpass_iq_create_taps(taps: *const f32, ntaps: usize, out: *mut *mut orion_b200_block) -> c_int;
    pub fn orion_b200_fir_lowpass_iq_filter_aligned(b: *mut orion_b200_block, io: *mut orion_b200_c32, n: usize) -> c_int;
    pub fn orion_b200_half_cosine_mf_taps(sps: usize, taps: *mut f32, cap: usize) -> usize;
    pub fn orion_b200_half_cosine_mf_create(sps: usize, out: *mut *mut orion_b200_block) -> c_int;
    pub fn orion_b200_rotator_create(freq_hz: f32, fs: f32, out: *mut *mut orion_b200_block) -> c_int;
    pub fn orion_b200_rotator_usb_create(freq_hz: f32, fs: f32, out: *mut *mut orion_b200_block) -> c_int;
    pub fn orion_b200_nco_mixer_create(freq_hz: f32, fs: f32, out: *mut *mut orion_b200_block) -> c_int;
    pub fn orion_b200_oscillator_set_freq(b: *mut orion_b200_block, freq_hz: f32, fs: f32) -> c_int;
    pub fn orion_b200_oscillator_reset_phase(b: *mut orion_b200_block) -> c_int;
    pub fn orion_b200_biquad_create(b0: f32, b1: f32, b2: f32, a1: f32, a2: f32, out: *mut *mut orion_b200_block) -> c_int;
    pub fn orion_b200_lp_cascade_create(fs: f32, fc: f32, out: *mut *mut orion_b200_block) -> c_int;
    pub fn orion_b200_fm_mod_create(sample_rate: f32, deviation_hz: f32, rf_hz: f32, out: *mut *mut orion_b200_block) -> c_int;
    pub fn orion_b200_fm_mod_set_deviation(b: *mut orion_b200_block, deviation_hz: f32) -> c_int;
    pub fn orion_b200_cw_mod_create(sample_rate: f32, tone_hz: f32, rise_ms: f32, fall_ms: f32, out: *mut *mut orion_b200_block) -> c_int;
    pub fn orion_b200_ssb_mod_create(fs: f32, audio_bw_hz: f32, audio_if_hz: f32, rf_hz: f32, usb: c_int, out: *mut *mut orion_b200_block) -> c_int;
    pub fn orion_b200_symbol_gain_create(gain: f32, out: *mut *mut orion_b200_block) -> c_int;
    pub fn orion_b200_symbol_gain_set(b: *mut orion_b200_block, gain: f32) -> c_int;
    pub fn orion_b200_decider_create(bits_per_symbol: c_int, out: *mut *mut orion_b200_block) -> c_int;
    pub fn orion_b200_cfo_derotate(cfo_hz: f32, fs: f32, input: *const orion_b200_c32, output: *mut orion_b200_c32, n: usize) -> c_int;
    pub fn orion_b200_agc_rms_create(fs: f32, attack_ms: f32, release_ms: f32, target_rms: f32, out: *mut *mut orion_b200_block) -> c_int;
    pub fn orion_b200_agc_rms_iq_create(fs: f32, attack_ms: f32, release_ms: f32, target_rms: f32, out: *mut *mut orion_b200_block) -> c_int;
    pub fn orion_b200_agc_env(b: *mut orion_b200_block) -> f32;
    pub fn orion_b200_lp_dc_cascade_create(fs: f32, lp_fc: f32, dc_cut_hz: f32, map_sqrt: c_int, out: *mut *mut orion_b200_block) -> c_int;
    pub fn orion_b200_dc_blocker_create(fs: f32, cut_hz: f32, out: *mut *mut orion_b200_block) -> c_int;
    pub fn orion_b200_iir_cascade_create(sos: *const f32, nsections: usize, out: *mut *mut orion_b200_block) -> c_int;
    pub fn orion_b200_fm_demod_create(fs: f32, dev_hz: f32, audio_bw_hz: f32, out: *mut *mut orion_b200_block) -> c_int;
    pub fn orion_b200_fm_demod_with_translate(b: *mut orion_b200_block, freq_hz: f32) -> c_int;
    pub fn orion_b200_pm_demod_create(fs: f32, k: f32, audio_bw_hz: f32, out: *mut *mut orion_b200_block) -> c_int;
    pub fn orion_b200_am_demod_create(fs: f32, audio_bw_hz: f32, out: *mut *mut orion_b200_block) -> c_int;
    pub fn orion_b200_am_demod_with_abs_approx(b: *mut orion_b200_block, k1: f32, k2: f32) -> c_int;
    pub fn orion_b200_ssb_demod_create(fs: f32, bfo_hz: f32, audio_bw_hz: f32, out: *mut *mut orion_b200_block) -> c_int;
    pub fn orion_b200_cw_demod_create(sample_rate: f32, tone_hz: f32, env_bw_hz: f32, out: *mut *mut orion_b200_block) -> c_int;
    pub fn orion_b200_cw_demod_set_gain(b: *mut orion_b200_block, gain: f32) -> c_int;
    pub fn orion_b200_am_mod_create(fs: f32, rf_hz: f32, carrier_level: f32, modulation_index: f32, out: *mut *mut orion_b200_block) -> c_int;
    pub fn orion_b200_am_mod_set_clamp(b: *mut orion_b200_block, on: c_int) -> c_int;
    pub fn orion_b200_pm_mod_create(fs: f32, kp_rad_per_unit: f32, rf_hz: f32, out: *mut *mut orion_b200_block) -> c_int;
    pub fn orion_b200_mod_set_gain(b: *mut orion_b200_block, gain: f32) -> c_int;
    pub fn orion_b200_chain_create(spec: *const orion_b200_chain_spec, out: *mut *mut orion_b200_block) -> c_int;

    pub fn orion_b200_bank_create(specs: *const orion_b200_chain_spec, n_channels: usize, out: *mut *mut orion_b200_bank) -> c_int;
    pub fn orion_b200_bank_set_stream(bank: *mut orion_b200_bank, cuda_stream: *mut c_void) -> c_int;
    pub fn orion_b200_bank_destroy(k: *mut orion_b200_bank);
    pub fn orion_b200_bank_reset(k: *mut orion_b200_bank) -> c_int;
    pub fn orion_b200_bank_channels(k: *const orion_b200_bank) -> usize;
    pub fn orion_b200_bank_last_error(k: *const orion_b200_bank) -> *const c_char;
    pub fn orion_b200_bank_process(k: *mut orion_b200_bank, input: *const c_void, n_in: usize, output: *mut c_void,
                                   out_stride: usize, in_read: *mut usize, out_written: *mut usize) -> c_int;
    pub fn orion_b200_bank_process_dev(k: *mut orion_b200_bank, d_in: *const c_void, n_in: usize, d_out: *mut c_void,
                                       out_stride: usize, in_read: *mut usize, out_written: *mut usize) -> c_int;
    pub fn orion_b200_bank_synchronize(k: *mut orion_b200_bank) -> c_int;
    pub fn orion_b200_bank_launch_count(k: *const orion_b200_bank) -> u64;

    pub fn orion_b200_block_destroy(b: *mut orion_b200_block);
    pub fn orion_b200_block_reset(b: *mut orion_b200_block) -> c_int;
    pub fn orion_b200_block_last_error(b: *const orion_b200_block) -> *const c_char;
    pub fn orion_b200_block_in_item(b: *const orion_b200_block) -> c_int;
    pub fn orion_b200_block_out_item(b: *const orion_b200_block) -> c_int;
    pub fn orion_b200_block_decimation(b: *const orion_b200_block) -> usize;
    pub fn orion_b200_block_plan(b: *const orion_b200_block, n_in: usize, out_cap: usize) -> orion_b200_work_report;
    pub fn orion_b200_block_process(b: *mut orion_b200_block, input: *const c_void, n_in: usize, output: *mut c_void,
                                    out_cap: usize, in_read: *mut usize, out_written: *mut usize) -> c_int;
    pub fn orion_b200_block_process_dev(b: *mut orion_b200_block, d_in: *const c_void, n_in: usize, d_out: *mut c_void,
                                        out_cap: usize, in_read: *mut usize, out_written: *mut usize) -> c_int;
    pub fn orion_b200_block_synchronize(b: *mut orion_b200_block) -> c_int;
    pub fn orion_b200_block_set_stream(b: *mut orion_b200_block, cuda_stream: *mut c_void) -> c_int;
    pub fn orion_b200_block_set_option(b: *mut orion_b200_block, option: c_int, value: f64) -> c_int;
    pub fn orion_b200_block_get_state(b: *mut orion_b200_block, state: *mut f32, cap: usize) -> usize;
    pub fn orion_b200_block_launch_count(b: *const orion_b200_block) -> u64;
    pub fn orion_b200_block_exact_host_ms(b: *const orion_b200_block) -> f64;
    pub fn orion_b200_block_prepare_oscillator(b: *mut orion_b200_block, n_in_per_call: usize, n_calls: usize) -> c_int;
    pub fn orion_b200_last_create_error() -> *const c_char;
    pub fn orion_b200_block_snapshot_size(b: *const orion_b200_block) -> usize;
    pub fn orion_b200_block_snapshot(b: *mut orion_b200_block, buf: *mut c_void, cap: usize) -> c_int;
    pub fn orion_b200_block_restore(b: *mut orion_b200_block, buf: *const c_void, size: usize) -> c_int;
}
