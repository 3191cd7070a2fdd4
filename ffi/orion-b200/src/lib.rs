//! GPU drop-ins for orion-sdr's sample-stream front end.
//!
//! Every type here implements `orion_sdr::core::Block` (reference `src/core.rs:12-22`) with the In/Out
//! item types of the CPU block it replaces, so it slots into `IqToAudioChain::new(..)`,
//! `IqToIqChain::new(..)`, `util::run_block(..)` and any user graph unchanged:
//!
//! ```ignore
//! use orion_b200::GpuFmQuadratureDemod;                       // instead of orion_sdr::demodulate::FmQuadratureDemod
//! let mut chain = IqToAudioChain::new(GpuFmQuadratureDemod::new(48_000.0, 2_500.0, 5_000.0)?);
//! let audio = chain.process(iq);
//! ```
//!
//! NOT COMPILED in this repository's build environment (no Rust toolchain); all executable
//! verification goes through the same C ABI from Python (`tests/`).  There is no CPU fallback:
//! constructors return `Err(Error::NoDevice)` when no CUDA device is usable.
use core::ffi::c_void;
use num_complex::Complex32 as C32;
use orion_b200_sys as sys;
use orion_sdr::core::{Block, WorkReport};
use std::ffi::CStr;
use std::marker::PhantomData;
use std::ptr;

#[derive(Debug, Clone, PartialEq, Eq)]
pub enum Error {
    Invalid,
    NoDevice,
    Cuda(String),
    Alloc,
    Unsupported,
    Internal(String),
}

fn status_to_error(st: i32, b: *const sys::orion_b200_block) -> Error {
    let msg = || unsafe {
        if b.is_null() {
            CStr::from_ptr(sys::orion_b200_status_string(st)).to_string_lossy().into_owned()
        } else {
            CStr::from_ptr(sys::orion_b200_block_last_error(b)).to_string_lossy().into_owned()
        }
    };
    match st {
        sys::ORION_B200_ERR_INVALID => Error::Invalid,
        sys::ORION_B200_ERR_NO_DEVICE => Error::NoDevice,
        sys::ORION_B200_ERR_CUDA => Error::Cuda(msg()),
        sys::ORION_B200_ERR_ALLOC => Error::Alloc,
        sys::ORION_B200_ERR_UNSUPPORTED => Error::Unsupported,
        _ => Error::Internal(msg()),
    }
}

/// Owning handle of one `orion_b200_block`; `In`/`Out` are the reference block's item types.
pub struct GpuBlock<In, Out> {
    h: *mut sys::orion_b200_block,
    /// sticky: the last non-zero status of `process` (the trait's `process` is infallible, core.rs:15)
    last_error: Option<Error>,
    _io: PhantomData<(In, Out)>,
}

// One CUDA stream per handle, no thread affinity, no shared mutable state between handles:
// the handle may move between threads like the plain-data CPU blocks (SURVEY.md 8b "Threading").
unsafe impl<In, Out> Send for GpuBlock<In, Out> {}

impl<In, Out> GpuBlock<In, Out> {
    fn from_create(f: impl FnOnce(*mut *mut sys::orion_b200_block) -> i32) -> Result<Self, Error> {
        let mut h = ptr::null_mut();
        let st = f(&mut h);
        if st != sys::ORION_B200_OK {
            return Err(status_to_error(st, ptr::null()));
        }
        Ok(Self { h, last_error: None, _io: PhantomData })
    }
    /// zero the streaming state (`reset()` / a freshly constructed block)
    pub fn reset(&mut self) {
        unsafe { sys::orion_b200_block_reset(self.h) };
    }
    pub fn take_error(&mut self) -> Option<Error> {
        self.last_error.take()
    }
    /// The complete streaming state as an opaque blob (what `Clone` gives the CPU blocks).
    pub fn snapshot(&mut self) -> Result<Vec<u8>, Error> {
        let n = unsafe { sys::orion_b200_block_snapshot_size(self.h) };
        let mut buf = vec![0u8; n];
        let st = unsafe { sys::orion_b200_block_snapshot(self.h, buf.as_mut_ptr() as *mut c_void, n) };
        if st == sys::ORION_B200_OK { Ok(buf) } else { Err(status_to_error(st, self.h)) }
    }
    /// Restore a blob taken from a block built with the same constructor arguments.
    pub fn restore(&mut self, blob: &[u8]) -> Result<(), Error> {
        let st = unsafe { sys::orion_b200_block_restore(self.h, blob.as_ptr() as *const c_void, blob.len()) };
        if st == sys::ORION_B200_OK { Ok(()) } else { Err(status_to_error(st, self.h)) }
    }
    pub fn raw(&mut self) -> *mut sys::orion_b200_block {
        self.h
    }
}

impl<In, Out> Drop for GpuBlock<In, Out> {
    fn drop(&mut self) {
        unsafe { sys::orion_b200_block_destroy(self.h) };
    }
}

impl<In, Out> Block for GpuBlock<In, Out> {
    type In = In;
    type Out = Out;
    /// Same contract as the CPU block: synchronous, outputs `[0, out_written)` complete on return.
    /// A failed call reports `WorkReport { 0, 0 }` and leaves the error in `take_error()`
    /// (debug builds panic), because `Block::process` has no error channel.
    fn process(&mut self, input: &[In], output: &mut [Out]) -> WorkReport {
        let (mut r, mut w) = (0usize, 0usize);
        let st = unsafe {
            sys::orion_b200_block_process(
                self.h,
                input.as_ptr() as *const c_void,
                input.len(),
                output.as_mut_ptr() as *mut c_void,
                output.len(),
                &mut r,
                &mut w,
            )
        };
        if st != sys::ORION_B200_OK {
            let e = status_to_error(st, self.h);
            debug_assert!(false, "orion_b200_block_process failed: {e:?}");
            self.last_error = Some(e);
            return WorkReport { in_read: 0, out_written: 0 };
        }
        WorkReport { in_read: r, out_written: w }
    }
}

macro_rules! gpu_block {
    ($(#[$doc:meta])* $name:ident : $inp:ty => $out:ty) => {
        $(#[$doc])*
        pub struct $name(GpuBlock<$inp, $out>);
        impl Block for $name {
            type In = $inp;
            type Out = $out;
            #[inline]
            fn process(&mut self, input: &[$inp], output: &mut [$out]) -> WorkReport {
                self.0.process(input, output)
            }
        }
        impl $name {
            pub fn reset(&mut self) { self.0.reset() }
            pub fn take_error(&mut self) -> Option<Error> { self.0.take_error() }
        }
    };
}

gpu_block!(/// `dsp::FirDecimator` (decim.rs:24-76): consumes all input, writes `min(ceil(n/m), out.len())`.
    GpuFirDecimator: C32 => C32);
gpu_block!(/// `dsp::FirLowpassIq` as a streaming `Block` (fir.rs:279-297).
    GpuFirLowpassIq: C32 => C32);
gpu_block!(/// `dsp::Rotator::rotate_block` (rotator.rs:74-84).
    GpuRotator: C32 => C32);
gpu_block!(/// `dsp::LpCascade` / `LpDcCascade` / `DcBlocker` / `Biquad` cascades (iir.rs, dc.rs).
    GpuIirCascade: f32 => f32);
gpu_block!(/// `demodulate::FmQuadratureDemod` (fm.rs:11-78).
    GpuFmQuadratureDemod: C32 => f32);
gpu_block!(/// `demodulate::PmQuadratureDemod` (pm.rs:12-67).
    GpuPmQuadratureDemod: C32 => f32);
gpu_block!(/// `demodulate::AmEnvelopeDemod` (am.rs:10-130).
    GpuAmEnvelopeDemod: C32 => f32);
gpu_block!(/// `demodulate::SsbProductDemod` (ssb.rs:9-72).
    GpuSsbProductDemod: C32 => f32);
gpu_block!(/// `demodulate::CwEnvelopeDemod` (cw.rs:8-47).
    GpuCwEnvelopeDemod: C32 => f32);
gpu_block!(/// The fused chain [mixer] -> FIR -> decimate -> demod -> [sections] in one kernel launch.
    GpuChain: C32 => f32);

impl GpuFirDecimator {
    pub fn new(fs: f32, m: usize, cutoff_hz: f32, trans_hz: f32) -> Result<Self, Error> {
        GpuBlock::from_create(|h| unsafe { sys::orion_b200_fir_decimator_create(fs, m, cutoff_hz, trans_hz, h) }).map(Self)
    }
}
impl GpuFirLowpassIq {
    pub fn design(num_taps: usize, cutoff_norm: f32, stopband_db: f32) -> Result<Self, Error> {
        GpuBlock::from_create(|h| unsafe { sys::orion_b200_fir_lowpass_iq_create(num_taps, cutoff_norm, stopband_db, h) }).map(Self)
    }
    pub fn from_taps(taps: &[f32]) -> Result<Self, Error> {
        GpuBlock::from_create(|h| unsafe { sys::orion_b200_fir_lowpass_iq_create_taps(taps.as_ptr(), taps.len(), h) }).map(Self)
    }
    /// fir.rs:260-276
    pub fn filter_aligned(&mut self, io: &mut [C32]) -> Result<(), Error> {
        let st = unsafe { sys::orion_b200_fir_lowpass_iq_filter_aligned(self.0.h, io.as_mut_ptr() as *mut sys::orion_b200_c32, io.len()) };
        if st == sys::ORION_B200_OK { Ok(()) } else { Err(status_to_error(st, self.0.h)) }
    }
}
impl GpuRotator {
    pub fn new(freq_hz: f32, fs: f32) -> Result<Self, Error> {
        GpuBlock::from_create(|h| unsafe { sys::orion_b200_rotator_create(freq_hz, fs, h) }).map(Self)
    }
    pub fn set_freq(&mut self, freq_hz: f32, fs: f32) {
        unsafe { sys::orion_b200_oscillator_set_freq(self.0.h, freq_hz, fs) };
    }
    pub fn reset_phase(&mut self) {
        unsafe { sys::orion_b200_oscillator_reset_phase(self.0.h) };
    }
}
impl GpuIirCascade {
    pub fn lp_cascade(fs: f32, fc: f32) -> Result<Self, Error> {
        GpuBlock::from_create(|h| unsafe { sys::orion_b200_lp_cascade_create(fs, fc, h) }).map(Self)
    }
    pub fn lp_dc_cascade(fs: f32, lp_fc: f32, dc_cut_hz: f32, map_sqrt: bool) -> Result<Self, Error> {
        GpuBlock::from_create(|h| unsafe { sys::orion_b200_lp_dc_cascade_create(fs, lp_fc, dc_cut_hz, map_sqrt as i32, h) }).map(Self)
    }
    pub fn dc_blocker(fs: f32, cut_hz: f32) -> Result<Self, Error> {
        GpuBlock::from_create(|h| unsafe { sys::orion_b200_dc_blocker_create(fs, cut_hz, h) }).map(Self)
    }
    pub fn from_sos(sos: &[[f32; 5]]) -> Result<Self, Error> {
        GpuBlock::from_create(|h| unsafe { sys::orion_b200_iir_cascade_create(sos.as_ptr() as *const f32, sos.len(), h) }).map(Self)
    }
}
impl GpuFmQuadratureDemod {
    pub fn new(fs: f32, dev_hz: f32, audio_bw_hz: f32) -> Result<Self, Error> {
        GpuBlock::from_create(|h| unsafe { sys::orion_b200_fm_demod_create(fs, dev_hz, audio_bw_hz, h) }).map(Self)
    }
    pub fn with_translate(mut self, freq_hz: f32) -> Self {
        unsafe { sys::orion_b200_fm_demod_with_translate(self.0.h, freq_hz) };
        self
    }
}
impl GpuPmQuadratureDemod {
    pub fn new(fs: f32, k: f32, audio_bw_hz: f32) -> Result<Self, Error> {
        GpuBlock::from_create(|h| unsafe { sys::orion_b200_pm_demod_create(fs, k, audio_bw_hz, h) }).map(Self)
    }
}
impl GpuAmEnvelopeDemod {
    pub fn new(fs: f32, audio_bw_hz: f32) -> Result<Self, Error> {
        GpuBlock::from_create(|h| unsafe { sys::orion_b200_am_demod_create(fs, audio_bw_hz, h) }).map(Self)
    }
    pub fn with_abs_approx(mut self, k1: f32, k2: f32) -> Self {
        unsafe { sys::orion_b200_am_demod_with_abs_approx(self.0.h, k1, k2) };
        self
    }
}
impl GpuSsbProductDemod {
    pub fn new(fs: f32, bfo_hz: f32, audio_bw_hz: f32) -> Result<Self, Error> {
        GpuBlock::from_create(|h| unsafe { sys::orion_b200_ssb_demod_create(fs, bfo_hz, audio_bw_hz, h) }).map(Self)
    }
}
impl GpuCwEnvelopeDemod {
    pub fn new(sample_rate: f32, tone_hz: f32, env_bw_hz: f32) -> Result<Self, Error> {
        GpuBlock::from_create(|h| unsafe { sys::orion_b200_cw_demod_create(sample_rate, tone_hz, env_bw_hz, h) }).map(Self)
    }
    pub fn set_gain(&mut self, g: f32) {
        unsafe { sys::orion_b200_cw_demod_set_gain(self.0.h, g) };
    }
}
impl GpuChain {
    /// The C1 chain: `FirDecimator::new(fs, m, cutoff, trans)` -> `FmQuadratureDemod::new(fs/m, dev, bw).with_translate(f)`.
    pub fn fir_decim_fm(fs: f32, m: usize, cutoff_hz: f32, trans_hz: f32, dev_hz: f32, audio_bw_hz: f32,
                        translate_hz: Option<f32>) -> Result<Self, Error> {
        let n = unsafe { sys::orion_b200_fir_lowpass_design(fs, cutoff_hz, trans_hz, ptr::null_mut(), 0) };
        let mut taps = vec![0.0f32; n];
        unsafe { sys::orion_b200_fir_lowpass_design(fs, cutoff_hz, trans_hz, taps.as_mut_ptr(), n) };
        let spec = sys::orion_b200_chain_spec {
            struct_size: core::mem::size_of::<sys::orion_b200_chain_spec>() as u32,
            mix: sys::ORION_B200_MIX_NONE, mix_freq_hz: 0.0, mix_fs: 0.0,
            fir: sys::ORION_B200_FIR_DECIM, taps: taps.as_ptr(), ntaps: n, decim: m,
            demod: sys::ORION_B200_DEMOD_FM, fs_demod: fs / m as f32, p0: dev_hz, p1: 0.0, audio_bw_hz,
            translate: translate_hz.is_some() as i32, translate_hz: translate_hz.unwrap_or(0.0),
            post_sos: ptr::null(), n_post: 0,
        };
        GpuBlock::from_create(|h| unsafe { sys::orion_b200_chain_create(&spec, h) }).map(Self)
    }
}

// ---- next-row blocks: AGC, the remaining modulators, symbol gain and deciders --------------------
gpu_block!(/// `dsp::AgcRms` (agc.rs:8-75).
    GpuAgcRms: f32 => f32);
gpu_block!(/// `dsp::AgcRmsIq` (agc.rs:81-150).
    GpuAgcRmsIq: C32 => C32);
gpu_block!(/// `modulate::FmPhaseAccumMod` (modulate/fm.rs:11-75).
    GpuFmPhaseAccumMod: f32 => C32);
gpu_block!(/// `modulate::CwKeyedMod` (modulate/cw.rs:10-102).
    GpuCwKeyedMod: f32 => C32);
gpu_block!(/// `modulate::SsbPhasingMod` (modulate/ssb.rs:11-114).
    GpuSsbPhasingMod: f32 => C32);
gpu_block!(/// `BpskDemod` / `QpskDemod` / `QamDemod` (demodulate/{bpsk,qpsk,qam}.rs): soft symbols scaled by a gain.
    GpuSymbolGain: C32 => C32);
gpu_block!(/// `BpskDecider` / `QpskDecider` / `QamDecider<BITS>`: one hard-decision bit per output byte.
    GpuDecider: C32 => u8);

impl GpuAgcRms {
    pub fn new(fs: f32, attack_ms: f32, release_ms: f32, target_rms: f32) -> Result<Self, Error> {
        GpuBlock::from_create(|h| unsafe { sys::orion_b200_agc_rms_create(fs, attack_ms, release_ms, target_rms, h) }).map(Self)
    }
}
impl GpuAgcRmsIq {
    pub fn new(fs: f32, attack_ms: f32, release_ms: f32, target_rms: f32) -> Result<Self, Error> {
        GpuBlock::from_create(|h| unsafe { sys::orion_b200_agc_rms_iq_create(fs, attack_ms, release_ms, target_rms, h) }).map(Self)
    }
}
impl GpuFmPhaseAccumMod {
    pub fn new(sample_rate: f32, deviation_hz: f32, rf_hz: f32) -> Result<Self, Error> {
        GpuBlock::from_create(|h| unsafe { sys::orion_b200_fm_mod_create(sample_rate, deviation_hz, rf_hz, h) }).map(Self)
    }
    pub fn set_deviation(&mut self, deviation_hz: f32) { unsafe { sys::orion_b200_fm_mod_set_deviation(self.0.h, deviation_hz) }; }
    pub fn set_gain(&mut self, g: f32) { unsafe { sys::orion_b200_mod_set_gain(self.0.h, g) }; }
}
impl GpuCwKeyedMod {
    pub fn new(sample_rate: f32, tone_hz: f32, rise_ms: f32, fall_ms: f32) -> Result<Self, Error> {
        GpuBlock::from_create(|h| unsafe { sys::orion_b200_cw_mod_create(sample_rate, tone_hz, rise_ms, fall_ms, h) }).map(Self)
    }
    pub fn set_gain(&mut self, g: f32) { unsafe { sys::orion_b200_mod_set_gain(self.0.h, g) }; }
}
impl GpuSsbPhasingMod {
    pub fn new(fs: f32, audio_bw_hz: f32, audio_if_hz: f32, rf_hz: f32, usb: bool) -> Result<Self, Error> {
        GpuBlock::from_create(|h| unsafe { sys::orion_b200_ssb_mod_create(fs, audio_bw_hz, audio_if_hz, rf_hz, usb as i32, h) }).map(Self)
    }
}
impl GpuSymbolGain {
    pub fn new(gain: f32) -> Result<Self, Error> {
        GpuBlock::from_create(|h| unsafe { sys::orion_b200_symbol_gain_create(gain, h) }).map(Self)
    }
    pub fn set_gain(&mut self, g: f32) { unsafe { sys::orion_b200_symbol_gain_set(self.0.h, g) }; }
}
impl GpuDecider {
    /// bits per symbol: 1 BPSK, 2 QPSK, 4 / 6 / 8 QAM-16 / 64 / 256 (qam.rs:13-18)
    pub fn new(bits_per_symbol: usize) -> Result<Self, Error> {
        GpuBlock::from_create(|h| unsafe { sys::orion_b200_decider_create(bits_per_symbol as i32, h) }).map(Self)
    }
}
/// `Rotator::new(-cfo_hz, fs).rotate_block(iq, out)` (sync/ofdm_sync.rs:527-528) in one call.
pub fn cfo_derotate(iq: &[C32], out: &mut [C32], cfo_hz: f32, fs: f32) -> Result<(), Error> {
    let n = iq.len().min(out.len());
    let st = unsafe { sys::orion_b200_cfo_derotate(cfo_hz, fs, iq.as_ptr() as *const sys::orion_b200_c32, out.as_mut_ptr() as *mut sys::orion_b200_c32, n) };
    if st == sys::ORION_B200_OK { Ok(()) } else { Err(status_to_error(st, ptr::null())) }
}
