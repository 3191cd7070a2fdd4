/*
 * orion_b200.h -- C ABI of the B200-native orion-sdr sample-stream front end.
 *
 * This is the drop-in boundary: the entry points a thin Rust FFI crate (or the PyO3
 * layer, or any other host) binds in place of the reference's CPU blocks.  Every
 * block kind mirrors one reference `impl Block` and keeps its contract
 * (`fn process(&mut self, &[In], &mut [Out]) -> WorkReport`, src/core.rs:12-22):
 *
 *   - `in`/`out` are raw, contiguous, interleaved buffers: C32 = {f32 re, f32 im}
 *     (num_complex::Complex32 is #[repr(C)]), audio = f32.
 *   - the caller owns both buffers; the handle owns all streaming state (FIR history,
 *     oscillator phase, discriminator `prev`, IIR state) in device memory;
 *   - `process` is synchronous: outputs [0, out_written) are complete on return;
 *   - length rules are the reference's: rate-1 blocks consume n = min(n_in, out_cap);
 *     a decimating block consumes ALL n_in and writes min(ceil(n_in/m), out_cap), and its
 *     decimation phase restarts at 0 on every call (src/dsp/decim.rs:44-76);
 *   - nothing throws or unwinds across this boundary; every call returns a status.
 *
 * There is no CPU fallback: if no CUDA device is usable, create() fails with
 * ORION_B200_ERR_NO_DEVICE.  Functions marked [host-only] never touch the GPU.
 *
 * Pointer classes: `orion_b200_block_process` takes HOST pointers (pageable or pinned;
 * H2D/D2H staging inside), `orion_b200_block_process_dev` takes DEVICE pointers
 * (16-byte aligned) and only enqueues work on the block's stream -- call
 * `orion_b200_block_synchronize` (or sync the stream you attached) before reading.
 *
 * Stream contract of `_process_dev`: work is enqueued on the block's stream (its own non-blocking stream, or the one
 * attached with `orion_b200_block_set_stream`).  The block's own stream has NO ordering with the stream that produced
 * `d_in` or will read `d_out` (the legacy default stream included): the input must be complete when the call is
 * enqueued, and the caller synchronises before reading.  Attach your stream to get stream ordering instead.
 * Consecutive calls on one block may overlap on the device (long calls; see ORION_B200_OPT_OVERLAP_LAUNCHES): do not
 * hand two un-synchronised consecutive calls the same output buffer unless overwriting it in any order is acceptable.
 */
#ifndef ORION_B200_H
#define ORION_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define ORION_B200_ABI_VERSION 1

/* status codes */
#define ORION_B200_OK               0
#define ORION_B200_ERR_INVALID      1   /* bad argument (null pointer, zero taps, ...) */
#define ORION_B200_ERR_NO_DEVICE    2   /* no usable CUDA device: there is no CPU fallback */
#define ORION_B200_ERR_CUDA         3   /* a CUDA call failed; see orion_b200_block_last_error */
#define ORION_B200_ERR_ALLOC        4
#define ORION_B200_ERR_UNSUPPORTED  5
#define ORION_B200_ERR_INTERNAL     6   /* device-side watchdog tripped (inter-tile link timeout) */

/* item types */
#define ORION_B200_ITEM_F32 1
#define ORION_B200_ITEM_C32 2
#define ORION_B200_ITEM_U8  3   /* one hard-decision bit per byte (the deciders of src/demodulate/{bpsk,qpsk,qam}.rs) */

typedef struct orion_b200_block orion_b200_block;      /* opaque; one per reference block instance */

typedef struct orion_b200_c32 { float re, im; } orion_b200_c32;

/* mirrors WorkReport, src/core.rs:7-10 */
typedef struct orion_b200_work_report { size_t in_read, out_written; } orion_b200_work_report;

/* ------------------------------------------------------------------------------------
 * Library-level
 * ---------------------------------------------------------------------------------- */
int         orion_b200_abi_version(void);                       /* [host-only] */
const char *orion_b200_last_create_error(void);                 /* [host-only] why the last constructor on this thread failed */
const char *orion_b200_build_info(void);                        /* [host-only] arch, flags */
int         orion_b200_device_count(void);                      /* 0 if no driver/GPU */
int         orion_b200_set_device(int ordinal);                 /* device for blocks created afterwards (per thread) */
const char *orion_b200_status_string(int status);               /* [host-only] */

/* pinned host memory for callers that want H2D/D2H at full PCIe rate */
int  orion_b200_host_alloc(void **ptr, size_t bytes);
void orion_b200_host_free(void *ptr);

/* ------------------------------------------------------------------------------------
 * Host-side design helpers [host-only] -- the reference's design-time math, f32 + libm,
 * quirks included.  Each returns the tap count; pass taps=NULL to size the buffer.
 * ---------------------------------------------------------------------------------- */
size_t orion_b200_fir_lowpass_design(float fs, float pass_hz, float trans_hz,
                                     float *taps, size_t cap);        /* FirLowpass::design, src/dsp/fir.rs:16-44 */
size_t orion_b200_kaiser_lowpass_taps(size_t num_taps, float cutoff_norm, float stopband_db,
                                      float *taps, size_t cap);       /* src/dsp/fir.rs:113-141 */
float  orion_b200_kaiser_transition_norm(size_t num_taps, float stopband_db);   /* src/dsp/fir.rs:147-150 */
size_t orion_b200_kaiser_num_taps(float transition_norm, float stopband_db);    /* src/dsp/fir.rs:154-157 */
void   orion_b200_lp_biquad_design(float fs, float fc, float coeffs[5]);        /* LpCascade::design, src/dsp/iir.rs:49-71 */
float  orion_b200_dc_pole(float fs, float cut_hz);                              /* src/dsp/dc.rs:15-17 */
float  orion_b200_cw_alpha(float fs, float env_bw_hz);                          /* src/demodulate/cw.rs:15-18 */

/* ------------------------------------------------------------------------------------
 * Block constructors.  Each writes a handle to *out and returns a status.
 * ---------------------------------------------------------------------------------- */

/* FirDecimator::new, src/dsp/decim.rs:24-37.  C32 -> C32, ratio 1/m. */
int orion_b200_fir_decimator_create(float fs, size_t m, float cutoff_hz, float trans_hz,
                                    orion_b200_block **out);
/* same block from caller-supplied FirLowpass taps (same tap pairing, fir.rs:57-66) */
int orion_b200_fir_decimator_create_taps(const float *taps, size_t ntaps, size_t m,
                                         orion_b200_block **out);

/* FirLowpassIq::design / from_taps, src/dsp/fir.rs:186-204.  C32 -> C32 streaming (Block impl :279-297). */
int orion_b200_fir_lowpass_iq_create(size_t num_taps, float cutoff_norm, float stopband_db,
                                     orion_b200_block **out);
int orion_b200_fir_lowpass_iq_create_taps(const float *taps, size_t ntaps, orion_b200_block **out);
/* FirLowpassIq::filter_aligned, src/dsp/fir.rs:260-276: resets, filters `io` in place
 * (host pointer), same length, group delay compensated. */
int orion_b200_fir_lowpass_iq_filter_aligned(orion_b200_block *b, orion_b200_c32 *io, size_t n);

/* HalfCosineMf::new + one push() per input sample, src/dsp/fir.rs:317-376 (PSK31 matched filter).
 * C32 -> C32, sps taps, unit energy.  `..._taps` is the design alone [host-only]; returns the tap count. */
size_t orion_b200_half_cosine_mf_taps(size_t sps, float *taps, size_t cap);
int orion_b200_half_cosine_mf_create(size_t sps, orion_b200_block **out);

/* Rotator::new + rotate_block, src/dsp/rotator.rs:16-24,74-84.  C32 -> C32. */
int orion_b200_rotator_create(float freq_hz, float fs, orion_b200_block **out);
/* Rotator + mix_usb_block, src/dsp/rotator.rs:88-94.  C32 -> f32. */
int orion_b200_rotator_usb_create(float freq_hz, float fs, orion_b200_block **out);
/* Nco::new + mix_with_nco per sample, src/dsp/nco.rs:20-31,63-66.  C32 -> C32. */
int orion_b200_nco_mixer_create(float freq_hz, float fs, orion_b200_block **out);
/* Rotator::set_freq / Nco::set_freq (rotator.rs:35-39, nco.rs:34-38): phase is kept. */
int orion_b200_oscillator_set_freq(orion_b200_block *b, float freq_hz, float fs);
/* Rotator::reset_phase, rotator.rs:28-31 */
int orion_b200_oscillator_reset_phase(orion_b200_block *b);

/* IIR blocks, f32 -> f32.  Biquad (iir.rs:17-40), LpCascade (:44-84), LpDcCascade
 * (:90-187; map_sqrt!=0 selects process_mapped(x, sqrt)), DcBlocker (dc.rs:8-59), and a
 * general N-section cascade of Biquad::process stages (sos = n x {b0,b1,b2,a1,a2}). */
int orion_b200_biquad_create(float b0, float b1, float b2, float a1, float a2, orion_b200_block **out);
int orion_b200_lp_cascade_create(float fs, float fc, orion_b200_block **out);
int orion_b200_lp_dc_cascade_create(float fs, float lp_fc, float dc_cut_hz, int map_sqrt,
                                    orion_b200_block **out);
int orion_b200_dc_blocker_create(float fs, float cut_hz, orion_b200_block **out);
int orion_b200_iir_cascade_create(const float *sos, size_t nsections, orion_b200_block **out);

/* Demodulators, C32 -> f32 (src/demodulate/{fm,pm,am,ssb,cw}.rs). */
int orion_b200_fm_demod_create(float fs, float dev_hz, float audio_bw_hz, orion_b200_block **out);   /* fm.rs:22-32 */
int orion_b200_fm_demod_with_translate(orion_b200_block *b, float freq_hz);                           /* fm.rs:34-37 */
int orion_b200_pm_demod_create(float fs, float k, float audio_bw_hz, orion_b200_block **out);        /* pm.rs:22-32 */
int orion_b200_am_demod_create(float fs, float audio_bw_hz, orion_b200_block **out);                 /* am.rs:24-30 */
int orion_b200_am_demod_with_abs_approx(orion_b200_block *b, float k1, float k2);                    /* am.rs:33-36 */
int orion_b200_ssb_demod_create(float fs, float bfo_hz, float audio_bw_hz, orion_b200_block **out);  /* ssb.rs:15-20 */
int orion_b200_cw_demod_create(float sample_rate, float tone_hz, float env_bw_hz, orion_b200_block **out); /* cw.rs:15-24 */
int orion_b200_cw_demod_set_gain(orion_b200_block *b, float gain);                                   /* cw.rs:25-27 */

/* Modulators, f32 audio -> C32 IQ (next-row scope: the step before the path in the reference's round-trip tests).
 * AmDsbMod (src/modulate/am.rs:10-120) and PmDirectPhaseMod (src/modulate/pm.rs:10-47); the FM / SSB / CW modulators
 * are not built on the GPU. */
int orion_b200_am_mod_create(float fs, float rf_hz, float carrier_level, float modulation_index, orion_b200_block **out);
int orion_b200_am_mod_set_clamp(orion_b200_block *b, int on);                                         /* am.rs:34-36 */
int orion_b200_pm_mod_create(float fs, float kp_rad_per_unit, float rf_hz, orion_b200_block **out);
int orion_b200_mod_set_gain(orion_b200_block *b, float gain);

/* ------------------------------------------------------------------------------------
 * Fused chain: [input-rate mixer] -> [FIR, decimate by m] -> [demodulator] -> [extra IIR
 * sections], one streaming kernel.  It is what a user composes today by running the
 * reference blocks back to back through intermediate Vecs (docs/demodulate.md:128-136);
 * results equal that composition within the stated tolerance.
 * ---------------------------------------------------------------------------------- */
#define ORION_B200_MIX_NONE     0
#define ORION_B200_MIX_ROTATE   1   /* Rotator::rotate_block, x*p with FMAs   (rotator.rs:74-84) */
#define ORION_B200_MIX_NCO      2   /* mix_with_nco, x*p unfused              (nco.rs:63-66)     */

#define ORION_B200_FIR_NONE     0
#define ORION_B200_FIR_DECIM    1   /* FirDecimator / FirLowpass tap pairing  (fir.rs:57-66)     */
#define ORION_B200_FIR_IQ       2   /* FirLowpassIq tap pairing, then keep every m-th output of the call */

#define ORION_B200_DEMOD_NONE   0   /* C32 out */
#define ORION_B200_DEMOD_FM     1
#define ORION_B200_DEMOD_PM     2
#define ORION_B200_DEMOD_AM     3   /* PowerSqrt */
#define ORION_B200_DEMOD_AM_ABS 4   /* AbsApprox{k1,k2} */
#define ORION_B200_DEMOD_SSB    5
#define ORION_B200_DEMOD_CW     6
#define ORION_B200_DEMOD_USB    7   /* Rotator::mix_usb_block only, no filter */

typedef struct orion_b200_chain_spec {
    uint32_t struct_size;        /* sizeof(orion_b200_chain_spec) */
    /* 1. input-rate mixer */
    int32_t  mix;                /* ORION_B200_MIX_* */
    float    mix_freq_hz, mix_fs;
    /* 2. FIR + decimation */
    int32_t  fir;                /* ORION_B200_FIR_* */
    const float *taps;           /* reference-order taps (as FirLowpass / FirLowpassIq hold them) */
    size_t   ntaps;
    size_t   decim;              /* m >= 1 */
    /* 3. demodulator, running at fs_demod (= input rate / m) */
    int32_t  demod;              /* ORION_B200_DEMOD_* */
    float    fs_demod;
    float    p0;                 /* FM: dev_hz | PM: k | AM_ABS: k1 | SSB/USB: bfo_hz | CW: env_bw_hz */
    float    p1;                 /* AM_ABS: k2 | CW: gain (0 -> 1.0) */
    float    audio_bw_hz;        /* FM/PM/AM/SSB */
    int32_t  translate;          /* FM only: non-zero = with_translate(translate_hz) */
    float    translate_hz;
    /* 4. extra Biquad::process sections after the demodulator (f32), sos = n x 5 */
    const float *post_sos;
    size_t   n_post;
} orion_b200_chain_spec;

int orion_b200_chain_create(const orion_b200_chain_spec *spec, orion_b200_block **out);

/* ------------------------------------------------------------------------------------
 * Modulators, continued (next-row scope; AmDsbMod and PmDirectPhaseMod are declared further up).
 *  - FmPhaseAccumMod (src/modulate/fm.rs:11-75), f32 -> C32: the running phasor of the reference is a prefix sum of the
 *    phase; the GPU accumulates it exactly in 64-bit fixed point (tile sums, scan, per-item phasor) and mixes with the rf
 *    oscillator replayed bit for bit.  The reference's own f32 rounding walk (~4e-8 rad per step) is not reproducible in
 *    parallel: outputs stay within 1e-4 of the reference for about 10^6 samples after a reset.
 *  - CwKeyedMod (src/modulate/cw.rs:10-102), f32 key envelope -> C32: chunked evaluation of the rise / fall envelope.
 *  - SsbPhasingMod (src/modulate/ssb.rs:11-114), f32 -> C32: audio oscillator, two LpCascade filters, rf oscillator.
 * orion_b200_mod_set_gain (above) applies to all of them.
 * ---------------------------------------------------------------------------------- */
int orion_b200_fm_mod_create(float sample_rate, float deviation_hz, float rf_hz, orion_b200_block **out);
int orion_b200_fm_mod_set_deviation(orion_b200_block *b, float deviation_hz);
int orion_b200_cw_mod_create(float sample_rate, float tone_hz, float rise_ms, float fall_ms, orion_b200_block **out);
int orion_b200_ssb_mod_create(float fs, float audio_bw_hz, float audio_if_hz, float rf_hz, int usb, orion_b200_block **out);

/* ------------------------------------------------------------------------------------
 * Soft-symbol gain blocks and hard-decision slicers (next-row scope, src/demodulate/{bpsk,qpsk,qam}.rs).
 *  - symbol gain: BpskDemod / QpskDemod / QamDemod::process, C32 -> C32, out = (g*re, g*im).
 *  - decider: C32 -> U8, bits_per_symbol = 1 (BpskDecider: re < 0), 2 (QpskDecider: re < 0, im < 0), 4 / 6 / 8
 *    (QamDecider<BITS>: per-axis thresholds, Gray coded, MSB first).  n_syms = min(n_in, out_cap / bits); the WorkReport
 *    is (n_syms, n_syms * bits) like the reference's.
 *  - CFO de-rotation: the one-shot Rotator::new(-cfo_hz, fs).rotate_block(in, out) of sync/ofdm_sync.rs:527-528.
 * ---------------------------------------------------------------------------------- */
int orion_b200_symbol_gain_create(float gain, orion_b200_block **out);
int orion_b200_symbol_gain_set(orion_b200_block *b, float gain);
int orion_b200_decider_create(int bits_per_symbol, orion_b200_block **out);
int orion_b200_cfo_derotate(float cfo_hz, float fs, const orion_b200_c32 *in, orion_b200_c32 *out, size_t n);

/* ------------------------------------------------------------------------------------
 * AGC (next-row scope): AgcRms (f32 -> f32, src/dsp/agc.rs:8-75) and AgcRmsIq (C32 -> C32, agc.rs:81-150).
 * attack_ms / release_ms in milliseconds, target_rms the desired RMS amplitude; gain limits 0.05 .. 20 as in the
 * reference.  The envelope is seeded from the first sample of a call while it is exactly 0 (agc.rs:58-61).
 * The tracker is a data-dependent recurrence; the GPU evaluates it in chunks with a warm-up long enough that the
 * envelope agrees with the reference's to 2^-26 relative (exact for the first chunk of every call).
 * ---------------------------------------------------------------------------------- */
int   orion_b200_agc_rms_create(float fs, float attack_ms, float release_ms, float target_rms, orion_b200_block **out);
int   orion_b200_agc_rms_iq_create(float fs, float attack_ms, float release_ms, float target_rms, orion_b200_block **out);
float orion_b200_agc_env(orion_b200_block *b);      /* the tracked power `env` after the last call (synchronises) */

/* ------------------------------------------------------------------------------------
 * Channel bank: C independent narrowband chains fed by ONE wideband input (BASELINE config 5).
 * Channel c is exactly the block `orion_b200_chain_create(&specs[c])` would build -- typically
 * Rotator(-f_c).rotate_block -> FirDecimator -> FM/AM demod -- with its own streaming state, i.e. what the
 * reference does today with C separate Block chains run one after another over the same input slice.
 * All channels must share the input item type, the decimation factor and the output item type.  Output is
 * row-major [channel][out_stride] items.  Channels are independent, so a bank over a channel sub-range
 * is how the workload is sharded across GPUs (one bank per process / GPU, no collective).
 * ---------------------------------------------------------------------------------- */
typedef struct orion_b200_bank orion_b200_bank;
int    orion_b200_bank_create(const orion_b200_chain_spec *specs, size_t n_channels, orion_b200_bank **out);
/* Run the bank on a caller's CUDA stream (NULL: back to its own); only for banks on the shared front-end path. */
int    orion_b200_bank_set_stream(orion_b200_bank *bank, void *cuda_stream);
void   orion_b200_bank_destroy(orion_b200_bank *k);
int    orion_b200_bank_reset(orion_b200_bank *k);
size_t orion_b200_bank_channels(const orion_b200_bank *k);                         /* [host-only] */
const char *orion_b200_bank_last_error(const orion_b200_bank *k);                  /* [host-only] */
/* host pointers: H2D of the wideband slice once, C kernel launches, one D2H of the [C][out_stride] block */
int orion_b200_bank_process(orion_b200_bank *k, const void *in, size_t n_in, void *out, size_t out_stride,
                            size_t *in_read, size_t *out_written);
/* device pointers; asynchronous on the bank's streams */
int orion_b200_bank_process_dev(orion_b200_bank *k, const void *d_in, size_t n_in, void *d_out, size_t out_stride,
                                size_t *in_read, size_t *out_written);
int orion_b200_bank_synchronize(orion_b200_bank *k);
uint64_t orion_b200_bank_launch_count(const orion_b200_bank *k);                   /* [host-only] */

/* ------------------------------------------------------------------------------------
 * Common block operations
 * ---------------------------------------------------------------------------------- */
void        orion_b200_block_destroy(orion_b200_block *b);
int         orion_b200_block_reset(orion_b200_block *b);     /* zero streaming state (reset()/new state) */
const char *orion_b200_block_last_error(const orion_b200_block *b);   /* [host-only] sticky message */

int orion_b200_block_in_item(const orion_b200_block *b);     /* [host-only] ORION_B200_ITEM_* */
int orion_b200_block_out_item(const orion_b200_block *b);    /* [host-only] */
size_t orion_b200_block_decimation(const orion_b200_block *b);        /* [host-only] m */
/* [host-only] the WorkReport a call with (n_in, out_cap) will return -- pure length rules */
orion_b200_work_report orion_b200_block_plan(const orion_b200_block *b, size_t n_in, size_t out_cap);

/* Block::process with host buffers. */
int orion_b200_block_process(orion_b200_block *b, const void *in, size_t n_in,
                             void *out, size_t out_cap, size_t *in_read, size_t *out_written);
/* Same with device buffers; asynchronous on the block's stream. */
int orion_b200_block_process_dev(orion_b200_block *b, const void *d_in, size_t n_in,
                                 void *d_out, size_t out_cap, size_t *in_read, size_t *out_written);
int orion_b200_block_synchronize(orion_b200_block *b);
/* Attach an existing CUDA stream (cudaStream_t / CUstream as void*); NULL restores the block's own. */
int orion_b200_block_set_stream(orion_b200_block *b, void *cuda_stream);

/* options */
#define ORION_B200_OPT_FIR_GLOBAL    1  /* 1: evaluate the FIR straight from global memory in the
                                           reference's accumulation order and rounding (FirLowpass::dot
                                           unfused, FirLowpassIq::push fused): bit-faithful, slow.
                                           0 (default): shared-memory staged polyphase FIR, FMA */
#define ORION_B200_OPT_USE_TMA       2  /* 1 (default): interior tiles are staged with cp.async.bulk.tensor;
                                           0: every tile uses the cooperative loader */
#define ORION_B200_OPT_SERIAL_TILES  3  /* 1: debug -- one CTA walks the tiles in order */
#define ORION_B200_OPT_OVERLAP_LAUNCHES 4 /* 1: on a stream attached with orion_b200_block_set_stream, let the kernel of call N+1
                                           start while call N drains (programmatic dependent launch; state handed over between
                                           calls is still waited for).  The caller promises that the input of a call is complete
                                           when the call is enqueued (not produced by the kernel enqueued just before it).  On the
                                           block's own stream this overlap is always on: nothing foreign can be enqueued there. */
#define ORION_B200_OPT_EXACT_NCO     5  /* oscillator arithmetic.  -1 (default): by block kind -- the reference's f32 phasor recurrence
                                           (renormalised every 1024 steps, rotator.rs:44-61) is replayed bit for bit wherever the
                                           ABSOLUTE phase reaches the output (Rotator, NcoMixer, mix_usb_block, SsbProductDemod,
                                           the modulators, mixer -> FIR -> C32 / SSB chains); a closed-form 64-bit phase is used
                                           where only phase differences or magnitudes matter (FM translate, mixer -> FIR -> FM / PM /
                                           AM / CW).  1: replay everywhere it applies; 0: closed form everywhere (drifts from the
                                           reference by its rounding: ~5e-2 rad after 24 M samples).  The replay is a sequential walk
                                           on the host (~3 ns per item, orion_b200_block_exact_host_ms) plus a parallel expansion
                                           kernel.  Select it before the first call after a reset. */
int orion_b200_block_set_option(orion_b200_block *b, int option, double value);
/* host milliseconds this block has spent walking the oscillator recurrence (exact mode) since creation [host-only] */
double orion_b200_block_exact_host_ms(const orion_b200_block *b);
/* Walk the exact-mode oscillator(s) of the block AHEAD of the stream: enough for the next `n_calls` process() calls of
 * `n_in_per_call` input items each (the phasor sequence of rotator.rs:44-61 does not depend on the data, so it can be
 * produced before the samples arrive -- e.g. while the previous buffer is still being captured).  process() then only
 * copies the anchors it needs; without this call it walks its own n_in items inline.  Host time goes to
 * orion_b200_block_exact_host_ms.  No-op for blocks that use the closed-form phase.  [host-only] */
int orion_b200_block_prepare_oscillator(orion_b200_block *b, size_t n_in_per_call, size_t n_calls);

/* Streaming-state snapshot for the parity harness.  Layout (20 floats):
 * [0..1] discriminator prev (re,im); [2..3] oscillator call counters (input-rate, demod-rate,
 * modulo 2^24); [4 ..] recursive-section states, 2 per section in section order (8 sections). */
size_t orion_b200_block_get_state(orion_b200_block *b, float *state, size_t cap);

/* Checkpoint / resume: the complete streaming state of a block (FIR history, oscillator phase and counters,
 * discriminator `prev`, section states) as an opaque blob.  The reference's blocks are plain data and derive
 * Clone (fm.rs:10, iir.rs:4,43,89, rotator.rs:7, fir.rs:176): snapshot + restore into a block built with the
 * same constructor arguments is that clone.  Both calls drain the block's stream first. */
size_t orion_b200_block_snapshot_size(const orion_b200_block *b);
int    orion_b200_block_snapshot(orion_b200_block *b, void *buf, size_t cap);
int    orion_b200_block_restore(orion_b200_block *b, const void *buf, size_t size);

/* number of kernels this block has launched since creation (for bench accounting) [host-only] */
uint64_t orion_b200_block_launch_count(const orion_b200_block *b);

/* Debug (only in a -DORION_TRACE=1 build): per-tile SM-clock stamps of the kernel's phases, 16 x int64 per tile
 * {consume, slot ready, FIR done, front done, section phase done, globaltimer, smid, warp, refill, front end, look-back
 * done, scan start, scan end, publish end, active mask lane 0, active mask lane 31} plus 2 trailing words {kernel start,
 * kernel end in globaltimer ns}: size the buffer (16 * ntiles + 2) * 8 bytes.  d_trace is a device pointer (NULL disables). */
int orion_b200_debug_set_trace(orion_b200_block *b, void *d_trace);

/* Plan introspection for the host-logic tests [host-only].  info[12] = {front, R, U, Mb, O, P,
 * P_pad, HR, row_samples, row_pitch, rows, H}; returns the polyphase tap-table length (floats). */
size_t orion_b200_debug_fir_plan(int fir_kind, const float *taps, size_t ntaps, size_t m, int info[12],
                                 float *table, size_t cap, float *g, size_t gcap);
/* Scan tables of one section group [host-only]: sections = nsec x {type, c0..c4}  (nsec <= 2; type 1 biquad
 * {b0,b1,b2,a1,a2}, 2 DC {r}, 3 one-pole {a, 1-a}); npt = items per lane.  out receives
 * {D, depth, agg_only, 0, imp[16][4], lv[5][16], lane[32][16], lb[32][16], lb32[16], tile[16]};
 * returns the length in floats. */
size_t orion_b200_debug_group_tables(const float *sections, size_t nsec, int npt, float *out, size_t cap);

#ifdef __cplusplus
}
#endif
#endif /* ORION_B200_H */
