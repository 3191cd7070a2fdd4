/*
 * orion_oracle.h -- CPU restatement of the orion-sdr sample-stream front end.
 *
 * TEST INFRASTRUCTURE ONLY.  Nothing in the shipped GPU path links, loads or calls
 * this library; it exists so tests/, __graft_entry__.smoke() and bench.py's
 * cpu_baseline leg can check the CUDA kernels against the reference's arithmetic.
 *
 * PARITY UNPINNED: the reference (skynavga/orion-sdr v0.0.63) is Rust; no Rust
 * toolchain exists in this image, the reference cannot be built or imported here,
 * and its own test-suite holds no golden vectors or known-answer tests for this
 * path (only behavioural thresholds).  The restatement below
 * follows the reference loop-for-loop (citations are file:line under
 * /root/reference) and is pinned by (i) the reference's behavioural tests
 * transcribed in tests/test_oracle_behaviour.py and (ii) an independent numpy-f32
 * restatement (oracle/np_oracle.py) that must agree bit-for-bit.
 *
 * Arithmetic conventions: IEEE f32 throughout, fmaf()
 * exactly where the reference calls mul_add, every other operation a separately
 * rounded f32 op (compile with -ffp-contract=off), glibc sinf/cosf/expf/powf/sqrtf
 * for design-time math (what Rust's f32 methods call on linux-gnu).
 */
#ifndef ORION_ORACLE_H
#define ORION_ORACLE_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct { float re, im; } oo_c32;
typedef struct { size_t in_read, out_written; } oo_work_report; /* src/core.rs:7-10 */

/* ---- design-time helpers ------------------------------------------------ */
size_t oo_fir_lowpass_ntaps(float fs, float pass_hz, float trans_hz);          /* src/dsp/fir.rs:17-19 */
size_t oo_fir_lowpass_design(float fs, float pass_hz, float trans_hz,
                             float *taps, size_t cap);                         /* src/dsp/fir.rs:16-44 */
float  oo_kaiser_beta(float a_db);                                             /* src/dsp/fir.rs:74-82 */
float  oo_bessel_i0(float x);                                                  /* src/dsp/fir.rs:86-99 */
size_t oo_kaiser_lowpass_taps(size_t num_taps, float cutoff_norm, float stopband_db,
                              float *taps, size_t cap);                        /* src/dsp/fir.rs:113-141 */
float  oo_kaiser_transition_norm(size_t num_taps, float stopband_db);          /* src/dsp/fir.rs:147-150 */
size_t oo_kaiser_num_taps(float transition_norm, float stopband_db);           /* src/dsp/fir.rs:154-157 */
void   oo_lp_biquad_design(float fs, float fc, float coeffs[5]);               /* src/dsp/iir.rs:49-71 */
float  oo_dc_pole(float fs, float cut_hz);                                     /* src/dsp/dc.rs:15-17 */
float  oo_cw_alpha(float fs, float env_bw_hz);                                 /* src/demodulate/cw.rs:15-18 */
float  oo_atan2_approx(float y, float x);                                      /* src/util.rs:305-322 */

/* ---- stateful blocks (opaque; every *_new has a matching oo_free) -------- */
typedef struct oo_block oo_block;
void oo_free(oo_block *b);
void oo_reset(oo_block *b);                    /* zero streaming state where the reference has reset() */

/* Block::process, src/core.rs:12-22.  in/out element types depend on the block. */
oo_work_report oo_process(oo_block *b, const void *in, size_t n_in, void *out, size_t out_cap);

oo_block *oo_fir_lowpass_new(float fs, float pass_hz, float trans_hz);          /* f32 -> f32 */
oo_block *oo_fir_lowpass_from_taps(const float *taps, size_t n);                /* test helper */
oo_block *oo_fir_decimator_new(float fs, size_t m, float cutoff_hz, float trans_hz); /* c32 -> c32 */
oo_block *oo_fir_decimator_from_taps(const float *taps, size_t n, size_t m);    /* test helper */
oo_block *oo_fir_iq_design(size_t num_taps, float cutoff_norm, float stopband_db);   /* c32 -> c32 */
oo_block *oo_fir_iq_from_taps(const float *taps, size_t n);
size_t    oo_fir_iq_group_delay(const oo_block *b);
void      oo_fir_iq_filter_aligned(oo_block *b, oo_c32 *io, size_t n);          /* src/dsp/fir.rs:260-276 */
size_t    oo_get_taps(const oo_block *b, float *taps, size_t cap);              /* FIR blocks */

oo_block *oo_rotator_new(float freq_hz, float fs);                              /* src/dsp/rotator.rs:16-24 */
void      oo_rotator_set_freq(oo_block *b, float freq_hz, float fs);
void      oo_rotator_reset_phase(oo_block *b);
oo_c32    oo_rotator_next(oo_block *b);                                         /* rotator.rs:44-61 */
void      oo_rotator_phasors(oo_block *b, oo_c32 *out, size_t n);               /* n x next() */
void      oo_rotator_rotate_block(oo_block *b, const oo_c32 *in, oo_c32 *out, size_t n);  /* :74-84 */
void      oo_rotator_mix_usb_block(oo_block *b, const oo_c32 *in, float *out, size_t n);  /* :88-94 */
/* process() on a rotator == rotate_block */

oo_block *oo_nco_new(float freq_hz, float fs);                                  /* src/dsp/nco.rs:20-31 */
void      oo_nco_set_freq(oo_block *b, float freq_hz);
void      oo_nco_mix(oo_block *b, const oo_c32 *in, oo_c32 *out, size_t n);     /* mix_with_nco, nco.rs:63-66 */

oo_block *oo_biquad_new(float b0, float b1, float b2, float a1, float a2);      /* f32 -> f32, iir.rs:17-40 */
oo_block *oo_lp_cascade_new(float fs, float fc);                                /* iir.rs:44-84 */
oo_block *oo_lp_dc_cascade_new(float fs, float lp_fc, float dc_cut_hz, int map_sqrt); /* iir.rs:90-187 */
oo_block *oo_dc_blocker_new(float fs, float cut_hz);                            /* dc.rs:8-59 */

oo_block *oo_fm_demod_new(float fs, float dev_hz, float audio_bw_hz);           /* fm.rs:22-32 */
void      oo_fm_demod_with_translate(oo_block *b, float freq_hz);               /* fm.rs:34-37 */
oo_block *oo_pm_demod_new(float fs, float k, float audio_bw_hz);                /* pm.rs:22-32 */
oo_block *oo_am_demod_new(float fs, float audio_bw_hz);                         /* am.rs:24-30 */
void      oo_am_demod_with_abs_approx(oo_block *b, float k1, float k2);         /* am.rs:33-36 */
oo_block *oo_ssb_demod_new(float fs, float bfo_hz, float audio_bw_hz);          /* ssb.rs:15-20 */
oo_block *oo_cw_demod_new(float fs, float tone_hz, float env_bw_hz);            /* cw.rs:15-24 */
void      oo_cw_demod_set_gain(oo_block *b, float g);                           /* cw.rs:25-27 */

/* HalfCosineMf (src/dsp/fir.rs:317-376) as a streaming c32 -> c32 block: one push() per sample */
size_t    oo_half_cosine_taps(size_t sps, float *taps, size_t cap);
oo_block *oo_half_cosine_mf_new(size_t sps);

/* modulators, f32 -> c32 (src/modulate/{fm,pm,am,ssb,cw}.rs) -- SURVEY.md 8(f) row 1: the step before the path in the
 * reference's round-trip tests; CPU only */
oo_block *oo_fm_mod_new(float fs, float deviation_hz, float rf_hz);             /* modulate/fm.rs:22-75 */
oo_block *oo_pm_mod_new(float fs, float kp_rad_per_unit, float rf_hz);          /* modulate/pm.rs:17-47 */
oo_block *oo_am_mod_new(float fs, float rf_hz, float carrier_level, float modulation_index); /* modulate/am.rs:21-120 */
void      oo_am_mod_set_clamp(oo_block *b, int on);
oo_block *oo_ssb_mod_new(float fs, float audio_bw_hz, float audio_if_hz, float rf_hz, int usb); /* modulate/ssb.rs:23-114 */
oo_block *oo_cw_mod_new(float fs, float tone_hz, float rise_ms, float fall_ms); /* modulate/cw.rs:21-102 */
void      oo_mod_set_gain(oo_block *b, float g);
oo_block *oo_agc_rms_new(float fs, float attack_ms, float release_ms, float target_rms);     /* src/dsp/agc.rs:8-75 (f32 -> f32) */
oo_block *oo_agc_rms_iq_new(float fs, float attack_ms, float release_ms, float target_rms);  /* src/dsp/agc.rs:81-150 (c32 -> c32) */
float     oo_agc_env(const oo_block *b);

/* test accelerators (see the .c): the outputs [j0, j1) a keep-every-m-th caller retains from a FRESH
 * FirDecimator / FirLowpassIq fed x[0, n) -- bit-identical to running the block loop-for-loop */
void oo_fir_decim_kept(const float *taps, size_t L, size_t m, const oo_c32 *x, size_t n,
                       oo_c32 *out, size_t j0, size_t j1);
void oo_fir_iq_kept(const float *taps, size_t L, size_t m, const oo_c32 *x, size_t n,
                    oo_c32 *out, size_t j0, size_t j1);

/* streaming-state snapshot for tests (floats; layout documented per block in the .c) */
size_t oo_get_state(const oo_block *b, float *state, size_t cap);

#ifdef __cplusplus
}
#endif
#endif
