/*
 * orion_oracle.c -- CPU restatement of the orion-sdr sample-stream front end.
 *
 * TEST INFRASTRUCTURE ONLY (see orion_oracle.h).  PARITY UNPINNED by reference
 * golden vectors (none exist for this path); pinned by transcribed behavioural
 * tests and by an independent numpy restatement.
 *
 * Every function cites the reference lines it follows (paths relative to
 * /root/reference).  Build: see oracle/Makefile (-ffp-contract=off is mandatory).
 */
#include "orion_oracle.h"

#include <math.h>
#include <stdlib.h>
#include <string.h>

#define OO_PI   3.14159265358979323846f   /* core::f32::consts::PI  == 0x40490FDB */
#define OO_TAU  6.28318530717958647692f   /* core::f32::consts::TAU == 0x40C90FDB */
#define OO_FRAC_PI_2 1.57079632679489661923f
#define OO_FRAC_PI_4 0.78539816339744830962f
#define OO_EPS  1.1920929e-07f            /* f32::EPSILON */

static inline float oo_maxf(float a, float b) { return (a > b || b != b) ? a : b; } /* f32::max */
static inline float oo_minf(float a, float b) { return (a < b || b != b) ? a : b; }

/* ======================================================================== */
/* design-time                                                              */
/* ======================================================================== */

/* src/dsp/fir.rs:17-19 */
size_t oo_fir_lowpass_ntaps(float fs, float pass_hz, float trans_hz) {
    pass_hz = oo_maxf(pass_hz, 10.0f);
    trans_hz = oo_maxf(trans_hz, pass_hz * 0.2f);
    float c = ceilf(fs / trans_hz);
    size_t n = (c > 0.0f) ? (size_t)c : 0;   /* `as usize` saturates at 0 */
    if (n < 31) n = 31;
    return n | 1;
}

/* src/dsp/fir.rs:16-44 (sinc x Hann, extra 2*fc on the off-centre taps, DC-normalised) */
size_t oo_fir_lowpass_design(float fs, float pass_hz, float trans_hz, float *taps, size_t cap) {
    size_t ntaps = oo_fir_lowpass_ntaps(fs, pass_hz, trans_hz);
    if (!taps || cap < ntaps) return ntaps;
    pass_hz = oo_maxf(pass_hz, 10.0f);
    float fc = pass_hz / fs;
    long m0 = (long)ntaps / 2;
    for (size_t n = 0; n < ntaps; ++n) {
        long m = (long)n - m0;
        float sinc;
        if (m == 0) {
            sinc = 2.0f * fc;
        } else {
            float x = OO_PI * (float)m;
            sinc = (2.0f * fc) * sinf(2.0f * OO_PI * fc * (float)m) / x;
        }
        float w = 0.5f - 0.5f * cosf(2.0f * OO_PI * (float)n / ((float)ntaps - 1.0f));
        taps[n] = sinc * w;
    }
    float s = 0.0f;
    for (size_t n = 0; n < ntaps; ++n) s += taps[n];
    for (size_t n = 0; n < ntaps; ++n) taps[n] /= s;
    return ntaps;
}

/* src/dsp/fir.rs:74-82 */
float oo_kaiser_beta(float a_db) {
    if (a_db > 50.0f) return 0.1102f * (a_db - 8.7f);
    if (a_db >= 21.0f) return 0.5842f * powf(a_db - 21.0f, 0.4f) + 0.07886f * (a_db - 21.0f);
    return 0.0f;
}

/* src/dsp/fir.rs:86-99 */
float oo_bessel_i0(float x) {
    float half = 0.5f * x;
    float term = 1.0f, sum = 1.0f;
    for (unsigned k = 1; k <= 40; ++k) {
        term *= half / (float)k;
        float t = term * term;
        sum += t;
        if (t < 1e-12f * sum) break;
    }
    return sum;
}

/* src/dsp/fir.rs:113-141 */
size_t oo_kaiser_lowpass_taps(size_t num_taps, float cutoff_norm, float stopband_db,
                              float *taps, size_t cap) {
    size_t m = (num_taps < 3 ? 3 : num_taps) | 1;
    if (!taps || cap < m) return m;
    float mid = (float)(m / 2);
    float fc = cutoff_norm;
    if (fc < 1e-4f) fc = 1e-4f;
    if (fc > 0.4999f) fc = 0.4999f;
    float beta = oo_kaiser_beta(stopband_db);
    float i0_beta = oo_bessel_i0(beta);
    for (size_t n = 0; n < m; ++n) {
        float d = (float)n - mid;
        float ideal = (d == 0.0f) ? 2.0f * fc : sinf(OO_TAU * fc * d) / (OO_PI * d);
        float r = d / mid;
        float w = oo_bessel_i0(beta * sqrtf(oo_maxf(1.0f - r * r, 0.0f))) / i0_beta;
        taps[n] = ideal * w;
    }
    float s = 0.0f;
    for (size_t n = 0; n < m; ++n) s += taps[n];
    if (fabsf(s) > OO_EPS)
        for (size_t n = 0; n < m; ++n) taps[n] /= s;
    return m;
}

/* src/dsp/fir.rs:147-150 */
float oo_kaiser_transition_norm(size_t num_taps, float stopband_db) {
    float m = (float)((num_taps < 3 ? 3 : num_taps) | 1);
    return (oo_maxf(stopband_db, 21.0f) - 8.0f) / (14.36f * m);
}

/* src/dsp/fir.rs:154-157 */
size_t oo_kaiser_num_taps(float transition_norm, float stopband_db) {
    float m = ceilf((oo_maxf(stopband_db, 21.0f) - 8.0f) / (14.36f * oo_maxf(transition_norm, 1e-4f)));
    return ((size_t)oo_maxf(m, 3.0f)) | 1;
}

/* src/dsp/iir.rs:49-71 (identical arithmetic at iir.rs:111-137) -> {b0,b1,b2,a1,a2} */
void oo_lp_biquad_design(float fs, float fc, float c[5]) {
    float w0 = OO_TAU * fc / fs;
    float sn = sinf(w0), cs = cosf(w0);
    float alpha = sn / (2.0f * sqrtf(0.5f));
    float b0 = (1.0f - cs) * 0.5f;
    float b1 = 1.0f - cs;
    float b2 = (1.0f - cs) * 0.5f;
    float a0 = 1.0f + alpha;
    float a1 = -2.0f * cs;
    float a2 = 1.0f - alpha;
    float norm = 1.0f / a0;
    c[0] = b0 * norm; c[1] = b1 * norm; c[2] = b2 * norm; c[3] = a1 * norm; c[4] = a2 * norm;
}

/* src/dsp/dc.rs:15-17 and src/dsp/iir.rs:122 */
float oo_dc_pole(float fs, float cut_hz) {
    float r = 1.0f - 2.0f * OO_PI * (oo_maxf(cut_hz, 0.1f) / fs);
    if (r < 0.0f) r = 0.0f;
    if (r > 0.9999f) r = 0.9999f;
    return r;
}

/* src/demodulate/cw.rs:15-18 */
float oo_cw_alpha(float fs, float env_bw_hz) {
    float fc = oo_maxf(env_bw_hz, 1.0f);
    return expf(-OO_TAU * fc / fs);
}

/* src/util.rs:305-322 */
float oo_atan2_approx(float y, float x) {
    float ax = fabsf(x), ay = fabsf(y);
    float mn, mx;
    if (ax < ay) { mn = ax; mx = ay; } else { mn = ay; mx = ax; }
    float r = mn / (mx + OO_EPS);
    float r2 = r * r;
    float phi = r * (OO_FRAC_PI_4 + r2 * (-0.2447f + r2 * 0.0663f));
    if (ax < ay) phi = OO_FRAC_PI_2 - phi;
    float sgn = (y < 0.0f) ? -1.0f : 1.0f;
    if (x < 0.0f) return (OO_PI - phi) * sgn;
    return phi * sgn;
}

/* ======================================================================== */
/* block state                                                              */
/* ======================================================================== */

enum {
    K_FIR_LOWPASS = 1, K_FIR_DECIM, K_FIR_IQ, K_ROTATOR, K_NCO, K_BIQUAD, K_LP_CASCADE,
    K_LP_DC_CASCADE, K_DC_BLOCKER, K_FM, K_PM, K_AM, K_SSB, K_CW,
    K_MOD_FM, K_MOD_PM, K_MOD_AM, K_MOD_SSB, K_MOD_CW,     /* f32 -> c32; SURVEY.md 8(f) row 1 */
    K_HCMF,                                                /* HalfCosineMf, SURVEY.md 8(f) row 2 */
    K_AGC, K_AGC_IQ                                        /* AgcRms / AgcRmsIq, SURVEY.md 8(f) row 3 */
};

typedef struct { float *taps; float *delay; size_t len, idx; } fir_real;     /* fir.rs:7-12 */
typedef struct { float b0, b1, b2, a1, a2, z1, z2; } biquad;                 /* iir.rs:4-13 */
typedef struct { oo_c32 z, w; uint32_t ctr; } rotator;                       /* rotator.rs:7-12 */
typedef struct {                                                             /* iir.rs:90-108 */
    float z0_1, z0_2, z1_1, z1_2, dc_x1, dc_y1, b0, b1, b2, a1, a2, r;
} lpdc;

struct oo_block {
    int kind;
    /* FIR */
    fir_real fi, fq;                 /* fir_lowpass uses fi; decimator uses both (decim.rs:12-13) */
    size_t m;                        /* decimation factor (decim.rs:27) */
    float *ri, *rq, *yi, *yq; size_t scratch;   /* decim.rs:15-18 */
    float *iq_taps; oo_c32 *iq_delay; size_t iq_len, iq_idx;                 /* fir.rs:177-181 */
    /* oscillators */
    rotator rot; int has_rot;
    float nco_fs;
    /* IIR */
    biquad bq[2];
    lpdc ld; int map_sqrt;
    float dc_r, dc_x1, dc_y1;
    /* demods */
    float fs, k; oo_c32 prev;
    int abs_approx; float k1, k2;
    float alpha, y, gain;
    /* modulators (src/modulate/{fm,pm,am,ssb,cw}.rs) */
    rotator rot2;                    /* second oscillator: SSB rf_nco */
    biquad bq2[2];                   /* SSB lp_q */
    oo_c32 mz; uint32_t mctr;        /* FM running phasor + its renorm counter */
    float carrier, mindex; int clamp, usb;
    float env, a_rise, a_fall;
    /* AGC (src/dsp/agc.rs): env above, a_rise = attack_a, a_fall = release_a */
    float target_rms, min_gain, max_gain;
};

static void fir_real_init(fir_real *f, const float *taps, size_t n) {
    f->taps = (float *)malloc(n * sizeof(float));
    f->delay = (float *)calloc(n, sizeof(float));
    memcpy(f->taps, taps, n * sizeof(float));
    f->len = n; f->idx = 0;
}

/* src/dsp/fir.rs:47-66: delay[idx]=x; acc = sum_t delay[(idx+len-1-t)%len]*taps[t]; idx++ */
static inline float fir_real_step(fir_real *f, float x) {
    const size_t len = f->len;
    f->delay[f->idx] = x;
    float acc = 0.0f;
    for (size_t t = 0; t < len; ++t) {
        size_t d = (f->idx + len - 1 - t) % len;
        acc += f->delay[d] * f->taps[t];
    }
    f->idx = (f->idx + 1) % len;
    return acc;
}

/* src/dsp/iir.rs:34-40 */
static inline float biquad_step(biquad *q, float x) {
    float y = fmaf(x, q->b0, q->z1);
    q->z1 = fmaf(x, q->b1, q->z2) - q->a1 * y;
    q->z2 = x * q->b2 - q->a2 * y;
    return y;
}

/* src/dsp/rotator.rs:44-61 == src/dsp/nco.rs:42-58 */
static inline oo_c32 rotator_step(rotator *r) {
    float zr = fmaf(r->z.re, r->w.re, -(r->z.im * r->w.im));
    float zi = fmaf(r->z.im, r->w.re, r->z.re * r->w.im);
    r->z.re = zr; r->z.im = zi;
    r->ctr += 1u;
    if ((r->ctr & 0x3FFu) == 0u) {
        float r2 = r->z.re * r->z.re + r->z.im * r->z.im;
        float inv = 1.0f / sqrtf(r2);
        r->z.re *= inv; r->z.im *= inv;
    }
    return r->z;
}

static void rotator_init(rotator *r, float freq_hz, float fs) {   /* rotator.rs:16-24 */
    float phi = OO_TAU * freq_hz / fs;
    r->z.re = 1.0f; r->z.im = 0.0f;
    r->w.re = cosf(phi); r->w.im = sinf(phi);
    r->ctr = 0;
}

/* src/dsp/iir.rs:151-165 (map_sqrt=0) and :170-186 (map_sqrt=1, f = sqrt) */
static inline float lpdc_step(lpdc *s, float x, int map_sqrt) {
    float y0 = fmaf(x, s->b0, s->z0_1);
    s->z0_1 = fmaf(x, s->b1, s->z0_2) - s->a1 * y0;
    s->z0_2 = x * s->b2 - s->a2 * y0;
    float y1 = fmaf(y0, s->b0, s->z1_1);
    s->z1_1 = fmaf(y0, s->b1, s->z1_2) - s->a1 * y1;
    s->z1_2 = y0 * s->b2 - s->a2 * y1;
    float mapped = map_sqrt ? sqrtf(y1) : y1;
    float y = mapped - s->dc_x1 + s->r * s->dc_y1;
    s->dc_x1 = mapped;
    s->dc_y1 = y;
    return y;
}

static void lpdc_init(lpdc *s, float fs, float lp_fc, float dc_cut) {   /* iir.rs:111-137 */
    float c[5];
    oo_lp_biquad_design(fs, lp_fc, c);
    memset(s, 0, sizeof(*s));
    s->b0 = c[0]; s->b1 = c[1]; s->b2 = c[2]; s->a1 = c[3]; s->a2 = c[4];
    s->r = oo_dc_pole(fs, dc_cut);
}

static oo_block *blk_new(int kind) {
    oo_block *b = (oo_block *)calloc(1, sizeof(oo_block));
    b->kind = kind;
    return b;
}

void oo_free(oo_block *b) {
    if (!b) return;
    free(b->fi.taps); free(b->fi.delay); free(b->fq.taps); free(b->fq.delay);
    free(b->ri); free(b->rq); free(b->yi); free(b->yq);
    free(b->iq_taps); free(b->iq_delay);
    free(b);
}

/* ---- constructors -------------------------------------------------------- */

oo_block *oo_fir_lowpass_from_taps(const float *taps, size_t n) {
    oo_block *b = blk_new(K_FIR_LOWPASS);
    fir_real_init(&b->fi, taps, n);
    return b;
}

oo_block *oo_fir_lowpass_new(float fs, float pass_hz, float trans_hz) {
    size_t n = oo_fir_lowpass_ntaps(fs, pass_hz, trans_hz);
    float *t = (float *)malloc(n * sizeof(float));
    oo_fir_lowpass_design(fs, pass_hz, trans_hz, t, n);
    oo_block *b = oo_fir_lowpass_from_taps(t, n);
    free(t);
    return b;
}

oo_block *oo_fir_decimator_from_taps(const float *taps, size_t n, size_t m) {
    oo_block *b = blk_new(K_FIR_DECIM);
    fir_real_init(&b->fi, taps, n);
    fir_real_init(&b->fq, taps, n);
    b->m = m < 1 ? 1 : m;                                    /* decim.rs:30 */
    return b;
}

/* src/dsp/decim.rs:24-37 */
oo_block *oo_fir_decimator_new(float fs, size_t m, float cutoff_hz, float trans_hz) {
    size_t n = oo_fir_lowpass_ntaps(fs, cutoff_hz, trans_hz);
    float *t = (float *)malloc(n * sizeof(float));
    oo_fir_lowpass_design(fs, cutoff_hz, trans_hz, t, n);
    oo_block *b = oo_fir_decimator_from_taps(t, n, m);
    b->fs = fs;
    free(t);
    return b;
}

/* src/dsp/fir.rs:192-204 */
oo_block *oo_fir_iq_from_taps(const float *taps, size_t n) {
    oo_block *b = blk_new(K_FIR_IQ);
    float one = 1.0f;
    if (n == 0) { taps = &one; n = 1; }
    b->iq_taps = (float *)malloc(n * sizeof(float));
    memcpy(b->iq_taps, taps, n * sizeof(float));
    b->iq_delay = (oo_c32 *)calloc(n, sizeof(oo_c32));
    b->iq_len = n; b->iq_idx = 0;
    return b;
}

/* src/dsp/fir.rs:186-188 */
oo_block *oo_fir_iq_design(size_t num_taps, float cutoff_norm, float stopband_db) {
    size_t n = oo_kaiser_lowpass_taps(num_taps, cutoff_norm, stopband_db, NULL, 0);
    float *t = (float *)malloc(n * sizeof(float));
    oo_kaiser_lowpass_taps(num_taps, cutoff_norm, stopband_db, t, n);
    oo_block *b = oo_fir_iq_from_taps(t, n);
    free(t);
    return b;
}

size_t oo_fir_iq_group_delay(const oo_block *b) { return (b->iq_len - 1) / 2; }   /* fir.rs:216-218 */

size_t oo_get_taps(const oo_block *b, float *taps, size_t cap) {
    const float *src = NULL; size_t n = 0;
    if (b->kind == K_FIR_LOWPASS || b->kind == K_FIR_DECIM) { src = b->fi.taps; n = b->fi.len; }
    else if (b->kind == K_FIR_IQ) { src = b->iq_taps; n = b->iq_len; }
    if (taps && cap >= n && n) memcpy(taps, src, n * sizeof(float));
    return n;
}

oo_block *oo_rotator_new(float freq_hz, float fs) {
    oo_block *b = blk_new(K_ROTATOR);
    rotator_init(&b->rot, freq_hz, fs);
    return b;
}
void oo_rotator_set_freq(oo_block *b, float freq_hz, float fs) {    /* rotator.rs:35-39 */
    float phi = OO_TAU * freq_hz / fs;
    b->rot.w.re = cosf(phi); b->rot.w.im = sinf(phi);
}
void oo_rotator_reset_phase(oo_block *b) {                          /* rotator.rs:28-31 */
    b->rot.z.re = 1.0f; b->rot.z.im = 0.0f; b->rot.ctr = 0;
}
oo_c32 oo_rotator_next(oo_block *b) { return rotator_step(&b->rot); }
void oo_rotator_phasors(oo_block *b, oo_c32 *out, size_t n) {
    for (size_t i = 0; i < n; ++i) out[i] = rotator_step(&b->rot);
}
/* src/dsp/rotator.rs:74-84 */
void oo_rotator_rotate_block(oo_block *b, const oo_c32 *in, oo_c32 *out, size_t n) {
    for (size_t i = 0; i < n; ++i) {
        oo_c32 p = rotator_step(&b->rot);
        float a = in[i].re, bb = in[i].im;
        out[i].re = fmaf(a, p.re, -(bb * p.im));
        out[i].im = fmaf(bb, p.re, a * p.im);
    }
}
/* src/dsp/rotator.rs:88-94 */
void oo_rotator_mix_usb_block(oo_block *b, const oo_c32 *in, float *out, size_t n) {
    for (size_t i = 0; i < n; ++i) {
        oo_c32 p = rotator_step(&b->rot);
        out[i] = fmaf(in[i].re, p.re, in[i].im * p.im);
    }
}

oo_block *oo_nco_new(float freq_hz, float fs) {                     /* nco.rs:20-31 */
    oo_block *b = blk_new(K_NCO);
    rotator_init(&b->rot, freq_hz, fs);
    b->nco_fs = fs;
    return b;
}
void oo_nco_set_freq(oo_block *b, float freq_hz) {                  /* nco.rs:34-38 */
    float dphi = OO_TAU * freq_hz / b->nco_fs;
    b->rot.w.re = cosf(dphi); b->rot.w.im = sinf(dphi);
}
/* src/dsp/nco.rs:63-66 */
void oo_nco_mix(oo_block *b, const oo_c32 *in, oo_c32 *out, size_t n) {
    for (size_t i = 0; i < n; ++i) {
        oo_c32 p = rotator_step(&b->rot);
        float c = p.re, s = p.im;
        oo_c32 x = in[i];
        out[i].re = x.re * c - x.im * s;
        out[i].im = x.re * s + x.im * c;
    }
}

oo_block *oo_biquad_new(float b0, float b1, float b2, float a1, float a2) {
    oo_block *b = blk_new(K_BIQUAD);
    biquad q = { b0, b1, b2, a1, a2, 0.0f, 0.0f };
    b->bq[0] = q;
    return b;
}
oo_block *oo_lp_cascade_new(float fs, float fc) {
    oo_block *b = blk_new(K_LP_CASCADE);
    float c[5];
    oo_lp_biquad_design(fs, fc, c);
    biquad q = { c[0], c[1], c[2], c[3], c[4], 0.0f, 0.0f };
    b->bq[0] = q; b->bq[1] = q;
    return b;
}
oo_block *oo_lp_dc_cascade_new(float fs, float lp_fc, float dc_cut_hz, int map_sqrt) {
    oo_block *b = blk_new(K_LP_DC_CASCADE);
    lpdc_init(&b->ld, fs, lp_fc, dc_cut_hz);
    b->map_sqrt = map_sqrt;
    return b;
}
oo_block *oo_dc_blocker_new(float fs, float cut_hz) {
    oo_block *b = blk_new(K_DC_BLOCKER);
    b->dc_r = oo_dc_pole(fs, cut_hz);
    return b;
}

/* src/demodulate/fm.rs:22-32 */
oo_block *oo_fm_demod_new(float fs, float dev_hz, float audio_bw_hz) {
    oo_block *b = blk_new(K_FM);
    b->fs = fs;
    b->k = 1.0f / oo_maxf(dev_hz, 1.0f);
    float c[5];
    oo_lp_biquad_design(fs, audio_bw_hz * 0.9f, c);
    biquad q = { c[0], c[1], c[2], c[3], c[4], 0.0f, 0.0f };
    b->bq[0] = q; b->bq[1] = q;
    b->prev.re = 1.0f; b->prev.im = 0.0f;
    return b;
}
void oo_fm_demod_with_translate(oo_block *b, float freq_hz) {       /* fm.rs:34-37 */
    rotator_init(&b->rot, freq_hz, b->fs);
    b->has_rot = 1;
}
/* src/demodulate/pm.rs:22-32 */
oo_block *oo_pm_demod_new(float fs, float k, float audio_bw_hz) {
    oo_block *b = oo_fm_demod_new(fs, 1.0f, audio_bw_hz);
    b->kind = K_PM;
    b->k = k;
    return b;
}
/* src/demodulate/am.rs:24-36 */
oo_block *oo_am_demod_new(float fs, float audio_bw_hz) {
    oo_block *b = blk_new(K_AM);
    lpdc_init(&b->ld, fs, audio_bw_hz * 0.9f, 2.0f);
    return b;
}
void oo_am_demod_with_abs_approx(oo_block *b, float k1, float k2) {
    b->abs_approx = 1; b->k1 = k1; b->k2 = k2;
}
/* src/demodulate/ssb.rs:15-20 */
oo_block *oo_ssb_demod_new(float fs, float bfo_hz, float audio_bw_hz) {
    oo_block *b = blk_new(K_SSB);
    lpdc_init(&b->ld, fs, audio_bw_hz * 0.9f, 2.0f);
    rotator_init(&b->rot, bfo_hz, fs);
    return b;
}
/* src/demodulate/cw.rs:15-27 */
oo_block *oo_cw_demod_new(float fs, float tone_hz, float env_bw_hz) {
    (void)tone_hz;
    oo_block *b = blk_new(K_CW);
    b->alpha = oo_cw_alpha(fs, env_bw_hz);
    b->y = 0.0f; b->gain = 1.0f;
    return b;
}
void oo_cw_demod_set_gain(oo_block *b, float g) { b->gain = g; }

/* ---- reset --------------------------------------------------------------- */

/* ---- HalfCosineMf, src/dsp/fir.rs:317-376: split delay lines, y[n] = sum_t taps[t] x[n-t], unfused, t ascending ---- */
size_t oo_half_cosine_taps(size_t sps, float *taps, size_t cap) {               /* fir.rs:325-346 */
    size_t n = sps < 1 ? 1 : sps;
    if (!taps || cap < n) return n;
    if (sps <= 1) { taps[0] = 1.0f; }
    else {
        float denom = (float)(sps - 1);
        for (size_t i = 0; i < sps; ++i) taps[i] = 0.5f - 0.5f * cosf(OO_PI * (float)i / denom);
    }
    float energy = 0.0f;
    for (size_t i = 0; i < n; ++i) energy += taps[i] * taps[i];
    float scale = (energy > 0.0f) ? 1.0f / sqrtf(energy) : 1.0f;
    for (size_t i = 0; i < n; ++i) taps[i] = taps[i] * scale;
    return n;
}
oo_block *oo_half_cosine_mf_new(size_t sps) {
    oo_block *b = blk_new(K_HCMF);
    size_t n = oo_half_cosine_taps(sps, NULL, 0);
    float *t = (float *)malloc(n * sizeof(float));
    oo_half_cosine_taps(sps, t, n);
    fir_real_init(&b->fi, t, n);                 /* delay_re */
    fir_real_init(&b->fq, t, n);                 /* delay_im; both share idx (kept equal) */
    free(t);
    return b;
}

/* ---- modulators (f32 -> c32): the step before the path in the reference's round-trip tests ---- */

/* mix_with_nco, src/dsp/nco.rs:63-66 (num-complex Mul is unfused) */
static inline oo_c32 mix_nco_step(rotator *r, oo_c32 x) {
    oo_c32 p = rotator_step(r);
    oo_c32 y = { x.re * p.re - x.im * p.im, x.re * p.im + x.im * p.re };
    return y;
}

oo_block *oo_fm_mod_new(float fs, float deviation_hz, float rf_hz) {            /* modulate/fm.rs:22-31 */
    oo_block *b = blk_new(K_MOD_FM);
    b->fs = fs; b->k = deviation_hz; b->gain = 1.0f;
    b->mz.re = 1.0f; b->mz.im = 0.0f; b->mctr = 0;
    rotator_init(&b->rot, rf_hz, fs);
    return b;
}
oo_block *oo_pm_mod_new(float fs, float kp_rad_per_unit, float rf_hz) {         /* modulate/pm.rs:17-23 */
    oo_block *b = blk_new(K_MOD_PM);
    b->k = kp_rad_per_unit; b->gain = 1.0f;
    rotator_init(&b->rot, rf_hz, fs);
    return b;
}
oo_block *oo_am_mod_new(float fs, float rf_hz, float carrier_level, float modulation_index) {   /* modulate/am.rs:21-30 */
    oo_block *b = blk_new(K_MOD_AM);
    b->gain = 1.0f; b->carrier = carrier_level; b->mindex = modulation_index; b->clamp = 0;
    rotator_init(&b->rot, rf_hz, fs);
    return b;
}
void oo_am_mod_set_clamp(oo_block *b, int on) { b->clamp = on; }                 /* modulate/am.rs:34-36 */
oo_block *oo_ssb_mod_new(float fs, float audio_bw_hz, float audio_if_hz, float rf_hz, int usb) { /* modulate/ssb.rs:23-35 */
    oo_block *b = blk_new(K_MOD_SSB);
    float c[5];
    oo_lp_biquad_design(fs, audio_bw_hz * 0.9f, c);
    for (int i = 0; i < 2; ++i) {
        biquad q = { c[0], c[1], c[2], c[3], c[4], 0.0f, 0.0f };
        b->bq[i] = q; b->bq2[i] = q;
    }
    b->usb = usb;
    rotator_init(&b->rot, audio_if_hz, fs);
    rotator_init(&b->rot2, rf_hz, fs);
    return b;
}
oo_block *oo_cw_mod_new(float fs, float tone_hz, float rise_ms, float fall_ms) { /* modulate/cw.rs:21-34 */
    oo_block *b = blk_new(K_MOD_CW);
    float tau_r = (oo_maxf(rise_ms, 0.1f) * 1e-3f) * fs;
    float tau_f = (oo_maxf(fall_ms, 0.1f) * 1e-3f) * fs;
    b->a_rise = expf(-1.0f / tau_r);
    b->a_fall = expf(-1.0f / tau_f);
    b->env = 0.0f; b->gain = 1.0f;
    rotator_init(&b->rot, tone_hz, fs);
    return b;
}
void oo_mod_set_gain(oo_block *b, float g) { b->gain = g; }

/* AgcRms::new / AgcRmsIq::new, src/dsp/agc.rs:20-31,93-104 */
static oo_block *agc_new(int kind, float fs, float attack_ms, float release_ms, float target_rms) {
    oo_block *b = blk_new(kind);
    b->fs = fs;
    b->a_rise = expf(-1.0f / (fs * (oo_maxf(attack_ms, 1e-3f) / 1000.0f)));
    b->a_fall = expf(-1.0f / (fs * (oo_maxf(release_ms, 1e-3f) / 1000.0f)));
    b->target_rms = oo_maxf(target_rms, 1e-6f);
    b->min_gain = 0.05f; b->max_gain = 20.0f;
    b->env = 0.0f;
    return b;
}
oo_block *oo_agc_rms_new(float fs, float attack_ms, float release_ms, float target_rms) {
    return agc_new(K_AGC, fs, attack_ms, release_ms, target_rms);
}
oo_block *oo_agc_rms_iq_new(float fs, float attack_ms, float release_ms, float target_rms) {
    return agc_new(K_AGC_IQ, fs, attack_ms, release_ms, target_rms);
}
float oo_agc_env(const oo_block *b) { return b->env; }
/* agc.rs:33-40 (update_env) + :64-69: gain from the tracked power */
static inline float agc_gain_step(oo_block *b, float x2) {
    float a = (x2 > b->env) ? b->a_rise : b->a_fall;
    b->env = a * b->env + (1.0f - a) * x2;
    float rms = oo_maxf(sqrtf(b->env), 1e-6f);
    float g = b->target_rms / rms;
    g = oo_minf(oo_maxf(g, b->min_gain), b->max_gain);       /* f32::clamp */
    return g;
}

void oo_reset(oo_block *b) {
    if (b->fi.delay) { memset(b->fi.delay, 0, b->fi.len * sizeof(float)); b->fi.idx = 0; }
    if (b->fq.delay) { memset(b->fq.delay, 0, b->fq.len * sizeof(float)); b->fq.idx = 0; }
    if (b->iq_delay) { memset(b->iq_delay, 0, b->iq_len * sizeof(oo_c32)); b->iq_idx = 0; } /* fir.rs:221-224 */
    b->rot.z.re = 1.0f; b->rot.z.im = 0.0f; b->rot.ctr = 0;
    b->bq[0].z1 = b->bq[0].z2 = b->bq[1].z1 = b->bq[1].z2 = 0.0f;                  /* iir.rs:29-32,74-77 */
    b->ld.z0_1 = b->ld.z0_2 = b->ld.z1_1 = b->ld.z1_2 = b->ld.dc_x1 = b->ld.dc_y1 = 0.0f; /* iir.rs:140-147 */
    b->dc_x1 = b->dc_y1 = 0.0f;
    b->prev.re = 1.0f; b->prev.im = 0.0f;
    b->y = 0.0f;
}

/* ---- FirLowpassIq -------------------------------------------------------- */
/* src/dsp/fir.rs:229-247 */
static inline oo_c32 fir_iq_push(oo_block *b, oo_c32 s) {
    const size_t len = b->iq_len;
    const size_t idx = b->iq_idx;
    b->iq_delay[idx] = s;
    float re = 0.0f, im = 0.0f;
    for (size_t j = 0; j <= idx; ++j) {
        oo_c32 d = b->iq_delay[idx - j];
        float t = b->iq_taps[j];
        re = fmaf(d.re, t, re);
        im = fmaf(d.im, t, im);
    }
    for (size_t k = 0; k + idx + 1 < len; ++k) {
        oo_c32 d = b->iq_delay[len - 1 - k];
        float t = b->iq_taps[idx + 1 + k];
        re = fmaf(d.re, t, re);
        im = fmaf(d.im, t, im);
    }
    b->iq_idx = (idx + 1 == len) ? 0 : idx + 1;
    oo_c32 y = { re, im };
    return y;
}

/* src/dsp/fir.rs:260-276 */
void oo_fir_iq_filter_aligned(oo_block *b, oo_c32 *io, size_t n) {
    size_t d = oo_fir_iq_group_delay(b);
    memset(b->iq_delay, 0, b->iq_len * sizeof(oo_c32));
    b->iq_idx = 0;
    oo_c32 zero = { 0.0f, 0.0f };
    for (size_t i = 0; i < d; ++i) (void)fir_iq_push(b, i < n ? io[i] : zero);
    for (size_t i = 0; i < n; ++i) {
        oo_c32 x = (i + d < n) ? io[i + d] : zero;
        io[i] = fir_iq_push(b, x);
    }
}

/* ---- test accelerators: only the outputs a decimating caller keeps -------------------------
 * FirDecimator::process (decim.rs:44-76) runs both FirLowpass filters at the INPUT rate and then keeps
 * yi[j*m], yq[j*m]; every filter output is an independent dot product, so the kept ones can be evaluated
 * alone, in the same accumulation order and rounding as fir_real_step above (t ascending; taps[t] pairs
 * with x[n-1-t], taps[L-1] with x[n]; samples before the start of a FRESH block are the zeros of its delay
 * line).  Bit-identical to the loop-for-loop path (tests/test_oracle_crosscheck.py pins that); m times
 * cheaper, which is what lets the full-size BASELINE configurations be checked in seconds.
 * Outputs [j0, j1) of a fresh block fed x[0, n). */
void oo_fir_decim_kept(const float *taps, size_t L, size_t m, const oo_c32 *x, size_t n,
                       oo_c32 *out, size_t j0, size_t j1) {
    for (size_t j = j0; j < j1; ++j) {
        const size_t p = j * m;                      /* index of the newest sample */
        if (p >= n) { out[j - j0].re = 0.0f; out[j - j0].im = 0.0f; continue; }
        float ar = 0.0f, ai = 0.0f;
        for (size_t t = 0; t + 1 < L; ++t) {
            if (t + 1 > p) { ar += 0.0f * taps[t]; ai += 0.0f * taps[t]; continue; }
            const oo_c32 d = x[p - 1 - t];
            ar += d.re * taps[t];
            ai += d.im * taps[t];
        }
        ar += x[p].re * taps[L - 1];
        ai += x[p].im * taps[L - 1];
        out[j - j0].re = ar; out[j - j0].im = ai;
    }
}
/* The same for FirLowpassIq::push (fir.rs:229-247: y[n] = sum_j taps[j] x[n-j], fused, j ascending) followed
 * by a keep-every-m-th pick: outputs y[j*m], j in [j0, j1), of a fresh block. */
void oo_fir_iq_kept(const float *taps, size_t L, size_t m, const oo_c32 *x, size_t n,
                    oo_c32 *out, size_t j0, size_t j1) {
    for (size_t j = j0; j < j1; ++j) {
        const size_t p = j * m;
        if (p >= n) { out[j - j0].re = 0.0f; out[j - j0].im = 0.0f; continue; }
        float re = 0.0f, im = 0.0f;
        for (size_t k = 0; k < L; ++k) {
            oo_c32 d = { 0.0f, 0.0f };
            if (k <= p) d = x[p - k];
            re = fmaf(d.re, taps[k], re);
            im = fmaf(d.im, taps[k], im);
        }
        out[j - j0].re = re; out[j - j0].im = im;
    }
}

/* ---- Block::process ------------------------------------------------------ */
oo_work_report oo_process(oo_block *b, const void *in, size_t n_in, void *out, size_t out_cap) {
    oo_work_report wr = { 0, 0 };
    const oo_c32 *cin = (const oo_c32 *)in;
    const float *fin = (const float *)in;
    oo_c32 *cout = (oo_c32 *)out;
    float *fout = (float *)out;
    size_t n = n_in < out_cap ? n_in : out_cap;

    switch (b->kind) {
    case K_FIR_LOWPASS:                                      /* fir.rs:47-54 */
        for (size_t i = 0; i < n; ++i) fout[i] = fir_real_step(&b->fi, fin[i]);
        wr.in_read = n; wr.out_written = n;
        break;

    case K_FIR_DECIM: {                                      /* decim.rs:44-76 */
        size_t nn = n_in;
        if (b->scratch < nn) {
            b->ri = (float *)realloc(b->ri, nn * sizeof(float));
            b->rq = (float *)realloc(b->rq, nn * sizeof(float));
            b->yi = (float *)realloc(b->yi, nn * sizeof(float));
            b->yq = (float *)realloc(b->yq, nn * sizeof(float));
            b->scratch = nn;
        }
        for (size_t k = 0; k < nn; ++k) { b->ri[k] = cin[k].re; b->rq[k] = cin[k].im; }
        for (size_t k = 0; k < nn; ++k) b->yi[k] = fir_real_step(&b->fi, b->ri[k]);
        for (size_t k = 0; k < nn; ++k) b->yq[k] = fir_real_step(&b->fq, b->rq[k]);
        size_t m = b->m;
        size_t n_out = (nn + m - 1) / m;
        size_t n_write = n_out < out_cap ? n_out : out_cap;
        for (size_t j = 0; j < n_write; ++j) {
            cout[j].re = b->yi[j * m];
            cout[j].im = b->yq[j * m];
        }
        wr.in_read = nn; wr.out_written = n_write;
        break;
    }

    case K_FIR_IQ:                                           /* fir.rs:287-296 */
        for (size_t i = 0; i < n; ++i) cout[i] = fir_iq_push(b, cin[i]);
        wr.in_read = n; wr.out_written = n;
        break;

    case K_ROTATOR:
        oo_rotator_rotate_block(b, cin, cout, n);
        wr.in_read = n; wr.out_written = n;
        break;

    case K_NCO:
        oo_nco_mix(b, cin, cout, n);
        wr.in_read = n; wr.out_written = n;
        break;

    case K_BIQUAD:
        for (size_t i = 0; i < n; ++i) fout[i] = biquad_step(&b->bq[0], fin[i]);
        wr.in_read = n; wr.out_written = n;
        break;

    case K_LP_CASCADE:                                       /* iir.rs:79-83 */
        for (size_t i = 0; i < n; ++i) {
            float x = biquad_step(&b->bq[0], fin[i]);
            fout[i] = biquad_step(&b->bq[1], x);
        }
        wr.in_read = n; wr.out_written = n;
        break;

    case K_LP_DC_CASCADE:
        for (size_t i = 0; i < n; ++i) fout[i] = lpdc_step(&b->ld, fin[i], b->map_sqrt);
        wr.in_read = n; wr.out_written = n;
        break;

    case K_DC_BLOCKER: {                                     /* dc.rs:40-58 */
        float x1 = b->dc_x1, y1 = b->dc_y1, r = b->dc_r;
        for (size_t i = 0; i < n; ++i) {
            float x = fin[i];
            float y = x - x1 + r * y1;
            fout[i] = y;
            x1 = x; y1 = y;
        }
        b->dc_x1 = x1; b->dc_y1 = y1;
        wr.in_read = n; wr.out_written = n;
        break;
    }

    case K_FM:                                               /* fm.rs:45-77 */
        for (size_t i = 0; i < n; ++i) {
            oo_c32 z = cin[i];
            if (b->has_rot) {
                /* input[i] * r.next().conj()  (num-complex 0.4.6 Mul, unfused) */
                oo_c32 p = rotator_step(&b->rot);
                float cr = p.re, ci = -p.im;
                oo_c32 t = { z.re * cr - z.im * ci, z.re * ci + z.im * cr };
                z = t;
            }
            float pr = z.re * b->prev.re + z.im * b->prev.im;
            float pi = z.im * b->prev.re - z.re * b->prev.im;
            float d = oo_atan2_approx(pi, pr) * b->k;
            float x = biquad_step(&b->bq[0], d);
            fout[i] = biquad_step(&b->bq[1], x);
            b->prev = z;
        }
        wr.in_read = n; wr.out_written = n;
        break;

    case K_PM:                                               /* pm.rs:39-66 */
        for (size_t i = 0; i < n; ++i) {
            oo_c32 z = cin[i];
            /* z * prev.conj() (num-complex Mul) */
            float cr = b->prev.re, ci = -b->prev.im;
            float wre = z.re * cr - z.im * ci;
            float wim = z.re * ci + z.im * cr;
            float d = b->k * oo_atan2_approx(wim, wre);
            float x = biquad_step(&b->bq[0], d);
            fout[i] = biquad_step(&b->bq[1], x);
            b->prev = z;
        }
        wr.in_read = n; wr.out_written = n;
        break;

    case K_AM:                                               /* am.rs:43-129 */
        for (size_t i = 0; i < n; ++i) {
            oo_c32 z = cin[i];
            if (!b->abs_approx) {
                float p = fmaf(z.re, z.re, z.im * z.im);
                fout[i] = lpdc_step(&b->ld, p, 1);
            } else {
                float e = fmaf(b->k1, fabsf(z.re), b->k2 * fabsf(z.im));
                fout[i] = lpdc_step(&b->ld, e, 0);
            }
        }
        wr.in_read = n; wr.out_written = n;
        break;

    case K_SSB:                                              /* ssb.rs:28-71 */
        for (size_t i = 0; i < n; ++i) {
            oo_c32 p = rotator_step(&b->rot);
            oo_c32 z = cin[i];
            float y = fmaf(z.re, p.re, z.im * p.im);
            fout[i] = lpdc_step(&b->ld, y, 0);
        }
        wr.in_read = n; wr.out_written = n;
        break;

    case K_CW: {                                             /* cw.rs:34-46 */
        float a = b->alpha;
        for (size_t i = 0; i < n; ++i) {
            float mag = sqrtf(cin[i].re * cin[i].re + cin[i].im * cin[i].im);
            b->y = a * b->y + (1.0f - a) * mag;
            fout[i] = b->y * b->gain;
        }
        wr.in_read = n; wr.out_written = n;
        break;
    }
    case K_HCMF:                                             /* fir.rs:358-371 */
        for (size_t i = 0; i < n; ++i) {
            const size_t len = b->fi.len;
            b->fi.delay[b->fi.idx] = cin[i].re;
            b->fq.delay[b->fi.idx] = cin[i].im;
            float re = 0.0f, im = 0.0f;
            for (size_t t = 0; t < len; ++t) {
                size_t d = (b->fi.idx + len - t) % len;
                float w = b->fi.taps[t];
                re += b->fi.delay[d] * w;
                im += b->fq.delay[d] * w;
            }
            b->fi.idx = (b->fi.idx + 1) % len;
            cout[i].re = re; cout[i].im = im;
        }
        wr.in_read = n; wr.out_written = n;
        break;

    case K_MOD_FM: {                                         /* modulate/fm.rs:45-72 */
        float kf = OO_TAU * b->k / b->fs;
        for (size_t i = 0; i < n; ++i) {
            float dphi = kf * fin[i];
            float ds = sinf(dphi), dc = cosf(dphi);
            float zr = fmaf(b->mz.re, dc, -(b->mz.im * ds));
            float zi = fmaf(b->mz.im, dc, b->mz.re * ds);
            b->mz.re = zr; b->mz.im = zi;
            b->mctr += 1u;
            if ((b->mctr & 0x3FFu) == 0u) {
                float inv = 1.0f / sqrtf(b->mz.re * b->mz.re + b->mz.im * b->mz.im);
                b->mz.re *= inv; b->mz.im *= inv;
            }
            oo_c32 base = { b->mz.re * b->gain, b->mz.im * b->gain };
            cout[i] = mix_nco_step(&b->rot, base);
        }
        wr.in_read = n; wr.out_written = n;
        break;
    }
    case K_MOD_PM:                                           /* modulate/pm.rs:37-47 */
        for (size_t i = 0; i < n; ++i) {
            float phi = b->k * fin[i];
            oo_c32 base = { cosf(phi) * b->gain, sinf(phi) * b->gain };
            cout[i] = mix_nco_step(&b->rot, base);
        }
        wr.in_read = n; wr.out_written = n;
        break;

    case K_MOD_AM:                                           /* modulate/am.rs:44-120 (the 4x unroll is order-preserving) */
        for (size_t i = 0; i < n; ++i) {
            float m = b->carrier + b->mindex * fin[i];
            if (b->clamp) m = oo_minf(oo_maxf(m, -1.0f), 1.0f);
            m = m * b->gain;
            oo_c32 r = rotator_step(&b->rot);
            cout[i].re = m * r.re; cout[i].im = m * r.im;
        }
        wr.in_read = n; wr.out_written = n;
        break;

    case K_MOD_SSB: {                                        /* modulate/ssb.rs:42-114 */
        float side = b->usb ? 1.0f : -1.0f;
        for (size_t i = 0; i < n; ++i) {
            oo_c32 p = rotator_step(&b->rot);
            float ii = biquad_step(&b->bq[1], biquad_step(&b->bq[0], fin[i] * p.re));
            float qq = biquad_step(&b->bq2[1], biquad_step(&b->bq2[0], fin[i] * p.im));
            oo_c32 z = { ii, side * qq };
            oo_c32 r = rotator_step(&b->rot2);
            cout[i].re = fmaf(z.re, r.re, -(z.im * r.im));
            cout[i].im = fmaf(z.im, r.re, z.re * r.im);
        }
        wr.in_read = n; wr.out_written = n;
        break;
    }
    case K_MOD_CW:                                           /* modulate/cw.rs:44-102 */
        for (size_t i = 0; i < n; ++i) {
            float tgt = oo_minf(oo_maxf(fin[i], 0.0f), 1.0f);
            b->env = (tgt >= b->env) ? b->a_rise * b->env + (1.0f - b->a_rise) * tgt
                                     : b->a_fall * b->env + (1.0f - b->a_fall) * tgt;
            oo_c32 base = { b->env * b->gain, 0.0f };
            cout[i] = mix_nco_step(&b->rot, base);
        }
        wr.in_read = n; wr.out_written = n;
        break;

    case K_AGC:                                              /* agc.rs:47-74 */
        if (n == 0) break;
        if (b->env == 0.0f) b->env = oo_maxf(fin[0] * fin[0], 1e-12f);
        for (size_t i = 0; i < n; ++i) {
            float x = fin[i];
            float g = agc_gain_step(b, x * x);
            fout[i] = g * x;
        }
        wr.in_read = n; wr.out_written = n;
        break;

    case K_AGC_IQ:                                           /* agc.rs:121-149 */
        if (n == 0) break;
        if (b->env == 0.0f) b->env = oo_maxf(cin[0].re * cin[0].re + cin[0].im * cin[0].im, 1e-12f);
        for (size_t i = 0; i < n; ++i) {
            oo_c32 x = cin[i];
            float g = agc_gain_step(b, x.re * x.re + x.im * x.im);
            cout[i].re = g * x.re; cout[i].im = g * x.im;
        }
        wr.in_read = n; wr.out_written = n;
        break;

    default: break;
    }
    return wr;
}

/* ---- state snapshot (tests) ----------------------------------------------
 * layout: rotator {z.re,z.im,w.re,w.im,ctr}; biquads {z1,z2}x2; lpdc 6 states;
 * dc {x1,y1}; prev {re,im}; cw {y}.  Always the same 18 floats, unused = 0. */
size_t oo_get_state(const oo_block *b, float *s, size_t cap) {
    const size_t n = 18;
    if (!s || cap < n) return n;
    s[0] = b->rot.z.re; s[1] = b->rot.z.im; s[2] = b->rot.w.re; s[3] = b->rot.w.im;
    s[4] = (float)(b->rot.ctr & 0x3FFu);
    s[5] = b->bq[0].z1; s[6] = b->bq[0].z2; s[7] = b->bq[1].z1; s[8] = b->bq[1].z2;
    s[9] = b->ld.z0_1; s[10] = b->ld.z0_2; s[11] = b->ld.z1_1; s[12] = b->ld.z1_2;
    s[13] = b->ld.dc_x1; s[14] = b->ld.dc_y1;
    if (b->kind == K_DC_BLOCKER) { s[13] = b->dc_x1; s[14] = b->dc_y1; }
    s[15] = b->prev.re; s[16] = b->prev.im; s[17] = b->y;
    return n;
}
