// orion_b200_api.cu -- the C ABI (include/orion_b200.h): block handles, the reference's
// design-time math restated on the host, launch planning, and the host<->device plumbing.
//
// No CPU fallback lives here: every process() call ends in a launch of the sm_100a chain
// kernel (chain_kernels.cuh); without a usable CUDA device the constructors fail.
#include "../../include/orion_b200.h"
#include "chain_args.h"

#include <cuda.h>
#include <cuda_runtime.h>

#include <algorithm>
#include <atomic>
#include <cmath>
#include <condition_variable>
#include <mutex>
#include <thread>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <new>
#include <string>
#include <vector>
#if defined(__AVX2__)
#include <immintrin.h>
#endif

namespace orion {
typedef void (*chain_kernel_t)(const ChainArgs, const CUtensorMap);
chain_kernel_t select_kernel(int front, int R, int U, int sp, int dm, int batch);
cudaError_t chain_kernel_prepare(chain_kernel_t k, size_t dyn_smem, int warps, int *ctas_per_sm);
cudaError_t chain_kernel_launch(chain_kernel_t k, const ChainArgs &args, const CUtensorMap &tmap, int grid, int warps,
                                size_t dyn_smem, cudaStream_t stream, int overlap);
struct BankFirArgs {             // bank_kernels.cu
    const float2 *in; long long n_in, n_out;
    const float2 *hist_in; float2 *hist_out; int H;
    int mix; const NcoParam *osc; unsigned long long kbase;
    int M, Lg, PM, BT, NS;
    const float *gt; float g0;
    float2 *z; long long z_stride; int nch;
    long long tiles_total; int nranges;
    int *err_flag;
};
size_t bank_fir_smem_bytes(const BankFirArgs &a);
int bank_fir_consumer_warps(int nch);
cudaError_t bank_fir_prepare(const BankFirArgs &a, int nwc, int *ctas_per_sm);
cudaError_t bank_fir_launch(const BankFirArgs &a, int nranges, int nwc, cudaStream_t stream);
chain_kernel_t get_kernel_ws(int dm);                       // chain_inst_ws.cu: warp-specialised decimate-by-8 chain
size_t ws_dyn_smem(int nstages, int ntaps2, int Lg);
int ws_warps();
int ws_max_stages(size_t smem_limit, int ntaps2, int Lg);
struct AgcArgs {                 // agc_kernels.cu
    const void *in; void *out; long long n; int iq;
    float attack_a, release_a, target_rms, min_gain, max_gain;
    long long L, W;
    const CarryState *carry_in; CarryState *carry_out;
    NcoParam osc; float gain;    // iq == 2 (CwKeyedMod): tone oscillator and output gain
};
cudaError_t agc_launch(const AgcArgs &a, cudaStream_t stream);
struct SliceArgs { const float2 *in; unsigned char *out; long long n_syms; int bits; float th[15]; };     // aux_kernels.cu
struct FmApplyArgs { const float *x; float2 *out; long long n; float kf, gain; const long long *tile_off; NcoParam rf;
                     const CarryState *carry_in; CarryState *carry_out; };
struct SsbSplitArgs { const float *x; float *xi, *xq; long long n; NcoParam aud; };
struct SsbCombineArgs { const float *yi, *yq; float2 *out; long long n; float side; NcoParam rf; };
cudaError_t gain_c32_launch(const void *in, void *out, long long n, float g, int sms, cudaStream_t st);
cudaError_t slice_launch(const SliceArgs &a, int sms, cudaStream_t st);
cudaError_t fm_mod_launch(const FmApplyArgs &a, long long *d_tile, cudaStream_t st);
cudaError_t ssb_split_launch(const SsbSplitArgs &a, cudaStream_t st);
cudaError_t ssb_combine_launch(const SsbCombineArgs &a, cudaStream_t st);
cudaError_t osc_expand_launch(const OscAnchor *d_an, int n_an, float2 *d_fine, unsigned long long c0, long long fine_len,
                              cudaStream_t stream);
}  // namespace orion

using namespace orion;

// ================================================================================================
// design-time math (host, f32 + libm) -- the reference's constructors, quirks included
// ================================================================================================
namespace {

const float kPi  = 3.14159265358979323846f;
const float kTau = 6.28318530717958647692f;
const float kEps = 1.1920929e-07f;

inline float maxf_rs(float a, float b) { return (a > b || b != b) ? a : b; }   // f32::max

size_t fir_lowpass_ntaps(float fs, float pass_hz, float trans_hz) {            // fir.rs:17-19
    pass_hz = maxf_rs(pass_hz, 10.0f);
    trans_hz = maxf_rs(trans_hz, pass_hz * 0.2f);
    float c = ceilf(fs / trans_hz);
    size_t n = (c > 0.0f) ? (size_t)c : 0;
    if (n < 31) n = 31;
    return n | 1;
}

void fir_lowpass_design(float fs, float pass_hz, float *taps, size_t ntaps) {  // fir.rs:20-44
    pass_hz = maxf_rs(pass_hz, 10.0f);
    const float fc = pass_hz / fs;
    const long m0 = (long)ntaps / 2;
    for (size_t n = 0; n < ntaps; ++n) {
        const long m = (long)n - m0;
        float sinc;
        if (m == 0) sinc = 2.0f * fc;
        else {
            const float x = kPi * (float)m;
            sinc = (2.0f * fc) * sinf(2.0f * kPi * fc * (float)m) / x;
        }
        const float w = 0.5f - 0.5f * cosf(2.0f * kPi * (float)n / ((float)ntaps - 1.0f));
        taps[n] = sinc * w;
    }
    float s = 0.0f;
    for (size_t n = 0; n < ntaps; ++n) s += taps[n];
    for (size_t n = 0; n < ntaps; ++n) taps[n] /= s;
}

float kaiser_beta(float a_db) {                                                // fir.rs:74-82
    if (a_db > 50.0f) return 0.1102f * (a_db - 8.7f);
    if (a_db >= 21.0f) return 0.5842f * powf(a_db - 21.0f, 0.4f) + 0.07886f * (a_db - 21.0f);
    return 0.0f;
}
float bessel_i0(float x) {                                                     // fir.rs:86-99
    const float half = 0.5f * x;
    float term = 1.0f, sum = 1.0f;
    for (unsigned k = 1; k <= 40; ++k) {
        term *= half / (float)k;
        const float t = term * term;
        sum += t;
        if (t < 1e-12f * sum) break;
    }
    return sum;
}
size_t kaiser_len(size_t num_taps) { return (num_taps < 3 ? 3 : num_taps) | 1; }
void kaiser_taps(size_t m, float cutoff_norm, float stopband_db, float *taps) { // fir.rs:113-141
    const float mid = (float)(m / 2);
    float fc = cutoff_norm;
    if (fc < 1e-4f) fc = 1e-4f;
    if (fc > 0.4999f) fc = 0.4999f;
    const float beta = kaiser_beta(stopband_db);
    const float i0b = bessel_i0(beta);
    for (size_t n = 0; n < m; ++n) {
        const float d = (float)n - mid;
        const float ideal = (d == 0.0f) ? 2.0f * fc : sinf(kTau * fc * d) / (kPi * d);
        const float r = d / mid;
        const float w = bessel_i0(beta * sqrtf(maxf_rs(1.0f - r * r, 0.0f))) / i0b;
        taps[n] = ideal * w;
    }
    float s = 0.0f;
    for (size_t n = 0; n < m; ++n) s += taps[n];
    if (fabsf(s) > kEps)
        for (size_t n = 0; n < m; ++n) taps[n] /= s;
}

void lp_biquad_design(float fs, float fc, float c[5]) {                        // iir.rs:49-71
    const float w0 = kTau * fc / fs;
    const float sn = sinf(w0), cs = cosf(w0);
    const float alpha = sn / (2.0f * sqrtf(0.5f));
    const float b0 = (1.0f - cs) * 0.5f, b1 = 1.0f - cs, b2 = (1.0f - cs) * 0.5f;
    const float a0 = 1.0f + alpha, a1 = -2.0f * cs, a2 = 1.0f - alpha;
    const float norm = 1.0f / a0;
    c[0] = b0 * norm; c[1] = b1 * norm; c[2] = b2 * norm; c[3] = a1 * norm; c[4] = a2 * norm;
}
float dc_pole(float fs, float cut_hz) {                                        // dc.rs:15-17
    float r = 1.0f - 2.0f * kPi * (maxf_rs(cut_hz, 0.1f) / fs);
    if (r < 0.0f) r = 0.0f;
    if (r > 0.9999f) r = 0.9999f;
    return r;
}
float cw_alpha(float fs, float env_bw_hz) {                                    // cw.rs:15-18
    return expf(-kTau * maxf_rs(env_bw_hz, 1.0f) / fs);
}

// ---- section groups: dense state-space powers in f64 for the scan tables -------------------------
typedef std::vector<double> Mat;                       // row-major D x D
Mat mat_mul(const Mat &x, const Mat &y, int D) {
    Mat r((size_t)D * D, 0.0);
    for (int i = 0; i < D; ++i)
        for (int k = 0; k < D; ++k) {
            const double xv = x[i * D + k];
            if (xv == 0.0) continue;
            for (int j = 0; j < D; ++j) r[i * D + j] += xv * y[k * D + j];
        }
    return r;
}
Mat mat_pow(Mat base, unsigned long long e, int D) {
    Mat r((size_t)D * D, 0.0);
    for (int i = 0; i < D; ++i) r[i * D + i] = 1.0;
    while (e) {
        if (e & 1ull) r = mat_mul(r, base, D);
        base = mat_mul(base, base, D);
        e >>= 1;
    }
    return r;
}
bool mat_is_zero(const Mat &m) {
    for (double v : m)
        if (std::fabs(v) >= 1e-30) return false;       // far below anything an f32 state sum can resolve
    return true;
}
void mat_store(const Mat &m, float *dst) { for (size_t i = 0; i < m.size(); ++i) dst[i] = (float)m[i]; }

// one step of a section cascade in f64 (the linear model of sec_step_t; state = (s0, s1) per section)
void cascade_step(const SecParam *secs, int count, double *st, double u) {
    double v = u;
    for (int q = 0; q < count; ++q) {
        const SecParam &p = secs[q];
        double &s0 = st[2 * q], &s1 = st[2 * q + 1];
        double y;
        if (p.type == SEC_BIQUAD) {
            y = v * p.c[0] + s0;
            const double n0 = v * p.c[1] + s1 - (double)p.c[3] * y;
            const double n1 = v * p.c[2] - (double)p.c[4] * y;
            s0 = n0; s1 = n1;
        } else if (p.type == SEC_DC) {
            y = v - s0 + (double)p.c[0] * s1;
            s0 = v; s1 = y;
        } else {
            y = (double)p.c[0] * s0 + (double)p.c[1] * v;
            s0 = y; s1 = 0.0;
        }
        v = y;
    }
}

struct GroupHost { int first = 0, count = 0; };
std::vector<GroupHost> split_groups(const std::vector<SecParam> &secs) {
    std::vector<GroupHost> gs;
    GroupHost cur;
    for (size_t s = 0; s < secs.size(); ++s) {
        if (cur.count == 0) cur.first = (int)s;
        cur.count += 1;
        if (secs[s].post_op != OP_NONE || cur.count == kMaxGroupDim / 2) { gs.push_back(cur); cur = GroupHost(); }
    }
    if (cur.count) gs.push_back(cur);
    return gs;
}

// fills the launch-plan data of one group: parameter-bank part (gp) and global tables (gt)
void build_group(const SecParam *secs, const GroupHost &gh, int npt, GroupParam *gp, GroupTables *gt) {
    const int D = 2 * gh.count;
    memset(gp, 0, sizeof(*gp));
    memset(gt, 0, sizeof(*gt));
    gp->first = gh.first; gp->count = gh.count; gp->D = D;
    Mat A((size_t)D * D, 0.0);
    std::vector<double> B(D, 0.0), st(D);
    for (int k = 0; k < D; ++k) {                      // column k of Ac: one step from the unit state e_k
        std::fill(st.begin(), st.end(), 0.0);
        st[k] = 1.0;
        cascade_step(secs + gh.first, gh.count, st.data(), 0.0);
        for (int r = 0; r < D; ++r) A[r * D + k] = st[r];
    }
    std::fill(st.begin(), st.end(), 0.0);
    cascade_step(secs + gh.first, gh.count, st.data(), 1.0);
    B = st;
    const unsigned long long n = (unsigned long long)npt, T = n * kThreads;
    for (int i = 0; i < npt; ++i) {
        const Mat Ak = mat_pow(A, (unsigned long long)(npt - 1 - i), D);
        for (int r = 0; r < D; ++r) {
            double acc = 0.0;
            for (int c = 0; c < D; ++c) acc += Ak[r * D + c] * B[c];
            gp->imp[i][r] = (float)acc;
        }
    }
    for (int l = 0; l < 5; ++l) mat_store(mat_pow(A, n << l, D), gp->lv[l]);
    gp->scan_levels = 5;
    while (gp->scan_levels > 1) {                      // trailing levels whose power is below 2^-40 everywhere
        const Mat P = mat_pow(A, n << (gp->scan_levels - 1), D);
        double mx = 0.0;
        for (double v : P) mx = std::max(mx, std::fabs(v));
        if (mx >= 9.094947017729282e-13) break;
        gp->scan_levels -= 1;
    }
    for (int k = 0; k < 32; ++k) mat_store(mat_pow(A, n * k, D), gt->lane[k]);
    for (int k = 0; k < 32; ++k) mat_store(mat_pow(A, T * k, D), gt->lb[k]);
    mat_store(mat_pow(A, T * 32ull, D), gt->lb32);
    for (int k = 0; k <= 32; ++k) mat_store(mat_pow(A, T * 32ull * (unsigned long long)k, D), gt->lbb[k]);
    mat_store(mat_pow(A, T, D), gt->tile);
    // look-back depth: first k with Ac^(T*k) == 0; geometric search then bisection (the decay of the
    // norm is not strictly monotone for complex poles, so a margin of 4 further powers is verified)
    int depth = 1 << 20;
    {
        int lo = 0, hi = 1;
        while (hi < (1 << 20) && !mat_is_zero(mat_pow(A, T * (unsigned long long)hi, D))) hi <<= 1;
        if (hi < (1 << 20)) {
            while (hi - lo > 1) {
                const int mid = (lo + hi) / 2;
                if (mat_is_zero(mat_pow(A, T * (unsigned long long)mid, D))) hi = mid; else lo = mid;
            }
            bool ok = true;
            for (int e = 1; e <= 4; ++e) ok = ok && mat_is_zero(mat_pow(A, T * (unsigned long long)(hi + e), D));
            if (ok) depth = hi;
        }
    }
    gt->depth = depth;
    gp->agg_only = depth <= 32 ? 1 : 0;
}

// ---- exact-replay oscillator: the reference recurrence walked on the host --------------------------------------
// rotator.rs:44-61 / nco.rs:42-58 restated: z <- (fma(zr, wr, -(zi*wi)), fma(zi, wr, zr*wi)); every 1024 steps
// z *= 1 / sqrt(|z|^2).  The walk is inherently sequential (each step rounds), ~3 ns per step on a host core; it leaves
// one anchor per 1024 steps, from which the device replays in parallel (chain_kernels.cuh).  The sequence is independent
// of the data, so a caller may have it walked ahead of the stream (orion_b200_block_prepare_oscillator).
struct ExactOsc {
    bool enabled = false;
    float wre = 1.f, wim = 0.f;               // the reference's f32 step (cosf(phi), sinf(phi))
    unsigned long long ctr = 0;               // next() calls since the last reset
    float zr = 1.f, zi = 0.f;                 // Z(ctr)
    std::vector<float2> recent;               // the phasors applied to the most recent items (oldest first), for the FIR history
    size_t keep = 0;                          // how many of them to keep
    // device side
    OscAnchor *h_an2[2] = { nullptr, nullptr };   // pinned staging, two sets: the host prepares call N+1 while call N's upload is in flight
    OscAnchor *h_an = nullptr;                // the set in use
    // device tables, two sets like the staging: call N+1's tables are uploaded and expanded on the block's side stream
    // while call N's kernel still reads its own set
    OscAnchor *d_an2[2] = { nullptr, nullptr };
    size_t an_cap = 0;
    float2 *d_fine2[2] = { nullptr, nullptr }, *d_hist2[2] = { nullptr, nullptr };
    size_t fine_cap = 0, hist_cap = 0;
    cudaEvent_t ready2[2] = { nullptr, nullptr };    // set i is uploaded and expanded (side stream)
    cudaEvent_t used2[2] = { nullptr, nullptr };     // the kernel that reads set i has been enqueued before this point of the block's stream
    bool used_valid[2] = { false, false };
    float2 *h_hist2[2] = { nullptr, nullptr };
    float2 *h_hist = nullptr;
    cudaEvent_t staged2[2] = { nullptr, nullptr };   // the last upload from staging set i has completed
    int sbuf = 0;

    // Look-ahead: anc[i] = Z(origin_ctr + 1024 i).  The sequence does not depend on the data, so it may be walked any
    // distance ahead of the items consumed so far (orion_b200_block_prepare_oscillator); a call only copies the anchors
    // its items need.  (ctr, zr, zi) stays the state after the last CONSUMED item.
    unsigned long long origin_ctr = 0;
    std::vector<float2> anc;

    void rebase() {                           // the consumed state becomes the origin of a fresh walk
        origin_ctr = ctr;
        anc.assign(1, make_float2(zr, zi));
    }
    void set_freq(float freq_hz, float fs) {  // rotator.rs:16-18,35-38
        const float phi = kTau * freq_hz / fs;
        wre = cosf(phi); wim = sinf(phi);
        rebase();
    }
    void reset_phase() { zr = 1.f; zi = 0.f; ctr = 0; rebase(); }     // rotator.rs:28-31; the FIR history keeps what it was mixed with
    static inline void step1(float &r, float &i, unsigned long long &c, const float wr, const float wi) {
        const float nr = fmaf(r, wr, -(i * wi));
        const float ni = fmaf(i, wr, r * wi);
        r = nr; i = ni;
        c += 1;
        if ((c & 0x3FFull) == 0) {
            const float r2 = r * r + i * i;
            const float inv = 1.0f / sqrtf(r2);
            r *= inv; i *= inv;
        }
    }
    // anchors up to and including the one at or below counter c
    void ensure(unsigned long long c) {
        if (anc.empty()) rebase();
        const size_t need = (size_t)((c - origin_ctr) >> 10) + 1;
        if (anc.size() >= need) return;
        anc.reserve(need + need / 8);
        float r = anc.back().x, i = anc.back().y;
        unsigned long long cc = origin_ctr + (((unsigned long long)anc.size() - 1ull) << 10);
        const float wr = wre, wi = wim;
        while (anc.size() < need) {
            for (int k = 0; k < 1024; ++k) step1(r, i, cc, wr, wi);
            anc.push_back(make_float2(r, i));
        }
    }
    // Z(c), c >= origin_ctr: replay from the anchor below it
    float2 at(unsigned long long c) {
        ensure(c);
        const size_t a = (size_t)((c - origin_ctr) >> 10);
        float r = anc[a].x, i = anc[a].y;
        unsigned long long cc = origin_ctr + ((unsigned long long)a << 10);
        while (cc < c) step1(r, i, cc, wre, wim);
        return make_float2(r, i);
    }
    // the consumed state moves to counter c (>= ctr); anchors far behind it are dropped
    void consume_to(unsigned long long c) {
        const float2 z = at(c);
        ctr = c; zr = z.x; zi = z.y;
        const size_t a = (size_t)((c - origin_ctr) >> 10);
        if (a >= 65536) {
            anc.erase(anc.begin(), anc.begin() + (long)a);
            origin_ctr += (unsigned long long)a << 10;
        }
    }
    void free_device() {
        for (int i = 0; i < 2; ++i) {
            if (h_an2[i]) cudaFreeHost(h_an2[i]);
            if (h_hist2[i]) cudaFreeHost(h_hist2[i]);
            if (staged2[i]) cudaEventDestroy(staged2[i]);
            h_an2[i] = nullptr; h_hist2[i] = nullptr; staged2[i] = nullptr;
        }
        for (int i = 0; i < 2; ++i) {
            cudaFree(d_an2[i]); cudaFree(d_fine2[i]); cudaFree(d_hist2[i]);
            d_an2[i] = nullptr; d_fine2[i] = nullptr; d_hist2[i] = nullptr;
            if (ready2[i]) cudaEventDestroy(ready2[i]);
            if (used2[i]) cudaEventDestroy(used2[i]);
            ready2[i] = used2[i] = nullptr; used_valid[i] = false;
        }
        h_an = nullptr; h_hist = nullptr;
        an_cap = fine_cap = hist_cap = 0;
    }
};

// ---- oscillator ------------------------------------------------------------------------------
struct Osc {
    bool on = false;
    unsigned long long step = 0, phase0 = 0, k0 = 0;
    float wre = 1.f, wim = 0.f, amp_delta = 0.f;
    ExactOsc x;                                                // the exact-replay twin (same frequency, same resets)
    void set(float freq_hz, float fs, unsigned long long k_now) {
        x.set_freq(freq_hz, fs);
        // keep the phase reached so far (rotator.rs:35-39), then change the step
        phase0 = phase0 + step * (k_now - k0);
        k0 = k_now;
        const float phi = kTau * freq_hz / fs;                 // rotator.rs:17
        const float c = cosf(phi), s = sinf(phi);              // the reference's f32 step w
        const double th = atan2((double)s, (double)c);
        const double turns = th / 6.283185307179586476925286766559;
        if (turns >= 0.5) step = 1ull << 63;
        else step = (unsigned long long)(long long)llround(turns * 18446744073709551616.0);
        amp_delta = (float)(0.5 * log((double)c * (double)c + (double)s * (double)s));
        wre = (float)cos(th);
        wim = (float)sin(th);
        on = true;
    }
    void reset_phase() { phase0 = 0; k0 = 0; x.reset_phase(); }
    NcoParam param(unsigned long long kbase) const {
        NcoParam p;
        memset(&p, 0, sizeof(p));
        p.step = step; p.phase0 = phase0; p.k0 = k0; p.kbase = kbase; p.wre = wre; p.wim = wim;
        p.amp_delta = amp_delta;
        p.xwre = x.wre; p.xwim = x.wim;
        return p;
    }
};

// ---- staged polyphase FIR plan -------------------------------------------------------------------
struct FirPlan {
    int front = FRONT_DIRECT;
    int R = 8, U = 1, Mb = 0, O = 0, P = 0, P_pad = 0, HR = 0, row_samples = 0, row_pitch = 0, rows = 0;
    int H = 0;
    int nstages = 0;
    int warps = kMaxWarpsPerCta;
    size_t stage_bytes = 0;
    size_t dyn_smem = 0;
    std::vector<float> g;          // generic causal taps
    std::vector<float2> taps2;     // [u][q][c]
};

enum : int { AUX_NONE = 0, AUX_GAIN = 1, AUX_SLICE = 2, AUX_FMMOD = 3, AUX_SSBMOD = 4 };
const size_t kMaxStageBytes = 100 * 1024;      // one staged tile
const size_t kRingBudget = 200 * 1024;         // upper bound of the stage ring; finalize_plan trims it to what fits next to the
                                               // per-warp areas and tables (C1: 11 slots -- 45.4 us against 46.3 with 10, 46.8 with 9)

void plan_fir(int fir_kind, const std::vector<float> &taps, size_t M, bool force_global, FirPlan *pl, bool small_tiles = false) {
    const int L = (int)taps.size();
    pl->g.assign(L, 0.f);
    if (fir_kind == FIR_DECIM) {               // y[n] = taps[L-1] x[n] + sum_{t<L-1} taps[t] x[n-1-t]   (fir.rs:57-66)
        pl->g[0] = taps[L - 1];
        for (int k = 1; k < L; ++k) pl->g[k] = taps[k - 1];
    } else {                                   // y[n] = sum_j taps[j] x[n-j]                           (fir.rs:229-247)
        pl->g = taps;
    }
    const int Lg = L;
    const int Mi = (int)M;
    const int U = (Mi % 2 == 0) ? 1 : 2;
    const long long Mb = (long long)Mi * U;
    int O = Mi * (U - 1);
    if (O & 1) O += 1;
    int R = 0, rmax = 8;
    if (const char *e = getenv("ORION_B200_ROWR")) rmax = std::max(1, atoi(e));      // experiment: smaller tiles, more ring slots
    // A chain with an input-rate mixer rotates every staged tile in place before its FIR (stage_mix): the work per slot
    // roughly doubles, so such chains take smaller tiles -- more slots in the ring, more warps with a tile in hand
    // (C2, 201 taps / 25: 28.5 KB tiles and 7 slots -> 15.4 KB and 12; 192 -> 161 us, profiles/r02_c2_steps.txt).
    for (int r : { 8, 4, 2, 1 }) {
        if (r > rmax || Mb * r > kMaxRowSamples) continue;
        if (small_tiles && r > 1) {
            const int Pp = (int)((Lg + 1 + O + Mb - 1) / Mb);
            const int hr = (Pp + r - 1) / r;
            const int rs = (int)(r * Mb);
            const size_t bytes = (size_t)(kThreads + hr) * (size_t)(rs * 8 + (((rs / 2) % 2 == 0) ? 16 : 0));
            if (bytes > 20 * 1024) continue;
        }
        R = r;
        break;
    }
    bool staged = R > 0 && M <= 4096;
    if (staged) {
        const int P = (int)((Lg + 1 + O + Mb - 1) / Mb);
        const int P_pad = ((P + R - 1) / R) * R;
        const int HR = P_pad / R;
        const int row_samples = (int)(R * Mb);
        const int pad = ((row_samples / 2) % 2 == 0) ? 16 : 0;
        const int pitch = row_samples * 8 + pad;
        const int rows = kThreads + HR;
        const size_t table = (size_t)U * (size_t)(Mb / 2) * (size_t)P_pad;
        const size_t smem = (size_t)rows * pitch;
        if (table * 2 > (size_t)kMaxTapTable || rows > 256 || smem > kMaxStageBytes) staged = false;
        else {
            pl->front = FRONT_STAGED;
            pl->R = R; pl->U = U; pl->Mb = (int)Mb; pl->O = O; pl->P = P; pl->P_pad = P_pad; pl->HR = HR;
            const size_t stride = (smem + 127) & ~(size_t)127;       // TMA destinations are 128-byte aligned
            pl->row_samples = row_samples; pl->row_pitch = pitch; pl->rows = rows; pl->stage_bytes = stride;
            pl->nstages = (int)std::min<size_t>(std::max<size_t>(kRingBudget / stride, 1), (size_t)kMaxStages);
            // tuning overrides (experiments): ORION_B200_WARPS = warps per CTA, ORION_B200_STAGES = ring slots
            if (const char *e = getenv("ORION_B200_WARPS")) pl->warps = std::max(1, std::min(kMaxWarpsPerCta, atoi(e)));
            if (const char *e = getenv("ORION_B200_STAGES")) pl->nstages = std::max(1, std::min(kMaxStages, atoi(e)));
            pl->dyn_smem = stride * pl->nstages;       // + the per-warp park area, added in finalize_plan
            pl->taps2.assign(table, make_float2(0.f, 0.f));
            for (int u = 0; u < U; ++u)
                for (int q = 0; q < (int)(Mb / 2); ++q)
                    for (int c = 0; c < P_pad; ++c) {
                        const long long p = P_pad - 1 - c;
                        const long long t0 = Mb * p + (long long)Mi * u - O + 2 * q;   // pairs with sample Mb*b + O - 2q
                        const long long t1 = t0 - 1;                                    // ... and the one after it
                        float2 v;
                        v.x = (t0 >= 0 && t0 < Lg) ? pl->g[t0] : 0.f;
                        v.y = (t1 >= 0 && t1 < Lg) ? pl->g[t1] : 0.f;
                        pl->taps2[((size_t)u * (Mb / 2) + q) * P_pad + c] = v;
                    }
        }
    }
    // history: enough for either front, whichever one runs (the option can flip between calls)
    long long H = Lg;                                            // direct evaluation needs Lg - 1
    if (staged) H = std::max<long long>(H, (long long)pl->row_samples * pl->HR + pl->Mb - 2 - pl->O);
    if (H & 1) H += 1;
    pl->H = (int)H;
    if (!staged || force_global) {
        pl->front = FRONT_GLOBAL;
        pl->R = 8; pl->U = 1; pl->Mb = 0; pl->O = 0; pl->P = pl->P_pad = pl->HR = 0;
        pl->row_samples = pl->row_pitch = pl->rows = 0; pl->dyn_smem = 0; pl->nstages = 0; pl->stage_bytes = 0;
        pl->taps2.clear();
    }
}

typedef CUresult (*encode_tiled_t)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *,
                                   const cuuint64_t *, const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave,
                                   CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
encode_tiled_t get_encode_tiled() {
    static encode_tiled_t fn = nullptr;
    static bool tried = false;
    if (!tried) {
        tried = true;
        void *p = nullptr;
        cudaDriverEntryPointQueryResult qres;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qres) == cudaSuccess &&
            qres == cudaDriverEntryPointSuccess)
            fn = (encode_tiled_t)p;
    }
    return fn;
}

// Streaming copy for the staging ring: the destination is written with non-temporal stores, so the copy moves 2 bytes of
// memory traffic per byte instead of 3 (no read-for-ownership of the destination lines) -- the staging copy of a pageable
// caller buffer is bound by host memory bandwidth, not by PCIe.
static void stream_copy(void *dst, const void *src, size_t bytes) {
#if defined(__AVX2__)
    char *d = (char *)dst;
    const char *s = (const char *)src;
    const size_t head = (32 - ((uintptr_t)d & 31)) & 31;
    if (bytes < 4096 + head) { memcpy(dst, src, bytes); return; }
    if (head) { memcpy(d, s, head); d += head; s += head; bytes -= head; }
    size_t n128 = bytes / 128;
    for (size_t i = 0; i < n128; ++i, d += 128, s += 128) {
        _mm_prefetch(s + 1024, _MM_HINT_NTA);
        const __m256i a = _mm256_loadu_si256((const __m256i *)s), b = _mm256_loadu_si256((const __m256i *)(s + 32));
        const __m256i c = _mm256_loadu_si256((const __m256i *)(s + 64)), e = _mm256_loadu_si256((const __m256i *)(s + 96));
        _mm256_stream_si256((__m256i *)d, a); _mm256_stream_si256((__m256i *)(d + 32), b);
        _mm256_stream_si256((__m256i *)(d + 64), c); _mm256_stream_si256((__m256i *)(d + 96), e);
    }
    _mm_sfence();
    if (bytes & 127) memcpy(d, s, bytes & 127);
#else
    memcpy(dst, src, bytes);
#endif
}

// ---- host copy pool: pageable caller buffers are moved to / from the pinned staging ring by a few threads ------
class CopyPool {
public:
    static CopyPool &get() { static CopyPool p; return p; }
    void copy(void *dst, const void *src, size_t bytes) {
        if (bytes < (1u << 20) || workers_.empty()) { stream_copy(dst, src, bytes); return; }
        std::lock_guard<std::mutex> job(job_m_);                  // one job at a time (blocks on different threads share the pool)
        const size_t parts = workers_.size() + 1;
        const size_t per = ((bytes / parts) + 4095) & ~(size_t)4095;
        {
            std::lock_guard<std::mutex> lk(m_);
            dst_ = (char *)dst; src_ = (const char *)src; bytes_ = bytes; per_ = per;
            next_ = 1; pending_ = (int)workers_.size(); ++gen_;
        }
        cv_.notify_all();
        stream_copy(dst, src, std::min(per, bytes));               // the caller's share
        std::unique_lock<std::mutex> lk(m_);
        done_.wait(lk, [this] { return pending_ == 0; });
    }
private:
    CopyPool() {
        unsigned hw = std::thread::hardware_concurrency();
        int n = hw >= 16 ? 7 : (hw >= 8 ? 3 : (hw >= 4 ? 1 : 0));
        if (const char *e = getenv("ORION_B200_COPY_THREADS")) n = std::max(0, std::min(15, atoi(e) - 1));
        for (int i = 0; i < n; ++i) workers_.emplace_back([this] { run(); });
    }
    ~CopyPool() {
        { std::lock_guard<std::mutex> lk(m_); stop_ = true; ++gen_; }
        cv_.notify_all();
        for (std::thread &t : workers_) t.join();
    }
    void run() {
        unsigned long long seen = 0;
        for (;;) {
            size_t part;
            {
                std::unique_lock<std::mutex> lk(m_);
                cv_.wait(lk, [&] { return gen_ != seen; });
                seen = gen_;
                if (stop_) return;
                part = next_++;
            }
            const size_t off = part * per_;
            if (off < bytes_) stream_copy(dst_ + off, src_ + off, std::min(per_, bytes_ - off));
            {
                std::lock_guard<std::mutex> lk(m_);
                if (--pending_ == 0) done_.notify_all();
            }
        }
    }
    std::vector<std::thread> workers_;
    std::mutex m_, job_m_;
    std::condition_variable cv_, done_;
    char *dst_ = nullptr; const char *src_ = nullptr;
    size_t bytes_ = 0, per_ = 0, next_ = 0;
    int pending_ = 0;
    unsigned long long gen_ = 0;
    bool stop_ = false;
};
bool is_pageable(const void *p) {
    cudaPointerAttributes at;
    if (cudaPointerGetAttributes(&at, p) != cudaSuccess) { cudaGetLastError(); return true; }
    return at.type == cudaMemoryTypeUnregistered;
}

thread_local int t_device = 0;
thread_local std::string t_create_error;       // why the last constructor on this thread failed

}  // namespace

// ================================================================================================
// the block handle
// ================================================================================================
struct orion_b200_block {
    // ---- specification ----
    int in_item = ORION_B200_ITEM_C32, out_item = ORION_B200_ITEM_C32;
    int mix = MIX_NONE;
    int fir = FIR_NONE;
    std::vector<float> taps;
    size_t M = 1;
    int demod = DEMOD_NONE;
    int translate = 0;
    float k = 0.f, k1 = 0.f, k2 = 0.f, k3 = 0.f;
    float fs_demod = 0.f;
    std::vector<SecParam> secs;
    Osc pre, post;
    int nbatch = 1;                       // > 1: this block runs nbatch independent, equally long streams per call (a bank's
    int *d_batch_chan = nullptr;          //      demodulator group); member m reads / writes row d_batch_chan[m] of strided buffers
    int opt_exact = -1;                   // -1 auto (by block kind), 0 closed form everywhere, 1 exact everywhere
    double exact_host_ms = 0.0;           // host time spent walking the recurrence (reported separately from kernel time)
    int cw_gain_sec = -1;
    int agc = 0;                          // 1: AgcRms (f32), 2: AgcRmsIq (C32), 3: CwKeyedMod -- agc_kernels.cu instead of the chain kernel
    int aux = 0;                          // AUX_*: the elementwise / scan kernels of aux_kernels.cu
    int slice_bits = 0;                   // AUX_SLICE: 1 BPSK, 2 QPSK, 4 / 6 / 8 QAM
    float slice_th[15] = { 0 };
    float aux_gain = 1.0f, aux_p0 = 0.f, aux_side = 1.0f;
    long long *d_fm_tiles = nullptr; size_t fm_tiles_cap = 0;
    orion_b200_block *child[2] = { nullptr, nullptr };   // AUX_SSBMOD: the two LpCascade filters
    float *d_ssb[4] = { nullptr, nullptr, nullptr, nullptr }; size_t ssb_cap = 0;
    float agc_attack_a = 0.f, agc_release_a = 0.f, agc_target = 0.f, agc_min_gain = 0.05f, agc_max_gain = 20.0f;
    // ---- plan ----
    FirPlan plan;
    bool plan_dirty = true;
    chain_kernel_t kernel = nullptr;
    int ctas_per_sm = 1, sm_count = 1;
    int ws = 0;                           // the warp-specialised instance is selected
    int pipe_park_slots = 2, pipe_u_slots = 0;   // per-warp pipeline area of the chain kernel (multi-group chains)
    int nstage = 0, stage_group[8] = { 0 };     // pipeline stages: the group each one finishes (-1: idle)
    int split = 1;                        // warps that share the FIR of one tile (FIR-only staged instance, long filters)
    // ---- options ----
    int opt_force_global = 0, opt_use_tma = 1, opt_serial = 0, opt_overlap = 0;
    long long *trace = nullptr;           // debug: device buffer of 8 x int64 per tile
    // ---- device ----
    int device = 0;
    cudaStream_t own_stream = nullptr, stream = nullptr;
    cudaStream_t osc_stream = nullptr;    // exact oscillators: table upload + expansion run here, ahead of the block's stream
    int osc_side = 1;                     // 1: side stream; 0: the block's own stream (ORION_B200_OSC_STREAM)
    std::vector<cudaEvent_t> osc_used_pending;   // recorded on `stream` right after the launch that reads the tables (launch())
    float *d_g = nullptr;
    GroupTables *d_gtabs = nullptr;
    std::vector<GroupParam> groups;
    std::vector<int> group_depth;
    // FIR history and carried state rotate through THREE buffers: call N reads [pp] and writes [pp+1]; the overlapped
    // call N+1 writes [pp+2], which nobody has read since call N-1 (over before N+1 writes: ChainArgs::depth_target)
    float2 *d_hist[3] = { nullptr, nullptr, nullptr };
    size_t hist_cap = 0;
    CarryState *d_carry[3] = { nullptr, nullptr, nullptr };
    int pp = 0;
    unsigned int ctas_par[2] = { 0, 0 };  // CTAs launched so far by the calls of each parity: targets of the CTA-done counters handoff[2 + parity]
    TileLink *d_links = nullptr;
    size_t links_cap = 0;
    unsigned long long *d_ticket = nullptr;
    unsigned long long ticket_base = 0;
    unsigned epoch = 0;
    int *d_err = nullptr;
    int *h_err = nullptr;                 // pinned
    int *d_err_ext = nullptr;             // a channel bank's shared watchdog word (not owned)
    unsigned int *d_handoff = nullptr;    // hand-over counters between consecutive calls (chain_kernels.cuh)
    unsigned int calls_since_reset = 0;
    void *d_in = nullptr, *d_out = nullptr;
    size_t d_in_cap = 0, d_out_cap = 0;
    // host-pointer calls are pipelined in chunks: copies in, kernels and copies out overlap (orion_b200_block_process)
    cudaStream_t s_h2d = nullptr, s_d2h = nullptr;
    static const int kPipeSlots = 3;
    cudaEvent_t ev_h2d[kPipeSlots] = { nullptr, nullptr, nullptr }, ev_k[kPipeSlots] = { nullptr, nullptr, nullptr },
                ev_d2h[kPipeSlots] = { nullptr, nullptr, nullptr };
    void *h_stage_in[kPipeSlots] = { nullptr, nullptr, nullptr }, *h_stage_out[kPipeSlots] = { nullptr, nullptr, nullptr };
    size_t h_stage_in_cap = 0, h_stage_out_cap = 0;      // pinned staging for pageable callers
    // ---- counters ----
    unsigned long long k_pre = 0, k_post = 0;     // items consumed at the input rate / demod rate since reset
    uint64_t launches = 0;
    std::string err;
};

namespace {

int fail(orion_b200_block *b, int status, const char *what, cudaError_t e = cudaSuccess) {
    if (b) {
        b->err = what;
        if (e != cudaSuccess) { b->err += ": "; b->err += cudaGetErrorString(e); }
    }
    return status;
}
#define CK(call) do { cudaError_t _e = (call); if (_e != cudaSuccess) return fail(b, ORION_B200_ERR_CUDA, #call, _e); } while (0)

// Device memory is initialised ON THE BLOCK'S STREAM.  The legacy default stream does not order against the
// non-blocking streams the kernels run on: a 45 MB cudaMemset of the link records, queued on it behind a caller's own
// default-stream work, was seen to run while launches were already publishing into those records -- and wiped them.
cudaError_t dev_memset(orion_b200_block *b, void *p, int v, size_t n) { return cudaMemsetAsync(p, v, n, b->stream); }
cudaError_t dev_upload(orion_b200_block *b, void *dst, const void *src, size_t n) {      // src may be a temporary
    cudaError_t e = cudaMemcpyAsync(dst, src, n, cudaMemcpyHostToDevice, b->stream);
    return e == cudaSuccess ? cudaStreamSynchronize(b->stream) : e;
}
cudaError_t dev_copy(orion_b200_block *b, void *dst, const void *src, size_t n) {
    return cudaMemcpyAsync(dst, src, n, cudaMemcpyDeviceToDevice, b->stream);
}

size_t item_bytes(int item) { return item == ORION_B200_ITEM_C32 ? 8 : (item == ORION_B200_ITEM_U8 ? 1 : 4); }
size_t in_item_bytes(const orion_b200_block *b) { return item_bytes(b->in_item); }
size_t out_item_bytes(const orion_b200_block *b) { return item_bytes(b->out_item); }

int npt_of(const orion_b200_block *b) { return b->plan.R * b->plan.U; }

// (re)build everything derived from the specification and upload it
int finalize_plan(orion_b200_block *b) {
    CK(cudaSetDevice(b->device));
    b->plan.warps = kMaxWarpsPerCta;               // (a previous plan may have been the 24-warp warp-specialised one)
    if (b->fir != FIR_NONE) plan_fir(b->fir, b->taps, b->M, b->opt_force_global != 0, &b->plan, b->mix != MIX_NONE && !getenv("ORION_B200_BIG_TILES"));
    else { b->plan = FirPlan(); b->plan.front = FRONT_DIRECT; b->plan.R = 16; b->plan.U = 1; }   // chain_kernel<DIRECT,16,1>
    // shape / demodulator specialisations of the kernel family (chain_kernels.cuh, Geo<SP> and Dm<DM>)
    int sp = 0, dm = -1;
    const bool lr4_only = b->secs.size() == 2 && b->secs[0].type == SEC_BIQUAD && b->secs[1].type == SEC_BIQUAD &&
                          b->secs[0].post_op == OP_NONE && b->secs[1].post_op == OP_NONE;
    if (b->plan.front == FRONT_STAGED && b->plan.R == 8 && b->plan.U == 1 && b->plan.Mb == 8 && b->plan.HR == 1 &&
        b->plan.P_pad == 8 && b->plan.row_pitch == 528) {
        sp = 1;
        if (b->demod == DEMOD_NONE) dm = DEMOD_NONE;
        else if (b->demod == DEMOD_FM && lr4_only) dm = 100 + DEMOD_FM;              // DM_LR4 + kind
        else if (b->demod == DEMOD_AM) dm = DEMOD_AM;
    } else if (b->plan.front == FRONT_STAGED && b->plan.R == 4 && b->plan.U == 1 && b->plan.Mb == 32 && b->plan.HR == 8 &&
               b->plan.P_pad == 32 && b->plan.row_pitch == 1040 && b->demod == DEMOD_NONE && b->mix == MIX_NONE) {
        sp = 2; dm = DEMOD_NONE;                                                     // C4 shape, FIR alone
    } else if (b->plan.front == FRONT_DIRECT) {
        if (b->demod == DEMOD_NONE) dm = DEMOD_NONE;
        else if ((b->demod == DEMOD_FM || b->demod == DEMOD_PM || b->demod == DEMOD_F32) && lr4_only) dm = 100 + b->demod;
    }
    if (getenv("ORION_B200_NO_SPECIALIZE")) { sp = 0; dm = -1; }
    if (b->nbatch > 1) {                           // one CTA per member: 4 warps when the members fill the machine, more for small banks
        int w = 4;
        while (w < kMaxWarpsPerCta && (long long)b->nbatch * (2 * w) <= 148LL * kMaxWarpsPerCta) w *= 2;
        b->plan.warps = w;
    }
    b->kernel = select_kernel(b->plan.front, b->plan.R, b->plan.U, sp, dm, b->nbatch > 1);
    if (!b->kernel) return fail(b, ORION_B200_ERR_INTERNAL, "no kernel instance for plan");
    // section groups + scan tables
    b->groups.clear();
    b->group_depth.clear();
    if (!b->secs.empty()) {
        const std::vector<GroupHost> gh = split_groups(b->secs);
        if (gh.size() > (size_t)kMaxGroups) return fail(b, ORION_B200_ERR_UNSUPPORTED, "too many section groups");
        std::vector<GroupTables> tabs(gh.size());
        b->groups.resize(gh.size());
        for (size_t g = 0; g < gh.size(); ++g) build_group(b->secs.data(), gh[g], npt_of(b), &b->groups[g], &tabs[g]);
        b->group_depth.clear();
        for (size_t g = 0; g < gh.size(); ++g) b->group_depth.push_back(tabs[g].depth);
        if (b->d_gtabs) { cudaFree(b->d_gtabs); b->d_gtabs = nullptr; }
        CK(cudaMalloc(&b->d_gtabs, tabs.size() * sizeof(GroupTables)));
        CK(dev_upload(b, b->d_gtabs, tabs.data(), tabs.size() * sizeof(GroupTables)));
    }
    // per-warp pipeline area (chain_kernels.cuh): chains with several section groups keep one tile per group in flight
    // Stage schedule of the pipeline: stage k finishes group stage_group[k-1] (and forms the aggregate of the next one); a
    // slow-pole group CAN get an idle stage in front of its finish (ORION_B200_IDLE_STAGE=1), so that its block records --
    // which wait for the slowest of 32 concurrent tiles -- have two iterations to arrive instead of one.  Experiment only:
    // measured on C3 it tripped the look-back watchdog (profiles/r02_experiments.txt), so the default schedule has no idle stage.
    {
        const int ng = (int)b->groups.size();
        const bool multi = ng >= 2 && dm < 100;
        b->nstage = 0;
        if (multi)
            for (int g = 0; g < ng; ++g) {
                if (g > 0 && !b->groups[g].agg_only && getenv("ORION_B200_IDLE_STAGE")) b->stage_group[b->nstage++] = -1;
                b->stage_group[b->nstage++] = g;
            }
        b->pipe_park_slots = multi ? b->nstage + 1 : 2;
        b->pipe_u_slots = multi ? b->nstage : 0;
    }
    const size_t warp_pipe = 16 + (size_t)b->pipe_park_slots * 33 * kMaxGroupDim * sizeof(float) +
                             (size_t)b->pipe_u_slots * kThreads * npt_of(b) * sizeof(float);
    // long filters on the FIR-only staged instance: several warps share the FIR of one tile (slices of the tap rows), so
    // that more than one warp per ring slot does arithmetic (the 1023-tap /32 shape fits only four 41.6 KB slots)
    b->split = 1;
    if (b->plan.front == FRONT_STAGED && b->demod == DEMOD_NONE && b->mix == MIX_NONE && b->nbatch == 1 && b->plan.HR >= 4 && !b->opt_serial)
        b->split = (b->plan.HR % 4 == 0) ? 4 : ((b->plan.HR % 2 == 0) ? 2 : 1);
    if (const char *e = getenv("ORION_B200_SPLIT")) {
        const int v = atoi(e);
        if ((v == 1 || v == 2 || v == 4) && b->plan.front == FRONT_STAGED && b->demod == DEMOD_NONE && b->mix == MIX_NONE && b->plan.HR % v == 0) b->split = v;
    }
    const size_t part_bytes = b->split > 1 ? (size_t)3 * kThreads * npt_of(b) * sizeof(float2) : 0;     // per ring slot
    auto rest_of = [&](int warps) {
        return (size_t)warps * warp_pipe + (size_t)b->plan.nstages * part_bytes + b->plan.taps2.size() * sizeof(float2) +
               ((b->plan.g.size() * sizeof(float) + 15) & ~(size_t)15) +
               sizeof(GroupParam) * kMaxGroups + sizeof(SecParam) * kMaxSections + 32 +
               (2 * 32 * 16 + kMaxNpt * 4) * sizeof(float) +
               (b->plan.front == FRONT_DIRECT ? (size_t)warps * 32 * 144 + 16 : 0);
    };
    const size_t smem_limit = (size_t)(227 - 3) * 1024;
    // deep pipelines of wide tiles: fewer ring slots first (down to 4), then fewer warps
    while (b->plan.nstages > 4 && b->plan.stage_bytes * b->plan.nstages + rest_of(b->plan.warps) > smem_limit) b->plan.nstages -= 1;
    while (b->plan.warps > 4 && b->plan.stage_bytes * b->plan.nstages + rest_of(b->plan.warps) > smem_limit) b->plan.warps -= 2;
    while (b->plan.nstages > 1 && b->plan.stage_bytes * b->plan.nstages + rest_of(b->plan.warps) > smem_limit) b->plan.nstages -= 1;
    b->plan.dyn_smem = b->plan.stage_bytes * b->plan.nstages +
                       (size_t)b->plan.warps * warp_pipe +                                  // stage ring + per-warp pipeline area
                       b->plan.taps2.size() * sizeof(float2) +                              // + tap table
                       ((b->plan.g.size() * sizeof(float) + 15) & ~(size_t)15) +            // + generic taps
                       sizeof(GroupParam) * kMaxGroups + sizeof(SecParam) * kMaxSections + 32 +  // + section/group data
                       (2 * 32 * 16 + kMaxNpt * 4) * sizeof(float) +                             // + per-lane scan tables (LR4 instance)
                       (size_t)b->plan.nstages * part_bytes +                                    // + partial sums of the tap-split FIR
                       (b->plan.front == FRONT_DIRECT ? (size_t)b->plan.warps * 32 * 144 + 16 : 0);  // + transposing scratch (rate-1 blocks)
    // FIR /8 + FM | PM + LR4 (the C1 chain): the warp-specialised instance (chain_inst_ws.cu) is an experiment, selected with
    // ORION_B200_WS=1 only -- measured 4-7 % SLOWER than the unified kernel on B200 (profiles/r02_experiments.txt)
    b->ws = 0;
    if (sp == 1 && (dm == 100 + DEMOD_FM || dm == 100 + DEMOD_PM) && b->nbatch == 1 && !b->opt_serial && getenv("ORION_B200_WS")) {
        chain_kernel_t kw = get_kernel_ws(dm);
        if (kw) {
            b->kernel = kw;
            b->ws = 1;
            b->plan.warps = ws_warps();
            int ns = ws_max_stages(227 * 1024, (int)b->plan.taps2.size(), (int)b->plan.g.size());
            if (const char *e = getenv("ORION_B200_STAGES")) ns = std::max(1, std::min(ns, atoi(e)));
            b->plan.nstages = ns;
            b->plan.dyn_smem = ws_dyn_smem(ns, (int)b->plan.taps2.size(), (int)b->plan.g.size());
        }
    }
    CK(chain_kernel_prepare(b->kernel, b->plan.dyn_smem, b->plan.warps, &b->ctas_per_sm));
    if (b->ctas_per_sm < 1) return fail(b, ORION_B200_ERR_INTERNAL, "kernel does not fit on an SM");
    // FIR taps + history
    if (b->fir != FIR_NONE) {
        if (b->d_g) { cudaFree(b->d_g); b->d_g = nullptr; }
        CK(cudaMalloc(&b->d_g, b->plan.g.size() * sizeof(float)));
        CK(dev_upload(b, b->d_g, b->plan.g.data(), b->plan.g.size() * sizeof(float)));
        if ((size_t)b->plan.H > b->hist_cap) {
            // growing the history keeps the most recent samples at the end
            for (int i = 0; i < 3; ++i) {
                float2 *nh = nullptr;
                CK(cudaMalloc(&nh, (size_t)b->plan.H * sizeof(float2)));
                CK(dev_memset(b, nh, 0, (size_t)b->plan.H * sizeof(float2)));
                if (b->d_hist[i]) {
                    CK(dev_copy(b, nh + (b->plan.H - b->hist_cap), b->d_hist[i], b->hist_cap * sizeof(float2)));
                    CK(cudaStreamSynchronize(b->stream));
                    cudaFree(b->d_hist[i]);
                }
                b->d_hist[i] = nh;
            }
            b->hist_cap = (size_t)b->plan.H;
        }
    }
    b->plan_dirty = false;
    return ORION_B200_OK;
}

int reset_state(orion_b200_block *b) {
    CK(cudaSetDevice(b->device));
    CK(cudaStreamSynchronize(b->stream));
    CarryState cs;
    memset(&cs, 0, sizeof(cs));
    cs.prev = make_float2(1.0f, 0.0f);                         // fm.rs:29, pm.rs:29
    std::vector<CarryState> csv((size_t)b->nbatch, cs);
    for (int i = 0; i < 3; ++i) {
        CK(dev_upload(b, b->d_carry[i], csv.data(), csv.size() * sizeof(CarryState)));
        if (b->d_hist[i]) CK(dev_memset(b, b->d_hist[i], 0, b->hist_cap * sizeof(float2)));
    }
    CK(dev_memset(b, b->d_handoff, 0, 4 * sizeof(unsigned int)));
    CK(cudaStreamSynchronize(b->stream));
    b->calls_since_reset = 0;
    b->ctas_par[0] = b->ctas_par[1] = 0;
    b->k_pre = b->k_post = 0;
    b->pre.reset_phase();
    b->post.reset_phase();
    b->pre.x.recent.clear();
    b->post.x.recent.clear();
    for (orion_b200_block *c : b->child)
        if (c) { const int st = reset_state(c); if (st != ORION_B200_OK) return st; }
    return ORION_B200_OK;
}

int new_block(orion_b200_block **out, orion_b200_block **pb) {
    if (!out) return ORION_B200_ERR_INVALID;
    *out = nullptr;
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev <= 0) { cudaGetLastError(); return ORION_B200_ERR_NO_DEVICE; }
    orion_b200_block *b = new (std::nothrow) orion_b200_block();
    if (!b) return ORION_B200_ERR_ALLOC;
    b->device = t_device;
    *pb = b;
    return ORION_B200_OK;
}

int init_device_side(orion_b200_block *b) {
    CK(cudaSetDevice(b->device));
    cudaDeviceProp prop;
    CK(cudaGetDeviceProperties(&prop, b->device));
    b->sm_count = prop.multiProcessorCount;
    CK(cudaStreamCreateWithFlags(&b->own_stream, cudaStreamNonBlocking));
    b->stream = b->own_stream;
    for (int i = 0; i < 3; ++i) CK(cudaMalloc(&b->d_carry[i], (size_t)b->nbatch * sizeof(CarryState)));
    CK(cudaMalloc(&b->d_ticket, 2 * sizeof(unsigned long long)));        // {ticket, done}
    CK(dev_memset(b, b->d_ticket, 0, 2 * sizeof(unsigned long long)));
    CK(cudaMalloc(&b->d_handoff, 4 * sizeof(unsigned int)));
    CK(dev_memset(b, b->d_handoff, 0, 4 * sizeof(unsigned int)));
    CK(cudaMalloc(&b->d_err, sizeof(int)));
    CK(dev_memset(b, b->d_err, 0, sizeof(int)));
    CK(cudaMallocHost(&b->h_err, sizeof(int)));
    *b->h_err = 0;
    int st = finalize_plan(b);
    if (st != ORION_B200_OK) return st;
    return reset_state(b);
}

int finish_create(orion_b200_block *b, orion_b200_block **out) {
    if (b->secs.size() > (size_t)kMaxSections) { delete b; return ORION_B200_ERR_UNSUPPORTED; }
    if (!b->aux && b->agc != 3) {
        b->in_item = kind_f32_in(b->demod) ? ORION_B200_ITEM_F32 : ORION_B200_ITEM_C32;
        b->out_item = kind_c32_out(b->demod) ? ORION_B200_ITEM_C32 : ORION_B200_ITEM_F32;
    }
    int st = init_device_side(b);
    if (st != ORION_B200_OK) {
        t_create_error = b->err;                 // orion_b200_last_create_error()
        orion_b200_block_destroy(b);
        return st;
    }
    *out = b;
    return ORION_B200_OK;
}

SecParam sec_biquad(const float c[5]) {
    SecParam p; memset(&p, 0, sizeof(p));
    p.type = SEC_BIQUAD; p.post_op = OP_NONE;
    for (int i = 0; i < 5; ++i) p.c[i] = c[i];
    return p;
}
SecParam sec_dc(float r) {
    SecParam p; memset(&p, 0, sizeof(p));
    p.type = SEC_DC; p.c[0] = r;
    return p;
}
void add_lr4(orion_b200_block *b, float fs, float fc) {
    float c[5];
    lp_biquad_design(fs, fc, c);
    b->secs.push_back(sec_biquad(c));
    b->secs.push_back(sec_biquad(c));
}

void length_rules(const orion_b200_block *b, size_t n_in, size_t out_cap, size_t *consume, size_t *produce) {
    if (b->aux == AUX_SLICE) {                         // qpsk.rs:70 / qam.rs:150: n_syms = min(len(in), len(out) / BITS)
        const size_t n = std::min(n_in, out_cap / (size_t)b->slice_bits);
        *consume = n;
        *produce = n * (size_t)b->slice_bits;
        return;
    }
    if (b->fir == FIR_DECIM || (b->fir != FIR_NONE && b->M > 1)) {   // decim.rs:45,66-75 (any m, m = 1 included): all input read, ceil(n/m) capped
        *consume = n_in;
        const size_t n_out = (n_in + b->M - 1) / b->M;
        *produce = std::min(n_out, out_cap);
    } else {                                           // rate-1: n = min(len(in), len(out))
        const size_t n = std::min(n_in, out_cap);
        *consume = n;
        *produce = n;
    }
}

// Where the absolute oscillator phase reaches the output the reference's own recurrence is replayed (bit-exact
// phasors); where only phase differences or magnitudes matter (FM / PM / AM / CW behind the mixer, FM translate) the
// closed form is kept: a slowly drifting common rotation cancels there (DESIGN.md "oscillator").
bool want_exact_pre(const orion_b200_block *b) {
    if (b->mix == MIX_NONE) return false;
    if (b->opt_exact >= 0) return b->opt_exact != 0;
    return b->demod == DEMOD_NONE || b->demod == DEMOD_SSB || b->demod == DEMOD_USB;
}
bool want_exact_post(const orion_b200_block *b) {
    if (!(b->demod == DEMOD_SSB || b->demod == DEMOD_USB || b->demod == MOD_AM || b->demod == MOD_PM)) return false;
    return b->opt_exact < 0 || b->opt_exact != 0;
}

// Walk the reference recurrence over the n_items of this call (host, sequential), leave one anchor per 1024 items and
// the phasors of the previous hist_len items, and enqueue the expansion kernel.  Fills the exact-mode fields of *np.
int prepare_exact(orion_b200_block *b, Osc &o, unsigned long long kbase, size_t n_items, size_t hist_len, NcoParam *np) {
    ExactOsc &x = o.x;
    if (x.ctr != kbase)
        return fail(b, ORION_B200_ERR_INVALID, "exact oscillator mode must be selected before the first call after a reset");
    if (n_items == 0) return ORION_B200_OK;
    const size_t n_an = (n_items + 1023) / 1024 + 2;      // capacity: the consumed state + one per 1024-grid point
    size_t n_an_used = 0;
    const long long fine_len = (long long)((n_items - 1) >> 4) + 1;
    x.sbuf ^= 1;                                               // staging + table set of this call; its previous upload (two calls ago) must be done
    const int p = x.sbuf;
    if (!b->osc_stream) CK(cudaStreamCreateWithFlags(&b->osc_stream, cudaStreamNonBlocking));
    // Measured (profiles/r02_c2_steps.txt): the side stream hides the uploads of blocks with ONE exact oscillator (Rotator
    // 133 -> 101 us, FIR + SSB 161 -> 148 us); a chain with two (C2: input-rate mixer + BFO) is slower with it (288 -> 363 us:
    // its large expansion kernel competes with the SM-filling chain kernel for a place), so such chains stay in-stream.
    b->osc_side = !(want_exact_pre(b) && want_exact_post(b));
    if (const char *e = getenv("ORION_B200_OSC_STREAM")) b->osc_side = atoi(e);
    if (!x.staged2[p]) CK(cudaEventCreateWithFlags(&x.staged2[p], cudaEventDisableTiming));
    else CK(cudaEventSynchronize(x.staged2[p]));
    if (!x.ready2[p]) CK(cudaEventCreateWithFlags(&x.ready2[p], cudaEventDisableTiming));
    if (!x.used2[p]) CK(cudaEventCreateWithFlags(&x.used2[p], cudaEventDisableTiming));
    if (n_an > x.an_cap) {
        CK(cudaStreamSynchronize(b->stream));
        CK(cudaStreamSynchronize(b->osc_stream));
        const size_t cap = n_an + n_an / 4 + 16;
        for (int i = 0; i < 2; ++i) {
            if (x.h_an2[i]) cudaFreeHost(x.h_an2[i]);
            x.h_an2[i] = nullptr;
            cudaFree(x.d_an2[i]); x.d_an2[i] = nullptr;
            CK(cudaMallocHost(&x.h_an2[i], cap * sizeof(OscAnchor)));
            CK(cudaMalloc(&x.d_an2[i], cap * sizeof(OscAnchor)));
        }
        x.an_cap = cap;
    }
    if ((size_t)fine_len > x.fine_cap) {
        CK(cudaStreamSynchronize(b->stream));
        CK(cudaStreamSynchronize(b->osc_stream));
        const size_t cap = (size_t)fine_len + (size_t)fine_len / 4 + 64;
        for (int i = 0; i < 2; ++i) {
            cudaFree(x.d_fine2[i]); x.d_fine2[i] = nullptr;
            CK(cudaMalloc(&x.d_fine2[i], cap * sizeof(float2)));
        }
        x.fine_cap = cap;
    }
    if (hist_len > x.hist_cap) {
        CK(cudaStreamSynchronize(b->stream));
        CK(cudaStreamSynchronize(b->osc_stream));
        for (int i = 0; i < 2; ++i) {
            if (x.h_hist2[i]) cudaFreeHost(x.h_hist2[i]);
            x.h_hist2[i] = nullptr;
            cudaFree(x.d_hist2[i]); x.d_hist2[i] = nullptr;
            CK(cudaMallocHost(&x.h_hist2[i], hist_len * sizeof(float2)));
            CK(cudaMalloc(&x.d_hist2[i], hist_len * sizeof(float2)));
        }
        x.hist_cap = hist_len;
    }
    x.h_an = x.h_an2[x.sbuf];
    x.h_hist = x.h_hist2[x.sbuf];
    // phasors of the items before this call (newest last)
    const size_t nh = std::min(x.recent.size(), hist_len);
    for (size_t i = 0; i < nh; ++i) x.h_hist[i] = x.recent[x.recent.size() - nh + i];
    // the walk (nothing to do for the part orion_b200_block_prepare_oscillator has covered already)
    timespec t0, t1;
    clock_gettime(CLOCK_MONOTONIC, &t0);
    const unsigned long long ctr0 = x.ctr;
    const unsigned long long c_last = ctr0 + 1ull + 16ull * (unsigned long long)(fine_len - 1);   // last counter the expansion writes
    x.ensure(ctr0 + (unsigned long long)n_items);
    clock_gettime(CLOCK_MONOTONIC, &t1);
    b->exact_host_ms += (t1.tv_sec - t0.tv_sec) * 1e3 + (t1.tv_nsec - t0.tv_nsec) * 1e-6;
    // device anchors: anchor 0 is the consumed state itself, the others sit on the walk's 1024-grid
    {
        size_t k = 0;
        OscAnchor &A0 = x.h_an[k++];
        const unsigned long long g1 = x.origin_ctr + ((((ctr0 - x.origin_ctr) >> 10) + 1ull) << 10);   // first grid point above ctr0
        A0.ctr = ctr0; A0.z = make_float2(x.zr, x.zi); A0.w = make_float2(x.wre, x.wim);
        A0.nsteps = (unsigned)(std::min(g1, c_last) - ctr0); A0.pad = 0;
        for (unsigned long long g = g1; g < c_last; g += 1024ull) {
            OscAnchor &A = x.h_an[k++];
            A.ctr = g; A.z = x.anc[(size_t)((g - x.origin_ctr) >> 10)]; A.w = make_float2(x.wre, x.wim);
            A.nsteps = (unsigned)std::min<unsigned long long>(1024ull, c_last - g); A.pad = 0;
        }
        n_an_used = k;
    }
    if (hist_len) {                                            // the last hist_len phasors of the call feed the next call's history
        const size_t n_tail = std::min(n_items, hist_len);
        std::vector<float2> tail(n_tail);
        const unsigned long long c_first = ctr0 + (unsigned long long)(n_items - n_tail) + 1ull;   // item i sees Z(ctr0 + i + 1)
        float2 z = x.at(c_first);
        unsigned long long cc = c_first;
        for (size_t i = 0; i < n_tail; ++i) {
            tail[i] = z;
            ExactOsc::step1(z.x, z.y, cc, x.wre, x.wim);
        }
        if (n_tail >= hist_len) x.recent.swap(tail);
        else {
            x.recent.insert(x.recent.end(), tail.begin(), tail.end());
            if (x.recent.size() > hist_len) x.recent.erase(x.recent.begin(), x.recent.end() - (long)hist_len);
        }
    }
    x.consume_to(ctr0 + (unsigned long long)n_items);
    // Side stream: the tables do not depend on the data, so their upload and expansion overlap whatever the block's stream
    // is still running (the previous call's kernel).  Set p was last read by the kernel of two calls ago (used2[p]).
    const bool side = b->osc_side != 0;
    cudaStream_t os = side ? b->osc_stream : b->stream;
    if (side && x.used_valid[p]) CK(cudaStreamWaitEvent(os, x.used2[p], 0));
    CK(cudaMemcpyAsync(x.d_an2[p], x.h_an, n_an_used * sizeof(OscAnchor), cudaMemcpyHostToDevice, os));
    if (nh) CK(cudaMemcpyAsync(x.d_hist2[p], x.h_hist, nh * sizeof(float2), cudaMemcpyHostToDevice, os));
    CK(cudaEventRecord(x.staged2[p], os));
    CK(osc_expand_launch(x.d_an2[p], (int)n_an_used, x.d_fine2[p], ctr0 + 1ull, fine_len, os));
    if (side) {
        CK(cudaEventRecord(x.ready2[p], os));
        CK(cudaStreamWaitEvent(b->stream, x.ready2[p], 0));
    }
    b->osc_used_pending.push_back(x.used2[p]);           // recorded in either mode: the mode may change between calls
    x.used_valid[p] = true;
    b->launches += 1;
    np->exact = 1;
    np->xfine = x.d_fine2[p]; np->xfine_len = (int)fine_len;
    np->xhist = x.d_hist2[p]; np->xhist_len = (int)nh;
    np->xwre = x.wre; np->xwim = x.wim;
    return ORION_B200_OK;
}

// AgcRms / AgcRmsIq: chunked evaluation of the data-dependent envelope recurrence (agc_kernels.cu)
int launch_agc(orion_b200_block *b, const void *d_in, size_t n, void *d_out) {
    AgcArgs a;
    memset(&a, 0, sizeof(a));
    a.in = d_in; a.out = d_out; a.n = (long long)n; a.iq = b->agc == 2;
    a.attack_a = b->agc_attack_a; a.release_a = b->agc_release_a; a.target_rms = b->agc_target;
    a.min_gain = b->agc_min_gain; a.max_gain = b->agc_max_gain;
    // warm-up: amax^W <= 2^-26  ->  W = 18.03 / -ln(amax)
    const double amax = std::max((double)b->agc_attack_a, (double)b->agc_release_a);
    double w = (amax > 0.0 && amax < 1.0) ? std::ceil(18.03 / -std::log(amax)) : 1.0;
    if (amax >= 1.0) w = (double)n;                               // no contraction: one exact sequential chunk
    a.W = (long long)std::min<double>(std::max(w, 1.0), (double)n);
    a.L = std::max<long long>(std::max<long long>(a.W / 4, 64), ((long long)n + 65535) / 65536);
    if (const char *e = getenv("ORION_B200_AGC_CHUNK")) a.L = std::max(1, atoi(e));        // experiments / tests
    a.carry_in = b->d_carry[b->pp]; a.carry_out = b->d_carry[(b->pp + 1) % 3];
    if (b->agc == 3) {                                            // CwKeyedMod: rise / fall envelope + tone oscillator (replayed bit-exactly)
        a.iq = 2;
        a.gain = b->aux_gain;
        a.osc = b->post.param(b->k_post);
        const int st = prepare_exact(b, b->post, b->k_post, n, 0, &a.osc);
        if (st != ORION_B200_OK) return st;
    }
    cudaError_t e = agc_launch(a, b->stream);
    if (e != cudaSuccess) return fail(b, ORION_B200_ERR_CUDA, "agc kernel launch", e);
    b->launches += 1;
    b->pp = (b->pp + 1) % 3;
    if (b->agc == 3) b->k_post += n;
    return ORION_B200_OK;
}

int launch(orion_b200_block *b, const void *d_in, size_t n_in, void *d_out, size_t n_out, long long batch_in_stride = 0, long long batch_out_stride = 0);

// the blocks of aux_kernels.cu: symbol gain, slicers, FM and SSB modulators
int launch_aux(orion_b200_block *b, const void *d_in, size_t n_in, void *d_out, size_t n_out) {
    cudaError_t e = cudaSuccess;
    if (b->aux == AUX_GAIN) {
        e = gain_c32_launch(d_in, d_out, (long long)n_in, b->aux_gain, b->sm_count, b->stream);
        b->launches += 1;
    } else if (b->aux == AUX_SLICE) {
        SliceArgs a;
        memset(&a, 0, sizeof(a));
        a.in = (const float2 *)d_in; a.out = (unsigned char *)d_out; a.n_syms = (long long)n_in; a.bits = b->slice_bits;
        memcpy(a.th, b->slice_th, sizeof(a.th));
        e = slice_launch(a, b->sm_count, b->stream);
        b->launches += 1;
    } else if (b->aux == AUX_FMMOD) {
        const size_t ntiles = (n_in + 1023) / 1024;
        if (ntiles > b->fm_tiles_cap) {
            CK(cudaStreamSynchronize(b->stream));
            cudaFree(b->d_fm_tiles); b->d_fm_tiles = nullptr;
            CK(cudaMalloc(&b->d_fm_tiles, (ntiles + ntiles / 4 + 64) * sizeof(long long)));
            b->fm_tiles_cap = ntiles + ntiles / 4 + 64;
        }
        FmApplyArgs a;
        memset(&a, 0, sizeof(a));
        a.x = (const float *)d_in; a.out = (float2 *)d_out; a.n = (long long)n_in;
        a.kf = kTau * b->aux_p0 / b->fs_demod;                          // fm.rs:50
        a.gain = b->aux_gain;
        a.tile_off = b->d_fm_tiles;
        a.rf = b->post.param(b->k_post);
        const int st = prepare_exact(b, b->post, b->k_post, n_in, 0, &a.rf);
        if (st != ORION_B200_OK) return st;
        a.carry_in = b->d_carry[b->pp]; a.carry_out = b->d_carry[(b->pp + 1) % 3];
        e = fm_mod_launch(a, b->d_fm_tiles, b->stream);
        b->launches += 3;
        b->pp = (b->pp + 1) % 3;
        b->k_post += n_in;
    } else if (b->aux == AUX_SSBMOD) {
        if (n_in > b->ssb_cap) {
            CK(cudaStreamSynchronize(b->stream));
            for (float *&p : b->d_ssb) { cudaFree(p); p = nullptr; }
            const size_t cap = n_in + n_in / 4 + 64;
            for (float *&p : b->d_ssb) CK(cudaMalloc(&p, cap * sizeof(float)));
            b->ssb_cap = cap;
        }
        SsbSplitArgs sa;
        memset(&sa, 0, sizeof(sa));
        sa.x = (const float *)d_in; sa.xi = b->d_ssb[0]; sa.xq = b->d_ssb[1]; sa.n = (long long)n_in;
        sa.aud = b->pre.param(b->k_pre);
        int st = prepare_exact(b, b->pre, b->k_pre, n_in, 0, &sa.aud);
        if (st != ORION_B200_OK) return st;
        e = ssb_split_launch(sa, b->stream);
        if (e != cudaSuccess) return fail(b, ORION_B200_ERR_CUDA, "ssb split launch", e);
        for (int c = 0; c < 2; ++c) {                                    // lp_i, lp_q: the library's own LpCascade blocks
            b->child[c]->stream = b->stream;
            st = launch(b->child[c], b->d_ssb[c], n_in, b->d_ssb[2 + c], n_in, 0, 0);
            if (st != ORION_B200_OK) return fail(b, st, b->child[c]->err.c_str());
        }
        SsbCombineArgs ca;
        memset(&ca, 0, sizeof(ca));
        ca.yi = b->d_ssb[2]; ca.yq = b->d_ssb[3]; ca.out = (float2 *)d_out; ca.n = (long long)n_in; ca.side = b->aux_side;
        ca.rf = b->post.param(b->k_post);
        st = prepare_exact(b, b->post, b->k_post, n_in, 0, &ca.rf);
        if (st != ORION_B200_OK) return st;
        e = ssb_combine_launch(ca, b->stream);
        b->launches += 2;
        b->k_pre += n_in; b->k_post += n_in;
    }
    if (e != cudaSuccess) return fail(b, ORION_B200_ERR_CUDA, "aux kernel launch", e);
    (void)n_out;
    return ORION_B200_OK;
}

int launch_inner(orion_b200_block *b, const void *d_in, size_t n_in, void *d_out, size_t n_out,
                 long long batch_in_stride, long long batch_out_stride);
int launch(orion_b200_block *b, const void *d_in, size_t n_in, void *d_out, size_t n_out,
           long long batch_in_stride, long long batch_out_stride) {
    const int st = launch_inner(b, d_in, n_in, d_out, n_out, batch_in_stride, batch_out_stride);
    // everything that reads this call's oscillator tables is enqueued: from here on the block's stream their set may be rewritten
    for (cudaEvent_t ev : b->osc_used_pending) cudaEventRecord(ev, b->stream);
    b->osc_used_pending.clear();
    return st;
}
int launch_inner(orion_b200_block *b, const void *d_in, size_t n_in, void *d_out, size_t n_out,
                 long long batch_in_stride, long long batch_out_stride) {
    if (b->plan_dirty) { int st = finalize_plan(b); if (st) return st; }
    if (n_in == 0) return ORION_B200_OK;
    CK(cudaSetDevice(b->device));
    if (b->agc) return launch_agc(b, d_in, n_in, d_out);
    if (b->aux) return launch_aux(b, d_in, n_in, d_out, n_out);
    const int npt = npt_of(b);
    const long long tile_items = (long long)kThreads * npt;
    long long ntiles = ((long long)n_out + tile_items - 1) / tile_items;
    if (ntiles < 1) ntiles = 1;
    if (ntiles > 0x7fffffffLL) return fail(b, ORION_B200_ERR_UNSUPPORTED, "call too long");
    const int nsec = (int)b->secs.size();
    if (nsec > 0 && (size_t)ntiles > b->links_cap) {
        CK(cudaStreamSynchronize(b->stream));
        if (b->d_links) cudaFree(b->d_links);
        size_t cap = std::max<size_t>((size_t)ntiles, 4096);
        cap += cap / 4;
        CK(cudaMalloc(&b->d_links, 2 * cap * b->nbatch * kMaxGroups * sizeof(TileLink)));      // two halves, alternating between calls
        CK(dev_memset(b, b->d_links, 0, 2 * cap * b->nbatch * kMaxGroups * sizeof(TileLink)));     // stream-ordered before the launch below
        b->links_cap = cap;
        b->epoch = 0;
    }
    b->epoch += 1;
    if (b->epoch >= (1u << 30)) {                      // epoch wrap: clear the links once
        CK(cudaStreamSynchronize(b->stream));
        CK(cudaMemsetAsync(b->d_links, 0, 2 * b->links_cap * b->nbatch * kMaxGroups * sizeof(TileLink), b->stream));
        b->epoch = 1;
    }

    ChainArgs *ap = new (std::nothrow) ChainArgs();    // ~11 KB: keep it off the stack
    if (!ap) return fail(b, ORION_B200_ERR_ALLOC, "ChainArgs");
    ChainArgs &a = *ap;
    memset(&a, 0, sizeof(a));
    a.in = d_in; a.out = d_out;
    a.n_in = (long long)n_in; a.n_out = (long long)n_out;
    a.hist_in = b->d_hist[b->pp]; a.hist_out = b->d_hist[(b->pp + 1) % 3];
    a.H = (b->fir != FIR_NONE) ? b->plan.H : 0;
    a.mix = b->mix;
    a.pre = b->pre.param(b->k_pre);
    a.fir = b->fir; a.M = (int)b->M; a.Lg = (int)b->plan.g.size(); a.g = b->d_g;
    a.Mb = b->plan.Mb; a.O = b->plan.O; a.P_pad = b->plan.P_pad; a.HR = b->plan.HR;
    a.row_samples = b->plan.row_samples; a.row_pitch = b->plan.row_pitch;
    a.row_shift = -1;
    for (int sh = 0; sh < 20; ++sh) if ((1 << sh) == b->plan.row_samples) a.row_shift = sh;
    a.nstages = b->opt_serial ? std::min(b->plan.nstages, 1) : b->plan.nstages;
    a.demod = b->demod; a.translate = b->translate; a.k = b->k; a.k1 = b->k1; a.k2 = b->k2; a.k3 = b->k3;
    a.post = b->post.param(b->k_post);
    bool exact_used = false;
    if (want_exact_pre(b)) {
        const int st = prepare_exact(b, b->pre, b->k_pre, n_in, (b->fir != FIR_NONE) ? (size_t)b->plan.H : 0, &a.pre);
        if (st != ORION_B200_OK) { delete ap; return st; }
        exact_used = true;
    }
    if (want_exact_post(b)) {
        const int st = prepare_exact(b, b->post, b->k_post, n_out, 0, &a.post);
        if (st != ORION_B200_OK) { delete ap; return st; }
        exact_used = true;
    }
    a.nsec = nsec;
    for (int s = 0; s < nsec; ++s) a.sec[s] = b->secs[s];
    a.ngroups = (int)b->groups.size();
    a.pipe_park_slots = b->pipe_park_slots; a.pipe_u_slots = b->pipe_u_slots;
    a.nstage = b->nstage;
    for (int i = 0; i < 8; ++i) a.stage_group[i] = b->stage_group[i];
    a.split = b->opt_serial ? 1 : b->split;
    for (int g = 0; g < a.ngroups; ++g) a.grp[g] = b->groups[g];
    a.gtabs = b->d_gtabs;
    a.carry_in = b->d_carry[b->pp]; a.carry_out = b->d_carry[(b->pp + 1) % 3];
    a.links = b->d_links ? b->d_links + (size_t)(b->epoch & 1u) * b->links_cap * b->nbatch * kMaxGroups : nullptr;   // consecutive calls may overlap
    a.batch = b->nbatch;
    a.batch_chan = b->d_batch_chan;
    a.batch_in_stride = batch_in_stride; a.batch_out_stride = batch_out_stride;
    a.batch_links_stride = (long long)b->links_cap * kMaxGroups;
    a.epoch = b->epoch; a.ntiles = (int)ntiles; a.serial = b->opt_serial; a.err_flag = b->d_err_ext ? b->d_err_ext : b->d_err;
    a.trace = b->trace;
    {   // tile t's look-back reaches the carried state iff t < depth of a group; tile 0 also reads history and `prev`
        int guard = 1;
        for (size_t g = 0; g < b->groups.size(); ++g) guard = std::max(guard, std::min(b->group_depth[g], 32) + 1);
        a.pdl_guard = guard;
    }
    a.handoff = b->d_handoff;
    a.hist_target = b->calls_since_reset;                 // every earlier call has written its history ...
    a.carry_target = 2u * b->calls_since_reset;           // ... and both of its hand-over signals
    a.depth_slot = 2 + (int)(b->calls_since_reset & 1u);  // the call before the previous one has the same parity:
    a.depth_target = b->ctas_par[b->calls_since_reset & 1u];   // ... every CTA of it (and of its same-parity predecessors) has run to its end
    if (!b->plan.taps2.empty()) memcpy(a.taps2, b->plan.taps2.data(), b->plan.taps2.size() * sizeof(float2));
    a.ntaps2 = (int)b->plan.taps2.size();

    CUtensorMap tmap;
    memset(&tmap, 0, sizeof(tmap));
    a.tile_int_lo = 1; a.tile_int_hi = 0;              // no interior tiles unless a tensor map is set up below
    a.ns_magic = (unsigned)(((1ull << 32) + (unsigned long long)std::max(a.nstages, 1) - 1) / (unsigned long long)std::max(a.nstages, 1));
    if (b->plan.front == FRONT_STAGED && b->opt_use_tma && ((reinterpret_cast<uintptr_t>(d_in) & 15u) == 0)) {
        encode_tiled_t enc = get_encode_tiled();
        const long long rs = b->plan.row_samples;
        const long long off = (long long)b->plan.O - b->plan.Mb + 2;        // start sample of global row 0
        long long row0 = 0;
        if (off < 0) row0 = (-off + rs - 1) / rs;
        const long long start = rs * row0 + off;
        const long long nrows = ((long long)n_in - start) / rs;
        if (enc && nrows >= b->plan.rows) {
            // 8-byte elements (one complex sample each); the box is 2 elements wider than the tensor
            // row, so the TMA zero-fills the 16-byte pad that keeps the rows bank-conflict free
            const cuuint64_t gdim[2] = { (cuuint64_t)rs, (cuuint64_t)nrows };
            const cuuint64_t gstr[1] = { (cuuint64_t)(rs * 8) };
            const cuuint32_t box[2] = { (cuuint32_t)(b->plan.row_pitch / 8), (cuuint32_t)b->plan.rows };
            const cuuint32_t estr[2] = { 1, 1 };
            void *base = (void *)(reinterpret_cast<const char *>(d_in) + start * 8);
            CUresult r = enc(&tmap, CU_TENSOR_MAP_DATA_TYPE_FLOAT64, 2, base, gdim, gstr, box, estr,
                             CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE,
                             CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
            if (r == CUDA_SUCCESS) {
                a.use_tma = 1; a.tma_row0 = row0; a.tma_rows = nrows;
                const long long hr = b->plan.HR;
                a.tile_int_lo = (row0 + hr + kThreads - 1) / kThreads;                 // 32 t - HR >= row0
                a.tile_int_hi = (row0 + nrows - kThreads) / kThreads;                  // 32 t + 32 <= row0 + nrows
                if (row0 + nrows - kThreads < 0) a.tile_int_hi = -1;
            }
            // L2 look-ahead of the stage ring, in fills (tiles of this CTA): ~2 ring depths; 0 = off
            a.l2_prefetch = 0;                 // measured on B200: any distance costs 3-10 % (profiles/r02_experiments.txt); kept for experiments
            if (const char *e2 = getenv("ORION_B200_L2_PREFETCH")) a.l2_prefetch = std::max(0, atoi(e2));
        }
    }

    if (b->plan.front == FRONT_DIRECT) {                   // L2 look-ahead of the rate-1 blocks, in tickets of a CTA (two per warp)
        a.l2_prefetch = 0;
        if (const char *e2 = getenv("ORION_B200_L2_PREFETCH")) a.l2_prefetch = std::max(0, atoi(e2));
    }
    int grid = 1;
    if (!b->opt_serial) {
        const long long resident = (long long)b->sm_count * b->ctas_per_sm;
        const long long want = (ntiles + b->plan.warps - 1) / b->plan.warps;      // one tile per warp at least
        grid = (int)std::max<long long>(1, std::min<long long>(want, resident));
        if (const char *e = getenv("ORION_B200_GRID")) grid = std::max(1, std::min(grid, atoi(e)));   // experiments
        if (b->nbatch > 1) {
            // One CTA per member when the members fill the machine (its tiles then only ever wait for tiles of the same
            // CTA).  A small bank (the 128-channel shard of an 8-GPU run: 64 members per demodulator kind) would leave
            // more than half of the SMs idle and every warp with eight latency-bound iterations; it gets two CTAs per
            // member -- CTAs of one member are neighbours in launch order, so both are resident together.
            int g = 1;
            while (g < 4 && (long long)b->nbatch * (2 * g) <= resident && ntiles >= (long long)(4 * g) * b->plan.warps) g *= 2;
            if (const char *e = getenv("ORION_B200_BATCH_GRID")) g = std::max(1, std::min(8, atoi(e)));
            grid = g;
        }
    }
    // Overlap with the previous launch on the stream (programmatic dependent launch): only long calls (the link
    // records and the output of call N are far from what call N+1 touches first), only section groups whose
    // look-back cannot reach the carried state past the guarded tiles, and only on the block's own stream unless
    // the caller opted in (ORION_B200_OPT_OVERLAP_LAUNCHES) -- on an attached stream the predecessor may be a
    // foreign kernel that is still producing this call's input.
    bool overlap = !b->opt_serial && ntiles >= 1024 && b->nbatch == 1 && (b->stream == b->own_stream || b->opt_overlap);
    for (const GroupParam &gp : b->groups) overlap = overlap && gp.agg_only;
    if (exact_used) overlap = false;                       // the expansion kernel just enqueued must have finished
    if (getenv("ORION_B200_NO_OVERLAP")) overlap = false;
    cudaError_t e = chain_kernel_launch(b->kernel, a, tmap, grid, b->plan.warps, b->plan.dyn_smem, b->stream, overlap ? 1 : 0);
    delete ap;
    if (e != cudaSuccess) return fail(b, ORION_B200_ERR_CUDA, "chain kernel launch", e);
    b->launches += 1;
    b->ctas_par[b->calls_since_reset & 1u] += (unsigned)grid * (unsigned)b->nbatch;
    b->calls_since_reset += 1;
    if (b->calls_since_reset >= (1u << 30) || b->ctas_par[0] >= (1u << 30) || b->ctas_par[1] >= (1u << 30)) {   // counter wrap: drain, start over
        CK(cudaStreamSynchronize(b->stream));
        CK(dev_memset(b, b->d_handoff, 0, 4 * sizeof(unsigned int)));
        b->calls_since_reset = 0;
        b->ctas_par[0] = b->ctas_par[1] = 0;
    }
    b->pp = (b->pp + 1) % 3;
    b->k_pre += n_in;
    b->k_post += n_out;
    return ORION_B200_OK;
}

int check_device_error(orion_b200_block *b) {
    CK(cudaMemcpyAsync(b->h_err, b->d_err, sizeof(int), cudaMemcpyDeviceToHost, b->stream));
    CK(cudaStreamSynchronize(b->stream));
    if (*b->h_err != 0) {
        char msg[96];
        snprintf(msg, sizeof(msg), "device watchdog tripped (code %d): inter-tile link or TMA wait timed out", *b->h_err);
        *b->h_err = 0;
        dev_memset(b, b->d_err, 0, sizeof(int));
        cudaStreamSynchronize(b->stream);
        return fail(b, ORION_B200_ERR_INTERNAL, msg);
    }
    return ORION_B200_OK;
}

}  // namespace

// ================================================================================================
// C ABI
// ================================================================================================
extern "C" {

int orion_b200_abi_version(void) { return ORION_B200_ABI_VERSION; }
const char *orion_b200_last_create_error(void) { return t_create_error.c_str(); }
const char *orion_b200_build_info(void) {
    return "orion_b200 sm_100a (compute_100a) -fmad=false; kernels: chain_kernel<front,R,U>; no CPU fallback";
}
int orion_b200_device_count(void) {
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess) { cudaGetLastError(); return 0; }
    return n;
}
int orion_b200_set_device(int ordinal) {
    int n = orion_b200_device_count();
    if (ordinal < 0 || ordinal >= n) return ORION_B200_ERR_NO_DEVICE;
    t_device = ordinal;
    return ORION_B200_OK;
}
const char *orion_b200_status_string(int status) {
    switch (status) {
        case ORION_B200_OK: return "ok";
        case ORION_B200_ERR_INVALID: return "invalid argument";
        case ORION_B200_ERR_NO_DEVICE: return "no usable CUDA device (there is no CPU fallback)";
        case ORION_B200_ERR_CUDA: return "CUDA error";
        case ORION_B200_ERR_ALLOC: return "allocation failed";
        case ORION_B200_ERR_UNSUPPORTED: return "unsupported configuration";
        case ORION_B200_ERR_INTERNAL: return "internal error";
    }
    return "unknown status";
}
int orion_b200_host_alloc(void **ptr, size_t bytes) {
    if (!ptr) return ORION_B200_ERR_INVALID;
    *ptr = nullptr;
    if (orion_b200_device_count() <= 0) return ORION_B200_ERR_NO_DEVICE;
    if (cudaMallocHost(ptr, bytes ? bytes : 1) != cudaSuccess) { cudaGetLastError(); return ORION_B200_ERR_ALLOC; }
    return ORION_B200_OK;
}
void orion_b200_host_free(void *ptr) { if (ptr) cudaFreeHost(ptr); }

// ---- design helpers [host-only] ------------------------------------------------------------------
size_t orion_b200_fir_lowpass_design(float fs, float pass_hz, float trans_hz, float *taps, size_t cap) {
    const size_t n = fir_lowpass_ntaps(fs, pass_hz, trans_hz);
    if (taps && cap >= n) fir_lowpass_design(fs, pass_hz, taps, n);
    return n;
}
size_t orion_b200_kaiser_lowpass_taps(size_t num_taps, float cutoff_norm, float stopband_db, float *taps, size_t cap) {
    const size_t m = kaiser_len(num_taps);
    if (taps && cap >= m) kaiser_taps(m, cutoff_norm, stopband_db, taps);
    return m;
}
float orion_b200_kaiser_transition_norm(size_t num_taps, float stopband_db) {   // fir.rs:147-150
    const float m = (float)kaiser_len(num_taps);
    return (maxf_rs(stopband_db, 21.0f) - 8.0f) / (14.36f * m);
}
size_t orion_b200_kaiser_num_taps(float transition_norm, float stopband_db) {   // fir.rs:154-157
    const float m = ceilf((maxf_rs(stopband_db, 21.0f) - 8.0f) / (14.36f * maxf_rs(transition_norm, 1e-4f)));
    return ((size_t)maxf_rs(m, 3.0f)) | 1;
}
void orion_b200_lp_biquad_design(float fs, float fc, float coeffs[5]) { lp_biquad_design(fs, fc, coeffs); }
float orion_b200_dc_pole(float fs, float cut_hz) { return dc_pole(fs, cut_hz); }
float orion_b200_cw_alpha(float fs, float env_bw_hz) { return cw_alpha(fs, env_bw_hz); }

// ---- constructors ------------------------------------------------------------------------------
#define NEW_BLOCK() orion_b200_block *b = nullptr; { int _s = new_block(out, &b); if (_s) return _s; }

int orion_b200_fir_decimator_create_taps(const float *taps, size_t ntaps, size_t m, orion_b200_block **out) {
    if (!taps || ntaps == 0) return ORION_B200_ERR_INVALID;
    NEW_BLOCK();
    b->fir = FIR_DECIM;
    b->taps.assign(taps, taps + ntaps);
    b->M = m < 1 ? 1 : m;                                           // decim.rs:30
    return finish_create(b, out);
}
int orion_b200_fir_decimator_create(float fs, size_t m, float cutoff_hz, float trans_hz, orion_b200_block **out) {
    const size_t n = fir_lowpass_ntaps(fs, cutoff_hz, trans_hz);
    std::vector<float> t(n);
    fir_lowpass_design(fs, cutoff_hz, t.data(), n);
    return orion_b200_fir_decimator_create_taps(t.data(), n, m, out);
}
int orion_b200_fir_lowpass_iq_create_taps(const float *taps, size_t ntaps, orion_b200_block **out) {
    NEW_BLOCK();
    b->fir = FIR_IQ;
    if (!taps || ntaps == 0) b->taps.assign(1, 1.0f);               // fir.rs:193-196: empty -> identity
    else b->taps.assign(taps, taps + ntaps);
    b->M = 1;
    return finish_create(b, out);
}
// HalfCosineMf::new, fir.rs:325-346: unit-energy half-cosine pulse
size_t orion_b200_half_cosine_mf_taps(size_t sps, float *taps, size_t cap) {
    const size_t n = sps < 1 ? 1 : sps;
    if (!taps || cap < n) return n;
    if (sps <= 1) taps[0] = 1.0f;
    else {
        const float denom = (float)(sps - 1);
        for (size_t i = 0; i < sps; ++i) taps[i] = 0.5f - 0.5f * cosf(kPi * (float)i / denom);
    }
    float energy = 0.0f;
    for (size_t i = 0; i < n; ++i) energy += taps[i] * taps[i];
    const float scale = (energy > 0.0f) ? 1.0f / sqrtf(energy) : 1.0f;
    for (size_t i = 0; i < n; ++i) taps[i] = taps[i] * scale;
    return n;
}
int orion_b200_half_cosine_mf_create(size_t sps, orion_b200_block **out) {
    const size_t n = orion_b200_half_cosine_mf_taps(sps, nullptr, 0);
    std::vector<float> t(n);
    orion_b200_half_cosine_mf_taps(sps, t.data(), n);
    NEW_BLOCK();
    b->fir = FIR_IQ_UNFUSED;
    b->taps = t;
    b->M = 1;
    return finish_create(b, out);
}
int orion_b200_fir_lowpass_iq_create(size_t num_taps, float cutoff_norm, float stopband_db, orion_b200_block **out) {
    const size_t m = kaiser_len(num_taps);
    std::vector<float> t(m);
    kaiser_taps(m, cutoff_norm, stopband_db, t.data());
    return orion_b200_fir_lowpass_iq_create_taps(t.data(), m, out);
}

int orion_b200_rotator_create(float freq_hz, float fs, orion_b200_block **out) {
    NEW_BLOCK();
    b->mix = MIX_ROTATE;
    b->pre.set(freq_hz, fs, 0);
    return finish_create(b, out);
}
int orion_b200_rotator_usb_create(float freq_hz, float fs, orion_b200_block **out) {
    NEW_BLOCK();
    b->demod = DEMOD_USB;
    b->post.set(freq_hz, fs, 0);
    return finish_create(b, out);
}
int orion_b200_nco_mixer_create(float freq_hz, float fs, orion_b200_block **out) {
    NEW_BLOCK();
    b->mix = MIX_NCO;
    b->pre.set(freq_hz, fs, 0);
    return finish_create(b, out);
}
int orion_b200_oscillator_set_freq(orion_b200_block *b, float freq_hz, float fs) {
    if (!b) return ORION_B200_ERR_INVALID;
    if (b->pre.on) b->pre.set(freq_hz, fs, b->k_pre);
    else if (b->post.on) b->post.set(freq_hz, fs, b->k_post);
    else return fail(b, ORION_B200_ERR_INVALID, "block has no oscillator");
    return ORION_B200_OK;
}
int orion_b200_oscillator_reset_phase(orion_b200_block *b) {
    if (!b) return ORION_B200_ERR_INVALID;
    // rotator.rs:28-31: z = 1, ctr = 0 -> the call counter restarts as well
    if (b->pre.on) { b->pre.reset_phase(); b->k_pre = 0; }
    if (b->post.on) { b->post.reset_phase(); b->k_post = 0; }
    return ORION_B200_OK;
}

int orion_b200_iir_cascade_create(const float *sos, size_t nsections, orion_b200_block **out) {
    if (!sos || nsections == 0 || nsections > (size_t)kMaxSections) return ORION_B200_ERR_INVALID;
    NEW_BLOCK();
    b->demod = DEMOD_F32;
    for (size_t s = 0; s < nsections; ++s) b->secs.push_back(sec_biquad(sos + 5 * s));
    return finish_create(b, out);
}
int orion_b200_biquad_create(float b0, float b1, float b2, float a1, float a2, orion_b200_block **out) {
    const float c[5] = { b0, b1, b2, a1, a2 };
    return orion_b200_iir_cascade_create(c, 1, out);
}
int orion_b200_lp_cascade_create(float fs, float fc, orion_b200_block **out) {
    NEW_BLOCK();
    b->demod = DEMOD_F32;
    add_lr4(b, fs, fc);
    return finish_create(b, out);
}
int orion_b200_lp_dc_cascade_create(float fs, float lp_fc, float dc_cut_hz, int map_sqrt, orion_b200_block **out) {
    NEW_BLOCK();
    b->demod = DEMOD_F32;
    add_lr4(b, fs, lp_fc);
    if (map_sqrt) b->secs.back().post_op = OP_SQRT;                 // iir.rs:170-186 process_mapped(x, sqrt)
    b->secs.push_back(sec_dc(dc_pole(fs, dc_cut_hz)));
    return finish_create(b, out);
}
int orion_b200_dc_blocker_create(float fs, float cut_hz, orion_b200_block **out) {
    NEW_BLOCK();
    b->demod = DEMOD_F32;
    b->secs.push_back(sec_dc(dc_pole(fs, cut_hz)));
    return finish_create(b, out);
}

static void demod_setup(orion_b200_block *b, int demod, float fs, float p0, float p1, float audio_bw_hz) {
    b->demod = demod;
    b->fs_demod = fs;
    switch (demod) {
        case DEMOD_FM:                                              // fm.rs:22-32
            b->k = 1.0f / maxf_rs(p0, 1.0f);
            add_lr4(b, fs, audio_bw_hz * 0.9f);
            break;
        case DEMOD_PM:                                              // pm.rs:22-32
            b->k = p0;
            add_lr4(b, fs, audio_bw_hz * 0.9f);
            break;
        case DEMOD_AM:                                              // am.rs:24-30 (PowerSqrt)
            add_lr4(b, fs, audio_bw_hz * 0.9f);
            b->secs.back().post_op = OP_SQRT;
            b->secs.push_back(sec_dc(dc_pole(fs, 2.0f)));
            break;
        case DEMOD_AM_ABS:                                          // am.rs:33-36
            b->k1 = p0; b->k2 = p1;
            add_lr4(b, fs, audio_bw_hz * 0.9f);
            b->secs.push_back(sec_dc(dc_pole(fs, 2.0f)));
            break;
        case DEMOD_SSB:                                             // ssb.rs:15-20
            b->post.set(p0, fs, 0);
            add_lr4(b, fs, audio_bw_hz * 0.9f);
            b->secs.push_back(sec_dc(dc_pole(fs, 2.0f)));
            break;
        case DEMOD_CW: {                                            // cw.rs:15-24
            SecParam p; memset(&p, 0, sizeof(p));
            const float a = cw_alpha(fs, p0);
            p.type = SEC_ONEPOLE; p.c[0] = a; p.c[1] = 1.0f - a;
            p.post_op = OP_SCALE; p.post_scale = (p1 == 0.0f) ? 1.0f : p1;
            b->cw_gain_sec = (int)b->secs.size();
            b->secs.push_back(p);
            break;
        }
        case DEMOD_USB:
            b->post.set(p0, fs, 0);
            break;
        default: break;
    }
}

int orion_b200_fm_demod_create(float fs, float dev_hz, float audio_bw_hz, orion_b200_block **out) {
    NEW_BLOCK();
    demod_setup(b, DEMOD_FM, fs, dev_hz, 0.f, audio_bw_hz);
    return finish_create(b, out);
}
int orion_b200_fm_demod_with_translate(orion_b200_block *b, float freq_hz) {
    if (!b || b->demod != DEMOD_FM) return ORION_B200_ERR_INVALID;
    b->translate = 1;
    b->post = Osc();
    b->post.set(freq_hz, b->fs_demod, 0);                           // fm.rs:35: a fresh Rotator
    b->k_post = 0;
    return ORION_B200_OK;
}
int orion_b200_pm_demod_create(float fs, float k, float audio_bw_hz, orion_b200_block **out) {
    NEW_BLOCK();
    demod_setup(b, DEMOD_PM, fs, k, 0.f, audio_bw_hz);
    return finish_create(b, out);
}
int orion_b200_am_demod_create(float fs, float audio_bw_hz, orion_b200_block **out) {
    NEW_BLOCK();
    demod_setup(b, DEMOD_AM, fs, 0.f, 0.f, audio_bw_hz);
    return finish_create(b, out);
}
int orion_b200_am_demod_with_abs_approx(orion_b200_block *b, float k1, float k2) {
    if (!b || (b->demod != DEMOD_AM && b->demod != DEMOD_AM_ABS)) return ORION_B200_ERR_INVALID;
    b->demod = DEMOD_AM_ABS;
    b->k1 = k1; b->k2 = k2;
    for (auto &s : b->secs) if (s.post_op == OP_SQRT) s.post_op = OP_NONE;
    b->plan_dirty = true;            // the demodulator kind and the section groups changed: reselect the kernel instance
    return ORION_B200_OK;
}
int orion_b200_ssb_demod_create(float fs, float bfo_hz, float audio_bw_hz, orion_b200_block **out) {
    NEW_BLOCK();
    demod_setup(b, DEMOD_SSB, fs, bfo_hz, 0.f, audio_bw_hz);
    return finish_create(b, out);
}
int orion_b200_cw_demod_create(float sample_rate, float tone_hz, float env_bw_hz, orion_b200_block **out) {
    (void)tone_hz;                                                  // cw.rs:16: unused by the reference too
    NEW_BLOCK();
    demod_setup(b, DEMOD_CW, sample_rate, env_bw_hz, 1.0f, 0.f);
    return finish_create(b, out);
}
int orion_b200_cw_demod_set_gain(orion_b200_block *b, float gain) {
    if (!b || b->demod != DEMOD_CW || b->cw_gain_sec < 0) return ORION_B200_ERR_INVALID;
    b->secs[b->cw_gain_sec].post_scale = gain;
    return ORION_B200_OK;
}

int orion_b200_chain_create(const orion_b200_chain_spec *spec, orion_b200_block **out) {
    if (!spec || spec->struct_size != sizeof(orion_b200_chain_spec)) return ORION_B200_ERR_INVALID;
    if (spec->fir != FIR_NONE && (!spec->taps || spec->ntaps == 0)) return ORION_B200_ERR_INVALID;
    if (spec->demod < 0 || spec->demod > DEMOD_USB) return ORION_B200_ERR_INVALID;
    NEW_BLOCK();
    b->mix = spec->mix;
    if (spec->mix != MIX_NONE) b->pre.set(spec->mix_freq_hz, spec->mix_fs, 0);
    b->fir = spec->fir;
    if (spec->fir != FIR_NONE) {
        b->taps.assign(spec->taps, spec->taps + spec->ntaps);
        b->M = spec->decim < 1 ? 1 : spec->decim;
    }
    demod_setup(b, spec->demod, spec->fs_demod, spec->p0, spec->p1, spec->audio_bw_hz);
    if (spec->demod == DEMOD_FM && spec->translate) {
        b->translate = 1;
        b->post.set(spec->translate_hz, spec->fs_demod, 0);
    }
    if (spec->n_post) {
        if (!spec->post_sos || spec->demod == DEMOD_NONE) { delete b; return ORION_B200_ERR_INVALID; }
        for (size_t s = 0; s < spec->n_post; ++s) b->secs.push_back(sec_biquad(spec->post_sos + 5 * s));
    }
    return finish_create(b, out);
}

// ---- modulators (next-row scope): f32 audio -> C32 IQ ------------------------------------
int orion_b200_am_mod_create(float fs, float rf_hz, float carrier_level, float modulation_index, orion_b200_block **out) {
    NEW_BLOCK();                                                     // AmDsbMod::new, modulate/am.rs:21-30
    b->demod = MOD_AM;
    b->k = 0.f; b->k1 = carrier_level; b->k2 = modulation_index; b->k3 = 1.0f;
    b->post.set(rf_hz, fs, 0);
    return finish_create(b, out);
}
int orion_b200_am_mod_set_clamp(orion_b200_block *b, int on) {      // modulate/am.rs:34-36
    if (!b || b->demod != MOD_AM) return b ? fail(b, ORION_B200_ERR_INVALID, "not an AM modulator") : ORION_B200_ERR_INVALID;
    b->k = on ? 1.0f : 0.0f;
    return ORION_B200_OK;
}
int orion_b200_pm_mod_create(float fs, float kp_rad_per_unit, float rf_hz, orion_b200_block **out) {
    NEW_BLOCK();                                                     // PmDirectPhaseMod::new, modulate/pm.rs:17-23
    b->demod = MOD_PM;
    b->k1 = kp_rad_per_unit; b->k2 = 1.0f;
    b->post.set(rf_hz, fs, 0);
    return finish_create(b, out);
}
int orion_b200_mod_set_gain(orion_b200_block *b, float gain) {      // set_gain, modulate/am.rs:31-33, pm.rs:24-26
    if (!b) return ORION_B200_ERR_INVALID;
    if (b->aux == AUX_FMMOD || b->agc == 3) b->aux_gain = gain;      // modulate/fm.rs:35-37, cw.rs:35-37
    else if (b->demod == MOD_AM) b->k3 = gain;
    else if (b->demod == MOD_PM) b->k2 = gain;
    else return fail(b, ORION_B200_ERR_INVALID, "not a modulator");
    return ORION_B200_OK;
}

// ---- AGC (next-row scope) ------------------------------------------------------------------
static int agc_create(int kind, float fs, float attack_ms, float release_ms, float target_rms, orion_b200_block **out) {
    NEW_BLOCK();                                                     // AgcRms::new / AgcRmsIq::new, agc.rs:20-31, :93-104
    b->agc = kind;
    b->demod = kind == 1 ? DEMOD_F32 : DEMOD_NONE;                   // item types: f32 -> f32 | C32 -> C32
    b->agc_attack_a = expf(-1.0f / (fs * (maxf_rs(attack_ms, 1e-3f) / 1000.0f)));
    b->agc_release_a = expf(-1.0f / (fs * (maxf_rs(release_ms, 1e-3f) / 1000.0f)));
    b->agc_target = maxf_rs(target_rms, 1e-6f);
    return finish_create(b, out);
}
int orion_b200_agc_rms_create(float fs, float attack_ms, float release_ms, float target_rms, orion_b200_block **out) {
    return agc_create(1, fs, attack_ms, release_ms, target_rms, out);
}
int orion_b200_agc_rms_iq_create(float fs, float attack_ms, float release_ms, float target_rms, orion_b200_block **out) {
    return agc_create(2, fs, attack_ms, release_ms, target_rms, out);
}
float orion_b200_agc_env(orion_b200_block *b) {                     // the tracked power (agc.rs `env`), for the parity harness
    if (!b || !b->agc) return 0.0f;
    cudaSetDevice(b->device);
    cudaStreamSynchronize(b->stream);
    CarryState cs;
    if (cudaMemcpy(&cs, b->d_carry[b->pp], sizeof(cs), cudaMemcpyDeviceToHost) != cudaSuccess) return 0.0f;
    return cs.pad.x;
}

// ---- modulators, continued: FM (phase prefix sum), CW (chunked envelope), SSB (phasing) ------
int orion_b200_fm_mod_create(float sample_rate, float deviation_hz, float rf_hz, orion_b200_block **out) {
    NEW_BLOCK();                                                     // FmPhaseAccumMod::new, modulate/fm.rs:22-31
    b->aux = AUX_FMMOD;
    b->in_item = ORION_B200_ITEM_F32; b->out_item = ORION_B200_ITEM_C32;
    b->demod = DEMOD_F32;
    b->fs_demod = sample_rate;
    b->aux_p0 = deviation_hz;
    b->post.set(rf_hz, sample_rate, 0);
    return finish_create(b, out);
}
int orion_b200_fm_mod_set_deviation(orion_b200_block *b, float deviation_hz) {    // modulate/fm.rs:32-34
    if (!b || b->aux != AUX_FMMOD) return b ? fail(b, ORION_B200_ERR_INVALID, "not an FM modulator") : ORION_B200_ERR_INVALID;
    b->aux_p0 = deviation_hz;
    return ORION_B200_OK;
}
int orion_b200_cw_mod_create(float sample_rate, float tone_hz, float rise_ms, float fall_ms, orion_b200_block **out) {
    NEW_BLOCK();                                                     // CwKeyedMod::new, modulate/cw.rs:21-34
    b->agc = 3;
    b->in_item = ORION_B200_ITEM_F32; b->out_item = ORION_B200_ITEM_C32;
    b->demod = DEMOD_F32;
    const float tau_r = (maxf_rs(rise_ms, 0.1f) * 1e-3f) * sample_rate;
    const float tau_f = (maxf_rs(fall_ms, 0.1f) * 1e-3f) * sample_rate;
    b->agc_attack_a = expf(-1.0f / tau_r);
    b->agc_release_a = expf(-1.0f / tau_f);
    b->post.set(tone_hz, sample_rate, 0);
    return finish_create(b, out);
}
int orion_b200_ssb_mod_create(float fs, float audio_bw_hz, float audio_if_hz, float rf_hz, int usb, orion_b200_block **out) {
    NEW_BLOCK();                                                     // SsbPhasingMod::new, modulate/ssb.rs:23-35
    b->aux = AUX_SSBMOD;
    b->in_item = ORION_B200_ITEM_F32; b->out_item = ORION_B200_ITEM_C32;
    b->demod = DEMOD_F32;
    b->aux_side = usb ? 1.0f : -1.0f;
    b->pre.set(audio_if_hz, fs, 0);
    b->post.set(rf_hz, fs, 0);
    for (int c = 0; c < 2; ++c) {
        const int st = orion_b200_lp_cascade_create(fs, audio_bw_hz * 0.9f, &b->child[c]);     // ssb.rs:25-30
        if (st != ORION_B200_OK) { orion_b200_block_destroy(b->child[0]); delete b; return st; }
    }
    return finish_create(b, out);
}

// ---- soft-symbol gain blocks and hard-decision slicers ------------------------------------
int orion_b200_symbol_gain_create(float gain, orion_b200_block **out) {       // BpskDemod / QpskDemod / QamDemod::new
    NEW_BLOCK();
    b->aux = AUX_GAIN;
    b->in_item = ORION_B200_ITEM_C32; b->out_item = ORION_B200_ITEM_C32;
    b->aux_gain = gain;
    return finish_create(b, out);
}
int orion_b200_symbol_gain_set(orion_b200_block *b, float gain) {             // set_gain, bpsk.rs:22-24
    if (!b || b->aux != AUX_GAIN) return b ? fail(b, ORION_B200_ERR_INVALID, "not a symbol gain block") : ORION_B200_ERR_INVALID;
    b->aux_gain = gain;
    return ORION_B200_OK;
}
int orion_b200_decider_create(int bits_per_symbol, orion_b200_block **out) {
    if (!(bits_per_symbol == 1 || bits_per_symbol == 2 || bits_per_symbol == 4 || bits_per_symbol == 6 || bits_per_symbol == 8)) {
        if (out) *out = nullptr;
        return ORION_B200_ERR_INVALID;                               // qam.rs:13-18 check_bits
    }
    NEW_BLOCK();
    b->aux = AUX_SLICE;
    b->in_item = ORION_B200_ITEM_C32; b->out_item = ORION_B200_ITEM_U8;
    b->slice_bits = bits_per_symbol;
    if (bits_per_symbol >= 4) {                                      // qam.rs:20-31 with modulate/qam.rs:27-31 axis_scale
        const int m = 1 << (bits_per_symbol / 2);
        const double avg_e_total = 2.0 * (double)(m * m - 1) / 3.0;
        const float scale = (float)(1.0 / sqrt(avg_e_total));
        for (int j = 0; j < m - 1; ++j) b->slice_th[j] = ((float)(2 * j) - (float)(m - 2)) * scale;
    }
    return finish_create(b, out);
}
// The CFO de-rotation call sites (sync/ofdm_sync.rs:527-528, demodulate/ofdm_frame.rs:1489, dvb_t_frame.rs:389): a fresh
// Rotator::new(-cfo_hz, fs).rotate_block(in, out) over one buffer.
int orion_b200_cfo_derotate(float cfo_hz, float fs, const orion_b200_c32 *in, orion_b200_c32 *out, size_t n) {
    orion_b200_block *r = nullptr;
    int st = orion_b200_rotator_create(-cfo_hz, fs, &r);
    if (st != ORION_B200_OK) return st;
    size_t ir = 0, ow = 0;
    st = orion_b200_block_process(r, in, n, out, n, &ir, &ow);
    orion_b200_block_destroy(r);
    return st;
}

// ---- channel bank --------------------------------------------------------------------------------
struct BankGroup {                       // channels that share one demodulator configuration: one batched launch
    orion_b200_block *proto = nullptr;   // rate-1 block with nbatch members (kernel instance, section tables, per-member state)
    std::vector<int> chans;              // bank channel indices of the members
};
struct orion_b200_bank {
    // ---- fast path (K5): one shared front-end kernel + one batched demodulator launch per group ----
    bool fast = false;
    size_t nch = 0;
    int in_item = ORION_B200_ITEM_C32, out_item = ORION_B200_ITEM_F32;
    int mix = MIX_NONE, fir = FIR_NONE;
    size_t M = 1;
    std::vector<float> g;                // generic causal taps
    std::vector<Osc> osc;                // per channel
    NcoParam *d_osc = nullptr;
    float *d_gt = nullptr;
    int PM = 0, H = 0;
    float2 *d_hist[2] = { nullptr, nullptr };
    int pp = 0;
    unsigned long long k_pre = 0;
    float2 *d_z = nullptr;
    size_t z_cap = 0;
    std::vector<BankGroup> groups;
    cudaStream_t stream = nullptr, own_stream = nullptr;
    cudaEvent_t ev_fork = nullptr;
    std::vector<cudaEvent_t> ev_join;
    uint64_t launches = 0;
    // ---- general path: one block per channel ----
    std::vector<orion_b200_block *> ch;
    std::vector<cudaStream_t> streams;
    cudaEvent_t ev_in = nullptr;
    std::vector<cudaEvent_t> ev_done;
    int device = 0;
    void *d_in = nullptr, *d_out = nullptr;
    size_t d_in_cap = 0, d_out_cap = 0;
    int *d_err = nullptr, *h_err = nullptr;       // one watchdog word for all channels
    std::string err;
};
namespace {
const int kBankStreams = 8;
int bank_fail(orion_b200_bank *k, int st, const std::string &what) { if (k) k->err = what; return st; }

// ---- K5 fast path ------------------------------------------------------------------------------------------------
// Eligible when every channel is  [Rotator | Nco mixer] -> FIR with the SAME taps and even decimation -> FM / PM / AM / CW
// (demodulators that only see phase differences or magnitudes, so the closed-form oscillator is exact enough; see
// want_exact_pre).  Anything else keeps the general one-block-per-channel path.
bool bank_fast_eligible(const orion_b200_chain_spec *specs, size_t n) {
    const orion_b200_chain_spec &s0 = specs[0];
    if (getenv("ORION_B200_BANK_GENERAL")) return false;
    if (s0.fir == FIR_NONE || !s0.taps || s0.ntaps == 0 || s0.decim < 2 || (s0.decim & 1)) return false;
    if (s0.ntaps - 1 > 8 * s0.decim || s0.decim > 4096) return false;
    for (size_t c = 0; c < n; ++c) {
        const orion_b200_chain_spec &s = specs[c];
        if (s.struct_size != sizeof(orion_b200_chain_spec)) return false;
        if (s.mix != s0.mix || s.fir != s0.fir || s.decim != s0.decim || s.ntaps != s0.ntaps) return false;
        if (memcmp(s.taps, s0.taps, s0.ntaps * sizeof(float)) != 0) return false;
        if (!(s.demod == DEMOD_FM || s.demod == DEMOD_PM || s.demod == DEMOD_AM || s.demod == DEMOD_AM_ABS || s.demod == DEMOD_CW)) return false;
        if (s.n_post && !s.post_sos) return false;
    }
    return true;
}
std::string bank_group_key(const orion_b200_chain_spec &s) {
    std::string k;
    auto put = [&k](const void *p, size_t n) { k.append((const char *)p, n); };
    put(&s.demod, sizeof(s.demod)); put(&s.fs_demod, 4); put(&s.p0, 4); put(&s.p1, 4); put(&s.audio_bw_hz, 4);
    put(&s.translate, sizeof(s.translate));
    if (s.translate) put(&s.translate_hz, 4);
    put(&s.n_post, sizeof(s.n_post));
    if (s.n_post) put(s.post_sos, s.n_post * 5 * sizeof(float));
    return k;
}
int bank_fast_create(orion_b200_bank *k, const orion_b200_chain_spec *specs, size_t n) {
    const orion_b200_chain_spec &s0 = specs[0];
    if (cudaSetDevice(k->device) != cudaSuccess) return ORION_B200_ERR_CUDA;
    k->fast = true;
    k->nch = n;
    k->mix = s0.mix; k->fir = s0.fir; k->M = s0.decim;
    k->in_item = ORION_B200_ITEM_C32; k->out_item = ORION_B200_ITEM_F32;
    FirPlan pl;
    plan_fir(s0.fir, std::vector<float>(s0.taps, s0.taps + s0.ntaps), s0.decim, true, &pl);   // only the generic taps g[] are used
    k->g = pl.g;
    const int Lg = (int)k->g.size(), M = (int)k->M;
    k->PM = std::max(1, (Lg - 1 + M - 1) / M);
    k->H = ((Lg + 1) & ~1);
    // polyphase table gt[i][p-1] = g[M*p - i]
    std::vector<float> gt((size_t)M * k->PM, 0.f);
    for (int i = 0; i < M; ++i)
        for (int p = 1; p <= k->PM; ++p) {
            const long long t = (long long)M * p - i;
            if (t >= 1 && t < Lg) gt[(size_t)i * k->PM + (p - 1)] = k->g[t];
        }
    k->osc.resize(n);
    std::vector<NcoParam> op(n);
    for (size_t c = 0; c < n; ++c) {
        if (k->mix != MIX_NONE) k->osc[c].set(specs[c].mix_freq_hz, specs[c].mix_fs, 0);
        op[c] = k->osc[c].param(0);
    }
    if (cudaStreamCreateWithFlags(&k->own_stream, cudaStreamNonBlocking) != cudaSuccess) return ORION_B200_ERR_CUDA;
    k->stream = k->own_stream;
    if (cudaMalloc(&k->d_osc, n * sizeof(NcoParam)) != cudaSuccess || cudaMalloc(&k->d_gt, gt.size() * sizeof(float)) != cudaSuccess ||
        cudaMalloc(&k->d_hist[0], k->H * sizeof(float2)) != cudaSuccess || cudaMalloc(&k->d_hist[1], k->H * sizeof(float2)) != cudaSuccess ||
        cudaMalloc(&k->d_err, sizeof(int)) != cudaSuccess || cudaMallocHost(&k->h_err, sizeof(int)) != cudaSuccess)
        return ORION_B200_ERR_ALLOC;
    *k->h_err = 0;
    cudaMemcpyAsync(k->d_osc, op.data(), n * sizeof(NcoParam), cudaMemcpyHostToDevice, k->stream);
    cudaMemcpyAsync(k->d_gt, gt.data(), gt.size() * sizeof(float), cudaMemcpyHostToDevice, k->stream);
    cudaMemsetAsync(k->d_hist[0], 0, k->H * sizeof(float2), k->stream);
    cudaMemsetAsync(k->d_hist[1], 0, k->H * sizeof(float2), k->stream);
    cudaMemsetAsync(k->d_err, 0, sizeof(int), k->stream);
    if (cudaStreamSynchronize(k->stream) != cudaSuccess) return ORION_B200_ERR_CUDA;
    // demodulator groups
    std::vector<std::string> keys;
    for (size_t c = 0; c < n; ++c) {
        const std::string key = bank_group_key(specs[c]);
        size_t gi = 0;
        while (gi < keys.size() && keys[gi] != key) ++gi;
        if (gi == keys.size()) { keys.push_back(key); k->groups.push_back(BankGroup()); }
        k->groups[gi].chans.push_back((int)c);
    }
    for (BankGroup &g : k->groups) {
        const orion_b200_chain_spec &s = specs[g.chans[0]];
        orion_b200_block *b = nullptr;
        int st = new_block(&g.proto, &b);
        if (st != ORION_B200_OK) return st;
        g.proto = nullptr;
        b->nbatch = (int)g.chans.size();
        demod_setup(b, s.demod, s.fs_demod, s.p0, s.p1, s.audio_bw_hz);
        if (s.demod == DEMOD_FM && s.translate) { b->translate = 1; b->post.set(s.translate_hz, s.fs_demod, 0); }
        for (size_t q = 0; q < s.n_post; ++q) b->secs.push_back(sec_biquad(s.post_sos + 5 * q));
        orion_b200_block *made = nullptr;
        st = finish_create(b, &made);
        if (st != ORION_B200_OK) return st;
        g.proto = made;
        made->d_err_ext = k->d_err;
        if (cudaMalloc(&made->d_batch_chan, g.chans.size() * sizeof(int)) != cudaSuccess) return ORION_B200_ERR_ALLOC;
        if (cudaMemcpy(made->d_batch_chan, g.chans.data(), g.chans.size() * sizeof(int), cudaMemcpyHostToDevice) != cudaSuccess) return ORION_B200_ERR_CUDA;
        st = orion_b200_block_set_stream(made, (void *)k->stream);
        if (st != ORION_B200_OK) return st;
    }
    if (cudaEventCreateWithFlags(&k->ev_fork, cudaEventDisableTiming) != cudaSuccess) return ORION_B200_ERR_CUDA;
    k->ev_join.assign(k->groups.size(), nullptr);
    for (cudaEvent_t &e : k->ev_join)
        if (cudaEventCreateWithFlags(&e, cudaEventDisableTiming) != cudaSuccess) return ORION_B200_ERR_CUDA;
    return ORION_B200_OK;
}
int bank_fast_launch(orion_b200_bank *k, const void *d_in, size_t n_in, void *d_out, size_t out_stride,
                     size_t *in_read, size_t *out_written) {
    const size_t M = k->M;
    const size_t n_out_all = (n_in + M - 1) / M;
    const size_t n_out = std::min(n_out_all, out_stride);              // decim.rs:66-75: capped by the output slice
    if (in_read) *in_read = n_in;
    if (out_written) *out_written = n_out;
    if (n_in == 0) return ORION_B200_OK;
    if (n_out_all * k->nch > k->z_cap) {
        cudaStreamSynchronize(k->stream);
        cudaFree(k->d_z); k->d_z = nullptr; k->z_cap = 0;
        const size_t cap = n_out_all * k->nch + 1024;
        if (cudaMalloc(&k->d_z, cap * sizeof(float2)) != cudaSuccess) return bank_fail(k, ORION_B200_ERR_ALLOC, "bank scratch");
        k->z_cap = cap;
    }
    BankFirArgs a;
    memset(&a, 0, sizeof(a));
    a.in = (const float2 *)d_in; a.n_in = (long long)n_in; a.n_out = (long long)n_out_all;
    a.hist_in = k->d_hist[k->pp]; a.hist_out = k->d_hist[k->pp ^ 1]; a.H = k->H;
    a.mix = k->mix; a.osc = k->d_osc; a.kbase = k->k_pre;
    a.M = (int)M; a.Lg = (int)k->g.size(); a.PM = k->PM;
    a.BT = std::max(k->PM, std::max(1, 2048 / (int)M));               // 16 KB tiles (16 blocks of 128 samples)
    const int nwc = bank_fir_consumer_warps((int)k->nch);
    a.NS = nwc > 4 ? 4 : 3;                                            // small CTAs: smaller rings, more CTAs per SM
    if (const char *e = getenv("ORION_B200_BANK_BT")) a.BT = std::max(k->PM, atoi(e));       // experiments
    if (const char *e = getenv("ORION_B200_BANK_NS")) a.NS = std::max(2, std::min(6, atoi(e)));
    a.gt = k->d_gt; a.g0 = k->g[0];
    a.z = k->d_z; a.z_stride = (long long)n_out_all; a.nch = (int)k->nch;
    a.tiles_total = ((long long)n_out_all + a.BT - 1) / a.BT;
    a.err_flag = k->d_err;
    int sms = 148;
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, k->device);
    int per_sm = 2;
    if (bank_fir_prepare(a, nwc, &per_sm) != cudaSuccess || per_sm < 1) return bank_fail(k, ORION_B200_ERR_CUDA, "bank front-end set-up");
    const int ngroups = (int)((k->nch + 32 * nwc - 1) / (32 * nwc));
    long long want = std::max<long long>(1, ((long long)per_sm * sms + ngroups - 1) / ngroups);    // fill every SM once
    if (const char *e = getenv("ORION_B200_BANK_RANGES")) want = std::max(1, atoi(e));
    const int nranges = (int)std::max<long long>(1, std::min<long long>(want, a.tiles_total));   // balanced: ranges differ by at most one tile
    a.nranges = nranges;
    cudaError_t e = bank_fir_launch(a, nranges, nwc, k->stream);
    if (e != cudaSuccess) return bank_fail(k, ORION_B200_ERR_CUDA, std::string("bank front-end launch: ") + cudaGetErrorString(e));
    k->launches += 1;
    k->pp ^= 1;
    k->k_pre += n_in;
    const size_t ob = 4;
    // the demodulator groups are independent: with more than one they run side by side on their own streams between two
    // events on the bank's stream (a small bank's group does not fill the machine by itself)
    const bool fork = k->groups.size() > 1 && k->ev_fork != nullptr && !getenv("ORION_B200_BANK_NO_FORK");
    if (fork) cudaEventRecord(k->ev_fork, k->stream);
    const bool rev = getenv("ORION_B200_BANK_REVERSE") != nullptr;          // experiment: launch order of the demodulator groups
    for (size_t gq = 0; gq < k->groups.size(); ++gq) {
        const size_t gi = rev ? k->groups.size() - 1 - gq : gq;
        BankGroup &g = k->groups[gi];
        if (fork && gi > 0) {
            g.proto->stream = g.proto->own_stream;
            cudaStreamWaitEvent(g.proto->stream, k->ev_fork, 0);
        } else {
            g.proto->stream = k->stream;
        }
        const int st = launch(g.proto, k->d_z, n_out, d_out, n_out, (long long)(n_out_all * sizeof(float2)), (long long)(out_stride * ob));
        if (st != ORION_B200_OK) return bank_fail(k, st, "bank demodulator group: " + g.proto->err);
        if (fork && gi > 0) {
            cudaEventRecord(k->ev_join[gi], g.proto->stream);
            cudaStreamWaitEvent(k->stream, k->ev_join[gi], 0);
        }
    }
    return ORION_B200_OK;
}
int bank_launch_all(orion_b200_bank *k, const void *d_in, size_t n_in, void *d_out, size_t out_stride,
                    size_t *in_read, size_t *out_written) {
    if (k->fast) return bank_fast_launch(k, d_in, n_in, d_out, out_stride, in_read, out_written);
    const size_t ob = out_item_bytes(k->ch[0]);
    size_t r = 0, w = 0;
    for (size_t c = 0; c < k->ch.size(); ++c) {
        const int st = orion_b200_block_process_dev(k->ch[c], d_in, n_in, (char *)d_out + c * out_stride * ob, out_stride, &r, &w);
        if (st != ORION_B200_OK) return bank_fail(k, st, "channel " + std::to_string(c) + ": " + k->ch[c]->err);
    }
    if (in_read) *in_read = r;
    if (out_written) *out_written = w;
    return ORION_B200_OK;
}
}  // namespace

int orion_b200_bank_create(const orion_b200_chain_spec *specs, size_t n_channels, orion_b200_bank **out) {
    if (!out) return ORION_B200_ERR_INVALID;
    *out = nullptr;
    if (!specs || n_channels == 0) return ORION_B200_ERR_INVALID;
    orion_b200_bank *k = new (std::nothrow) orion_b200_bank();
    if (!k) return ORION_B200_ERR_ALLOC;
    k->device = t_device;
    int st = ORION_B200_OK;
    if (bank_fast_eligible(specs, n_channels)) {
        int ndev = 0;
        if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev <= 0) { cudaGetLastError(); delete k; return ORION_B200_ERR_NO_DEVICE; }
        st = bank_fast_create(k, specs, n_channels);
        if (st != ORION_B200_OK) { orion_b200_bank_destroy(k); return st; }
        k->streams.assign(1, k->stream);
        *out = k;
        return ORION_B200_OK;
    }
    for (size_t c = 0; c < n_channels && st == ORION_B200_OK; ++c) {
        orion_b200_block *b = nullptr;
        st = orion_b200_chain_create(&specs[c], &b);
        if (st == ORION_B200_OK) {
            k->ch.push_back(b);
            if (b->in_item != k->ch[0]->in_item || b->out_item != k->ch[0]->out_item ||
                orion_b200_block_decimation(b) != orion_b200_block_decimation(k->ch[0]))
                st = ORION_B200_ERR_UNSUPPORTED;
        }
    }
    if (st == ORION_B200_OK) {
        cudaSetDevice(k->device);
        const int ns = (int)std::min<size_t>(kBankStreams, n_channels);
        k->streams.resize(ns, nullptr);
        for (int i = 0; i < ns && st == ORION_B200_OK; ++i)
            if (cudaStreamCreateWithFlags(&k->streams[i], cudaStreamNonBlocking) != cudaSuccess) st = ORION_B200_ERR_CUDA;
        k->ev_done.resize(ns, nullptr);
        for (int i = 0; i < ns && st == ORION_B200_OK; ++i)
            if (cudaEventCreateWithFlags(&k->ev_done[i], cudaEventDisableTiming) != cudaSuccess) st = ORION_B200_ERR_CUDA;
        if (st == ORION_B200_OK && cudaEventCreateWithFlags(&k->ev_in, cudaEventDisableTiming) != cudaSuccess) st = ORION_B200_ERR_CUDA;
        if (st == ORION_B200_OK && (cudaMalloc(&k->d_err, sizeof(int)) != cudaSuccess || cudaMemset(k->d_err, 0, sizeof(int)) != cudaSuccess || cudaDeviceSynchronize() != cudaSuccess ||
                                    cudaMallocHost(&k->h_err, sizeof(int)) != cudaSuccess)) st = ORION_B200_ERR_ALLOC;
        for (size_t c = 0; c < k->ch.size() && st == ORION_B200_OK; ++c) {
            k->ch[c]->d_err_ext = k->d_err;
            k->ch[c]->opt_overlap = 1;                 // neighbours on a bank stream are different, independent blocks
            st = orion_b200_block_set_stream(k->ch[c], (void *)k->streams[c % ns]);
        }
    }
    if (st != ORION_B200_OK) { orion_b200_bank_destroy(k); return st; }
    *out = k;
    return ORION_B200_OK;
}
void orion_b200_bank_destroy(orion_b200_bank *k) {
    if (!k) return;
    cudaSetDevice(k->device);
    if (k->fast) {
        if (k->stream) cudaStreamSynchronize(k->stream);
        for (BankGroup &g : k->groups) if (g.proto) { g.proto->stream = g.proto->own_stream; orion_b200_block_destroy(g.proto); }
        if (k->ev_fork) cudaEventDestroy(k->ev_fork);
        for (cudaEvent_t e : k->ev_join) if (e) cudaEventDestroy(e);
        cudaFree(k->d_osc); cudaFree(k->d_gt); cudaFree(k->d_hist[0]); cudaFree(k->d_hist[1]); cudaFree(k->d_z);
        cudaFree(k->d_in); cudaFree(k->d_out); cudaFree(k->d_err);
        if (k->h_err) cudaFreeHost(k->h_err);
        if (k->own_stream) cudaStreamDestroy(k->own_stream);
        cudaGetLastError();
        delete k;
        return;
    }
    for (cudaStream_t s : k->streams) if (s) cudaStreamSynchronize(s);
    for (orion_b200_block *b : k->ch) { if (b) { b->stream = b->own_stream; orion_b200_block_destroy(b); } }
    for (cudaStream_t s : k->streams) if (s) cudaStreamDestroy(s);
    for (cudaEvent_t e : k->ev_done) if (e) cudaEventDestroy(e);
    if (k->ev_in) cudaEventDestroy(k->ev_in);
    cudaFree(k->d_in); cudaFree(k->d_out); cudaFree(k->d_err);
    if (k->h_err) cudaFreeHost(k->h_err);
    cudaGetLastError();
    delete k;
}
int orion_b200_bank_reset(orion_b200_bank *k) {
    if (!k) return ORION_B200_ERR_INVALID;
    if (k->fast) {
        if (cudaSetDevice(k->device) != cudaSuccess) return bank_fail(k, ORION_B200_ERR_CUDA, "cudaSetDevice");
        cudaStreamSynchronize(k->stream);
        cudaMemsetAsync(k->d_hist[0], 0, k->H * sizeof(float2), k->stream);
        cudaMemsetAsync(k->d_hist[1], 0, k->H * sizeof(float2), k->stream);
        cudaStreamSynchronize(k->stream);
        k->k_pre = 0;
        for (Osc &o : k->osc) o.reset_phase();
        for (BankGroup &g : k->groups) { const int st = reset_state(g.proto); if (st) return bank_fail(k, st, g.proto->err); }
        return ORION_B200_OK;
    }
    for (orion_b200_block *b : k->ch) { const int st = reset_state(b); if (st) return bank_fail(k, st, b->err); }
    return ORION_B200_OK;
}
int orion_b200_bank_set_stream(orion_b200_bank *k, void *cuda_stream) {
    if (!k) return ORION_B200_ERR_INVALID;
    if (!k->fast) return cuda_stream ? bank_fail(k, ORION_B200_ERR_UNSUPPORTED, "only a bank on the shared front-end path runs on one caller stream") : ORION_B200_OK;
    if (cudaSetDevice(k->device) != cudaSuccess) return bank_fail(k, ORION_B200_ERR_CUDA, "cudaSetDevice");
    cudaStreamSynchronize(k->stream);
    k->stream = cuda_stream ? (cudaStream_t)cuda_stream : k->own_stream;
    k->streams.assign(1, k->stream);
    for (BankGroup &g : k->groups) {
        const int st = orion_b200_block_set_stream(g.proto, (void *)k->stream);
        if (st != ORION_B200_OK) return bank_fail(k, st, g.proto->err);
    }
    return ORION_B200_OK;
}
size_t orion_b200_bank_channels(const orion_b200_bank *k) { return k ? (k->fast ? k->nch : k->ch.size()) : 0; }
const char *orion_b200_bank_last_error(const orion_b200_bank *k) { return k ? k->err.c_str() : "null bank"; }
uint64_t orion_b200_bank_launch_count(const orion_b200_bank *k) {
    uint64_t n = 0;
    if (k) for (const orion_b200_block *b : k->ch) n += b->launches;
    if (k && k->fast) { n = k->launches; for (const BankGroup &g : k->groups) n += g.proto->launches; }
    return n;
}
int orion_b200_bank_process_dev(orion_b200_bank *k, const void *d_in, size_t n_in, void *d_out, size_t out_stride,
                                size_t *in_read, size_t *out_written) {
    if (in_read) *in_read = 0;
    if (out_written) *out_written = 0;
    if (!k || (n_in && !d_in) || (out_stride && !d_out)) return k ? bank_fail(k, ORION_B200_ERR_INVALID, "null buffer") : ORION_B200_ERR_INVALID;
    if (cudaSetDevice(k->device) != cudaSuccess) return bank_fail(k, ORION_B200_ERR_CUDA, "cudaSetDevice");
    return bank_launch_all(k, d_in, n_in, d_out, out_stride, in_read, out_written);
}
int orion_b200_bank_synchronize(orion_b200_bank *k) {
    if (!k) return ORION_B200_ERR_INVALID;
    if (cudaSetDevice(k->device) != cudaSuccess) return bank_fail(k, ORION_B200_ERR_CUDA, "cudaSetDevice");
    for (cudaStream_t s : k->streams)
        if (cudaStreamSynchronize(s) != cudaSuccess) return bank_fail(k, ORION_B200_ERR_CUDA, "cudaStreamSynchronize");
    if (cudaMemcpy(k->h_err, k->d_err, sizeof(int), cudaMemcpyDeviceToHost) != cudaSuccess) return bank_fail(k, ORION_B200_ERR_CUDA, "watchdog read");
    if (*k->h_err != 0) {
        char msg[96];
        snprintf(msg, sizeof(msg), "device watchdog tripped (code %d) in a bank channel", *k->h_err);
        *k->h_err = 0;
        cudaMemset(k->d_err, 0, sizeof(int));
        cudaDeviceSynchronize();
        return bank_fail(k, ORION_B200_ERR_INTERNAL, msg);
    }
    return ORION_B200_OK;
}
int orion_b200_bank_process(orion_b200_bank *k, const void *in, size_t n_in, void *out, size_t out_stride,
                            size_t *in_read, size_t *out_written) {
    if (in_read) *in_read = 0;
    if (out_written) *out_written = 0;
    if (!k || (n_in && !in) || (out_stride && !out)) return k ? bank_fail(k, ORION_B200_ERR_INVALID, "null buffer") : ORION_B200_ERR_INVALID;
    if (n_in == 0) return ORION_B200_OK;
    if (cudaSetDevice(k->device) != cudaSuccess) return bank_fail(k, ORION_B200_ERR_CUDA, "cudaSetDevice");
    const size_t ib = n_in * (k->fast ? 8 : in_item_bytes(k->ch[0]));
    const size_t ob = (k->fast ? k->nch * out_stride * 4 : k->ch.size() * out_stride * out_item_bytes(k->ch[0]));
    if (ib > k->d_in_cap) {
        for (cudaStream_t s : k->streams) cudaStreamSynchronize(s);
        cudaFree(k->d_in); k->d_in = nullptr; k->d_in_cap = 0;
        if (cudaMalloc(&k->d_in, ib + ib / 4 + 256) != cudaSuccess) return bank_fail(k, ORION_B200_ERR_ALLOC, "bank input");
        k->d_in_cap = ib + ib / 4 + 256;
    }
    if (ob > k->d_out_cap) {
        for (cudaStream_t s : k->streams) cudaStreamSynchronize(s);
        cudaFree(k->d_out); k->d_out = nullptr; k->d_out_cap = 0;
        if (cudaMalloc(&k->d_out, ob + 256) != cudaSuccess) return bank_fail(k, ORION_B200_ERR_ALLOC, "bank output");
        k->d_out_cap = ob + 256;
    }
    // the wideband slice goes up once on stream 0; every other stream waits for it
    if (cudaMemcpyAsync(k->d_in, in, ib, cudaMemcpyHostToDevice, k->streams[0]) != cudaSuccess) return bank_fail(k, ORION_B200_ERR_CUDA, "H2D");
    if (k->ev_in) cudaEventRecord(k->ev_in, k->streams[0]);
    for (size_t i = 1; i < k->streams.size(); ++i) cudaStreamWaitEvent(k->streams[i], k->ev_in, 0);
    int st = bank_launch_all(k, k->d_in, n_in, k->d_out, out_stride, in_read, out_written);
    if (st != ORION_B200_OK) return st;
    for (size_t i = 1; i < k->streams.size(); ++i) { cudaEventRecord(k->ev_done[i], k->streams[i]); cudaStreamWaitEvent(k->streams[0], k->ev_done[i], 0); }
    if (ob && cudaMemcpyAsync(out, k->d_out, ob, cudaMemcpyDeviceToHost, k->streams[0]) != cudaSuccess) return bank_fail(k, ORION_B200_ERR_CUDA, "D2H");
    return orion_b200_bank_synchronize(k);
}

// ---- common operations ---------------------------------------------------------------------------
void orion_b200_block_destroy(orion_b200_block *b) {
    if (!b) return;
    cudaSetDevice(b->device);
    if (b->stream) cudaStreamSynchronize(b->stream);
    if (b->osc_stream) cudaStreamSynchronize(b->osc_stream);
    cudaFree(b->d_g); cudaFree(b->d_gtabs);
    for (int i = 0; i < 3; ++i) { cudaFree(b->d_hist[i]); cudaFree(b->d_carry[i]); }
    b->pre.x.free_device(); b->post.x.free_device();
    cudaFree(b->d_links); cudaFree(b->d_ticket); cudaFree(b->d_err); cudaFree(b->d_handoff);
    cudaFree(b->d_in); cudaFree(b->d_out); cudaFree(b->d_batch_chan);
    cudaFree(b->d_fm_tiles);
    for (float *p : b->d_ssb) cudaFree(p);
    for (orion_b200_block *c : b->child) if (c) { c->stream = c->own_stream; orion_b200_block_destroy(c); }
    if (b->h_err) cudaFreeHost(b->h_err);
    for (int i = 0; i < orion_b200_block::kPipeSlots; ++i) {
        if (b->h_stage_in[i]) cudaFreeHost(b->h_stage_in[i]);
        if (b->h_stage_out[i]) cudaFreeHost(b->h_stage_out[i]);
        if (b->ev_h2d[i]) cudaEventDestroy(b->ev_h2d[i]);
        if (b->ev_k[i]) cudaEventDestroy(b->ev_k[i]);
        if (b->ev_d2h[i]) cudaEventDestroy(b->ev_d2h[i]);
    }
    if (b->s_h2d) cudaStreamDestroy(b->s_h2d);
    if (b->s_d2h) cudaStreamDestroy(b->s_d2h);
    if (b->osc_stream) cudaStreamDestroy(b->osc_stream);
    if (b->own_stream) cudaStreamDestroy(b->own_stream);
    cudaGetLastError();
    delete b;
}
int orion_b200_block_reset(orion_b200_block *b) {
    if (!b) return ORION_B200_ERR_INVALID;
    return reset_state(b);
}
const char *orion_b200_block_last_error(const orion_b200_block *b) { return b ? b->err.c_str() : "null block"; }
int orion_b200_block_in_item(const orion_b200_block *b) { return b ? b->in_item : 0; }
int orion_b200_block_out_item(const orion_b200_block *b) { return b ? b->out_item : 0; }
size_t orion_b200_block_decimation(const orion_b200_block *b) { return (b && b->fir != FIR_NONE) ? b->M : 1; }
orion_b200_work_report orion_b200_block_plan(const orion_b200_block *b, size_t n_in, size_t out_cap) {
    orion_b200_work_report wr = { 0, 0 };
    if (b) length_rules(b, n_in, out_cap, &wr.in_read, &wr.out_written);
    return wr;
}

int orion_b200_block_process_dev(orion_b200_block *b, const void *d_in, size_t n_in, void *d_out, size_t out_cap,
                                 size_t *in_read, size_t *out_written) {
    if (in_read) *in_read = 0;
    if (out_written) *out_written = 0;
    if (!b || (n_in && !d_in) || (out_cap && !d_out)) return b ? fail(b, ORION_B200_ERR_INVALID, "null buffer") : ORION_B200_ERR_INVALID;
    size_t consume, produce;
    length_rules(b, n_in, out_cap, &consume, &produce);
    const int st = launch(b, d_in, consume, d_out, produce);
    if (st != ORION_B200_OK) return st;
    if (in_read) *in_read = consume;
    if (out_written) *out_written = produce;
    return ORION_B200_OK;
}

int orion_b200_block_process(orion_b200_block *b, const void *in, size_t n_in, void *out, size_t out_cap,
                             size_t *in_read, size_t *out_written) {
    if (in_read) *in_read = 0;
    if (out_written) *out_written = 0;
    if (!b || (n_in && !in) || (out_cap && !out)) return b ? fail(b, ORION_B200_ERR_INVALID, "null buffer") : ORION_B200_ERR_INVALID;
    size_t consume, produce;
    length_rules(b, n_in, out_cap, &consume, &produce);
    if (consume == 0) return ORION_B200_OK;
    CK(cudaSetDevice(b->device));
    const size_t ib = consume * in_item_bytes(b), ob = produce * out_item_bytes(b);
    if (ib > b->d_in_cap) {
        CK(cudaStreamSynchronize(b->stream));
        cudaFree(b->d_in); b->d_in = nullptr; b->d_in_cap = 0;
        const size_t cap = ib + ib / 4 + 256;
        CK(cudaMalloc(&b->d_in, cap));
        b->d_in_cap = cap;
    }
    if (ob > b->d_out_cap) {
        CK(cudaStreamSynchronize(b->stream));
        cudaFree(b->d_out); b->d_out = nullptr; b->d_out_cap = 0;
        const size_t cap = ob + ob / 4 + 256;
        CK(cudaMalloc(&b->d_out, cap));
        b->d_out_cap = cap;
    }
    // Chunked pipeline: the copy in of chunk i+1, the kernel on chunk i and the copy out of chunk i-1 overlap (three
    // streams, events between them).  Chunks are multiples of the decimation factor (and of the warp tile), so the
    // per-call decimation phase of the reference (decim.rs:44-76) holds across the internal launches, whose streaming
    // state carries over like between process() calls.  Pageable caller memory goes through a pinned staging ring.
    const size_t M = (b->fir != FIR_NONE) ? b->M : 1;
    const bool whole_output = produce == (consume + M - 1) / M;           // a short output slice keeps the one-launch path
    size_t chunk = 0;
    if (whole_output && !getenv("ORION_B200_NO_PIPELINE")) {
        const size_t unit = M * (size_t)kThreads * (size_t)npt_of(b);       // input items of one warp tile
        // ~4 MB of input per chunk from pinned memory; 16 MB when the caller's buffer is pageable (every chunk is one job
        // of the copy pool, whose wake-up cost is amortised over the chunk) ...
        size_t want = (size_t)((is_pageable(in) ? 16u : 4u) << 20) / in_item_bytes(b);
        if (const char *e = getenv("ORION_B200_PIPE_CHUNK_BYTES")) want = std::max<size_t>(1, (size_t)atoll(e) / in_item_bytes(b));
        want = std::max(want, consume / 64 + 1);                            // ... but at most ~64 chunks
        chunk = std::max<size_t>(1, (want + unit - 1) / unit) * unit;
        if (chunk * 2 > consume) chunk = 0;                                 // short calls: one launch
    }
    if (chunk == 0) {
        CK(cudaMemcpyAsync(b->d_in, in, ib, cudaMemcpyHostToDevice, b->stream));
        const int st = launch(b, b->d_in, consume, b->d_out, produce);
        if (st != ORION_B200_OK) return st;
        if (ob) CK(cudaMemcpyAsync(out, b->d_out, ob, cudaMemcpyDeviceToHost, b->stream));
    } else {
        if (!b->s_h2d) {
            CK(cudaStreamCreateWithFlags(&b->s_h2d, cudaStreamNonBlocking));
            CK(cudaStreamCreateWithFlags(&b->s_d2h, cudaStreamNonBlocking));
            for (int i = 0; i < orion_b200_block::kPipeSlots; ++i) {
                CK(cudaEventCreateWithFlags(&b->ev_h2d[i], cudaEventDisableTiming));
                CK(cudaEventCreateWithFlags(&b->ev_k[i], cudaEventDisableTiming));
                CK(cudaEventCreateWithFlags(&b->ev_d2h[i], cudaEventDisableTiming));
            }
        }
        const int NSL = orion_b200_block::kPipeSlots;
        const bool in_pageable = is_pageable(in), out_pageable = ob && is_pageable(out);
        const size_t cib = chunk * in_item_bytes(b), cob = (chunk / M) * out_item_bytes(b);
        if (in_pageable && cib > b->h_stage_in_cap) {
            for (int i = 0; i < NSL; ++i) { if (b->h_stage_in[i]) cudaFreeHost(b->h_stage_in[i]); b->h_stage_in[i] = nullptr; }
            b->h_stage_in_cap = 0;
            for (int i = 0; i < NSL; ++i) CK(cudaMallocHost(&b->h_stage_in[i], cib));
            b->h_stage_in_cap = cib;
        }
        if (out_pageable && cob > b->h_stage_out_cap) {
            for (int i = 0; i < NSL; ++i) { if (b->h_stage_out[i]) cudaFreeHost(b->h_stage_out[i]); b->h_stage_out[i] = nullptr; }
            b->h_stage_out_cap = 0;
            for (int i = 0; i < NSL; ++i) CK(cudaMallocHost(&b->h_stage_out[i], cob));
            b->h_stage_out_cap = cob;
        }
        // the staging streams start after whatever the block's stream still has queued (earlier process_dev calls)
        CK(cudaEventRecord(b->ev_k[0], b->stream));
        CK(cudaStreamWaitEvent(b->s_h2d, b->ev_k[0], 0));
        const size_t nchunks = (consume + chunk - 1) / chunk;
        struct Pending { size_t off_out, n_out; bool live; } pend[orion_b200_block::kPipeSlots] = {};
        auto drain_out = [&](int slot) -> cudaError_t {                     // staged output of an earlier chunk -> the caller's buffer
            if (!pend[slot].live) return cudaSuccess;
            cudaError_t e = cudaEventSynchronize(b->ev_d2h[slot]);
            if (e != cudaSuccess) return e;
            CopyPool::get().copy((char *)out + pend[slot].off_out * out_item_bytes(b), b->h_stage_out[slot], pend[slot].n_out * out_item_bytes(b));
            pend[slot].live = false;
            return cudaSuccess;
        };
        for (size_t c = 0; c < nchunks; ++c) {
            const int slot = (int)(c % NSL);
            const size_t off = c * chunk, cn = std::min(chunk, consume - off);
            const size_t off_out = off / M, pn = (cn + M - 1) / M;
            const char *src = (const char *)in + off * in_item_bytes(b);
            if (in_pageable) {
                if (c >= (size_t)NSL) CK(cudaEventSynchronize(b->ev_h2d[slot]));       // the slot's previous copy in has left it
                CopyPool::get().copy(b->h_stage_in[slot], src, cn * in_item_bytes(b));
                src = (const char *)b->h_stage_in[slot];
            }
            CK(cudaMemcpyAsync((char *)b->d_in + off * in_item_bytes(b), src, cn * in_item_bytes(b), cudaMemcpyHostToDevice, b->s_h2d));
            CK(cudaEventRecord(b->ev_h2d[slot], b->s_h2d));
            CK(cudaStreamWaitEvent(b->stream, b->ev_h2d[slot], 0));
            const int st = launch(b, (char *)b->d_in + off * in_item_bytes(b), cn, (char *)b->d_out + off_out * out_item_bytes(b), pn);
            if (st != ORION_B200_OK) { cudaStreamSynchronize(b->s_h2d); cudaStreamSynchronize(b->s_d2h); return st; }
            if (ob) {
                CK(cudaEventRecord(b->ev_k[slot], b->stream));
                CK(cudaStreamWaitEvent(b->s_d2h, b->ev_k[slot], 0));
                char *dst = (char *)out + off_out * out_item_bytes(b);
                if (out_pageable) {
                    CK(drain_out(slot));
                    dst = (char *)b->h_stage_out[slot];
                    pend[slot].off_out = off_out; pend[slot].n_out = pn; pend[slot].live = true;
                }
                CK(cudaMemcpyAsync(dst, (char *)b->d_out + off_out * out_item_bytes(b), pn * out_item_bytes(b), cudaMemcpyDeviceToHost, b->s_d2h));
                CK(cudaEventRecord(b->ev_d2h[slot], b->s_d2h));
            }
        }
        for (int i = 0; i < NSL; ++i) CK(drain_out(i));
        CK(cudaStreamSynchronize(b->s_d2h));
        CK(cudaStreamSynchronize(b->s_h2d));
    }
    const int es = check_device_error(b);                          // also synchronises the stream
    if (es != ORION_B200_OK) return es;
    if (in_read) *in_read = consume;
    if (out_written) *out_written = produce;
    return ORION_B200_OK;
}

int orion_b200_block_synchronize(orion_b200_block *b) {
    if (!b) return ORION_B200_ERR_INVALID;
    CK(cudaSetDevice(b->device));
    return check_device_error(b);
}
int orion_b200_block_set_stream(orion_b200_block *b, void *cuda_stream) {
    if (!b) return ORION_B200_ERR_INVALID;
    CK(cudaSetDevice(b->device));
    CK(cudaStreamSynchronize(b->stream));
    b->stream = cuda_stream ? (cudaStream_t)cuda_stream : b->own_stream;
    return ORION_B200_OK;
}

int orion_b200_fir_lowpass_iq_filter_aligned(orion_b200_block *b, orion_b200_c32 *io, size_t n) {
    if (!b || b->fir != FIR_IQ || b->M != 1 || b->demod != DEMOD_NONE || b->mix != MIX_NONE)
        return b ? fail(b, ORION_B200_ERR_INVALID, "filter_aligned needs a FirLowpassIq block") : ORION_B200_ERR_INVALID;
    if (n && !io) return fail(b, ORION_B200_ERR_INVALID, "null buffer");
    // fir.rs:260-276: reset, prime with the first d samples, then emit n outputs while feeding
    // io[i+d] (zeros past the end)  ==  stream io ++ zeros(d) through the filter, drop d outputs.
    int st = reset_state(b);
    if (st != ORION_B200_OK || n == 0) return st;
    const size_t d = (b->taps.size() - 1) / 2;                      // fir.rs:216-218
    const size_t tot = n + d;
    float2 *din = nullptr, *dout = nullptr;
    CK(cudaMalloc(&din, tot * sizeof(float2)));
    if (cudaMalloc(&dout, tot * sizeof(float2)) != cudaSuccess) { cudaFree(din); return fail(b, ORION_B200_ERR_ALLOC, "filter_aligned"); }
    cudaMemsetAsync(din, 0, tot * sizeof(float2), b->stream);
    cudaMemcpyAsync(din, io, n * sizeof(float2), cudaMemcpyHostToDevice, b->stream);
    st = launch(b, din, tot, dout, tot);
    if (st == ORION_B200_OK) {
        cudaMemcpyAsync(io, dout + d, n * sizeof(float2), cudaMemcpyDeviceToHost, b->stream);
        st = check_device_error(b);
    }
    cudaStreamSynchronize(b->stream);
    cudaFree(din); cudaFree(dout);
    return st;
}

int orion_b200_block_set_option(orion_b200_block *b, int option, double value) {
    if (!b) return ORION_B200_ERR_INVALID;
    const int v = value != 0.0;
    switch (option) {
        case ORION_B200_OPT_FIR_GLOBAL:
            if (b->opt_force_global != v) { b->opt_force_global = v; b->plan_dirty = true; }
            break;
        case ORION_B200_OPT_USE_TMA: b->opt_use_tma = v; break;
        case ORION_B200_OPT_SERIAL_TILES:
            if (b->opt_serial != v) { b->opt_serial = v; b->plan_dirty = true; }     // the warp-specialised instance has no one-warp mode
            break;
        case ORION_B200_OPT_OVERLAP_LAUNCHES: b->opt_overlap = v; break;
        case ORION_B200_OPT_EXACT_NCO: b->opt_exact = value < 0.0 ? -1 : v; break;
        default: return fail(b, ORION_B200_ERR_INVALID, "unknown option");
    }
    if (b->plan_dirty) {
        // the history length may change with the front; finalize now so errors surface here
        CK(cudaSetDevice(b->device));
        CK(cudaStreamSynchronize(b->stream));
        return finalize_plan(b);
    }
    return ORION_B200_OK;
}

size_t orion_b200_block_get_state(orion_b200_block *b, float *state, size_t cap) {
    const size_t n = 4 + 2 * kMaxSections;
    if (!b || !state || cap < n) return n;
    cudaSetDevice(b->device);
    cudaStreamSynchronize(b->stream);
    CarryState cs;
    if (cudaMemcpy(&cs, b->d_carry[b->pp], sizeof(cs), cudaMemcpyDeviceToHost) != cudaSuccess) return 0;
    state[0] = cs.prev.x; state[1] = cs.prev.y;
    state[2] = (float)(b->k_pre & 0xFFFFFFull); state[3] = (float)(b->k_post & 0xFFFFFFull);
    for (int s = 0; s < kMaxSections; ++s) { state[4 + 2 * s] = cs.sec[s].x; state[5 + 2 * s] = cs.sec[s].y; }
    return n;
}

// ---- checkpoint / resume: the block's complete streaming state as an opaque blob ----------------------------
// (the reference's blocks are plain data and derive Clone: fm.rs:10, iir.rs:4,43,89, rotator.rs:7, fir.rs:176)
namespace {
struct SnapshotHeader {
    uint32_t magic, version;
    uint32_t fir, demod, mix, nsec;
    uint64_t M, ntaps, hist_len;
    unsigned long long k_pre, k_post;
    unsigned long long pre_step, pre_phase0, pre_k0, post_step, post_phase0, post_k0;
    float pre_w[3], post_w[3];
    uint32_t pre_on, post_on;
};
struct SnapshotExact {                         // the exact-replay twin of one oscillator (ExactOsc)
    unsigned long long ctr, nrecent;
    float zr, zi, wre, wim;
};
const uint32_t kSnapMagic = 0x4F423230u;      // "OB20"
const uint32_t kSnapVersion = 2;
}  // namespace

size_t orion_b200_block_snapshot_size(const orion_b200_block *b) {
    if (!b) return 0;
    return sizeof(SnapshotHeader) + sizeof(CarryState) + b->hist_cap * sizeof(float2) +
           2 * sizeof(SnapshotExact) + b->hist_cap * sizeof(float2);
}
int orion_b200_block_snapshot(orion_b200_block *b, void *buf, size_t cap) {
    if (!b || !buf) return ORION_B200_ERR_INVALID;
    if (cap < orion_b200_block_snapshot_size(b)) return fail(b, ORION_B200_ERR_INVALID, "snapshot buffer too small");
    CK(cudaSetDevice(b->device));
    CK(cudaStreamSynchronize(b->stream));
    SnapshotHeader h;
    memset(&h, 0, sizeof(h));
    h.magic = kSnapMagic; h.version = kSnapVersion;
    h.fir = (uint32_t)b->fir; h.demod = (uint32_t)b->demod; h.mix = (uint32_t)b->mix; h.nsec = (uint32_t)b->secs.size();
    h.M = b->M; h.ntaps = b->taps.size(); h.hist_len = b->hist_cap;
    h.k_pre = b->k_pre; h.k_post = b->k_post;
    h.pre_step = b->pre.step; h.pre_phase0 = b->pre.phase0; h.pre_k0 = b->pre.k0;
    h.post_step = b->post.step; h.post_phase0 = b->post.phase0; h.post_k0 = b->post.k0;
    h.pre_w[0] = b->pre.wre; h.pre_w[1] = b->pre.wim; h.pre_w[2] = b->pre.amp_delta;
    h.post_w[0] = b->post.wre; h.post_w[1] = b->post.wim; h.post_w[2] = b->post.amp_delta;
    h.pre_on = b->pre.on; h.post_on = b->post.on;
    char *p = (char *)buf;
    memcpy(p, &h, sizeof(h)); p += sizeof(h);
    CK(cudaMemcpy(p, b->d_carry[b->pp], sizeof(CarryState), cudaMemcpyDeviceToHost)); p += sizeof(CarryState);
    if (b->hist_cap) CK(cudaMemcpy(p, b->d_hist[b->pp], b->hist_cap * sizeof(float2), cudaMemcpyDeviceToHost));
    p += b->hist_cap * sizeof(float2);
    for (const Osc *o : { &b->pre, &b->post }) {
        SnapshotExact xs;
        memset(&xs, 0, sizeof(xs));
        xs.ctr = o->x.ctr; xs.zr = o->x.zr; xs.zi = o->x.zi; xs.wre = o->x.wre; xs.wim = o->x.wim;
        xs.nrecent = (o == &b->pre) ? std::min(o->x.recent.size(), b->hist_cap) : 0;
        memcpy(p, &xs, sizeof(xs)); p += sizeof(xs);
        if (o == &b->pre) {
            memset(p, 0, b->hist_cap * sizeof(float2));
            if (xs.nrecent) memcpy(p, o->x.recent.data() + (o->x.recent.size() - xs.nrecent), xs.nrecent * sizeof(float2));
            p += b->hist_cap * sizeof(float2);
        }
    }
    return ORION_B200_OK;
}
int orion_b200_block_restore(orion_b200_block *b, const void *buf, size_t size) {
    if (!b || !buf || size < sizeof(SnapshotHeader)) return ORION_B200_ERR_INVALID;
    SnapshotHeader h;
    memcpy(&h, buf, sizeof(h));
    if (h.magic != kSnapMagic || h.version != kSnapVersion) return fail(b, ORION_B200_ERR_INVALID, "not a snapshot of this library version");
    if (h.fir != (uint32_t)b->fir || h.demod != (uint32_t)b->demod || h.mix != (uint32_t)b->mix || h.nsec != b->secs.size() ||
        h.M != b->M || h.ntaps != b->taps.size() || h.hist_len != b->hist_cap ||
        size < sizeof(h) + sizeof(CarryState) + 2 * h.hist_len * sizeof(float2) + 2 * sizeof(SnapshotExact))
        return fail(b, ORION_B200_ERR_INVALID, "snapshot was taken from a block of a different shape");
    int st = reset_state(b);                              // drains the stream, zeroes the hand-over counters
    if (st != ORION_B200_OK) return st;
    b->k_pre = h.k_pre; b->k_post = h.k_post;
    b->pre.step = h.pre_step; b->pre.phase0 = h.pre_phase0; b->pre.k0 = h.pre_k0;
    b->post.step = h.post_step; b->post.phase0 = h.post_phase0; b->post.k0 = h.post_k0;
    b->pre.wre = h.pre_w[0]; b->pre.wim = h.pre_w[1]; b->pre.amp_delta = h.pre_w[2];
    b->post.wre = h.post_w[0]; b->post.wim = h.post_w[1]; b->post.amp_delta = h.post_w[2];
    b->pre.on = h.pre_on != 0; b->post.on = h.post_on != 0;
    const char *p = (const char *)buf + sizeof(h);
    CK(dev_upload(b, b->d_carry[b->pp], p, sizeof(CarryState))); p += sizeof(CarryState);
    if (b->hist_cap) CK(dev_upload(b, b->d_hist[b->pp], p, b->hist_cap * sizeof(float2)));
    p += b->hist_cap * sizeof(float2);
    for (Osc *o : { &b->pre, &b->post }) {
        SnapshotExact xs;
        memcpy(&xs, p, sizeof(xs)); p += sizeof(xs);
        o->x.ctr = xs.ctr; o->x.zr = xs.zr; o->x.zi = xs.zi; o->x.wre = xs.wre; o->x.wim = xs.wim;
        o->x.rebase();
        o->x.recent.clear();
        if (o == &b->pre) {
            const float2 *r = reinterpret_cast<const float2 *>(p);
            if (xs.nrecent <= b->hist_cap) o->x.recent.assign(r, r + xs.nrecent);
            p += b->hist_cap * sizeof(float2);
        }
    }
    return ORION_B200_OK;
}

uint64_t orion_b200_block_launch_count(const orion_b200_block *b) { return b ? b->launches : 0; }
double orion_b200_block_exact_host_ms(const orion_b200_block *b) { return b ? b->exact_host_ms : 0.0; }
int orion_b200_block_prepare_oscillator(orion_b200_block *b, size_t n_in_per_call, size_t n_calls) {
    if (!b) return ORION_B200_ERR_INVALID;
    size_t consume = 0, produce = 0;
    length_rules(b, n_in_per_call, (size_t)-1, &consume, &produce);
    timespec t0, t1;
    clock_gettime(CLOCK_MONOTONIC, &t0);
    if (want_exact_pre(b)) b->pre.x.ensure(b->pre.x.ctr + (unsigned long long)consume * n_calls);
    if (want_exact_post(b)) b->post.x.ensure(b->post.x.ctr + (unsigned long long)produce * n_calls);
    clock_gettime(CLOCK_MONOTONIC, &t1);
    b->exact_host_ms += (t1.tv_sec - t0.tv_sec) * 1e3 + (t1.tv_nsec - t0.tv_nsec) * 1e-6;
    return ORION_B200_OK;
}

// debug: per-tile SM clock stamps (8 x int64 per tile, device pointer; NULL disables)
int orion_b200_debug_set_trace(orion_b200_block *b, void *d_trace) {
    if (!b) return ORION_B200_ERR_INVALID;
    b->trace = (long long *)d_trace;
    return ORION_B200_OK;
}

// plan introspection for the host-logic tests [host-only]: fills `info` (12 ints) and, when
// `table` has room, the polyphase tap table as floats.  Returns the table length in floats.
size_t orion_b200_debug_fir_plan(int fir_kind, const float *taps, size_t ntaps, size_t m, int info[12],
                                 float *table, size_t cap, float *g, size_t gcap) {
    if (!taps || !ntaps || !info) return 0;
    FirPlan pl;
    std::vector<float> t(taps, taps + ntaps);
    plan_fir(fir_kind, t, m < 1 ? 1 : m, false, &pl);
    info[0] = pl.front; info[1] = pl.R; info[2] = pl.U; info[3] = pl.Mb; info[4] = pl.O; info[5] = pl.P;
    info[6] = pl.P_pad; info[7] = pl.HR; info[8] = pl.row_samples; info[9] = pl.row_pitch; info[10] = pl.rows;
    info[11] = pl.H;
    const size_t nf = pl.taps2.size() * 2;
    if (table && cap >= nf && nf) memcpy(table, pl.taps2.data(), nf * sizeof(float));
    if (g && gcap >= pl.g.size()) memcpy(g, pl.g.data(), pl.g.size() * sizeof(float));
    return nf;
}
// scan tables of one section group [host-only]: sections = nsec x {type, b0|r|a, b1|1-a, b2, a1, a2};
// out receives {D, depth, agg_only, 0, imp[16][4], lv[5][16], lane[32][16], lb[32][16], lb32[16], tile[16]}.
size_t orion_b200_debug_group_tables(const float *sections, size_t nsec, int npt, float *out, size_t cap) {
    const int MM = kMaxGroupDim * kMaxGroupDim;
    const size_t nf = 4 + kMaxNpt * kMaxGroupDim + 5 * MM + 32 * MM + 32 * MM + MM + MM;
    if (!sections || nsec == 0 || nsec > (size_t)(kMaxGroupDim / 2) || npt < 1 || npt > kMaxNpt || !out || cap < nf) return nf;
    std::vector<SecParam> secs(nsec);
    for (size_t q = 0; q < nsec; ++q) {
        memset(&secs[q], 0, sizeof(SecParam));
        secs[q].type = (int)sections[6 * q];
        for (int i = 0; i < 5; ++i) secs[q].c[i] = sections[6 * q + 1 + i];
    }
    GroupHost gh; gh.first = 0; gh.count = (int)nsec;
    GroupParam *gp = new GroupParam();
    GroupTables *gt = new GroupTables();
    build_group(secs.data(), gh, npt, gp, gt);
    float *o = out;
    *o++ = (float)gp->D; *o++ = (float)gt->depth; *o++ = (float)gp->agg_only; *o++ = 0.f;
    memcpy(o, gp->imp, sizeof(gp->imp)); o += kMaxNpt * kMaxGroupDim;
    memcpy(o, gp->lv, sizeof(gp->lv)); o += 5 * MM;
    memcpy(o, gt->lane, sizeof(gt->lane)); o += 32 * MM;
    memcpy(o, gt->lb, sizeof(gt->lb)); o += 32 * MM;
    memcpy(o, gt->lb32, sizeof(gt->lb32)); o += MM;
    memcpy(o, gt->tile, sizeof(gt->tile)); o += MM;
    delete gp; delete gt;
    return nf;
}

}  // extern "C"
