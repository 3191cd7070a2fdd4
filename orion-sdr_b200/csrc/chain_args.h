// chain_args.h -- launch-argument structures shared by the host code and the sm_100a kernels.
//
// One streaming "chain" = [input-rate mixer] -> [FIR, decimate by M] -> [demod-rate oscillator]
// -> [demodulator front map] -> [recursive sections].  Every
// reference block on the hot path is a degenerate chain, so there is exactly one kernel family.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace orion {

constexpr int kThreads     = 32;    // lanes per (warp) tile: one tile = 32 * NPT outputs
constexpr int kWarpsPerCta = 8;     // default warps per CTA; they share a ring of staged-tile slots
#ifndef ORION_MAX_WARPS
#define ORION_MAX_WARPS 16
#endif
constexpr int kMaxWarpsPerCta = ORION_MAX_WARPS;   // __launch_bounds__(32 * this, 1): 16 -> 128 registers per thread
constexpr int kMaxStages   = 12;    // slots in the ring
constexpr int kMaxSections = 8;     // recursive sections per chain (LR4 = 2, LpDc = 3, + post sections)
constexpr int kMaxTapTable = 2560;  // float capacity of the polyphase tap table held in the parameter bank
constexpr int kMaxRowSamples = 128; // R * Mb limit of the shared-memory staged FIR (row <= 1024 B + pad)

enum : int { SEC_BIQUAD = 1, SEC_DC = 2, SEC_ONEPOLE = 3 };
enum : int { OP_NONE = 0, OP_SQRT = 1, OP_SCALE = 2 };
enum : int { MIX_NONE = 0, MIX_ROTATE = 1, MIX_NCO = 2 };
enum : int { FIR_NONE = 0, FIR_DECIM = 1, FIR_IQ = 2,
             FIR_IQ_UNFUSED = 3 /* FIR_IQ pairing; the bit-faithful front accumulates unfused (HalfCosineMf, fir.rs:358-371) */ };
enum : int { DEMOD_NONE = 0, DEMOD_FM = 1, DEMOD_PM = 2, DEMOD_AM = 3, DEMOD_AM_ABS = 4,
             DEMOD_SSB = 5, DEMOD_CW = 6, DEMOD_USB = 7, DEMOD_F32 = 8 /* f32 in -> sections only */,
             MOD_AM = 9, MOD_PM = 10 /* modulators: f32 in -> C32 out, no sections (modulate/am.rs, pm.rs) */ };
__host__ __device__ inline bool kind_f32_in(int d) { return d == DEMOD_F32 || d == MOD_AM || d == MOD_PM; }
__host__ __device__ inline bool kind_c32_out(int d) { return d == DEMOD_NONE || d == MOD_AM || d == MOD_PM; }
enum : int { FRONT_DIRECT = 0,   // no FIR: items are read straight from global memory
             FRONT_STAGED = 1,   // polyphase FIR on a TMA / cooperatively staged shared-memory tile
             FRONT_GLOBAL = 2 }; // any-shape FIR evaluated from global memory (large M, huge tap sets)

constexpr int kMaxNpt      = 16;    // items per lane (R * U)
constexpr int kMaxGroups   = 4;     // linear section groups per chain
constexpr int kMaxGroupDim = 4;     // state dimension of one group (two second-order sections: 4x4 powers stay in registers)

// One first/second-order recursive section.  Coefficients are the reference's f32 values;
// the per-sample arithmetic in sec_step_t() follows the reference op for op.
struct SecParam {
    int   type;        // SEC_*
    int   post_op;     // OP_* applied to the section's output before the next section
    float c[5];        // BIQUAD: b0,b1,b2,a1,a2 | DC: r | ONEPOLE: a, (1-a)
    float post_scale;  // OP_SCALE factor
};

// A group = a maximal run of sections handed over linearly (post_op == NONE inside the run).
// For the chunked scan the whole run is ONE linear system with state x = (s^0, s^1, ...) of
// dimension D = 2 * count, x' = Ac x + Bc u: its zero-state tile aggregate depends on the
// tile's own input only, so one look-back per group and tile suffices and no aggregate ever
// waits for a predecessor.  All powers are computed on the host in f64 and rounded once;
// n = items per lane, T = 32 * n = items per (warp) tile.
struct GroupParam {
    int   first, count;      // sections [first, first + count)
    int   D;                 // 2 * count
    int   agg_only;          // depth <= 32: predecessors' aggregates alone determine the start state
    int   scan_levels;       // warp-scan levels whose transition power is not negligible (<= 5)
    int   pad_[3];
    float imp[kMaxNpt][kMaxGroupDim];                  // Ac^(n-1-i) Bc: zero-state end state = sum_i imp[i] * u[i]
    float lv[5][kMaxGroupDim * kMaxGroupDim];          // Ac^(n*2^l), l = 0..4, row-major D x D (stride D)
};
struct GroupTables {         // global memory, one per group
    float lane[32][kMaxGroupDim * kMaxGroupDim];       // Ac^(n*lane)   (carry into a lane's chunk)
    float lb[32][kMaxGroupDim * kMaxGroupDim];         // Ac^(T*k)      (inter-tile look-back)
    float lb32[kMaxGroupDim * kMaxGroupDim];           // Ac^(32*T)
    float lbb[33][kMaxGroupDim * kMaxGroupDim];        // Ac^(32*T*k): block-level look-back of slow poles (k = 32: the superblock chain)
    float tile[kMaxGroupDim * kMaxGroupDim];           // Ac^T
    int   depth;             // predecessor tiles with a non-zero weight: Ac^(T*k) == 0 in f32 for k >= depth
    int   pad[3];
};

// Oscillator (Rotator / Nco).  Phase is a 64-bit fraction of a turn:
//   phase(k) = phase0 + step * (k - k0),
// where k counts next() calls since the last reset (the reference advances BEFORE use, so item i
// of a fresh block sees k = i + 1; rotator.rs:44-61).  step = angle(w_f32) / 2pi, w_f32 being the
// reference's f32-rounded (cos, sin) step, so the closed form tracks the reference recurrence to
// its rounding noise.
struct NcoParam {
    unsigned long long step;
    unsigned long long phase0;
    unsigned long long k0;
    unsigned long long kbase;   // k of item 0 of this call, minus 1  (k = kbase + idx + 1)
    float wre, wim;             // unit (cos, sin) of the step angle, for short in-thread recurrences
    float amp_delta;            // ln|w_f32|: |z_k| = 1 + (k mod 1024) * amp_delta (renormalised every 1024)
    int   exact;                // 1: replay the reference's f32 recurrence from the checkpoint tables below (bit-exact phasors)
    // exact-replay mode (rotator.rs:44-61 restated; DESIGN.md "exact oscillator").  Z(c) = the phasor returned by the c-th
    // next() since the last reset; item i of this call sees Z(kbase + i + 1).
    const float2 *xfine;        // xfine[e] = Z(kbase + 1 + 16 e): one checkpoint per 16 items of this call (expanded on the device)
    const float2 *xhist;        // xhist[xhist_len + i] = the phasor that was applied to item i < 0 (FIR history re-mix)
    int   xhist_len;
    int   xfine_len;
    float xwre, xwim;           // the reference's own f32 step w = (cosf(phi), sinf(phi))
};

// exact-replay oscillator: one anchor per 1024 items of a call, produced by the host's sequential walk of the
// reference recurrence; osc_expand_kernel replays from it (chain_kernels.cuh "exact-replay oscillator")
struct OscAnchor { unsigned long long ctr; float2 z; float2 w; unsigned nsteps; unsigned pad; };

struct CarryState {              // streaming state carried between process() calls (device memory)
    float2 prev;                 // discriminator previous sample (fm.rs:16, pm.rs:16)
    float2 sec[kMaxSections];    // recursive-section states
    float2 pad;
};

constexpr int kLinkRecs = 3;     // 16-byte records per link value: {x0, x1, x2, epoch tag} each
struct TileLink {                // one per (tile, group): decoupled look-back records; every record is
    uint4 agg[kLinkRecs];        //   written/read as ONE 128-bit access, so payload and tag travel together.
    uint4 incl[kLinkRecs];       //   agg = tile end state from a zero start state, incl = true end state
};

struct ChainArgs {
    const void *in;              // C32 (or f32 when demod == DEMOD_F32), call-relative item 0
    void       *out;             // f32, or C32 when demod == DEMOD_NONE
    long long   n_in;            // input items consumed by this call
    long long   n_out;           // outputs produced by this call
    // FIR history: the H input samples preceding item 0 (ping-pong across calls)
    const float2 *hist_in;
    float2       *hist_out;
    int   H;
    // input-rate mixer
    int   mix;
    NcoParam pre;
    // FIR in generic causal form  y[j] = sum_{t < Lg} g[t] * x[M*j - t]
    int   fir;
    int   M;
    int   Lg;
    const float *g;              // device copy of g[]
    // staged polyphase plan (FRONT_STAGED)
    int   Mb;                    // samples per block (even) = M * U
    int   O;                     // block origin offset (even, >= M*(U-1))
    int   P_pad;                 // tap steps (multiple of R)
    int   HR;                    // halo rows = P_pad / R
    int   row_samples;           // R * Mb
    int   row_pitch;             // bytes, odd multiple of 16
    int   row_shift;             // log2(row_samples) when it is a power of two, else -1
    int   nstages;               // slots in the CTA's stage ring (1 in serial mode)
    int   ntaps2;                // entries of taps2[] in use
    int   use_tma;               // interior tiles are staged with one cp.async.bulk.tensor
    long long tma_row0;          // global row index of tensor-map row 0
    long long tma_rows;          // rows the tensor map covers
    long long tile_int_lo, tile_int_hi;   // tiles in [lo, hi] are staged by TMA (interior); lo > hi: none
    unsigned int ns_magic;       // ceil(2^32 / nstages): ticket / nstages by multiplication
    int   split;                 // FIR-only staged instance: warps that share the FIR of one tile (1, 2 or 4: slices of the tap rows)
    int   l2_prefetch;           // > 0: every fill also prefetches into L2 the tile this CTA stages that many fills later
    // demodulator
    int   demod;
    int   translate;             // FM: multiply by conj(post phasor) first (fm.rs:34-37)
    float k, k1, k2;             // FM/PM: k | AM_ABS: k1, k2 | MOD_AM: carrier level, modulation index, gain (k = clamp flag) | MOD_PM: k1 = kp, k2 = gain
    float k3;
    NcoParam post;               // demod-rate oscillator (FM translate / SSB BFO / USB mix)
    // recursive sections
    int   nsec;
    int   ngroups;
    int   pipe_park_slots;       // per-warp scan-state slots (2, or ngroups + 1 for the multi-group pipeline)
    int   pipe_u_slots;          // per-warp item slots of the multi-group pipeline (0: single-group path)
    int   nstage;                // stages of that pipeline (= pipe_u_slots)
    int   stage_group[8];        // stage k finishes group stage_group[k-1]; -1: idle stage (slack in front of a slow-pole group)
    GroupParam grp[kMaxGroups];  // grp immediately followed by sec: the kernels view the pair as one block
    SecParam sec[kMaxSections];
    const GroupTables *gtabs;    // [ngroups]
    // carried state, ping-pong across calls
    const CarryState *carry_in;
    CarryState       *carry_out;
    // inter-tile links
    TileLink *links;             // [ntiles][kMaxGroups]
    unsigned int epoch;
    int   ntiles;
    int   serial;                // debug: tiles run one after another (grid = 1)
    int  *err_flag;
    int   pdl_guard;             // tiles below this index read state handed over by the previous call (prev, look-back)
    unsigned int *handoff;       // [0] calls whose FIR history is written, [1] 2 x calls whose carried state is complete,
                                 // [2], [3] CTAs that have run to their end, one counter per launch parity: a CTA of call
                                 // N+1 may well finish before the last CTA of call N, so one counter for all calls could
                                 // reach call N+2's target while call N still had CTAs running (seen as link records of
                                 // epoch N+2 under the eyes of call N: scripts/stress_overlap.py)
    unsigned int hist_target;    // values of the two counters this call needs before it reads the hand-over (0: none)
    unsigned int carry_target;
    unsigned int depth_target;   // value of handoff[depth_slot] that proves the launch before the previous one has ended
                                 // (all of its tiles: nobody reads its buffers or its half of the link records any more)
    int depth_slot;              // 2 + launch parity: the CTA-done counter this launch waits on AND counts into
    // batched launch (channel bank): blockIdx.y = member; member m works on channel batch_chan[m]
    int   batch;                 // members (1: plain launch)
    const int *batch_chan;       // [batch] channel index of every member (row of the strided input / output)
    long long batch_in_stride;   // bytes between the inputs of two channels
    long long batch_out_stride;  // bytes between the outputs of two channels
    long long batch_links_stride;// TileLink records between two members; carry_in / carry_out are arrays indexed by member
    long long *trace;            // debug: 8 x int64 per tile {consume, ready, fir_done, front_done, finish_done, -, smid, warp}
    float2 taps2[kMaxTapTable / 2];   // [u][q][c] -> (g[t0], g[t0-1]); see DESIGN.md "staged FIR"
};

}  // namespace orion
