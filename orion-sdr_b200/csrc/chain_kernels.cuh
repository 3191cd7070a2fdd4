#pragma once
// chain_kernels.cuh -- the sm_100a streaming-chain kernel family (templates; instantiated per translation unit
// by chain_inst_*.cu so the instances compile in parallel).
//
//   [input-rate mixer] -> [FIR, keep every M-th] -> [demod-rate oscillator] -> [demod front map]
//   -> [recursive sections]                                       one launch, one pass over HBM.
//
// Work decomposition: the output stream is cut into WARP TILES of 32*NPT items.
// One persistent CTA of up to 16 warps per SM; the CTA's tiles (static round-robin over the grid) are taken by
// its warps in ticket order, and every lane owns NPT = R*U CONSECUTIVE output items, so
//   * the polyphase FIR slides an R-deep register window over the staged input (each staged
//     sample is read once per lane and reused R times),
//   * the discriminator needs one neighbour value per lane (a shuffle),
//   * every recursive section is a per-lane serial recursion (the reference's exact arithmetic)
//     stitched together by a warp-level state-space scan and an inter-tile decoupled look-back.
// There is no block-level barrier in the steady state: a warp that waits (TMA, look-back) stalls alone.
// Input staging: one cp.async.bulk.tensor (TMA) per interior tile into a slot of a ring shared by the CTA's
// warps (padded row layout); the warp that has run the FIR on a slot refills it at once with the tile NS
// places ahead; edge tiles (FIR history, ragged tail) use a cooperative loader into the same layout.
// Rate-1 blocks (FRONT_DIRECT) move whole warp tiles with lane-consecutive 16-byte accesses through a per-warp
// transposing scratch instead.  Consecutive launches of a block may overlap (programmatic dependent launch);
// what one call hands to the next is guarded by two counters (handoff_wait / handoff_signal).
//
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -fmad=false (FMA only where written).
#include "chain_args.h"

#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <string.h>

#ifndef ORION_TRACE
#define ORION_TRACE 0
#endif

namespace orion {

#define DEV __device__ __forceinline__
#define FULLMASK 0xffffffffu

// Shape specialisation.  SP = 0 keeps every staged-geometry value a launch argument; SP = 1 fixes the geometry
// of the decimate-by-8, <= 64-tap shape (C1 / C3: Mb = 8, one halo row, 8 tap steps, 528-byte rows), so the FIR
// loops unroll and every shared-memory offset folds into the instruction.
template <int SP> struct Geo { static constexpr bool fixed = false; static constexpr int Mb = 0, HR = 0, P_pad = 0, pitch = 0; };
template <> struct Geo<1> { static constexpr bool fixed = true; static constexpr int Mb = 8, HR = 1, P_pad = 8, pitch = 528; };
// SP = 2: the decimate-by-32, <= 1024-tap shape (C4: R = 4, Mb = 32, eight halo rows, 32 tap steps, 1040-byte rows)
template <> struct Geo<2> { static constexpr bool fixed = true; static constexpr int Mb = 32, HR = 8, P_pad = 32, pitch = 1040; };
// Demodulator specialisation.  DM = -1: demodulator kind and section structure are launch arguments;
// DM = DEMOD_NONE (0): C32 out, no sections; DM = DM_FM_LR4: FM discriminator followed by exactly one group of
// two biquads (the LR4 of fm.rs:27) and nothing else -- the C1 chain.
// DM = DM_LR4 + kind: that demodulator kind followed by exactly one group of two biquads and nothing else (the LR4
// of fm.rs:27 / pm.rs:27, or LpCascade itself when the kind is DEMOD_F32).
constexpr int DM_LR4 = 100;
constexpr int DM_FM_LR4 = DM_LR4 + DEMOD_FM;
template <int DM> struct Dm {
    static constexpr bool fixed = DM >= 0;
    static constexpr bool lr4 = DM >= DM_LR4;
    static DEV int demod(const ChainArgs &a) { return DM < 0 ? a.demod : (DM >= DM_LR4 ? DM - DM_LR4 : DM); }
};

// ----------------------------------------------------------------------------------------------
// small helpers
// ----------------------------------------------------------------------------------------------
DEV float2 mv(const float4 m, const float2 v) {              // 2x2 (row-major) times vector
    return make_float2(fmaf(m.x, v.x, m.y * v.y), fmaf(m.z, v.x, m.w * v.y));
}
DEV float4 mm(const float4 a, const float4 b) {              // a * b
    return make_float4(fmaf(a.x, b.x, a.y * b.z), fmaf(a.x, b.y, a.y * b.w),
                       fmaf(a.z, b.x, a.w * b.z), fmaf(a.z, b.y, a.w * b.w));
}
DEV float2 add2(float2 a, float2 b) { return make_float2(a.x + b.x, a.y + b.y); }
DEV float2 scale2(float2 z, float a) { return make_float2(z.x * a, z.y * a); }
DEV float2 shfl_up2(float2 v, int d) {
    return make_float2(__shfl_up_sync(FULLMASK, v.x, d), __shfl_up_sync(FULLMASK, v.y, d));
}
DEV float2 shfl2(float2 v, int src) {
    return make_float2(__shfl_sync(FULLMASK, v.x, src), __shfl_sync(FULLMASK, v.y, src));
}

DEV uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }

// Packed FP32x2 FMA (sm_100 FFMA2): d = a * b + c on both halves of a 64-bit register pair, each
// half a correctly rounded IEEE fma.  With a = (t, t) the SASS form takes t as a scalar operand.
typedef unsigned long long f32x2;
DEV f32x2 pack2(float lo, float hi) {
    f32x2 r;
    asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi));
    return r;
}
DEV float2 unpack2(f32x2 v) {
    float2 r;
    asm("mov.b64 {%0, %1}, %2;" : "=f"(r.x), "=f"(r.y) : "l"(v));
    return r;
}
DEV f32x2 ffma2(f32x2 a, f32x2 b, f32x2 c) {
    f32x2 d;
    asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c));
    return d;
}

// mbarrier / TMA (PTX ISA: mbarrier, cp.async.bulk.tensor)
DEV void mbar_init(uint32_t mbar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(mbar), "r"(count) : "memory");
}
DEV void mbar_expect_tx(uint32_t mbar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(mbar), "r"(bytes) : "memory");
}
DEV bool mbar_try_wait(uint32_t mbar, uint32_t parity) {
    uint32_t ok;
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(ok) : "r"(mbar), "r"(parity) : "memory");
    return ok != 0;
}
// the same with a suspend-time hint (ns): the waiting warp stays off the issue slots for up to that long per try
DEV bool mbar_try_wait_hint(uint32_t mbar, uint32_t parity, uint32_t ns) {
    uint32_t ok;
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2, %3;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(ok) : "r"(mbar), "r"(parity), "r"(ns) : "memory");
    return ok != 0;
}
DEV bool mbar_test_wait(uint32_t mbar, uint32_t parity) {          // one non-blocking look at the phase
    uint32_t ok;
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.test_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(ok) : "r"(mbar), "r"(parity) : "memory");
    return ok != 0;
}
DEV void tma_load_2d(uint32_t dst, const CUtensorMap *map, int c0, int c1, uint32_t mbar) {
    asm volatile(
        "cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes"
        " [%0], [%1, {%2, %3}], [%4];"
        ::"r"(dst), "l"(map), "r"(c0), "r"(c1), "r"(mbar) : "memory");
}
// Overlapping consecutive calls (programmatic dependent launch).  With the launch attribute set, the CTAs of call
// N+1 are scheduled as SMs drain from call N.  Nothing waits for the whole previous grid: the few items call N
// hands to call N+1 -- FIR history (written when N starts), discriminator `prev` and section states (written by
// N's last tile) -- are guarded by two monotonic counters in global memory, bumped with release semantics by the
// writers and polled with acquire semantics by the handful of warps of N+1 that read them; those reads bypass L1.
DEV void griddep_launch_dependents() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
DEV unsigned ld_acquire_u32(const unsigned *p) {
    unsigned v;
    asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
DEV void handoff_wait(const ChainArgs &a, int which, unsigned target) {          // whole warp; which: 0 history, 1 carried state
    if (target == 0u) return;
    int spins = 0;
    while ((int)(ld_acquire_u32(a.handoff + which) - target) < 0) {
        if (++spins > (1 << 22)) { atomicExch(a.err_flag, 4); break; }           // watchdog: never hang the device
        __nanosleep(100);
    }
    __syncwarp();
}
DEV void handoff_signal(const ChainArgs &a, int which, int lane) {               // after the warp's hand-over stores
    __threadfence();
    __syncwarp();
    if (lane == 0) atomicAdd(a.handoff + which, 1u);
    __syncwarp();
}
DEV void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

// Predicated forms for work that one elected lane does.  A C++ `if (lane == 0) { ... }` around these
// would be a divergent branch; ptxas was seen to leave lane 0 split from the warp for the rest of the
// loop iteration, after which every warp-synchronous shuffle takes its divergent slow path (~300
// cycles each).  Predication keeps the warp converged.
DEV void tma_fill_pred(bool p, uint32_t dst, const CUtensorMap *map, int c0, int c1, uint32_t mbar, uint32_t bytes) {
    asm volatile(
        "{\n\t.reg .pred p;\n\tsetp.ne.u32 p, %0, 0;\n\t"
        "@p fence.proxy.async.shared::cta;\n\t"
        "@p mbarrier.arrive.expect_tx.shared::cta.b64 _, [%5], %6;\n\t"
        "@p cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%1], [%2, {%3, %4}], [%5];\n\t}"
        ::"r"((uint32_t)p), "r"(dst), "l"(map), "r"(c0), "r"(c1), "r"(mbar), "r"(bytes) : "memory");
}
// L2 prefetch of a tile the stage ring will ask for later: the ring holds ~10 tiles per SM, which at 16 warps is only
// ~2.5 k cycles of look-ahead -- about one loaded HBM round trip, so warps used to wait for their slot (trace: 2.5 k
// cycles per tile).  With the tile already in L2 the TMA fill is an L2 hit.
DEV void tma_prefetch_l2_pred(bool p, const CUtensorMap *map, int c0, int c1) {
    asm volatile(
        "{\n\t.reg .pred p;\n\tsetp.ne.u32 p, %0, 0;\n\t"
        "@p cp.async.bulk.prefetch.tensor.2d.L2.global.tile [%1, {%2, %3}];\n\t}"
        ::"r"((uint32_t)p), "l"(map), "r"(c0), "r"(c1) : "memory");
}
DEV void mbar_arrive_pred(bool p, uint32_t mbar) {
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.u32 p, %0, 0;\n\t@p mbarrier.arrive.shared::cta.b64 _, [%1];\n\t}"
                 ::"r"((uint32_t)p), "r"(mbar) : "memory");
}
DEV void st_shared_volatile_pred(bool p, uint32_t addr, int v) {
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.u32 p, %0, 0;\n\t@p st.volatile.shared.s32 [%1], %2;\n\t}"
                 ::"r"((uint32_t)p), "r"(addr), "r"(v) : "memory");
}
DEV unsigned atom_add_shared_pred(bool p, uint32_t addr, unsigned v) {
    unsigned old = 0;
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.u32 p, %1, 0;\n\t@p atom.shared.add.u32 %0, [%2], %3;\n\t}"
                 : "+r"(old) : "r"((uint32_t)p), "r"(addr), "r"(v) : "memory");
    return old;
}

// 128-bit single-copy-atomic global accesses for the look-back records
DEV uint4 ld_relaxed_b128(const void *p) {
    uint4 v;
    asm volatile(
        "{\n\t.reg .b128 t;\n\t.reg .b64 lo, hi;\n\t"
        "ld.relaxed.gpu.global.b128 t, [%4];\n\t"
        "mov.b128 {lo, hi}, t;\n\t"
        "mov.b64 {%0, %1}, lo;\n\t"
        "mov.b64 {%2, %3}, hi;\n\t}"
        : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(p) : "memory");
    return v;
}
DEV void st_relaxed_b128(bool pred, void *p, uint4 v) {        // predicated: see tma_fill_pred
    asm volatile(
        "{\n\t.reg .b128 t;\n\t.reg .b64 lo, hi;\n\t.reg .pred q;\n\t"
        "setp.ne.u32 q, %5, 0;\n\t"
        "mov.b64 lo, {%1, %2};\n\t"
        "mov.b64 hi, {%3, %4};\n\t"
        "mov.b128 t, {lo, hi};\n\t"
        "@q st.relaxed.gpu.global.b128 [%0], t;\n\t}"
        ::"l"(p), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w), "r"((uint32_t)pred) : "memory");
}

// ----------------------------------------------------------------------------------------------
// oscillator: closed-form phase from the absolute call counter (NcoParam in chain_args.h)
// ----------------------------------------------------------------------------------------------
DEV float2 nco_unit(const NcoParam &p, unsigned long long k) {
    const unsigned long long ph = p.phase0 + p.step * (k - p.k0);
    // top 32 bits as a signed fraction of pi in [-1, 1); the float conversion keeps 24 of them,
    // the dropped bits come back as a first-order rotation
    const int hi = (int)(unsigned)(ph >> 32);
    const float xh = (float)hi;
    const long long rem = ((long long)ph >> 8) - ((long long)xh << 24);   // exact, 2^-56 turn units
    float s, c;
    __sincosf(xh * (4.656612873077393e-10f * 3.14159265358979323846f), &s, &c);   // MUFU: |err| <= 2^-21.4 on [-pi, pi]
    const float d = (float)rem * (3.14159265358979f * 2.7755575615628914e-17f);   // rem * 2^-55 * pi
    return make_float2(fmaf(-s, d, c), fmaf(c, d, s));
}
// |z_k| of the reference recurrence: |w_f32|^(k mod 1024), renormalised to 1 every 1024 steps
DEV float nco_amp(const NcoParam &p, unsigned long long k) {
    return fmaf((float)(unsigned)(k & 1023ull), p.amp_delta, 1.0f);
}
DEV float2 nco_phasor(const NcoParam &p, unsigned long long k) { return scale2(nco_unit(p, k), nco_amp(p, k)); }
DEV float2 cmul_fma(float2 z, float2 w) {   // phasor advance, reference form (rotator.rs:46-47)
    return make_float2(fmaf(z.x, w.x, -(z.y * w.y)), fmaf(z.y, w.x, z.x * w.y));
}

// ---- exact-replay oscillator -----------------------------------------------------------------------------------
// The reference phasor is the f32 recurrence z <- z*w with a renormalisation every 1024 steps (rotator.rs:44-61,
// nco.rs:42-58).  Its rounding makes it drift away from any closed form (5e-2 rad after 24 M steps), which matters
// wherever the ABSOLUTE phase reaches the output (Rotator, NcoMixer, mix_usb_block, SsbProductDemod, the modulators).
// The sequence does not depend on the data, so the host walks the recurrence once (one checkpoint per 1024 steps,
// orion_b200_api.cu ExactOsc), a small kernel expands the checkpoints to one per 16 items (osc_expand_kernel below),
// and every lane replays the reference's own arithmetic from the nearest one: phasors are bit-identical.
DEV void nco_step_exact(float2 &z, const float2 w, unsigned &ctr) {
    z = cmul_fma(z, w);                                            // rotator.rs:46-48
    ctr += 1u;
    if ((ctr & 0x3FFu) == 0u) {                                    // rotator.rs:51-59: r2.sqrt().recip()
        const float r2 = z.x * z.x + z.y * z.y;
        const float inv = 1.0f / sqrtf(r2);
        z.x *= inv;
        z.y *= inv;
    }
}
// Z(kbase + 1 + idx), idx >= 0: the phasor item idx of this call sees; ctr is left at kbase + 1 + idx.
// `steps` (>= idx & 15) is the replay trip count: pass the warp maximum where the warp must stay converged.
DEV float2 nco_exact_at(const NcoParam &p, long long idx, unsigned &ctr, int steps) {
    long long e = idx >> 4;
    int r = (int)(idx & 15);
    if (e >= p.xfine_len) { e = p.xfine_len - 1; r = 0; }         // past the end of the call: the item is zero padding
    float2 z = __ldg(p.xfine + e);
    ctr = (unsigned)(p.kbase + 1ull + ((unsigned long long)e << 4));
    const float2 w = make_float2(p.xwre, p.xwim);
    for (int i = 0; i < steps; ++i) {
        float2 zn = z;
        unsigned cn = ctr;
        nco_step_exact(zn, w, cn);
        if (i < r) { z = zn; ctr = cn; }
    }
    return z;
}
DEV int warp_max_replay(long long idx) {                           // whole warp: max over lanes of idx & 15
    return (int)__reduce_max_sync(FULLMASK, (unsigned)(idx & 15));
}
// the phasor that was applied to item idx < 0 when it was first consumed (FIR history re-mix)
DEV float2 nco_exact_hist(const NcoParam &p, long long idx) {
    const long long h = idx + p.xhist_len;
    return h >= 0 ? __ldcg(p.xhist + h) : make_float2(1.f, 0.f);
}
DEV float2 nco_exact_any(const NcoParam &p, long long idx) {      // any lane, any context (slow paths)
    unsigned ctr;
    return idx < 0 ? nco_exact_hist(p, idx) : nco_exact_at(p, idx, ctr, (int)(idx & 15));
}

// checkpoint expansion: thread t replays up to `nsteps` steps from anchor t and writes Z(c0 + 16 e) for every e it passes
__global__ void osc_expand_kernel(const OscAnchor *an, int n_an, float2 *fine, unsigned long long c0, long long fine_len);

// input-rate mixers
DEV float2 mix_apply(int mix, float2 x, float2 p) {
    if (mix == MIX_ROTATE)                                         // rotator.rs:74-84
        return make_float2(fmaf(x.x, p.x, -(x.y * p.y)), fmaf(x.y, p.x, x.x * p.y));
    return make_float2(x.x * p.x - x.y * p.y, x.x * p.y + x.y * p.x);   // nco.rs:63-66
}

// n / d rounded to nearest like the IEEE division the reference performs, without the compiler's out-of-line slow
// path (a CALL with spills around it in the middle of the demodulator loop): reciprocal refined by one Newton step,
// quotient corrected with the exact FMA remainder (Markstein).  Operands here are finite with d >= 2^-23 (an epsilon
// is added to the denominator) and 0 <= n <= d, where the sequence returns the correctly rounded quotient.
DEV float div_rn_fast(float n, float d) {
    float r0;
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r0) : "f"(d));
    const float r1 = fmaf(fmaf(-d, r0, 1.0f), r0, r0);
    const float q = n * r1;
    const float rem = fmaf(-d, q, n);
    return fmaf(rem, r1, q);
}

// util.rs:305-322, op for op
DEV float atan2_approx(float y, float x) {
    const float ax = fabsf(x), ay = fabsf(y);
    const bool sw = ax < ay;
    const float mn = sw ? ax : ay, mx = sw ? ay : ax;
    const float r = div_rn_fast(mn, mx + 1.1920929e-07f);
    const float r2 = r * r;
    float phi = r * (0.78539816339744830962f + r2 * (-0.2447f + r2 * 0.0663f));
    if (sw) phi = 1.57079632679489661923f - phi;
    const float sgn = (y < 0.0f) ? -1.0f : 1.0f;
    if (x < 0.0f) return (3.14159265358979323846f - phi) * sgn;
    return phi * sgn;
}

DEV float post_apply(const SecParam &P, float y) {
    if (P.post_op == OP_SQRT) return sqrtf(y);
    if (P.post_op == OP_SCALE) return y * P.post_scale;
    return y;
}

// Warp-tile transposing copies for the rate-1 blocks.  A lane owns CHUNKS consecutive 16-byte chunks of the warp
// tile (its NPT items), but global memory is touched with LANE-consecutive chunks -- one instruction moves 512
// contiguous bytes -- and the re-layout goes through a per-warp shared-memory scratch whose row pitch is an odd
// number of chunks, so both directions are bank-conflict free.
constexpr int kXposeBytes = kThreads * (8 * 16 + 16);          // per warp: 32 rows of 8 chunks + pad
template <int CHUNKS>
DEV void warp_tile_load(const void *gsrc, unsigned char *xs, int lane, float4 (&v)[CHUNKS]) {
    constexpr int PITCH = CHUNKS * 16 + 16;
    const float4 *g = reinterpret_cast<const float4 *>(gsrc);
    float4 t[CHUNKS];
#pragma unroll
    for (int k = 0; k < CHUNKS; ++k) t[k] = __ldg(g + k * kThreads + lane);
#pragma unroll
    for (int k = 0; k < CHUNKS; ++k) {
        const int c = k * kThreads + lane;
        *reinterpret_cast<float4 *>(xs + (c / CHUNKS) * PITCH + (c % CHUNKS) * 16) = t[k];
    }
    __syncwarp();
#pragma unroll
    for (int k = 0; k < CHUNKS; ++k) v[k] = *reinterpret_cast<const float4 *>(xs + lane * PITCH + k * 16);
    __syncwarp();
}
template <int CHUNKS>
DEV void warp_tile_store(void *gdst, unsigned char *xs, int lane, const float4 (&v)[CHUNKS]) {
    constexpr int PITCH = CHUNKS * 16 + 16;
    float4 *g = reinterpret_cast<float4 *>(gdst);
#pragma unroll
    for (int k = 0; k < CHUNKS; ++k) *reinterpret_cast<float4 *>(xs + lane * PITCH + k * 16) = v[k];
    __syncwarp();
#pragma unroll
    for (int k = 0; k < CHUNKS; ++k) {
        const int c = k * kThreads + lane;
        g[c] = *reinterpret_cast<const float4 *>(xs + (c / CHUNKS) * PITCH + (c % CHUNKS) * 16);
    }
    __syncwarp();
}

// virtual input stream: FIR history for negative indices, zeros past the end of the call
DEV float2 load_x(const ChainArgs &a, long long s) {
    if (s < 0) {
        const long long h = s + a.H;
        return h >= 0 ? __ldcg(a.hist_in + h) : make_float2(0.f, 0.f);
    }
    if (s < a.n_in) return __ldg(reinterpret_cast<const float2 *>(a.in) + s);
    return make_float2(0.f, 0.f);
}
DEV float2 load_x_mixed(const ChainArgs &a, long long s) {
    float2 x = load_x(a, s);
    if (a.mix != MIX_NONE)
        x = mix_apply(a.mix, x, a.pre.exact ? nco_exact_any(a.pre, s) : nco_phasor(a.pre, a.pre.kbase + (unsigned long long)s + 1ull));
    return x;
}

// Shared-memory copy of the per-section / per-group launch data.  Indexed reads of the kernel
// parameter bank (LDC with a register offset) are not pipelined; broadcast LDS reads are.
struct Hot {
    GroupParam grp[kMaxGroups];
    SecParam sec[kMaxSections];
};

// ----------------------------------------------------------------------------------------------
// section groups: D-dimensional state-space scan (GroupParam / GroupTables in chain_args.h)
// ----------------------------------------------------------------------------------------------
template <int D>
DEV void matvec(const float *__restrict__ M, const float (&v)[D], float (&out)[D]) {   // row-major D x D
#pragma unroll
    for (int r = 0; r < D; ++r) {
        float acc = 0.f;
#pragma unroll
        for (int c = 0; c < D; ++c) acc = fmaf(M[r * D + c], v[c], acc);
        out[r] = acc;
    }
}
// A cascade's transition matrix is block lower-triangular in the state order (s^0, s^1): section 0 never sees section 1,
// and every power keeps that shape -- the upper-right 2x2 block of a D = 4 power is exactly zero.  12 FMAs instead of 16.
DEV void matvec4_tri(const float *__restrict__ M, const float (&v)[4], float (&out)[4]) {
    out[0] = fmaf(M[1], v[1], fmaf(M[0], v[0], 0.f));          // the same operation sequence as matvec<4>, minus the exact zeros
    out[1] = fmaf(M[5], v[1], fmaf(M[4], v[0], 0.f));
    out[2] = fmaf(M[11], v[3], fmaf(M[10], v[2], fmaf(M[9], v[1], fmaf(M[8], v[0], 0.f))));
    out[3] = fmaf(M[15], v[3], fmaf(M[14], v[2], fmaf(M[13], v[1], fmaf(M[12], v[0], 0.f))));
}
template <int D>
DEV void load_mat(const float *__restrict__ g, float (&M)[D * D]) {      // D*D floats, 16-byte aligned
#pragma unroll
    for (int i = 0; i < D * D; i += 4) {
        const float4 v = __ldg(reinterpret_cast<const float4 *>(g + i));
        M[i] = v.x; M[i + 1] = v.y; M[i + 2] = v.z; M[i + 3] = v.w;
    }
}

// A link value is D floats in ceil(D/3) records of 16 bytes {x, x, x, epoch tag}; every record is
// written and read with ONE 128-bit access, so payload and tag can never be observed apart (no
// fences, one L2 round trip per window of 32 predecessors).
template <int D>
DEV void publish(bool pred, const ChainArgs &a, long long tile, int g, const float (&v)[D], bool inclusive) {
    TileLink *lk = a.links + tile * kMaxGroups + g;
    uint4 *dst = inclusive ? lk->incl : lk->agg;
#pragma unroll
    for (int r = 0; r < (D + 2) / 3; ++r) {
        uint4 rec = make_uint4(0u, 0u, 0u, a.epoch);
        rec.x = __float_as_uint(v[3 * r]);
        if (3 * r + 1 < D) rec.y = __float_as_uint(v[3 * r + 1]);
        if (3 * r + 2 < D) rec.z = __float_as_uint(v[3 * r + 2]);
        st_relaxed_b128(pred, dst + r, rec);
    }
}
template <int D>
DEV bool read_link(const uint4 *src, unsigned epoch, float (&v)[D]) {
    bool ok = true;
#pragma unroll
    for (int r = 0; r < (D + 2) / 3; ++r) {
        const uint4 rec = ld_relaxed_b128(src + r);
        ok = ok && rec.w == epoch;
        v[3 * r] = __uint_as_float(rec.x);
        if (3 * r + 1 < D) v[3 * r + 1] = __uint_as_float(rec.y);
        if (3 * r + 2 < D) v[3 * r + 2] = __uint_as_float(rec.z);
    }
    return ok;
}

// Inter-tile decoupled look-back for one group (the whole warp): the group state at the start of
// `tile`.  T->depth is the number of predecessor tiles whose transition power Ac^(T*k) is still
// non-zero in f32: tiles further back contribute exactly nothing, so for fast-decaying groups
// (the LR4 biquads) the look-back reads a few aggregates -- published one whole tile earlier,
// because the section phase of a tile runs one loop iteration behind its front -- and never waits
// for an inclusive value; slow poles (the DC blocker) chain through one inclusive value per block of tiles.
template <int D>
DEV void lookback(const ChainArgs &a, const Hot *hot, int g, long long tile, int lane, float (&sin)[D]) {
    const GroupParam &G = hot->grp[g];
    const GroupTables *T = a.gtabs + g;
#pragma unroll
    for (int d = 0; d < D; ++d) sin[d] = 0.f;
    const int depth = __ldg(&T->depth);
    const bool agg_only = G.agg_only != 0;
    long long base = tile - 1;
    int dist0 = 0;                                   // predecessor distance of lane 0 in this window
    int window = 0;
    float Mw[D * D];                                 // Ac^(T * 32 * window), only used past the first window
    // Slow poles (no truncation): a FIXED recipe, so that the sum has the same shape in every run.  Tiles are dealt
    // round-robin in blocks of gridDim.x; tile t takes the aggregates of its p = t mod gridDim.x predecessors of the
    // same block and the INCLUSIVE value of the last tile of the previous block (or the state carried into the call),
    // which was published about one block-time earlier.
    const int p_blk = (int)(tile % (long long)gridDim.x);
    for (;;) {
        const long long idx = base - lane;
        const int delta = dist0 + lane;                   // predecessor distance, 0-based
        const bool beyond = agg_only ? (delta >= depth)   // weight is exactly zero from here on
                                     : (delta > p_blk);   // past the designated inclusive predecessor
        const TileLink *lk = a.links + (idx >= 0 ? idx : 0) * kMaxGroups + g;
        float pa[D], pi[D];
#pragma unroll
        for (int d = 0; d < D; ++d) { pa[d] = 0.f; pi[d] = 0.f; }
        int first_incl = 32;
        int spins = 0;
        for (;;) {
            bool ready, incl;
            if (beyond || idx < 0) {       // virtual terminators: zero weight / the state carried in
                ready = true;
                incl = true;
            } else if (agg_only || delta < p_blk) {
                incl = false;
                ready = read_link<D>(lk->agg, a.epoch, pa);
            } else {                       // delta == p_blk: the designated inclusive predecessor
                incl = read_link<D>(lk->incl, a.epoch, pi);
                ready = incl;
            }
            const unsigned incl_mask = __ballot_sync(FULLMASK, incl);
            const unsigned ready_mask = __ballot_sync(FULLMASK, ready);
            first_incl = incl_mask ? (__ffs(incl_mask) - 1) : 32;
            const unsigned need = (first_incl >= 31) ? FULLMASK : ((2u << first_incl) - 1u);
            if ((ready_mask & need) == need) break;
            if (++spins > (1 << 21)) {     // watchdog: never hang the device
                if (lane == 0) atomicExch(a.err_flag, 1);
                return;
            }
            __nanosleep(32);
        }
        float term[D];
#pragma unroll
        for (int d = 0; d < D; ++d) term[d] = 0.f;
        if (lane <= first_incl && !beyond) {
            float pay[D];
            if (idx < 0) {
#pragma unroll
                for (int d = 0; d < D; d += 2) {
                    const float2 c = (idx == -1) ? __ldcg(&a.carry_in->sec[G.first + d / 2]) : make_float2(0.f, 0.f);
                    pay[d] = c.x; pay[d + 1] = c.y;
                }
            } else {
#pragma unroll
                for (int d = 0; d < D; ++d) pay[d] = (lane == first_incl) ? pi[d] : pa[d];
            }
            float lbk[D * D], t1[D];
            load_mat<D>(T->lb[lane], lbk);
            matvec<D>(lbk, pay, t1);
            if (window == 0) {
#pragma unroll
                for (int d = 0; d < D; ++d) term[d] = t1[d];
            } else {
                matvec<D>(Mw, t1, term);
            }
        }
#pragma unroll
        for (int d = 0; d < D; ++d) {
            float v = term[d];
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(FULLMASK, v, o);
            sin[d] += v;
        }
        if (first_incl < 32) break;
        // next window of 32 predecessors: Mw <- Mw * Ac^(32 T)
        float L32[D * D];
        load_mat<D>(T->lb32, L32);
        if (window == 0) {
#pragma unroll
            for (int i = 0; i < D * D; ++i) Mw[i] = L32[i];
        } else {
            float Mn[D * D];
#pragma unroll
            for (int r = 0; r < D; ++r)
#pragma unroll
                for (int c = 0; c < D; ++c) {
                    float acc = 0.f;
#pragma unroll
                    for (int k = 0; k < D; ++k) acc = fmaf(Mw[r * D + k], L32[k * D + c], acc);
                    Mn[r * D + c] = acc;
                }
#pragma unroll
            for (int i = 0; i < D * D; ++i) Mw[i] = Mn[i];
        }
        ++window;
        base -= 32;
        dist0 += 32;
    }
}

// Slow poles (the DC blocker: no truncation) -- two-level look-back with a FIXED recipe, so results repeat to the bit.
// Tiles are grouped in blocks of 32 and blocks in superblocks of 32 (1024 tiles).  Three kinds of records:
//   * the zero-state aggregate of every tile (lk[t].agg, published by the tile's front),
//   * B_b, the zero-state aggregate of block b = tiles 32b .. 32b+31: a function of that block's tile aggregates alone,
//     formed by the block's last tile right after it has published its own aggregate (slow_block_publish; lk[32b+31].incl),
//   * SS_s, the TRUE state at the end of superblock s, chained SS_s = Ac^(1024T) SS_(s-1) + sum_i Ac^(32T(31-i)) B_(32s+i)
//     by the superblock's last tile at the same point (lk[1024s+1022].incl -- a slot no block uses).
// The state at the start of tile t (block b, position j; b = 32s + jb) is then
//   Ac^(Tj) [ Ac^(32T jb) SS_(s-1) + sum_(i<jb) Ac^(32T(jb-1-i)) B_(32s+i) ]  +  sum_(i<j) Ac^(T(j-1-i)) agg_(32b+i):
// two windows and one record, all loads in flight together, and the only serial chain runs over superblocks
// (18 steps for the 38.4 M-sample AM configuration, where the earlier per-grid-block chain had 127 links of five windows).
template <int D>
DEV void lookback_slow(const ChainArgs &a, const Hot *hot, int g, long long tile, int lane, const float (&agg_own)[D], float (&sin)[D]) {
    const GroupParam &G = hot->grp[g];
    const GroupTables *T = a.gtabs + g;
    const int j = (int)(tile & 31);
    const long long b = tile >> 5;
    const int jb = (int)(b & 31);
    const long long sb = b >> 5;
    // ---- tile level: lane l < j reads the aggregate of tile t-1-l
    float pa[D];
#pragma unroll
    for (int d = 0; d < D; ++d) pa[d] = 0.f;
    const bool want1 = lane < j;
    {
        const TileLink *lk = a.links + (want1 ? (tile - 1 - lane) : tile) * kMaxGroups + g;
        int spins = 0;
        for (;;) {
            const bool ready = !want1 || read_link<D>(lk->agg, a.epoch, pa);
            if (__all_sync(FULLMASK, ready)) break;
            if (++spins > (1 << 21)) { if (lane == 0) atomicExch(a.err_flag, 9); break; }   // watchdog: never hang the device
            __nanosleep(32);
        }
    }
    float term1[D];
    {
        float t1[D];
#pragma unroll
        for (int d = 0; d < D; ++d) t1[d] = 0.f;
        if (want1) {
            float m[D * D];
            load_mat<D>(T->lb[lane], m);
            matvec<D>(m, pa, t1);
        }
#pragma unroll
        for (int d = 0; d < D; ++d) {
            float v = t1[d];
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(FULLMASK, v, o);
            term1[d] = v;
        }
    }
    // ---- block level: lane l < jb reads B of block b-1-l; lane 31 reads SS of the previous superblock (or the carried state)
    float pb[D];
#pragma unroll
    for (int d = 0; d < D; ++d) pb[d] = 0.f;
    const bool want2 = lane < jb;
    const bool want_ss = lane == 31;
    {
        const long long rec_tile = want2 ? ((b - 1 - lane) * 32 + 31) : (want_ss && sb > 0 ? (sb - 1) * 1024 + 1022 : tile);
        const TileLink *lk = a.links + rec_tile * kMaxGroups + g;
        int spins = 0;
        for (;;) {
            bool ready = true;
            if (want2 || (want_ss && sb > 0)) ready = read_link<D>(lk->incl, a.epoch, pb);
            if (__all_sync(FULLMASK, ready)) break;
            if (++spins > (1 << 21)) { if (lane == 0) atomicExch(a.err_flag, 10); break; }
            __nanosleep(32);
        }
        if (want_ss && sb == 0) {                    // the state carried into this call sits before superblock 0
#pragma unroll
            for (int d = 0; d < D; d += 2) {
                const float2 c = __ldcg(&a.carry_in->sec[G.first + d / 2]);
                pb[d] = c.x; pb[d + 1] = c.y;
            }
        }
    }
    float base[D];                                   // state at the start of block b
    {
        float t2[D];
#pragma unroll
        for (int d = 0; d < D; ++d) t2[d] = 0.f;
        if (want2 || want_ss) {
            float m[D * D];
            load_mat<D>(T->lbb[want_ss ? jb : lane], m);
            matvec<D>(m, pb, t2);
        }
#pragma unroll
        for (int d = 0; d < D; ++d) {
            float v = t2[d];
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(FULLMASK, v, o);
            base[d] = v;
        }
    }
    {
        float m[D * D], t3[D];
        load_mat<D>(T->lb[j], m);
        matvec<D>(m, base, t3);
#pragma unroll
        for (int d = 0; d < D; ++d) sin[d] = t3[d] + term1[d];
    }
    (void)agg_own;
}

// The block and superblock records of a slow-pole group are formed by the block's LAST tile right after it has published
// its own aggregate -- in the same pipeline stage, i.e. a whole loop iteration before the tiles of the next block look for
// them.  (Forming them in the finish stage, as a by-product of that tile's own look-back, made 32 tiles wait for one
// warp that was running concurrently with them.)  Only this one warp per 32 tiles waits for concurrent publishes here.
template <int D>
DEV void slow_block_publish(const ChainArgs &a, const Hot *hot, int g, long long tile, int lane, const float (&agg_own)[D]) {
    const GroupParam &G = hot->grp[g];
    const GroupTables *T = a.gtabs + g;
    const long long b = tile >> 5;
    const int jb = (int)(b & 31);
    const long long sb = b >> 5;
    float pa[D];
#pragma unroll
    for (int d = 0; d < D; ++d) pa[d] = 0.f;
    const bool want1 = lane < 31;
    {
        const TileLink *lk = a.links + (want1 ? (tile - 1 - lane) : tile) * kMaxGroups + g;
        int spins = 0;
        for (;;) {
            const bool ready = !want1 || read_link<D>(lk->agg, a.epoch, pa);
            if (__all_sync(FULLMASK, ready)) break;
            if (++spins > (1 << 21)) { if (lane == 0) atomicExch(a.err_flag, 11); break; }   // watchdog: never hang the device
            __nanosleep(64);
        }
    }
    float Bb[D];
    {
        float t1[D], term1[D];
#pragma unroll
        for (int d = 0; d < D; ++d) t1[d] = 0.f;
        if (want1) {
            float m[D * D];
            load_mat<D>(T->lb[lane], m);
            matvec<D>(m, pa, t1);
        }
#pragma unroll
        for (int d = 0; d < D; ++d) {
            float v = t1[d];
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(FULLMASK, v, o);
            term1[d] = v;
        }
        float m[D * D];
        load_mat<D>(T->tile, m);
        matvec<D>(m, term1, Bb);                     // B_b = Ac^T (sum over the 31 tiles before this one) + own aggregate
#pragma unroll
        for (int d = 0; d < D; ++d) Bb[d] += agg_own[d];
    }
    publish<D>(lane == 0, a, tile, g, Bb, true);
    if (jb != 31) return;
    // superblock-last tile: SS_s = Ac^(32T) [ Ac^(32T 31) SS_(s-1) + sum_(i<31) Ac^(32T(30-i)) B_(32s+i) ] + B_b
    float pb[D];
#pragma unroll
    for (int d = 0; d < D; ++d) pb[d] = 0.f;
    const bool want2 = lane < 31, want_ss = lane == 31;
    {
        const long long rec_tile = want2 ? ((b - 1 - lane) * 32 + 31) : (sb > 0 ? (sb - 1) * 1024 + 1022 : tile);
        const TileLink *lk = a.links + rec_tile * kMaxGroups + g;
        int spins = 0;
        for (;;) {
            bool ready = true;
            if (want2 || (want_ss && sb > 0)) ready = read_link<D>(lk->incl, a.epoch, pb);
            if (__all_sync(FULLMASK, ready)) break;
            if (++spins > (1 << 21)) { if (lane == 0) atomicExch(a.err_flag, 12); break; }
            __nanosleep(64);
        }
        if (want_ss && sb == 0) {
#pragma unroll
            for (int d = 0; d < D; d += 2) {
                const float2 c = __ldcg(&a.carry_in->sec[G.first + d / 2]);
                pb[d] = c.x; pb[d + 1] = c.y;
            }
        }
    }
    float base[D], t2[D];
    {
        float m[D * D];
        load_mat<D>(T->lbb[want_ss ? 31 : lane], m);
        matvec<D>(m, pb, t2);
    }
#pragma unroll
    for (int d = 0; d < D; ++d) {
        float v = t2[d];
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(FULLMASK, v, o);
        base[d] = v;
    }
    float m[D * D], ss[D];
    load_mat<D>(T->lbb[1], m);
    matvec<D>(m, base, ss);
#pragma unroll
    for (int d = 0; d < D; ++d) ss[d] += Bb[d];
    publish<D>(lane == 0, a, sb * 1024 + 1022, g, ss, true);
}

// one recursive-section step, reference arithmetic
template <int TYPE>
DEV float sec_step_t(const SecParam &P, float x, float &s0, float &s1) {
    float y;
    if (TYPE == SEC_BIQUAD) {                 // iir.rs:34-40
        y = fmaf(x, P.c[0], s0);
        s0 = fmaf(x, P.c[1], s1) - P.c[3] * y;
        s1 = x * P.c[2] - P.c[4] * y;
    } else if (TYPE == SEC_DC) {              // iir.rs:160-163 / dc.rs:49-52: y = x - x1 + r*y1
        y = (x - s0) + P.c[0] * s1;
        s0 = x;
        s1 = y;
    } else {                                  // cw.rs:38: y = a*y + (1-a)*x
        y = P.c[0] * s0 + P.c[1] * x;
        s0 = y;
    }
    return y;
}

// pass 2 of one section over this lane's items (section by section == sample by sample for a
// cascade); writes the carried state when this lane owns the last item of the call
template <int TYPE, int NPT, bool NOPOST = false>
DEV void sec_pass2(const ChainArgs &a, const SecParam &P, int s, float (&u)[NPT], float &s0, float &s1,
                   long long jt, bool full) {
    if (full) {
#pragma unroll
        for (int i = 0; i < NPT; ++i) {
            const float y = sec_step_t<TYPE>(P, u[i], s0, s1);
            u[i] = (NOPOST || P.post_op == OP_NONE) ? y : post_apply(P, y);
        }
        if (jt + NPT == a.n_out) a.carry_out->sec[s] = make_float2(s0, s1);
    } else {
#pragma unroll
        for (int i = 0; i < NPT; ++i) {
            if (jt + i < a.n_out) {
                const float y = sec_step_t<TYPE>(P, u[i], s0, s1);
                u[i] = post_apply(P, y);
                if (jt + i == a.n_out - 1) a.carry_out->sec[s] = make_float2(s0, s1);
            }
        }
    }
}

template <int D, bool TRI = false>
DEV void group_scan(const ChainArgs &a, const Hot *hot, int g, long long tile, int lane, float (&E)[D], float (&X)[D], float (&agg)[D]);
// group front: zero-state end state of this lane's chunk (dot products with the impulse
// responses), warp scan with constant transition powers, publish the tile aggregate.
// X = state contribution of the lanes before this one; agg = whole-tile aggregate.
template <int D, int NPT>
DEV void group_front(const ChainArgs &a, const Hot *hot, int g, long long tile, int lane, const float (&u)[NPT], bool full,
                     float (&X)[D], float (&agg)[D]) {
    const GroupParam &G = hot->grp[g];
    __syncwarp();
    float E[D];
#pragma unroll
    for (int d = 0; d < D; ++d) E[d] = 0.f;
    if (full) {
#pragma unroll
        for (int i = 0; i < NPT; ++i)
#pragma unroll
            for (int d = 0; d < D; ++d) E[d] = fmaf(G.imp[i][d], u[i], E[d]);
    }
    group_scan<D>(a, hot, g, tile, lane, E, X, agg);
    if (!G.agg_only && (tile & 31) == 31) slow_block_publish<D>(a, hot, g, tile, lane, agg);
}
// warp scan of the lanes' zero-state end states E with constant transition powers; publishes the tile aggregate
template <int D, bool TRI>
DEV void group_scan(const ChainArgs &a, const Hot *hot, int g, long long tile, int lane, float (&E)[D], float (&X)[D], float (&agg)[D]) {
    const GroupParam &G = hot->grp[g];
    if ((ORION_TRACE && a.trace) && g == 0) {
        const unsigned am = __activemask();
        if (lane == 0) { a.trace[tile * 16 + 11] = clock64(); }
        if (lane == 31) a.trace[tile * 16 + 15] = am;
    }
    __syncwarp();                                  // converged here: the shuffles below take the fast path
#pragma unroll
    for (int l = 0; l < 5; ++l) {
        // levels whose transition power Ac^(n 2^l) is below 2^-40 in every entry cannot change an f32 state sum: skipped
        // (warp-uniform; for the C1 LR4 that is the last level)
        if (l >= G.scan_levels) break;
        float o[D], t[D];
#pragma unroll
        for (int d = 0; d < D; ++d) o[d] = __shfl_up_sync(FULLMASK, E[d], 1 << l);
        if constexpr (TRI && D == 4) matvec4_tri(G.lv[l], o, t);
        else matvec<D>(G.lv[l], o, t);
        const bool take = lane >= (1 << l);        // select, not a branch: the warp stays converged
#pragma unroll
        for (int d = 0; d < D; ++d) E[d] += take ? t[d] : 0.f;
    }
#pragma unroll
    for (int d = 0; d < D; ++d) {
        X[d] = __shfl_up_sync(FULLMASK, E[d], 1);
        if (lane == 0) X[d] = 0.f;
        agg[d] = __shfl_sync(FULLMASK, E[d], 31);
    }
    if ((ORION_TRACE && a.trace) && lane == 0 && g == 0) a.trace[tile * 16 + 12] = clock64();
    publish<D>(lane == 0, a, tile, g, agg, false);
    if ((ORION_TRACE && a.trace) && lane == 0 && g == 0) a.trace[tile * 16 + 13] = clock64();
}

// group finish: look-back, true start state of this lane's chunk, the reference recursion
template <int D, int NPT, bool ALLBIQ = false>
DEV void group_finish(const ChainArgs &a, const Hot *hot, int g, long long tile, int lane, float (&u)[NPT], bool full, long long jt,
                      const float (&X)[D], const float (&agg)[D]) {
    const GroupParam &G = hot->grp[g];
    const GroupTables *T = a.gtabs + g;
    float sin[D];
    if (G.agg_only) lookback<D>(a, hot, g, tile, lane, sin);
    else lookback_slow<D>(a, hot, g, tile, lane, agg, sin);        // slow poles: two-level fixed recipe (publishes B_b / SS_s)
    if ((ORION_TRACE && a.trace) && lane == 0 && g == 0) a.trace[tile * 16 + 10] = clock64();
    float lm[D * D], st[D];
    load_mat<D>(T->lane[lane], lm);
    matvec<D>(lm, sin, st);
#pragma unroll
    for (int d = 0; d < D; ++d) st[d] += X[d];
#pragma unroll
    for (int q = 0; q < D / 2; ++q) {
        const int s = G.first + q;
        const SecParam &P = hot->sec[s];
        if (ALLBIQ || P.type == SEC_BIQUAD) sec_pass2<SEC_BIQUAD, NPT, ALLBIQ>(a, P, s, u, st[2 * q], st[2 * q + 1], jt, full);
        else if (P.type == SEC_DC) sec_pass2<SEC_DC, NPT>(a, P, s, u, st[2 * q], st[2 * q + 1], jt, full);
        else sec_pass2<SEC_ONEPOLE, NPT>(a, P, s, u, st[2 * q], st[2 * q + 1], jt, full);
    }
}

// dispatch on the group dimension
template <int NPT>
DEV void group_front_park(const ChainArgs &a, const Hot *hot, int g, long long tile, int lane, const float (&u)[NPT], bool full,
                          float *park) {
    // park layout (floats): [lane][8] X, then [8] agg
#define ORION_GF(DD) { float X[DD], agg[DD]; group_front<DD, NPT>(a, hot, g, tile, lane, u, full, X, agg); \
        _Pragma("unroll") for (int d = 0; d < DD; ++d) park[lane * kMaxGroupDim + d] = X[d]; \
        if (lane == 0) { _Pragma("unroll") for (int d = 0; d < DD; ++d) park[32 * kMaxGroupDim + d] = agg[d]; } }
    switch (hot->grp[g].D) {
        case 2: ORION_GF(2) break;
        default: ORION_GF(4) break;
    }
#undef ORION_GF
    __syncwarp();
}
template <int NPT>
DEV void group_finish_parked(const ChainArgs &a, const Hot *hot, int g, long long tile, int lane, float (&u)[NPT], bool full,
                             long long jt, const float *park) {
#define ORION_GP(DD) { float X[DD], agg[DD]; \
        _Pragma("unroll") for (int d = 0; d < DD; ++d) { X[d] = park[lane * kMaxGroupDim + d]; agg[d] = park[32 * kMaxGroupDim + d]; } \
        group_finish<DD, NPT>(a, hot, g, tile, lane, u, full, jt, X, agg); }
    switch (hot->grp[g].D) {
        case 2: ORION_GP(2) break;
        default: ORION_GP(4) break;
    }
#undef ORION_GP
}
template <int NPT>
DEV void group_whole(const ChainArgs &a, const Hot *hot, int g, long long tile, int lane, float (&u)[NPT], bool full, long long jt) {
#define ORION_GW(DD) { float X[DD], agg[DD]; group_front<DD, NPT>(a, hot, g, tile, lane, u, full, X, agg); \
        group_finish<DD, NPT>(a, hot, g, tile, lane, u, full, jt, X, agg); }
    switch (hot->grp[g].D) {
        case 2: ORION_GW(2) break;
        default: ORION_GW(4) break;
    }
#undef ORION_GW
}

// ----------------------------------------------------------------------------------------------
// staged-tile geometry (FRONT_STAGED)
//   global row G covers samples [row_samples*G + O - Mb + 2, +row_samples)   (call-relative)
//   tile t stages rows G0 .. G0+rows-1 with G0 = t*32 - HR, rows = 32 + HR; lane l reads rows
//   l .. l+HR.  row_pitch is an odd multiple of 16 bytes, so the 8 lanes of a quarter-warp hit 8
//   different 16-byte bank groups on every LDS.128.
// ----------------------------------------------------------------------------------------------
DEV long long row_start_sample(const ChainArgs &a, long long G) {
    return (long long)a.row_samples * G + (a.O - a.Mb + 2);
}
DEV bool tile_is_interior(const ChainArgs &a, long long tile) {
    // the host's form of: use_tma && G0 >= tma_row0 && G0 + 32 + HR <= tma_row0 + tma_rows, G0 = 32 tile - HR
    return tile >= a.tile_int_lo && tile <= a.tile_int_hi;
}

// cooperative (whole warp) load of an edge tile -- FIR history / ragged tail -- into the staged layout.
// Loads are issued in batches of 16 independent chunks per lane: an edge tile sits in the same ring
// as the TMA-staged ones, so a latency-serialised loader would stall the whole CTA behind its slot.
static __device__ __noinline__ void stage_load_generic(const ChainArgs &a, long long tile, unsigned char *smem, int lane) {
    constexpr int B = 16;
    const int rows = kThreads + a.HR;
    const long long G0 = tile * kThreads - a.HR;
    const int cpr = a.row_samples >> 1;                 // 16-byte chunks per row
    const int total = rows * cpr;
    const long long s_base = row_start_sample(a, G0);   // rows are contiguous in the stream: chunk c starts at s_base + 2c
    const bool al16 = ((reinterpret_cast<uintptr_t>(a.in) & 15u) == 0);
    const float2 *in = reinterpret_cast<const float2 *>(a.in);
    for (int cb = 0; cb < total; cb += kThreads * B) {      // uniform trip count
        const int c0 = cb + lane;
        float4 v[B];
#pragma unroll
        for (int k = 0; k < B; ++k) {
            const int c = c0 + k * kThreads;
            const long long s = s_base + 2 * (long long)c;
            v[k] = make_float4(0.f, 0.f, 0.f, 0.f);
            if (c < total) {
                if (al16 && s >= 0 && s + 1 < a.n_in) {
                    v[k] = __ldg(reinterpret_cast<const float4 *>(in + s));
                } else {
                    const float2 x0 = load_x(a, s), x1 = load_x(a, s + 1);
                    v[k] = make_float4(x0.x, x0.y, x1.x, x1.y);
                }
            }
        }
#pragma unroll
        for (int k = 0; k < B; ++k) {
            const int c = c0 + k * kThreads;
            if (c < total) {
                const int rho = c / cpr;
                const int cc = c - rho * cpr;
                *reinterpret_cast<float4 *>(smem + (size_t)rho * a.row_pitch + (size_t)cc * 16) = v[k];
            }
        }
    }
    __syncwarp();
}

// in-place input-rate mixer on the staged samples: x[s] * p(kbase + s + 1).  A lane walks one row (the phasor is a
// recurrence along the samples); the rows left over after the last full pass of 32 (the HR halo rows) are cut into
// segments so that they occupy all lanes for a fraction of a pass instead of a few lanes for a whole one.
static __device__ __noinline__ void stage_mix(const ChainArgs &a, long long tile, unsigned char *smem, int lane) {
    const int rows = kThreads + a.HR;
    const long long G0 = tile * kThreads - a.HR;
    const int cpr = a.row_samples >> 1;
    const float2 w = make_float2(a.pre.wre, a.pre.wim);
    for (int rb = 0; rb < rows; rb += kThreads) {           // uniform trip count
        const int nrem = min(kThreads, rows - rb);
        const int S = kThreads / nrem;                      // segments per row in this pass (1 while more than 16 rows are left)
        const int rho = rb + lane / S;
        const int len = (cpr + S - 1) / S;
        const int cc0 = (lane % S) * len, cc1 = min(cpr, cc0 + len);
        if (lane >= nrem * S || cc0 >= cc1) continue;
        const long long s0 = row_start_sample(a, G0 + rho);
        unsigned char *rp = smem + (size_t)rho * a.row_pitch;
        float2 p = make_float2(1.f, 0.f);
        if (a.pre.exact) {                                     // the reference recurrence itself, replayed from a checkpoint
            const float2 wx = make_float2(a.pre.xwre, a.pre.xwim);
            unsigned ctr = 0;
            bool walking = false;
            for (int cc = cc0; cc < cc1; ++cc) {
                float4 v = *reinterpret_cast<float4 *>(rp + cc * 16);
                float2 y[2];
#pragma unroll
                for (int h = 0; h < 2; ++h) {
                    const long long s = s0 + 2 * cc + h;
                    if (s < 0) p = nco_exact_hist(a.pre, s);
                    else if (!walking) { p = nco_exact_at(a.pre, s, ctr, (int)(s & 15)); walking = true; }
                    else nco_step_exact(p, wx, ctr);
                    y[h] = mix_apply(a.mix, h == 0 ? make_float2(v.x, v.y) : make_float2(v.z, v.w), p);
                }
                *reinterpret_cast<float4 *>(rp + cc * 16) = make_float4(y[0].x, y[0].y, y[1].x, y[1].y);
            }
            continue;
        }
        for (int cc = cc0; cc < cc1; ++cc) {
            const unsigned long long k = a.pre.kbase + (unsigned long long)(s0 + 2 * cc) + 1ull;
            if (cc == cc0 || (cc & 7) == 0) p = nco_unit(a.pre, k);
            float4 v = *reinterpret_cast<float4 *>(rp + cc * 16);
            const float2 y0 = mix_apply(a.mix, make_float2(v.x, v.y), scale2(p, nco_amp(a.pre, k)));
            p = cmul_fma(p, w);
            const float2 y1 = mix_apply(a.mix, make_float2(v.z, v.w), scale2(p, nco_amp(a.pre, k + 1ull)));
            p = cmul_fma(p, w);
            *reinterpret_cast<float4 *>(rp + cc * 16) = make_float4(y0.x, y0.y, y1.x, y1.y);
        }
    }
    __syncwarp();
}

// shared-memory address of call-relative sample s inside the staged tile
DEV const float2 *staged_sample(const ChainArgs &a, const unsigned char *smem, long long G0, long long s) {
    const int d = (int)(s - row_start_sample(a, G0));           // tile-local: fits 32 bits
    const int rho = (a.row_shift >= 0) ? (d >> a.row_shift) : (d / a.row_samples);
    const int w = d - rho * a.row_samples;
    return reinterpret_cast<const float2 *>(smem + (size_t)rho * a.row_pitch + (size_t)w * 8);
}

// polyphase FIR over the staged tile: lane l produces the outputs of blocks l*R .. l*R+R-1.
// Per tap step: one LDS.128 (two adjacent samples), one tap pair (broadcast LDS.64),
// 2*R*U packed FP32x2 FMAs (FFMA2: re and im of one output in one instruction) on an R-deep
// sliding register window.
// Build-time variants (measured on the B200; DESIGN.md "tuning log"):
//   ORION_FIR_PACKED  1: FFMA2 (packed f32x2) inner loop, loads hoisted per row; 0: scalar FFMA, rotating window
//   ORION_TAPS_SMEM   1: tap pairs from a shared-memory copy; 0: from the kernel parameter bank
//   ORION_HOT_SMEM    1: section / group launch data from a shared-memory copy; 0: from the parameter bank
//   ORION_TRACE       1: per-tile SM-clock stamps into ChainArgs::trace (debug builds only: the stamps sit in the hot path)
#ifndef ORION_FIR_PACKED
#define ORION_FIR_PACKED 1
#endif
#ifndef ORION_FIR_Q_UNROLL
#define ORION_FIR_Q_UNROLL 1  // unroll factor of the FIR's loop over sample pairs (4 = full for the decimate-by-8 shape)
#endif
constexpr int kFirQUnroll = ORION_FIR_Q_UNROLL;
#ifndef ORION_EARLY_FINISH
#define ORION_EARLY_FINISH 0  // 1: a warp whose stage slot is not filled yet runs its pending tile's section phase first
                              // (measured: C1 46.1 -> 59.1 us -- the look-back then waits instead; profiles/r02_experiments.txt)
#endif
#ifndef ORION_FM_ROLLED
#define ORION_FM_ROLLED 0     // 1: the FM front of a lane is one rolled loop (small code); 0: unrolled over the lane's items
#endif
#ifndef ORION_TAPS_SMEM
#define ORION_TAPS_SMEM 0
#endif
#ifndef ORION_HOT_SMEM
#define ORION_HOT_SMEM 0
#endif
#if ORION_TAPS_SMEM
#define ORION_TAPS taps_sh
#else
#define ORION_TAPS a.taps2
#endif

#if ORION_FIR_PACKED
// rr0 .. rr1: the range of halo rows (tap steps rr*R .. rr*R + R - 1 of every pair) this call accumulates -- the whole
// filter by default; a tile whose FIR is split over several warps (ChainArgs::split) gives each of them a slice.
template <int R, int U, int SP, int QU>
DEV void fir_staged(const ChainArgs &a, const unsigned char *smem, const float2 *taps_sh, int lane, float2 (&z)[R * U],
                    int rr0 = 0, int rr1 = -1) {
    typedef Geo<SP> GE;
    f32x2 acc[U][R];                                 // (re, im) of every output, one packed register pair each
#pragma unroll
    for (int u = 0; u < U; ++u)
#pragma unroll
        for (int i = 0; i < R; ++i) acc[u][i] = pack2(0.f, 0.f);

    const int Mb = GE::fixed ? GE::Mb : a.Mb, pitch = GE::fixed ? GE::pitch : a.row_pitch;
    const int P_pad = GE::fixed ? GE::P_pad : a.P_pad, HR = GE::fixed ? GE::HR : a.HR;
    const int blk_bytes = Mb * 8;
    if (rr1 < 0) rr1 = HR;
    const unsigned char *row_own = smem + (size_t)(lane + rr0) * pitch;
    const int npairs = Mb >> 1;
#pragma unroll QU
    for (int q = 0; q < npairs; ++q) {
        const int off_q = (Mb - 2 - 2 * q) * 8;
        // sliding window of sample pairs (x[s], x[s+1]), each sample a packed (re, im) pair:
        // W[0 .. R-2] = carried in, W[R-1 .. 2R-2] = the R blocks of the next row
        f32x2 wlo[2 * R - 1], whi[2 * R - 1];
#pragma unroll
        for (int k = 0; k + 1 < R; ++k) {
            const float4 v = *reinterpret_cast<const float4 *>(row_own + (k + 1) * blk_bytes + off_q);
            wlo[k] = pack2(v.x, v.y);
            whi[k] = pack2(v.z, v.w);
        }
        const float2 *tp0 = ORION_TAPS + (size_t)q * P_pad;
        const float2 *tp1 = ORION_TAPS + (size_t)(npairs + q) * P_pad;     // U == 2 only
        const unsigned char *rb = row_own + pitch + off_q;
#pragma unroll (GE::fixed ? 2 : 1)
        for (int rr = rr0; rr < rr1; ++rr, rb += pitch) {
            // all loads of this row first (R LDS.128 + R*U broadcast LDS.64), then 2*R*R*U packed FMAs
            // in an order that touches every accumulator once per tap: dependent FMAs are R apart
            float2 t[U][R];
#pragma unroll
            for (int kk = 0; kk < R; ++kk) {
                const float4 v = *reinterpret_cast<const float4 *>(rb + kk * blk_bytes);
                wlo[R - 1 + kk] = pack2(v.x, v.y);
                whi[R - 1 + kk] = pack2(v.z, v.w);
                t[0][kk] = tp0[rr * R + kk];
                if (U == 2) t[U - 1][kk] = tp1[rr * R + kk];
            }
#pragma unroll
            for (int kk = 0; kk < R; ++kk) {
#pragma unroll
                for (int u = 0; u < U; ++u) {
                    const f32x2 t0 = pack2(t[u][kk].x, t[u][kk].x), t1 = pack2(t[u][kk].y, t[u][kk].y);
#pragma unroll
                    for (int i = 0; i < R; ++i) acc[u][i] = ffma2(t0, wlo[kk + i], acc[u][i]);
#pragma unroll
                    for (int i = 0; i < R; ++i) acc[u][i] = ffma2(t1, whi[kk + i], acc[u][i]);
                }
            }
#pragma unroll
            for (int k = 0; k + 1 < R; ++k) { wlo[k] = wlo[k + R]; whi[k] = whi[k + R]; }
        }
    }
#pragma unroll
    for (int i = 0; i < R; ++i)
#pragma unroll
        for (int u = 0; u < U; ++u) z[i * U + u] = unpack2(acc[u][i]);
}

#else
template <int R, int U, int SP, int QU>
DEV void fir_staged(const ChainArgs &a, const unsigned char *smem, const float2 *taps_sh, int lane, float2 (&z)[R * U]) {
    float2 acc[U][R];
#pragma unroll
    for (int u = 0; u < U; ++u)
#pragma unroll
        for (int i = 0; i < R; ++i) acc[u][i] = make_float2(0.f, 0.f);

    const int Mb = a.Mb, pitch = a.row_pitch, P_pad = a.P_pad, HR = a.HR;
    const int blk_bytes = Mb * 8;
    const unsigned char *row_own = smem + (size_t)lane * pitch;
    const int npairs = Mb >> 1;
    for (int q = 0; q < npairs; ++q) {
        const int off_q = (Mb - 2 - 2 * q) * 8;
        float4 w[R];
#pragma unroll
        for (int k = 0; k + 1 < R; ++k)
            w[k] = *reinterpret_cast<const float4 *>(row_own + (k + 1) * blk_bytes + off_q);
        const float2 *tp0 = ORION_TAPS + (size_t)q * P_pad;
        const float2 *tp1 = ORION_TAPS + (size_t)(npairs + q) * P_pad;     // U == 2 only
        const unsigned char *rb = row_own + pitch + off_q;
        for (int rr = 0; rr < HR; ++rr, rb += pitch) {
#pragma unroll
            for (int kk = 0; kk < R; ++kk) {
                w[(kk + R - 1) % R] = *reinterpret_cast<const float4 *>(rb + kk * blk_bytes);
                const int c = rr * R + kk;
#pragma unroll
                for (int u = 0; u < U; ++u) {
                    const float2 t = (u == 0) ? tp0[c] : tp1[c];
#pragma unroll
                    for (int i = 0; i < R; ++i) {
                        const float4 wv = w[(kk + i) % R];
                        acc[u][i].x = fmaf(t.x, wv.x, acc[u][i].x);
                        acc[u][i].y = fmaf(t.x, wv.y, acc[u][i].y);
                        acc[u][i].x = fmaf(t.y, wv.z, acc[u][i].x);
                        acc[u][i].y = fmaf(t.y, wv.w, acc[u][i].y);
                    }
                }
            }
        }
    }
#pragma unroll
    for (int i = 0; i < R; ++i)
#pragma unroll
        for (int u = 0; u < U; ++u) z[i * U + u] = acc[u][i];
}

#endif

// one FIR output evaluated straight from the virtual stream (any shape), in the reference's
// accumulation order and rounding (fir.rs:57-66 unfused; fir.rs:229-247 fused)
DEV float2 fir_global_one(const ChainArgs &a, long long j) {
    const long long n = (long long)a.M * j;
    float re = 0.f, im = 0.f;
    if (a.fir == FIR_DECIM) {
        // g[k] = taps[k-1] (k >= 1), g[0] = taps[L-1]; reference order: k = 1 .. L-1, then 0
        for (int k = 1; k < a.Lg; ++k) {
            const float2 x = load_x_mixed(a, n - k);
            const float t = __ldg(a.g + k);
            re = re + x.x * t;
            im = im + x.y * t;
        }
        const float2 x = load_x_mixed(a, n);
        const float t = __ldg(a.g);
        re = re + x.x * t;
        im = im + x.y * t;
    } else if (a.fir == FIR_IQ_UNFUSED) {
        for (int k = 0; k < a.Lg; ++k) {
            const float2 x = load_x_mixed(a, n - k);
            const float t = __ldg(a.g + k);
            re = re + x.x * t;
            im = im + x.y * t;
        }
    } else {
        for (int k = 0; k < a.Lg; ++k) {
            const float2 x = load_x_mixed(a, n - k);
            const float t = __ldg(a.g + k);
            re = fmaf(x.x, t, re);
            im = fmaf(x.y, t, im);
        }
    }
    return make_float2(re, im);
}

// warp-cooperative evaluation of one FIR output from the staged tile (discriminator halo)
DEV float2 fir_staged_one(const ChainArgs &a, const unsigned char *smem, const float *g_sh, long long G0, long long j, int lane) {
    const long long n = (long long)a.M * j;
    float re = 0.f, im = 0.f;
    // uniform trip count: a lane-dependent loop bound would leave the warp split, and every
    // warp-synchronous shuffle after it would then take the compiler's (very slow) divergent path
    for (int k0 = 0; k0 < a.Lg; k0 += 32) {
        const int k = k0 + lane;
        if (k < a.Lg) {
            const float2 x = *staged_sample(a, smem, G0, n - k);
            const float t = g_sh[k];
            re = fmaf(x.x, t, re);
            im = fmaf(x.y, t, im);
        }
    }
    __syncwarp();
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        re += __shfl_xor_sync(FULLMASK, re, o);
        im += __shfl_xor_sync(FULLMASK, im, o);
    }
    return make_float2(re, im);
}

// ----------------------------------------------------------------------------------------------
// after the front: demod-rate oscillator + demodulator front map  (z -> u), C32 store for
// DEMOD_NONE, end-of-call duties
// ----------------------------------------------------------------------------------------------
template <int NPT, int DM>
DEV void front_map(const ChainArgs &a, long long tile, int lane, float2 (&z)[NPT], float (&u)[NPT], float2 zhalo,
                   unsigned char *xs = nullptr) {
    const long long j0 = tile * (long long)(kThreads * NPT);
    const long long jt = j0 + (long long)lane * NPT;
    const int demod = Dm<DM>::demod(a);
    const bool is_c32_in = !kind_f32_in(demod);
    const bool need_prev = demod == DEMOD_FM || demod == DEMOD_PM;
    const bool post_osc = (demod == DEMOD_FM && a.translate) || demod == DEMOD_SSB || demod == DEMOD_USB ||
                          demod == MOD_AM || demod == MOD_PM;
    const bool full = jt + NPT <= a.n_out;

    if (is_c32_in && demod != DEMOD_NONE) {
        float2 p = make_float2(1.f, 0.f);
        const float2 w = make_float2(a.post.wre, a.post.wim);
        const unsigned long long kp0 = a.post.kbase + (unsigned long long)jt + 1ull;
        if (post_osc && !(a.post.exact && demod != DEMOD_FM)) p = nco_unit(a.post, kp0);
        if (demod == DEMOD_FM && a.translate) {
            // z = in * conj(p)   (num-complex Mul, unfused; fm.rs:49)
            // the unit phasor is enough: |p| = 1 + O(1e-5) (rotator.rs renormalises every 1024 steps) scales
            // both discriminator arguments alike, and atan2_approx only sees their ratio.
            // The phasor of the item before the lane's first one is one step back, p * conj(w): lane 0 needs it for the
            // item before the tile (every lane computes it -- four flops -- so that the warp stays converged).
            {
                const float2 pb = make_float2(fmaf(p.x, w.x, p.y * w.y), fmaf(p.y, w.x, -(p.x * w.y)));
                const float cr = pb.x, ci = -pb.y;
                zhalo = make_float2(zhalo.x * cr - zhalo.y * ci, zhalo.x * ci + zhalo.y * cr);
            }
#pragma unroll
            for (int i = 0; i < NPT; ++i) {
                const float cr = p.x, ci = -p.y;
                z[i] = make_float2(z[i].x * cr - z[i].y * ci, z[i].x * ci + z[i].y * cr);
                p = cmul_fma(p, w);
            }
        }
        if (need_prev) {
            float2 prev = shfl_up2(z[NPT - 1], 1);
            if (lane == 0) prev = (j0 > 0) ? zhalo : __ldcg(&a.carry_in->prev);
            __syncwarp();
            // carried discriminator state for the next call
            if (a.n_out > 0 && jt <= a.n_out - 1 && a.n_out - 1 < jt + NPT) {
                float2 last = z[0];
#pragma unroll
                for (int i = 1; i < NPT; ++i)
                    if (jt + i == a.n_out - 1) last = z[i];
                a.carry_out->prev = last;
            }
            if (demod == DEMOD_FM) {
#pragma unroll
                for (int i = 0; i < NPT; ++i) {             // fm.rs:50-56
                    const float pr = z[i].x * prev.x + z[i].y * prev.y;
                    const float pi = z[i].y * prev.x - z[i].x * prev.y;
                    u[i] = atan2_approx(pi, pr) * a.k;
                    prev = z[i];
                }
            } else {
#pragma unroll
                for (int i = 0; i < NPT; ++i) {             // pm.rs:54-58 (z * prev.conj())
                    const float cr = prev.x, ci = -prev.y;
                    const float wre = z[i].x * cr - z[i].y * ci;
                    const float wim = z[i].x * ci + z[i].y * cr;
                    u[i] = a.k * atan2_approx(wim, wre);
                    prev = z[i];
                }
            }
        } else if (demod == DEMOD_AM) {
#pragma unroll
            for (int i = 0; i < NPT; ++i) u[i] = fmaf(z[i].x, z[i].x, z[i].y * z[i].y);          // am.rs:53
        } else if (demod == DEMOD_AM_ABS) {
#pragma unroll
            for (int i = 0; i < NPT; ++i) u[i] = fmaf(a.k1, fabsf(z[i].x), a.k2 * fabsf(z[i].y)); // am.rs:86
        } else if ((demod == DEMOD_SSB || demod == DEMOD_USB) && a.post.exact) {
            unsigned ctr;
            float2 px = nco_exact_at(a.post, jt, ctr, warp_max_replay(jt));
            const float2 wx = make_float2(a.post.xwre, a.post.xwim);
#pragma unroll
            for (int i = 0; i < NPT; ++i) {                                                      // ssb.rs:36-37
                u[i] = fmaf(z[i].x, px.x, z[i].y * px.y);
                nco_step_exact(px, wx, ctr);
            }
        } else if (demod == DEMOD_SSB || demod == DEMOD_USB) {
#pragma unroll
            for (int i = 0; i < NPT; ++i) {                                                      // ssb.rs:36-37
                const float2 pa = scale2(p, nco_amp(a.post, kp0 + i));
                u[i] = fmaf(z[i].x, pa.x, z[i].y * pa.y);
                p = cmul_fma(p, w);
            }
        } else if (demod == DEMOD_CW) {
#pragma unroll
            for (int i = 0; i < NPT; ++i) u[i] = sqrtf(z[i].x * z[i].x + z[i].y * z[i].y);        // cw.rs:37
        }
    }

    if (demod == MOD_AM || demod == MOD_PM) {             // modulators: u[] holds the audio, z[] receives the IQ
        float2 p = make_float2(1.f, 0.f);
        const float2 w = make_float2(a.post.wre, a.post.wim);
        const unsigned long long kp0 = a.post.kbase + (unsigned long long)jt + 1ull;
        unsigned ctr = 0;
        const bool ex = a.post.exact != 0;
        const float2 wx = make_float2(a.post.xwre, a.post.xwim);
        if (ex) p = nco_exact_at(a.post, jt, ctr, warp_max_replay(jt));
        else p = nco_unit(a.post, kp0);
#pragma unroll
        for (int i = 0; i < NPT; ++i) {
            const float2 r = ex ? p : scale2(p, nco_amp(a.post, kp0 + i));
            if (demod == MOD_AM) {                         // modulate/am.rs:61-118: m = (cl + mi*x) [clamped] * g; out = m * rot.next()
                float m = a.k1 + a.k2 * u[i];
                if (a.k != 0.f) m = fminf(fmaxf(m, -1.0f), 1.0f);
                m = m * a.k3;
                z[i] = make_float2(m * r.x, m * r.y);
            } else {                                       // modulate/pm.rs:37-47: base = (cos phi, sin phi) * gain; mix_with_nco (unfused)
                const float phi = a.k1 * u[i];
                const float br = cosf(phi) * a.k2, bi = sinf(phi) * a.k2;
                z[i] = make_float2(br * r.x - bi * r.y, br * r.y + bi * r.x);
            }
            if (ex) nco_step_exact(p, wx, ctr);
            else p = cmul_fma(p, w);
        }
    }

    if (kind_c32_out(demod)) {
        float2 *out = reinterpret_cast<float2 *>(a.out);
        const bool al16 = ((reinterpret_cast<uintptr_t>(a.out) & 15u) == 0) && (NPT % 2 == 0);
        bool done = false;
        if constexpr (NPT == 16) {
            if (xs && al16 && j0 + kThreads * NPT <= a.n_out) {                 // whole warp tile: coalesced through the scratch
                float4 v[NPT / 2];
#pragma unroll
                for (int i = 0; i < NPT; i += 2) v[i / 2] = make_float4(z[i].x, z[i].y, z[i + 1].x, z[i + 1].y);
                warp_tile_store<NPT / 2>(out + j0, xs, lane, v);
                done = true;
            }
        }
        if (done) {
        } else if (al16 && full) {
#pragma unroll
            for (int i = 0; i < NPT; i += 2)
                *reinterpret_cast<float4 *>(out + jt + i) = make_float4(z[i].x, z[i].y, z[i + 1].x, z[i + 1].y);
        } else {
#pragma unroll
            for (int i = 0; i < NPT; ++i)
                if (jt + i < a.n_out) out[jt + i] = z[i];
        }
    }

}

// end-of-call duties that depend on the input only: the FIR history for the next call and the carried
// state no stage of this call touches.  Run by one warp at the START of the kernel (off the tail).
static __device__ __noinline__ void end_of_call_duties(const ChainArgs &a, int lane) {
    const bool need_prev = a.demod == DEMOD_FM || a.demod == DEMOD_PM;
    if (a.H > 0)
        for (int k0 = 0; k0 < a.H; k0 += kThreads)
            if (k0 + lane < a.H) a.hist_out[k0 + lane] = load_x(a, a.n_in - a.H + k0 + lane);
    __syncwarp();
    if (lane == 0) {
        if (!need_prev || a.n_out == 0) a.carry_out->prev = __ldcg(&a.carry_in->prev);
        for (int s = 0; s < kMaxSections; ++s)
            if (s >= a.nsec || a.n_out == 0 || kind_c32_out(a.demod)) a.carry_out->sec[s] = __ldcg(&a.carry_in->sec[s]);
    }
}

// store the f32 outputs of a tile
template <int NPT>
DEV void store_f32(const ChainArgs &a, long long tile, int lane, const float (&u)[NPT], unsigned char *xs = nullptr) {
    const long long j0 = tile * (long long)(kThreads * NPT);
    const long long jt = j0 + (long long)lane * NPT;
    float *out = reinterpret_cast<float *>(a.out);
    const bool al16 = ((reinterpret_cast<uintptr_t>(a.out) & 15u) == 0) && (NPT % 4 == 0);
    bool done = false;
    if constexpr (NPT == 16) {
        if (xs && al16 && j0 + kThreads * NPT <= a.n_out) {                     // whole warp tile: coalesced through the scratch
            float4 v[NPT / 4];
#pragma unroll
            for (int i = 0; i < NPT; i += 4) v[i / 4] = make_float4(u[i], u[i + 1], u[i + 2], u[i + 3]);
            warp_tile_store<NPT / 4>(out + j0, xs, lane, v);
            done = true;
        }
    }
    if (done) {
    } else if (al16 && jt + NPT <= a.n_out) {
#pragma unroll
        for (int i = 0; i < NPT; i += 4)
            *reinterpret_cast<float4 *>(out + jt + i) = make_float4(u[i], u[i + 1], u[i + 2], u[i + 3]);
    } else {
#pragma unroll
        for (int i = 0; i < NPT; ++i)
            if (jt + i < a.n_out) out[jt + i] = u[i];
    }
}

// DM_FM_LR4: the one group is known to be two biquads (D = 4, sections 0 and 1) -- no dispatch, no section-type
// tests, section coefficients at compile-time offsets of the parameter bank, scan tables in shared memory.
struct Lr4Tabs {                                   // shared-memory copy of GroupTables::lane / ::lb of group 0
    float lane[32][16];
    float lb[32][16];
    float imp[kMaxNpt][4];                         // ... and of GroupParam::imp (indexed at run time by the rolled front loop)
};
DEV void load_mat4_sh(const float *m, float (&M)[16]) {
#pragma unroll
    for (int i = 0; i < 16; i += 4) {
        const float4 v = *reinterpret_cast<const float4 *>(m + i);
        M[i] = v.x; M[i + 1] = v.y; M[i + 2] = v.z; M[i + 3] = v.w;
    }
}
// The FM front of one lane as ONE rolled loop over its NPT items: translate (fm.rs:49), conj-product discriminator
// (fm.rs:50-56), atan2_approx, the group's zero-state dot product.  The item arrays rotate through registers
// (z[0] is always the current item, the new u enters at the top), so the loop body exists once in the
// instruction cache instead of NPT times -- the kernel is instruction-fetch bound, not issue bound.
template <int NPT>
DEV void fm_front_rolled(const ChainArgs &a, const Lr4Tabs *tabs, long long tile, int lane, float2 (&z)[NPT], float (&u)[NPT],
                         float2 zhalo, float (&E)[4]) {
    const long long j0 = tile * (long long)(kThreads * NPT);
    const long long jt = j0 + (long long)lane * NPT;
    const bool xlate = a.translate != 0;
    const float2 w = make_float2(a.post.wre, a.post.wim);
    float2 p = make_float2(1.f, 0.f);
    if (xlate) {
        p = nco_unit(a.post, a.post.kbase + (unsigned long long)jt + 1ull);
        if (j0 > 0 && lane == 0) {
            const float2 ph = nco_unit(a.post, a.post.kbase + (unsigned long long)j0);
            const float cr = ph.x, ci = -ph.y;
            zhalo = make_float2(zhalo.x * cr - zhalo.y * ci, zhalo.x * ci + zhalo.y * cr);
        }
        __syncwarp();
    }
    // previous item of the lane's first one: the neighbour lane's last item, translated there
    float2 zl = z[NPT - 1];
    if (xlate) {
        const float2 pl = nco_unit(a.post, a.post.kbase + (unsigned long long)jt + (unsigned long long)NPT);
        const float cr = pl.x, ci = -pl.y;
        zl = make_float2(zl.x * cr - zl.y * ci, zl.x * ci + zl.y * cr);
    }
    float2 prev = shfl_up2(zl, 1);
    if (lane == 0) prev = (j0 > 0) ? zhalo : __ldcg(&a.carry_in->prev);
    __syncwarp();
    const int last = (int)max(min(a.n_out - 1 - jt, (long long)NPT), -1ll);      // item of this lane that ends the call
#pragma unroll
    for (int d = 0; d < 4; ++d) E[d] = 0.f;
#pragma unroll 1
    for (int it = 0; it < NPT; ++it) {
        float2 zz = z[0];
        if (xlate) {
            const float cr = p.x, ci = -p.y;
            zz = make_float2(zz.x * cr - zz.y * ci, zz.x * ci + zz.y * cr);
            p = cmul_fma(p, w);
        }
        const float pr = zz.x * prev.x + zz.y * prev.y;
        const float pi = zz.y * prev.x - zz.x * prev.y;
        const float uu = atan2_approx(pi, pr) * a.k;
        prev = zz;
        if (it == last) a.carry_out->prev = zz;
        const float4 im = *reinterpret_cast<const float4 *>(tabs->imp[it]);
        E[0] = fmaf(im.x, uu, E[0]); E[1] = fmaf(im.y, uu, E[1]); E[2] = fmaf(im.z, uu, E[2]); E[3] = fmaf(im.w, uu, E[3]);
#pragma unroll
        for (int i = 0; i + 1 < NPT; ++i) { z[i] = z[i + 1]; u[i] = u[i + 1]; }
        u[NPT - 1] = uu;
    }
}
template <int NPT>
DEV void lr4_front_park(const ChainArgs &a, const Hot *hot, long long tile, int lane, float (&E)[4], float *park) {
    float X[4], agg[4];
    __syncwarp();
    group_scan<4, true>(a, hot, 0, tile, lane, E, X, agg);     // the LR4 instance: always two cascaded biquads
    if (!hot->grp[0].agg_only && (tile & 31) == 31) slow_block_publish<4>(a, hot, 0, tile, lane, agg);
    *reinterpret_cast<float4 *>(park + lane * kMaxGroupDim) = make_float4(X[0], X[1], X[2], X[3]);
    if (lane == 0) *reinterpret_cast<float4 *>(park + 32 * kMaxGroupDim) = make_float4(agg[0], agg[1], agg[2], agg[3]);
    __syncwarp();
}
// Look-back when every predecessor weight beyond `depth` (<= 32) tiles is exactly zero: lane l reads the
// aggregate of tile - 1 - l (published a whole loop iteration ago), weighs it with Ac^(T*l), one warp sum.
DEV void lr4_lookback_short(const ChainArgs &a, const Lr4Tabs *tabs, long long tile, int lane, int depth, float (&sin)[4]) {
    const long long idx = tile - 1 - lane;
    float pay[4] = { 0.f, 0.f, 0.f, 0.f };
    const bool want = lane < depth && idx >= -1;
    if (want) {
        if (idx < 0) {                              // the state carried into this call sits "before tile 0"
            const float2 c0 = __ldcg(&a.carry_in->sec[0]), c1 = __ldcg(&a.carry_in->sec[1]);
            pay[0] = c0.x; pay[1] = c0.y; pay[2] = c1.x; pay[3] = c1.y;
        } else {
            const TileLink *lk = a.links + idx * kMaxGroups;
            int spins = 0;
            while (!read_link<4>(lk->agg, a.epoch, pay)) {
                if (++spins > (1 << 21)) {                                        // watchdog: never hang the device
                    // diagnostic code: 1 | (epoch tag seen & 0x7FF) << 4 | (epoch expected & 0x7FF) << 15 | lane << 26
                    const uint4 r0 = ld_relaxed_b128(lk->agg);
                    atomicExch(a.err_flag, 1 | (int)((r0.w & 0x7FFu) << 4) | (int)((a.epoch & 0x7FFu) << 15) | (int)(((unsigned)lane & 0x1Fu) << 26));
                    break;
                }
                __nanosleep(32);
            }
        }
    }
    __syncwarp();
    float term[4] = { 0.f, 0.f, 0.f, 0.f };
    if (want) {
        float m[16];
        load_mat4_sh(tabs->lb[lane], m);
        matvec4_tri(m, pay, term);
    }
#pragma unroll
    for (int d = 0; d < 4; ++d) {
        float v = term[d];
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(FULLMASK, v, o);
        sin[d] = v;
    }
}
template <int NPT>
DEV void lr4_finish_parked(const ChainArgs &a, const Hot *hot, const Lr4Tabs *tabs, long long tile, int lane, float (&u)[NPT],
                           const float *park, unsigned char *xs = nullptr) {
    const long long jt = tile * (long long)(kThreads * NPT) + (long long)lane * NPT;
    const float4 xv = *reinterpret_cast<const float4 *>(park + lane * kMaxGroupDim);
    float sin[4];
    const GroupParam &G = hot->grp[0];
    if ((ORION_TRACE && a.trace) && lane == 0) a.trace[tile * 16 + 14] = clock64();       // finish phase entered
    if (G.agg_only) {
        lr4_lookback_short(a, tabs, tile, lane, __ldg(&a.gtabs->depth), sin);
        if ((ORION_TRACE && a.trace) && lane == 0) a.trace[tile * 16 + 10] = clock64();   // look-back done
    } else {                                       // slow poles: the two-level look-back
        const float4 av = *reinterpret_cast<const float4 *>(park + 32 * kMaxGroupDim);
        const float agg_own[4] = { av.x, av.y, av.z, av.w };
        lookback_slow<4>(a, hot, 0, tile, lane, agg_own, sin);
    }
    float lm[16], st[4];
    load_mat4_sh(tabs->lane[lane], lm);
    matvec4_tri(lm, sin, st);
    st[0] += xv.x; st[1] += xv.y; st[2] += xv.z; st[3] += xv.w;
    // the reference recursion, sample by sample through both sections (iir.rs:78-83); items past the end of the
    // call are computed from zero-padded input and never stored, the carried state is taken at the last real one
    const SecParam &P0 = hot->sec[0], &P1 = hot->sec[1];
    const long long last = a.n_out - 1 - jt;
    if (last < 0 || last >= NPT) {                   // every lane but one of the whole call: the bare recursion
#pragma unroll
        for (int i = 0; i < NPT; ++i) {
            const float y = sec_step_t<SEC_BIQUAD>(P0, u[i], st[0], st[1]);
            u[i] = sec_step_t<SEC_BIQUAD>(P1, y, st[2], st[3]);
        }
    } else {                                         // the lane that holds the call's last item also leaves the carried state
        const int li = (int)last;
#pragma unroll
        for (int i = 0; i < NPT; ++i) {
            const float y = sec_step_t<SEC_BIQUAD>(P0, u[i], st[0], st[1]);
            u[i] = sec_step_t<SEC_BIQUAD>(P1, y, st[2], st[3]);
            if (li == i) {
                a.carry_out->sec[0] = make_float2(st[0], st[1]);
                a.carry_out->sec[1] = make_float2(st[2], st[3]);
            }
        }
    }
    __syncwarp();
    store_f32<NPT>(a, tile, lane, u, xs);
}

// the section phase of a tile whose group-0 front ran one loop iteration earlier (state parked in
// shared memory): group 0 finish, then the remaining groups front + finish, then the store
template <int NPT>
DEV void finish_sections(const ChainArgs &a, const Hot *hot, long long tile, int lane, float (&u)[NPT], const float *park,
                         unsigned char *xs = nullptr) {
    const long long jt = tile * (long long)(kThreads * NPT) + (long long)lane * NPT;
    const bool full = jt + NPT <= a.n_out;
    if (a.ngroups > 0) group_finish_parked<NPT>(a, hot, 0, tile, lane, u, full, jt, park);      // (chains with more groups: stage_step)
    store_f32<NPT>(a, tile, lane, u, xs);
}

// Chains with more than one section group run a software pipeline that is as deep as they have groups: in every loop
// iteration a warp runs the front of a new tile (up to the aggregate of group 0) and then, for each older tile it still
// holds, ONE stage: stage k finishes group k-1 (look-back, the reference recursion) and, unless that was the last
// group, forms and publishes the aggregate of group k.  Every aggregate is therefore published a whole iteration before
// the tiles behind it look for it -- for every group, not just the first.  (Before, groups 1.. were handled front +
// finish back to back, so each tile waited for aggregates its 31 predecessors were publishing at that very moment:
// the AM chain of BASELINE config 3 spent 56 % of its stall samples polling link records.)  The items of the tiles in
// flight live in shared memory between stages (item-major, conflict free).
// (inlined at both call sites on purpose: as a called function it cost 40 % on the AM chain -- 295 -> 415 us)
template <int NPT>
DEV void stage_step(const ChainArgs &a, const Hot *hot, int k, int S, long long tile, int lane, float *us, float *park,
                    unsigned char *xs) {
    float u[NPT];
#pragma unroll
    for (int i = 0; i < NPT; ++i) u[i] = us[i * kThreads + lane];
    const long long jt = tile * (long long)(kThreads * NPT) + (long long)lane * NPT;
    const bool full = jt + NPT <= a.n_out;
    const int g = a.stage_group[k - 1];
    if (g < 0) return;                               // idle stage: the tile only ages (slack for a slow-pole group's records)
    group_finish_parked<NPT>(a, hot, g, tile, lane, u, full, jt, park);
    if (g + 1 < a.ngroups) {
        group_front_park<NPT>(a, hot, g + 1, tile, lane, u, full, park);
#pragma unroll
        for (int i = 0; i < NPT; ++i) us[i * kThreads + lane] = u[i];
        __syncwarp();
    } else {
        store_f32<NPT>(a, tile, lane, u, xs);
        if (tile == a.ntiles - 1) handoff_signal(a, 1, lane);
    }
}

// direct front: items straight from global memory (rate-1 blocks)
template <int NPT>
DEV void front_direct(const ChainArgs &a, long long tile, int lane, float2 (&z)[NPT], float (&u)[NPT], float2 &zhalo,
                      unsigned char *xs) {
    const long long j0 = tile * (long long)(kThreads * NPT);
    const long long jt = j0 + (long long)lane * NPT;
    const bool need_prev = a.demod == DEMOD_FM || a.demod == DEMOD_PM;
    const bool whole = NPT == 16 && j0 + kThreads * NPT <= a.n_out;             // warp-uniform: the warp tile is full
    if (!kind_f32_in(a.demod)) {
        const float2 *in = reinterpret_cast<const float2 *>(a.in);
        const bool al16 = ((reinterpret_cast<uintptr_t>(a.in) & 15u) == 0) && (NPT % 2 == 0);
        bool done = false;
        if constexpr (NPT == 16) {
            if (xs && al16 && whole) {
                float4 v[NPT / 2];
                warp_tile_load<NPT / 2>(in + j0, xs, lane, v);
#pragma unroll
                for (int i = 0; i < NPT; i += 2) {
                    z[i] = make_float2(v[i / 2].x, v[i / 2].y);
                    z[i + 1] = make_float2(v[i / 2].z, v[i / 2].w);
                }
                done = true;
            }
        }
        if (done) {
        } else if (al16 && jt + NPT <= a.n_out) {
#pragma unroll
            for (int i = 0; i < NPT; i += 2) {
                const float4 v = __ldg(reinterpret_cast<const float4 *>(in + jt + i));
                z[i] = make_float2(v.x, v.y);
                z[i + 1] = make_float2(v.z, v.w);
            }
        } else {
#pragma unroll
            for (int i = 0; i < NPT; ++i)
                if (jt + i < a.n_out) z[i] = __ldg(in + jt + i);
        }
        if (a.mix != MIX_NONE && a.pre.exact) {
            unsigned ctr;
            float2 p = nco_exact_at(a.pre, jt, ctr, warp_max_replay(jt));
            const float2 w = make_float2(a.pre.xwre, a.pre.xwim);
#pragma unroll
            for (int i = 0; i < NPT; ++i) {
                z[i] = mix_apply(a.mix, z[i], p);
                nco_step_exact(p, w, ctr);
            }
        } else if (a.mix != MIX_NONE) {
            const unsigned long long k0 = a.pre.kbase + (unsigned long long)jt + 1ull;
            float2 p = nco_unit(a.pre, k0);
            const float2 w = make_float2(a.pre.wre, a.pre.wim);
#pragma unroll
            for (int i = 0; i < NPT; ++i) {
                z[i] = mix_apply(a.mix, z[i], scale2(p, nco_amp(a.pre, k0 + i)));
                p = cmul_fma(p, w);
            }
        }
        if (need_prev && j0 > 0 && lane == 0) zhalo = load_x_mixed(a, j0 - 1);
    } else {
        const float *in = reinterpret_cast<const float *>(a.in);
        const bool al16 = ((reinterpret_cast<uintptr_t>(a.in) & 15u) == 0) && (NPT % 4 == 0);
        bool done = false;
        if constexpr (NPT == 16) {
            if (xs && al16 && whole) {
                float4 v[NPT / 4];
                warp_tile_load<NPT / 4>(in + j0, xs, lane, v);
#pragma unroll
                for (int i = 0; i < NPT; i += 4) { u[i] = v[i / 4].x; u[i + 1] = v[i / 4].y; u[i + 2] = v[i / 4].z; u[i + 3] = v[i / 4].w; }
                done = true;
            }
        }
        if (done) {
        } else if (al16 && jt + NPT <= a.n_out) {
#pragma unroll
            for (int i = 0; i < NPT; i += 4) {
                const float4 v = __ldg(reinterpret_cast<const float4 *>(in + jt + i));
                u[i] = v.x; u[i + 1] = v.y; u[i + 2] = v.z; u[i + 3] = v.w;
            }
        } else {
#pragma unroll
            for (int i = 0; i < NPT; ++i)
                if (jt + i < a.n_out) u[i] = __ldg(in + jt + i);
        }
    }
}

// CTA-level control block of the stage ring (FRONT_STAGED)
struct __align__(16) RingCtl {
    unsigned long long full[kMaxStages];     // mbarrier per slot: "the tile of this fill has landed"
    int gen[kMaxStages];                     // index of the slot's latest fill (use k may only look at fill k)
    unsigned int cons;                       // consume counter of the CTA
    unsigned int done;                       // warps of the CTA that have run out of tiles
    int pgen[kMaxStages][3];                 // tap-split FIR: fill index for which partial sum q of the slot's tile is in place
    int lgen[kMaxStages];                    // ... and for which an edge tile has been loaded cooperatively (by sub-warp 0)
};

// The kernel.  Tiles are assigned statically and round-robin: CTA b owns tiles b, b+G, b+2G, ...
// (G = gridDim.x, all CTAs co-resident) and its NW warps take them in that order through a consume
// counter; the i-th tile of the CTA is staged in ring slot i % NS by its (i / NS)-th fill.
//   * NS < NW: a slot is busy only while its tile is in flight from HBM and under the FIR, about a
//     quarter of a tile's life; the rest of the time a warp works from registers, so more warps than
//     slots are resident and the register file -- not shared memory -- bounds the occupancy;
//   * the warp that has run the FIR on a slot refills it at once with the CTA's tile NS places ahead
//     (one cp.async.bulk.tensor), so loads run ahead of the consumers;
//   * each warp is software-pipelined over its tiles: iteration i runs the front of tile_i (FIR,
//     demod map, group-0 scan, publish) and then the section phase of tile_(i-1), whose look-back
//     reads aggregates that the neighbouring CTAs published a whole iteration ago (round-robin
//     assignment keeps neighbouring tiles in lock-step);
//   * no deadlock by construction: the oldest unfinished tile of the stream is either in a warp's
//     hands, and that warp never waits on a younger tile, or next in line at its CTA, whose warps
//     all hold older tiles that wait on still older, finished ones.
// BATCH = 1: blockIdx.y selects one of ChainArgs::batch independent, equally long streams (the channels of a bank after
// the shared front end, bank_kernels.cu).  The launch arguments are copied to shared memory once per CTA and the
// per-stream pointers patched there; everything below reads them through the same reference.
template <int FRONT, int R, int U, int SP, int DM, int BATCH = 0>
__global__ void __launch_bounds__(kThreads * kMaxWarpsPerCta, 1)
chain_kernel(const __grid_constant__ ChainArgs a_param, const __grid_constant__ CUtensorMap tmap) {
    constexpr int NPT = R * U;
    __shared__ __align__(16) unsigned char args_sh[BATCH ? offsetof(ChainArgs, taps2) : 16];
    if (BATCH) {
        const uint32_t *src = reinterpret_cast<const uint32_t *>(&a_param);
        uint32_t *dst = reinterpret_cast<uint32_t *>(args_sh);
        for (int i = threadIdx.x; i < (int)(offsetof(ChainArgs, taps2) / 4); i += blockDim.x) dst[i] = src[i];
        __syncthreads();
        if (threadIdx.x == 0) {
            ChainArgs *as = reinterpret_cast<ChainArgs *>(args_sh);
            const long long m = blockIdx.y;
            const long long ch = a_param.batch_chan ? a_param.batch_chan[m] : m;
            as->in = reinterpret_cast<const unsigned char *>(a_param.in) + ch * a_param.batch_in_stride;
            as->out = reinterpret_cast<unsigned char *>(a_param.out) + ch * a_param.batch_out_stride;
            as->carry_in = a_param.carry_in + m;
            as->carry_out = a_param.carry_out + m;
            as->links = a_param.links ? a_param.links + m * a_param.batch_links_stride : nullptr;
        }
        __syncthreads();
    }
    const ChainArgs &a = BATCH ? *reinterpret_cast<const ChainArgs *>(args_sh) : a_param;
    typedef Geo<SP> GE;
    const int HRc = GE::fixed ? GE::HR : a.HR;
    const int demod = Dm<DM>::demod(a);
    extern __shared__ __align__(1024) unsigned char smem[];
    __shared__ RingCtl ring;
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    const int NW = blockDim.x >> 5;
    const int NS = a.nstages;
    const long long G = gridDim.x, cta = blockIdx.x;
    const size_t stage_bytes = (size_t)(kThreads + HRc) * (GE::fixed ? GE::pitch : a.row_pitch);      // bytes one TMA load delivers
    const size_t stage_stride = (stage_bytes + 127) & ~(size_t)127;          // TMA destinations are 128-byte aligned
    // per-warp pipeline area behind the stage ring: [tile ids of the tiles in flight (16 B)] [scan state ("park") slots]
    // [item slots of the multi-group pipeline]
    const size_t park_bytes = 33 * kMaxGroupDim * sizeof(float);
    const size_t warp_pipe = 16 + (size_t)a.pipe_park_slots * park_bytes + (size_t)a.pipe_u_slots * (kThreads * NPT * sizeof(float));
    unsigned char *pipe = smem + (size_t)NS * stage_stride + (size_t)wid * warp_pipe;
    int *qtile = reinterpret_cast<int *>(pipe);
    float (*park)[33 * kMaxGroupDim] = reinterpret_cast<float (*)[33 * kMaxGroupDim]>(pipe + 16);
    float *uslots = reinterpret_cast<float *>(pipe + 16 + (size_t)a.pipe_park_slots * park_bytes);
    // polyphase tap table: behind the pipeline areas
    float2 *taps_sh = reinterpret_cast<float2 *>(smem + (size_t)NS * stage_stride + (size_t)NW * warp_pipe);
    // ... then the generic taps g[] (discriminator halo) and the section / group launch data
    float *g_sh = reinterpret_cast<float *>(taps_sh + a.ntaps2);
    Hot *hot_sh = reinterpret_cast<Hot *>(reinterpret_cast<unsigned char *>(g_sh) + (((size_t)a.Lg * sizeof(float) + 15) & ~(size_t)15));
    // Hot's layout equals ChainArgs::grp followed by ChainArgs::sec (static_assert below)
    const Hot *hot = ORION_HOT_SMEM ? hot_sh : reinterpret_cast<const Hot *>(a.grp);
    // The ring first: barriers, then the initial fills go out BEFORE the table copies below, so the first tiles
    // are already in flight from HBM while the CTA sets itself up.
    if (FRONT == FRONT_STAGED) {
        if (threadIdx.x == 0) {
            for (int s = 0; s < NS; ++s) {
                mbar_init(smem_u32(&ring.full[s]), 1); ring.gen[s] = -1; ring.lgen[s] = -1;
                ring.pgen[s][0] = ring.pgen[s][1] = ring.pgen[s][2] = -1;
            }
            asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        }
    }
    if (threadIdx.x == 0) { ring.cons = 0; ring.done = 0; }
    __syncthreads();
    griddep_launch_dependents();
    // fill k of slot s = the CTA's tile number i = k*NS + s (lane 0): TMA for interior tiles, a plain
    // arrive for edge tiles (their consumer loads them cooperatively); nothing past the end
    auto fill_slot = [&](int s, int k) {
        const long long t = cta + G * ((long long)k * NS + s);
        const bool pred = lane == 0 && t < a.ntiles;
        const uint32_t bar = smem_u32(&ring.full[s]);
        st_shared_volatile_pred(pred, smem_u32(&ring.gen[s]), k);        // before the arrive below (release)
        if (tile_is_interior(a, t)) {                                     // warp-uniform
            tma_fill_pred(pred, smem_u32(smem + (size_t)s * stage_stride), &tmap, 0,
                          (int)(t * kThreads - HRc - a.tma_row0), bar, (uint32_t)stage_bytes);
        } else {
            mbar_arrive_pred(pred, bar);
        }
        if (a.l2_prefetch > 0) {                                          // the tile this CTA stages l2_prefetch fills from now
            const long long tp = t + G * (long long)a.l2_prefetch;
            if (tp < a.ntiles && tile_is_interior(a, tp))
                tma_prefetch_l2_pred(lane == 0, &tmap, 0, (int)(tp * kThreads - HRc - a.tma_row0));
        }
    };

    if (FRONT == FRONT_STAGED)
        for (int s = wid; s < NS; s += NW) fill_slot(s, 0);            // initial fill of the ring

    // DM_FM_LR4: per-lane scan tables of the one group, behind the section data
    Lr4Tabs *tabs_sh = reinterpret_cast<Lr4Tabs *>(reinterpret_cast<unsigned char *>(hot_sh) + ((sizeof(Hot) + 15) & ~(size_t)15));
    // rate-1 blocks: per-warp transposing scratch behind the tables
    unsigned char *xs = (FRONT == FRONT_DIRECT) ? reinterpret_cast<unsigned char *>(tabs_sh + 1) + (size_t)wid * kXposeBytes : nullptr;
    // tap-split FIR: three partial-sum buffers per ring slot, behind the tables
    unsigned char *part_sh = reinterpret_cast<unsigned char *>(tabs_sh + 1);
    if (Dm<DM>::lr4) {
        const float *src_l = &a.gtabs->lane[0][0], *src_b = &a.gtabs->lb[0][0];
        for (int i = threadIdx.x; i < 32 * 16; i += blockDim.x) {
            (&tabs_sh->lane[0][0])[i] = __ldg(src_l + i);
            (&tabs_sh->lb[0][0])[i] = __ldg(src_b + i);
        }
        for (int i = threadIdx.x; i < kMaxNpt * 4; i += blockDim.x) (&tabs_sh->imp[0][0])[i] = a.grp[0].imp[i >> 2][i & 3];
    }
    if (FRONT == FRONT_STAGED) {
        for (int i = threadIdx.x; i < a.ntaps2; i += blockDim.x) taps_sh[i] = a.taps2[i];
        for (int i = threadIdx.x; i < a.Lg; i += blockDim.x) g_sh[i] = __ldg(a.g + i);
    }
    if (ORION_HOT_SMEM) {
        const int n_grp = (int)(sizeof(GroupParam) * kMaxGroups / sizeof(float)), n_sec = (int)(sizeof(SecParam) * kMaxSections / sizeof(float));
        const float *src_g = reinterpret_cast<const float *>(a.grp), *src_s = reinterpret_cast<const float *>(a.sec);
        float *dst_g = reinterpret_cast<float *>(hot_sh->grp), *dst_s = reinterpret_cast<float *>(hot_sh->sec);
        for (int i = threadIdx.x; i < n_grp; i += blockDim.x) dst_g[i] = src_g[i];
        for (int i = threadIdx.x; i < n_sec; i += blockDim.x) dst_s[i] = src_s[i];
    }

    const bool need_prev = demod == DEMOD_FM || demod == DEMOD_PM;
    const bool has_sections = !kind_c32_out(demod);
    if ((ORION_TRACE && a.trace) && threadIdx.x == 0) {                  // debug: kernel span in %globaltimer ns
        unsigned long long gt;
        asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(gt));
        atomicMin(reinterpret_cast<unsigned long long *>(a.trace + (long long)a.ntiles * 16), gt);
    }

    __syncthreads();                                       // tables in shared memory are complete
    // At most two launches of a block are ever in flight.  The link records alternate between two halves by launch
    // parity and the history / carried-state buffers rotate through three, so what this launch WRITES was last READ by
    // the launch before the previous one: nothing is published or handed over before every CTA of that launch has run
    // to its end (handoff[2 + launch parity] counts them; observed on long streams: with only the hardware's launch ordering a third
    // launch did start while the first still had tiles to finish).
    handoff_wait(a, a.depth_slot, a.depth_target);
    if (cta == (long long)(a.ntiles - 1) % G && wid == NW - 1) {
        handoff_wait(a, 1, a.carry_target);                // copies what the previous call writes up to its very end
        end_of_call_duties(a, lane);
        handoff_signal(a, 0, lane);                        // FIR history for the next call is in place
        handoff_signal(a, 1, lane);                        // ... and this warp's share of the carried state
    }

    float u_pend[NPT];
#pragma unroll
    for (int i = 0; i < NPT; ++i) u_pend[i] = 0.f;
    long long pend_tile = -1;
    int slot_pp = 0;
    int pipe_it = 0;                                       // tiles this warp has put into the multi-group pipeline
    auto stamp = [&](long long t, int k) {
        if ((ORION_TRACE && a.trace) && lane == 0) a.trace[t * 16 + k] = clock64();
        __syncwarp();
    };


    for (;;) {
        unsigned c = atom_add_shared_pred(lane == 0, smem_u32(&ring.cons), 1u);
        c = __shfl_sync(FULLMASK, c, 0);
        // tap-split FIR (long filters, FIR-only instance): `split` consecutive tickets share one tile, each takes a slice of
        // the tap rows; the last one adds the partial sums (fixed order) and goes on with the tile
        const int split = (FRONT == FRONT_STAGED && DM == DEMOD_NONE) ? a.split : 1;
        const int sub = (int)(c & (unsigned)(split - 1));           // split is 1, 2 or 4
        c >>= (split >> 1);
        const long long tile = cta + G * (long long)c;
        if (tile >= a.ntiles) break;
        // the first tiles read what the previous call carried over (FIR history, discriminator `prev`, section
        // states through the look-back), the last one writes what this call carries out
        const bool handoff_tile = tile < a.pdl_guard || tile == a.ntiles - 1;
        if (tile == 0) handoff_wait(a, 0, a.hist_target);  // its edge loader reads the history the previous call left
        stamp(tile, 0);
        if ((ORION_TRACE && a.trace) && lane == 0) {
            unsigned long long gt;
            asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(gt));
            a.trace[tile * 16 + 5] = (long long)gt;
        }
        int s = 0, k = 0;
        unsigned char *stage = smem;
        if (FRONT == FRONT_STAGED) {
            // c / NS by multiplication with ceil(2^32 / NS): exact while c * NS < 2^32 (tickets of one CTA)
            k = (NS > 1 && c < (1u << 27)) ? (int)__umulhi(c, a.ns_magic) : (int)(c / (unsigned)NS);   // (NS = 1: the magic number would be 2^32)
            s = (int)(c - (unsigned)k * (unsigned)NS);
            const uint32_t bar = smem_u32(&ring.full[s]);
            // The slot's tile is not there yet and a finished front is pending: run that tile's section phase NOW instead
            // of waiting (it would otherwise run after this tile's front) -- the wait for HBM becomes useful work.
            if (ORION_EARLY_FINISH && has_sections && pend_tile >= 0 && (Dm<DM>::lr4 || a.pipe_u_slots == 0) &&
                !(mbar_test_wait(bar, (unsigned)k & 1u) && *reinterpret_cast<volatile int *>(&ring.gen[s]) == k)) {
                if (Dm<DM>::lr4) lr4_finish_parked<NPT>(a, hot, tabs_sh, pend_tile, lane, u_pend, park[slot_pp ^ 1], xs);
                else finish_sections<NPT>(a, hot, pend_tile, lane, u_pend, park[slot_pp ^ 1], xs);
                if (pend_tile == a.ntiles - 1) handoff_signal(a, 1, lane);
                pend_tile = -1;
            }
            // Use k of the slot waits for fill k: first on the mbarrier phase (a hardware-suspended wait, no
            // polling traffic), then ONE look at the fill index -- a warp more than one lap ahead of the
            // slot's current user sees the phase of fill k-2 as complete (the parity is one bit), finds the
            // index short, and only then falls back to a sleeping poll until fill k is announced.
            int spins = 0;
            for (;;) {
                while (!mbar_try_wait(bar, (unsigned)k & 1u)) {
                    if (++spins > (1 << 22)) { atomicExch(a.err_flag, 2); break; }
                }
                if (*reinterpret_cast<volatile int *>(&ring.gen[s]) == k || spins > (1 << 22)) break;
                while (*reinterpret_cast<volatile int *>(&ring.gen[s]) != k) {
                    if (++spins > (1 << 22)) { atomicExch(a.err_flag, 3); break; }
                    __nanosleep(64);
                }
            }
            stage = smem + (size_t)s * stage_stride;
        }
        stamp(tile, 1);

        float2 z[NPT];
        float  u[NPT];
#pragma unroll
        for (int i = 0; i < NPT; ++i) { z[i] = make_float2(0.f, 0.f); u[i] = 0.f; }
        float2 zhalo = make_float2(0.f, 0.f);           // item j0-1 (lane 0 only)
        const long long j0 = tile * (long long)(kThreads * NPT);
        const long long jt = j0 + (long long)lane * NPT;

        if (FRONT == FRONT_STAGED) {
            if (split > 1) {
                if (!tile_is_interior(a, tile)) {                 // edge tile: sub-warp 0 loads it, the others wait for that
                    if (sub == 0) {
                        stage_load_generic(a, tile, stage, lane);
                        __threadfence_block();
                        __syncwarp();
                        if (lane == 0) *reinterpret_cast<volatile int *>(&ring.lgen[s]) = k;
                    } else {
                        int spins = 0;
                        while (*reinterpret_cast<volatile int *>(&ring.lgen[s]) != k) {
                            if (++spins > (1 << 22)) { atomicExch(a.err_flag, 13); break; }
                            __nanosleep(64);
                        }
                        __threadfence_block();
                    }
                    __syncwarp();
                }
                const int rows = HRc / split;
                fir_staged<R, U, SP, kFirQUnroll>(a, stage, taps_sh, lane, z, sub * rows, sub * rows + rows);
                float2 *part = reinterpret_cast<float2 *>(part_sh + ((size_t)s * 3) * (kThreads * NPT * sizeof(float2)));
                if (sub < split - 1) {                            // a partial sum: park it, announce it, take the next ticket
                    float2 *mine = part + (size_t)sub * (kThreads * NPT);
#pragma unroll
                    for (int i = 0; i < NPT; ++i) mine[i * kThreads + lane] = z[i];
                    __threadfence_block();
                    __syncwarp();
                    if (lane == 0) *reinterpret_cast<volatile int *>(&ring.pgen[s][sub]) = k;
                    continue;
                }
                for (int q = 0; q < split - 1; ++q) {             // the last slice: add the others in a fixed order
                    int spins = 0;
                    while (*reinterpret_cast<volatile int *>(&ring.pgen[s][q]) != k) {
                        if (++spins > (1 << 22)) { atomicExch(a.err_flag, 14); break; }
                        __nanosleep(64);
                    }
                }
                __threadfence_block();
                __syncwarp();
                {
                    float2 sum[NPT];
#pragma unroll
                    for (int i = 0; i < NPT; ++i) sum[i] = part[i * kThreads + lane];
                    for (int q = 1; q < split - 1; ++q)
#pragma unroll
                        for (int i = 0; i < NPT; ++i) sum[i] = add2(sum[i], part[(size_t)q * (kThreads * NPT) + i * kThreads + lane]);
#pragma unroll
                    for (int i = 0; i < NPT; ++i) z[i] = add2(sum[i], z[i]);
                }
            } else {
            if (!tile_is_interior(a, tile)) stage_load_generic(a, tile, stage, lane);
            if (a.mix != MIX_NONE) stage_mix(a, tile, stage, lane);
            // the FIR-only instance affords the fully unrolled pair loop (loads hoisted across pairs); with a post phase
            // behind it the smaller rolled body wins (instruction cache)
            fir_staged<R, U, SP, (SP == 1 && DM == DEMOD_NONE) ? 4 : kFirQUnroll>(a, stage, taps_sh, lane, z);
            }
            if (need_prev && j0 > 0) zhalo = fir_staged_one(a, stage, g_sh, tile * kThreads - HRc, j0 - 1, lane);
            __syncwarp();                               // every lane is done with the slot: refill it
            stamp(tile, 2);
            fill_slot(s, k + 1);
            stamp(tile, 8);
        } else if (FRONT == FRONT_GLOBAL) {
#pragma unroll
            for (int i = 0; i < NPT; ++i)
                if (jt + i < a.n_out) z[i] = fir_global_one(a, jt + i);
            if (need_prev && j0 > 0 && lane == 0) zhalo = fir_global_one(a, j0 - 1);
        } else {
            if (a.l2_prefetch > 0) {
                // rate-1 blocks read their tile straight from global memory, one round trip per iteration and warp: pull
                // the tile this CTA takes l2_prefetch tickets from now into L2 (one 128-byte line per lane)
                const long long tp = tile + G * (long long)a.l2_prefetch;
                const int item = kind_f32_in(demod) ? 4 : 8;
                const long long off = tp * (long long)(kThreads * NPT) * item + (long long)lane * 128;
                if (off + 128 <= a.n_in * item && lane * 128 < kThreads * NPT * item)
                    asm volatile("prefetch.global.L2 [%0];" ::"l"(reinterpret_cast<const unsigned char *>(a.in) + off));
            }
            front_direct<NPT>(a, tile, lane, z, u, zhalo, xs);
        }
        if (handoff_tile) handoff_wait(a, 1, a.carry_target);   // `prev` now, section states in this tile's look-back
        float E4[4] = { 0.f, 0.f, 0.f, 0.f };
        if (DM == DM_FM_LR4 && ORION_FM_ROLLED) fm_front_rolled<NPT>(a, tabs_sh, tile, lane, z, u, zhalo, E4);
        else {
            front_map<NPT, DM>(a, tile, lane, z, u, zhalo, xs);
            if (Dm<DM>::lr4) {                           // zero-state dot product of the unrolled front
#pragma unroll
                for (int i = 0; i < NPT; ++i) {
                    const float4 im = *reinterpret_cast<const float4 *>(tabs_sh->imp[i]);
                    E4[0] = fmaf(im.x, u[i], E4[0]); E4[1] = fmaf(im.y, u[i], E4[1]);
                    E4[2] = fmaf(im.z, u[i], E4[2]); E4[3] = fmaf(im.w, u[i], E4[3]);
                }
            }
        }
        if ((ORION_TRACE && a.trace) && lane == 0) {
            unsigned smid;
            asm volatile("mov.u32 %0, %%smid;" : "=r"(smid));
            a.trace[tile * 16 + 6] = smid;
            a.trace[tile * 16 + 7] = wid + 100 * (long long)blockIdx.x;
            a.trace[tile * 16 + 9] = clock64();
            if (!has_sections) a.trace[tile * 16 + 3] = clock64();
        }
        if (!has_sections && tile == a.ntiles - 1) handoff_signal(a, 1, lane);      // the call's carried state is complete
        if (has_sections && !Dm<DM>::lr4 && a.pipe_u_slots > 0) {
            // multi-group chain: pipeline of depth S = ngroups (stage_step)
            const int S = a.nstage;
            group_front_park<NPT>(a, hot, 0, tile, lane, u, jt + NPT <= a.n_out, park[pipe_it % (S + 1)]);
            // youngest first: a stage that only publishes (k small) must never sit behind a stage that may wait for the
            // records of older tiles -- otherwise the block records of a slow-pole group chain through those waits
            for (int k = 1; k <= S; ++k) {
                if (pipe_it - k < 0) continue;
                const int e = (pipe_it - k) % S;
                const long long t = qtile[e];
                stage_step<NPT>(a, hot, k, S, t, lane, uslots + (size_t)e * (kThreads * NPT), park[(pipe_it - k) % (S + 1)], xs);
            }
            {
                const int e = pipe_it % S;
                float *us = uslots + (size_t)e * (kThreads * NPT);
#pragma unroll
                for (int i = 0; i < NPT; ++i) us[i * kThreads + lane] = u[i];
                if (lane == 0) qtile[e] = (int)tile;
                __syncwarp();
            }
            ++pipe_it;
        } else if (has_sections) {
            if (Dm<DM>::lr4) lr4_front_park<NPT>(a, hot, tile, lane, E4, park[slot_pp]);
            else if (a.ngroups > 0) group_front_park<NPT>(a, hot, 0, tile, lane, u, jt + NPT <= a.n_out, park[slot_pp]);
            stamp(tile, 3);
            if (pend_tile >= 0) {
                if (Dm<DM>::lr4) lr4_finish_parked<NPT>(a, hot, tabs_sh, pend_tile, lane, u_pend, park[slot_pp ^ 1], xs);
                else finish_sections<NPT>(a, hot, pend_tile, lane, u_pend, park[slot_pp ^ 1], xs);
                if (pend_tile == a.ntiles - 1) handoff_signal(a, 1, lane);
                stamp(pend_tile, 4);
            }
#pragma unroll
            for (int i = 0; i < NPT; ++i) u_pend[i] = u[i];
            pend_tile = tile;
            slot_pp ^= 1;
        }
    }
    if (pend_tile >= 0) {
        if (Dm<DM>::lr4) lr4_finish_parked<NPT>(a, hot, tabs_sh, pend_tile, lane, u_pend, park[slot_pp ^ 1], xs);
        else finish_sections<NPT>(a, hot, pend_tile, lane, u_pend, park[slot_pp ^ 1], xs);
        if (pend_tile == a.ntiles - 1) handoff_signal(a, 1, lane);
    }
    if (!Dm<DM>::lr4 && a.pipe_u_slots > 0 && pipe_it > 0) {           // drain the multi-group pipeline
        const int S = a.nstage;
        for (int j = pipe_it; j < pipe_it + S; ++j)
            for (int k = 1; k <= S; ++k) {
                const int en = j - k;                                    // entry iteration of the tile at stage k
                if (en < 0 || en >= pipe_it) continue;
                const int e = en % S;
                const long long t = qtile[e];
                stage_step<NPT>(a, hot, k, S, t, lane, uslots + (size_t)e * (kThreads * NPT), park[en % (S + 1)], xs);
            }
    }
    if ((ORION_TRACE && a.trace) && lane == 0) {
        unsigned long long gt;
        asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(gt));
        atomicMax(reinterpret_cast<unsigned long long *>(a.trace + (long long)a.ntiles * 16 + 1), gt);
    }
    // this warp reads nothing of the call any more; the last warp of the CTA reports the CTA as done
    __threadfence();
    __syncwarp();
    if (lane == 0 && atomicAdd(&ring.done, 1u) == (unsigned)NW - 1u) atomicAdd(a.handoff + a.depth_slot, 1u);
}

typedef void (*chain_kernel_t)(const ChainArgs, const CUtensorMap);

template <int FRONT, int R, int U, int SP = 0, int DM = -1, int BATCH = 0>
static chain_kernel_t kptr() { return chain_kernel<FRONT, R, U, SP, DM, BATCH>; }

// one getter per translation unit (nullptr: no such instance there)
chain_kernel_t get_kernel_hot(int front, int sp, int dm);   // fixed geometries (SP = 1, 2) + FRONT_GLOBAL
chain_kernel_t get_kernel_direct(int dm);                   // rate-1 blocks
chain_kernel_t get_kernel_direct_batch(int dm);             // rate-1 blocks, batched over blockIdx.y (channel bank)
chain_kernel_t get_kernel_staged_u1(int R);                 // generic staged instances, even M
chain_kernel_t get_kernel_staged_u2(int R);                 // generic staged instances, odd M

}  // namespace orion

static_assert(offsetof(orion::ChainArgs, sec) == offsetof(orion::ChainArgs, grp) + sizeof(orion::GroupParam) * orion::kMaxGroups,
              "ChainArgs::grp must be followed directly by ChainArgs::sec");
