// chain_inst_ws.cu -- the warp-specialised instance of the decimate-by-8 chain (C1 / C3 geometry, Geo<1>).
//
// The unified chain_kernel runs FIR and post phase one after the other in the same 128-register warps, so the FIR's
// register window caps the SM at 16 warps and the kernel is latency-bound (round 1: issue slots 46 % busy, 0.64 of
// HBM peak).  Here a CTA is split by role:
//
//   * FIR warps (two warpgroups, 112 registers per thread after setmaxnreg.inc): take the CTA's tiles in ticket
//     order from the TMA stage ring (same ring, same refill rule as chain_kernel), run the packed polyphase FIR and the
//     discriminator halo, and hand the 256 decimated outputs of the tile (2 KB) to the post warps through a ring of z slots
//     in shared memory (full / empty mbarriers).  The stage slot is refilled as soon as the FIR has read it.
//   * post warps (four warpgroups, 64 registers per thread after setmaxnreg.dec): take z tiles in the same ticket
//     order and run everything behind the FIR -- translate, discriminator, atan2_approx, the group scan, publish, and,
//     software-pipelined one tile behind, the look-back and the reference recursion (the functions of chain_kernels.cuh,
//     unchanged: results are bit-identical to the unified kernel).
//
// 24 warps per SM instead of 16, and each role only ever waits on its own kind of latency.  Register pool: the kernel
// is launched at 80 registers x 768 threads = 61 440; 8 x 32 x 112 + 16 x 32 x 64 = 61 440.
//
// Ring discipline: a slow warp can keep an old ticket while its siblings go round the ring, so two fills of one slot
// may be waited for at the same time and the one-bit phase parity alone is ambiguous; every slot therefore also carries
// a fill index / release count that the waiter checks after the barrier (ring_wait).
// No deadlock: a post warp waits only for z of its own CTA (produced by FIR warps, which wait only for TMA data and for
// the z slot of an OLDER ticket to be read) and for look-back records of strictly older tiles; all CTAs are co-resident.
#include "chain_kernels.cuh"

namespace orion {

#ifndef ORION_WS_FIR_REGS
#define ORION_WS_FIR_REGS 112
#define ORION_WS_POST_REGS 64
#endif
#define ORION_STR2(x) #x
#define ORION_STR(x) ORION_STR2(x)
#ifndef ORION_WS_WAIT_NS
#define ORION_WS_WAIT_NS 20000
#endif
#ifndef ORION_WS_FIR_WARPS
#define ORION_WS_FIR_WARPS 8
#define ORION_WS_POST_WARPS 16
#endif
constexpr int kWsFirWarps = ORION_WS_FIR_WARPS;
constexpr int kWsPostWarps = ORION_WS_POST_WARPS;
constexpr int kWsWarps = kWsFirWarps + kWsPostWarps;
constexpr int kWsZSlots = 16;                       // >= kWsPostWarps (see above)
constexpr int kWsZPitch = 80;                       // bytes per lane: 8 complex outputs + 16 (conflict-free LDS.128 / STS.128)
constexpr int kWsZSlotBytes = 32 * kWsZPitch + 16;  // + the discriminator halo item

struct __align__(16) WsCtl {
    unsigned long long full[kMaxStages];            // stage ring: "the tile of this fill has landed"
    unsigned long long zfull[kWsZSlots], zempty[kWsZSlots];
    int gen[kMaxStages];
    int zgen[kWsZSlots], zrel[kWsZSlots];           // index of a z slot's latest fill / number of its releases (see ring_wait)
    unsigned int cons, zcons, done;
};

// Wait for one event of a ring slot.  First on the mbarrier phase (a hardware-suspended wait, no polling traffic), then
// ONE look at the slot's event counter: the phase parity is a single bit, so a warp that is a whole lap ahead of the
// slot's current user sees the phase of two events ago as complete; it then finds the counter short and falls back to a
// sleeping poll until the event it waits for has at least been announced, and takes the barrier again.
DEV void ring_wait(uint32_t bar, unsigned parity, const int *counter, int want, int *err_flag, int code) {
    int spins = 0;
    for (;;) {
        while (!mbar_try_wait_hint(bar, parity, ORION_WS_WAIT_NS)) {
            if (++spins > (1 << 20)) { atomicExch(err_flag, code); return; }     // watchdog: never hang the device
        }
        if (*reinterpret_cast<const volatile int *>(counter) == want) return;
        while (*reinterpret_cast<const volatile int *>(counter) != want) {
            if (++spins > (1 << 22)) { atomicExch(err_flag, code + 100); return; }
            __nanosleep(64);
        }
    }
}

size_t ws_dyn_smem(int nstages, int ntaps2, int Lg) {
    const size_t stage_stride = ((size_t)(kThreads + 1) * 528 + 127) & ~(size_t)127;
    return (size_t)nstages * stage_stride + (size_t)kWsZSlots * kWsZSlotBytes +
           (size_t)kWsPostWarps * 2 * 33 * kMaxGroupDim * sizeof(float) + (size_t)ntaps2 * sizeof(float2) +
           (((size_t)Lg * sizeof(float) + 15) & ~(size_t)15) + ((sizeof(Hot) + 15) & ~(size_t)15) + sizeof(Lr4Tabs) + 64;
}
int ws_warps() { return kWsWarps; }
int ws_max_stages(size_t smem_limit, int ntaps2, int Lg) {
    int ns = kMaxStages;
    while (ns > 1 && ws_dyn_smem(ns, ntaps2, Lg) + sizeof(WsCtl) + 1024 > smem_limit) --ns;
    return ns;
}

template <int DM>
__global__ void __launch_bounds__(kThreads * kWsWarps, 1)
chain_ws_kernel(const __grid_constant__ ChainArgs a, const __grid_constant__ CUtensorMap tmap) {
    constexpr int R = 8, NPT = 8, SP = 1;
    typedef Geo<SP> GE;
    static_assert(Dm<DM>::lr4, "the warp-specialised instance serves the demodulator + LR4 chains");
    extern __shared__ __align__(1024) unsigned char smem[];
    __shared__ WsCtl ctl;
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    const int NS = a.nstages;
    const long long G = gridDim.x, cta = blockIdx.x;
    constexpr size_t stage_bytes = (size_t)(kThreads + GE::HR) * GE::pitch;
    constexpr size_t stage_stride = (stage_bytes + 127) & ~(size_t)127;
    unsigned char *zring = smem + (size_t)NS * stage_stride;
    float (*park)[33 * kMaxGroupDim] = reinterpret_cast<float (*)[33 * kMaxGroupDim]>(zring + (size_t)kWsZSlots * kWsZSlotBytes);
    float2 *taps_sh = reinterpret_cast<float2 *>(reinterpret_cast<unsigned char *>(park) +
                                                 (size_t)kWsPostWarps * 2 * 33 * kMaxGroupDim * sizeof(float));
    float *g_sh = reinterpret_cast<float *>(taps_sh + a.ntaps2);
    Hot *hot_sh = reinterpret_cast<Hot *>(reinterpret_cast<unsigned char *>(g_sh) + (((size_t)a.Lg * sizeof(float) + 15) & ~(size_t)15));
    const Hot *hot = reinterpret_cast<const Hot *>(a.grp);
    Lr4Tabs *tabs_sh = reinterpret_cast<Lr4Tabs *>(reinterpret_cast<unsigned char *>(hot_sh) + ((sizeof(Hot) + 15) & ~(size_t)15));

    if (threadIdx.x == 0) {
        for (int s = 0; s < NS; ++s) { mbar_init(smem_u32(&ctl.full[s]), 1); ctl.gen[s] = -1; }
        for (int s = 0; s < kWsZSlots; ++s) {
            mbar_init(smem_u32(&ctl.zfull[s]), 1); mbar_init(smem_u32(&ctl.zempty[s]), 1);
            ctl.zgen[s] = -1; ctl.zrel[s] = 0;
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        ctl.cons = 0; ctl.zcons = 0; ctl.done = 0;
    }
    __syncthreads();
    griddep_launch_dependents();

    auto fill_slot = [&](int s, int k) {
        const long long t = cta + G * ((long long)k * NS + s);
        const bool pred = lane == 0 && t < a.ntiles;
        const uint32_t bar = smem_u32(&ctl.full[s]);
        st_shared_volatile_pred(pred, smem_u32(&ctl.gen[s]), k);
        if (tile_is_interior(a, t)) {
            tma_fill_pred(pred, smem_u32(smem + (size_t)s * stage_stride), &tmap, 0,
                          (int)(t * kThreads - GE::HR - a.tma_row0), bar, (uint32_t)stage_bytes);
        } else {
            mbar_arrive_pred(pred, bar);
        }
        if (a.l2_prefetch > 0) {
            const long long tp = t + G * (long long)a.l2_prefetch;
            if (tp < a.ntiles && tile_is_interior(a, tp))
                tma_prefetch_l2_pred(lane == 0, &tmap, 0, (int)(tp * kThreads - GE::HR - a.tma_row0));
        }
    };
    if (wid < kWsFirWarps)
        for (int s = wid; s < NS; s += kWsFirWarps) fill_slot(s, 0);           // the first tiles are in flight while the CTA sets up

    {
        const float *src_l = &a.gtabs->lane[0][0], *src_b = &a.gtabs->lb[0][0];
        for (int i = threadIdx.x; i < 32 * 16; i += blockDim.x) {
            (&tabs_sh->lane[0][0])[i] = __ldg(src_l + i);
            (&tabs_sh->lb[0][0])[i] = __ldg(src_b + i);
        }
        for (int i = threadIdx.x; i < kMaxNpt * 4; i += blockDim.x) (&tabs_sh->imp[0][0])[i] = a.grp[0].imp[i >> 2][i & 3];
        for (int i = threadIdx.x; i < a.ntaps2; i += blockDim.x) taps_sh[i] = a.taps2[i];
        for (int i = threadIdx.x; i < a.Lg; i += blockDim.x) g_sh[i] = __ldg(a.g + i);
    }
    __syncthreads();
    handoff_wait(a, a.depth_slot, a.depth_target);               // see chain_kernel: what this launch writes was read two launches ago

    const int demod = Dm<DM>::demod(a);
    const bool need_prev = demod == DEMOD_FM || demod == DEMOD_PM;

    if (wid < kWsFirWarps) {
        // ------------------------------------------ FIR warps ------------------------------------------
        asm volatile("setmaxnreg.inc.sync.aligned.u32 " ORION_STR(ORION_WS_FIR_REGS) ";");
        if (cta == (long long)(a.ntiles - 1) % G && wid == kWsFirWarps - 1) {      // (a noinline callee: called where registers are plentiful)
            handoff_wait(a, 1, a.carry_target);
            end_of_call_duties(a, lane);
            handoff_signal(a, 0, lane);
            handoff_signal(a, 1, lane);
        }
        for (;;) {
            unsigned c = atom_add_shared_pred(lane == 0, smem_u32(&ctl.cons), 1u);
            c = __shfl_sync(FULLMASK, c, 0);
            const long long tile = cta + G * (long long)c;
            if (tile >= a.ntiles) break;
            if (tile == 0) handoff_wait(a, 0, a.hist_target);
            const int s = (int)(c % (unsigned)NS), k = (int)(c / (unsigned)NS);
            {
                const uint32_t bar = smem_u32(&ctl.full[s]);
                int spins = 0;
                for (;;) {
                    while (!mbar_try_wait_hint(bar, (unsigned)k & 1u, ORION_WS_WAIT_NS)) {
                        if (++spins > (1 << 20)) { atomicExch(a.err_flag, 2); break; }
                    }
                    if (*reinterpret_cast<volatile int *>(&ctl.gen[s]) == k || spins > (1 << 20)) break;
                    while (*reinterpret_cast<volatile int *>(&ctl.gen[s]) != k) {
                        if (++spins > (1 << 22)) { atomicExch(a.err_flag, 3); break; }
                        __nanosleep(64);
                    }
                }
            }
            unsigned char *stage = smem + (size_t)s * stage_stride;
            const long long j0 = tile * (long long)(kThreads * NPT);
            float2 z[NPT];
            if (!tile_is_interior(a, tile)) stage_load_generic(a, tile, stage, lane);
            if (a.mix != MIX_NONE) stage_mix(a, tile, stage, lane);
            fir_staged<R, 1, SP, 4>(a, stage, taps_sh, lane, z);
            float2 zhalo = make_float2(0.f, 0.f);
            if (need_prev && j0 > 0) zhalo = fir_staged_one(a, stage, g_sh, tile * kThreads - GE::HR, j0 - 1, lane);
            __syncwarp();
            fill_slot(s, k + 1);
            // hand the tile's outputs to the post warps
            const int zs = (int)(c % (unsigned)kWsZSlots);
            const unsigned zk = c / (unsigned)kWsZSlots;
            ring_wait(smem_u32(&ctl.zempty[zs]), (zk & 1u) ^ 1u, &ctl.zrel[zs], (int)zk, a.err_flag, 7);    // fills 0 .. zk-1 have been read
            unsigned char *zp = zring + (size_t)zs * kWsZSlotBytes;
#pragma unroll
            for (int i = 0; i < NPT; i += 2)
                *reinterpret_cast<float4 *>(zp + lane * kWsZPitch + i * 8) = make_float4(z[i].x, z[i].y, z[i + 1].x, z[i + 1].y);
            if (lane == 0) *reinterpret_cast<float2 *>(zp + 32 * kWsZPitch) = zhalo;
            __syncwarp();
            st_shared_volatile_pred(lane == 0, smem_u32(&ctl.zgen[zs]), (int)zk);       // before the arrive below (release)
            mbar_arrive_pred(lane == 0, smem_u32(&ctl.zfull[zs]));
        }
    } else {
        // ------------------------------------------ post warps -----------------------------------------
        asm volatile("setmaxnreg.dec.sync.aligned.u32 " ORION_STR(ORION_WS_POST_REGS) ";");
        const int pw = wid - kWsFirWarps;
        float u_pend[NPT];
#pragma unroll
        for (int i = 0; i < NPT; ++i) u_pend[i] = 0.f;
        long long pend_tile = -1;
        int slot_pp = 0;
        float (*mypark)[33 * kMaxGroupDim] = park + 2 * pw;
        for (;;) {
            unsigned d = atom_add_shared_pred(lane == 0, smem_u32(&ctl.zcons), 1u);
            d = __shfl_sync(FULLMASK, d, 0);
            const long long tile = cta + G * (long long)d;
            if (tile >= a.ntiles) break;
            const bool handoff_tile = tile < a.pdl_guard || tile == a.ntiles - 1;
            const int zs = (int)(d % (unsigned)kWsZSlots);
            const unsigned zk = d / (unsigned)kWsZSlots;
            ring_wait(smem_u32(&ctl.zfull[zs]), zk & 1u, &ctl.zgen[zs], (int)zk, a.err_flag, 8);
            const unsigned char *zp = zring + (size_t)zs * kWsZSlotBytes;
            float2 z[NPT];
            float u[NPT];
#pragma unroll
            for (int i = 0; i < NPT; i += 2) {
                const float4 v = *reinterpret_cast<const float4 *>(zp + lane * kWsZPitch + i * 8);
                z[i] = make_float2(v.x, v.y);
                z[i + 1] = make_float2(v.z, v.w);
            }
            const float2 zhalo = *reinterpret_cast<const float2 *>(zp + 32 * kWsZPitch);
            __syncwarp();
            st_shared_volatile_pred(lane == 0, smem_u32(&ctl.zrel[zs]), (int)zk + 1);
            mbar_arrive_pred(lane == 0, smem_u32(&ctl.zempty[zs]));
#pragma unroll
            for (int i = 0; i < NPT; ++i) u[i] = 0.f;
            if (handoff_tile) handoff_wait(a, 1, a.carry_target);
            front_map<NPT, DM>(a, tile, lane, z, u, zhalo, nullptr);
            float E4[4] = { 0.f, 0.f, 0.f, 0.f };
#pragma unroll
            for (int i = 0; i < NPT; ++i) {
                const float4 im = *reinterpret_cast<const float4 *>(tabs_sh->imp[i]);
                E4[0] = fmaf(im.x, u[i], E4[0]); E4[1] = fmaf(im.y, u[i], E4[1]);
                E4[2] = fmaf(im.z, u[i], E4[2]); E4[3] = fmaf(im.w, u[i], E4[3]);
            }
            lr4_front_park<NPT>(a, hot, tile, lane, E4, mypark[slot_pp]);
            if (pend_tile >= 0) {
                lr4_finish_parked<NPT>(a, hot, tabs_sh, pend_tile, lane, u_pend, mypark[slot_pp ^ 1], nullptr);
                if (pend_tile == a.ntiles - 1) handoff_signal(a, 1, lane);
            }
#pragma unroll
            for (int i = 0; i < NPT; ++i) u_pend[i] = u[i];
            pend_tile = tile;
            slot_pp ^= 1;
        }
        if (pend_tile >= 0) {
            lr4_finish_parked<NPT>(a, hot, tabs_sh, pend_tile, lane, u_pend, mypark[slot_pp ^ 1], nullptr);
            if (pend_tile == a.ntiles - 1) handoff_signal(a, 1, lane);
        }
    }
    __threadfence();
    __syncwarp();
    if (lane == 0 && atomicAdd(&ctl.done, 1u) == (unsigned)kWsWarps - 1u) atomicAdd(a.handoff + a.depth_slot, 1u);
}

chain_kernel_t get_kernel_ws(int dm) {
    if (dm == DM_LR4 + DEMOD_FM) return chain_ws_kernel<DM_LR4 + DEMOD_FM>;
    if (dm == DM_LR4 + DEMOD_PM) return chain_ws_kernel<DM_LR4 + DEMOD_PM>;
    return nullptr;
}

}  // namespace orion
