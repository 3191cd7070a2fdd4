// bank_kernels.cu -- K5a: the channel-bank front end (BASELINE config 5).
//
// C independent narrowband channels are cut out of ONE wideband stream: channel c is
//     Rotator(-f_c).rotate_block  (rotator.rs:74-84)  ->  FirDecimator(L taps, keep every M-th)  (decim.rs:44-76)
// followed by its demodulator.  The reference runs C separate block chains over the same slice; round 1 did the
// same with C kernel launches, each staging the whole wideband buffer again and spending its time on tile
// hand-over instead of arithmetic.  Here the wideband stream is staged ONCE per CTA and every resident warp
// cuts 32 channels out of it:
//
//   * lane = channel.  All lanes of a warp walk the same input samples in the same order, so the sample and the
//     polyphase taps are shared-memory BROADCAST reads (one wavefront each) and the only per-lane state lives in
//     registers: the channel's phasor (the reference recurrence z <- z*w with its own f32 step, re-anchored on the
//     closed-form 64-bit phase every 32 steps, amplitude sawtooth of the 1024-step renormalisation included) and
//     PM running accumulators, one per output the sample contributes to (polyphase: only retained outputs).
//   * a CTA = NWC consumer warps (32*NWC channels) + one producer warp that streams tiles of BT*M input samples
//     through a ring of NS slots with cp.async.bulk + mbarrier (edge tiles -- FIR history of the previous call,
//     ragged tail -- with a cooperative loader).  grid = (time ranges, channel groups): a warp owns a contiguous
//     range of outputs, so the PM-block warm-up is paid once per range, not per tile.
//   * per sample and channel: 4 FMA rotate + 4 FMA phasor advance + 2*PM FMA taps (PM = 4 for 513 taps / 128):
//     16 FMA, all register-to-register.  Nothing is re-read from HBM: the wideband buffer (65 MB for one second)
//     stays in L2 across the channel groups.
//
// The decimated complex outputs Z[c][j] go to a [C][n_out] scratch in HBM (0.5 GB/s of wideband: noise next to the
// arithmetic); the demodulators + recursive sections then run as ONE batched launch of the rate-1 chain kernel per
// demodulator kind (chain_kernel<..., BATCH>, blockIdx.y = channel), with the same look-back scan as everywhere.
#include "chain_kernels.cuh"

namespace orion {

struct BankFirArgs {
    const float2 *in;            // wideband input, call-relative sample 0
    long long n_in, n_out;
    const float2 *hist_in;       // the H samples before sample 0
    float2 *hist_out;            // ... and for the next call
    int H;
    int mix;                     // MIX_ROTATE / MIX_NCO / MIX_NONE
    const NcoParam *osc;         // [nch] closed-form oscillator of every channel (its kbase field is ignored)
    unsigned long long kbase;    // input items consumed before this call (the counter of sample s is kbase + s + 1)
    int M, Lg, PM, BT, NS;       // decimation, taps, main polyphase branches, blocks per tile, ring slots
    const float *gt;             // [M][PM]: gt[i*PM + p-1] = g[M*p - i] (0 outside [1, Lg-1]); g = generic causal taps
    float g0;                    // g[0]: pairs with the newest sample (taps[L-1] of FirLowpass, fir.rs:57-66)
    float2 *z;                   // [nch][z_stride] decimated outputs
    long long z_stride;
    int nch;
    long long tiles_total;       // ceil(n_out / BT)
    int nranges;                 // time ranges (gridDim.x): range r covers blocks [r*n_out/nranges, (r+1)*n_out/nranges)
    int *err_flag;
};

// virtual wideband stream: history for negative indices, zeros past the end of the call
DEV float2 bank_load_x(const BankFirArgs &a, long long s) {
    if (s < 0) {
        const long long h = s + a.H;
        return h >= 0 ? __ldcg(a.hist_in + h) : make_float2(0.f, 0.f);
    }
    if (s < a.n_in) return __ldg(a.in + s);
    return make_float2(0.f, 0.f);
}

DEV void bulk_load(uint32_t dst, const void *src, uint32_t bytes, uint32_t mbar) {
    asm volatile(
        "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
        ::"r"(dst), "l"(src), "r"(bytes), "r"(mbar) : "memory");
}
DEV void mbar_arrive(uint32_t mbar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(mbar) : "memory");
}
DEV void mbar_wait(uint32_t mbar, uint32_t parity, int *err_flag, int code) {
    int spins = 0;
    while (!mbar_try_wait(mbar, parity)) {
        if (++spins > (1 << 24)) { atomicExch(err_flag, code); break; }      // watchdog: never hang the device
    }
}

// the producer's wait: it is NS tiles ahead and in no hurry, so it sleeps between polls instead of taking issue slots from
// the consumer warps of its SM (small banks run four producers per SM: their polling was 7.6 % of all instructions)
DEV void mbar_wait_sleepy(uint32_t mbar, uint32_t parity, int *err_flag, int code) {
    int spins = 0;
    while (!mbar_try_wait(mbar, parity)) {
        if (++spins > (1 << 22)) { atomicExch(err_flag, code); break; }      // watchdog: never hang the device
        __nanosleep(400);
    }
}

// packed f32x2 multiply (sm_100 FMUL2; FFMA2 is in chain_kernels.cuh).  NOTE: ptxas 12.9 contracts mul.rn.f32x2 followed by
// add.rn.f32x2 into one FFMA2 -- unlike the scalar .rn forms, which it never fuses -- so the packed add is not used here:
// where the reference rounds a product and a sum separately the arithmetic is scalar.
DEV f32x2 fmul2(f32x2 a, f32x2 b) {
    f32x2 d;
    asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b));
    return d;
}
// a channel's phasor p = (re, im) as one packed pair
struct Ph { f32x2 p; };
DEV void ph_set(Ph &P, float2 z) { P.p = pack2(z.x, z.y); }
template <int MIXK>
DEV void ph_step(Ph &P, float wr, float wi) {                      // p <- p * w  (the reference recurrence, rotator.rs:46-47)
    // (fma(zr, wr, -(zi*wi)), fma(zi, wr, zr*wi)): the reference's own products and rounding order.  The step enters as two
    // scalars and the phasor as the pair operand, so no pair of w / j*w has to be kept in (or copied into) an aligned
    // register pair; the two products are ONE packed multiply of the phasor with its halves swapped -- ptxas folds swap and
    // negation into the operand (FMUL2 -p.F32x2.LO_HI.NP, wi.F32)
    const float2 z = unpack2(P.p);
    P.p = ffma2(pack2(wr, wr), P.p, fmul2(pack2(z.y, z.x), pack2(-wi, wi)));
}
template <int MIXK>
DEV f32x2 bank_mix(float2 x, const Ph &P) {
    const float2 pz = unpack2(P.p);
    if (MIXK == MIX_ROTATE) {                                      // rotator.rs:74-84: x * p with FMAs
        // (xi * -pi, xi * pr) by two scalar multiplies (the negation is an operand modifier), then xr * (pr, pi) + that:
        // the reference's products and roundings, without forming j*p after every step (that was a negation and two
        // register moves per sample).  Same-box A/B: the packed form of these two multiplies is slower (FMA pipe).
        return ffma2(pack2(x.x, x.x), P.p, pack2(x.y * (-pz.y), x.y * pz.x));
    }
    if (MIXK == MIX_NCO)                                           // nco.rs:63-66: (xr*c - xi*s, xr*s + xi*c), every product and sum rounded
        return pack2(x.x * pz.x - x.y * pz.y, x.x * pz.y + x.y * pz.x);
    return pack2(x.x, x.y);
}
// one pair of samples into the PM running outputs: taps g[M*(p+1) - i] (sample i) and g[M*(p+1) - i - 1] (sample i+1)
template <int PM>
DEV void bank_mac(f32x2 (&acc)[PM], const float *gp, f32x2 x0, f32x2 x1) {
    float t0[PM], t1[PM];
    if (PM == 4) {
        const float4 a = *reinterpret_cast<const float4 *>(gp), b = *reinterpret_cast<const float4 *>(gp + 4);
        t0[0] = a.x; t0[1 % PM] = a.y; t0[2 % PM] = a.z; t0[3 % PM] = a.w;
        t1[0] = b.x; t1[1 % PM] = b.y; t1[2 % PM] = b.z; t1[3 % PM] = b.w;
    } else {
#pragma unroll
        for (int p = 0; p < PM; ++p) { t0[p] = gp[p]; t1[p] = gp[PM + p]; }
    }
#pragma unroll
    for (int p = 0; p < PM; ++p) acc[p] = ffma2(pack2(t0[p], t0[p]), x0, acc[p]);
#pragma unroll
    for (int p = 0; p < PM; ++p) acc[p] = ffma2(pack2(t1[p], t1[p]), x1, acc[p]);
}

constexpr int kBankConsumerWarps = 8;        // at most 256 channels per CTA (blockDim.x / 32 - 1 consumer warps at run time)
constexpr int kBankMaxSlots = 6;

template <int PM, int MIXK>
__global__ void __launch_bounds__(32 * (kBankConsumerWarps + 1), 2)      // (…, 3) fits in 71 registers without spills, but 24 consumer warps per SM are no faster than 16: the loop is bound by the FMA pipe and issue slots, not by latency
bank_fir_kernel(const __grid_constant__ BankFirArgs a) {
    extern __shared__ __align__(128) unsigned char smem[];
    __shared__ __align__(8) unsigned long long full[kBankMaxSlots], empty[kBankMaxSlots];
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    const int NS = a.NS, M = a.M, BT = a.BT;
    const int NWC = (int)(blockDim.x >> 5) - 1;      // consumer warps of this launch (small banks use small CTAs, more of them per SM)
    const size_t slot_bytes = (size_t)BT * M * sizeof(float2);
    float *gt_sh = reinterpret_cast<float *>(smem + (size_t)NS * slot_bytes);

    if (threadIdx.x == 0) {
        for (int s = 0; s < NS; ++s) {
            mbar_init(smem_u32(&full[s]), 1);
            mbar_init(smem_u32(&empty[s]), NWC);
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    for (int i = threadIdx.x; i < M * PM; i += blockDim.x) gt_sh[i] = __ldg(a.gt + i);
    __syncthreads();

    // this CTA's range of BLOCKS (block b = samples [M*b, M*b + M) = output b): ranges differ by at most one block, so the
    // CTAs of the single wave end together.  The range is walked in tiles of BT blocks from its first block; the last tile
    // may be cut short (it is still staged whole), tile 0 is the warm-up (the PM blocks before the range).
    const long long Ba = (long long)blockIdx.x * a.n_out / a.nranges;
    const long long Bb = ((long long)blockIdx.x + 1) * a.n_out / a.nranges;
    if (Ba >= Bb) return;
    const long long ntile = (Bb - Ba + BT - 1) / BT + 1;

    if (wid == NWC) {
        // ---------------- producer warp: stream the tiles of the range through the ring ----------------
        if (blockIdx.x == 0 && blockIdx.y == 0 && a.H > 0) {                 // FIR history for the next call
            for (int k0 = 0; k0 < a.H; k0 += 32)
                if (k0 + lane < a.H) a.hist_out[k0 + lane] = bank_load_x(a, a.n_in - a.H + k0 + lane);
        }
        const bool al16 = (reinterpret_cast<uintptr_t>(a.in) & 15u) == 0 && ((size_t)M * sizeof(float2)) % 16 == 0;
        for (long long q = 0; q < ntile; ++q) {
            const long long blk0 = Ba + (q - 1) * BT;                        // first block of the tile
            const int s = (int)(q % NS);
            const unsigned lap = (unsigned)(q / NS);
            mbar_wait_sleepy(smem_u32(&empty[s]), (lap & 1u) ^ 1u, a.err_flag, 5);  // first lap: passes at once
            const long long s0 = blk0 * (long long)M;                        // first sample of the tile
            unsigned char *dst = smem + (size_t)s * slot_bytes;
            const bool interior = al16 && s0 >= 0 && s0 + (long long)BT * M <= a.n_in;
            if (interior) {
                if (lane == 0) {
                    fence_proxy_async();
                    mbar_expect_tx(smem_u32(&full[s]), (uint32_t)slot_bytes);
                    bulk_load(smem_u32(dst), a.in + s0, (uint32_t)slot_bytes, smem_u32(&full[s]));
                }
            } else {
                const int total = BT * M;
                for (int c = lane; c < total; c += 32) reinterpret_cast<float2 *>(dst)[c] = bank_load_x(a, s0 + c);
                __threadfence_block();
                __syncwarp();
                if (lane == 0) mbar_arrive(smem_u32(&full[s]));
            }
            __syncwarp();
        }
        return;
    }

    // ---------------- consumer warps: lane = channel ----------------
    const int ch_raw = (blockIdx.y * NWC + wid) * 32 + lane;
    const bool ch_ok = ch_raw < a.nch;
    const int ch = ch_ok ? ch_raw : a.nch - 1;
    NcoParam osc = a.osc[ch];
    osc.kbase = a.kbase;
    // the reference's own f32 step w = (cosf(phi), sinf(phi)): |w| = 1 + amp_delta, so the walk between two anchors
    // grows in amplitude exactly like the reference's phasor does between two renormalisations
    const float wv = osc.xwre, wj = osc.xwim;                             // the step w = (wv, wj) as two scalars (ph_step)
    const float g0 = a.g0;
    const bool have_g0 = g0 != 0.0f;
    float2 *zrow = a.z + (long long)ch * a.z_stride;
    const long long ja = Ba, jb = min(Bb, a.n_out);                       // outputs this range stores

    f32x2 acc[PM];
#pragma unroll
    for (int p = 0; p < PM; ++p) acc[p] = pack2(0.f, 0.f);
    Ph P;
    ph_set(P, make_float2(1.f, 0.f));
    // Anchors (closed-form 64-bit phase, amplitude 1 + (ctr mod 1024) * amp_delta) sit where the ABSOLUTE step counter
    // is a multiple of 32, so the reference's renormalisation every 1024 steps always coincides with one.  The counter
    // of call-relative sample s is kbase + s + 1; M is even, so within a pair (i, i+1) the even counter always
    // belongs to the same half: par = 0 the first sample, par = 1 the second.
    const unsigned par = (unsigned)((osc.kbase + 1ull) & 1ull);

    for (long long q = 0; q < ntile; ++q) {
        const long long blk0 = Ba + (q - 1) * BT;
        const int s = (int)(q % NS);
        mbar_wait(smem_u32(&full[s]), (unsigned)(q / NS) & 1u, a.err_flag, 6);
        const float4 *slot = reinterpret_cast<const float4 *>(smem + (size_t)s * slot_bytes);
        const int bi0 = (q == 0) ? BT - PM : 0;                    // warm-up tile: only its last PM blocks matter
        for (int bi = bi0; bi < BT; ++bi) {
            const long long b = blk0 + bi;                         // block index; its first sample is M*b
            if (b >= Bb) break;                                    // the range's last tile may be cut short (warp-uniform)
            const unsigned long long c0 = osc.kbase + (unsigned long long)(b * (long long)M) + 1ull;   // counter of sample M*b
            const float4 *xs = slot + (size_t)bi * (M >> 1);
            const float *gp = gt_sh;
            // offset (even) of the first pair of this block that holds a multiple of 32
            const int ia = (MIXK == MIX_NONE) ? M : (int)((0u - ((unsigned)c0 + par)) & 31u);
            if (MIXK != MIX_NONE && q == 0 && bi == bi0) {
                // start of the walk: from the anchor before the first sample, so that every sample's phasor is a
                // function of its absolute counter alone (a channel's outputs do not depend on how the time axis or
                // the channels are partitioned: sharded banks are bit-identical to the full bank)
                const unsigned long long ca = c0 & ~31ull;
                ph_set(P, nco_phasor(osc, ca));
                for (int st = 0; st < (int)(c0 - ca); ++st) ph_step<MIXK>(P, wv, wj);
            }

            {   // sample M*b completes output j = b with its newest-sample tap g[0] (fir.rs:57-66: taps[L-1] * x[n])
                float2 o = unpack2(acc[0]);
                if (have_g0) {
                    Ph P0 = P;
                    if (MIXK != MIX_NONE && ia == 0) ph_set(P0, nco_phasor(osc, c0));
                    const float4 xx = xs[0];
                    const float2 r = unpack2(bank_mix<MIXK>(make_float2(xx.x, xx.y), P0));
                    o.x = fmaf(g0, r.x, o.x);
                    o.y = fmaf(g0, r.y, o.y);
                }
                if (ch_ok && b >= ja && b < jb) zrow[b] = o;
#pragma unroll
                for (int p = 0; p + 1 < PM; ++p) acc[p] = acc[p + 1];
                acc[PM - 1] = pack2(0.f, 0.f);
            }

            int i = 0, nexta = ia;
            while (i < M) {
                const int e = min(M, nexta);
#pragma unroll 2
                for (; i < e; i += 2, gp += 2 * PM) {              // walking pairs: no checks
                    const float4 xx = xs[i >> 1];                  // two samples, one broadcast LDS.128
                    const f32x2 x0 = bank_mix<MIXK>(make_float2(xx.x, xx.y), P);
                    if (MIXK != MIX_NONE) ph_step<MIXK>(P, wv, wj);
                    const f32x2 x1 = bank_mix<MIXK>(make_float2(xx.z, xx.w), P);
                    if (MIXK != MIX_NONE) ph_step<MIXK>(P, wv, wj);
                    bank_mac<PM>(acc, gp, x0, x1);
                }
                if (i < M) {                                       // the pair at i holds a multiple of 32: anchor it
                    const float4 xx = xs[i >> 1];
                    ph_set(P, nco_phasor(osc, c0 + (unsigned long long)i));
                    const f32x2 x0 = bank_mix<MIXK>(make_float2(xx.x, xx.y), P);
                    if (par == 0u) ph_step<MIXK>(P, wv, wj);
                    else ph_set(P, nco_phasor(osc, c0 + (unsigned long long)i + 1ull));
                    const f32x2 x1 = bank_mix<MIXK>(make_float2(xx.z, xx.w), P);
                    ph_step<MIXK>(P, wv, wj);
                    bank_mac<PM>(acc, gp, x0, x1);
                    i += 2; gp += 2 * PM; nexta += 32;
                }
            }
        }
        __syncwarp();
        if (lane == 0) mbar_arrive(smem_u32(&empty[s]));
    }
}

typedef void (*bank_fir_kernel_t)(const BankFirArgs);
template <int MIXK>
static bank_fir_kernel_t bank_fir_instance_pm(int PM) {
    switch (PM) {
        case 1: return bank_fir_kernel<1, MIXK>;
        case 2: return bank_fir_kernel<2, MIXK>;
        case 3: return bank_fir_kernel<3, MIXK>;
        case 4: return bank_fir_kernel<4, MIXK>;
        case 5: return bank_fir_kernel<5, MIXK>;
        case 6: return bank_fir_kernel<6, MIXK>;
        case 7: return bank_fir_kernel<7, MIXK>;
        case 8: return bank_fir_kernel<8, MIXK>;
    }
    return nullptr;
}
static bank_fir_kernel_t bank_fir_instance(int PM, int mix) {
    if (mix == MIX_ROTATE) return bank_fir_instance_pm<MIX_ROTATE>(PM);
    if (mix == MIX_NCO) return bank_fir_instance_pm<MIX_NCO>(PM);
    return bank_fir_instance_pm<MIX_NONE>(PM);
}

size_t bank_fir_smem_bytes(const BankFirArgs &a) {
    return (size_t)a.NS * a.BT * a.M * sizeof(float2) + (size_t)a.M * a.PM * sizeof(float) + 16;
}

// consumer warps per CTA for a bank of nch channels: as few channel groups as possible, evenly filled
int bank_fir_consumer_warps(int nch) {
    const int w = (nch + 31) / 32;
    const int groups = (w + kBankConsumerWarps - 1) / kBankConsumerWarps;
    return (w + groups - 1) / groups;
}
// occupancy set-up once per (PM, smem); grid = (time ranges, channel groups of 32 * nwc channels)
cudaError_t bank_fir_prepare(const BankFirArgs &a, int nwc, int *ctas_per_sm) {
    bank_fir_kernel_t k = bank_fir_instance(a.PM, a.mix);
    if (!k) return cudaErrorInvalidValue;
    const size_t smem = bank_fir_smem_bytes(a);
    cudaError_t e = cudaFuncSetAttribute((const void *)k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    e = cudaFuncSetAttribute((const void *)k, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
    if (e != cudaSuccess) return e;
    return cudaOccupancyMaxActiveBlocksPerMultiprocessor(ctas_per_sm, (const void *)k, 32 * (nwc + 1), smem);
}
cudaError_t bank_fir_launch(const BankFirArgs &a, int nranges, int nwc, cudaStream_t stream) {
    bank_fir_kernel_t k = bank_fir_instance(a.PM, a.mix);
    if (!k) return cudaErrorInvalidValue;
    const size_t smem = bank_fir_smem_bytes(a);
    const int ngroups = (a.nch + 32 * nwc - 1) / (32 * nwc);
    dim3 grid((unsigned)nranges, (unsigned)ngroups);
    k<<<grid, 32 * (nwc + 1), smem, stream>>>(a);
    return cudaGetLastError();
}

}  // namespace orion
