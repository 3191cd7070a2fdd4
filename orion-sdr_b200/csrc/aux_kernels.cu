// aux_kernels.cu -- the "next" rows of SURVEY.md section 8(f) that are not chains of the fused kernel family:
//   * FmPhaseAccumMod (src/modulate/fm.rs:45-72), SsbPhasingMod (src/modulate/ssb.rs:42-114): row 1
//   * BPSK / QPSK / QAM soft-symbol gain blocks and hard-decision slicers (src/demodulate/{bpsk,qpsk,qam}.rs): row 4
// (CwKeyedMod shares the chunked envelope kernel of the AGC, agc_kernels.cu.)
// All are rate-1 and HBM-bound; grids are sized in multiples of the SM count by the callers.
#include "chain_kernels.cuh"

namespace orion {

// ---------------------------------------------------------------------------------------------------------------
// soft-symbol gain (BpskDemod / QpskDemod / QamDemod::process: out = (g*re, g*im)) and the hard-decision slicers
// ---------------------------------------------------------------------------------------------------------------
__global__ void gain_c32_kernel(const float2 *__restrict__ in, float2 *__restrict__ out, long long n, float g) {
    const long long stride = (long long)gridDim.x * blockDim.x * 2;
    for (long long i = ((long long)blockIdx.x * blockDim.x + threadIdx.x) * 2; i < n; i += stride) {
        if (i + 1 < n && ((reinterpret_cast<uintptr_t>(in) | reinterpret_cast<uintptr_t>(out)) & 15u) == 0) {
            const float4 v = __ldg(reinterpret_cast<const float4 *>(in + i));
            *reinterpret_cast<float4 *>(out + i) = make_float4(g * v.x, g * v.y, g * v.z, g * v.w);
        } else {
            for (long long k = i; k < min(i + 2, n); ++k) { const float2 v = __ldg(in + k); out[k] = make_float2(g * v.x, g * v.y); }
        }
    }
}

struct SliceArgs {
    const float2 *in;
    unsigned char *out;
    long long n_syms;
    int bits;                    // 1 BPSK (bpsk.rs:67-88), 2 QPSK (qpsk.rs:68-98), 4 / 6 / 8 QAM (qam.rs:106-178)
    float th[15];                // QAM per-axis thresholds, ascending (qam.rs:20-31)
};
// qam.rs:122-137: natural index = number of thresholds below v, Gray-coded, MSB first
DEV void qam_axis(const SliceArgs &a, float v, unsigned char *o, int k) {
    const int m = 1 << k;
    int nat = 0;
    for (int t = 0; t < m - 1; ++t) nat += (v > a.th[t]) ? 1 : 0;
    const int gray = nat ^ (nat >> 1);
    for (int b = 0; b < k; ++b) o[b] = (unsigned char)((gray >> (k - 1 - b)) & 1);
}
__global__ void slice_kernel(const __grid_constant__ SliceArgs a) {
    const long long stride = (long long)gridDim.x * blockDim.x;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < a.n_syms; i += stride) {
        const float2 x = __ldg(a.in + i);
        if (a.bits == 1) {
            a.out[i] = x.x < 0.0f ? 1 : 0;
        } else if (a.bits == 2) {
            a.out[2 * i] = x.x < 0.0f ? 1 : 0;
            a.out[2 * i + 1] = x.y < 0.0f ? 1 : 0;
        } else {
            unsigned char o[8];
            const int k = a.bits / 2;
            qam_axis(a, x.x, o, k);
            qam_axis(a, x.y, o + k, k);
            for (int b = 0; b < a.bits; ++b) a.out[i * a.bits + b] = o[b];
        }
    }
}
cudaError_t gain_c32_launch(const void *in, void *out, long long n, float g, int sms, cudaStream_t st) {
    const long long want = (n / 2 + 255) / 256;
    const int blocks = (int)std::max<long long>(1, std::min<long long>(want, (long long)sms * 8));
    gain_c32_kernel<<<blocks, 256, 0, st>>>((const float2 *)in, (float2 *)out, n, g);
    return cudaGetLastError();
}
cudaError_t slice_launch(const SliceArgs &a, int sms, cudaStream_t st) {
    const long long want = (a.n_syms + 255) / 256;
    const int blocks = (int)std::max<long long>(1, std::min<long long>(want, (long long)sms * 8));
    slice_kernel<<<blocks, 256, 0, st>>>(a);
    return cudaGetLastError();
}

// ---------------------------------------------------------------------------------------------------------------
// FmPhaseAccumMod (modulate/fm.rs:45-72).  The reference keeps a running phasor z <- z * e^{j dphi_i}, dphi_i = kf * x_i,
// renormalised every 1024 steps: a data-dependent product with no forgetting, i.e. a prefix sum of the phase.  Here
// the phase is accumulated EXACTLY in 64-bit fixed point (turns * 2^64; integer adds are associative, so the result
// does not depend on how the stream is cut), tile sums -> exclusive scan -> per-item phasor.  What this cannot
// reproduce is the reference's own f32 rounding walk (~4e-8 rad per step, random): the two stay within 1e-4 rad for
// about 10^6 samples after a reset -- the tests say so.  base = z * gain; out = mix_with_nco(base, rf_nco) (nco.rs:63-66,
// unfused) with the rf oscillator replayed bit-exactly.
// ---------------------------------------------------------------------------------------------------------------
constexpr int kFmTile = 1024;    // items per block: 256 threads x 4

DEV long long fm_dphi_q(float kf, float x) {
    const float dphi = kf * x;                                       // fm.rs:53
    return __double2ll_rn((double)dphi * 2935890503282001226.2);    // dphi / 2pi * 2^64
}
__global__ void __launch_bounds__(256) fm_reduce_kernel(const float *__restrict__ x, long long n, float kf, long long *tile_sum) {
    __shared__ long long wsum[8];
    const long long base = (long long)blockIdx.x * kFmTile + threadIdx.x * 4;
    long long s = 0;
#pragma unroll
    for (int i = 0; i < 4; ++i)
        if (base + i < n) s += fm_dphi_q(kf, __ldg(x + base + i));
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(FULLMASK, s, o);
    if ((threadIdx.x & 31) == 0) wsum[threadIdx.x >> 5] = s;
    __syncthreads();
    if (threadIdx.x == 0) {
        long long t = 0;
        for (int w = 0; w < 8; ++w) t += wsum[w];
        tile_sum[blockIdx.x] = t;
    }
}
// exclusive scan of the tile sums (one block; ntiles is n / 1024), seeded with the carried phase
__global__ void __launch_bounds__(1024) fm_scan_kernel(long long *tile_sum, long long ntiles, const CarryState *carry_in) {
    __shared__ long long part[1024];
    const long long per = (ntiles + 1023) / 1024;
    const long long a0 = (long long)threadIdx.x * per, a1 = min(a0 + per, ntiles);
    long long s = 0;
    for (long long i = a0; i < a1; ++i) s += tile_sum[i];
    part[threadIdx.x] = s;
    __syncthreads();
    if (threadIdx.x == 0) {
        long long run = *reinterpret_cast<const long long *>(&carry_in->pad);
        for (int i = 0; i < 1024; ++i) { const long long v = part[i]; part[i] = run; run += v; }
    }
    __syncthreads();
    long long run = part[threadIdx.x];
    for (long long i = a0; i < a1; ++i) { const long long v = tile_sum[i]; tile_sum[i] = run; run += v; }
}
DEV float2 unit_from_turns(unsigned long long ph) {                 // (cos, sin) of 2 pi ph / 2^64, as nco_unit does it
    const int hi = (int)(unsigned)(ph >> 32);
    const float xh = (float)hi;
    const long long rem = ((long long)ph >> 8) - ((long long)xh << 24);
    float s, c;
    __sincosf(xh * (4.656612873077393e-10f * 3.14159265358979323846f), &s, &c);
    const float d = (float)rem * (3.14159265358979f * 2.7755575615628914e-17f);
    return make_float2(fmaf(-s, d, c), fmaf(c, d, s));
}
struct FmApplyArgs {
    const float *x; float2 *out; long long n;
    float kf, gain;
    const long long *tile_off;
    NcoParam rf;
    const CarryState *carry_in; CarryState *carry_out;
};
__global__ void __launch_bounds__(256) fm_apply_kernel(const __grid_constant__ FmApplyArgs a) {
    __shared__ long long wsum[8];
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    const long long base = (long long)blockIdx.x * kFmTile + threadIdx.x * 4;
    long long q[4], s = 0;
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        q[i] = (base + i < a.n) ? fm_dphi_q(a.kf, __ldg(a.x + base + i)) : 0;
        s += q[i];
        q[i] = s;                                                   // inclusive within the thread
    }
    long long incl = s;                                             // inclusive scan over the warp
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const long long v = __shfl_up_sync(FULLMASK, incl, o);
        if (lane >= o) incl += v;
    }
    if (lane == 31) wsum[wid] = incl;
    __syncthreads();
    long long off = a.tile_off[blockIdx.x] + (incl - s);
    for (int w = 0; w < wid; ++w) off += wsum[w];
    // rf oscillator: the reference recurrence replayed from the checkpoint table (bit-exact phasors)
    unsigned ctr;
    const long long idx0 = min(base, a.n - 1);
    float2 p = nco_exact_at(a.rf, idx0, ctr, warp_max_replay(idx0));
    const float2 w = make_float2(a.rf.xwre, a.rf.xwim);
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        if (base + i < a.n) {
            const float2 z = unit_from_turns((unsigned long long)(off + q[i]));
            const float br = z.x * a.gain, bi = z.y * a.gain;       // fm.rs:68
            a.out[base + i] = make_float2(br * p.x - bi * p.y, br * p.y + bi * p.x);     // nco.rs:63-66
            if (base + i == a.n - 1) {
                CarryState cs = *a.carry_in;
                *reinterpret_cast<long long *>(&cs.pad) = off + q[i];
                *a.carry_out = cs;
            }
        }
        nco_step_exact(p, w, ctr);
    }
}
cudaError_t fm_mod_launch(const FmApplyArgs &a, long long *d_tile, cudaStream_t st) {
    const long long ntiles = (a.n + kFmTile - 1) / kFmTile;
    fm_reduce_kernel<<<(unsigned)ntiles, 256, 0, st>>>(a.x, a.n, a.kf, d_tile);
    fm_scan_kernel<<<1, 1024, 0, st>>>(d_tile, ntiles, a.carry_in);
    fm_apply_kernel<<<(unsigned)ntiles, 256, 0, st>>>(a);
    return cudaGetLastError();
}

// ---------------------------------------------------------------------------------------------------------------
// SsbPhasingMod (modulate/ssb.rs:42-114): p = aud_nco.next(); I = lp_i(x * p.re); Q = lp_q(x * p.im);
// z = (I, side * Q); out = z * rf_nco.next() (FMA form).  The two LpCascade filters run as the library's own LpCascade
// blocks (look-back scan); the kernels here do the products before and after them with both oscillators replayed
// bit-exactly.  One thread per 16 items: exactly one checkpoint of the replay table, no replay loop.
// ---------------------------------------------------------------------------------------------------------------
struct SsbSplitArgs { const float *x; float *xi, *xq; long long n; NcoParam aud; };
__global__ void __launch_bounds__(128) ssb_split_kernel(const __grid_constant__ SsbSplitArgs a) {
    const long long g = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    const long long i0 = g * 16;
    if (i0 >= a.n) return;
    float2 p = __ldg(a.aud.xfine + g);
    unsigned ctr = (unsigned)(a.aud.kbase + 1ull + (unsigned long long)i0);
    const float2 w = make_float2(a.aud.xwre, a.aud.xwim);
    for (int i = 0; i < 16 && i0 + i < a.n; ++i) {
        const float x = __ldg(a.x + i0 + i);
        a.xi[i0 + i] = x * p.x;                                     // ssb.rs:54
        a.xq[i0 + i] = x * p.y;                                     // ssb.rs:55
        nco_step_exact(p, w, ctr);
    }
}
struct SsbCombineArgs { const float *yi, *yq; float2 *out; long long n; float side; NcoParam rf; };
__global__ void __launch_bounds__(128) ssb_combine_kernel(const __grid_constant__ SsbCombineArgs a) {
    const long long g = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    const long long i0 = g * 16;
    if (i0 >= a.n) return;
    float2 r = __ldg(a.rf.xfine + g);
    unsigned ctr = (unsigned)(a.rf.kbase + 1ull + (unsigned long long)i0);
    const float2 w = make_float2(a.rf.xwre, a.rf.xwim);
    for (int i = 0; i < 16 && i0 + i < a.n; ++i) {
        const float zr = __ldg(a.yi + i0 + i), zi = a.side * __ldg(a.yq + i0 + i);            // ssb.rs:56
        a.out[i0 + i] = make_float2(fmaf(zr, r.x, -(zi * r.y)), fmaf(zi, r.x, zr * r.y));      // ssb.rs:58-61
        nco_step_exact(r, w, ctr);
    }
}
cudaError_t ssb_split_launch(const SsbSplitArgs &a, cudaStream_t st) {
    const long long groups = (a.n + 15) / 16;
    ssb_split_kernel<<<(unsigned)((groups + 127) / 128), 128, 0, st>>>(a);
    return cudaGetLastError();
}
cudaError_t ssb_combine_launch(const SsbCombineArgs &a, cudaStream_t st) {
    const long long groups = (a.n + 15) / 16;
    ssb_combine_kernel<<<(unsigned)((groups + 127) / 128), 128, 0, st>>>(a);
    return cudaGetLastError();
}

}  // namespace orion
