// agc_kernels.cu -- AgcRms / AgcRmsIq (src/dsp/agc.rs:8-150) on the GPU.  SURVEY.md section 8(f) row 3.
//
// The envelope tracker  env <- a*env + (1-a)*x2,  a = attack_a if x2 > env else release_a  (agc.rs:33-40)  is a
// data-dependent recurrence: the coefficient of every step depends on the state, so there is no fixed-size monoid to scan
// with.  It is, however, a monotone contraction: both branches have slope a < 1 and meet at env = x2, so two
// trajectories over the same samples approach each other by a factor <= max(attack_a, release_a) per step whatever
// branches they take.  The stream is therefore cut into chunks, one per thread; a thread starts W samples before its
// chunk from a guessed envelope and runs the reference recursion over the warm-up (no output), after which its
// envelope differs from the reference's by at most amax^W of the initial error -- W is chosen by the host so that this
// is below 2^-26 (relative).  Chunks whose warm-up would start before sample 0 start AT sample 0 from the state carried
// in from the previous call, with the reference's seeding rule (agc.rs:58-61): they are exact, and so is the carried
// state as long as a call is not longer than one chunk.  Gain and output per sample follow agc.rs:64-69 op for op.
#include "chain_kernels.cuh"

namespace orion {

struct AgcArgs {
    const void *in;
    void *out;
    long long n;
    int iq;                          // 0: f32 -> f32 (AgcRms), 1: C32 -> C32 (AgcRmsIq), 2: f32 -> C32 (CwKeyedMod, below)
    float attack_a, release_a, target_rms, min_gain, max_gain;
    long long L, W;                  // chunk length, warm-up length
    const CarryState *carry_in;
    CarryState *carry_out;
    NcoParam osc; float gain;        // CwKeyedMod: tone oscillator (replayed bit-exactly) and output gain
};

template <bool IQ>
DEV float agc_x2(const AgcArgs &a, long long i, float &xr, float &xi) {
    if (IQ) {
        const float2 x = __ldg(reinterpret_cast<const float2 *>(a.in) + i);
        xr = x.x; xi = x.y;
        return x.x * x.x + x.y * x.y;                           // agc.rs:138 (unfused)
    }
    xr = __ldg(reinterpret_cast<const float *>(a.in) + i);
    xi = 0.f;
    return xr * xr;                                             // agc.rs:65
}
DEV float agc_env_step(const AgcArgs &a, float env, float x2) { // agc.rs:33-40
    const float c = (x2 > env) ? a.attack_a : a.release_a;
    return c * env + (1.0f - c) * x2;
}

template <bool IQ>
__global__ void __launch_bounds__(128) agc_kernel(const __grid_constant__ AgcArgs a) {
    const long long c = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    const long long s0 = c * a.L;
    if (s0 >= a.n) return;
    const long long s1 = min(s0 + a.L, a.n);
    float xr, xi;
    float env;
    long long i = s0 - a.W;
    if (i <= 0) {                                               // exact: from the carried state, seeding rule included
        i = 0;
        env = __ldcg(&a.carry_in->pad.x);
        if (env == 0.0f) env = fmaxf(agc_x2<IQ>(a, 0, xr, xi), 1e-12f);      // agc.rs:58-61 / :132-135
    } else {
        env = fmaxf(agc_x2<IQ>(a, i, xr, xi), 1e-12f);          // any start will do: the warm-up forgets it
    }
    for (; i < s0; ++i) env = agc_env_step(a, env, agc_x2<IQ>(a, i, xr, xi));
    for (; i < s1; ++i) {
        const float x2 = agc_x2<IQ>(a, i, xr, xi);
        env = agc_env_step(a, env, x2);
        const float rms = fmaxf(sqrtf(env), 1e-6f);             // agc.rs:66
        float g = a.target_rms / rms;
        g = fminf(fmaxf(g, a.min_gain), a.max_gain);            // f32::clamp
        if (IQ) reinterpret_cast<float2 *>(a.out)[i] = make_float2(g * xr, g * xi);
        else reinterpret_cast<float *>(a.out)[i] = g * xr;
    }
    if (s1 == a.n) {                                            // the thread that owns the last sample carries the state out
        CarryState cs = *a.carry_in;
        cs.pad.x = env;
        *a.carry_out = cs;
    }
}

// CwKeyedMod (src/modulate/cw.rs:44-102): tgt = clamp(x, 0, 1); env <- a*env + (1-a)*tgt with a = alpha_rise if tgt >= env
// else alpha_fall; out = mix_with_nco((env * gain, 0), nco).  The same kind of recurrence as the AGC's tracker (both
// branches contract, they meet at env = tgt), so the same chunking applies; attack_a / release_a carry alpha_rise / alpha_fall.
__global__ void __launch_bounds__(128) cw_mod_kernel(const __grid_constant__ AgcArgs a) {
    const long long c = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    const long long s0 = c * a.L;
    if (s0 >= a.n) return;
    const long long s1 = min(s0 + a.L, a.n);
    const float *x = reinterpret_cast<const float *>(a.in);
    float env;
    long long i = s0 - a.W;
    if (i <= 0) { i = 0; env = __ldcg(&a.carry_in->pad.x); }
    else env = fminf(fmaxf(__ldg(x + i), 0.0f), 1.0f);
    for (; i < s0; ++i) {
        const float tgt = fminf(fmaxf(__ldg(x + i), 0.0f), 1.0f);
        const float al = (tgt >= env) ? a.attack_a : a.release_a;
        env = al * env + (1.0f - al) * tgt;
    }
    unsigned ctr;
    float2 p = nco_exact_any(a.osc, s0);
    ctr = (unsigned)(a.osc.kbase + 1ull + (unsigned long long)s0);
    const float2 w = make_float2(a.osc.xwre, a.osc.xwim);
    for (; i < s1; ++i) {
        const float tgt = fminf(fmaxf(__ldg(x + i), 0.0f), 1.0f);                // cw.rs:52
        const float al = (tgt >= env) ? a.attack_a : a.release_a;
        env = al * env + (1.0f - al) * tgt;                                       // cw.rs:53-57
        const float xr = env * a.gain, xi = 0.0f;
        reinterpret_cast<float2 *>(a.out)[i] = make_float2(xr * p.x - xi * p.y, xr * p.y + xi * p.x);   // nco.rs:63-66
        nco_step_exact(p, w, ctr);
    }
    if (s1 == a.n) {
        CarryState cs = *a.carry_in;
        cs.pad.x = env;
        *a.carry_out = cs;
    }
}

cudaError_t agc_launch(const AgcArgs &a, cudaStream_t stream) {
    const long long nchunks = (a.n + a.L - 1) / a.L;
    const int threads = 128;
    const unsigned blocks = (unsigned)((nchunks + threads - 1) / threads);
    if (a.iq == 2) cw_mod_kernel<<<blocks, threads, 0, stream>>>(a);
    else if (a.iq) agc_kernel<true><<<blocks, threads, 0, stream>>>(a);
    else agc_kernel<false><<<blocks, threads, 0, stream>>>(a);
    return cudaGetLastError();
}

}  // namespace orion
