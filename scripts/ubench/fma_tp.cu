// FP32 FMA issue throughput per SM: FFMA (3 registers), FFMA with a uniform/constant operand, FFMA2.
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdlib>
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("%s: %s\n", #x, cudaGetErrorString(e)); exit(1); } } while (0)
typedef unsigned long long f32x2;
__device__ __forceinline__ f32x2 pack2(float a, float b) { f32x2 r; asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(a), "f"(b)); return r; }
__device__ __forceinline__ f32x2 ffma2(f32x2 a, f32x2 b, f32x2 c) { f32x2 d; asm volatile("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c)); return d; }
template <int MODE>
__global__ void k(float *out, int iters, float t0, float t1) {
    float a[16];
    f32x2 p[16];
    for (int i = 0; i < 16; ++i) { a[i] = threadIdx.x * 0.001f + i; p[i] = pack2(a[i], a[i] + 1.f); }
    float x = out[threadIdx.x & 7], y = out[8 + (threadIdx.x & 7)];
    const f32x2 tt0 = pack2(t0, t0), xy = pack2(x, y);
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int r = 0; r < 8; ++r) {
            if (MODE == 0) {           // FFMA, 3 vector registers
#pragma unroll
                for (int i = 0; i < 16; ++i) a[i] = fmaf(a[i], x, y);
            } else if (MODE == 1) {    // FFMA with a kernel-parameter (constant bank / uniform) multiplier
#pragma unroll
                for (int i = 0; i < 16; ++i) a[i] = fmaf(x, t0, a[i]);
            } else if (MODE == 2) {    // FFMA2, scalar-broadcast multiplier
#pragma unroll
                for (int i = 0; i < 16; ++i) p[i] = ffma2(tt0, xy, p[i]);
            } else {                   // FFMA2, three packed registers
#pragma unroll
                for (int i = 0; i < 16; ++i) p[i] = ffma2(p[i], xy, xy);
            }
        }
    }
    float s = 0.f;
    for (int i = 0; i < 16; ++i) { s += a[i]; s += __uint_as_float((unsigned)(p[i] & 0xffffffffu)); }
    if (s == 12345.678f) out[0] = s;
}
int main() {
    float *out; CK(cudaMalloc(&out, 1024)); CK(cudaMemset(out, 0, 1024));
    cudaEvent_t e0, e1; CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
    const int iters = 4096;
    const char *names[4] = {"FFMA reg,reg,reg", "FFMA reg,const,reg", "FFMA2 scalar-bcast", "FFMA2 packed regs"};
    for (int mode = 0; mode < 4; ++mode)
        for (int warps : {1, 2, 4, 8, 16, 32}) {
            auto launch = [&]() {
                if (mode == 0) k<0><<<148, warps * 32>>>(out, iters, 1.0001f, 0.5f);
                if (mode == 1) k<1><<<148, warps * 32>>>(out, iters, 1.0001f, 0.5f);
                if (mode == 2) k<2><<<148, warps * 32>>>(out, iters, 1.0001f, 0.5f);
                if (mode == 3) k<3><<<148, warps * 32>>>(out, iters, 1.0001f, 0.5f);
            };
            launch(); CK(cudaDeviceSynchronize());
            CK(cudaEventRecord(e0)); launch(); CK(cudaEventRecord(e1)); CK(cudaEventSynchronize(e1));
            float ms; CK(cudaEventElapsedTime(&ms, e0, e1));
            const double inst = (double)iters * 8 * 16 * warps;            // warp-instructions per SM
            const double lanes = inst * 32 * (mode >= 2 ? 2 : 1);
            printf("%-20s warps/SM %2d: %7.1f us  %6.2f warp-instr/ns/SM  %7.1f FMA-lanes/ns/SM  %6.1f TFLOP/s chip\n", names[mode], warps, ms * 1e3,
                   inst / (ms * 1e6), lanes / (ms * 1e6), lanes * 148 * 2 / (ms * 1e-3) / 1e12);
        }
    return 0;
}
