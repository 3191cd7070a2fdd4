#include <cuda_runtime.h>
#include <cstdio>
__global__ void k(long long *o, int iters) {
    unsigned long long g0, g1; long long c0, c1;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(g0));
    c0 = clock64();
    float a = threadIdx.x;
    for (int i = 0; i < iters; ++i) a = fmaf(a, 1.0001f, 0.5f);
    c1 = clock64();
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(g1));
    if (threadIdx.x == 0 && blockIdx.x == 0) { o[0] = c1 - c0; o[1] = (long long)(g1 - g0); o[2] = (long long)a; }
}
int main() {
    long long *d, h[3];
    cudaMalloc(&d, 64);
    for (int rep = 0; rep < 4; ++rep) {
        k<<<148, 128>>>(d, 200000);
        cudaMemcpy(h, d, 24, cudaMemcpyDeviceToHost);
        printf("clock64 cycles %lld, globaltimer ns %lld -> %.3f GHz\n", h[0], h[1], (double)h[0] / h[1]);
    }
    return 0;
}
