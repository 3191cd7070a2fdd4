// Kernel selection, occupancy set-up and launch of the chain kernel family (instances: chain_inst_*.cu).
#include "chain_kernels.cuh"

#include <cuda.h>
#include <cuda_runtime.h>
#include <string.h>

namespace orion {

typedef void (*chain_kernel_t)(const ChainArgs, const CUtensorMap);
chain_kernel_t get_kernel_hot(int front, int sp, int dm);
chain_kernel_t get_kernel_direct(int dm);
chain_kernel_t get_kernel_direct_batch(int dm);
chain_kernel_t get_kernel_staged_u1(int R);
chain_kernel_t get_kernel_staged_u2(int R);

// sp: 1 = the staged geometry is the fixed decimate-by-8 shape (Geo<1>); dm: -1 generic, a DEMOD_* kind, DM_LR4 + kind
chain_kernel_t select_kernel(int front, int R, int U, int sp, int dm, int batch) {
    if (batch) return front == FRONT_DIRECT ? get_kernel_direct_batch(dm) : nullptr;
    if (front == FRONT_DIRECT) return get_kernel_direct(dm);
    if (front == FRONT_GLOBAL) return get_kernel_hot(front, 0, dm);
    if (sp == 1 && R == 8 && U == 1) return get_kernel_hot(front, 1, dm);
    if (sp == 2 && R == 4 && U == 1 && dm == DEMOD_NONE) return get_kernel_hot(front, 2, dm);
    if (U == 1) return get_kernel_staged_u1(R);
    if (U == 2) return get_kernel_staged_u2(R);
    return nullptr;
}

// Checkpoint expansion of the exact-replay oscillator: thread t starts at anchor t = (ctr, Z(ctr), w), replays the
// reference recurrence (renormalisation included) and writes fine[e] = Z(c0 + 16 e) for every such counter value in
// (ctr, ctr + nsteps].  The walk is cut into runs that end at the next event (a table entry every 16 counter values, the
// renormalisation every 1024): inside a run there is nothing but the dependent multiply-add chain of the recurrence
// itself (8 cycles per step instead of the ~85 of a loop that tests both events after every step).
__global__ void __launch_bounds__(64)
osc_expand_kernel(const OscAnchor *an, int n_an, float2 *fine, unsigned long long c0, long long fine_len) {
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= n_an) return;
    const OscAnchor A = an[t];
    float2 z = A.z;
    const float2 w = A.w;
    unsigned long long c = A.ctr;
    unsigned left = A.nsteps;
    while (left > 0u) {
        const unsigned to_renorm = 1024u - (unsigned)(c & 1023ull);                       // 1 .. 1024
        const unsigned long long to_write = c < c0 ? c0 - c : 16ull - ((c - c0) & 15ull);   // >= 1
        unsigned run = left < to_renorm ? left : to_renorm;
        if ((unsigned long long)run > to_write) run = (unsigned)to_write;
        if (run == 16u) {
#pragma unroll
            for (int i = 0; i < 16; ++i) z = cmul_fma(z, w);                              // rotator.rs:46-48
        } else {
            for (unsigned i = 0; i < run; ++i) z = cmul_fma(z, w);
        }
        c += run;
        left -= run;
        if ((c & 1023ull) == 0ull) {                                                      // rotator.rs:51-59: r2.sqrt().recip()
            const float r2 = z.x * z.x + z.y * z.y;
            const float inv = 1.0f / sqrtf(r2);
            z.x *= inv;
            z.y *= inv;
        }
        if (c >= c0 && ((c - c0) & 15ull) == 0ull) {
            const long long e = (long long)((c - c0) >> 4);
            if (e < fine_len) fine[e] = z;
        }
    }
}
cudaError_t osc_expand_launch(const OscAnchor *d_an, int n_an, float2 *d_fine, unsigned long long c0, long long fine_len,
                              cudaStream_t stream) {
    if (n_an <= 0) return cudaSuccess;
    osc_expand_kernel<<<(n_an + 63) / 64, 64, 0, stream>>>(d_an, n_an, d_fine, c0, fine_len);
    return cudaGetLastError();
}

cudaError_t chain_kernel_prepare(chain_kernel_t k, size_t dyn_smem, int warps, int *ctas_per_sm) {
    // the opt-in limit is a property of the FUNCTION, and blocks of different shapes share one instance: always raise it to
    // the device maximum (a per-block value would be overwritten by the next block that is created)
    int dev = 0, optin = 227 * 1024;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&optin, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev);
    cudaFuncAttributes fa;
    cudaError_t e = cudaFuncGetAttributes(&fa, (const void *)k);
    if (e != cudaSuccess) return e;
    optin -= (int)fa.sharedSizeBytes;                        // static shared memory counts against the same limit
    if (dyn_smem > (size_t)optin) return cudaErrorInvalidValue;
    e = cudaFuncSetAttribute((const void *)k, cudaFuncAttributeMaxDynamicSharedMemorySize, optin);
    if (e != cudaSuccess) return e;
    e = cudaFuncSetAttribute((const void *)k, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
    if (e != cudaSuccess) return e;
    return cudaOccupancyMaxActiveBlocksPerMultiprocessor(ctas_per_sm, (const void *)k, kThreads * warps, dyn_smem);
}

cudaError_t chain_kernel_launch(chain_kernel_t k, const ChainArgs &args, const CUtensorMap &tmap, int grid, int warps,
                                size_t dyn_smem, cudaStream_t stream, int overlap) {
    // serial debug mode: one warp walks the tiles in order
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(grid, args.batch > 1 ? args.batch : 1);
    cfg.blockDim = dim3(args.serial ? kThreads : kThreads * warps);
    cfg.dynamicSmemBytes = dyn_smem;
    cfg.stream = stream;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr;
    cfg.numAttrs = overlap ? 1 : 0;
    return cudaLaunchKernelEx(&cfg, k, args, tmap);
}

}  // namespace orion
