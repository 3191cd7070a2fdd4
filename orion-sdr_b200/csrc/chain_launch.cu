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
// (ctr, ctr + nsteps].
__global__ void osc_expand_kernel(const OscAnchor *an, int n_an, float2 *fine, unsigned long long c0, long long fine_len) {
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= n_an) return;
    const OscAnchor A = an[t];
    float2 z = A.z;
    unsigned ctr = (unsigned)A.ctr;
    unsigned long long c = A.ctr;
    for (unsigned s = 0; s < A.nsteps; ++s) {
        nco_step_exact(z, A.w, ctr);
        c += 1ull;
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
