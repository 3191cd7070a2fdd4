// Streaming-read bandwidth of the staging options for the chain kernel's input tiles.
//   mode 0: LDG.128 grid-stride read
//   mode 1: TMA 2D tensor load, box 66 x 33 of 8-byte elements over a [rows][64] tensor (zero-filled pad)
//   mode 2: one 1D bulk copy of 33*512 B per tile (dense layout)
//   mode 3: 33 1D bulk copies of 512 B per tile into 528-byte-pitch rows
//   mode 4: TMA 2D, box 64 x 33 (no pad)
// Every warp owns one slot of (33*528 rounded to 128) bytes and loops over its tiles.
#include <cuda.h>
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("%s: %s\n", #x, cudaGetErrorString(e)); exit(1); } } while (0)
__device__ __forceinline__ uint32_t s32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint32_t b, uint32_t c) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(b), "r"(c) : "memory"); }
__device__ __forceinline__ void mbar_expect(uint32_t b, uint32_t n) { asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(b), "r"(n) : "memory"); }
__device__ __forceinline__ bool mbar_try(uint32_t b, uint32_t par) {
    uint32_t ok;
    asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(ok) : "r"(b), "r"(par) : "memory");
    return ok != 0;
}
__device__ __forceinline__ void tma2d(uint32_t dst, const CUtensorMap *m, int c0, int c1, uint32_t bar) {
    asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];" ::"r"(dst), "l"(m), "r"(c0), "r"(c1), "r"(bar) : "memory");
}
__device__ __forceinline__ void bulk1d(uint32_t dst, const void *src, uint32_t bytes, uint32_t bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst), "l"(src), "r"(bytes), "r"(bar) : "memory");
}
constexpr int SLOT = 17536;
__global__ void k_ldg(const float4 *in, size_t n16, float *out) {
    float acc = 0.f;
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n16; i += (size_t)gridDim.x * blockDim.x) {
        float4 v = __ldg(in + i);
        acc += v.x + v.w;
    }
    if (acc == 1234.5f) out[0] = acc;
}
__global__ void k_stage(const __grid_constant__ CUtensorMap tm, const char *in, long long ntiles, int mode, int depth, float *out) {
    extern __shared__ __align__(1024) unsigned char smem[];
    __shared__ unsigned long long bars[64];
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5, nw = blockDim.x >> 5;
    // each warp owns `depth` slots used as a ring
    unsigned char *base = smem + (size_t)wid * depth * SLOT;
    if (lane == 0) for (int d = 0; d < depth; ++d) mbar_init(s32(&bars[wid * depth + d]), 1);
    __syncwarp();
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    const long long W = (long long)gridDim.x * nw, me = (long long)blockIdx.x * nw + wid;
    auto issue = [&](long long t, int d) {
        uint32_t bar = s32(&bars[wid * depth + d]);
        uint32_t dst = s32(base + (size_t)d * SLOT);
        if (mode == 1) { if (lane == 0) { mbar_expect(bar, 33 * 528); tma2d(dst, &tm, 0, (int)(t * 32), bar); } }
        else if (mode == 4) { if (lane == 0) { mbar_expect(bar, 33 * 512); tma2d(dst, &tm, 0, (int)(t * 32), bar); } }
        else if (mode == 2) { if (lane == 0) { mbar_expect(bar, 33 * 512); bulk1d(dst, in + t * 32 * 512, 33 * 512, bar); } }
        else if (mode == 3) {
            if (lane == 0) mbar_expect(bar, 33 * 512);
            __syncwarp();
            bulk1d(dst + lane * 528, in + (t * 32 + lane) * 512, 512, bar);
            if (lane == 0) bulk1d(dst + 32 * 528, in + (t * 32 + 32) * 512, 512, bar);
        }
    };
    float acc = 0.f;
    long long t = me; int i = 0;
    for (; i < depth && t < ntiles; ++i, t += W) issue(t, i);
    long long tc = me; int u = 0;
    for (; tc < ntiles; tc += W, ++u) {
        const int d = u % depth; const unsigned par = (u / depth) & 1u;
        uint32_t bar = s32(&bars[wid * depth + d]);
        while (!mbar_try(bar, par)) {}
        acc += *reinterpret_cast<float *>(base + (size_t)d * SLOT + lane * 528);
        __syncwarp();
        if (t < ntiles) { issue(t, d); t += W; }
    }
    if (acc == 1234.5f) out[0] = acc;
}
typedef CUresult (*enc_t)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *, const cuuint64_t *, const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
int main() {
    const size_t bytes = 192ull << 20;      // 192 MiB
    const long long rows = bytes / 512, ntiles = (rows - 33) / 32;
    char *in[3]; float *out;
    for (int i = 0; i < 3; ++i) { CK(cudaMalloc(&in[i], bytes + 4096)); CK(cudaMemset(in[i], 1, bytes)); }
    CK(cudaMalloc(&out, 64));
    void *p = nullptr; cudaDriverEntryPointQueryResult q;
    CK(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q));
    enc_t enc = (enc_t)p;
    cudaEvent_t e0, e1; CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
    auto time_it = [&](const char *name, auto launch) {
        for (int i = 0; i < 3; ++i) launch(i % 3);
        CK(cudaDeviceSynchronize());
        CK(cudaEventRecord(e0));
        for (int i = 0; i < 9; ++i) launch(i % 3);
        CK(cudaEventRecord(e1)); CK(cudaEventSynchronize(e1));
        float ms; CK(cudaEventElapsedTime(&ms, e0, e1)); ms /= 9;
        printf("%-46s %8.1f us  %8.1f GB/s\n", name, ms * 1e3, bytes / ms / 1e6); fflush(stdout);
    };
    time_it("LDG.128 grid-stride 148x8x256", [&](int b) { k_ldg<<<148 * 8, 256>>>((const float4 *)in[b], bytes / 16, out); });
    time_it("LDG.128 grid-stride 148x4x512", [&](int b) { k_ldg<<<148 * 4, 512>>>((const float4 *)in[b], bytes / 16, out); });
    CK(cudaFuncSetAttribute(k_stage, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
    for (int mode : {1, 4, 2, 3}) {
        CUtensorMap tms[3];
        for (int b = 0; b < 3; ++b) {
            cuuint64_t gdim[2] = {64, (cuuint64_t)rows}; cuuint64_t gstr[1] = {512};
            cuuint32_t box[2] = {(cuuint32_t)(mode == 1 ? 66 : 64), 33}; cuuint32_t es[2] = {1, 1};
            CUresult r = enc(&tms[b], CU_TENSOR_MAP_DATA_TYPE_FLOAT64, 2, in[b], gdim, gstr, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                             CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
            if (r != CUDA_SUCCESS) { printf("encode failed %d\n", (int)r); return 1; }
        }
        for (int cfg = 0; cfg < 4; ++cfg) {
            const int warps = (cfg == 0) ? 5 : (cfg == 1) ? 10 : (cfg == 2) ? 5 : 2, depth = (cfg == 2) ? 2 : (cfg == 3) ? 5 : 1;
            const int ctas = (cfg == 3) ? 1 : 1;
            char name[96]; snprintf(name, sizeof(name), "mode %d: %d warps x %d slots per CTA, %d CTA/SM", mode, warps, depth, ctas);
            const size_t sm = (size_t)warps * depth * SLOT;
            time_it(name, [&](int b) { k_stage<<<148 * ctas, warps * 32, sm>>>(tms[b], in[b], ntiles, mode, depth, out); });
        }
    }
    return 0;
}
