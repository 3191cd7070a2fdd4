// Instances for the rate-1 blocks: 16 items per lane, whole warp tiles moved through the transposing scratch.
#include "chain_kernels.cuh"
namespace orion {
chain_kernel_t get_kernel_direct(int dm) {
    switch (dm) {
        case DEMOD_NONE: return kptr<FRONT_DIRECT, 16, 1, 0, DEMOD_NONE>();                  // Rotator, NcoMixer
        case DM_LR4 + DEMOD_FM: return kptr<FRONT_DIRECT, 16, 1, 0, DM_LR4 + DEMOD_FM>();    // FmQuadratureDemod
        case DM_LR4 + DEMOD_PM: return kptr<FRONT_DIRECT, 16, 1, 0, DM_LR4 + DEMOD_PM>();    // PmQuadratureDemod
        case DM_LR4 + DEMOD_F32: return kptr<FRONT_DIRECT, 16, 1, 0, DM_LR4 + DEMOD_F32>();  // LpCascade
    }
    return kptr<FRONT_DIRECT, 16, 1>();
}
chain_kernel_t get_kernel_direct_batch(int dm) {
    switch (dm) {
        case DM_LR4 + DEMOD_FM: return kptr<FRONT_DIRECT, 16, 1, 0, DM_LR4 + DEMOD_FM, 1>();
        case DM_LR4 + DEMOD_PM: return kptr<FRONT_DIRECT, 16, 1, 0, DM_LR4 + DEMOD_PM, 1>();
    }
    return kptr<FRONT_DIRECT, 16, 1, 0, -1, 1>();
}
}  // namespace orion
