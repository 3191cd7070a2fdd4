// Instances for the fixed geometries (decimate-by-8 C1 / C3 shape, decimate-by-32 C4 shape) and the any-shape global-memory FIR.
#include "chain_kernels.cuh"
namespace orion {
chain_kernel_t get_kernel_hot(int front, int sp, int dm) {
    if (front == FRONT_GLOBAL) return kptr<FRONT_GLOBAL, 8, 1>();
    if (sp == 2) return kptr<FRONT_STAGED, 4, 1, 2, DEMOD_NONE>();                               // C4: FIR1023 / 32 alone
    if (dm == DEMOD_NONE) return kptr<FRONT_STAGED, 8, 1, 1, DEMOD_NONE>();                      // FirDecimator alone
    if (dm == DM_FM_LR4) return kptr<FRONT_STAGED, 8, 1, 1, DM_FM_LR4>();                        // the C1 chain
    if (dm == DEMOD_AM) return kptr<FRONT_STAGED, 8, 1, 1, DEMOD_AM>();                          // C3: sections stay generic
    return kptr<FRONT_STAGED, 8, 1, 1, -1>();
}
}  // namespace orion
