// Generic staged instances, odd decimation factors (U = 2: two outputs per block so that blocks hold an even number of samples).
#include "chain_kernels.cuh"
namespace orion {
chain_kernel_t get_kernel_staged_u2(int R) {
    switch (R) {
        case 8: return kptr<FRONT_STAGED, 8, 2>();
        case 4: return kptr<FRONT_STAGED, 4, 2>();
        case 2: return kptr<FRONT_STAGED, 2, 2>();
        case 1: return kptr<FRONT_STAGED, 1, 2>();
    }
    return nullptr;
}
}  // namespace orion
