// Generic staged instances, even decimation factors (U = 1).
#include "chain_kernels.cuh"
namespace orion {
chain_kernel_t get_kernel_staged_u1(int R) {
    switch (R) {
        case 8: return kptr<FRONT_STAGED, 8, 1>();
        case 4: return kptr<FRONT_STAGED, 4, 1>();
        case 2: return kptr<FRONT_STAGED, 2, 1>();
        case 1: return kptr<FRONT_STAGED, 1, 1>();
    }
    return nullptr;
}
}  // namespace orion
