// Complex-valued trace (anisotropic media): placeholder until the crystal
// kernels land; reports PYR_E_UNSUPPORTED loudly instead of falling back.
#include "pyr_device.cuh"

namespace pyr {
int trace_complex(const PyrStep *, int32_t, const PyrRaysIn *, int64_t, uint32_t, cudaStream_t) {
    return PYR_E_UNSUPPORTED;
}
}  // namespace pyr
