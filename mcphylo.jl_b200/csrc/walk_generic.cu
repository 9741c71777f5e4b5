// walk_generic.cu — the runtime-K walk kernel (6 < K <= 32) as its own translation unit (kernel_api.hpp).
#include "schedule.hpp"

#include <cuda_runtime.h>

#include "device_layout.cuh"
#include "smem_layout.cuh"
#include "kernel_api.hpp"
#include "model_const.cuh"
#include "device_math.cuh"
#include "kernel_generic.cuh"

namespace {

cudaError_t ensure_generic_smem(size_t smem) {
    return cudaFuncSetAttribute(felsenstein_walk_generic, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
}
cudaError_t launch_generic(const LaunchCfg& c, const WalkParams& wp, bool, bool) {
    if (c.smem > 48 * 1024) {
        cudaError_t e = ensure_generic_smem(c.smem);
        if (e != cudaSuccess) return e;
    }
    felsenstein_walk_generic<<<c.grid, c.block, c.smem, c.stream>>>(wp, c.K);
    return cudaGetLastError();
}
cudaError_t occupancy_generic(const LaunchCfg& c, int* out) {
    if (c.smem > 48 * 1024) {
        cudaError_t e = ensure_generic_smem(c.smem);
        if (e != cudaSuccess) return e;
    }
    return cudaOccupancyMaxActiveBlocksPerMultiprocessor(out, felsenstein_walk_generic, c.block, c.smem);
}

}  // namespace

namespace mcpdev {
const KernelTable* kernels_generic() {
    static const KernelTable t{launch_generic, occupancy_generic, nullptr, nullptr, nullptr};
    return &t;
}
}  // namespace mcpdev
