// walk_generic.cu — the kernels for large alphabets (6 < K <= 32) as their own translation unit (kernel_api.hpp):
// the tile-cooperative FP64 tensor-core walk (kernel_mma.cuh) and the runtime-K fallback (kernel_generic.cuh).
#include "schedule.hpp"

#include <cuda_runtime.h>

#include <type_traits>

#include "device_layout.cuh"
#include "smem_layout.cuh"
#include "kernel_api.hpp"
#include "model_const.cuh"
#include "device_math.cuh"
#include "kernel_generic.cuh"
#include "kernel_mma.cuh"

namespace {

template <class Kern>
cudaError_t raise_smem(Kern kern, size_t smem) {
    return smem > 48 * 1024 ? cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) : cudaSuccess;
}
template <int KP>
cudaError_t launch_mma(const LaunchCfg& c, const WalkParams& wp) {
    cudaError_t e = raise_smem(felsenstein_walk_mma<KP>, c.smem);
    if (e != cudaSuccess) return e;
    felsenstein_walk_mma<KP><<<c.grid, MMA_WARPS * 32, c.smem, c.stream>>>(wp, c.K);
    return cudaGetLastError();
}
template <int KP>
cudaError_t occupancy_mma(const LaunchCfg& c, int* out) {
    cudaError_t e = raise_smem(felsenstein_walk_mma<KP>, c.smem);
    if (e != cudaSuccess) return e;
    return cudaOccupancyMaxActiveBlocksPerMultiprocessor(out, felsenstein_walk_mma<KP>, MMA_WARPS * 32, c.smem);
}
cudaError_t launch_generic(const LaunchCfg& c, const WalkParams& wp, bool, bool) {
    if (c.mma) {
        const int KP = (c.K + 7) & ~7;
        return KP == 8 ? launch_mma<8>(c, wp) : KP == 16 ? launch_mma<16>(c, wp) : KP == 24 ? launch_mma<24>(c, wp) : launch_mma<32>(c, wp);
    }
    cudaError_t e = raise_smem(felsenstein_walk_generic, c.smem);
    if (e != cudaSuccess) return e;
    felsenstein_walk_generic<<<c.grid, c.block, c.smem, c.stream>>>(wp, c.K);
    return cudaGetLastError();
}
cudaError_t occupancy_generic(const LaunchCfg& c, int* out) {
    if (c.mma) {
        const int KP = (c.K + 7) & ~7;
        return KP == 8 ? occupancy_mma<8>(c, out) : KP == 16 ? occupancy_mma<16>(c, out) : KP == 24 ? occupancy_mma<24>(c, out) : occupancy_mma<32>(c, out);
    }
    cudaError_t e = raise_smem(felsenstein_walk_generic, c.smem);
    if (e != cudaSuccess) return e;
    return cudaOccupancyMaxActiveBlocksPerMultiprocessor(out, felsenstein_walk_generic, c.block, c.smem);
}

}  // namespace

namespace mcpdev {
const KernelTable* kernels_generic() {
    static const KernelTable t{launch_generic, occupancy_generic, nullptr, nullptr, nullptr};
    return &t;
}
}  // namespace mcpdev
