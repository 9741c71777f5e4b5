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
template <int KP, bool DST>
cudaError_t launch_mma_inst(const LaunchCfg& c, const WalkParams& wp) {
    cudaError_t e = raise_smem(felsenstein_walk_mma<KP, DST>, c.smem);
    if (e != cudaSuccess) return e;
    felsenstein_walk_mma<KP, DST><<<c.grid, MMA_WARPS * 32, c.smem, c.stream>>>(wp, c.K);
    return cudaGetLastError();
}
template <int KP>
cudaError_t launch_mma(const LaunchCfg& c, const WalkParams& wp) {
    return wp.want_grad ? launch_mma_inst<KP, true>(c, wp) : launch_mma_inst<KP, false>(c, wp);
}
template <int KP>
cudaError_t occupancy_mma(const LaunchCfg& c, int* out) {   // the persistent grid is sized for the more demanding of the two
    int o1 = 0, o2 = 0;
    cudaError_t e = raise_smem(felsenstein_walk_mma<KP, true>, c.smem);
    if (e == cudaSuccess) e = raise_smem(felsenstein_walk_mma<KP, false>, c.smem);
    if (e == cudaSuccess) e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&o1, felsenstein_walk_mma<KP, true>, MMA_WARPS * 32, c.smem);
    if (e == cudaSuccess) e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&o2, felsenstein_walk_mma<KP, false>, MMA_WARPS * 32, c.smem);
    *out = o1 < o2 ? o1 : o2;
    return e;
}
cudaError_t launch_generic(const LaunchCfg& c, const WalkParams& wp, bool, bool) {
    if (c.mma) {
        const int KP = (c.K + 7) & ~7;
        return KP == 8 ? launch_mma<8>(c, wp) : KP == 16 ? launch_mma<16>(c, wp) : KP == 24 ? launch_mma<24>(c, wp) : launch_mma<32>(c, wp);
    }
    cudaError_t e = raise_smem(felsenstein_walk_generic, c.smem);
    if (e != cudaSuccess) return e;
    felsenstein_walk_generic<<<c.grid, c.block, c.smem, c.stream>>>(wp, c.K, c.mg);
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
