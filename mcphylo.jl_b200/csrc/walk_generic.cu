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
// state counts 2 .. 6 reach this unit only for model-gradient evaluations: compile-time K (kernel_generic.cuh)
template <int KT>
cudaError_t launch_generic_inst(const LaunchCfg& c, const WalkParams& wp) {
    cudaError_t e = raise_smem(felsenstein_walk_generic<KT>, c.smem);
    if (e != cudaSuccess) return e;
    felsenstein_walk_generic<KT><<<c.grid, c.block, c.smem, c.stream>>>(wp, c.K, c.mg, c.mg_rep, c.mg_stride);
    return cudaGetLastError();
}
template <int KT>
cudaError_t occupancy_generic_inst(const LaunchCfg& c, int* out) {
    cudaError_t e = raise_smem(felsenstein_walk_generic<KT>, c.smem);
    if (e != cudaSuccess) return e;
    return cudaOccupancyMaxActiveBlocksPerMultiprocessor(out, felsenstein_walk_generic<KT>, c.block, c.smem);
}
cudaError_t launch_generic(const LaunchCfg& c, const WalkParams& wp, bool, bool) {
    if (c.mma) {
        const int KP = (c.K + 7) & ~7;
        return KP == 8 ? launch_mma<8>(c, wp) : KP == 16 ? launch_mma<16>(c, wp) : KP == 24 ? launch_mma<24>(c, wp) : launch_mma<32>(c, wp);
    }
    switch (c.K) {
        case 2: return launch_generic_inst<2>(c, wp);
        case 3: return launch_generic_inst<3>(c, wp);
        case 4: return launch_generic_inst<4>(c, wp);
        case 5: return launch_generic_inst<5>(c, wp);
        case 6: return launch_generic_inst<6>(c, wp);
        default: return launch_generic_inst<0>(c, wp);
    }
}
cudaError_t occupancy_generic(const LaunchCfg& c, int* out) {
    if (c.mma) {
        const int KP = (c.K + 7) & ~7;
        return KP == 8 ? occupancy_mma<8>(c, out) : KP == 16 ? occupancy_mma<16>(c, out) : KP == 24 ? occupancy_mma<24>(c, out) : occupancy_mma<32>(c, out);
    }
    switch (c.K) {
        case 2: return occupancy_generic_inst<2>(c, out);
        case 3: return occupancy_generic_inst<3>(c, out);
        case 4: return occupancy_generic_inst<4>(c, out);
        case 5: return occupancy_generic_inst<5>(c, out);
        case 6: return occupancy_generic_inst<6>(c, out);
        default: return occupancy_generic_inst<0>(c, out);
    }
}

}  // namespace

namespace mcpdev {
const KernelTable* kernels_generic() {
    static const KernelTable t{launch_generic, occupancy_generic, nullptr, nullptr, nullptr};
    return &t;
}
}  // namespace mcpdev
