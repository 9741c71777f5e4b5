// walk_inst.cuh — instantiates the walk kernels for ONE state count (MCP_INST_K, set by walk_k*.cu) and
// exports them to the host translation unit as a KernelTable (kernel_api.hpp).
#pragma once
#include "schedule.hpp"

#include <cuda_runtime.h>

#include <algorithm>
#include <cstring>
#include <mutex>
#include <vector>

#include "device_layout.cuh"
#include "smem_layout.cuh"
#include "kernel_api.hpp"
#include "model_const.cuh"
#include "device_math.cuh"
#include "kernel_walk.cuh"
#include "epilogue_prior.cuh"
#include "kernel_levels.cuh"

namespace {

#define MCP_CU(expr)                              \
    do {                                          \
        cudaError_t _e = (expr);                  \
        if (_e != cudaSuccess) return _e;         \
    } while (0)

// cudaFuncSetAttribute is only needed when a kernel's dynamic shared memory grows.  The attribute
// belongs to the (device, function) pair and is shared by every context of the process, so the
// high-water marks are process-global and only ever raised.
template <class Kern>
cudaError_t ensure_smem_attr(int device, Kern kern, size_t smem) {
    struct Mark { int device; const void* fn; size_t bytes; };
    static std::mutex mu;
    static std::vector<Mark> marks;
    std::lock_guard<std::mutex> lock(mu);
    for (auto& m : marks)
        if (m.device == device && m.fn == (const void*)kern) {
            if (m.bytes >= smem) return cudaSuccess;
            MCP_CU(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
            m.bytes = smem;
            return cudaSuccess;
        }
    MCP_CU(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    marks.push_back({device, (const void*)kern, smem});
    return cudaSuccess;
}

template <int K, int CPT, bool DYN, bool SSCR, int NE, bool ACCG, int RD = 0>
cudaError_t launch_walk_inst(const LaunchCfg& c, const WalkParams& wp) {
    MCP_CU(ensure_smem_attr(c.device, felsenstein_walk<K, CPT, DYN, SSCR, NE, ACCG, RD>, c.smem));
    felsenstein_walk<K, CPT, DYN, SSCR, NE, ACCG, RD><<<c.grid, c.block, c.smem, c.stream>>>(wp);
    return cudaGetLastError();
}
// null_last: every model of the batch has a null eigenvalue, moved to the last position by the host
// (true for any rate matrix) -> kernels with K - 1 active eigen-components.  Otherwise (a caller
// passing some other decomposition) the full-K kernels.  Those, and the kernels that accumulate the
// gradient in global memory (acc_global: very large trees), exist in the constant-memory (DYN)
// flavour only -- and the latter with one column per thread only, which the planner arranges.
template <int K>
cudaError_t launch_walk_k(const LaunchCfg& c, const WalkParams& wp, bool dyn_model, bool null_last) {
    constexpr int NE = K - 1;
    if constexpr (walk_ring_supported(K)) {
        // operand ring (host: only for one shared model with a null eigenvalue, shared-memory accumulator, scratch in HBM)
        if (c.ring && !c.acc_global && !c.smem_scratch && null_last && !dyn_model)
            return c.cpt == 2 ? launch_walk_inst<K, 2, false, false, NE, false, WALK_RING_DEPTH>(c, wp)
                              : launch_walk_inst<K, 1, false, false, NE, false, WALK_RING_DEPTH>(c, wp);
    }
    if (c.ring) return cudaErrorInvalidValue;
    if (c.acc_global)
        return null_last ? launch_walk_inst<K, 1, true, false, NE, true>(c, wp) : launch_walk_inst<K, 1, true, false, K, true>(c, wp);
    if (!null_last)
        return c.cpt == 2 ? launch_walk_inst<K, 2, true, false, K, false>(c, wp) : launch_walk_inst<K, 1, true, false, K, false>(c, wp);
    if (c.smem_scratch && !dyn_model && c.cpt == 1) return launch_walk_inst<K, 1, false, true, NE, false>(c, wp);
    if (c.cpt == 2) return dyn_model ? launch_walk_inst<K, 2, true, false, NE, false>(c, wp) : launch_walk_inst<K, 2, false, false, NE, false>(c, wp);
    return dyn_model ? launch_walk_inst<K, 1, true, false, NE, false>(c, wp) : launch_walk_inst<K, 1, false, false, NE, false>(c, wp);
}

template <class Kern>
cudaError_t occ_of(const LaunchCfg& c, Kern kern, int* out) {
    MCP_CU(ensure_smem_attr(c.device, kern, c.smem));
    return cudaOccupancyMaxActiveBlocksPerMultiprocessor(out, kern, c.block, c.smem);
}
template <int K, int CPT>
cudaError_t occupancy_inst(const LaunchCfg& c, int* out) {
    constexpr int NE = K - 1;
    if (c.ring) {
        if constexpr (walk_ring_supported(K)) return occ_of(c, felsenstein_walk<K, CPT, false, false, NE, false, WALK_RING_DEPTH>, out);
        else return cudaErrorInvalidValue;
    }
    if (c.acc_global) {
        int o1 = 0, o2 = 0;
        MCP_CU(occ_of(c, felsenstein_walk<K, 1, true, false, K, true>, &o1));
        MCP_CU(occ_of(c, felsenstein_walk<K, 1, true, false, NE, true>, &o2));
        *out = std::min(o1, o2);
        return cudaSuccess;
    }
    if (c.smem_scratch) return occ_of(c, felsenstein_walk<K, 1, false, true, NE, false>, out);
    // the variants differ by a few registers: size the persistent grid for the most demanding one
    int o1 = 0, o2 = 0, o3 = 0;
    MCP_CU(occ_of(c, felsenstein_walk<K, CPT, true, false, NE, false>, &o1));
    MCP_CU(occ_of(c, felsenstein_walk<K, CPT, false, false, NE, false>, &o2));
    MCP_CU(occ_of(c, felsenstein_walk<K, CPT, true, false, K, false>, &o3));
    *out = std::min(o1, std::min(o2, o3));
    return cudaSuccess;
}
template <int K>
cudaError_t occupancy_walk_k(const LaunchCfg& c, int* out) {
    if (c.cpt == 2) {
        LaunchCfg c2 = c;
        c2.smem_scratch = false;
        return occupancy_inst<K, 2>(c2, out);
    }
    return occupancy_inst<K, 1>(c, out);
}
template <int K>
cudaError_t occupancy_levels_k(const LaunchCfg& c, int* out) {
    int o1 = 0, o2 = 0;
    MCP_CU(occ_of(c, felsenstein_walk_levels<K, true>, &o1));
    MCP_CU(occ_of(c, felsenstein_walk_levels<K, false>, &o2));
    *out = std::min(o1, o2);
    return cudaSuccess;
}
template <int K>
cudaError_t launch_levels_k(const LaunchCfg& c, const WalkParams& wp, bool dyn_model, const double* dyn_inline, size_t n_inline) {
    static thread_local LevelParams lp;            // 5 KB: not on the stack of whatever thread calls in
    lp.w = wp;
    if (dyn_inline) {
        if (n_inline > (size_t)LEVEL_DYN_INLINE) return cudaErrorInvalidValue;
        std::memcpy(lp.dyn_inline, dyn_inline, n_inline * sizeof(double));
        lp.w.dyn = nullptr;
    }
    if (dyn_model) felsenstein_walk_levels<K, true><<<c.grid, c.block, c.smem, c.stream>>>(lp);
    else felsenstein_walk_levels<K, false><<<c.grid, c.block, c.smem, c.stream>>>(lp);
    return cudaGetLastError();
}
cudaError_t upload_model_slots(const double* h_slots, size_t bytes, cudaStream_t stream) {
    return cudaMemcpyToSymbolAsync(c_model, h_slots, bytes, 0, cudaMemcpyHostToDevice, stream);
}

}  // namespace

#define MCP_DEFINE_KERNEL_TABLE(K)                                                                          \
    namespace mcpdev {                                                                                      \
    const KernelTable* kernels_k##K() {                                                                     \
        static const KernelTable t{launch_walk_k<K>, occupancy_walk_k<K>, launch_levels_k<K>,               \
                                   occupancy_levels_k<K>, upload_model_slots};                              \
        return &t;                                                                                          \
    }                                                                                                       \
    }
