// kernel_api.hpp — the seam between the host translation unit (mcphylo_b200.cu: C ABI, planner) and the
// kernel translation units (walk_k2.cu .. walk_k6.cu, walk_generic.cu).  Each state count K is compiled in
// its own unit so that the units build in parallel and a host-side edit does not recompile any kernel.
// The host only ever sees this table of plain function pointers.
#pragma once
#include <cuda_runtime.h>

#include <cstddef>

namespace mcpdev {

struct WalkParams;

struct LaunchCfg {
    int device = 0;
    int K = 0;                 // state count (only the runtime-K kernel reads it)
    int grid = 0, block = 0;   // persistent CTAs, threads per CTA
    int cpt = 1;               // alignment columns per thread
    size_t smem = 0;           // dynamic shared memory per CTA
    cudaStream_t stream = nullptr;
    bool smem_scratch = false; // partials scratch in shared memory (small inputs)
    bool acc_global = false;   // gradient accumulator in global memory (very large trees)
    bool ring = false;         // gradient pass fetches its stored operands through the per-warp operand ring (cp.async.bulk)
    bool mma = false;          // K > 6: tile-cooperative FP64 tensor-core kernel instead of the runtime-K fallback
    double* mg = nullptr;      // runtime-K kernel only: per-(branch, rate) moment matrices of a model-gradient evaluation
    int mg_rep = 1;            // ... kept in mg_rep replicas mg_stride doubles apart (CTA c adds to replica c % mg_rep), folded afterwards
    long long mg_stride = 0;
};

// Every entry returns cudaSuccess or the error of the CUDA call that failed.
struct KernelTable {
    // depth-first walk (kernel_walk.cuh / kernel_generic.cuh)
    cudaError_t (*launch_walk)(const LaunchCfg&, const WalkParams&, bool dyn_model, bool null_last);
    cudaError_t (*occupancy_walk)(const LaunchCfg&, int* ctas_per_sm);
    // level-parallel small-tree kernel (kernel_levels.cuh); null for the runtime-K unit
    // dyn_inline != nullptr: the per-evaluation parameter block (n_inline doubles) travels in the kernel arguments
    cudaError_t (*launch_levels)(const LaunchCfg&, const WalkParams&, bool dyn_model, const double* dyn_inline, size_t n_inline);
    cudaError_t (*occupancy_levels)(const LaunchCfg&, int* ctas_per_sm);
    // fills this unit's constant-memory model slots (batches with several models, full-K kernels);
    // null for the runtime-K unit, which reads its tables from global memory
    cudaError_t (*upload_model)(const double* h_slots, size_t bytes, cudaStream_t stream);
};

const KernelTable* kernels_k2();
const KernelTable* kernels_k3();
const KernelTable* kernels_k4();
const KernelTable* kernels_k5();
const KernelTable* kernels_k6();
const KernelTable* kernels_generic();

}  // namespace mcpdev
