// smem_layout.cuh — shared-memory carve-ups of the walk kernels and the constants the host needs to
// size a launch (chunk length, accumulator placement).  Shared by the host translation unit and the
// kernel translation units.
#pragma once

namespace mcpdev {

constexpr int CH = 16;

// Where a CTA accumulates its branch-gradient sums (template parameter ACCG, decided on the host):
// in shared memory without atomics (below) whenever the per-branch accumulator fits next to the staging
// buffers without costing a resident CTA -- trees of up to WALK_ACC_SHARED_MAX_NODES nodes -- else
// (ACCG) directly in the CTA's accumulator row in global memory with fire-and-forget RED.ADD.F64 (the
// row stays in L2; any tree size, but the sums arrive in no fixed order).  The two are separate
// instantiations: with both paths in one kernel the K = 4 op loop spills again.
constexpr int WALK_ACC_SHARED_MAX_NODES = 4096;
inline bool walk_acc_global(int n_nodes, int mode /* -1 auto, 0 shared, 1 global */) {
    return mode < 0 ? n_nodes > WALK_ACC_SHARED_MAX_NODES : mode != 0;
}

// Shared-memory accumulator: the per-warp sums of a chunk's 2 * CH branch terms are parked here
// and folded into the accumulator by one thread per term after the chunk barrier, in fixed warp order
// -- no atomics (a shared fp64 atomic add is a compare-and-swap loop, ~10 instructions, 38 % retries
// with 8 warps on one address) and a run-to-run reproducible gradient.
//   [2 buffers][CH ops][2 children][8 warps] doubles, then [2][CH][2] branch ids
__host__ __device__ constexpr int walk_part_doubles(int chn) { return 2 * chn * 2 * 8; }
__host__ __device__ constexpr int walk_part_bytes(int chn) { return walk_part_doubles(chn) * 8 + 2 * chn * 2 * 4; }

// Operand ring of the gradient pass (template parameter RD of the walk, 0 = none): the stored child partials a
// warp will need are fetched RD entries ahead of their use by bulk asynchronous copies (one elected lane,
// cp.async.bulk + mbarrier) into a per-warp ring, so their DRAM latency never reaches the scoreboard.
//   [warps] x { [RD stages][columns per thread][32 lanes][K doubles], [RD] mbarriers }
// Kernels with a ring stage their per-op inputs WALK_RING_CH ops at a time instead of CH, which pays for the
// ring's shared memory.
#ifndef MCP_RING_DEPTH
#define MCP_RING_DEPTH 2
#endif
#ifndef MCP_RING_CH
#define MCP_RING_CH 8
#endif
// Flavour of the ring: 0 = bulk copies (cp.async.bulk + mbarrier, one lane per warp issues), 1 = every thread moves its
// own vectors with 16-byte cp.async copies (completion by cp.async.wait_group; no mbarrier, no elected lane).
// Measured (profiles/r2_ab_ring_flavour_*.json, r2_ab_ring_k2_*.json): equal at cfg4 / 1 M sites, the cp.async flavour
// 1.5 % faster on a 125 k-site shard and on cfg3, 3-4 % faster at K = 2 (no elected-lane issue path through uniform
// registers, no reloads of U / Uinv around the try_wait branch) -- it is the default.
#ifndef MCP_RING_LDGSTS
#define MCP_RING_LDGSTS 1
#endif
constexpr bool WALK_RING_LDGSTS = MCP_RING_LDGSTS != 0;
constexpr int WALK_RING_DEPTH = MCP_RING_DEPTH;
constexpr int WALK_RING_CH = MCP_RING_CH;
__host__ __device__ constexpr bool walk_ring_supported(int K) { return K == 2 || K == 4; }

template <int K, int CHN = CH>
struct WalkSmem {
    // dynamic shared memory carve-up (offsets in bytes)
    // branch-gradient accumulator of the CTA + parked per-warp sums (shared-accumulator kernels only:
    // callers pass want_grad && !ACCG)
    static __host__ __device__ size_t acc_bytes(int n_br, int shared_acc) {
        return shared_acc ? (((size_t)n_br * 8 + 15) & ~(size_t)15) + walk_part_bytes(CHN) : 0;
    }
    static __host__ __device__ size_t desc_bytes() { return 3 * CHN * 32; }
    static __host__ __device__ size_t e_bytes() { return 2 * CHN * 2 * 2 * K * 8; }   // (em1, de) per internal child
    // state codes of leaf children: [2 buffers][CHN ops][2 children][rows][tile sites]; a child slot holds two
    // rows (a cherry child of the gradient pass stages the rows of both leaves below it)
    static __host__ __device__ size_t code_bytes(int TS, int rows) { return (size_t)2 * CHN * 2 * rows * TS; }
    static __host__ __device__ size_t rec_bytes() { return 2 * CHN * 32; }
    // leaf children: P (and, in the gradient pass, dP) columns [(K+1)][K] of the child's branch
    static __host__ __device__ size_t tab_bytes() { return (size_t)2 * CHN * 2 * 2 * K * (K + 1) * 8; }
    static __host__ __device__ size_t total(int n_br, int shared_acc, int TS) {
        return acc_bytes(n_br, shared_acc) + desc_bytes() + e_bytes() + rec_bytes() + tab_bytes() + code_bytes(TS, 2);
    }
    // operand ring behind everything else (128-byte aligned offset from the start of dynamic shared memory)
    static __host__ __device__ size_t ring_offset(int n_br, int shared_acc, int TS) {
        return (total(n_br, shared_acc, TS) + 127) & ~(size_t)127;
    }
    static __host__ __device__ size_t ring_bytes(int warps, int depth, int cpt) {
        return (size_t)warps * depth * cpt * 32 * K * 8 + (size_t)warps * depth * 8;
    }
};

struct LevelSmem {
    // byte offsets into dynamic shared memory
    static __host__ __device__ size_t acc_bytes(int n_br, int want_grad) { return want_grad ? (((size_t)n_br * 8 + 127) & ~(size_t)127) : 0; }
    static __host__ __device__ size_t exp_bytes() { return 128; }
    static __host__ __device__ size_t code_bytes(int n_rows) { return (((size_t)n_rows * 32) + 127) & ~(size_t)127; }
    static __host__ __device__ size_t slot_bytes(int K) { return (size_t)32 * K * 8; }
    static __host__ __device__ size_t tab_bytes(int n_br, int K) { return (((size_t)n_br * bt_size(K) * 8) + 127) & ~(size_t)127; }
    // post + pre slots; the region doubles as the 4 KB scratch of the prior's block reduction at the very end
    static __host__ __device__ size_t slots_bytes(int n_slots, int n_stack, int K) {
        const size_t b = (size_t)(n_slots + n_stack) * slot_bytes(K);
        return b < 4096 ? 4096 : b;
    }
    // op program (post + pre: at most 2 ops of 32 bytes per device node) and level offsets of the tree
    static __host__ __device__ size_t prog_bytes(int n_br) { return (size_t)n_br * 64 + (((size_t)(2 * n_br + 4) * 4 + 15) & ~(size_t)15); }
    static __host__ __device__ size_t total(int n_br, int want_grad, int n_rows, int n_slots, int n_stack, int K) {
        return acc_bytes(n_br, want_grad) + exp_bytes() + code_bytes(n_rows) + tab_bytes(n_br, K) +
               slots_bytes(n_slots, n_stack, K) + prog_bytes(n_br);
    }
};

// Tile-cooperative walk for large alphabets (kernel_mma.cuh): 8 warps x 16 columns per CTA.
constexpr int MMA_WARPS = 8, MMA_MB = 2, MMA_WCOLS = 8 * MMA_MB, MMA_TILE = MMA_WARPS * MMA_WCOLS;   // 128 columns per CTA

template <int KP>
struct MmaSmem {
    static constexpr int STRIDE = KP + 2;
    static constexpr int TAB = (KP + 1) * STRIDE;                   // doubles of one staged table
    static __host__ __device__ size_t acc_bytes(int n_br, int want_grad) { return want_grad ? (((size_t)n_br * 8 + 15) & ~(size_t)15) : 0; }
    static __host__ __device__ size_t part_bytes() { return 2 * 2 * MMA_WARPS * 8 + 2 * 2 * 4; }    // parked sums + branch ids, 2 buffers
    static __host__ __device__ size_t desc_bytes() { return 3 * 32; }
    static __host__ __device__ size_t code_bytes() { return 2 * 2 * MMA_TILE; }
    // [2 buffers][4 tables] staged per op + one table per tile: Q' = mu * rate * Q (dP = Q' P), kernel_mma.cuh
    static __host__ __device__ size_t tab_bytes() { return (size_t)(2 * 4 + 1) * TAB * 8; }
    // gradient evaluations: the two stored child vectors of the NEXT family, fetched one op ahead with cp.async:
    // [8 warps][2 children][vector of 16 columns x KP states, fragment-major like the scratch]
    static __host__ __device__ size_t pref_bytes(int want_grad) { return want_grad ? (size_t)MMA_WARPS * 2 * MMA_WCOLS * KP * 8 : 0; }
    static __host__ __device__ size_t total(int n_br, int want_grad) {
        return acc_bytes(n_br, want_grad) + ((part_bytes() + 15) & ~(size_t)15) + desc_bytes() + code_bytes() + tab_bytes() +
               pref_bytes(want_grad);
    }
};

}  // namespace mcpdev
using namespace mcpdev;
