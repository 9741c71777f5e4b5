// device_layout.cuh — device-side data layout: tree / walk parameter blocks, per-evaluation parameter layout, branch-table layout, model constants.
// Part of libmcphylo_b200.so; shared by the host translation unit and the per-K kernel translation units
// (named namespace: these types cross translation-unit boundaries through kernel_api.hpp).
#pragma once

namespace mcpdev {

using mcp::PostOp;
using mcp::PreOp;
using mcp::Schedule;

// --------------------------------------------------------------------------------------------
// device-side descriptors
// --------------------------------------------------------------------------------------------
struct TreeDev {
    long long post_off;      // op index (32-byte units) of this tree's post program
    long long pre_off;       // ... pre program
    long long btab_off;      // doubles, into the branch-table buffer
    long long dyn_off;       // doubles, into the per-evaluation parameter buffer
    long long out_off;       // doubles, into the result buffer ([logL, grad(NN-1)] per tree)
    const unsigned char* codes;  // (rows, code_stride) state codes of this tree's alignment
    long long S;             // sites
    long long code_stride;
    int n_post, n_pre;
    int NN, n_br;            // real nodes; rows of the branch table (device nodes)
    int tile_begin, tiles_per_rate;
    int row_lo, row_hi;      // accumulator rows [lo, hi) holding this tree's partial sums
    int lvl_off, n_post_lvl, n_pre_lvl;   // level-ordered program: offsets into WalkParams::levels
    int n_rows;              // leaf rows of the alignment
    int fetch_off, n_fetch;  // operand ring of the gradient pass: this tree's fetch list, WalkParams::fetch[fetch_off .. + n_fetch)
};

struct LLRow {
    long long esum;  // sum of binary exponents removed by rescaling (exact)
    double logsum;   // sum of log(pi . L_root)
};

struct WalkParams {
    const TreeDev* trees;
    const int4* ops;
    const double* btab;
    const double* dyn;
    double* scratch;
    long long scratch_per_cta;  // doubles
    double* rows;               // [row][row_stride] gradient partial sums
    LLRow* rows_ll;
    const int* cta_row_base;
    const int* levels;          // level offsets of the level-ordered programs (small-tree kernel)
    const unsigned short* fetch; // fetch lists of the operand ring (post slots in the order the gradient pass reads them)
    double* out;                // small-tree kernel: [logL, grad] per tree, device or pinned host memory
    unsigned int* done_counter; // small-tree kernel: CTAs finished (the last one reduces the rows)
    long long row_stride;
    int n_slots, n_stack;
    int n_tiles, T, R, want_grad;
    int max_br;
    int max_rows;               // largest number of leaf rows in the batch
    // Streamed evaluation (mcp_eval_streamed, one tree): the walk is launched BEFORE the alignment has arrived.
    // Tiles are handed out by an atomic ticket in site order (all rate categories of a site tile are
    // consecutive tickets), and a CTA starts a tile once the copy stream has marked the tile's sites as
    // landed: ready_flags[site >> ready_shift] == ready_epoch.  All null for a resident alignment
    // (static contiguous tile ranges, rate-major tile order).
    const unsigned* ready_flags;
    unsigned* ticket;
    unsigned* error_flag;       // set when a tile's data did not arrive within the spin limit
    unsigned ready_epoch;
    int ready_shift;
    // Substitution-model constants when the whole batch shares ONE model (the common case): kernel
    // parameters live in constant bank 0, so they reach the FP64 pipe as uniform operands without
    // a separate host-to-device copy.  Layout as in c_model (below).
    double model[176];
};

// Kernel parameters of the level-parallel small-tree kernel: the walk parameters plus, for MCMC-sized inputs, the
// whole per-evaluation parameter block (branch lengths, priors: two doubles per node) INLINE -- kernel arguments are
// delivered with the launch (CUDA 12.1+: up to 32 KB), so such an evaluation is one launch with no staging kernel and
// no read of pinned host memory in front of it.  w.dyn == nullptr selects the inline copy.
constexpr int LEVEL_DYN_INLINE = 448;           // doubles (3.5 KB): a tree of up to ~210 nodes, or the NNI pair of cfg2's 99-node tree
struct LevelParams {
    WalkParams w;
    double dyn_inline[LEVEL_DYN_INLINE];
};

// per-tree layout of the per-evaluation parameter block (offsets in doubles from dyn_off)
__host__ __device__ inline long long dyn_blv(int) { return 0; }
__host__ __device__ inline long long dyn_U(int NN) { return NN - 1; }
__host__ __device__ inline long long dyn_D(int NN, int K) { return NN - 1 + (long long)K * K; }
__host__ __device__ inline long long dyn_Uinv(int NN, int K) { return NN - 1 + (long long)K * K + K; }
__host__ __device__ inline long long dyn_mu(int NN, int K) { return NN - 1 + 2LL * K * K + K; }
__host__ __device__ inline long long dyn_rates(int NN, int K) { return NN + 2LL * K * K + K; }
__host__ __device__ inline long long dyn_pi(int NN, int K, int R) { return NN + 2LL * K * K + K + R; }
__host__ __device__ inline long long dyn_slot(int NN, int K, int R) { return NN + 2LL * K * K + 2LL * K + R; }
// branch-length prior (fused posterior epilogue, mcp_eval_posterior): 4 header doubles
// [enabled, c0, beta, k4] and NN-1 per-branch weights w_j, for the prior written as
//   log p(t) = c0 - beta * T + sum_j w_j log t_j + k4 log T,   T = sum_j t_j
__host__ __device__ inline long long dyn_prior(int NN, int K, int R) { return NN + 2LL * K * K + 2LL * K + R + 1; }
__host__ __device__ inline long long dyn_size(int NN, int K, int R) {
    long long n = dyn_prior(NN, K, R) + 4 + (NN - 1);
    return (n + 3) & ~3LL;  // keep every tree's block 32-byte aligned
}
// Branch table, one entry per (device branch, rate category), BT(K) doubles:
//   [0, K)                    em1_i = expm1(mu * t * D_i * rate) (internal children: P = I + U diag(em1) Uinv)
//   [K, 2K)                   de_i = D_i mu rate * exp(mu t D_i rate)  (internal children: dP/dt = U diag(de) Uinv)
//   [2K, 2K + K*(K+1))        P columns 0..K for LEAF children: column j = P[:, j], column K = row sums
//                             (= P * all-ones leaf); each column holds the K parent-state entries
//   [2K + K*(K+1), 2K + 2K(K+1)) dP/dt columns, same layout
__host__ __device__ inline int bt_size(int K) { return 2 * K + 2 * K * (K + 1); }

// Model constants of the evaluation, read as CONSTANT-BANK operands (warp-uniform: no per-lane
// register delivery, DFMA takes them directly).  One slot per distinct substitution model in the
// batch.  Slot layout (doubles): U (K*K col-major) | Uinv (K*K col-major) | pi (K) | c[r][i] =
// D_i * mu * rate_r (R*K).
constexpr int MODEL_SLOT = 256;                 // doubles per slot
constexpr int MODEL_SLOTS = 32;                 // 64 KB of constant memory
constexpr int MAX_RATES = 16;
constexpr int KMAX_GENERIC = 32;                // largest state count of the runtime-K kernel

}  // namespace mcpdev
using namespace mcpdev;
