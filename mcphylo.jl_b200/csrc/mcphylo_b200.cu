// mcphylo_b200.cu — libmcphylo_b200.so: C ABI (include/mcphylo_b200.h) + sm_100a kernels.
//
// Replaces, for one PhyloDist evaluation, the reference's
//   my_repeat + parallel_transition_prob + FelsensteinFunction (post-order pruning with per-column
//   rescaling, then the reverse-post-order gradient pass)
//   /root/reference/src/distributions/Phylodist.jl:107-138
//   /root/reference/src/Likelihood/LikelihoodCalculator_Node.jl:3-114
//   /root/reference/src/Likelihood/VectorizedFunctions.jl:13-213
//
// Design: alignment columns (site x rate category) are independent through both
// passes, so ONE thread owns one (or two) columns for the whole evaluation and walks the flat op
// program produced by schedule.hpp.  No inter-thread data dependency exists, hence one fused
// persistent kernel (post pass + gradient pass); the only CTA-level synchronisation is the
// chunk-wise staging of per-op inputs (descriptors, leaf codes, branch data) in shared memory.
// Partials live in a CTA-private scratch region laid out [slot][column][thread][state] so that a
// warp's access is one contiguous run of 32*K doubles (256-bit vector ld/st per thread at K = 4).
// fp64 throughout.  Transitions are applied in eigen-space, P L = U (e * (Uinv L)), with U / Uinv as
// constant-bank operands.  Rescaling uses exact powers of two (exponent extraction) instead of the
// reference's divide-by-max + log per node; the integer exponent sum is exact and
// logL = ln2 * sum(exponents) + sum(log(pi . L_root)).
#include "../../include/mcphylo_b200.h"
#include "schedule.hpp"

#include <cuda_runtime.h>

#include <algorithm>
#include <cmath>
#include <cstdarg>
#include <cstdio>
#include <cstring>
#include <mutex>
#include <string>
#include <vector>

namespace {

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
    double* out;                // small-tree kernel: [logL, grad] per tree, device or pinned host memory
    unsigned int* done_counter; // small-tree kernel: CTAs finished (the last one reduces the rows)
    long long row_stride;
    int n_slots, n_stack;
    int n_tiles, T, R, want_grad;
    int max_br;
    int max_rows;               // largest number of leaf rows in the batch
    // Substitution-model constants when the whole batch shares ONE model (the common case): kernel
    // parameters live in constant bank 0, so they reach the FP64 pipe as uniform operands without
    // a separate host-to-device copy.  Layout as in c_model (below).
    double model[176];
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
__constant__ double c_model[MODEL_SLOT * MODEL_SLOTS];

// --------------------------------------------------------------------------------------------
// vector load/store helpers (K doubles per column)
// --------------------------------------------------------------------------------------------
// Partials: written and re-read by the SAME thread inside one kernel, so they must not go
// through the non-coherent path; .cg keeps this streaming data out of L1.
template <int K>
__device__ __forceinline__ void ld_partial(const double* p, double (&v)[K]) {
    if constexpr (K == 4) {
        asm volatile("ld.global.cg.v4.f64 {%0,%1,%2,%3}, [%4];"
                     : "=d"(v[0]), "=d"(v[1]), "=d"(v[2]), "=d"(v[3]) : "l"(p) : "memory");
    } else if constexpr (K == 2) {
        asm volatile("ld.global.cg.v2.f64 {%0,%1}, [%2];" : "=d"(v[0]), "=d"(v[1]) : "l"(p) : "memory");
    } else {
#pragma unroll
        for (int k = 0; k < K; ++k) v[k] = __ldcg(p + k);
    }
}
template <int K>
__device__ __forceinline__ void st_partial(double* p, const double (&v)[K]) {
    if constexpr (K == 4) {
        asm volatile("st.global.cg.v4.f64 [%0], {%1,%2,%3,%4};"
                     :: "l"(p), "d"(v[0]), "d"(v[1]), "d"(v[2]), "d"(v[3]) : "memory");
    } else if constexpr (K == 2) {
        asm volatile("st.global.cg.v2.f64 [%0], {%1,%2};" :: "l"(p), "d"(v[0]), "d"(v[1]) : "memory");
    } else {
#pragma unroll
        for (int k = 0; k < K; ++k) __stcg(p + k, v[k]);
    }
}

// Identity moves the compiler cannot see through: a value passed through them is kept in a register
// (or spilled as one word) instead of being RE-COMPUTED at every use.  ptxas otherwise rematerialises
// the per-thread scratch base (blockIdx * scratch_per_cta + tid * K * 8, ~13 instructions) in front of
// every partial load/store of the walk.
__device__ __forceinline__ unsigned char* keep_ptr(unsigned char* p) {
    asm volatile("mov.u64 %0, %0;" : "+l"(p));
    return p;
}
__device__ __forceinline__ double keep_f64(double v) {
    asm volatile("mov.f64 %0, %0;" : "+d"(v));
    return v;
}

// L2 prefetch of a line the thread will read a few ops later (HBM -> L2 ahead of the demand load)
__device__ __forceinline__ void prefetch_l2(const void* p) { asm volatile("prefetch.global.L2 [%0];" :: "l"(p)); }

__device__ __forceinline__ void cp_async16(void* smem_dst, const void* gmem_src) {
    unsigned d = (unsigned)__cvta_generic_to_shared(smem_dst);
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" :: "r"(d), "l"(gmem_src) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_group 0;" ::: "memory"); }

// Model view: MOFS is the slot offset into c_model; with a compile-time 0 the constant operands
// fold into the DFMA encodings.
// DYN = false: the single model embedded in the kernel parameters; DYN = true: slot `ofs` of c_model.
template <int K, bool DYN>
struct ModelT {
    const WalkParams& p;
    int ofs;
    __device__ __forceinline__ double U(int s, int i) const { return DYN ? c_model[ofs + s + K * i] : p.model[s + K * i]; }
    __device__ __forceinline__ double Ui(int i, int j) const { return DYN ? c_model[ofs + K * K + i + K * j] : p.model[K * K + i + K * j]; }
    __device__ __forceinline__ double pi(int k) const { return DYN ? c_model[ofs + 2 * K * K + k] : p.model[2 * K * K + k]; }
    __device__ __forceinline__ double c(int r, int i) const {
        return DYN ? c_model[ofs + 2 * K * K + K + r * K + i] : p.model[2 * K * K + K + r * K + i];
    }
};

// ---- eigen-space products for C columns at once (column index innermost, so one constant /
// uniform-register operand feeds C independent DFMAs) ----
//
// Transitions are applied as  P L = L + U (em1 * (Uinv L)),  em1_i = expm1(mu t D_i r),  not as
// U (e * (Uinv L)): the latter is accurate only relative to |L|_max, and a partial likelihood vector
// routinely holds components 1e-20 of its maximum that still decide the likelihood of a site further
// up (a mismatch selects exactly that component).  The reference multiplies by an explicit
// non-negative P, which is component-wise accurate; adding the (accurately formed) deviation P - I
// onto L keeps that property, costs no extra instruction (the leading multiply becomes an FMA onto
// L), and makes identity branches exact.
//
// z[c] = em1 * w[c],  w[c] = Uinv L[c];  WD also returns zd[c] = de * w[c]  (the eigen-coordinates of dP L)
template <int K, int C, bool WD, class M>
__device__ __forceinline__ void eig_project(const M& m, const double (&L)[C][K], const double (&em1)[K], const double* de,
                                            double (&z)[C][K], double (&zd)[C][K]) {
#pragma unroll
    for (int i = 0; i < K; ++i) {
        double w[C];
#pragma unroll
        for (int c = 0; c < C; ++c) w[c] = m.Ui(i, 0) * L[c][0];
#pragma unroll
        for (int j = 1; j < K; ++j)
#pragma unroll
            for (int c = 0; c < C; ++c) w[c] = fma(m.Ui(i, j), L[c][j], w[c]);
#pragma unroll
        for (int c = 0; c < C; ++c) {
            z[c][i] = em1[i] * w[c];
            if constexpr (WD) zd[c][i] = de[i] * w[c];
        }
    }
}
// out[c] = base[c] + U z[c]
template <int K, int C, class M>
__device__ __forceinline__ void eig_expand(const M& m, const double (&z)[C][K], const double (&base)[C][K], double (&out)[C][K]) {
#pragma unroll
    for (int s = 0; s < K; ++s) {
#pragma unroll
        for (int c = 0; c < C; ++c) out[c][s] = fma(m.U(s, 0), z[c][0], base[c][s]);
#pragma unroll
        for (int i = 1; i < K; ++i)
#pragma unroll
            for (int c = 0; c < C; ++c) out[c][s] = fma(m.U(s, i), z[c][i], out[c][s]);
    }
}
// out[c] = U z[c]
template <int K, int C, class M>
__device__ __forceinline__ void eig_expand0(const M& m, const double (&z)[C][K], double (&out)[C][K]) {
#pragma unroll
    for (int s = 0; s < K; ++s) {
#pragma unroll
        for (int c = 0; c < C; ++c) out[c][s] = m.U(s, 0) * z[c][0];
#pragma unroll
        for (int i = 1; i < K; ++i)
#pragma unroll
            for (int c = 0; c < C; ++c) out[c][s] = fma(m.U(s, i), z[c][i], out[c][s]);
    }
}
// out[c] = P^T q[c] = q[c] + Uinv^T (em1 * (U^T q[c]))
template <int K, int C, class M>
__device__ __forceinline__ void eig_transposed(const M& m, const double (&q)[C][K], const double (&em1)[K], double (&out)[C][K]) {
    double z[C][K];
#pragma unroll
    for (int i = 0; i < K; ++i) {
        double w[C];
#pragma unroll
        for (int c = 0; c < C; ++c) w[c] = m.U(0, i) * q[c][0];
#pragma unroll
        for (int s = 1; s < K; ++s)
#pragma unroll
            for (int c = 0; c < C; ++c) w[c] = fma(m.U(s, i), q[c][s], w[c]);
#pragma unroll
        for (int c = 0; c < C; ++c) z[c][i] = em1[i] * w[c];
    }
#pragma unroll
    for (int j = 0; j < K; ++j) {
#pragma unroll
        for (int c = 0; c < C; ++c) out[c][j] = fma(m.Ui(0, j), z[c][0], q[c][j]);
#pragma unroll
        for (int i = 1; i < K; ++i)
#pragma unroll
            for (int c = 0; c < C; ++c) out[c][j] = fma(m.Ui(i, j), z[c][i], out[c][j]);
    }
}

// Multiply a column by the exact power of two that brings its largest magnitude into [1,2);
// returns the removed binary exponent.  Works on the exponent fields with integer ops (fp64 has no
// native max instruction; fmax() costs ~10 instructions).  Zero / denormal / non-finite maxima are
// left alone.
template <int K>
__device__ __forceinline__ int rescale_pow2(double (&v)[K]) {
    unsigned m = (unsigned)__double2hiint(v[0]) & 0x7fffffffu;
#pragma unroll
    for (int k = 1; k < K; ++k) m = max(m, (unsigned)__double2hiint(v[k]) & 0x7fffffffu);
    const int e = (int)(m >> 20);
    if (e == 0 || e == 0x7ff) return 0;
    const double sc = __hiloint2double((2046 - e) << 20, 0);
#pragma unroll
    for (int k = 0; k < K; ++k) v[k] *= sc;
    return e - 1023;
}

// 1/x for a positive, normal x: hardware seed + two Newton steps (relative error ~1e-16; the
// quotient only scales a gradient term whose tolerance is 1e-8).
__device__ __forceinline__ double fast_rcp(double x) {
    double r;
    asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(x));
    r = fma(fma(-x, r, 1.0), r, r);
    r = fma(fma(-x, r, 1.0), r, r);
    r = fma(fma(-x, r, 1.0), r, r);
    return r;
}

// lane 0 ends with sum(va) over the warp, lane 16 with sum(vb)
__device__ __forceinline__ double warp_pair_reduce(double va, double vb, int lane) {
    const bool upper = (lane & 16) != 0;
    double send = upper ? va : vb;
    double keep = upper ? vb : va;
    double v = keep + __shfl_xor_sync(0xffffffffu, send, 16);
    v += __shfl_xor_sync(0xffffffffu, v, 8);
    v += __shfl_xor_sync(0xffffffffu, v, 4);
    v += __shfl_xor_sync(0xffffffffu, v, 2);
    v += __shfl_xor_sync(0xffffffffu, v, 1);
    return v;
}

// --------------------------------------------------------------------------------------------
// kernel 1: branch tables for every (tree, branch, rate)
//   e   = exp(mu t D r)
//   P   = U diag(e) Uinv                             VectorizedFunctions.jl:116-168
//   dP  = U diag(D r mu e) Uinv                      VectorizedFunctions.jl:89-113, 139-152
// (P, dP columns are only read for LEAF children; same operation order as the reference.)
// --------------------------------------------------------------------------------------------
constexpr int KMAX_TABLE = 32;

__global__ void build_branch_tables(const TreeDev* __restrict__ trees, const double* __restrict__ dyn,
                                    double* __restrict__ btab, int K, int R) {
    const TreeDev tr = trees[blockIdx.y];
    const int idx = blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= tr.n_br * R) return;
    const int br = idx / R, r = idx - br * R;
    const double* d = dyn + tr.dyn_off;
    const double* U = d + dyn_U(tr.NN);
    const double* D = d + dyn_D(tr.NN, K);
    const double* Uinv = d + dyn_Uinv(tr.NN, K);
    const double mu = d[dyn_mu(tr.NN, K)];
    const double rate = d[dyn_rates(tr.NN, K) + r];
    double* ev = btab + tr.btab_off + ((long long)br * R + r) * bt_size(K);
    double* P = ev + 2 * K;
    double* dP = P + K * (K + 1);
    if (br >= tr.NN - 1) {  // root row (unused) and virtual branches: identity, zero derivative
        for (int i = 0; i < K; ++i) { ev[i] = 0.0; ev[K + i] = 0.0; }   // expm1(0) and zero derivative: identity branch
        for (int n = 0; n <= K; ++n)
            for (int m = 0; m < K; ++m) {
                P[n * K + m] = (n == K || n == m) ? 1.0 : 0.0;
                dP[n * K + m] = 0.0;
            }
        return;
    }
    const double t = d[dyn_blv(tr.NN) + br];
    double em1[KMAX_TABLE], de[KMAX_TABLE];
    for (int i = 0; i < K; ++i) {
        const double x = mu * t * D[i] * rate;
        em1[i] = expm1(x);
        ev[i] = em1[i];
        de[i] = D[i] * rate * mu * exp(x);
        ev[K + i] = de[i];
    }
    for (int m = 0; m < K; ++m) {
        double rs = 0.0, drs = 0.0;
        for (int n = 0; n < K; ++n) {
            double c = 0.0, dc = 0.0;
            for (int k = 0; k < K; ++k) {
                const double u = U[m + K * k], ui = Uinv[k + K * n];
                c += (u * em1[k]) * ui;      // P - I, formed without cancellation against the identity
                dc += (u * de[k]) * ui;
            }
            c += (m == n) ? 1.0 : 0.0;
            P[n * K + m] = c;
            dP[n * K + m] = dc;
            rs += c;    // what P * (all-ones leaf) gives: sum_s1 1 * P[s, s1]
            drs += dc;
        }
        P[K * K + m] = rs;
        dP[K * K + m] = drs;
    }
}

// --------------------------------------------------------------------------------------------
// kernel 2: the fused walk.  grid = persistent CTAs, each takes a contiguous range of column
// tiles; a tile = blockDim.x columns of one rate category of one tree.
//
// Per-op inputs that are uniform over the tile or byte-sized per column are staged in shared
// memory CH ops at a time with cp.async, one chunk ahead of the compute:
//   sdesc  3 x CH op descriptors (ring of 3: descriptors must be resident one chunk before the
//          data they describe can be requested)
//   se     2 x CH x 2 x 2K doubles: (em1, de) eigen-coefficient vectors of INTERNAL children
//   scode  2 x CH x 2 x TW bytes: the state codes of LEAF children for the tile's columns
// so the only global accesses on the per-op critical path are the thread's own partials and the
// leaf-table gathers.  One __syncthreads per chunk.
// --------------------------------------------------------------------------------------------
constexpr int CH = 16;

// Per-op record derived by the staging threads from the raw descriptor (schedule.hpp): everything the
// compute threads need as ready-to-add byte offsets, so no warp repeats the uniform address math.
//   xa / xb  LEAF child: byte offset (from the branch-table base) of the child's P columns for this
//            tile's rate; MEM child: byte offset (from the thread's scratch base) of its stored partial
//   post: y0 = where to store the result          pre: y0 = pre[mother] on the LIFO, y1 / y2 = where
//                                                       pre[a] / pre[b] are pushed
// Offsets are 32-bit: a CTA's scratch region and one tree's branch table are far below 4 GB (checked
// on the host).
struct __align__(16) OpRec {
    int flags;
    unsigned xa, xb, y0;       // first half: needed at the start of the op
    int a_br, b_br;
    unsigned y1, y2;           // second half: needed at its end
};
static_assert(sizeof(OpRec) == 32, "OpRec is two 16-byte words");

// Where a CTA accumulates its branch-gradient sums.  K <= 3: directly in its accumulator row in
// global memory with fire-and-forget RED.ADD.F64 (the row stays in L2); a shared-memory fp64 atomic
// add is a compare-and-swap loop (~10 instructions, 38 % retries when the 8 warps of a CTA hit the
// same branch).  Measured (profiles/r1_walk_notes.md): 5 % faster at K = 2, but 2.5-5 % SLOWER at
// K = 4, where the kernel sits on a register knife-edge and the extra 64-bit row pointer spills.
__host__ __device__ constexpr bool grad_in_l2(int K) { return K <= 3; }

template <int K>
struct WalkSmem {
    // dynamic shared memory carve-up (offsets in bytes)
    // branch-gradient accumulator of the CTA: in shared memory for K >= 4; for K <= 3 it is the CTA's
    // row in global memory (see grad_in_l2)
    static __host__ __device__ size_t acc_bytes(int n_br, int want_grad) {
        return (want_grad && !grad_in_l2(K)) ? (((size_t)n_br * 8 + 15) & ~(size_t)15) : 0;
    }
    static __host__ __device__ size_t desc_bytes() { return 3 * CH * 32; }
    static __host__ __device__ size_t e_bytes() { return 2 * CH * 2 * 2 * K * 8; }   // (em1, de) per internal child
    static __host__ __device__ size_t code_bytes(int TW) { return (size_t)2 * CH * 2 * TW; }
    static __host__ __device__ size_t rec_bytes() { return 2 * CH * 32; }
    // leaf children: P (and, in the gradient pass, dP) columns [(K+1)][K] of the child's branch
    static __host__ __device__ size_t tab_bytes() { return (size_t)2 * CH * 2 * 2 * K * (K + 1) * 8; }
    static __host__ __device__ size_t total(int n_br, int want_grad, int TW) {
        return acc_bytes(n_br, want_grad) + desc_bytes() + e_bytes() + rec_bytes() + tab_bytes() + code_bytes(TW);
    }
};

// 3 resident CTAs of 256 threads per SM (<= 80 registers): the walk is latency-bound, 24 warps
// with a few spills beat 16 warps without (profiles/r1_walk_notes.md).
#ifndef MCP_WALK_MIN_BLOCKS
#define MCP_WALK_MIN_BLOCKS 3
#endif
// Heavier per-thread state (K * columns per thread > 4 doubles per vector) gets 2 CTAs per SM
// (<= 128 registers) instead of 3.
#ifndef MCP_WALK_MIN_BLOCKS2
#define MCP_WALK_MIN_BLOCKS2 2
#endif
// SSCR = true keeps the CTA's partials scratch in SHARED memory instead of HBM: the latency path for
// small problems (MCMC-sized trees), where a lone warp would otherwise wait an L2 round trip for
// every partial it has just written.
template <int K, int CPT, bool DYN_MODEL, bool SSCR>
#ifndef MCP_WALK_MAXT
#define MCP_WALK_MAXT 256
#endif
#ifndef MCP_PREFETCH_DIST
#define MCP_PREFETCH_DIST 0   // L2 prefetch hints for gradient-pass operands: measured slower (22.3 vs 21.0 ms), kept for experiments
#endif
__global__ void __launch_bounds__(MCP_WALK_MAXT, K * CPT <= 4 ? MCP_WALK_MIN_BLOCKS : MCP_WALK_MIN_BLOCKS2) felsenstein_walk(const __grid_constant__ WalkParams p) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    __shared__ long long s_e[8];
    __shared__ double s_l[8];

    const int tid = threadIdx.x, TW = blockDim.x, lane = tid & 31, warp = tid >> 5;
    const int TS = TW * CPT;                          // sites per tile; column c of a thread = site0 + c*TW + tid
    const int q = p.n_tiles / gridDim.x, rem = p.n_tiles - q * gridDim.x;
    int tile = blockIdx.x * q + min((int)blockIdx.x, rem);
    const int tile_end = tile + q + ((int)blockIdx.x < rem ? 1 : 0);
    if (tile >= tile_end) return;

    double* const s_acc = reinterpret_cast<double*>(smem_raw);
    constexpr bool GL2 = grad_in_l2(K);
    int4* const sdesc = reinterpret_cast<int4*>(smem_raw + WalkSmem<K>::acc_bytes(p.max_br, p.want_grad));
    double* const se = reinterpret_cast<double*>(reinterpret_cast<unsigned char*>(sdesc) + WalkSmem<K>::desc_bytes());
    OpRec* const srec = reinterpret_cast<OpRec*>(reinterpret_cast<unsigned char*>(se) + WalkSmem<K>::e_bytes());
    double* const stab = reinterpret_cast<double*>(reinterpret_cast<unsigned char*>(srec) + WalkSmem<K>::rec_bytes());
    unsigned char* const scode = reinterpret_cast<unsigned char*>(stab) + WalkSmem<K>::tab_bytes();
    constexpr int KK1 = K * (K + 1);                  // doubles of one leaf table (P or dP columns)

    // per-thread base of the CTA-private scratch, laid out [slot][column c][thread][state]; all
    // slot / LIFO offsets in the records are byte offsets from here
    unsigned char* const scr = SSCR
        ? scode + WalkSmem<K>::code_bytes(TS) + (size_t)tid * K * 8
        : keep_ptr(reinterpret_cast<unsigned char*>(p.scratch + (long long)blockIdx.x * p.scratch_per_cta + (long long)tid * K));
    const unsigned col_bytes = (unsigned)TW * K * 8;  // distance between a thread's columns within a slot
    const unsigned slot_bytes = col_bytes * CPT;
    const unsigned stack_base = (unsigned)p.n_slots * slot_bytes;
    int row = p.cta_row_base[blockIdx.x];
    const int R = p.R;
    constexpr int BT = 2 * K + 2 * K * (K + 1);

    int ti = 0;
    while (ti < p.T - 1 && tile >= p.trees[ti].tile_begin + R * p.trees[ti].tiles_per_rate) ++ti;

    while (tile < tile_end) {
        const TreeDev tr = p.trees[ti];
        const int tree_tile_end = min(tile_end, tr.tile_begin + R * tr.tiles_per_rate);
        // this CTA's gradient accumulator row for the tree (global memory, stays in L2)
        double* const grow = p.rows + (long long)row * p.row_stride;
        if (p.want_grad) {   // ordered before the first update by the barriers below
            for (int i = tid; i < tr.n_br; i += TW) (GL2 ? grow : s_acc)[i] = 0.0;
        }
        long long e_total = 0;
        double logsum = 0.0;
        const ModelT<K, DYN_MODEL> mdl{p, DYN_MODEL ? (int)__ldg(p.dyn + tr.dyn_off + dyn_slot(tr.NN, K, R)) * MODEL_SLOT : 0};
        const int4* const post_ops = p.ops + 2 * tr.post_off;
        const int4* const pre_ops = p.ops + 2 * tr.pre_off;

        for (; tile < tree_tile_end; ++tile) {
            const int local = tile - tr.tile_begin;
            const int r = local / tr.tiles_per_rate;
            const long long site0 = (long long)(local - r * tr.tiles_per_rate) * TS;
            bool valid[CPT];
            double vmask[CPT];                 // 1.0 for real columns, 0.0 for the padding of a ragged tile
#pragma unroll
            for (int c = 0; c < CPT; ++c) {
                valid[c] = site0 + c * TW + tid < tr.S;
                vmask[c] = keep_f64(valid[c] ? 1.0 : 0.0);
            }
            const unsigned char* const codes0 = tr.codes + site0;
            // this tree's branch table at (branch 0, rate r); record offsets are relative to it
            const unsigned char* const btab_b = reinterpret_cast<const unsigned char*>(p.btab + tr.btab_off + (long long)r * BT);
            const unsigned br_bytes = (unsigned)R * BT * 8;

            // ---- chunk staging (all threads of the CTA) ----
            auto stage_desc = [&](const int4* ops, int n_ops, int c) {
                const int base = c * CH, cnt = min(CH, n_ops - base);
                int4* dst = sdesc + (c % 3) * (CH * 2);
                for (int i = tid; i < cnt * 2; i += TW) cp_async16(dst + i, ops + 2 * base + i);
            };
            // Copies e vectors / leaf codes of chunk c and derives the per-op records (byte offsets),
            // once per CTA instead of once per warp.  `pre` selects the pre-program field meaning.
            auto stage_data = [&](int n_ops, int c, bool pre) {
                const int base = c * CH, cnt = min(CH, n_ops - base);
                const int4* d = sdesc + (c % 3) * (CH * 2);
                double* eb = se + (c & 1) * (CH * 2 * 2 * K);
                unsigned char* cb = scode + (size_t)(c & 1) * (CH * 2 * TS);
                OpRec* rb = srec + (c & 1) * CH;
                double* tb = stab + (size_t)(c & 1) * (CH * 2 * 2 * KK1);
                const int pieces = TS / 16;                  // 16-byte pieces of one code row segment
                const int tab_doubles = pre ? 2 * KK1 : KK1; // P columns (+ dP columns in the gradient pass)
                const int tpieces = (tab_doubles + 1) / 2;   // 16-byte pieces of one leaf table
                const int epieces = K;                       // em1 and de: 2K doubles = K 16-byte pieces
                const int per_child = pieces + tpieces > epieces ? pieces + tpieces : epieces;
                for (int w = tid; w < cnt * 2 * per_child; w += TW) {
                    const int piece = w % per_child, jc = w / per_child, j = jc >> 1, ch = jc & 1;
                    const int4 o0 = d[2 * j];
                    const int fl = d[2 * j + 1].y;
                    const int kind = ch ? ((fl >> 2) & 3) : (fl & 3);
                    const int src = ch ? o0.z : o0.x, br = ch ? o0.w : o0.y;
                    const double* bsrc = reinterpret_cast<const double*>(btab_b + (unsigned)br * br_bytes);
                    if (kind == mcp::OPK_LEAF) {
                        if (piece < pieces) {
                            unsigned char* dstp = cb + (size_t)(j * 2 + ch) * TS + piece * 16;
                            if (src >= 0) cp_async16(dstp, codes0 + (long long)src * tr.code_stride + piece * 16);
                            else *reinterpret_cast<uint4*>(dstp) = make_uint4(0x01010101u * K, 0x01010101u * K, 0x01010101u * K, 0x01010101u * K);
                        } else if (piece < pieces + tpieces) {
                            const int tp = piece - pieces;
                            double* dstp = tb + (size_t)(j * 2 + ch) * 2 * KK1 + tp * 2;
                            const double* srcp = bsrc + 2 * K + tp * 2;
                            if constexpr ((K * 8) % 16 == 0) {
                                cp_async16(dstp, srcp);
                            } else {
                                dstp[0] = __ldg(srcp);
                                if (tp * 2 + 1 < tab_doubles) dstp[1] = __ldg(srcp + 1);
                            }
                        }
                    } else if (piece < epieces) {
                        // 2K doubles = K 16-byte pieces; entries are 16-byte aligned (bt_size is even)
                        cp_async16(reinterpret_cast<unsigned char*>(eb + (j * 2 + ch) * 2 * K) + piece * 16,
                                   reinterpret_cast<const unsigned char*>(bsrc) + piece * 16);
                    }
                }
                for (int j = tid; j < cnt; j += TW) {
                    const int4 o0 = d[2 * j], o1 = d[2 * j + 1];
                    const int fl = o1.y, ka = fl & 3, kb = (fl >> 2) & 3;
                    OpRec rec;
                    rec.flags = fl;
                    rec.a_br = o0.y;
                    rec.b_br = o0.w;
                    rec.xa = ka == mcp::OPK_LEAF ? (unsigned)o0.y * br_bytes + 2 * K * 8 : (unsigned)o0.x * slot_bytes;
                    rec.xb = kb == mcp::OPK_LEAF ? (unsigned)o0.w * br_bytes + 2 * K * 8 : (unsigned)o0.z * slot_bytes;
                    if (pre) {
                        rec.y0 = stack_base + (unsigned)o1.x * slot_bytes;     // pre[mother] on the LIFO
                        rec.y1 = stack_base + (unsigned)o1.z * slot_bytes;     // where pre[a] is pushed
                        rec.y2 = stack_base + (unsigned)o1.w * slot_bytes;     // where pre[b] is pushed
                    } else {
                        rec.y0 = (unsigned)o1.x * slot_bytes;                  // where the result is stored
                        rec.y1 = 0;
                        rec.y2 = 0;
                    }
                    rb[j] = rec;
                }
            };
            auto prologue = [&](const int4* ops, int n_ops, bool pre) {
                __syncthreads();                              // previous pass / tile done with the buffers
                stage_desc(ops, n_ops, 0);
                if (n_ops > CH) stage_desc(ops, n_ops, 1);
                cp_async_commit();
                cp_async_wait_all();
                __syncthreads();
                stage_data(n_ops, 0, pre);
                cp_async_commit();
            };
            auto chunk_boundary = [&](const int4* ops, int n_ops, int c, int n_chunks, bool pre) {
                cp_async_wait_all();
                __syncthreads();                              // chunk c data + descriptors c, c+1 visible
                if (c + 1 < n_chunks) stage_data(n_ops, c + 1, pre);
                if (c + 2 < n_chunks) stage_desc(ops, n_ops, c + 2);
                cp_async_commit();
            };
            auto ld_cols = [&](unsigned off, double (&v)[CPT][K]) {
#pragma unroll
                for (int c = 0; c < CPT; ++c) {
                    if constexpr (SSCR) {
                        const double* src = reinterpret_cast<const double*>(scr + off + c * col_bytes);
#pragma unroll
                        for (int k = 0; k < K; ++k) v[c][k] = src[k];
                    } else {
                        ld_partial<K>(reinterpret_cast<const double*>(scr + off + c * col_bytes), v[c]);
                    }
                }
            };
            auto st_cols = [&](unsigned off, const double (&v)[CPT][K]) {
#pragma unroll
                for (int c = 0; c < CPT; ++c) {
                    if constexpr (SSCR) {
                        double* dst = reinterpret_cast<double*>(scr + off + c * col_bytes);
#pragma unroll
                        for (int k = 0; k < K; ++k) dst[k] = v[c][k];
                    } else {
                        st_partial<K>(reinterpret_cast<double*>(scr + off + c * col_bytes), v[c]);
                    }
                }
            };

            // ------------------------------ post pass ------------------------------
            double cur[CPT][K];
#pragma unroll
            for (int c = 0; c < CPT; ++c)
#pragma unroll
                for (int k = 0; k < K; ++k) cur[c][k] = 1.0;
            int e_col[CPT];
#pragma unroll
            for (int c = 0; c < CPT; ++c) e_col[c] = 0;
            {
                const int n_post = tr.n_post, n_chunks = (n_post + CH - 1) / CH;
                prologue(post_ops, n_post, false);
                for (int c = 0; c < n_chunks; ++c) {
                    chunk_boundary(post_ops, n_post, c, n_chunks, false);
                    const OpRec* rb = srec + (c & 1) * CH;
                    const double* eb = se + (c & 1) * (CH * 2 * 2 * K);
                    const unsigned char* cb = scode + (size_t)(c & 1) * (CH * 2 * TS) + tid;
                    const double* tb = stab + (size_t)(c & 1) * (CH * 2 * 2 * KK1);
                    const int cnt = min(CH, n_post - c * CH);
                    for (int j = 0; j < cnt; ++j) {
                        const uint4 rh = *reinterpret_cast<const uint4*>(rb + j);   // flags, xa, xb, y0
                        const int flags = (int)rh.x, ka = flags & 3, kb = (flags >> 2) & 3;
                        // stored operand (at most one per op) first: its latency overlaps the rest
                        double Lm[CPT][K];
                        if (ka == mcp::OPK_MEM) ld_cols(rh.y, Lm);
                        else if (kb == mcp::OPK_MEM) ld_cols(rh.z, Lm);
                        double Da[CPT][K], Db[CPT][K];
                        if (ka == mcp::OPK_LEAF) {
#pragma unroll
                            for (int cc = 0; cc < CPT; ++cc) {
                                const int code = min((int)cb[(j * 2 + 0) * TS + cc * TW], K);
                                const double* t = tb + (j * 2 + 0) * 2 * KK1 + code * K;
#pragma unroll
                                for (int k = 0; k < K; ++k) Da[cc][k] = t[k];
                            }
                        } else {
                            double e[K], z[CPT][K];
#pragma unroll
                            for (int k = 0; k < K; ++k) e[k] = eb[(j * 2 + 0) * 2 * K + k];
                            if (ka == mcp::OPK_REG) { eig_project<K, CPT, false>(mdl, cur, e, nullptr, z, z); eig_expand<K, CPT>(mdl, z, cur, Da); }
                            else { eig_project<K, CPT, false>(mdl, Lm, e, nullptr, z, z); eig_expand<K, CPT>(mdl, z, Lm, Da); }
                        }
                        if (kb == mcp::OPK_LEAF) {
#pragma unroll
                            for (int cc = 0; cc < CPT; ++cc) {
                                const int code = min((int)cb[(j * 2 + 1) * TS + cc * TW], K);
                                const double* t = tb + (j * 2 + 1) * 2 * KK1 + code * K;
#pragma unroll
                                for (int k = 0; k < K; ++k) Db[cc][k] = t[k];
                            }
                        } else {
                            double e[K], z[CPT][K];
#pragma unroll
                            for (int k = 0; k < K; ++k) e[k] = eb[(j * 2 + 1) * 2 * K + k];
                            if (kb == mcp::OPK_REG) { eig_project<K, CPT, false>(mdl, cur, e, nullptr, z, z); eig_expand<K, CPT>(mdl, z, cur, Db); }
                            else { eig_project<K, CPT, false>(mdl, Lm, e, nullptr, z, z); eig_expand<K, CPT>(mdl, z, Lm, Db); }
                        }
#pragma unroll
                        for (int cc = 0; cc < CPT; ++cc) {
#pragma unroll
                            for (int k = 0; k < K; ++k) cur[cc][k] = Da[cc][k] * Db[cc][k];
                            e_col[cc] += rescale_pow2<K>(cur[cc]);
                        }
                        if (flags & mcp::POST_STORE) st_cols(rh.w, cur);
                    }
                }
            }
#pragma unroll
            for (int cc = 0; cc < CPT; ++cc) {
                double rootv = mdl.pi(0) * cur[cc][0];
#pragma unroll
                for (int k = 1; k < K; ++k) rootv = fma(mdl.pi(k), cur[cc][k], rootv);
                if (valid[cc]) {
                    logsum += log(rootv);
                    e_total += e_col[cc];
                }
            }

            // ------------------------------ gradient pass ------------------------------
            if (p.want_grad) {
                const int n_pre = tr.n_pre, n_chunks = (n_pre + CH - 1) / CH;
                prologue(pre_ops, n_pre, true);
                for (int c = 0; c < n_chunks; ++c) {
                    chunk_boundary(pre_ops, n_pre, c, n_chunks, true);
                    const OpRec* rb = srec + (c & 1) * CH;
                    const double* eb = se + (c & 1) * (CH * 2 * 2 * K);
                    const unsigned char* cb = scode + (size_t)(c & 1) * (CH * 2 * TS) + tid;
                    const double* tb = stab + (size_t)(c & 1) * (CH * 2 * 2 * KK1);
                    const int cnt = min(CH, n_pre - c * CH);
                    for (int j = 0; j < cnt; ++j) {
                        const uint4 rh = *reinterpret_cast<const uint4*>(rb + j);   // flags, xa, xb, y0
                        const int flags = (int)rh.x;
                        const bool ai = (flags & 3) == mcp::OPK_MEM, bi = ((flags >> 2) & 3) == mcp::OPK_MEM;
                        if constexpr (!SSCR && MCP_PREFETCH_DIST > 0) {
                            // the children partials of a later family were written in the post pass, long ago:
                            // pull them from HBM into L2 now (one request per 128-byte line)
                            if (j + MCP_PREFETCH_DIST < cnt && (lane * K * 8) % 128 == 0) {
                                const uint4 rf = *reinterpret_cast<const uint4*>(rb + j + MCP_PREFETCH_DIST);
                                if (((int)rf.x & 3) == mcp::OPK_MEM) {
#pragma unroll
                                    for (int cc = 0; cc < CPT; ++cc) prefetch_l2(scr + rf.y + cc * col_bytes);
                                }
                                if ((((int)rf.x >> 2) & 3) == mcp::OPK_MEM) {
#pragma unroll
                                    for (int cc = 0; cc < CPT; ++cc) prefetch_l2(scr + rf.z + cc * col_bytes);
                                }
                            }
                        }
                        const int mk = (flags >> 8) & 3;
                        // all stored operands of the family are requested up front
                        double pm[CPT][K], La[CPT][K], Lb[CPT][K];
                        if (mk == mcp::PREM_STACK) ld_cols(rh.w, pm);
                        if (ai) ld_cols(rh.y, La);
                        if (bi) ld_cols(rh.z, Lb);
                        if (mk == mcp::PREM_ROOT) {
#pragma unroll
                            for (int cc = 0; cc < CPT; ++cc)
#pragma unroll
                                for (int k = 0; k < K; ++k) pm[cc][k] = mdl.pi(k);
                        } else if (mk == mcp::PREM_REG) {
#pragma unroll
                            for (int cc = 0; cc < CPT; ++cc)
#pragma unroll
                                for (int k = 0; k < K; ++k) pm[cc][k] = cur[cc][k];
                        }
                        double ea[K], ebv[K];
                        double Da[CPT][K], Ya[CPT][K], Db[CPT][K], Yb[CPT][K];
                        if (ai) {
                            double z[CPT][K], zd[CPT][K];
#pragma unroll
                            for (int k = 0; k < K; ++k) ea[k] = eb[(j * 2 + 0) * 2 * K + k];
                            eig_project<K, CPT, true>(mdl, La, ea, eb + (j * 2 + 0) * 2 * K + K, z, zd);
                            eig_expand<K, CPT>(mdl, z, La, Da);
                            eig_expand0<K, CPT>(mdl, zd, Ya);
                        } else {
#pragma unroll
                            for (int cc = 0; cc < CPT; ++cc) {
                                const int code = min((int)cb[(j * 2 + 0) * TS + cc * TW], K);
                                const double* t = tb + (j * 2 + 0) * 2 * KK1 + code * K;
#pragma unroll
                                for (int k = 0; k < K; ++k) { Da[cc][k] = t[k]; Ya[cc][k] = t[KK1 + k]; }
                            }
                        }
                        if (bi) {
                            double z[CPT][K], zd[CPT][K];
#pragma unroll
                            for (int k = 0; k < K; ++k) ebv[k] = eb[(j * 2 + 1) * 2 * K + k];
                            eig_project<K, CPT, true>(mdl, Lb, ebv, eb + (j * 2 + 1) * 2 * K + K, z, zd);
                            eig_expand<K, CPT>(mdl, z, Lb, Db);
                            eig_expand0<K, CPT>(mdl, zd, Yb);
                        } else {
#pragma unroll
                            for (int cc = 0; cc < CPT; ++cc) {
                                const int code = min((int)cb[(j * 2 + 1) * TS + cc * TW], K);
                                const double* t = tb + (j * 2 + 1) * 2 * KK1 + code * K;
#pragma unroll
                                for (int k = 0; k < K; ++k) { Db[cc][k] = t[k]; Yb[cc][k] = t[KK1 + k]; }
                            }
                        }
                        double qa[CPT][K], qb[CPT][K];
                        double ga = 0.0, gb = 0.0;
#pragma unroll
                        for (int cc = 0; cc < CPT; ++cc) {
                            double den = 0.0, na = 0.0, nb = 0.0;
#pragma unroll
                            for (int k = 0; k < K; ++k) {
                                qa[cc][k] = pm[cc][k] * Db[cc][k];
                                qb[cc][k] = pm[cc][k] * Da[cc][k];
                                den = fma(qa[cc][k], Da[cc][k], den);
                                na = fma(qa[cc][k], Ya[cc][k], na);
                                nb = fma(qb[cc][k], Yb[cc][k], nb);
                            }
                            const double inv = fast_rcp(den) * vmask[cc];
                            ga = fma(na, inv, ga);
                            gb = fma(nb, inv, gb);
                        }
                        const double red = warp_pair_reduce(ga, gb, lane);   // lane 0: sum of ga, lane 16: sum of gb
                        if constexpr (GL2) {
                            if ((lane & 15) == 0) atomicAdd(grow + ((lane >> 4) ? rb[j].b_br : rb[j].a_br), red);
                        } else {
                            if (lane == 0) atomicAdd(&s_acc[rb[j].a_br], red);
                            else if (lane == 16) atomicAdd(&s_acc[rb[j].b_br], red);
                        }

                        // pre[child] = P^T q, only internal children have one
                        const int a_out = (flags >> 10) & 3, b_out = (flags >> 12) & 3;
                        if (a_out != mcp::OUT_NONE) {
                            double pa[CPT][K];
                            eig_transposed<K, CPT>(mdl, qa, ea, pa);
#pragma unroll
                            for (int cc = 0; cc < CPT; ++cc) rescale_pow2<K>(pa[cc]);
                            if (a_out == mcp::OUT_KEEP) {
#pragma unroll
                                for (int cc = 0; cc < CPT; ++cc)
#pragma unroll
                                    for (int k = 0; k < K; ++k) cur[cc][k] = pa[cc][k];
                            } else {
                                st_cols(rb[j].y1, pa);
                            }
                        }
                        if (b_out != mcp::OUT_NONE) {
                            double pb[CPT][K];
                            eig_transposed<K, CPT>(mdl, qb, ebv, pb);
#pragma unroll
                            for (int cc = 0; cc < CPT; ++cc) rescale_pow2<K>(pb[cc]);
                            if (b_out == mcp::OUT_KEEP) {
#pragma unroll
                                for (int cc = 0; cc < CPT; ++cc)
#pragma unroll
                                    for (int k = 0; k < K; ++k) cur[cc][k] = pb[cc][k];
                            } else {
                                st_cols(rb[j].y2, pb);
                            }
                        }
                    }
                }
            }
        }  // tiles of this tree

        // ---- flush this CTA's sums for the tree into its accumulator row ----
        for (int off = 16; off > 0; off >>= 1) {
            e_total += __shfl_xor_sync(0xffffffffu, e_total, off);
            logsum += __shfl_xor_sync(0xffffffffu, logsum, off);
        }
        if (lane == 0) { s_e[warp] = e_total; s_l[warp] = logsum; }
        __syncthreads();
        if (tid == 0) {
            long long es = 0;
            double ls = 0.0;
            for (int w = 0; w < (TW + 31) / 32; ++w) { es += s_e[w]; ls += s_l[w]; }
            p.rows_ll[row].esum = es;
            p.rows_ll[row].logsum = ls;
        }
        if constexpr (!GL2) {
            if (p.want_grad)
                for (int i = tid; i < tr.n_br; i += TW) grow[i] = s_acc[i];
        }
        __syncthreads();
        ++row;
        ++ti;
    }
}

// --------------------------------------------------------------------------------------------
// Branch-length prior epilogue (CompoundDirichlet / exponentialBL,
// /root/reference/src/Likelihood/Prior.jl:1-57): the whole block reduces T = sum t_j and
// W = sum w_j log t_j in a fixed order (thread-strided partial sums, then a serial sum over the
// threads), so the value is reproducible.  s_red: 2 * blockDim.x doubles.  Returns {T, W} to
// every thread.
// --------------------------------------------------------------------------------------------
struct PriorSums { double T, W; };
__device__ inline PriorSums prior_block_sums(const double* __restrict__ blv, const double* __restrict__ w, int nb,
                                             int tid, int nt, double* s_red) {
    double a = 0.0, b = 0.0;
    for (int j = tid; j < nb; j += nt) {
        const double t = blv[j], wj = w[j];
        a += t;
        if (wj != 0.0) b += wj * log(t);
    }
    s_red[tid] = a;
    s_red[nt + tid] = b;
    __syncthreads();
    PriorSums r{0.0, 0.0};
    for (int i = 0; i < nt; ++i) { r.T += s_red[i]; r.W += s_red[nt + i]; }
    __syncthreads();
    return r;
}
// contribution of the prior to output slot j (0 = log density, j >= 1 = d/dt_j)
__device__ inline double prior_term(const double* __restrict__ hdr, const double* __restrict__ blv,
                                    const double* __restrict__ w, const PriorSums& ps, int j) {
    const double c0 = hdr[1], beta = hdr[2], k4 = hdr[3];
    if (j == 0) return c0 - beta * ps.T + ps.W + (k4 != 0.0 ? k4 * log(ps.T) : 0.0);
    const double wj = w[j - 1];
    return -beta + (wj != 0.0 ? wj / blv[j - 1] : 0.0) + (k4 != 0.0 ? k4 / ps.T : 0.0);
}

// --------------------------------------------------------------------------------------------
// kernel 2c: small-tree latency path.  A tile is ONE warp wide (32 columns) but is worked on by all
// W warps of the CTA: the level-ordered program (schedule.hpp, by_levels) lists ops of equal height
// (post pass) / depth (gradient pass) together, the warps split each level's ops, and a
// __syncthreads separates levels.  All partials and pre vectors of the tile live in shared memory.
// The critical path is the tree height instead of the node count; used when the whole input is only
// a few tiles per SM (MCMC-sized problems), where the depth-first walk runs at single-warp latency.
// --------------------------------------------------------------------------------------------
struct LevelSmem {
    // byte offsets into dynamic shared memory
    static __host__ __device__ size_t acc_bytes(int n_br, int want_grad) { return want_grad ? (((size_t)n_br * 8 + 127) & ~(size_t)127) : 0; }
    static __host__ __device__ size_t exp_bytes() { return 128; }
    static __host__ __device__ size_t code_bytes(int n_rows) { return (((size_t)n_rows * 32) + 127) & ~(size_t)127; }
    static __host__ __device__ size_t slot_bytes(int K) { return (size_t)32 * K * 8; }
    static __host__ __device__ size_t tab_bytes(int n_br, int K) { return (((size_t)n_br * bt_size(K) * 8) + 127) & ~(size_t)127; }
    static __host__ __device__ size_t total(int n_br, int want_grad, int n_rows, int n_slots, int n_stack, int K) {
        return acc_bytes(n_br, want_grad) + exp_bytes() + code_bytes(n_rows) + tab_bytes(n_br, K) +
               (size_t)(n_slots + n_stack) * slot_bytes(K);
    }
};

template <int K, bool DYN_MODEL>
__global__ void __launch_bounds__(256) felsenstein_walk_levels(const __grid_constant__ WalkParams p) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    __shared__ long long s_e[8];
    __shared__ double s_l[8];
    __shared__ double s_prior[2 * 256];   // block reduction of the branch-length prior (final reduction only)

    const int tid = threadIdx.x, NT = blockDim.x, lane = tid & 31, warp = tid >> 5, W = NT >> 5;
    const int q = p.n_tiles / gridDim.x, rem = p.n_tiles - q * gridDim.x;
    int tile = blockIdx.x * q + min((int)blockIdx.x, rem);
    const int tile_end = tile + q + ((int)blockIdx.x < rem ? 1 : 0);
    __shared__ unsigned s_ticket;

    double* const s_acc = reinterpret_cast<double*>(smem_raw);
    int* const s_exp = reinterpret_cast<int*>(smem_raw + LevelSmem::acc_bytes(p.max_br, p.want_grad));
    unsigned char* const s_code = reinterpret_cast<unsigned char*>(s_exp) + LevelSmem::exp_bytes();
    double* const s_tab = reinterpret_cast<double*>(s_code + LevelSmem::code_bytes(p.max_rows));
    double* const s_post = reinterpret_cast<double*>(reinterpret_cast<unsigned char*>(s_tab) + LevelSmem::tab_bytes(p.max_br, K)) + lane * K;
    double* const s_pre = s_post + (size_t)p.n_slots * 32 * K;
    constexpr int SLOT = 32 * K;                  // doubles per slot
    constexpr int BT = 2 * K + 2 * K * (K + 1), KK1 = K * (K + 1);
    int row = p.cta_row_base[blockIdx.x];
    const int R = p.R;

    int ti = 0;
    while (tile < tile_end && ti < p.T - 1 && tile >= p.trees[ti].tile_begin + R * p.trees[ti].tiles_per_rate) ++ti;

    while (tile < tile_end) {
        const TreeDev tr = p.trees[ti];
        const int tree_tile_end = min(tile_end, tr.tile_begin + R * tr.tiles_per_rate);
        if (p.want_grad) {
            for (int i = tid; i < tr.n_br; i += NT) s_acc[i] = 0.0;
        }
        long long e_total = 0;
        double logsum = 0.0;
        const ModelT<K, DYN_MODEL> mdl{p, DYN_MODEL ? (int)__ldg(p.dyn + tr.dyn_off + dyn_slot(tr.NN, K, R)) * MODEL_SLOT : 0};
        const int4* const post_ops = p.ops + 2 * tr.post_off;
        const int4* const pre_ops = p.ops + 2 * tr.pre_off;
        const int* const post_lvl = p.levels + tr.lvl_off;
        const int* const pre_lvl = post_lvl + tr.n_post_lvl + 1;
        int built_rate = -1;

        for (; tile < tree_tile_end; ++tile) {
            const int local = tile - tr.tile_begin;
            const int r = local / tr.tiles_per_rate;
            const long long site0 = (long long)(local - r * tr.tiles_per_rate) * 32;
            const bool valid = site0 + lane < tr.S;
            const double* const tab_r = s_tab;                 // this tree's branch table for rate r, in shared memory
            constexpr long long br_stride = BT;

            __syncthreads();                                   // previous tile done with the shared buffers
            if (r != built_rate) {
                // branch table of (tree, rate r), built here instead of by a separate kernel:
                // e = exp(t * D mu rate), P = U diag(e) Uinv, dP = U diag(D mu rate e) Uinv, plus the
                // row-sum columns (same operation order as build_branch_tables)
                const double* const blv = p.dyn + tr.dyn_off;
                for (int br = tid; br < tr.n_br; br += NT) {
                    double* ev = s_tab + (size_t)br * BT;
                    double* P = ev + 2 * K;
                    double* dP = P + KK1;
                    if (br >= tr.NN - 1) {
#pragma unroll
                        for (int i = 0; i < K; ++i) { ev[i] = 0.0; ev[K + i] = 0.0; }
#pragma unroll
                        for (int n = 0; n <= K; ++n)
#pragma unroll
                            for (int m = 0; m < K; ++m) { P[n * K + m] = (n == K || n == m) ? 1.0 : 0.0; dP[n * K + m] = 0.0; }
                        continue;
                    }
                    const double t = __ldg(blv + br);
                    double em1[K], de[K];
#pragma unroll
                    for (int i = 0; i < K; ++i) {
                        em1[i] = expm1(t * mdl.c(r, i));
                        de[i] = mdl.c(r, i) * exp(t * mdl.c(r, i));
                        ev[i] = em1[i];
                        ev[K + i] = de[i];
                    }
#pragma unroll
                    for (int m = 0; m < K; ++m) {
                        double rs = 0.0, drs = 0.0;
#pragma unroll
                        for (int n = 0; n < K; ++n) {
                            double c = 0.0, dc = 0.0;
#pragma unroll
                            for (int k = 0; k < K; ++k) {
                                c += (mdl.U(m, k) * em1[k]) * mdl.Ui(k, n);
                                dc += (mdl.U(m, k) * de[k]) * mdl.Ui(k, n);
                            }
                            c += (m == n) ? 1.0 : 0.0;
                            P[n * K + m] = c;
                            dP[n * K + m] = dc;
                            rs += c;
                            drs += dc;
                        }
                        P[K * K + m] = rs;
                        dP[K * K + m] = drs;
                    }
                }
                built_rate = r;
            }
            for (int i = tid; i < tr.n_rows * 32; i += NT) {   // this tile's state codes, all leaves
                const int rw = i >> 5, l = i & 31;
                s_code[i] = (site0 + l < tr.S) ? __ldg(tr.codes + (long long)rw * tr.code_stride + site0 + l) : (unsigned char)K;
            }
            if (tid < 32) s_exp[tid] = 0;
            __syncthreads();

            auto leaf_code = [&](int src) -> int { return src >= 0 ? min((int)s_code[src * 32 + lane], K) : K; };
            auto ld_slot = [&](const double* base, int slot, double (&v)[1][K]) {
#pragma unroll
                for (int k = 0; k < K; ++k) v[0][k] = base[(size_t)slot * SLOT + k];
            };
            auto st_slot = [&](double* base, int slot, const double (&v)[1][K]) {
#pragma unroll
                for (int k = 0; k < K; ++k) base[(size_t)slot * SLOT + k] = v[0][k];
            };
            auto ld_vec = [&](const double* g, double (&v)[K]) {
#pragma unroll
                for (int k = 0; k < K; ++k) v[k] = g[k];
            };

            // ------------------------------ post pass ------------------------------
            for (int lv = 0; lv < tr.n_post_lvl; ++lv) {
                const int lo = __ldg(post_lvl + lv), hi = __ldg(post_lvl + lv + 1);
                for (int i = lo + warp; i < hi; i += W) {
                    const int4 o0 = __ldg(post_ops + 2 * i), o1 = __ldg(post_ops + 2 * i + 1);
                    const int flags = o1.y, ka = flags & 3, kb = (flags >> 2) & 3;
                    double Da[1][K], Db[1][K];
                    if (ka == mcp::OPK_LEAF) {
                        ld_vec(tab_r + o0.y * br_stride + 2 * K + leaf_code(o0.x) * K, Da[0]);
                    } else {
                        double L[1][K], z[1][K], e[K];
                        ld_slot(s_post, o0.x, L);
                        ld_vec(tab_r + o0.y * br_stride, e);
                        eig_project<K, 1, false>(mdl, L, e, nullptr, z, z);
                        eig_expand<K, 1>(mdl, z, L, Da);
                    }
                    if (kb == mcp::OPK_LEAF) {
                        ld_vec(tab_r + o0.w * br_stride + 2 * K + leaf_code(o0.z) * K, Db[0]);
                    } else {
                        double L[1][K], z[1][K], e[K];
                        ld_slot(s_post, o0.z, L);
                        ld_vec(tab_r + o0.w * br_stride, e);
                        eig_project<K, 1, false>(mdl, L, e, nullptr, z, z);
                        eig_expand<K, 1>(mdl, z, L, Db);
                    }
                    double cur[1][K];
#pragma unroll
                    for (int k = 0; k < K; ++k) cur[0][k] = Da[0][k] * Db[0][k];
                    const int ex = rescale_pow2<K>(cur[0]);
                    if (ex != 0) atomicAdd(&s_exp[lane], ex);
                    if (flags & mcp::POST_STORE) st_slot(s_post, o1.x, cur);
                    if (flags & mcp::POST_ROOT) {
                        double rootv = mdl.pi(0) * cur[0][0];
#pragma unroll
                        for (int k = 1; k < K; ++k) rootv = fma(mdl.pi(k), cur[0][k], rootv);
                        if (valid) logsum += log(rootv);
                    }
                }
                __syncthreads();
            }
            if (warp == 0 && valid) e_total += s_exp[lane];

            // ------------------------------ gradient pass ------------------------------
            if (p.want_grad) {
                for (int lv = 0; lv < tr.n_pre_lvl; ++lv) {
                    const int lo = __ldg(pre_lvl + lv), hi = __ldg(pre_lvl + lv + 1);
                    for (int i = lo + warp; i < hi; i += W) {
                        const int4 o0 = __ldg(pre_ops + 2 * i), o1 = __ldg(pre_ops + 2 * i + 1);
                        const int flags = o1.y;
                        const int a_br = o0.y, b_br = o0.w;
                        const bool ai = (flags & 3) == mcp::OPK_MEM, bi = ((flags >> 2) & 3) == mcp::OPK_MEM;
                        double pm[1][K];
                        if (((flags >> 8) & 3) == mcp::PREM_ROOT) {
#pragma unroll
                            for (int k = 0; k < K; ++k) pm[0][k] = mdl.pi(k);
                        } else {
                            ld_slot(s_pre, o1.x, pm);
                        }
                        double ea[K], ebv[K];
                        double Da[1][K], Ya[1][K], Db[1][K], Yb[1][K];
                        if (ai) {
                            double L[1][K], z[1][K], zd[1][K];
                            ld_slot(s_post, o0.x, L);
                            ld_vec(tab_r + a_br * br_stride, ea);
                            eig_project<K, 1, true>(mdl, L, ea, tab_r + a_br * br_stride + K, z, zd);
                            eig_expand<K, 1>(mdl, z, L, Da);
                            eig_expand0<K, 1>(mdl, zd, Ya);
                        } else {
                            const double* t = tab_r + a_br * br_stride + 2 * K + leaf_code(o0.x) * K;
                            ld_vec(t, Da[0]);
                            ld_vec(t + KK1, Ya[0]);
                        }
                        if (bi) {
                            double L[1][K], z[1][K], zd[1][K];
                            ld_slot(s_post, o0.z, L);
                            ld_vec(tab_r + b_br * br_stride, ebv);
                            eig_project<K, 1, true>(mdl, L, ebv, tab_r + b_br * br_stride + K, z, zd);
                            eig_expand<K, 1>(mdl, z, L, Db);
                            eig_expand0<K, 1>(mdl, zd, Yb);
                        } else {
                            const double* t = tab_r + b_br * br_stride + 2 * K + leaf_code(o0.z) * K;
                            ld_vec(t, Db[0]);
                            ld_vec(t + KK1, Yb[0]);
                        }
                        double qa[1][K], qb[1][K];
                        double den = 0.0, na = 0.0, nb = 0.0;
#pragma unroll
                        for (int k = 0; k < K; ++k) {
                            qa[0][k] = pm[0][k] * Db[0][k];
                            qb[0][k] = pm[0][k] * Da[0][k];
                            den = fma(qa[0][k], Da[0][k], den);
                            na = fma(qa[0][k], Ya[0][k], na);
                            nb = fma(qb[0][k], Yb[0][k], nb);
                        }
                        const double inv = fast_rcp(den);
                        const double red = warp_pair_reduce(valid ? na * inv : 0.0, valid ? nb * inv : 0.0, lane);
                        if (lane == 0) atomicAdd(&s_acc[a_br], red);
                        else if (lane == 16) atomicAdd(&s_acc[b_br], red);
                        if (((flags >> 10) & 3) != mcp::OUT_NONE) {
                            double pa[1][K];
                            eig_transposed<K, 1>(mdl, qa, ea, pa);
                            rescale_pow2<K>(pa[0]);
                            st_slot(s_pre, o1.z, pa);
                        }
                        if (((flags >> 12) & 3) != mcp::OUT_NONE) {
                            double pb[1][K];
                            eig_transposed<K, 1>(mdl, qb, ebv, pb);
                            rescale_pow2<K>(pb[0]);
                            st_slot(s_pre, o1.w, pb);
                        }
                    }
                    __syncthreads();
                }
            }
        }  // tiles of this tree

        for (int off = 16; off > 0; off >>= 1) {
            e_total += __shfl_xor_sync(0xffffffffu, e_total, off);
            logsum += __shfl_xor_sync(0xffffffffu, logsum, off);
        }
        if (lane == 0) { s_e[warp] = e_total; s_l[warp] = logsum; }
        __syncthreads();
        if (tid == 0) {
            long long es = 0;
            double ls = 0.0;
            for (int w = 0; w < W; ++w) { es += s_e[w]; ls += s_l[w]; }
            p.rows_ll[row].esum = es;
            p.rows_ll[row].logsum = ls;
        }
        if (p.want_grad) {
            double* dst = p.rows + (long long)row * p.row_stride;
            for (int i = tid; i < tr.n_br; i += NT) dst[i] = s_acc[i];
        }
        __syncthreads();
        ++row;
        ++ti;
    }

    // ---- fused final reduction: the last CTA to finish sums the accumulator rows in fixed order
    // and writes [logL, grad] per tree to p.out (device memory, or pinned host memory for the
    // synchronous entry points: no separate kernel, no device-to-host copy) ----
    __threadfence();
    __syncthreads();
    if (tid == 0) s_ticket = atomicAdd(p.done_counter, 1u);
    __syncthreads();
    if (s_ticket == gridDim.x - 1) {
        __threadfence();
        // All 8 warps take part: warp w sums rows row_lo + w, row_lo + w + W, ... (lanes across the
        // branches, so a warp reads one contiguous run per row), the per-warp partial sums meet in
        // shared memory and are added in warp order.  Fixed order, hence reproducible; ~10x less
        // latency than one thread per output walking all rows (rows = CTAs of the launch).
        double* const s_fin = s_tab;               // W x NN doubles; the tile buffers are free now
        for (int t = 0; t < p.T; ++t) {
            const TreeDev tr = p.trees[t];
            double* o = p.out + tr.out_off;
            const double* d = p.dyn + tr.dyn_off;
            const double* hdr = d + dyn_prior(tr.NN, K, R);
            const bool prior = hdr[0] != 0.0;
            PriorSums ps{0.0, 0.0};
            if (prior) ps = prior_block_sums(d + dyn_blv(tr.NN), hdr + 4, tr.NN - 1, tid, NT, s_prior);
            {
                long long es = 0;
                double ls = 0.0;
                for (int rw = tr.row_lo + tid; rw < tr.row_hi; rw += NT) {
                    es += __ldcg(&p.rows_ll[rw].esum);
                    ls += __ldcg(&p.rows_ll[rw].logsum);
                }
                for (int off = 16; off > 0; off >>= 1) {
                    es += __shfl_xor_sync(0xffffffffu, es, off);
                    ls += __shfl_xor_sync(0xffffffffu, ls, off);
                }
                if (lane == 0) { s_e[warp] = es; s_l[warp] = ls; }
            }
            if (p.want_grad) {
                const int nb = tr.NN - 1;
                for (int j0 = lane; j0 < nb; j0 += 128) {
                    double acc[4] = {0.0, 0.0, 0.0, 0.0};
                    for (int rw = tr.row_lo + warp; rw < tr.row_hi; rw += W) {
                        const double* rp = p.rows + (long long)rw * p.row_stride + j0;
#pragma unroll
                        for (int u = 0; u < 4; ++u)
                            if (j0 + 32 * u < nb) acc[u] += __ldcg(rp + 32 * u);
                    }
#pragma unroll
                    for (int u = 0; u < 4; ++u)
                        if (j0 + 32 * u < nb) s_fin[warp * nb + j0 + 32 * u] = acc[u];
                }
            }
            __syncthreads();
            for (int j = tid; j < tr.NN; j += NT) {
                double v = 0.0;
                if (j == 0) {
                    long long es = 0;
                    double ls = 0.0;
                    for (int w = 0; w < W; ++w) { es += s_e[w]; ls += s_l[w]; }
                    v = (double)es * 0.693147180559945309417232121458 + ls;
                } else if (p.want_grad) {
                    for (int w = 0; w < W; ++w) v += s_fin[w * (tr.NN - 1) + (j - 1)];
                }
                if (prior && (j == 0 || p.want_grad)) v += prior_term(hdr, d + dyn_blv(tr.NN), hdr + 4, ps, j);
                o[j] = v;
            }
            __syncthreads();                       // before the next tree reuses s_fin / s_e / s_l
        }
        if (tid == 0) *p.done_counter = 0;   // ready for the next launch
    }
}

// --------------------------------------------------------------------------------------------
// kernel 2b: generic state count (6 < K <= KMAX_GENERIC), runtime K.  Same op program, same scratch
// layout and accumulator rows as the templated kernel, but dense-table arithmetic straight from the
// branch table (P / dP columns are stored for every branch) and per-thread vectors in local memory.
// Correctness path for large alphabets (e.g. 20-state protein models); not tuned.
// --------------------------------------------------------------------------------------------
constexpr int KMAX_GENERIC = 32;

__global__ void __launch_bounds__(128) felsenstein_walk_generic(const WalkParams p, const int K) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    __shared__ long long s_e[8];
    __shared__ double s_l[8];
    double* const s_acc = reinterpret_cast<double*>(smem_raw);

    const int tid = threadIdx.x, TW = blockDim.x, lane = tid & 31, warp = tid >> 5;
    const int q = p.n_tiles / gridDim.x, rem = p.n_tiles - q * gridDim.x;
    int tile = blockIdx.x * q + min((int)blockIdx.x, rem);
    const int tile_end = tile + q + ((int)blockIdx.x < rem ? 1 : 0);
    if (tile >= tile_end) return;

    double* const slots = p.scratch + (long long)blockIdx.x * p.scratch_per_cta + (long long)tid * K;
    const long long slot_stride = (long long)TW * K;
    double* const stack = slots + (long long)p.n_slots * slot_stride;
    int row = p.cta_row_base[blockIdx.x];
    const int R = p.R;
    const int BT = bt_size(K), KK1 = K * (K + 1);

    int ti = 0;
    while (ti < p.T - 1 && tile >= p.trees[ti].tile_begin + R * p.trees[ti].tiles_per_rate) ++ti;

    // out = T L for a stored/register operand, or the code column for a leaf
    auto down = [&](const double* __restrict__ tab, const double* L, double* out) {
        for (int s = 0; s < K; ++s) out[s] = 0.0;
        for (int j = 0; j < K; ++j) {
            const double lj = L[j];
            const double* col = tab + j * K;
            for (int s = 0; s < K; ++s) out[s] = fma(__ldg(col + s), lj, out[s]);
        }
    };
    auto rescale = [&](double* v) -> int {
        unsigned m = 0;
        for (int k = 0; k < K; ++k) m = max(m, (unsigned)__double2hiint(v[k]) & 0x7fffffffu);
        const int e = (int)(m >> 20);
        if (e == 0 || e == 0x7ff) return 0;
        const double sc = __hiloint2double((2046 - e) << 20, 0);
        for (int k = 0; k < K; ++k) v[k] *= sc;
        return e - 1023;
    };

    while (tile < tile_end) {
        const TreeDev tr = p.trees[ti];
        const int tree_tile_end = min(tile_end, tr.tile_begin + R * tr.tiles_per_rate);
        if (p.want_grad) {
            for (int i = tid; i < tr.n_br; i += TW) s_acc[i] = 0.0;
        }
        __syncthreads();
        long long e_total = 0;
        double logsum = 0.0;
        const double* const pi = p.dyn + tr.dyn_off + dyn_pi(tr.NN, K, R);
        const int4* const post_ops = p.ops + 2 * tr.post_off;
        const int4* const pre_ops = p.ops + 2 * tr.pre_off;

        for (; tile < tree_tile_end; ++tile) {
            const int local = tile - tr.tile_begin;
            const int r = local / tr.tiles_per_rate;
            const long long site = (long long)(local - r * tr.tiles_per_rate) * TW + tid;
            const bool valid = site < tr.S;
            const unsigned char* const codes = tr.codes + (valid ? site : 0);
            const double* const tab_r = p.btab + tr.btab_off + (long long)r * BT + 2 * K;   // P columns of (branch 0, rate r)
            const long long br_stride = (long long)R * BT;
            auto leaf_code = [&](int src) -> int {
                int code = (valid && src >= 0) ? (int)__ldg(codes + (long long)src * tr.code_stride) : K;
                return min(code, K);
            };

            double cur[KMAX_GENERIC], Da[KMAX_GENERIC], Db[KMAX_GENERIC], L[KMAX_GENERIC];
            for (int k = 0; k < K; ++k) cur[k] = 1.0;
            int e_col = 0;
            for (int i = 0; i < tr.n_post; ++i) {
                const int4 o0 = __ldg(post_ops + 2 * i), o1 = __ldg(post_ops + 2 * i + 1);
                const int flags = o1.y, ka = flags & 3, kb = (flags >> 2) & 3;
                const double* ta = tab_r + o0.y * br_stride;
                const double* tb = tab_r + o0.w * br_stride;
                if (ka == mcp::OPK_LEAF) {
                    const double* col = ta + leaf_code(o0.x) * K;
                    for (int s = 0; s < K; ++s) Da[s] = __ldg(col + s);
                } else if (ka == mcp::OPK_REG) {
                    down(ta, cur, Da);
                } else {
                    for (int k = 0; k < K; ++k) L[k] = __ldcg(slots + o0.x * slot_stride + k);
                    down(ta, L, Da);
                }
                if (kb == mcp::OPK_LEAF) {
                    const double* col = tb + leaf_code(o0.z) * K;
                    for (int s = 0; s < K; ++s) Db[s] = __ldg(col + s);
                } else if (kb == mcp::OPK_REG) {
                    down(tb, cur, Db);
                } else {
                    for (int k = 0; k < K; ++k) L[k] = __ldcg(slots + o0.z * slot_stride + k);
                    down(tb, L, Db);
                }
                for (int k = 0; k < K; ++k) cur[k] = Da[k] * Db[k];
                e_col += rescale(cur);
                if (flags & mcp::POST_STORE)
                    for (int k = 0; k < K; ++k) __stcg(slots + o1.x * slot_stride + k, cur[k]);
            }
            {
                double rootv = 0.0;
                for (int k = 0; k < K; ++k) rootv = fma(__ldg(pi + k), cur[k], rootv);
                if (valid) {
                    logsum += log(rootv);
                    e_total += e_col;
                }
            }

            if (p.want_grad) {
                double pm[KMAX_GENERIC], Ya[KMAX_GENERIC], Yb[KMAX_GENERIC];
                for (int i = 0; i < tr.n_pre; ++i) {
                    const int4 o0 = __ldg(pre_ops + 2 * i), o1 = __ldg(pre_ops + 2 * i + 1);
                    const int flags = o1.y;
                    const int a_br = o0.y, b_br = o0.w;
                    const bool ai = (flags & 3) == mcp::OPK_MEM, bi = ((flags >> 2) & 3) == mcp::OPK_MEM;
                    const int mk = (flags >> 8) & 3;
                    if (mk == mcp::PREM_ROOT) { for (int k = 0; k < K; ++k) pm[k] = __ldg(pi + k); }
                    else if (mk == mcp::PREM_REG) { for (int k = 0; k < K; ++k) pm[k] = cur[k]; }
                    else { for (int k = 0; k < K; ++k) pm[k] = __ldcg(stack + o1.x * slot_stride + k); }
                    const double* ta = tab_r + a_br * br_stride;
                    const double* tb = tab_r + b_br * br_stride;
                    if (ai) {
                        for (int k = 0; k < K; ++k) L[k] = __ldcg(slots + o0.x * slot_stride + k);
                        down(ta, L, Da);
                        down(ta + KK1, L, Ya);
                    } else {
                        const double* col = ta + leaf_code(o0.x) * K;
                        for (int s = 0; s < K; ++s) { Da[s] = __ldg(col + s); Ya[s] = __ldg(col + KK1 + s); }
                    }
                    if (bi) {
                        for (int k = 0; k < K; ++k) L[k] = __ldcg(slots + o0.z * slot_stride + k);
                        down(tb, L, Db);
                        down(tb + KK1, L, Yb);
                    } else {
                        const double* col = tb + leaf_code(o0.z) * K;
                        for (int s = 0; s < K; ++s) { Db[s] = __ldg(col + s); Yb[s] = __ldg(col + KK1 + s); }
                    }
                    double den = 0.0, na = 0.0, nb = 0.0;
                    for (int k = 0; k < K; ++k) {
                        const double qa = pm[k] * Db[k], qb = pm[k] * Da[k];
                        den = fma(qa, Da[k], den);
                        na = fma(qa, Ya[k], na);
                        nb = fma(qb, Yb[k], nb);
                        Ya[k] = qa;     // Ya / Yb now hold qa / qb for the transposed products
                        Yb[k] = qb;
                    }
                    const double inv = 1.0 / den;
                    const double red = warp_pair_reduce(valid ? na * inv : 0.0, valid ? nb * inv : 0.0, lane);
                    if (lane == 0) atomicAdd(&s_acc[a_br], red);
                    else if (lane == 16) atomicAdd(&s_acc[b_br], red);

                    const int a_out = (flags >> 10) & 3, b_out = (flags >> 12) & 3;
                    // pre[child][j] = sum_s P[s][j] q[s] = column j of the table dotted with q
                    if (b_out != mcp::OUT_NONE) {
                        for (int j = 0; j < K; ++j) {
                            double acc = 0.0;
                            for (int s = 0; s < K; ++s) acc = fma(__ldg(tb + j * K + s), Yb[s], acc);
                            Db[j] = acc;
                        }
                        rescale(Db);
                        if (b_out == mcp::OUT_PUSH)
                            for (int k = 0; k < K; ++k) __stcg(stack + o1.w * slot_stride + k, Db[k]);
                    }
                    if (a_out != mcp::OUT_NONE) {
                        for (int j = 0; j < K; ++j) {
                            double acc = 0.0;
                            for (int s = 0; s < K; ++s) acc = fma(__ldg(ta + j * K + s), Ya[s], acc);
                            Da[j] = acc;
                        }
                        rescale(Da);
                        if (a_out == mcp::OUT_PUSH)
                            for (int k = 0; k < K; ++k) __stcg(stack + o1.z * slot_stride + k, Da[k]);
                    }
                    if (a_out == mcp::OUT_KEEP) { for (int k = 0; k < K; ++k) cur[k] = Da[k]; }
                    else if (b_out == mcp::OUT_KEEP) { for (int k = 0; k < K; ++k) cur[k] = Db[k]; }
                }
            }
        }  // tiles of this tree

        for (int off = 16; off > 0; off >>= 1) {
            e_total += __shfl_xor_sync(0xffffffffu, e_total, off);
            logsum += __shfl_xor_sync(0xffffffffu, logsum, off);
        }
        if (lane == 0) { s_e[warp] = e_total; s_l[warp] = logsum; }
        __syncthreads();
        if (tid == 0) {
            long long es = 0;
            double ls = 0.0;
            for (int w = 0; w < (TW + 31) / 32; ++w) { es += s_e[w]; ls += s_l[w]; }
            p.rows_ll[row].esum = es;
            p.rows_ll[row].logsum = ls;
        }
        if (p.want_grad) {
            double* dst = p.rows + (long long)row * p.row_stride;
            for (int i = tid; i < tr.n_br; i += TW) dst[i] = s_acc[i];
        }
        __syncthreads();
        ++row;
        ++ti;
    }
}

// --------------------------------------------------------------------------------------------
// kernel 3: fixed-order reduction of the accumulator rows -> [logL, grad] per tree
// --------------------------------------------------------------------------------------------
constexpr int FIN_J = 32, FIN_G = 8;   // outputs per block x row groups (blockDim = 32 x 8)
__global__ void finalize_results(const TreeDev* __restrict__ trees, const double* __restrict__ rows,
                                 long long row_stride, const LLRow* __restrict__ rows_ll,
                                 double* __restrict__ out, int want_grad, const double* __restrict__ dyn, int K, int R) {
    // Block = 32 consecutive outputs x 8 row groups: group g sums rows row_lo + g, row_lo + g + 8, ...
    // (a warp reads 32 consecutive doubles of one row), the 8 partial sums meet in shared memory and
    // are added in group order: fixed order, 8x shorter dependent chain than one thread per output.
    __shared__ double s_red[2 * FIN_J * FIN_G];
    __shared__ double s_g[FIN_G][FIN_J];
    __shared__ long long s_es[FIN_G];
    const TreeDev tr = trees[blockIdx.y];
    if ((long long)blockIdx.x * FIN_J >= tr.NN) return;   // whole block idle (batch of unequal trees)
    const int jl = threadIdx.x, g = threadIdx.y, tid = g * FIN_J + jl;
    const double* d = dyn + tr.dyn_off;
    const double* hdr = d + dyn_prior(tr.NN, K, R);
    const bool prior = hdr[0] != 0.0;
    PriorSums ps{0.0, 0.0};
    if (prior) ps = prior_block_sums(d + dyn_blv(tr.NN), hdr + 4, tr.NN - 1, tid, FIN_J * FIN_G, s_red);
    const int j = blockIdx.x * FIN_J + jl;
    double v = 0.0;
    long long es = 0;
    if (j < tr.NN) {
        if (j == 0) {
            for (int rw = tr.row_lo + g; rw < tr.row_hi; rw += FIN_G) { es += rows_ll[rw].esum; v += rows_ll[rw].logsum; }
        } else if (want_grad) {
            for (int rw = tr.row_lo + g; rw < tr.row_hi; rw += FIN_G) v += rows[(long long)rw * row_stride + (j - 1)];
        }
    }
    s_g[g][jl] = v;
    if (j == 0) s_es[g] = es;
    __syncthreads();
    if (g != 0 || j >= tr.NN) return;
    v = 0.0;
    for (int gg = 0; gg < FIN_G; ++gg) v += s_g[gg][jl];
    if (j == 0) {
        es = 0;
        for (int gg = 0; gg < FIN_G; ++gg) es += s_es[gg];
        v += (double)es * 0.693147180559945309417232121458;
    }
    if (prior && (j == 0 || want_grad)) v += prior_term(hdr, d + dyn_blv(tr.NN), hdr + 4, ps, j);
    out[tr.out_off + j] = v;
}

// --------------------------------------------------------------------------------------------
// host side
// --------------------------------------------------------------------------------------------
thread_local std::string g_error;

struct DevBuf {
    void* p = nullptr;
    size_t cap = 0;
};
struct PinBuf {
    void* p = nullptr;
    size_t cap = 0;
};

}  // namespace

struct mcp_alignment {
    int K = 0;
    long long S = 0, stride = 0;
    int n_leaves = 0;
    unsigned char* d_codes = nullptr;
    std::vector<int32_t> leaf_nums;
    unsigned long long id = 0;
    // Re-uploads (mcp_alignment_update_codes) run on the context's copy stream so that they overlap
    // evaluations of OTHER alignments; these order them against the evaluations of THIS one.
    cudaEvent_t ev_uploaded = nullptr;
    cudaEvent_t ev_read_done = nullptr;      // recorded after every walk that read d_codes, once `streamed`
    mutable bool upload_pending = false;     // an upload has been enqueued that no evaluation has waited for yet
    mutable bool read_since_upload = false;  // an evaluation reading d_codes was enqueued after the last upload
    mutable bool streamed = false;           // has been re-uploaded at least once: evaluations record ev_read_done
};

struct mcp_ctx {
    int device = 0;
    int sm_count = 0;
    cudaStream_t own_stream = nullptr, stream = nullptr;
    cudaStream_t copy_stream = nullptr;     // alignment re-uploads (overlap with evaluations)
    cudaEvent_t ev_walk_done = nullptr;     // the walk kernel of the last evaluation has finished reading the codes
    cudaEvent_t ev[4] = {nullptr, nullptr, nullptr, nullptr};
    cudaEvent_t ev_staged = nullptr;   // the pinned staging buffers of the last evaluation have been consumed
    bool staged_pending = false;
    std::string error;
    bool pending_async = false;
    int opt_block = 0, opt_ctas_per_sm = 0, opt_cpt = 0;
    int cpt = 1;   // columns per thread of the cached launch
    bool smem_scratch = false;   // partials scratch in shared memory (small-problem latency path)
    bool level_mode = false;     // level-parallel small-tree kernel
    int opt_levels = -1;         // -1 automatic, 0 never, 1 whenever it fits
    int sig_levels = -1;
    int max_rows = 1;
    size_t off_levels = 0;
    int opt_smem_scratch = -1;   // -1 automatic, 0 off, 1 on when it fits
    unsigned long long next_aln_id = 1;

    DevBuf d_topo, d_dyn, d_btab, d_scratch, d_rows, d_rows_ll, d_out, d_counter;
    PinBuf h_topo, h_dyn, h_out, h_model;

    // cached topology
    struct TreeSig {
        unsigned long long aln_id;
        int NN;
        std::vector<int32_t> po, pa;
    };
    std::vector<TreeSig> sig;
    int sig_want_grad = -1, sig_block = 0, sig_K = 0, sig_R = 0, sig_cpt = 0;
    // derived launch state kept with the cached topology
    std::vector<TreeDev> trees;
    std::vector<Schedule> scheds;
    size_t topo_bytes = 0, off_trees = 0, off_ops = 0, off_rowbase = 0;
    int n_tiles = 0, grid = 0, block = 0, n_rows = 0, n_slots = 0, n_stack = 0, max_br = 0;
    long long total_out = 0, total_dyn = 0, total_btab = 0, scratch_per_cta = 0, row_stride = 0;
    size_t smem_bytes = 0;

    mcp_stats stats{};
};

namespace {

int fail(mcp_ctx* ctx, int code, const char* fmt, ...) {
    char buf[1024];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof buf, fmt, ap);
    va_end(ap);
    if (ctx) ctx->error = buf;
    else g_error = buf;
    return code;
}

#define CUDA_TRY(ctx, expr)                                                                      \
    do {                                                                                         \
        cudaError_t _e = (expr);                                                                 \
        if (_e != cudaSuccess)                                                                   \
            return fail(ctx, MCP_ERR_CUDA, "%s failed: %s (%s:%d)", #expr, cudaGetErrorString(_e), \
                        __FILE__, __LINE__);                                                     \
    } while (0)

int ensure_dev(mcp_ctx* ctx, DevBuf& b, size_t bytes) {
    if (bytes <= b.cap) return 0;
    if (ctx->pending_async) {
        CUDA_TRY(ctx, cudaStreamSynchronize(ctx->stream));
        ctx->pending_async = false;
    }
    if (b.p) CUDA_TRY(ctx, cudaFree(b.p));
    b.p = nullptr;
    b.cap = 0;
    size_t want = bytes + bytes / 8 + 256;
    cudaError_t e = cudaMalloc(&b.p, want);
    if (e != cudaSuccess) {
        cudaGetLastError();
        want = bytes;
        e = cudaMalloc(&b.p, want);
    }
    if (e != cudaSuccess) {
        b.p = nullptr;
        return fail(ctx, MCP_ERR_CUDA, "cudaMalloc of %zu bytes failed: %s", want, cudaGetErrorString(e));
    }
    b.cap = want;
    return 0;
}
int ensure_pin(mcp_ctx* ctx, PinBuf& b, size_t bytes) {
    if (bytes <= b.cap) return 0;
    if (b.p) CUDA_TRY(ctx, cudaFreeHost(b.p));
    b.p = nullptr;
    b.cap = 0;
    size_t want = bytes + bytes / 4 + 4096;
    CUDA_TRY(ctx, cudaMallocHost(&b.p, want));
    b.cap = want;
    return 0;
}

// cudaFuncSetAttribute is only needed when a kernel's dynamic shared memory grows.  The attribute
// belongs to the (device, function) pair and is shared by every context of the process, so the
// high-water marks are process-global and only ever raised.
template <class Kern>
int ensure_smem_attr(mcp_ctx* ctx, Kern kern, size_t smem) {
    struct Mark { int device; const void* fn; size_t bytes; };
    static std::mutex mu;
    static std::vector<Mark> marks;
    std::lock_guard<std::mutex> lock(mu);
    for (auto& m : marks)
        if (m.device == ctx->device && m.fn == (const void*)kern) {
            if (m.bytes >= smem) return 0;
            CUDA_TRY(ctx, cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
            m.bytes = smem;
            return 0;
        }
    CUDA_TRY(ctx, cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    marks.push_back({ctx->device, (const void*)kern, smem});
    return 0;
}
template <int K, int CPT, bool DYN, bool SSCR>
int launch_walk_inst(mcp_ctx* ctx, const WalkParams& wp) {
    int e = ensure_smem_attr(ctx, felsenstein_walk<K, CPT, DYN, SSCR>, ctx->smem_bytes);
    if (e) return e;
    felsenstein_walk<K, CPT, DYN, SSCR><<<ctx->grid, ctx->block, ctx->smem_bytes, ctx->stream>>>(wp);
    CUDA_TRY(ctx, cudaGetLastError());
    return 0;
}
template <int K>
int launch_walk(mcp_ctx* ctx, const WalkParams& wp, bool dyn_model) {
    if (ctx->smem_scratch && !dyn_model && ctx->cpt == 1) return launch_walk_inst<K, 1, false, true>(ctx, wp);
    if (ctx->cpt == 2) return dyn_model ? launch_walk_inst<K, 2, true, false>(ctx, wp) : launch_walk_inst<K, 2, false, false>(ctx, wp);
    return dyn_model ? launch_walk_inst<K, 1, true, false>(ctx, wp) : launch_walk_inst<K, 1, false, false>(ctx, wp);
}
template <int K, int CPT>
int occupancy_inst(mcp_ctx* ctx, int block, size_t smem, bool sscr, int* out) {
    int e;
    if (sscr) {
        if ((e = ensure_smem_attr(ctx, felsenstein_walk<K, 1, false, true>, smem))) return e;
        CUDA_TRY(ctx, cudaOccupancyMaxActiveBlocksPerMultiprocessor(out, felsenstein_walk<K, 1, false, true>, block, smem));
        return 0;
    }
    // the dynamic-model variant needs a few more registers: size the persistent grid for it
    if ((e = ensure_smem_attr(ctx, felsenstein_walk<K, CPT, true, false>, smem))) return e;
    if ((e = ensure_smem_attr(ctx, felsenstein_walk<K, CPT, false, false>, smem))) return e;
    int o1 = 0, o2 = 0;
    CUDA_TRY(ctx, cudaOccupancyMaxActiveBlocksPerMultiprocessor(&o1, felsenstein_walk<K, CPT, true, false>, block, smem));
    CUDA_TRY(ctx, cudaOccupancyMaxActiveBlocksPerMultiprocessor(&o2, felsenstein_walk<K, CPT, false, false>, block, smem));
    *out = std::min(o1, o2);
    return 0;
}
template <int K>
int occupancy_for(mcp_ctx* ctx, int block, int cpt, size_t smem, bool sscr, int* out) {
    return cpt == 2 ? occupancy_inst<K, 2>(ctx, block, smem, false, out) : occupancy_inst<K, 1>(ctx, block, smem, sscr, out);
}
template <int K>
int occupancy_levels(mcp_ctx* ctx, int block, size_t smem, int* out) {
    int e;
    if ((e = ensure_smem_attr(ctx, felsenstein_walk_levels<K, true>, smem))) return e;
    if ((e = ensure_smem_attr(ctx, felsenstein_walk_levels<K, false>, smem))) return e;
    int o1 = 0, o2 = 0;
    CUDA_TRY(ctx, cudaOccupancyMaxActiveBlocksPerMultiprocessor(&o1, felsenstein_walk_levels<K, true>, block, smem));
    CUDA_TRY(ctx, cudaOccupancyMaxActiveBlocksPerMultiprocessor(&o2, felsenstein_walk_levels<K, false>, block, smem));
    *out = std::min(o1, o2);
    return 0;
}
template <int K>
int launch_levels(mcp_ctx* ctx, const WalkParams& wp, bool dyn_model) {
    if (dyn_model) felsenstein_walk_levels<K, true><<<ctx->grid, ctx->block, ctx->smem_bytes, ctx->stream>>>(wp);
    else felsenstein_walk_levels<K, false><<<ctx->grid, ctx->block, ctx->smem_bytes, ctx->stream>>>(wp);
    CUDA_TRY(ctx, cudaGetLastError());
    return 0;
}
size_t walk_smem_bytes(int K, int max_br, int want_grad, int block, int cpt) {
    const size_t acc = (want_grad && !grad_in_l2(K)) ? (((size_t)max_br * 8 + 15) & ~(size_t)15) : 0;
    return acc + (size_t)3 * CH * 32 + (size_t)2 * CH * 2 * 2 * K * 8 + (size_t)2 * CH * 32 +
           (size_t)2 * CH * 2 * 2 * K * (K + 1) * 8 + (size_t)2 * CH * 2 * block * cpt;
}

#define MCP_DISPATCH_K(K, CALL)                      \
    switch (K) {                                     \
        case 2: { constexpr int KK = 2; CALL; break; } \
        case 3: { constexpr int KK = 3; CALL; break; } \
        case 4: { constexpr int KK = 4; CALL; break; } \
        case 5: { constexpr int KK = 5; CALL; break; } \
        case 6: { constexpr int KK = 6; CALL; break; } \
        default: rc = MCP_ERR_UNSUPPORTED;           \
    }

bool k_templated(int K) { return K >= 2 && K <= 6; }
bool k_supported(int K) { return K >= 2 && K <= KMAX_GENERIC; }

struct BatchArgs {
    int T;
    const mcp_alignment* const* alns;
    const int32_t* NN;
    const int32_t* const* po;
    const int32_t* const* pa;
    const double* const* blv;
    const double* const* U;
    const double* const* D;
    const double* const* Uinv;
    const double* mu;
    const double* const* rates;
    int R;
    const double* const* pi;
    int want_grad;
    // optional branch-length prior (mcp_eval_posterior); applies to every tree of the batch
    int prior_kind = MCP_PRIOR_NONE;
    const double* prior_params = nullptr;
};

// (Re)build schedules, tile/row assignment and the topology upload if anything structural changed.
int prepare_topology(mcp_ctx* ctx, const BatchArgs& a, int K, bool* rebuilt) {
    const int T = a.T, R = a.R;
    // choose the tile width first: it is part of the signature
    long long total_cols = 0;
    for (int t = 0; t < T; ++t) total_cols += a.alns[t]->S * R;
    int block = ctx->opt_block;
    if (block <= 0) {
        block = 256;
        while (block > 32 && (total_cols + block - 1) / block < 6LL * ctx->sm_count) block >>= 1;
    }
    // Two columns per thread amortise the per-op overhead (descriptor decode, constant loads, warp
    // reduction) once there is enough work to fill the GPU.  Measured: 1.2x at K = 2; at K = 4 the
    // doubled register state costs more occupancy than it saves (profiles/r1_walk_notes.md).
    int cpt = ctx->opt_cpt;
    if (cpt <= 0) cpt = (K <= 3 && total_cols / (2LL * block) >= 12LL * ctx->sm_count) ? 2 : 1;
    if (!k_templated(K)) {   // generic-K kernel: one column per thread, at most 128 threads per CTA
        cpt = 1;
        if (block > 128) block = 128;
    }
    // Small inputs (a few one-warp tiles per SM): level-parallel kernel, a tile is 32 columns wide
    // and is worked on by all 8 warps of a 256-thread CTA.
    const bool try_levels = k_templated(K) && ctx->opt_levels != 0 &&
                            (ctx->opt_levels == 1 || (ctx->opt_block == 0 && total_cols <= 32LL * 4 * ctx->sm_count));
    bool same = (int)ctx->sig.size() == T && ctx->sig_want_grad == a.want_grad && ctx->sig_block == block &&
                ctx->sig_K == K && ctx->sig_R == R && ctx->sig_cpt == cpt && ctx->sig_levels == (int)try_levels;
    for (int t = 0; same && t < T; ++t) {
        const auto& s = ctx->sig[t];
        same = s.aln_id == a.alns[t]->id && s.NN == a.NN[t] &&
               std::memcmp(s.po.data(), a.po[t], sizeof(int32_t) * a.NN[t]) == 0 &&
               std::memcmp(s.pa.data(), a.pa[t], sizeof(int32_t) * a.NN[t]) == 0;
    }
    *rebuilt = !same;
    if (same) return 0;

    ctx->sig.clear();
    long long n_ops = 0, out_off = 0, dyn_off = 0, btab_off = 0, n_lvl_ints = 0;
    int tile_cursor = 0, n_slots = 1, n_stack = 1, max_br = 1, max_rows = 1;
    std::vector<int32_t> leaf_row;
    bool level_mode = try_levels;
    auto build_all = [&](bool by_levels, int tile_w) -> int {
    ctx->sig.clear();
    ctx->scheds.assign(T, Schedule());
    ctx->trees.assign(T, TreeDev());
    n_ops = out_off = dyn_off = btab_off = n_lvl_ints = 0;
    tile_cursor = 0; n_slots = 1; n_stack = 1; max_br = 1; max_rows = 1;
    for (int t = 0; t < T; ++t) {
        const mcp_alignment* al = a.alns[t];
        const int NN = a.NN[t];
        if (NN < 2) return fail(ctx, MCP_ERR_ARG, "tree %d: NN must be >= 2", t);
        leaf_row.assign(NN, -1);
        for (int i = 0; i < al->n_leaves; ++i) {
            int num = al->leaf_nums[i];
            if (num >= 1 && num <= NN) leaf_row[num - 1] = i;
        }
        std::string err = mcp::build_schedule(NN, a.po[t], a.pa[t], leaf_row.data(), a.want_grad != 0, ctx->scheds[t], by_levels);
        if (!err.empty()) return fail(ctx, MCP_ERR_ARG, "tree %d: %s", t, err.c_str());
        const Schedule& sc = ctx->scheds[t];
        TreeDev& td = ctx->trees[t];
        td.post_off = n_ops;
        n_ops += (long long)sc.post.size();
        td.pre_off = n_ops;
        n_ops += (long long)sc.pre.size();
        td.n_post = (int)sc.post.size();
        td.n_pre = (int)sc.pre.size();
        td.NN = NN;
        td.n_br = sc.n_dnodes;
        td.codes = al->d_codes;
        td.S = al->S;
        td.code_stride = al->stride;
        td.out_off = out_off;
        out_off += NN;
        td.dyn_off = dyn_off;
        dyn_off += dyn_size(NN, K, R);
        td.btab_off = btab_off;
        btab_off += (long long)sc.n_dnodes * R * bt_size(K);
        td.tiles_per_rate = (int)((al->S + (long long)tile_w - 1) / (long long)tile_w);
        td.n_rows = al->n_leaves;
        td.lvl_off = (int)n_lvl_ints;
        td.n_post_lvl = sc.post_levels.empty() ? 0 : (int)sc.post_levels.size() - 1;
        td.n_pre_lvl = sc.pre_levels.empty() ? 0 : (int)sc.pre_levels.size() - 1;
        n_lvl_ints += (long long)sc.post_levels.size() + (long long)sc.pre_levels.size();
        max_rows = std::max(max_rows, al->n_leaves);
        td.tile_begin = tile_cursor;
        long long nt = (long long)td.tiles_per_rate * R;
        if (tile_cursor + nt > 0x7fffffffLL) return fail(ctx, MCP_ERR_ARG, "too many column tiles");
        tile_cursor += (int)nt;
        n_slots = std::max(n_slots, sc.n_slots);
        n_stack = std::max(n_stack, sc.n_stack);
        max_br = std::max(max_br, sc.n_dnodes);
        mcp_ctx::TreeSig sg;
        sg.aln_id = al->id;
        sg.NN = NN;
        sg.po.assign(a.po[t], a.po[t] + NN);
        sg.pa.assign(a.pa[t], a.pa[t] + NN);
        ctx->sig.push_back(std::move(sg));
    }
    return 0;
    };  // build_all
    int be = 0;
    if (level_mode) {
        if ((be = build_all(true, 32))) { ctx->sig.clear(); return be; }
        const size_t need = LevelSmem::total(max_br, a.want_grad ? 1 : 0, max_rows, n_slots, a.want_grad ? n_stack : 0, K);
        if (need > 160 * 1024) level_mode = false;     // tree too large for the shared-memory path
    }
    if (!level_mode && (be = build_all(false, block * cpt))) { ctx->sig.clear(); return be; }
    const int sig_block = block, sig_cpt = cpt;
    if (level_mode) { block = 256; cpt = 1; }
    ctx->level_mode = level_mode;
    ctx->sig_levels = (int)try_levels;
    ctx->max_rows = max_rows;
    ctx->sig_want_grad = a.want_grad;
    ctx->sig_block = sig_block;
    ctx->sig_cpt = sig_cpt;
    ctx->cpt = cpt;
    ctx->sig_K = K;
    ctx->sig_R = R;
    ctx->n_tiles = tile_cursor;
    ctx->block = block;
    ctx->n_slots = n_slots;
    ctx->n_stack = a.want_grad ? n_stack : 0;
    ctx->max_br = max_br;
    ctx->total_out = out_off;
    ctx->total_dyn = dyn_off;
    ctx->total_btab = btab_off;
    ctx->smem_bytes = level_mode ? LevelSmem::total(max_br, a.want_grad ? 1 : 0, max_rows, n_slots, a.want_grad ? n_stack : 0, K)
                      : k_templated(K) ? walk_smem_bytes(K, max_br, a.want_grad ? 1 : 0, block, cpt)
                                       : (a.want_grad ? (size_t)max_br * sizeof(double) : 0);
    // Small problems: keep the partials scratch in shared memory (latency path).  Automatic when the
    // whole input is a handful of tiles per SM and the scratch of one CTA fits next to the staging
    // buffers.
    {
        const size_t scr_bytes = (size_t)(ctx->n_slots + ctx->n_stack) * block * cpt * K * 8;
        const bool fits = !level_mode && k_templated(K) && cpt == 1 && ctx->smem_bytes + scr_bytes <= 96 * 1024;
        ctx->smem_scratch = fits && (ctx->opt_smem_scratch == 1 ||
                                     (ctx->opt_smem_scratch < 0 && ctx->n_tiles <= 4 * ctx->sm_count));
        if (ctx->smem_scratch) ctx->smem_bytes += scr_bytes;
    }
    if (ctx->smem_bytes > 200 * 1024)
        return fail(ctx, MCP_ERR_UNSUPPORTED, "tree with %d nodes exceeds the shared-memory gradient accumulator", max_br);

    // persistent grid
    int occ = 0, rc = 0;
    if (level_mode) {
        MCP_DISPATCH_K(K, rc = occupancy_levels<KK>(ctx, block, ctx->smem_bytes, &occ));
    } else if (k_templated(K)) {
        MCP_DISPATCH_K(K, rc = occupancy_for<KK>(ctx, block, cpt, ctx->smem_bytes, ctx->smem_scratch, &occ));
    } else {
        if ((rc = ensure_smem_attr(ctx, felsenstein_walk_generic, ctx->smem_bytes))) return rc;
        CUDA_TRY(ctx, cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, felsenstein_walk_generic, block, ctx->smem_bytes));
    }
    if (rc) return rc == MCP_ERR_UNSUPPORTED ? fail(ctx, rc, "no kernel compiled for K = %d states", K) : rc;
    if (occ < 1) return fail(ctx, MCP_ERR_CUDA, "walk kernel does not fit on an SM (block %d, smem %zu)", block, ctx->smem_bytes);
    if (ctx->opt_ctas_per_sm > 0) occ = std::min(occ, ctx->opt_ctas_per_sm);
    ctx->grid = (int)std::min<long long>((long long)ctx->n_tiles, (long long)occ * ctx->sm_count);
    if (ctx->grid < 1) ctx->grid = 1;
    {   // very large trees: fewer persistent CTAs rather than a scratch allocation that cannot succeed
        size_t free_b = 0, total_b = 0;
        const double per_cta = level_mode ? 0.0 : (double)(ctx->n_slots + ctx->n_stack) * block * cpt * K * 8.0;
        if (per_cta * ctx->grid <= (double)ctx->d_scratch.cap) {
            // fits the scratch already held: nothing to allocate, no need to ask the driver
            // (cudaMemGetInfo costs milliseconds on a GPU with many live allocations)
        } else if (cudaMemGetInfo(&free_b, &total_b) == cudaSuccess && per_cta > 0) {
            const double budget = 0.6 * ((double)free_b + (double)ctx->d_scratch.cap);
            if (per_cta * ctx->grid > budget) ctx->grid = (int)std::max(1.0, std::floor(budget / per_cta));
        } else {
            cudaGetLastError();
        }
    }

    // accumulator rows: one per (CTA, tree) pair in CTA order (also tree order)
    std::vector<int32_t> row_base(ctx->grid, 0);
    {
        const int q = ctx->n_tiles / ctx->grid, rem = ctx->n_tiles % ctx->grid;
        int row = 0, ti = 0;
        for (int t = 0; t < T; ++t) ctx->trees[t].row_lo = ctx->trees[t].row_hi = 0;
        std::vector<char> seen(T, 0);
        for (int c = 0; c < ctx->grid; ++c) {
            int t0 = c * q + std::min(c, rem), t1 = t0 + q + (c < rem ? 1 : 0);
            row_base[c] = row;
            int tile = t0;
            while (ti < T - 1 && tile >= ctx->trees[ti].tile_begin + R * ctx->trees[ti].tiles_per_rate) ++ti;
            int tj = ti;
            while (tile < t1) {
                int tend = std::min(t1, ctx->trees[tj].tile_begin + R * ctx->trees[tj].tiles_per_rate);
                if (!seen[tj]) { ctx->trees[tj].row_lo = row; seen[tj] = 1; }
                ++row;
                ctx->trees[tj].row_hi = row;
                tile = tend;
                ++tj;
            }
        }
        ctx->n_rows = row;
    }
    ctx->row_stride = (max_br + 3) & ~3;
    if ((!level_mode && (double)(ctx->n_slots + n_stack + 1) * block * cpt * K * 8.0 >= 4.0e9) || (double)max_br * R * bt_size(K) * 8.0 >= 4.0e9)
        return fail(ctx, MCP_ERR_UNSUPPORTED, "tree too large for 32-bit scratch offsets (%d nodes)", max_br);
    ctx->scratch_per_cta = level_mode ? 4 : (long long)(ctx->n_slots + ctx->n_stack) * block * cpt * K;

    // topology upload: [TreeDev x T][ops][row_base]
    ctx->off_trees = 0;
    ctx->off_ops = (sizeof(TreeDev) * T + 31) & ~(size_t)31;
    ctx->off_rowbase = ctx->off_ops + (size_t)n_ops * 32;
    ctx->off_levels = ctx->off_rowbase + sizeof(int32_t) * ctx->grid;
    ctx->topo_bytes = ctx->off_levels + sizeof(int32_t) * (size_t)std::max<long long>(n_lvl_ints, 1);
    int e;
    if ((e = ensure_pin(ctx, ctx->h_topo, ctx->topo_bytes))) return e;
    if ((e = ensure_dev(ctx, ctx->d_topo, ctx->topo_bytes))) return e;
    char* h = (char*)ctx->h_topo.p;
    std::memcpy(h + ctx->off_trees, ctx->trees.data(), sizeof(TreeDev) * T);
    char* ho = h + ctx->off_ops;
    for (int t = 0; t < T; ++t) {
        const Schedule& sc = ctx->scheds[t];
        std::memcpy(ho, sc.post.data(), sc.post.size() * 32);
        ho += sc.post.size() * 32;
        std::memcpy(ho, sc.pre.data(), sc.pre.size() * 32);
        ho += sc.pre.size() * 32;
    }
    std::memcpy(h + ctx->off_rowbase, row_base.data(), sizeof(int32_t) * ctx->grid);
    {
        int32_t* hl = (int32_t*)(h + ctx->off_levels);
        for (int t = 0; t < T; ++t) {
            const Schedule& sc = ctx->scheds[t];
            for (int32_t v : sc.post_levels) *hl++ = v;
            for (int32_t v : sc.pre_levels) *hl++ = v;
        }
    }
    return 0;
}

int eval_impl(mcp_ctx* ctx, const BatchArgs& a, double* d_out_user, double* ll_out, double* const* grad_out) {
    if (!ctx) return fail(nullptr, MCP_ERR_ARG, "null context");
    if (a.T < 1) return fail(ctx, MCP_ERR_ARG, "batch must hold at least one tree");
    if (a.R < 1) return fail(ctx, MCP_ERR_ARG, "need at least one rate category");
    for (int t = 0; t < a.T; ++t) {
        if (!a.alns[t] || !a.po[t] || !a.pa[t] || !a.blv[t] || !a.U[t] || !a.D[t] || !a.Uinv[t] || !a.rates[t] || !a.pi[t])
            return fail(ctx, MCP_ERR_ARG, "tree %d: null argument", t);
        if (a.alns[t]->K != a.alns[0]->K) return fail(ctx, MCP_ERR_ARG, "all alignments of a batch must share K");
    }
    const int K = a.alns[0]->K, R = a.R, T = a.T;
    if (!k_supported(K)) return fail(ctx, MCP_ERR_UNSUPPORTED, "no kernel compiled for K = %d states", K);
    CUDA_TRY(ctx, cudaSetDevice(ctx->device));
    if (ctx->staged_pending) {  // the previous (asynchronous) evaluation may still be reading the staging buffers
        CUDA_TRY(ctx, cudaEventSynchronize(ctx->ev_staged));
        ctx->staged_pending = false;
    }
    bool rebuilt = false;
    int e = prepare_topology(ctx, a, K, &rebuilt);
    if (e) { ctx->sig.clear(); return e; }

    // buffers
    if ((e = ensure_pin(ctx, ctx->h_dyn, sizeof(double) * ctx->total_dyn))) return e;
    if ((e = ensure_dev(ctx, ctx->d_dyn, sizeof(double) * ctx->total_dyn))) return e;
    if ((e = ensure_dev(ctx, ctx->d_btab, sizeof(double) * ctx->total_btab))) return e;
    if ((e = ensure_dev(ctx, ctx->d_scratch, sizeof(double) * ctx->scratch_per_cta * ctx->grid))) return e;
    if ((e = ensure_dev(ctx, ctx->d_rows, sizeof(double) * ctx->row_stride * ctx->n_rows))) return e;
    if ((e = ensure_dev(ctx, ctx->d_rows_ll, sizeof(LLRow) * ctx->n_rows))) return e;
    double* d_out = d_out_user;
    if (!d_out) {
        if ((e = ensure_dev(ctx, ctx->d_out, sizeof(double) * ctx->total_out))) return e;
        if ((e = ensure_pin(ctx, ctx->h_out, sizeof(double) * ctx->total_out))) return e;
        d_out = (double*)ctx->d_out.p;
    }

    // per-evaluation parameters
    double* hd = (double*)ctx->h_dyn.p;
    for (int t = 0; t < T; ++t) {
        const int NN = a.NN[t];
        double* d = hd + ctx->trees[t].dyn_off;
        std::memcpy(d + dyn_blv(NN), a.blv[t], sizeof(double) * (NN - 1));
        std::memcpy(d + dyn_U(NN), a.U[t], sizeof(double) * K * K);
        std::memcpy(d + dyn_D(NN, K), a.D[t], sizeof(double) * K);
        std::memcpy(d + dyn_Uinv(NN, K), a.Uinv[t], sizeof(double) * K * K);
        d[dyn_mu(NN, K)] = a.mu[t];
        std::memcpy(d + dyn_rates(NN, K), a.rates[t], sizeof(double) * R);
        std::memcpy(d + dyn_pi(NN, K, R), a.pi[t], sizeof(double) * K);
        double* pr = d + dyn_prior(NN, K, R);
        pr[0] = 0.0;
        if (a.prior_kind != MCP_PRIOR_NONE) {
            // The prior is brought to the form  c0 - beta*T + sum_j w_j log t_j + k4 log T  on the host
            // (topology-only constants); sums over the branch lengths and the gradient are formed
            // on the device in the final reduction.
            if (!a.prior_params) return fail(ctx, MCP_ERR_ARG, "branch-length prior without parameters");
            double* w = pr + 4;
            if (a.prior_kind == MCP_PRIOR_EXPONENTIAL) {
                const double scale = a.prior_params[0];
                if (!(scale > 0.0)) return fail(ctx, MCP_ERR_ARG, "exponentialBL: scale must be positive");
                pr[1] = -(double)(NN - 1) * std::log(scale);
                pr[2] = 1.0 / scale;
                pr[3] = 0.0;
                for (int j = 0; j < NN - 1; ++j) w[j] = 0.0;
            } else if (a.prior_kind == MCP_PRIOR_COMPOUND_DIRICHLET) {
                const double alpha = a.prior_params[0], aa = a.prior_params[1], beta = a.prior_params[2], c = a.prior_params[3];
                if (!(alpha > 0.0 && aa > 0.0 && beta > 0.0 && c > 0.0))
                    return fail(ctx, MCP_ERR_ARG, "CompoundDirichlet: alpha, a, beta, c must be positive");
                // internal_external: 1 = the branch leads to an internal node, 0 = to a leaf (Prior.jl:15-23)
                std::vector<char> internal(NN, 0);
                for (int j = 0; j < NN; ++j) {
                    const int m = a.pa[t][j];
                    if (m >= 1 && m <= NN) internal[m - 1] = 1;
                }
                double nterm = 0.0;
                for (int j = 0; j < NN - 1; ++j) {
                    w[j] = internal[j] ? aa * c - 1.0 : aa - 1.0;
                    if (!internal[j]) nterm += 1.0;
                }
                const double n_int = nterm - 3.0;
                pr[1] = alpha * std::log(beta) - std::lgamma(alpha) - std::lgamma(aa) - std::lgamma(c) + std::lgamma(aa + c);
                pr[2] = beta;
                pr[3] = alpha - aa * nterm - aa * c * n_int;
            } else {
                return fail(ctx, MCP_ERR_ARG, "unknown branch-length prior kind %d", a.prior_kind);
            }
            pr[0] = 1.0;
        }
    }

    // substitution-model constants -> constant memory, one slot per distinct model of the batch
    if (R > MAX_RATES) return fail(ctx, MCP_ERR_UNSUPPORTED, "more than %d rate categories", MAX_RATES);
    const int model_doubles = k_templated(K) ? 2 * K * K + K + R * K : 0;
    if (model_doubles > MODEL_SLOT) return fail(ctx, MCP_ERR_UNSUPPORTED, "model with K=%d, R=%d does not fit a constant slot", K, R);
    if ((e = ensure_pin(ctx, ctx->h_model, sizeof(double) * MODEL_SLOT * MODEL_SLOTS))) return e;
    double* hm = (double*)ctx->h_model.p;
    int n_models = 0;
    for (int t = 0; t < T && k_templated(K); ++t) {
        double cand[MODEL_SLOT];
        std::memcpy(cand, a.U[t], sizeof(double) * K * K);
        std::memcpy(cand + K * K, a.Uinv[t], sizeof(double) * K * K);
        std::memcpy(cand + 2 * K * K, a.pi[t], sizeof(double) * K);
        for (int r = 0; r < R; ++r)
            for (int i = 0; i < K; ++i) cand[2 * K * K + K + r * K + i] = a.D[t][i] * a.rates[t][r] * a.mu[t];
        int slot = -1;
        for (int m = 0; m < n_models && slot < 0; ++m)
            if (std::memcmp(hm + (size_t)m * MODEL_SLOT, cand, sizeof(double) * model_doubles) == 0) slot = m;
        if (slot < 0) {
            if (n_models == MODEL_SLOTS)
                return fail(ctx, MCP_ERR_UNSUPPORTED, "more than %d distinct substitution models in one batch", MODEL_SLOTS);
            slot = n_models++;
            std::memcpy(hm + (size_t)slot * MODEL_SLOT, cand, sizeof(double) * model_doubles);
        }
        hd[ctx->trees[t].dyn_off + dyn_slot(a.NN[t], K, R)] = (double)slot;
    }
    const bool dyn_model = n_models > 1;

    cudaStream_t st = ctx->stream;
    mcp_stats& s = ctx->stats;
    s = mcp_stats{};
    CUDA_TRY(ctx, cudaEventRecord(ctx->ev[0], st));
    if (n_models > 1) {   // several models: slots in constant memory; one model travels as a kernel parameter
        CUDA_TRY(ctx, cudaMemcpyToSymbolAsync(c_model, hm, sizeof(double) * MODEL_SLOT * n_models, 0, cudaMemcpyHostToDevice, st));
        s.h2d_bytes += (int64_t)(sizeof(double) * MODEL_SLOT * n_models);
    }
    if (rebuilt) {
        CUDA_TRY(ctx, cudaMemcpyAsync(ctx->d_topo.p, ctx->h_topo.p, ctx->topo_bytes, cudaMemcpyHostToDevice, st));
        s.h2d_bytes += (int64_t)ctx->topo_bytes;
    }
    CUDA_TRY(ctx, cudaMemcpyAsync(ctx->d_dyn.p, hd, sizeof(double) * ctx->total_dyn, cudaMemcpyHostToDevice, st));
    s.h2d_bytes += (int64_t)(sizeof(double) * ctx->total_dyn);

    CUDA_TRY(ctx, cudaEventRecord(ctx->ev_staged, st));
    ctx->staged_pending = true;
    for (int t = 0; t < T; ++t)   // re-uploads of these alignments that are still in flight on the copy stream
        if (a.alns[t]->upload_pending) {
            CUDA_TRY(ctx, cudaStreamWaitEvent(st, a.alns[t]->ev_uploaded, 0));
            a.alns[t]->upload_pending = false;
        }

    const TreeDev* d_trees = (const TreeDev*)((char*)ctx->d_topo.p + ctx->off_trees);
    const bool fused = ctx->level_mode;   // small-tree kernel: tables, walk and final reduction in ONE launch
    if (fused && !ctx->d_counter.p) {
        if ((e = ensure_dev(ctx, ctx->d_counter, 256))) return e;
        CUDA_TRY(ctx, cudaMemsetAsync(ctx->d_counter.p, 0, 256, st));
    }
    if (!fused) {
        dim3 grid((ctx->max_br * R + 127) / 128, T);
        build_branch_tables<<<grid, 128, 0, st>>>(d_trees, (const double*)ctx->d_dyn.p, (double*)ctx->d_btab.p, K, R);
        CUDA_TRY(ctx, cudaGetLastError());
    }
    WalkParams wp;
    wp.trees = d_trees;
    wp.ops = (const int4*)((char*)ctx->d_topo.p + ctx->off_ops);
    wp.btab = (const double*)ctx->d_btab.p;
    wp.dyn = (const double*)ctx->d_dyn.p;
    wp.scratch = (double*)ctx->d_scratch.p;
    wp.scratch_per_cta = ctx->scratch_per_cta;
    wp.rows = (double*)ctx->d_rows.p;
    wp.rows_ll = (LLRow*)ctx->d_rows_ll.p;
    wp.cta_row_base = (const int*)((char*)ctx->d_topo.p + ctx->off_rowbase);
    wp.levels = (const int*)((char*)ctx->d_topo.p + ctx->off_levels);
    // pinned host memory is device-addressable (unified addressing): the fused kernel writes results there
    wp.out = d_out_user ? d_out_user : (double*)ctx->h_out.p;
    wp.done_counter = (unsigned int*)ctx->d_counter.p;
    wp.row_stride = ctx->row_stride;
    wp.n_slots = ctx->n_slots;
    wp.n_stack = ctx->n_stack;
    wp.n_tiles = ctx->n_tiles;
    wp.T = T;
    wp.R = R;
    wp.want_grad = a.want_grad ? 1 : 0;
    wp.max_br = ctx->max_br;
    wp.max_rows = ctx->max_rows;
    static_assert(sizeof(wp.model) / sizeof(double) >= 2 * 6 * 6 + 6 + MAX_RATES * 6, "model parameter block too small");
    std::memset(wp.model, 0, sizeof wp.model);
    if (n_models == 1) std::memcpy(wp.model, hm, sizeof(double) * model_doubles);
    CUDA_TRY(ctx, cudaEventRecord(ctx->ev[1], st));
    int rc = 0;
    if (ctx->level_mode) {
        MCP_DISPATCH_K(K, rc = launch_levels<KK>(ctx, wp, dyn_model));
    } else if (k_templated(K)) {
        MCP_DISPATCH_K(K, rc = launch_walk<KK>(ctx, wp, dyn_model));
    } else {
        felsenstein_walk_generic<<<ctx->grid, ctx->block, ctx->smem_bytes, st>>>(wp, K);
        CUDA_TRY(ctx, cudaGetLastError());
    }
    if (rc) return rc;
    CUDA_TRY(ctx, cudaEventRecord(ctx->ev[2], st));
    CUDA_TRY(ctx, cudaEventRecord(ctx->ev_walk_done, st));
    for (int t = 0; t < T; ++t) {
        a.alns[t]->read_since_upload = true;
        if (a.alns[t]->streamed) CUDA_TRY(ctx, cudaEventRecord(a.alns[t]->ev_read_done, st));
    }
    if (!fused) {
        int maxNN = 0;
        for (int t = 0; t < T; ++t) maxNN = std::max(maxNN, a.NN[t]);
        dim3 grid((maxNN + FIN_J - 1) / FIN_J, T);
        finalize_results<<<grid, dim3(FIN_J, FIN_G), 0, st>>>(d_trees, (const double*)ctx->d_rows.p, ctx->row_stride,
                                               (const LLRow*)ctx->d_rows_ll.p, d_out, wp.want_grad,
                                               (const double*)ctx->d_dyn.p, K, R);
        CUDA_TRY(ctx, cudaGetLastError());
    }
    s.kernel_launches = fused ? 1 : 3;
    s.grid = ctx->grid;
    s.block = ctx->block;
    s.tiles = ctx->n_tiles;
    s.schedule_rebuilt = rebuilt ? 1 : 0;
    s.scratch_bytes = (int64_t)ctx->d_scratch.cap;
    if (d_out_user) {
        CUDA_TRY(ctx, cudaEventRecord(ctx->ev[3], st));
        ctx->pending_async = true;
        return 0;
    }
    if (!fused) CUDA_TRY(ctx, cudaMemcpyAsync(ctx->h_out.p, d_out, sizeof(double) * ctx->total_out, cudaMemcpyDeviceToHost, st));
    s.d2h_bytes = (int64_t)(sizeof(double) * ctx->total_out);
    CUDA_TRY(ctx, cudaEventRecord(ctx->ev[3], st));
    CUDA_TRY(ctx, cudaStreamSynchronize(st));
    ctx->staged_pending = false;
    ctx->pending_async = false;
    const double* ho = (const double*)ctx->h_out.p;
    for (int t = 0; t < T; ++t) {
        const double* o = ho + ctx->trees[t].out_off;
        if (ll_out) ll_out[t] = o[0];
        if (a.want_grad && grad_out && grad_out[t]) std::memcpy(grad_out[t], o + 1, sizeof(double) * (a.NN[t] - 1));
    }
    return 0;
}

int make_alignment(mcp_ctx* ctx, const unsigned char* codes, int K, long long S, const int32_t* leaf_nums,
                   int n_leaves, mcp_alignment** out) {
    mcp_alignment* al = new mcp_alignment();
    al->K = K;
    al->S = S;
    al->stride = (S + 1023) & ~1023LL;   // a tile (<= 256 threads x 4 columns) never reads past a row
    if (al->stride == 0) al->stride = 1024;
    al->n_leaves = n_leaves;
    al->leaf_nums.assign(leaf_nums, leaf_nums + n_leaves);
    al->id = ctx->next_aln_id++;
    size_t bytes = (size_t)al->stride * (size_t)std::max(n_leaves, 1);
    cudaError_t e = cudaMalloc((void**)&al->d_codes, bytes);
    if (e != cudaSuccess) {
        delete al;
        return fail(ctx, MCP_ERR_CUDA, "cudaMalloc of %zu bytes for the alignment failed: %s", bytes, cudaGetErrorString(e));
    }
    // On the context's own stream and synchronised: a "synchronous" pageable host-to-device copy
    // on the default stream may return before the DMA has landed, and evaluations run on a
    // non-blocking stream that does not wait for the default stream.
    e = cudaMemsetAsync(al->d_codes, K, bytes, ctx->stream);
    if (e == cudaSuccess && S > 0 && n_leaves > 0)
        e = cudaMemcpy2DAsync(al->d_codes, (size_t)al->stride, codes, (size_t)S, (size_t)S, (size_t)n_leaves,
                              cudaMemcpyHostToDevice, ctx->stream);
    if (e == cudaSuccess) e = cudaStreamSynchronize(ctx->stream);
    if (e == cudaSuccess) e = cudaEventCreateWithFlags(&al->ev_uploaded, cudaEventDisableTiming);
    if (e == cudaSuccess) e = cudaEventCreateWithFlags(&al->ev_read_done, cudaEventDisableTiming);
    if (e != cudaSuccess) {
        if (al->ev_uploaded) cudaEventDestroy(al->ev_uploaded);
        if (al->ev_read_done) cudaEventDestroy(al->ev_read_done);
        cudaFree(al->d_codes);
        delete al;
        return fail(ctx, MCP_ERR_CUDA, "alignment upload failed: %s", cudaGetErrorString(e));
    }
    *out = al;
    return 0;
}

}  // namespace

// --------------------------------------------------------------------------------------------
// C ABI
// --------------------------------------------------------------------------------------------
extern "C" {

int mcp_abi_version(void) { return MCP_ABI_VERSION; }

const char* mcp_last_error(const mcp_ctx* ctx) { return ctx ? ctx->error.c_str() : g_error.c_str(); }

int mcp_create(mcp_ctx** out, int device) {
    if (!out) return fail(nullptr, MCP_ERR_ARG, "mcp_create: null output pointer");
    *out = nullptr;
    int n = 0;
    cudaError_t e = cudaGetDeviceCount(&n);
    if (e != cudaSuccess || n == 0) {
        cudaGetLastError();
        return fail(nullptr, MCP_ERR_CUDA, "no CUDA device available (%s); this library has no CPU fallback",
                    e != cudaSuccess ? cudaGetErrorString(e) : "device count is 0");
    }
    if (device < 0 || device >= n) return fail(nullptr, MCP_ERR_ARG, "device %d out of range (have %d)", device, n);
    e = cudaSetDevice(device);
    if (e != cudaSuccess) return fail(nullptr, MCP_ERR_CUDA, "cudaSetDevice(%d): %s", device, cudaGetErrorString(e));
    cudaDeviceProp prop;
    e = cudaGetDeviceProperties(&prop, device);
    if (e != cudaSuccess) return fail(nullptr, MCP_ERR_CUDA, "cudaGetDeviceProperties: %s", cudaGetErrorString(e));
    if (prop.major < 10)
        return fail(nullptr, MCP_ERR_UNSUPPORTED, "device %d is sm_%d%d; this library is built for sm_100a (B200) only",
                    device, prop.major, prop.minor);
    mcp_ctx* ctx = new mcp_ctx();
    ctx->device = device;
    ctx->sm_count = prop.multiProcessorCount;
    e = cudaStreamCreateWithFlags(&ctx->own_stream, cudaStreamNonBlocking);
    for (int i = 0; e == cudaSuccess && i < 4; ++i) e = cudaEventCreate(&ctx->ev[i]);
    if (e == cudaSuccess) e = cudaEventCreateWithFlags(&ctx->ev_staged, cudaEventDisableTiming);
    if (e == cudaSuccess) e = cudaEventCreateWithFlags(&ctx->ev_walk_done, cudaEventDisableTiming);
    if (e == cudaSuccess) e = cudaStreamCreateWithFlags(&ctx->copy_stream, cudaStreamNonBlocking);
    if (e != cudaSuccess) {
        std::string msg = cudaGetErrorString(e);
        delete ctx;
        return fail(nullptr, MCP_ERR_CUDA, "stream/event creation failed: %s", msg.c_str());
    }
    ctx->stream = ctx->own_stream;
    *out = ctx;
    return 0;
}

int mcp_destroy(mcp_ctx* ctx) {
    if (!ctx) return 0;
    cudaSetDevice(ctx->device);
    cudaStreamSynchronize(ctx->stream);
    if (ctx->copy_stream) cudaStreamSynchronize(ctx->copy_stream);
    for (DevBuf* b : {&ctx->d_topo, &ctx->d_dyn, &ctx->d_btab, &ctx->d_scratch, &ctx->d_rows, &ctx->d_rows_ll, &ctx->d_out, &ctx->d_counter})
        if (b->p) cudaFree(b->p);
    for (PinBuf* b : {&ctx->h_topo, &ctx->h_dyn, &ctx->h_out, &ctx->h_model})
        if (b->p) cudaFreeHost(b->p);
    for (auto& ev : ctx->ev)
        if (ev) cudaEventDestroy(ev);
    if (ctx->ev_staged) cudaEventDestroy(ctx->ev_staged);
    if (ctx->ev_walk_done) cudaEventDestroy(ctx->ev_walk_done);
    if (ctx->copy_stream) cudaStreamDestroy(ctx->copy_stream);
    if (ctx->own_stream) cudaStreamDestroy(ctx->own_stream);
    delete ctx;
    return 0;
}

int mcp_set_stream(mcp_ctx* ctx, void* cuda_stream) {
    if (!ctx) return fail(nullptr, MCP_ERR_ARG, "null context");
    if (ctx->pending_async) {
        CUDA_TRY(ctx, cudaStreamSynchronize(ctx->stream));
        ctx->pending_async = false;
    }
    ctx->stream = (cudaStream_t)cuda_stream;   // NULL is the CUDA default stream, a valid choice
    return 0;
}

int mcp_synchronize(mcp_ctx* ctx) {
    if (!ctx) return fail(nullptr, MCP_ERR_ARG, "null context");
    CUDA_TRY(ctx, cudaSetDevice(ctx->device));
    CUDA_TRY(ctx, cudaStreamSynchronize(ctx->stream));
    CUDA_TRY(ctx, cudaStreamSynchronize(ctx->copy_stream));
    ctx->pending_async = false;
    return 0;
}

int mcp_use_own_stream(mcp_ctx* ctx) {
    if (!ctx) return fail(nullptr, MCP_ERR_ARG, "null context");
    if (ctx->pending_async) {
        CUDA_TRY(ctx, cudaStreamSynchronize(ctx->stream));
        ctx->pending_async = false;
    }
    ctx->stream = ctx->own_stream;
    return 0;
}

int mcp_set_launch(mcp_ctx* ctx, int block, int ctas_per_sm) {
    if (!ctx) return fail(nullptr, MCP_ERR_ARG, "null context");
    if (block != 0 && (block < 32 || block > 256 || (block & 31)))
        return fail(ctx, MCP_ERR_ARG, "block must be 0 or a multiple of 32 in [32, 256]");
    if (ctas_per_sm < 0) return fail(ctx, MCP_ERR_ARG, "ctas_per_sm must be >= 0");
    ctx->opt_block = block;
    ctx->opt_ctas_per_sm = ctas_per_sm;
    ctx->sig.clear();
    return 0;
}

int mcp_set_scratch_mode(mcp_ctx* ctx, int mode) {
    if (!ctx) return fail(nullptr, MCP_ERR_ARG, "null context");
    if (mode < -1 || mode > 1) return fail(ctx, MCP_ERR_ARG, "scratch mode must be -1 (automatic), 0 (HBM) or 1 (shared memory when it fits)");
    ctx->opt_smem_scratch = mode;
    ctx->sig.clear();
    return 0;
}

int mcp_set_level_mode(mcp_ctx* ctx, int mode) {
    if (!ctx) return fail(nullptr, MCP_ERR_ARG, "null context");
    if (mode < -1 || mode > 1) return fail(ctx, MCP_ERR_ARG, "level mode must be -1 (automatic), 0 (off) or 1 (whenever the tree fits)");
    ctx->opt_levels = mode;
    ctx->sig.clear();
    return 0;
}

int mcp_set_columns_per_thread(mcp_ctx* ctx, int cpt) {
    if (!ctx) return fail(nullptr, MCP_ERR_ARG, "null context");
    if (cpt != 0 && cpt != 1 && cpt != 2) return fail(ctx, MCP_ERR_ARG, "columns per thread must be 0 (automatic), 1 or 2");
    ctx->opt_cpt = cpt;
    ctx->sig.clear();
    return 0;
}

int mcp_alignment_from_codes(mcp_ctx* ctx, const uint8_t* codes, int K, int64_t S, const int32_t* leaf_nums,
                             int n_leaves, mcp_alignment** out) {
    if (!ctx) return fail(nullptr, MCP_ERR_ARG, "null context");
    if (!out || !codes || !leaf_nums) return fail(ctx, MCP_ERR_ARG, "mcp_alignment_from_codes: null argument");
    if (K < 1 || K > 254 || S < 0 || n_leaves < 1) return fail(ctx, MCP_ERR_ARG, "mcp_alignment_from_codes: bad K/S/n_leaves");
    CUDA_TRY(ctx, cudaSetDevice(ctx->device));
    return make_alignment(ctx, codes, K, S, leaf_nums, n_leaves, out);
}

int mcp_alignment_from_dense(mcp_ctx* ctx, const double* x, int K, int64_t S, int NN, const int32_t* leaf_nums,
                             int n_leaves, mcp_alignment** out) {
    if (!ctx) return fail(nullptr, MCP_ERR_ARG, "null context");
    if (!out || !x || !leaf_nums) return fail(ctx, MCP_ERR_ARG, "mcp_alignment_from_dense: null argument");
    if (K < 1 || K > 254 || S < 0 || n_leaves < 1 || NN < n_leaves) return fail(ctx, MCP_ERR_ARG, "mcp_alignment_from_dense: bad sizes");
    std::vector<unsigned char> codes((size_t)n_leaves * (size_t)S);
    for (int l = 0; l < n_leaves; ++l) {
        const int num = leaf_nums[l];
        if (num < 1 || num > NN) return fail(ctx, MCP_ERR_ARG, "leaf number %d out of range", num);
        const double* slab = x + (size_t)K * (size_t)S * (size_t)(num - 1);
        for (int64_t s = 0; s < S; ++s) {
            const double* col = slab + (size_t)K * s;
            int ones = 0, zeros = 0, first = -1;
            for (int k = 0; k < K; ++k) {
                if (col[k] == 1.0) { ++ones; if (first < 0) first = k; }
                else if (col[k] == 0.0) ++zeros;
            }
            unsigned char c;
            if (ones == K) c = (unsigned char)K;
            else if (ones == 1 && zeros == K - 1) c = (unsigned char)first;
            else
                return fail(ctx, MCP_ERR_DATA, "leaf %d, site %lld: column is neither one-hot nor all ones", num, (long long)s + 1);
            codes[(size_t)l * S + s] = c;
        }
    }
    CUDA_TRY(ctx, cudaSetDevice(ctx->device));
    return make_alignment(ctx, codes.data(), K, S, leaf_nums, n_leaves, out);
}

int mcp_alignment_update_codes(mcp_ctx* ctx, mcp_alignment* aln, const uint8_t* codes) {
    if (!ctx) return fail(nullptr, MCP_ERR_ARG, "null context");
    if (!aln || !codes) return fail(ctx, MCP_ERR_ARG, "mcp_alignment_update_codes: null argument");
    CUDA_TRY(ctx, cudaSetDevice(ctx->device));
    // On the copy stream: the transfer overlaps whatever the evaluation stream is doing (the
    // evaluation of another site block, typically).  It must not overtake an evaluation that still
    // reads this buffer, and the next evaluation of this alignment waits for it (eval_impl).
    if (aln->read_since_upload) {
        // first re-upload: only the context-wide "last walk finished" event exists (conservative);
        // afterwards every evaluation of this alignment records the alignment's own event
        CUDA_TRY(ctx, cudaStreamWaitEvent(ctx->copy_stream, aln->streamed ? aln->ev_read_done : ctx->ev_walk_done, 0));
        aln->read_since_upload = false;
    }
    aln->streamed = true;
    if (aln->S > 0)
        CUDA_TRY(ctx, cudaMemcpy2DAsync(aln->d_codes, (size_t)aln->stride, codes, (size_t)aln->S, (size_t)aln->S,
                                        (size_t)aln->n_leaves, cudaMemcpyHostToDevice, ctx->copy_stream));
    CUDA_TRY(ctx, cudaEventRecord(aln->ev_uploaded, ctx->copy_stream));
    aln->upload_pending = true;
    ctx->pending_async = true;   // only consulted before buffers are freed / the stream is changed
    return 0;
}

int mcp_alignment_destroy(mcp_ctx* ctx, mcp_alignment* aln) {
    if (!aln) return 0;
    if (ctx) {
        cudaSetDevice(ctx->device);
        cudaStreamSynchronize(ctx->stream);
        if (ctx->copy_stream) cudaStreamSynchronize(ctx->copy_stream);
        ctx->pending_async = false;
        ctx->sig.clear();
    }
    if (aln->ev_uploaded) cudaEventDestroy(aln->ev_uploaded);
    if (aln->ev_read_done) cudaEventDestroy(aln->ev_read_done);
    if (aln->d_codes) cudaFree(aln->d_codes);
    delete aln;
    return 0;
}

int mcp_eval(mcp_ctx* ctx, const mcp_alignment* aln, int NN, const int32_t* postorder_num, const int32_t* parent_num,
             const double* blv, const double* U, const double* D, const double* Uinv, double mu, const double* rates,
             int R, const double* pi, int want_grad, double* ll_out, double* grad_out) {
    BatchArgs a{1, &aln, &NN, &postorder_num, &parent_num, &blv, &U, &D, &Uinv, &mu, &rates, R, &pi, want_grad};
    double* g = grad_out;
    return eval_impl(ctx, a, nullptr, ll_out, &g);
}

int mcp_eval_posterior(mcp_ctx* ctx, const mcp_alignment* aln, int NN, const int32_t* postorder_num,
                       const int32_t* parent_num, const double* blv, const double* U, const double* D, const double* Uinv,
                       double mu, const double* rates, int R, const double* pi, int prior_kind, const double* prior_params,
                       double* lp_out, double* grad_out) {
    BatchArgs a{1, &aln, &NN, &postorder_num, &parent_num, &blv, &U, &D, &Uinv, &mu, &rates, R, &pi, grad_out ? 1 : 0};
    a.prior_kind = prior_kind;
    a.prior_params = prior_params;
    double* g = grad_out;
    return eval_impl(ctx, a, nullptr, lp_out, &g);
}

int mcp_eval_device(mcp_ctx* ctx, const mcp_alignment* aln, int NN, const int32_t* postorder_num,
                    const int32_t* parent_num, const double* blv, const double* U, const double* D, const double* Uinv,
                    double mu, const double* rates, int R, const double* pi, int want_grad, double* d_out) {
    if (!d_out) return fail(ctx, MCP_ERR_ARG, "mcp_eval_device: null device output pointer");
    BatchArgs a{1, &aln, &NN, &postorder_num, &parent_num, &blv, &U, &D, &Uinv, &mu, &rates, R, &pi, want_grad};
    return eval_impl(ctx, a, d_out, nullptr, nullptr);
}

int mcp_eval_batch(mcp_ctx* ctx, int T, const mcp_alignment* const* alns, const int32_t* NN,
                   const int32_t* const* postorder_num, const int32_t* const* parent_num, const double* const* blv,
                   const double* const* U, const double* const* D, const double* const* Uinv, const double* mu,
                   const double* const* rates, int R, const double* const* pi, int want_grad, double* ll_out,
                   double* const* grad_out) {
    if (!alns || !NN || !postorder_num || !parent_num || !blv || !U || !D || !Uinv || !mu || !rates || !pi)
        return fail(ctx, MCP_ERR_ARG, "mcp_eval_batch: null argument array");
    BatchArgs a{T, alns, NN, postorder_num, parent_num, blv, U, D, Uinv, mu, rates, R, pi, want_grad};
    return eval_impl(ctx, a, nullptr, ll_out, grad_out);
}

int mcp_wave_columns(mcp_ctx* ctx, int K, int n_nodes, int want_grad, int64_t* columns) {
    if (!ctx || !columns) return fail(ctx, MCP_ERR_ARG, "mcp_wave_columns: null argument");
    if (!k_supported(K)) return fail(ctx, MCP_ERR_UNSUPPORTED, "no kernel compiled for K = %d states", K);
    CUDA_TRY(ctx, cudaSetDevice(ctx->device));
    int block = ctx->opt_block > 0 ? ctx->opt_block : 256;
    int cpt = ctx->opt_cpt > 0 ? ctx->opt_cpt : (K <= 3 ? 2 : 1);
    int occ = 0, rc = 0;
    if (k_templated(K)) {
        const size_t smem = walk_smem_bytes(K, std::max(n_nodes, 1), want_grad ? 1 : 0, block, cpt);
        MCP_DISPATCH_K(K, rc = occupancy_for<KK>(ctx, block, cpt, smem, false, &occ));
        if (rc) return rc;
    } else {
        cpt = 1;
        block = std::min(block, 128);
        const size_t smem = want_grad ? (size_t)std::max(n_nodes, 1) * sizeof(double) : 0;
        if ((rc = ensure_smem_attr(ctx, felsenstein_walk_generic, smem))) return rc;
        CUDA_TRY(ctx, cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, felsenstein_walk_generic, block, smem));
    }
    if (occ < 1) return fail(ctx, MCP_ERR_CUDA, "walk kernel does not fit on an SM (block %d)", block);
    if (ctx->opt_ctas_per_sm > 0) occ = std::min(occ, ctx->opt_ctas_per_sm);
    *columns = (int64_t)occ * ctx->sm_count * block * cpt;
    return 0;
}

int mcp_get_stats(const mcp_ctx* ctx, mcp_stats* out) {
    if (!ctx || !out) return fail(nullptr, MCP_ERR_ARG, "mcp_get_stats: null argument");
    *out = ctx->stats;
    // event times are read lazily: they exist once the stream has passed the last event
    float ms = 0.f;
    if (ctx->ev[3] && cudaEventQuery(ctx->ev[3]) == cudaSuccess) {
        if (cudaEventElapsedTime(&ms, ctx->ev[1], ctx->ev[2]) == cudaSuccess) out->walk_ms = ms;
        if (cudaEventElapsedTime(&ms, ctx->ev[0], ctx->ev[3]) == cudaSuccess) out->device_ms = ms;
    } else {
        cudaGetLastError();
    }
    return 0;
}

int mcp_schedule_dump(int NN, const int32_t* postorder_num, const int32_t* parent_num, const int32_t* leaf_row,
                      int want_grad, int32_t* post_ops, int cap_post, int32_t* pre_ops, int cap_pre, int32_t* info) {
    if (!postorder_num || !parent_num || !leaf_row || !info) return fail(nullptr, MCP_ERR_ARG, "mcp_schedule_dump: null argument");
    Schedule sc;
    const bool by_levels = (want_grad & 2) != 0;   // bit 1 of want_grad selects the level-ordered program
    want_grad &= 1;
    std::string err = mcp::build_schedule(NN, postorder_num, parent_num, leaf_row, want_grad != 0, sc, by_levels);
    if (!err.empty()) return fail(nullptr, MCP_ERR_ARG, "%s", err.c_str());
    info[0] = (int32_t)sc.post.size();
    info[1] = (int32_t)sc.pre.size();
    info[2] = sc.n_slots;
    info[3] = sc.n_stack;
    info[4] = sc.n_dnodes;
    info[5] = (int32_t)(sc.post_levels.empty() ? 0 : sc.post_levels.size() - 1);
    info[6] = (int32_t)(sc.pre_levels.empty() ? 0 : sc.pre_levels.size() - 1);
    if ((int)sc.post.size() > cap_post || (int)sc.pre.size() > cap_pre)
        return fail(nullptr, MCP_ERR_ARG, "mcp_schedule_dump: output arrays too small");
    if (post_ops) std::memcpy(post_ops, sc.post.data(), sc.post.size() * 32);
    if (pre_ops) std::memcpy(pre_ops, sc.pre.data(), sc.pre.size() * 32);
    return 0;
}

}  // extern "C"
