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
// fp64 throughout.  Transitions are applied in eigen-space, P L = L + U (expm1(.) * (Uinv L)), with
// U / Uinv as constant-bank operands and the null eigenvalue of the rate matrix skipped.  Rescaling
// uses exact powers of two (exponent extraction) instead of the reference's divide-by-max + log per
// node; the integer exponent sum is exact and logL = ln2 * sum(exponents) + sum(log(pi . L_root)).
// Branch-gradient sums are accumulated without atomics, in a fixed order (reproducible bit for bit).
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
#include <type_traits>
#include <vector>

#include "device_layout.cuh"
#include "device_math.cuh"
#include "kernel_tables.cuh"
#include "kernel_walk.cuh"
#include "epilogue_prior.cuh"
#include "kernel_levels.cuh"
#include "kernel_generic.cuh"
#include "kernel_finalize.cuh"

namespace {

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
    int opt_acc_mode = -1;       // gradient accumulator of the walk: -1 automatic, 0 shared memory, 1 global memory (RED)
    bool acc_global = false;     // what the current plan uses
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
template <int K, int CPT, bool DYN, bool SSCR, int NE, bool ACCG>
int launch_walk_inst(mcp_ctx* ctx, const WalkParams& wp) {
    int e = ensure_smem_attr(ctx, felsenstein_walk<K, CPT, DYN, SSCR, NE, ACCG>, ctx->smem_bytes);
    if (e) return e;
    felsenstein_walk<K, CPT, DYN, SSCR, NE, ACCG><<<ctx->grid, ctx->block, ctx->smem_bytes, ctx->stream>>>(wp);
    CUDA_TRY(ctx, cudaGetLastError());
    return 0;
}
// null_last: every model of the batch has a null eigenvalue, moved to the last position by the host
// (true for any rate matrix) -> kernels with K - 1 active eigen-components.  Otherwise (a caller
// passing some other decomposition) the full-K kernels.  Those, and the kernels that accumulate the
// gradient in global memory (ctx->acc_global: very large trees), exist in the constant-memory (DYN)
// flavour only -- and the latter with one column per thread only, which prepare_topology arranges.
template <int K>
int launch_walk(mcp_ctx* ctx, const WalkParams& wp, bool dyn_model, bool null_last) {
    constexpr int NE = K - 1;
    if (ctx->acc_global)
        return null_last ? launch_walk_inst<K, 1, true, false, NE, true>(ctx, wp) : launch_walk_inst<K, 1, true, false, K, true>(ctx, wp);
    if (!null_last)
        return ctx->cpt == 2 ? launch_walk_inst<K, 2, true, false, K, false>(ctx, wp) : launch_walk_inst<K, 1, true, false, K, false>(ctx, wp);
    if (ctx->smem_scratch && !dyn_model && ctx->cpt == 1) return launch_walk_inst<K, 1, false, true, NE, false>(ctx, wp);
    if (ctx->cpt == 2) return dyn_model ? launch_walk_inst<K, 2, true, false, NE, false>(ctx, wp) : launch_walk_inst<K, 2, false, false, NE, false>(ctx, wp);
    return dyn_model ? launch_walk_inst<K, 1, true, false, NE, false>(ctx, wp) : launch_walk_inst<K, 1, false, false, NE, false>(ctx, wp);
}
template <int K, int CPT>
int occupancy_inst(mcp_ctx* ctx, int block, size_t smem, bool sscr, bool acc_global, int* out) {
    int e;
    constexpr int NE = K - 1;
    if (acc_global) {
        if ((e = ensure_smem_attr(ctx, felsenstein_walk<K, 1, true, false, K, true>, smem))) return e;
        if ((e = ensure_smem_attr(ctx, felsenstein_walk<K, 1, true, false, NE, true>, smem))) return e;
        int o1 = 0, o2 = 0;
        CUDA_TRY(ctx, cudaOccupancyMaxActiveBlocksPerMultiprocessor(&o1, felsenstein_walk<K, 1, true, false, K, true>, block, smem));
        CUDA_TRY(ctx, cudaOccupancyMaxActiveBlocksPerMultiprocessor(&o2, felsenstein_walk<K, 1, true, false, NE, true>, block, smem));
        *out = std::min(o1, o2);
        return 0;
    }
    if (sscr) {
        if ((e = ensure_smem_attr(ctx, felsenstein_walk<K, 1, false, true, NE, false>, smem))) return e;
        CUDA_TRY(ctx, cudaOccupancyMaxActiveBlocksPerMultiprocessor(out, felsenstein_walk<K, 1, false, true, NE, false>, block, smem));
        return 0;
    }
    // the variants differ by a few registers: size the persistent grid for the most demanding one
    if ((e = ensure_smem_attr(ctx, felsenstein_walk<K, CPT, true, false, NE, false>, smem))) return e;
    if ((e = ensure_smem_attr(ctx, felsenstein_walk<K, CPT, false, false, NE, false>, smem))) return e;
    if ((e = ensure_smem_attr(ctx, felsenstein_walk<K, CPT, true, false, K, false>, smem))) return e;
    int o1 = 0, o2 = 0, o3 = 0;
    CUDA_TRY(ctx, cudaOccupancyMaxActiveBlocksPerMultiprocessor(&o1, felsenstein_walk<K, CPT, true, false, NE, false>, block, smem));
    CUDA_TRY(ctx, cudaOccupancyMaxActiveBlocksPerMultiprocessor(&o2, felsenstein_walk<K, CPT, false, false, NE, false>, block, smem));
    CUDA_TRY(ctx, cudaOccupancyMaxActiveBlocksPerMultiprocessor(&o3, felsenstein_walk<K, CPT, true, false, K, false>, block, smem));
    *out = std::min(o1, std::min(o2, o3));
    return 0;
}
template <int K>
int occupancy_for(mcp_ctx* ctx, int block, int cpt, size_t smem, bool sscr, bool acc_global, int* out) {
    return cpt == 2 ? occupancy_inst<K, 2>(ctx, block, smem, false, acc_global, out) : occupancy_inst<K, 1>(ctx, block, smem, sscr, acc_global, out);
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
size_t walk_smem_bytes(int K, int max_br, int shared_acc, int block, int cpt) {
    const size_t acc = shared_acc ? (((size_t)max_br * 8 + 15) & ~(size_t)15) + WALK_PART_BYTES : 0;
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

// Copies the eigen-decomposition (U, Uinv column-major K x K, D) with the eigenvalue of smallest
// magnitude moved to the LAST position (columns of U, rows of Uinv, entries of D permuted alike:
// U diag(f(D)) Uinv is unchanged).  Returns true when that eigenvalue is null, |D_i| <= 8 eps max|D| --
// every rate matrix has one (rows sum to zero; LAPACK returns it as ~1e-17) -- so that its terms
// expm1(mu t D_i r) and D_i mu r exp(.) vanish to rounding and the walk kernel may skip them.
bool null_eigenvalue_last(const double* U, const double* D, const double* Uinv, int K, double* Uo, double* Do, double* Uinvo) {
    int i0 = 0;
    double dmax = 0.0;
    for (int i = 0; i < K; ++i) {
        if (std::fabs(D[i]) < std::fabs(D[i0])) i0 = i;
        dmax = std::max(dmax, std::fabs(D[i]));
    }
    for (int i = 0; i < K; ++i) {
        const int from = i == K - 1 ? i0 : (i < i0 ? i : i + 1);
        Do[i] = D[from];
        for (int s = 0; s < K; ++s) Uo[s + K * i] = U[s + K * from];
        for (int j = 0; j < K; ++j) Uinvo[i + K * j] = Uinv[from + K * j];
    }
    return std::fabs(D[i0]) <= 8.0 * 2.220446049250313e-16 * dmax;
}
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
    // reduction) once there is enough work to fill the GPU.  Measured: 1.28x at K = 2, 1.06x at K = 4
    // (cfg4: 78.2 -> 73.6 ms; 2 CTAs x 256 threads per SM at 128 registers), profiles/r1_walk_notes.md.
    int cpt = ctx->opt_cpt;
    if (cpt <= 0) cpt = (K <= 4 && total_cols / (2LL * block) >= 12LL * ctx->sm_count) ? 2 : 1;
    // Very large trees: the per-branch accumulator no longer fits in shared memory next to the staging
    // buffers; those kernels exist with one column per thread only (launch_walk).
    int max_nn = 0;
    for (int t = 0; t < T; ++t) max_nn = std::max(max_nn, (int)a.NN[t]);
    const bool acc_global = a.want_grad && k_templated(K) && walk_acc_global(max_nn, ctx->opt_acc_mode);
    if (acc_global) cpt = 1;
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
    ctx->acc_global = acc_global && !level_mode;
    ctx->smem_bytes = level_mode ? LevelSmem::total(max_br, a.want_grad ? 1 : 0, max_rows, n_slots, a.want_grad ? n_stack : 0, K)
                      : k_templated(K) ? walk_smem_bytes(K, max_br, a.want_grad && !acc_global ? 1 : 0, block, cpt)
                                       : (a.want_grad ? (size_t)max_br * sizeof(double) : 0);
    // Small problems: keep the partials scratch in shared memory (latency path).  Automatic when the
    // whole input is a handful of tiles per SM and the scratch of one CTA fits next to the staging
    // buffers.
    {
        const size_t scr_bytes = (size_t)(ctx->n_slots + ctx->n_stack) * block * cpt * K * 8;
        const bool fits = !level_mode && !acc_global && k_templated(K) && cpt == 1 && ctx->smem_bytes + scr_bytes <= 96 * 1024;
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
        MCP_DISPATCH_K(K, rc = occupancy_for<KK>(ctx, block, cpt, ctx->smem_bytes, ctx->smem_scratch, ctx->acc_global, &occ));
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
    bool all_null_last = true;
    for (int t = 0; t < T; ++t) {
        const int NN = a.NN[t];
        double* d = hd + ctx->trees[t].dyn_off;
        std::memcpy(d + dyn_blv(NN), a.blv[t], sizeof(double) * (NN - 1));
        // eigen-decomposition with a null eigenvalue (if any) moved to the last position
        all_null_last = null_eigenvalue_last(a.U[t], a.D[t], a.Uinv[t], K, d + dyn_U(NN), d + dyn_D(NN, K), d + dyn_Uinv(NN, K)) && all_null_last;
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
        const double* dt = hd + ctx->trees[t].dyn_off;       // the permuted decomposition stored above
        std::memcpy(cand, dt + dyn_U(a.NN[t]), sizeof(double) * K * K);
        std::memcpy(cand + K * K, dt + dyn_Uinv(a.NN[t], K), sizeof(double) * K * K);
        std::memcpy(cand + 2 * K * K, a.pi[t], sizeof(double) * K);
        for (int r = 0; r < R; ++r)
            for (int i = 0; i < K; ++i) cand[2 * K * K + K + r * K + i] = dt[dyn_D(a.NN[t], K) + i] * a.rates[t][r] * a.mu[t];
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
    // the walk kernels with all K eigen-components read their model from constant memory only
    const bool dyn_model = n_models > 1 || (k_templated(K) && (!all_null_last || ctx->acc_global));

    cudaStream_t st = ctx->stream;
    mcp_stats& s = ctx->stats;
    s = mcp_stats{};
    CUDA_TRY(ctx, cudaEventRecord(ctx->ev[0], st));
    if (dyn_model) {   // several models: slots in constant memory; one model travels as a kernel parameter
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
        MCP_DISPATCH_K(K, rc = launch_walk<KK>(ctx, wp, dyn_model, all_null_last));
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

int mcp_set_accumulator_mode(mcp_ctx* ctx, int mode) {
    if (!ctx) return fail(nullptr, MCP_ERR_ARG, "null context");
    if (mode < -1 || mode > 1) return fail(ctx, MCP_ERR_ARG, "accumulator mode must be -1 (automatic), 0 (shared memory) or 1 (global memory)");
    ctx->opt_acc_mode = mode;
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
    int cpt = ctx->opt_cpt > 0 ? ctx->opt_cpt : (K <= 4 ? 2 : 1);   // what prepare_topology picks for large inputs
    int occ = 0, rc = 0;
    if (k_templated(K)) {
        const bool acc_global = want_grad && walk_acc_global(n_nodes, ctx->opt_acc_mode);
        if (acc_global) cpt = 1;
        const size_t smem = walk_smem_bytes(K, std::max(n_nodes, 1), want_grad && !acc_global ? 1 : 0, block, cpt);
        MCP_DISPATCH_K(K, rc = occupancy_for<KK>(ctx, block, cpt, smem, false, acc_global, &occ));
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

int mcp_model_reorder(int K, const double* U, const double* D, const double* Uinv, double* U_out, double* D_out,
                      double* Uinv_out, int* null_last) {
    if (!U || !D || !Uinv || !U_out || !D_out || !Uinv_out || !null_last)
        return fail(nullptr, MCP_ERR_ARG, "mcp_model_reorder: null argument");
    if (!k_supported(K)) return fail(nullptr, MCP_ERR_UNSUPPORTED, "no kernel compiled for K = %d states", K);
    *null_last = null_eigenvalue_last(U, D, Uinv, K, U_out, D_out, Uinv_out) ? 1 : 0;
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
