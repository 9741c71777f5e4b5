// mcphylo_b200.cu — libmcphylo_b200.so: C ABI (include/mcphylo_b200.h), host side.
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
//
// Source map: the walk kernels are compiled per state count in walk_k2.cu .. walk_k6.cu and
// walk_generic.cu and reached through kernel_api.hpp; this unit holds the planner (planner.hpp), the
// evaluation driver, the multi-GPU group (site shards + one all-reduce of [logL, gradient]) and the
// small kernels around the walk (branch tables, final reductions).
#include "planner.hpp"

#include "kernel_tables.cuh"
#include "epilogue_prior.cuh"
#include "kernel_finalize.cuh"

namespace {

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

// Several contexts of one process may share a GPU (chains in threads, a streaming context next to a
// resident one).  The constant-memory model slots of a kernel translation unit are per device, so a
// context that is about to overwrite them waits (on its stream) for the last kernel of any other context
// that read them; the bookkeeping is guarded by a mutex, the waiting happens on the device.
struct ModelSlotGuard {
    std::mutex mu;
    struct Entry { int device; const void* unit; cudaEvent_t ev; mcp_ctx* last; };
    std::vector<Entry> entries;
    Entry* find(int device, const void* unit) {
        for (auto& e : entries)
            if (e.device == device && e.unit == unit) return &e;
        return nullptr;
    }
};
ModelSlotGuard g_slots;

// Branch-length prior -> the form  c0 - beta*T + sum_j w_j log t_j + k4 log T  (topology-only constants);
// sums over the branch lengths and the gradient are formed on the device in the final reduction.
// pr = [enabled, c0, beta, k4, w_1 .. w_{NN-1}].  Returns an error message or null.
const char* fill_prior(int kind, const double* params, int NN, const int32_t* pa, double* pr) {
    pr[0] = 0.0;
    if (kind == MCP_PRIOR_NONE) return nullptr;
    if (!params) return "branch-length prior without parameters";
    double* w = pr + 4;
    if (kind == MCP_PRIOR_EXPONENTIAL) {
        const double scale = params[0];
        if (!(scale > 0.0)) return "exponentialBL: scale must be positive";
        pr[1] = -(double)(NN - 1) * std::log(scale);
        pr[2] = 1.0 / scale;
        pr[3] = 0.0;
        for (int j = 0; j < NN - 1; ++j) w[j] = 0.0;
    } else if (kind == MCP_PRIOR_COMPOUND_DIRICHLET) {
        const double alpha = params[0], aa = params[1], beta = params[2], c = params[3];
        if (!(alpha > 0.0 && aa > 0.0 && beta > 0.0 && c > 0.0)) return "CompoundDirichlet: alpha, a, beta, c must be positive";
        // internal_external: 1 = the branch leads to an internal node, 0 = to a leaf (Prior.jl:15-23)
        std::vector<char> internal(NN, 0);
        for (int j = 0; j < NN; ++j) {
            const int m = pa[j];
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
        return "unknown branch-length prior kind";
    }
    pr[0] = 1.0;
    return nullptr;
}

int validate_batch(mcp_ctx* ctx, const BatchArgs& a, int* K_out) {
    if (a.T < 1) return fail(ctx, MCP_ERR_ARG, "batch must hold at least one tree");
    if (a.R < 1) return fail(ctx, MCP_ERR_ARG, "need at least one rate category");
    if (a.R > MAX_RATES) return fail(ctx, MCP_ERR_UNSUPPORTED, "more than %d rate categories", MAX_RATES);
    for (int t = 0; t < a.T; ++t) {
        if (!a.alns[t] || !a.po[t] || !a.pa[t] || !a.blv[t] || !a.U[t] || !a.D[t] || !a.Uinv[t] || !a.rates[t] || !a.pi[t])
            return fail(ctx, MCP_ERR_ARG, "tree %d: null argument", t);
        if (a.alns[t]->K != a.alns[0]->K) return fail(ctx, MCP_ERR_ARG, "all alignments of a batch must share K");
        if (a.alns[t]->owner != ctx) return fail(ctx, MCP_ERR_ARG, "tree %d: the alignment belongs to another context", t);
    }
    const int K = a.alns[0]->K;
    if (!k_supported(K)) return fail(ctx, MCP_ERR_UNSUPPORTED, "no kernel compiled for K = %d states", K);
    if (k_templated(K) && 2 * K * K + K + a.R * K > MODEL_SLOT)
        return fail(ctx, MCP_ERR_UNSUPPORTED, "model with K=%d, R=%d does not fit a constant slot", K, a.R);
    if (a.prior_kind != MCP_PRIOR_NONE) {   // parameter errors surface before anything is planned or enqueued
        if (a.prior_kind != MCP_PRIOR_EXPONENTIAL && a.prior_kind != MCP_PRIOR_COMPOUND_DIRICHLET)
            return fail(ctx, MCP_ERR_ARG, "unknown branch-length prior kind %d", a.prior_kind);
        if (!a.prior_params) return fail(ctx, MCP_ERR_ARG, "branch-length prior without parameters");
        const int np = a.prior_kind == MCP_PRIOR_EXPONENTIAL ? 1 : 4;
        for (int i = 0; i < np; ++i)
            if (!(a.prior_params[i] > 0.0))
                return fail(ctx, MCP_ERR_ARG, a.prior_kind == MCP_PRIOR_EXPONENTIAL ? "exponentialBL: scale must be positive"
                                                                                    : "CompoundDirichlet: alpha, a, beta, c must be positive");
    }
    if (a.model_grad && (a.T != 1 || !a.want_grad))
        return fail(ctx, MCP_ERR_ARG, "a model-gradient evaluation takes one tree and computes the branch gradient with it");
    *K_out = K;
    return 0;
}

// One evaluation on ONE device.  d_out_user != null: the result [logL, grad] per tree is left there
// (device-addressable memory: this GPU, a peer GPU, or pinned host memory) and nothing is synchronised;
// otherwise the call returns the results in ll_out / grad_out.
int eval_impl(mcp_ctx* ctx, const BatchArgs& a, double* d_out_user, double* ll_out, double* const* grad_out) {
    if (!ctx) return fail(nullptr, MCP_ERR_ARG, "null context");
    int K = 0, e;
    if ((e = validate_batch(ctx, a, &K))) return e;
    const int R = a.R, T = a.T;
    CUDA_TRY(ctx, cudaSetDevice(ctx->device));
    Plan* plp = nullptr;
    bool rebuilt = false;
    if ((e = get_plan(ctx, a, K, &plp, &rebuilt))) return e;
    Plan& pl = *plp;
    ctx->last_plan = plp;
    const bool templ = k_templated(K) && !a.model_grad;     // model-gradient evaluations run the runtime-K kernel
    const KernelTable* kt = a.model_grad ? mcpdev::kernels_generic() : kernels_for(K);
    if (a.model_grad && (d_out_user || ctx->sf)) return fail(ctx, MCP_ERR_ARG, "internal: model-gradient evaluation must be synchronous and resident");
    // model-gradient moments: [device branch][rate][K x K] and the root vector W[K] (kernel_generic.cuh)
    const size_t mg_doubles = a.model_grad ? (((size_t)pl.trees[0].n_br * R * K * K + K + 3) & ~(size_t)3) : 0;
    // replicas against same-address atomics (kernel_generic.cuh): as many as 64 MB hold, at most 32 and at most one per CTA
    static const int mg_rep_env = []() { const char* v = std::getenv("MCPHYLO_B200_MG_REPLICAS"); return v ? std::atoi(v) : 0; }();
    int mg_rep = 1;
    if (a.model_grad) {
        mg_rep = (int)std::min<size_t>(32, std::max<size_t>(1, ((size_t)64 << 20) / (sizeof(double) * mg_doubles)));
        if (mg_rep_env > 0) mg_rep = mg_rep_env;
        mg_rep = std::max(1, std::min(mg_rep, pl.grid));
        if ((e = ensure_dev(ctx, ctx->d_mg, sizeof(double) * mg_doubles * (mg_rep + 1)))) return e;   // + the folded result
        if ((e = ensure_pin(ctx, ctx->h_mg, sizeof(double) * mg_doubles))) return e;
    }

    // buffers
    const int slot = ctx->stage_next;
    ctx->stage_next = (slot + 1) % MCP_STAGE_SLOTS;
    if (ctx->staged_pending[slot]) {   // an earlier asynchronous evaluation may still be reading this staging slot
        CUDA_TRY(ctx, cudaEventSynchronize(ctx->ev_staged[slot]));
        ctx->staged_pending[slot] = false;
    }
    if ((e = ensure_pin(ctx, ctx->h_dyn[slot], sizeof(double) * pl.total_dyn))) return e;
    if ((e = ensure_pin(ctx, ctx->h_model[slot], sizeof(double) * MODEL_SLOT * MODEL_SLOTS))) return e;
    if ((e = ensure_dev(ctx, ctx->d_dyn, sizeof(double) * pl.total_dyn))) return e;
    if ((e = ensure_dev(ctx, ctx->d_btab, sizeof(double) * pl.total_btab))) return e;
    if ((e = ensure_dev(ctx, ctx->d_scratch, sizeof(double) * pl.scratch_per_cta * pl.grid))) return e;
    // + 64 rows: the group rows of the small-tree kernel's two-level final reduction (kernel_levels.cuh)
    if ((e = ensure_dev(ctx, ctx->d_rows, sizeof(double) * pl.row_stride * (pl.n_rows + 64)))) return e;
    if ((e = ensure_dev(ctx, ctx->d_rows_ll, sizeof(LLRow) * (pl.n_rows + 64)))) return e;
    const bool fused = pl.level_mode;   // small-tree kernel: tables, walk and final reduction in ONE launch
    const bool via_comm = !d_out_user && ctx->rank_comm != nullptr;
    double* d_out = d_out_user;
    if (!d_out) {
        if ((e = ensure_dev(ctx, ctx->d_out, sizeof(double) * pl.total_out))) return e;
        if ((e = ensure_pin(ctx, ctx->h_out, sizeof(double) * pl.total_out))) return e;
        d_out = (double*)ctx->d_out.p;
    }
    if (fused && !ctx->d_counter.p) {
        if ((e = ensure_dev(ctx, ctx->d_counter, 256))) return e;
        CUDA_TRY(ctx, cudaMemsetAsync(ctx->d_counter.p, 0, 256, ctx->stream));
    }

    // per-evaluation parameters
    double* hd = (double*)ctx->h_dyn[slot].p;
    bool all_null_last = true;
    const bool prior_here = a.prior_kind != MCP_PRIOR_NONE && (ctx->rank_comm == nullptr || ctx->rank == 0);
    for (int t = 0; t < T; ++t) {
        const int NN = a.NN[t];
        double* d = hd + pl.trees[t].dyn_off;
        std::memcpy(d + dyn_blv(NN), a.blv[t], sizeof(double) * (NN - 1));
        // eigen-decomposition with a null eigenvalue (if any) moved to the last position
        all_null_last = null_eigenvalue_last(a.U[t], a.D[t], a.Uinv[t], K, d + dyn_U(NN), d + dyn_D(NN, K), d + dyn_Uinv(NN, K)) && all_null_last;
        d[dyn_mu(NN, K)] = a.mu[t];
        std::memcpy(d + dyn_rates(NN, K), a.rates[t], sizeof(double) * R);
        std::memcpy(d + dyn_pi(NN, K, R), a.pi[t], sizeof(double) * K);
        const char* perr = fill_prior(prior_here ? a.prior_kind : MCP_PRIOR_NONE, a.prior_params, NN, a.pa[t], d + dyn_prior(NN, K, R));
        if (perr) return fail(ctx, MCP_ERR_ARG, "%s", perr);
    }

    // substitution-model constants: one slot per distinct model of the batch
    const int model_doubles = templ ? 2 * K * K + K + R * K : 0;
    double* hm = (double*)ctx->h_model[slot].p;
    int n_models = 0;
    for (int t = 0; t < T && templ; ++t) {
        double cand[MODEL_SLOT];
        const double* dt = hd + pl.trees[t].dyn_off;       // the permuted decomposition stored above
        std::memcpy(cand, dt + dyn_U(a.NN[t]), sizeof(double) * K * K);
        std::memcpy(cand + K * K, dt + dyn_Uinv(a.NN[t], K), sizeof(double) * K * K);
        std::memcpy(cand + 2 * K * K, a.pi[t], sizeof(double) * K);
        for (int r = 0; r < R; ++r)
            for (int i = 0; i < K; ++i) cand[2 * K * K + K + r * K + i] = dt[dyn_D(a.NN[t], K) + i] * a.rates[t][r] * a.mu[t];
        int ms = -1;
        for (int m = 0; m < n_models && ms < 0; ++m)
            if (std::memcmp(hm + (size_t)m * MODEL_SLOT, cand, sizeof(double) * model_doubles) == 0) ms = m;
        if (ms < 0) {
            if (n_models == MODEL_SLOTS)
                return fail(ctx, MCP_ERR_UNSUPPORTED, "more than %d distinct substitution models in one batch", MODEL_SLOTS);
            ms = n_models++;
            std::memcpy(hm + (size_t)ms * MODEL_SLOT, cand, sizeof(double) * model_doubles);
        }
        hd[pl.trees[t].dyn_off + dyn_slot(a.NN[t], K, R)] = (double)ms;
    }
    // the walk kernels with all K eigen-components read their model from constant memory only
    const bool dyn_model = n_models > 1 || (templ && (!all_null_last || pl.acc_global));

    // ---- everything below only enqueues; from the first enqueue on, an error leaves the plan marked
    // "not uploaded" so that a retry re-sends the topology block instead of trusting a half-done one ----
    cudaStream_t st = ctx->stream;
    mcp_stats& s = ctx->stats;
    s = mcp_stats{};
    struct UploadGuard {
        Plan& pl; bool ok = false; bool was;
        explicit UploadGuard(Plan& p) : pl(p), was(p.uploaded) {}
        ~UploadGuard() { if (!ok) pl.uploaded = false; }
    } guard(pl);
    if (ctx->opt_timing) CUDA_TRY(ctx, cudaEventRecord(ctx->ev[0], st));
    if (dyn_model) {   // several models: slots in constant memory; one model travels as a kernel parameter
        std::lock_guard<std::mutex> lock(g_slots.mu);
        ModelSlotGuard::Entry* en = g_slots.find(ctx->device, kt);
        if (en && en->last != ctx) CUDA_TRY(ctx, cudaStreamWaitEvent(st, en->ev, 0));
        cudaError_t ce = kt->upload_model(hm, sizeof(double) * MODEL_SLOT * n_models, st);
        if (ce != cudaSuccess) return fail(ctx, MCP_ERR_CUDA, "model upload failed: %s", cudaGetErrorString(ce));
        s.h2d_bytes += (int64_t)(sizeof(double) * MODEL_SLOT * n_models);
    }
    // host -> device through a kernel that reads the pinned staging buffers (see stage_from_host): the copy
    // engine stays free for bulk alignment transfers and cannot delay an evaluation behind them
    auto stage = [&](const void* h, void* d, size_t bytes) -> cudaError_t {
        const long long n16 = (long long)((bytes + 15) / 16);
        const unsigned blocks = (unsigned)std::min<long long>((n16 + 255) / 256, 4LL * ctx->sm_count);
        stage_from_host<<<std::max(blocks, 1u), 256, 0, st>>>((const uint4*)h, (uint4*)d, n16);
        return cudaGetLastError();
    };
    if (!pl.uploaded) {
        CUDA_TRY(ctx, stage(pl.h_topo.p, pl.d_topo.p, pl.topo_bytes));
        s.h2d_bytes += (int64_t)pl.topo_bytes;
        pl.uploaded = true;
    }
    // MCMC-sized inputs on the fused small-tree kernel: the parameter block rides in the kernel arguments -- one
    // launch per evaluation, nothing read from pinned host memory in front of it (LevelParams, device_layout.cuh)
    static const bool inline_allowed = []() { const char* v = std::getenv("MCPHYLO_B200_INLINE_PARAMS"); return !(v && v[0] == '0'); }();
    const bool inline_dyn = fused && pl.total_dyn <= (long long)LEVEL_DYN_INLINE && inline_allowed;
    if (!inline_dyn) {
        CUDA_TRY(ctx, stage(hd, ctx->d_dyn.p, sizeof(double) * pl.total_dyn));
        CUDA_TRY(ctx, cudaEventRecord(ctx->ev_staged[slot], st));
        ctx->staged_pending[slot] = true;
    }
    s.h2d_bytes += (int64_t)(sizeof(double) * pl.total_dyn);
    // resident alignment, one tree: tiles by atomic ticket (site-major order) instead of static ranges
    const bool dyn_tiles = !ctx->sf && ctx->opt_dynamic == 1 && !pl.level_mode && !pl.acc_global && templ && T == 1;
    if (dyn_tiles) {
        if ((e = ensure_dev(ctx, ctx->d_ticket, sizeof(unsigned) * STREAM_CTL_WORDS))) return e;
        CUDA_TRY(ctx, cudaMemsetAsync(ctx->d_ticket.p, 0, sizeof(unsigned) * STREAM_CTL_WORDS, st));
    }
    if (ctx->sf) {
        if (pl.level_mode || pl.acc_global || !k_templated(K) || T != 1)
            return fail(ctx, MCP_ERR_ARG, "internal: streamed launch planned for a kernel without ready flags");
        CUDA_TRY(ctx, cudaMemsetAsync(ctx->sf->ticket, 0, sizeof(unsigned) * STREAM_CTL_WORDS, st));
    }
    for (int t = 0; t < T && !ctx->sf; ++t)   // re-uploads of these alignments that are still in flight on the copy stream
        if (a.alns[t]->upload_pending) {
            CUDA_TRY(ctx, cudaStreamWaitEvent(st, a.alns[t]->ev_uploaded, 0));
            a.alns[t]->upload_pending = false;
        }

    const TreeDev* d_trees = (const TreeDev*)((char*)pl.d_topo.p + pl.off_trees);
    if (!fused) {
        dim3 grid((pl.max_br * R + 127) / 128, T);
        build_branch_tables<<<grid, 128, 0, st>>>(d_trees, (const double*)ctx->d_dyn.p, (double*)ctx->d_btab.p, K, R);
        CUDA_TRY(ctx, cudaGetLastError());
    }
    WalkParams wp;
    wp.trees = d_trees;
    wp.ops = (const int4*)((char*)pl.d_topo.p + pl.off_ops);
    wp.btab = (const double*)ctx->d_btab.p;
    wp.dyn = (const double*)ctx->d_dyn.p;
    wp.scratch = (double*)ctx->d_scratch.p;
    wp.scratch_per_cta = pl.scratch_per_cta;
    wp.rows = (double*)ctx->d_rows.p;
    wp.rows_ll = (LLRow*)ctx->d_rows_ll.p;
    wp.cta_row_base = (const int*)((char*)pl.d_topo.p + pl.off_rowbase);
    wp.levels = (const int*)((char*)pl.d_topo.p + pl.off_levels);
    wp.fetch = (const unsigned short*)((char*)pl.d_topo.p + pl.off_fetch);
    // pinned host memory is device-addressable (unified addressing): the fused kernel writes results there
    wp.out = d_out_user ? d_out_user : via_comm ? d_out : (double*)ctx->h_out.p;
    wp.done_counter = (unsigned int*)ctx->d_counter.p;
    wp.row_stride = pl.row_stride;
    wp.n_slots = pl.n_slots;
    wp.n_stack = pl.n_stack;
    wp.n_tiles = pl.n_tiles;
    wp.T = T;
    wp.R = R;
    wp.want_grad = a.want_grad ? 1 : 0;
    wp.max_br = pl.max_br;
    wp.max_rows = pl.max_rows;
    wp.ready_flags = ctx->sf ? ctx->sf->flags : nullptr;
    wp.ticket = ctx->sf ? ctx->sf->ticket : dyn_tiles ? (unsigned*)ctx->d_ticket.p : nullptr;
    wp.error_flag = ctx->sf ? ctx->sf->error : nullptr;
    wp.ready_epoch = ctx->sf ? ctx->sf->epoch : 0u;
    wp.ready_shift = ctx->sf ? ctx->sf->shift : 0;
    static_assert(sizeof(wp.model) / sizeof(double) >= 2 * 6 * 6 + 6 + MAX_RATES * 6, "model parameter block too small");
    std::memset(wp.model, 0, sizeof wp.model);
    if (n_models == 1) std::memcpy(wp.model, hm, sizeof(double) * model_doubles);
    if (ctx->opt_timing) CUDA_TRY(ctx, cudaEventRecord(ctx->ev[1], st));
    if (ctx->ev_walk_begin) CUDA_TRY(ctx, cudaEventRecord(ctx->ev_walk_begin, st));
    LaunchCfg lc;
    lc.device = ctx->device;
    lc.K = K;
    lc.grid = pl.grid;
    lc.block = pl.block;
    lc.cpt = pl.cpt;
    lc.ring = pl.ring && !dyn_model && all_null_last;
    lc.smem = lc.ring ? pl.smem_ring : pl.smem_bytes;
    lc.stream = st;
    lc.smem_scratch = pl.smem_scratch;
    lc.acc_global = pl.acc_global;
    lc.mma = pl.mma;
    if (a.model_grad) {
        CUDA_TRY(ctx, cudaMemsetAsync(ctx->d_mg.p, 0, sizeof(double) * mg_doubles * mg_rep, st));
        lc.mg = (double*)ctx->d_mg.p;
        lc.mg_rep = mg_rep;
        lc.mg_stride = (long long)mg_doubles;
    }
    {
        cudaError_t ce = pl.level_mode ? kt->launch_levels(lc, wp, dyn_model, inline_dyn ? hd : nullptr, (size_t)pl.total_dyn)
                                       : kt->launch_walk(lc, wp, dyn_model, all_null_last);
        if (ce != cudaSuccess) return fail(ctx, MCP_ERR_CUDA, "walk kernel launch failed: %s", cudaGetErrorString(ce));
    }
    if (ctx->opt_timing) CUDA_TRY(ctx, cudaEventRecord(ctx->ev[2], st));
    if (ctx->ev_walk_end) CUDA_TRY(ctx, cudaEventRecord(ctx->ev_walk_end, st));
    if (dyn_model) {   // the next writer of this unit's model slots on this device waits for this kernel
        std::lock_guard<std::mutex> lock(g_slots.mu);
        ModelSlotGuard::Entry* en = g_slots.find(ctx->device, kt);
        if (!en) {
            cudaEvent_t ev = nullptr;
            CUDA_TRY(ctx, cudaEventCreateWithFlags(&ev, cudaEventDisableTiming));
            g_slots.entries.push_back({ctx->device, kt, ev, ctx});
            en = &g_slots.entries.back();
        }
        CUDA_TRY(ctx, cudaEventRecord(en->ev, st));
        en->last = ctx;
    }
    double* const d_mg_result = a.model_grad ? (double*)ctx->d_mg.p + (mg_rep > 1 ? (size_t)mg_rep * mg_doubles : 0) : nullptr;
    if (a.model_grad && mg_rep > 1) {   // fold the replicas, in replica order, into the slot behind them
        sum_rows<<<(unsigned)((mg_doubles + 255) / 256), 256, 0, st>>>((const double*)ctx->d_mg.p, (long long)mg_doubles, mg_rep,
                                                                        (long long)mg_doubles, d_mg_result);
        CUDA_TRY(ctx, cudaGetLastError());
    }
    CUDA_TRY(ctx, cudaEventRecord(ctx->ev_walk_done, st));
    for (int t = 0; t < T; ++t) {
        a.alns[t]->read_since_upload = true;
        if (a.alns[t]->streamed) CUDA_TRY(ctx, cudaEventRecord(a.alns[t]->ev_read_done, st));
    }
    if (!fused) {
        int maxNN = 0;
        for (int t = 0; t < T; ++t) maxNN = std::max(maxNN, a.NN[t]);
        dim3 grid((maxNN + FIN_J - 1) / FIN_J, T);
        finalize_results<<<grid, dim3(FIN_J, FIN_G), 0, st>>>(d_trees, (const double*)ctx->d_rows.p, pl.row_stride,
                                               (const LLRow*)ctx->d_rows_ll.p, d_out, wp.want_grad,
                                               (const double*)ctx->d_dyn.p, K, R);
        CUDA_TRY(ctx, cudaGetLastError());
    }
    s.kernel_launches = (fused ? 1 : 3) + (inline_dyn ? 0 : 1) + (rebuilt ? 1 : 0) + (a.model_grad && mg_rep > 1 ? 1 : 0);   // + parameter staging (+ topology staging) (+ fold of the moment replicas)
    s.grid = pl.grid;
    s.block = pl.block;
    s.tiles = pl.n_tiles;
    s.columns_per_thread = pl.cpt;
    s.operand_ring = lc.ring ? WALK_RING_DEPTH : 0;
    s.schedule_rebuilt = rebuilt ? 1 : 0;
    s.scratch_bytes = (int64_t)ctx->d_scratch.cap;
    guard.ok = true;
    if (d_out_user) {
        if (ctx->opt_timing) CUDA_TRY(ctx, cudaEventRecord(ctx->ev[3], st));
        ctx->pending_async = true;
        return 0;
    }
    if (via_comm) {   // one rank of a multi-process group: sum [logL, grad] over the ranks, then read it back
        const mcpnccl::Api& nc = mcpnccl::api();
        mcpnccl::result_t nr = nc.AllReduce(d_out, d_out, (size_t)pl.total_out, mcpnccl::kFloat64, mcpnccl::kSum, ctx->rank_comm, st);
        if (nr) return fail(ctx, MCP_ERR_CUDA, "ncclAllReduce failed: %s", nc.GetErrorString(nr));
        if (a.model_grad) {
            nr = nc.AllReduce(d_mg_result, d_mg_result, mg_doubles, mcpnccl::kFloat64, mcpnccl::kSum, ctx->rank_comm, st);
            if (nr) return fail(ctx, MCP_ERR_CUDA, "ncclAllReduce (model-gradient moments) failed: %s", nc.GetErrorString(nr));
        }
    }
    if (!fused || via_comm) CUDA_TRY(ctx, cudaMemcpyAsync(ctx->h_out.p, d_out, sizeof(double) * pl.total_out, cudaMemcpyDeviceToHost, st));
    s.d2h_bytes = (int64_t)(sizeof(double) * pl.total_out);
    if (a.model_grad) {
        CUDA_TRY(ctx, cudaMemcpyAsync(ctx->h_mg.p, d_mg_result, sizeof(double) * mg_doubles, cudaMemcpyDeviceToHost, st));
        s.d2h_bytes += (int64_t)(sizeof(double) * mg_doubles);
    }
    if (ctx->opt_timing) CUDA_TRY(ctx, cudaEventRecord(ctx->ev[3], st));
    CUDA_TRY(ctx, cudaStreamSynchronize(st));
    for (bool& p : ctx->staged_pending) p = false;
    ctx->pending_async = false;
    const double* ho = (const double*)ctx->h_out.p;
    for (int t = 0; t < T; ++t) {
        const double* o = ho + pl.trees[t].out_off;
        if (ll_out) ll_out[t] = o[0];
        if (a.want_grad && grad_out && grad_out[t]) std::memcpy(grad_out[t], o + 1, sizeof(double) * (a.NN[t] - 1));
    }
    return 0;
}

int make_alignment(mcp_ctx* ctx, const unsigned char* codes, size_t src_pitch, int K, long long S, const int32_t* leaf_nums,
                   int n_leaves, mcp_alignment** out) {
    CUDA_TRY(ctx, cudaSetDevice(ctx->device));
    mcp_alignment* al = new mcp_alignment();
    al->K = K;
    al->S = S;
    al->owner = ctx;
    al->stride = (S + 1023) & ~1023LL;
    if (al->stride == 0) al->stride = 1024;
    al->n_leaves = n_leaves;
    al->leaf_nums.assign(leaf_nums, leaf_nums + n_leaves);
    al->id = ctx->next_aln_id++;
    // One tile of slack behind the last row: a tile whose width does not divide the row stride (widths are
    // multiples of 32, up to 512 columns) may stage codes past the end of a row; the values are masked, the
    // read must stay inside the allocation.
    size_t bytes = (size_t)al->stride * (size_t)std::max(n_leaves, 1) + 1024;
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
        e = cudaMemcpy2DAsync(al->d_codes, (size_t)al->stride, codes, src_pitch, (size_t)S, (size_t)n_leaves,
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

int update_codes_one(mcp_ctx* ctx, mcp_alignment* aln, const unsigned char* codes, size_t src_pitch) {
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
        CUDA_TRY(ctx, cudaMemcpy2DAsync(aln->d_codes, (size_t)aln->stride, codes, src_pitch, (size_t)aln->S,
                                        (size_t)aln->n_leaves, cudaMemcpyHostToDevice, ctx->copy_stream));
    CUDA_TRY(ctx, cudaEventRecord(aln->ev_uploaded, ctx->copy_stream));
    aln->upload_pending = true;
    ctx->pending_async = true;   // only consulted before buffers are freed / the stream is changed
    return 0;
}

void destroy_alignment_one(mcp_ctx* ctx, mcp_alignment* aln) {
    if (ctx) {
        cudaSetDevice(ctx->device);
        cudaStreamSynchronize(ctx->stream);
        if (ctx->copy_stream) cudaStreamSynchronize(ctx->copy_stream);
        ctx->pending_async = false;
        invalidate_plans(ctx, aln->id);
    }
    if (aln->ev_uploaded) cudaEventDestroy(aln->ev_uploaded);
    if (aln->ev_read_done) cudaEventDestroy(aln->ev_read_done);
    if (aln->d_codes) cudaFree(aln->d_codes);
    delete aln;
}

int dense_to_codes(mcp_ctx* ctx, const double* x, int K, int64_t S, int NN, const int32_t* leaf_nums, int n_leaves,
                   std::vector<unsigned char>& codes) {
    codes.resize((size_t)n_leaves * (size_t)S);
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
    return 0;
}

void shard_range(long long S, int G, int g, long long* lo, long long* hi) {
    const long long per = (S + G - 1) / G;
    *lo = std::min(S, (long long)g * per);
    *hi = std::min(S, *lo + per);
}

// --------------------------------------------------------------------------------------------
// multi-device groups
// --------------------------------------------------------------------------------------------
// Runs fn(g) for every member of a group, member g > 0 on its own persistent host thread (MemberPool).
// Returns the first non-zero result.
int for_each_member(mcp_ctx* ctx, const std::function<int(int)>& fn) {
    const int G = (int)ctx->members.size();
    std::vector<int> serial_rc;
    const std::vector<int>* rc = &serial_rc;
    if (G == 1 || !ctx->pool) {
        serial_rc.resize(G);
        for (int g = 0; g < G; ++g) serial_rc[g] = fn(g);
    } else {
        ctx->pool->run(fn);
        rc = &ctx->pool->rc;
    }
    for (int g = 0; g < G; ++g)
        if ((*rc)[g]) {
            ctx->error = "device " + std::to_string(ctx->members[g]->device) + ": " + ctx->members[g]->error;
            return (*rc)[g];
        }
    return 0;
}

// Sums the members' result vectors (`len` doubles each; member g's lives at part[g], see group_targets)
// and returns the sum in pinned host memory (*h_result).
int group_reduce(mcp_ctx* ctx, const std::vector<double*>& part, long long len, const double** h_result) {
    const int G = (int)ctx->members.size();
    mcp_ctx* m0 = ctx->members[0];
    int e;
    CUDA_TRY(ctx, cudaSetDevice(m0->device));
    if ((e = ensure_pin(m0, m0->h_out, sizeof(double) * len))) { ctx->error = m0->error; return e; }
    if (ctx->reduce_mode == MCP_REDUCE_NCCL) {
        const mcpnccl::Api& nc = mcpnccl::api();
        mcpnccl::result_t nr = nc.GroupStart();
        for (int g = 0; g < G && !nr; ++g)
            nr = nc.AllReduce(part[g], part[g], (size_t)len, mcpnccl::kFloat64, mcpnccl::kSum, ctx->comms[g], ctx->members[g]->stream);
        mcpnccl::result_t ne = nc.GroupEnd();
        if (nr || ne) return fail(ctx, MCP_ERR_CUDA, "ncclAllReduce failed: %s", nc.GetErrorString(nr ? nr : ne));
        CUDA_TRY(ctx, cudaSetDevice(m0->device));
        CUDA_TRY(ctx, cudaMemcpyAsync(m0->h_out.p, part[0], sizeof(double) * len, cudaMemcpyDeviceToHost, m0->stream));
        CUDA_TRY(ctx, cudaStreamSynchronize(m0->stream));
        m0->pending_async = false;
    } else if (ctx->reduce_mode == MCP_REDUCE_PEER) {
        // every member has written its part into device 0's gather buffer (NVLink peer stores issued by its
        // own final-reduction kernel); device 0 waits for them and adds the parts in member order
        for (int g = 1; g < G; ++g) {
            CUDA_TRY(ctx, cudaSetDevice(ctx->members[g]->device));
            CUDA_TRY(ctx, cudaEventRecord(ctx->ev_member[g], ctx->members[g]->stream));
        }
        CUDA_TRY(ctx, cudaSetDevice(m0->device));
        for (int g = 1; g < G; ++g) CUDA_TRY(ctx, cudaStreamWaitEvent(m0->stream, ctx->ev_member[g], 0));
        sum_rows<<<(unsigned)((len + 255) / 256), 256, 0, m0->stream>>>(part[0], len, G, len, (double*)m0->h_out.p);
        CUDA_TRY(ctx, cudaGetLastError());
        CUDA_TRY(ctx, cudaStreamSynchronize(m0->stream));
        for (mcp_ctx* m : ctx->members) m->pending_async = false;
    } else {   // MCP_REDUCE_HOST
        for (int g = 0; g < G; ++g) {
            mcp_ctx* m = ctx->members[g];
            CUDA_TRY(ctx, cudaSetDevice(m->device));
            if ((e = ensure_pin(m, ctx->h_member_out[g], sizeof(double) * len))) { ctx->error = m->error; return e; }
            CUDA_TRY(ctx, cudaMemcpyAsync(ctx->h_member_out[g].p, part[g], sizeof(double) * len, cudaMemcpyDeviceToHost, m->stream));
        }
        for (int g = 0; g < G; ++g) {
            CUDA_TRY(ctx, cudaStreamSynchronize(ctx->members[g]->stream));
            ctx->members[g]->pending_async = false;
        }
        double* o = (double*)m0->h_out.p;
        std::memcpy(o, ctx->h_member_out[0].p, sizeof(double) * len);
        for (int g = 1; g < G; ++g) {
            const double* pg = (const double*)ctx->h_member_out[g].p;
            for (long long j = 0; j < len; ++j) o[j] += pg[j];
        }
    }
    // parameter staging slots are free again wherever the member's stream is known to have drained
    for (mcp_ctx* m : ctx->members)
        if (m == m0 || ctx->reduce_mode != MCP_REDUCE_NCCL)
            for (bool& p : m->staged_pending) p = false;
    *h_result = (const double*)m0->h_out.p;
    return 0;
}

// Where each member leaves its [logL, grad] part of `len` doubles.  PEER: rows of one buffer on device 0
// (the other members reach it through peer access); NCCL / HOST: a buffer on the member's own device.
int group_targets(mcp_ctx* ctx, long long len, std::vector<double*>& part) {
    const int G = (int)ctx->members.size();
    part.assign(G, nullptr);
    int e;
    if (ctx->reduce_mode == MCP_REDUCE_PEER) {
        mcp_ctx* m0 = ctx->members[0];
        CUDA_TRY(ctx, cudaSetDevice(m0->device));
        if (sizeof(double) * len * G > ctx->d_gather.cap) {   // growing it: nobody may still be writing into the old one
            for (mcp_ctx* m : ctx->members) {
                cudaSetDevice(m->device);
                cudaStreamSynchronize(m->stream);
            }
            CUDA_TRY(ctx, cudaSetDevice(m0->device));
        }
        if ((e = ensure_dev(m0, ctx->d_gather, sizeof(double) * len * G))) { ctx->error = m0->error; return e; }
        for (int g = 0; g < G; ++g) part[g] = (double*)ctx->d_gather.p + (long long)g * len;
    } else {
        for (int g = 0; g < G; ++g) {
            mcp_ctx* m = ctx->members[g];
            CUDA_TRY(ctx, cudaSetDevice(m->device));
            if ((e = ensure_dev(m, m->d_out, sizeof(double) * len))) { ctx->error = m->error; return e; }
            part[g] = (double*)m->d_out.p;
        }
    }
    return 0;
}

void unpack_results(const BatchArgs& a, const double* h, double* ll_out, double* const* grad_out) {
    long long off = 0;
    for (int t = 0; t < a.T; ++t) {
        if (ll_out) ll_out[t] = h[off];
        if (a.want_grad && grad_out && grad_out[t]) std::memcpy(grad_out[t], h + off + 1, sizeof(double) * (a.NN[t] - 1));
        off += a.NN[t];
    }
}

int group_eval(mcp_ctx* ctx, const BatchArgs& a, double* ll_out, double* const* grad_out) {
    const int G = (int)ctx->members.size();
    int K = 0, e;
    if ((e = validate_batch(ctx, a, &K))) return e;
    long long len = 0;
    for (int t = 0; t < a.T; ++t) {
        if ((int)a.alns[t]->shards.size() != G) return fail(ctx, MCP_ERR_ARG, "tree %d: alignment was not created on this multi-device context", t);
        len += a.NN[t];
    }
    std::vector<double*> part;
    if ((e = group_targets(ctx, len, part))) return e;
    e = for_each_member(ctx, [&](int g) -> int {
        std::vector<const mcp_alignment*> alns(a.T);
        for (int t = 0; t < a.T; ++t) alns[t] = a.alns[t]->shards[g];
        BatchArgs ag = a;
        ag.alns = alns.data();
        if (g > 0) ag.prior_kind = MCP_PRIOR_NONE;   // the prior is added once, by member 0
        return eval_impl(ctx->members[g], ag, part[g], nullptr, nullptr);
    });
    if (e) return e;
    const double* h = nullptr;
    if ((e = group_reduce(ctx, part, len, &h))) return e;
    unpack_results(a, h, ll_out, grad_out);
    return 0;
}

int eval_any(mcp_ctx* ctx, const BatchArgs& a, double* ll_out, double* const* grad_out) {
    if (!ctx) return fail(nullptr, MCP_ERR_ARG, "null context");
    return ctx->members.empty() ? eval_impl(ctx, a, nullptr, ll_out, grad_out) : group_eval(ctx, a, ll_out, grad_out);
}

// --------------------------------------------------------------------------------------------
// streamed evaluation: alignment in host memory, uploaded block by block under the evaluation
// --------------------------------------------------------------------------------------------
// Site blocks of one device's range [lo, hi).  The transfer of block b+1 runs under the evaluation of block b,
// so it stays hidden as long as  size(b+1) * t_upload <= size(b) * t_eval  per site.  For cfg4 (1000 taxa, K = 4,
// Gamma-4) a site costs ~38 ns to upload (1000 bytes at ~26 GB/s) and ~74 ns to evaluate on one GPU, i.e. blocks
// may grow by at most ~1.9x: they grow by 1.5x, in whole waves of the persistent grid once the range is long
// enough (no ragged last round), the first block is one wave at most (its transfer is the only one exposed) and
// the last block takes the rest while that is under 1.8x its predecessor.
void plan_stream_blocks(long long lo, long long hi, long long wave_sites, std::vector<std::pair<long long, long long>>& out) {
    out.clear();
    const long long n = hi - lo;
    if (n <= 0) return;
    wave_sites = std::max<long long>(wave_sites, 512);
    const bool whole_waves = n >= 8 * wave_sites;
    long long sz = whole_waves ? wave_sites : std::min(wave_sites, std::max<long long>((n / 4 + 511) & ~511LL, 4096));
    long long at = lo;
    while ((int)out.size() < 15 && (double)(hi - at - sz) > 1.8 * (double)sz) {
        out.push_back({at, at + sz});
        at += sz;
        long long next = sz + sz / 2;
        if (whole_waves) next = std::max(sz + wave_sites, next / wave_sites * wave_sites);
        else next = (next + 511) & ~511LL;
        sz = next;
    }
    if ((double)(hi - at) > 1.8 * (double)sz && (int)out.size() < 15) {
        out.push_back({at, at + sz});
        at += sz;
    }
    out.push_back({at, hi});
}

int wave_columns_impl(mcp_ctx* ctx, int K, int n_nodes, int want_grad, int64_t* columns) {
    const KernelTable* kt = kernels_for(K);
    if (!kt) return fail(ctx, MCP_ERR_UNSUPPORTED, "no kernel compiled for K = %d states", K);
    CUDA_TRY(ctx, cudaSetDevice(ctx->device));
    int block = ctx->opt_block > 0 ? ctx->opt_block : 256;
    int cpt = ctx->opt_cpt > 0 ? ctx->opt_cpt : (K <= 4 ? 2 : 1);   // what the planner picks for large inputs
    int occ = 0, e;
    if (k_templated(K)) {
        const bool acc_global = want_grad && walk_acc_global(n_nodes, ctx->opt_acc_mode);
        if (acc_global) cpt = 1;
        const size_t smem = walk_smem_bytes(K, std::max(n_nodes, 1), want_grad && !acc_global ? 1 : 0, block, cpt);
        if ((e = walk_occupancy(ctx, kt, K, block, cpt, smem, false, acc_global, false, &occ))) return e;
    } else if (ctx->opt_mma != 0) {
        if ((e = walk_occupancy(ctx, kt, K, MMA_WARPS * 32, 1, mma_smem_bytes(K, std::max(n_nodes, 1), want_grad), false, false, false, &occ, true))) return e;
        if (occ < 1) return fail(ctx, MCP_ERR_CUDA, "walk kernel does not fit on an SM");
        if (ctx->opt_ctas_per_sm > 0) occ = std::min(occ, ctx->opt_ctas_per_sm);
        *columns = (int64_t)occ * ctx->sm_count * MMA_TILE;
        return 0;
    } else {
        cpt = 1;
        block = std::min(block, 128);
        if ((e = walk_occupancy(ctx, kt, K, block, cpt, generic_smem_bytes(std::max(n_nodes, 1), want_grad), false, false, false, &occ))) return e;
    }
    if (occ < 1) return fail(ctx, MCP_ERR_CUDA, "walk kernel does not fit on an SM (block %d)", block);
    if (ctx->opt_ctas_per_sm > 0) occ = std::min(occ, ctx->opt_ctas_per_sm);
    *columns = (int64_t)occ * ctx->sm_count * block * cpt;
    return 0;
}

void drop_stream_set(mcp_ctx* m) {
    if (!m->stream_set) return;
    for (auto& b : m->stream_set->blocks) {
        destroy_alignment_one(m, b.aln);
        for (cudaEvent_t ev : b.ev)
            if (ev) cudaEventDestroy(ev);
    }
    free_dev(m->stream_set->d_ctl);
    free_pin(m->stream_set->h_epoch);
    free_pin(m->stream_set->h_err);
    m->stream_set.reset();
}

// (Re)creates the block alignments of member `m` for the host alignment described by the arguments.
// Transfer units of the fused mode, in sites relative to the range start: small first units so that the first
// tiles can start after a few hundred microseconds, doubling up to a cap; every boundary is a multiple of the
// ready-flag granularity.
void plan_stream_chunks(long long n, std::vector<std::pair<long long, long long>>& out) {
    out.clear();
    const long long g = 1LL << STREAM_GROUP_SHIFT;
    const long long cap = std::min<long long>(65536, std::max<long long>(16384, ((n / 48 + g - 1) / g) * g));
    long long at = 0, sz = g;
    while (at < n) {
        const long long end = std::min(n, at + sz);
        out.push_back({at, end});
        at = end;
        sz = std::min(cap, sz * 2);
    }
}

int ensure_stream_set(mcp_ctx* m, const unsigned char* codes, int K, long long S, const int32_t* leaf_nums, int n_leaves,
                      long long lo, long long hi, int NN, int R, int want_grad, bool fused) {
    StreamSet* ss = m->stream_set.get();
    if (ss && ss->K == K && ss->S == S && ss->n_leaves == n_leaves && ss->R == R && ss->NN == NN && ss->want_grad == want_grad &&
        ss->fused == fused && std::memcmp(ss->leaf_nums.data(), leaf_nums, sizeof(int32_t) * n_leaves) == 0)
        return 0;
    drop_stream_set(m);
    std::unique_ptr<StreamSet> ns(new StreamSet());
    ns->K = K; ns->S = S; ns->n_leaves = n_leaves; ns->R = R; ns->NN = NN; ns->want_grad = want_grad;
    ns->fused = fused;
    ns->leaf_nums.assign(leaf_nums, leaf_nums + n_leaves);
    int64_t wave = 0;
    int e;
    if ((e = wave_columns_impl(m, K, NN, want_grad, &wave))) return e;
    std::vector<std::pair<long long, long long>> bounds;
    if (fused) {
        bounds.push_back({lo, hi});
        plan_stream_chunks(hi - lo, ns->chunks);
        const size_t n_groups = (size_t)((hi - lo + (1LL << STREAM_GROUP_SHIFT) - 1) >> STREAM_GROUP_SHIFT);
        if ((e = ensure_dev(m, ns->d_ctl, sizeof(unsigned) * (STREAM_CTL_WORDS + n_groups)))) return e;
        if ((e = ensure_pin(m, ns->h_epoch, sizeof(unsigned) * n_groups))) { free_dev(ns->d_ctl); return e; }
        if ((e = ensure_pin(m, ns->h_err, 64))) { free_dev(ns->d_ctl); free_pin(ns->h_epoch); return e; }
        cudaMemsetAsync(ns->d_ctl.p, 0, sizeof(unsigned) * (STREAM_CTL_WORDS + n_groups), m->stream);
    } else {
        plan_stream_blocks(lo, hi, std::max<long long>(1, wave / std::max(R, 1)), bounds);
    }
    if (bounds.empty()) bounds.push_back({lo, lo});   // an empty range still yields a (zero) result vector
    for (auto& b : bounds) {
        mcp_alignment* al = nullptr;
        if ((e = make_alignment(m, codes + b.first, (size_t)S, K, b.second - b.first, leaf_nums, n_leaves, &al))) {
            for (auto& bb : ns->blocks) destroy_alignment_one(m, bb.aln);
            free_dev(ns->d_ctl);
            free_pin(ns->h_epoch);
            free_pin(ns->h_err);
            return e;
        }
        StreamSet::Block blk;
        blk.aln = al;
        blk.lo = b.first;
        blk.hi = b.second;
        for (cudaEvent_t& ev : blk.ev) cudaEventCreate(&ev);
        ns->blocks.push_back(blk);
    }
    m->stream_set = std::move(ns);
    return 0;
}

int eval_streamed(mcp_ctx* ctx, const unsigned char* codes, int K, long long S, const int32_t* leaf_nums, int n_leaves,
                  int NN, const int32_t* po, const int32_t* pa, const double* blv, const double* U, const double* D,
                  const double* Uinv, double mu, const double* rates, int R, const double* pi, int want_grad,
                  double* ll_out, double* grad_out) {
    if (!ctx) return fail(nullptr, MCP_ERR_ARG, "null context");
    if (!codes || !leaf_nums || !po || !pa || !blv || !U || !D || !Uinv || !rates || !pi)
        return fail(ctx, MCP_ERR_ARG, "mcp_eval_streamed: null argument");
    if (K < 1 || K > 254 || S < 0 || n_leaves < 1 || NN < 2) return fail(ctx, MCP_ERR_ARG, "mcp_eval_streamed: bad K/S/n_leaves/NN");
    const bool multi = !ctx->members.empty();
    std::vector<mcp_ctx*> single{ctx};
    const std::vector<mcp_ctx*>& mem = multi ? ctx->members : single;
    const int G = (int)mem.size();
    const long long len = NN;
    int e;
    std::vector<double*> part;
    if (multi) {
        if ((e = group_targets(ctx, len, part))) return e;
    } else {
        CUDA_TRY(ctx, cudaSetDevice(ctx->device));
        if ((e = ensure_dev(ctx, ctx->d_out, sizeof(double) * len))) return e;
        if ((e = ensure_pin(ctx, ctx->h_out, sizeof(double) * len))) return e;
        part.assign(1, (double*)ctx->d_out.p);
    }
    auto run_member = [&](int g) -> int {
        mcp_ctx* m = mem[g];
        long long lo, hi;
        shard_range(S, G, g, &lo, &hi);
        int er;
        // One launch over the whole range with per-tile ready flags whenever the walk kernel that will run has
        // them (templated K, shared-memory accumulator, an input beyond the small-tree kernel's reach);
        // otherwise -- and with MCPHYLO_B200_STREAM_BLOCKS set, for comparison -- one launch per site block.
        const bool fused = k_templated(K) && !std::getenv("MCPHYLO_B200_STREAM_BLOCKS") && m->opt_levels != 1 &&
                           (hi - lo) * (long long)R > 2LL * 32 * 4 * m->sm_count &&
                           !(want_grad && walk_acc_global(NN, m->opt_acc_mode));
        if ((er = ensure_stream_set(m, codes, K, S, leaf_nums, n_leaves, lo, hi, NN, R, want_grad, fused))) return er;
        StreamSet& ss = *m->stream_set;
        const int B = (int)ss.blocks.size();
        if (cudaSetDevice(m->device) != cudaSuccess) return fail(m, MCP_ERR_CUDA, "cudaSetDevice failed");
        if ((er = ensure_dev(m, m->d_part, sizeof(double) * len * B))) return er;
        if (fused) {
            StreamSet::Block& blk = ss.blocks[0];
            mcp_alignment* al = blk.aln;
            if (++ss.epoch == 0) ss.epoch = 1;
            const size_t n_groups = (size_t)((hi - lo + (1LL << STREAM_GROUP_SHIFT) - 1) >> STREAM_GROUP_SHIFT);
            unsigned* he = (unsigned*)ss.h_epoch.p;
            for (size_t i = 0; i < n_groups; ++i) he[i] = ss.epoch;
            unsigned* ctl = (unsigned*)ss.d_ctl.p;
            if (al->read_since_upload) {   // the previous evaluation of this buffer must have finished reading it
                if (cudaStreamWaitEvent(m->copy_stream, al->streamed ? al->ev_read_done : m->ev_walk_done, 0) != cudaSuccess)
                    return fail(m, MCP_ERR_CUDA, "cudaStreamWaitEvent failed");
                al->read_since_upload = false;
            }
            al->streamed = true;
            al->upload_pending = false;
            m->pending_async = true;
            cudaEventRecord(blk.ev[0], m->copy_stream);
            auto send = [&](size_t c) -> int {
                const long long c0 = ss.chunks[c].first, c1 = ss.chunks[c].second;
                if (cudaMemcpy2DAsync(al->d_codes + c0, (size_t)al->stride, codes + lo + c0, (size_t)S, (size_t)(c1 - c0),
                                      (size_t)al->n_leaves, cudaMemcpyHostToDevice, m->copy_stream) != cudaSuccess)
                    return fail(m, MCP_ERR_CUDA, "alignment transfer failed: %s", cudaGetErrorString(cudaGetLastError()));
                const size_t g0 = (size_t)(c0 >> STREAM_GROUP_SHIFT), g1 = (size_t)((c1 + (1LL << STREAM_GROUP_SHIFT) - 1) >> STREAM_GROUP_SHIFT);
                // same stream, hence after the sites themselves: mark their groups as landed
                if (cudaMemcpyAsync(ctl + STREAM_CTL_WORDS + g0, he + g0, sizeof(unsigned) * (g1 - g0), cudaMemcpyHostToDevice,
                                    m->copy_stream) != cudaSuccess)
                    return fail(m, MCP_ERR_CUDA, "ready-flag transfer failed: %s", cudaGetErrorString(cudaGetLastError()));
                return 0;
            };
            // Every unit is enqueued BEFORE the walk is launched (a few microseconds of host time per unit, during
            // which the first units are already crossing PCIe): a launch that blocks the host until the kernel
            // has finished -- profilers, compute-sanitizer, CUDA_LAUNCH_BLOCKING -- then still finds its data
            // on the way instead of spinning until the time-out.
            const size_t n_first = ss.chunks.size();
            for (size_t c = 0; c < n_first; ++c)
                if ((er = send(c))) return er;
            StreamFlags sf{ctl + STREAM_CTL_WORDS, ctl, ctl + 1, ss.epoch, STREAM_GROUP_SHIFT};
            const mcp_alignment* alp = al;
            BatchArgs ab{1, &alp, &NN, &po, &pa, &blv, &U, &D, &Uinv, &mu, &rates, R, &pi, want_grad};
            cudaEventRecord(blk.ev[2], m->stream);
            m->sf = &sf;
            m->ev_walk_begin = blk.ev[3];
            m->ev_walk_end = blk.ev[4];
            er = eval_impl(m, ab, part[g], nullptr, nullptr);
            m->sf = nullptr;
            m->ev_walk_begin = m->ev_walk_end = nullptr;
            // whatever happened, every unit must be sent: a launched walk waits for all of them
            int er2 = 0;
            for (size_t c = n_first; c < ss.chunks.size() && !er2; ++c) er2 = send(c);
            cudaEventRecord(blk.ev[1], m->copy_stream);
            if (er || er2) return er ? er : er2;
            if (cudaMemcpyAsync(ss.h_err.p, ctl + 1, sizeof(unsigned), cudaMemcpyDeviceToHost, m->stream) != cudaSuccess)
                return fail(m, MCP_ERR_CUDA, "cudaMemcpyAsync failed");
            return 0;
        }
        // Issue order matters: host-to-device transfers are served first-in first-out, and every
        // evaluation starts with a small parameter upload of its own.  Transfer b, evaluation b,
        // transfer b+1, ...: the parameters of evaluation b queue right behind the block they need
        // anyway, and transfer b+1 then runs under the kernels of evaluation b.
        // experiments only (tools/e2e_probe.py): "noupload" re-uses the device copies of the previous call,
        // "serial" finishes every transfer before the first evaluation starts
        const char* probe = std::getenv("MCPHYLO_B200_STREAM_PROBE");
        const bool no_upload = probe && !std::strcmp(probe, "noupload") && ss.blocks[0].aln->streamed;
        const bool serial = probe && !std::strcmp(probe, "serial");
        if (serial) {
            for (int b = 0; b < B; ++b)
                if ((er = update_codes_one(m, ss.blocks[b].aln, codes + ss.blocks[b].lo, (size_t)S))) return er;
            cudaStreamSynchronize(m->copy_stream);
        }
        for (int b = 0; b < B; ++b) {
            StreamSet::Block& blk = ss.blocks[b];
            mcp_alignment* al = blk.aln;
            if (!no_upload && !serial) {
                cudaEventRecord(blk.ev[0], m->copy_stream);
                if ((er = update_codes_one(m, al, codes + blk.lo, (size_t)S))) return er;
                cudaEventRecord(blk.ev[1], m->copy_stream);
            }
            const mcp_alignment* alp = al;
            BatchArgs ab{1, &alp, &NN, &po, &pa, &blv, &U, &D, &Uinv, &mu, &rates, R, &pi, want_grad};
            double* dst = B == 1 ? part[g] : (double*)m->d_part.p + (long long)b * len;
            cudaEventRecord(blk.ev[2], m->stream);
            m->ev_walk_begin = blk.ev[3];
            m->ev_walk_end = blk.ev[4];
            er = eval_impl(m, ab, dst, nullptr, nullptr);
            m->ev_walk_begin = m->ev_walk_end = nullptr;
            if (er) return er;
        }
        if (B > 1) {
            sum_rows<<<(unsigned)((len + 255) / 256), 256, 0, m->stream>>>((const double*)m->d_part.p, len, B, len, part[g]);
            if (cudaGetLastError() != cudaSuccess) return fail(m, MCP_ERR_CUDA, "sum_rows launch failed");
        }
        return 0;
    };
    const double* h = nullptr;
    if (multi) {
        if ((e = for_each_member(ctx, run_member))) return e;
        if ((e = group_reduce(ctx, part, len, &h))) return e;
    } else {
        if ((e = run_member(0))) return e;
        if (ctx->rank_comm) {   // one rank of a multi-process group: S, codes describe this rank's shard
            const mcpnccl::Api& nc = mcpnccl::api();
            mcpnccl::result_t nr = nc.AllReduce(part[0], part[0], (size_t)len, mcpnccl::kFloat64, mcpnccl::kSum, ctx->rank_comm, ctx->stream);
            if (nr) return fail(ctx, MCP_ERR_CUDA, "ncclAllReduce failed: %s", nc.GetErrorString(nr));
        }
        CUDA_TRY(ctx, cudaMemcpyAsync(ctx->h_out.p, part[0], sizeof(double) * len, cudaMemcpyDeviceToHost, ctx->stream));
        CUDA_TRY(ctx, cudaStreamSynchronize(ctx->stream));
        for (bool& p : ctx->staged_pending) p = false;
        h = (const double*)ctx->h_out.p;
    }
    for (mcp_ctx* m : mem)
        if (m->stream_set && m->stream_set->fused) {
            cudaSetDevice(m->device);
            CUDA_TRY(ctx, cudaStreamSynchronize(m->stream));
            if (*(const unsigned*)m->stream_set->h_err.p != 0)
                return fail(ctx, MCP_ERR_CUDA, "mcp_eval_streamed: device %d gave up waiting for alignment data (transfer stalled)", m->device);
        }
    if (ll_out) *ll_out = h[0];
    if (want_grad && grad_out) std::memcpy(grad_out, h + 1, sizeof(double) * (NN - 1));
    return 0;
}

int create_single(mcp_ctx** out, int device) {
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
    for (int i = 0; e == cudaSuccess && i < MCP_STAGE_SLOTS; ++i) e = cudaEventCreateWithFlags(&ctx->ev_staged[i], cudaEventDisableTiming);
    if (e == cudaSuccess) e = cudaEventCreateWithFlags(&ctx->ev_walk_done, cudaEventDisableTiming);
    if (e == cudaSuccess) e = cudaEventCreateWithFlags(&ctx->ev_done, cudaEventDisableTiming);
    if (e == cudaSuccess) e = cudaStreamCreateWithFlags(&ctx->copy_stream, cudaStreamNonBlocking);
    if (e != cudaSuccess) {
        std::string msg = cudaGetErrorString(e);
        delete ctx;
        return fail(nullptr, MCP_ERR_CUDA, "stream/event creation failed: %s", msg.c_str());
    }
    ctx->stream = ctx->own_stream;
    {
        const char* v = std::getenv("MCPHYLO_B200_TIMING");
        if (v && v[0] == '0') ctx->opt_timing = 0;
    }
    *out = ctx;
    return 0;
}

void destroy_single(mcp_ctx* ctx) {
    cudaSetDevice(ctx->device);
    cudaStreamSynchronize(ctx->stream);
    if (ctx->copy_stream) cudaStreamSynchronize(ctx->copy_stream);
    drop_stream_set(ctx);
    if (ctx->rank_comm) mcpnccl::api().CommDestroy(ctx->rank_comm);
    for (DevBuf* b : {&ctx->d_dyn, &ctx->d_btab, &ctx->d_scratch, &ctx->d_rows, &ctx->d_rows_ll, &ctx->d_out, &ctx->d_counter, &ctx->d_part, &ctx->d_ticket, &ctx->d_mg})
        free_dev(*b);
    free_pin(ctx->h_out);
    free_pin(ctx->h_mg);
    for (int i = 0; i < MCP_STAGE_SLOTS; ++i) {
        free_pin(ctx->h_dyn[i]);
        free_pin(ctx->h_model[i]);
        if (ctx->ev_staged[i]) cudaEventDestroy(ctx->ev_staged[i]);
    }
    for (auto& pl : ctx->plans) {
        free_dev(pl->d_topo);
        free_pin(pl->h_topo);
    }
    for (auto& ev : ctx->ev)
        if (ev) cudaEventDestroy(ev);
    if (ctx->ev_walk_done) cudaEventDestroy(ctx->ev_walk_done);
    if (ctx->ev_done) cudaEventDestroy(ctx->ev_done);
    if (ctx->ev_t0) cudaEventDestroy(ctx->ev_t0);
    if (ctx->ev_t1) cudaEventDestroy(ctx->ev_t1);
    if (ctx->copy_stream) cudaStreamDestroy(ctx->copy_stream);
    if (ctx->own_stream) cudaStreamDestroy(ctx->own_stream);
    {   // this context may be remembered as the last user of a unit's constant model slots
        std::lock_guard<std::mutex> lock(g_slots.mu);
        for (auto& en : g_slots.entries)
            if (en.last == ctx) en.last = nullptr;
    }
    delete ctx;
}

template <class F>
int for_members_or_self(mcp_ctx* ctx, F f) {
    if (ctx->members.empty()) return f(ctx);
    for (mcp_ctx* m : ctx->members) {
        int e = f(m);
        if (e) { ctx->error = m->error; return e; }
    }
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
    return create_single(out, device);
}

int mcp_create_multi(mcp_ctx** out, int n_dev, const int* dev_ids, int reduce_mode) {
    if (!out) return fail(nullptr, MCP_ERR_ARG, "mcp_create_multi: null output pointer");
    *out = nullptr;
    if (n_dev < 1 || n_dev > 64 || !dev_ids) return fail(nullptr, MCP_ERR_ARG, "mcp_create_multi: need 1..64 device ids");
    if (reduce_mode < MCP_REDUCE_AUTO || reduce_mode > MCP_REDUCE_HOST) return fail(nullptr, MCP_ERR_ARG, "mcp_create_multi: unknown reduce mode %d", reduce_mode);
    bool distinct = true;
    for (int i = 0; i < n_dev; ++i)
        for (int j = 0; j < i; ++j) distinct = distinct && dev_ids[i] != dev_ids[j];
    std::unique_ptr<mcp_ctx> grp(new mcp_ctx());
    auto cleanup = [&]() {
        for (mcp_ctx* m : grp->members) destroy_single(m);
        grp->members.clear();
    };
    for (int i = 0; i < n_dev; ++i) {
        mcp_ctx* m = nullptr;
        int e = create_single(&m, dev_ids[i]);
        if (e) { cleanup(); return e; }
        grp->members.push_back(m);
    }
    grp->device = dev_ids[0];
    grp->sm_count = grp->members[0]->sm_count;
    // peer access towards device 0 (PEER reduction; harmless otherwise)
    bool peer_ok = true;
    for (int i = 1; i < n_dev; ++i) {
        if (dev_ids[i] == dev_ids[0]) continue;
        int can = 0;
        cudaDeviceCanAccessPeer(&can, dev_ids[i], dev_ids[0]);
        if (!can) { peer_ok = false; continue; }
        cudaSetDevice(dev_ids[i]);
        cudaError_t pe = cudaDeviceEnablePeerAccess(dev_ids[0], 0);
        if (pe != cudaSuccess && pe != cudaErrorPeerAccessAlreadyEnabled) peer_ok = false;
        cudaGetLastError();
    }
    std::string why;
    int mode = reduce_mode;
    if (mode == MCP_REDUCE_AUTO) {
        // the [logL, gradient] all-reduce over NVLink is NCCL's whenever it can be: distinct devices and a
        // loadable library; otherwise peer stores into device 0, otherwise a sum on the host
        mode = (n_dev > 1 && distinct && mcpnccl::api(&why).ok()) ? MCP_REDUCE_NCCL : peer_ok ? MCP_REDUCE_PEER : MCP_REDUCE_HOST;
    }
    if (mode == MCP_REDUCE_NCCL) {
        if (!distinct) { cleanup(); return fail(nullptr, MCP_ERR_ARG, "mcp_create_multi: the NCCL reduction needs distinct devices"); }
        const mcpnccl::Api& nc = mcpnccl::api(&why);
        if (!nc.ok()) { cleanup(); return fail(nullptr, MCP_ERR_UNSUPPORTED, "mcp_create_multi: NCCL is not available (%s)", why.c_str()); }
        grp->comms.assign(n_dev, nullptr);
        mcpnccl::result_t nr = nc.CommInitAll(grp->comms.data(), n_dev, dev_ids);
        if (nr) {
            grp->comms.clear();
            cleanup();
            return fail(nullptr, MCP_ERR_CUDA, "ncclCommInitAll failed: %s", nc.GetErrorString(nr));
        }
    } else if (mode == MCP_REDUCE_PEER) {
        if (!peer_ok) { cleanup(); return fail(nullptr, MCP_ERR_UNSUPPORTED, "mcp_create_multi: peer access to device %d is not available", dev_ids[0]); }
        grp->ev_member.assign(n_dev, nullptr);
        for (int i = 1; i < n_dev; ++i) {
            cudaSetDevice(dev_ids[i]);
            if (cudaEventCreateWithFlags(&grp->ev_member[i], cudaEventDisableTiming) != cudaSuccess) {
                cleanup();
                return fail(nullptr, MCP_ERR_CUDA, "event creation failed");
            }
        }
    } else {
        grp->h_member_out.assign(n_dev, PinBuf());
    }
    grp->reduce_mode = mode;
    if (n_dev > 1 && !std::getenv("MCPHYLO_B200_SERIAL_GROUP")) grp->pool.reset(new MemberPool(n_dev));
    *out = grp.release();
    return 0;
}

int mcp_nccl_unique_id(void* id128) {
    if (!id128) return fail(nullptr, MCP_ERR_ARG, "mcp_nccl_unique_id: null argument");
    std::string why;
    const mcpnccl::Api& nc = mcpnccl::api(&why);
    if (!nc.ok()) return fail(nullptr, MCP_ERR_UNSUPPORTED, "NCCL is not available (%s)", why.c_str());
    mcpnccl::unique_id id;
    mcpnccl::result_t nr = nc.GetUniqueId(&id);
    if (nr) return fail(nullptr, MCP_ERR_CUDA, "ncclGetUniqueId failed: %s", nc.GetErrorString(nr));
    std::memcpy(id128, &id, sizeof id);
    return 0;
}

int mcp_create_rank(mcp_ctx** out, int device, int n_ranks, int rank, const void* id128) {
    if (!out) return fail(nullptr, MCP_ERR_ARG, "mcp_create_rank: null output pointer");
    *out = nullptr;
    if (n_ranks < 1 || rank < 0 || rank >= n_ranks || !id128) return fail(nullptr, MCP_ERR_ARG, "mcp_create_rank: bad rank / id");
    std::string why;
    const mcpnccl::Api& nc = mcpnccl::api(&why);
    if (!nc.ok()) return fail(nullptr, MCP_ERR_UNSUPPORTED, "NCCL is not available (%s)", why.c_str());
    mcp_ctx* ctx = nullptr;
    int e = create_single(&ctx, device);
    if (e) return e;
    mcpnccl::unique_id id;
    std::memcpy(&id, id128, sizeof id);
    mcpnccl::result_t nr = nc.CommInitRank(&ctx->rank_comm, n_ranks, id, rank);
    if (nr) {
        ctx->rank_comm = nullptr;
        destroy_single(ctx);
        return fail(nullptr, MCP_ERR_CUDA, "ncclCommInitRank failed: %s", nc.GetErrorString(nr));
    }
    ctx->n_ranks = n_ranks;
    ctx->rank = rank;
    *out = ctx;
    return 0;
}

int mcp_device_count(const mcp_ctx* ctx) { return !ctx ? 0 : ctx->members.empty() ? 1 : (int)ctx->members.size(); }

int mcp_reduce_mode(const mcp_ctx* ctx) {
    if (!ctx) return MCP_REDUCE_AUTO;
    if (ctx->rank_comm) return MCP_REDUCE_NCCL;
    return ctx->members.empty() ? MCP_REDUCE_AUTO : ctx->reduce_mode;
}

int mcp_shard_bounds(int64_t S, int n_shards, int shard, int64_t* lo, int64_t* hi) {
    if (S < 0 || n_shards < 1 || shard < 0 || shard >= n_shards || !lo || !hi) return fail(nullptr, MCP_ERR_ARG, "mcp_shard_bounds: bad argument");
    long long l, h;
    shard_range(S, n_shards, shard, &l, &h);
    *lo = l;
    *hi = h;
    return 0;
}

int mcp_destroy(mcp_ctx* ctx) {
    if (!ctx) return 0;
    if (!ctx->members.empty()) {
        for (mcp_ctx* m : ctx->members) {
            cudaSetDevice(m->device);
            cudaStreamSynchronize(m->stream);
        }
        if (!ctx->comms.empty()) {
            const mcpnccl::Api& nc = mcpnccl::api();
            for (auto c : ctx->comms)
                if (c) nc.CommDestroy(c);
        }
        for (auto ev : ctx->ev_member)
            if (ev) cudaEventDestroy(ev);
        for (auto& b : ctx->h_member_out) free_pin(b);
        if (ctx->d_gather.p) {
            cudaSetDevice(ctx->members[0]->device);
            free_dev(ctx->d_gather);
        }
        ctx->pool.reset();
        for (mcp_ctx* m : ctx->members) destroy_single(m);
        delete ctx;
        return 0;
    }
    destroy_single(ctx);
    return 0;
}

int mcp_set_stream(mcp_ctx* ctx, void* cuda_stream) {
    if (!ctx) return fail(nullptr, MCP_ERR_ARG, "null context");
    if (!ctx->members.empty()) return fail(ctx, MCP_ERR_UNSUPPORTED, "mcp_set_stream: a multi-device context runs on its own streams");
    if (ctx->pending_async) {
        CUDA_TRY(ctx, cudaStreamSynchronize(ctx->stream));
        ctx->pending_async = false;
    }
    ctx->stream = (cudaStream_t)cuda_stream;   // NULL is the CUDA default stream, a valid choice
    return 0;
}

int mcp_synchronize(mcp_ctx* ctx) {
    if (!ctx) return fail(nullptr, MCP_ERR_ARG, "null context");
    return for_members_or_self(ctx, [](mcp_ctx* m) -> int {
        CUDA_TRY(m, cudaSetDevice(m->device));
        CUDA_TRY(m, cudaStreamSynchronize(m->stream));
        CUDA_TRY(m, cudaStreamSynchronize(m->copy_stream));
        m->pending_async = false;
        return 0;
    });
}

int mcp_use_own_stream(mcp_ctx* ctx) {
    if (!ctx) return fail(nullptr, MCP_ERR_ARG, "null context");
    if (!ctx->members.empty()) return 0;
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
    return for_members_or_self(ctx, [&](mcp_ctx* m) -> int {
        m->opt_block = block;
        m->opt_ctas_per_sm = ctas_per_sm;
        invalidate_plans(m);
        return 0;
    });
}

int mcp_set_scratch_mode(mcp_ctx* ctx, int mode) {
    if (!ctx) return fail(nullptr, MCP_ERR_ARG, "null context");
    if (mode < -1 || mode > 1) return fail(ctx, MCP_ERR_ARG, "scratch mode must be -1 (automatic), 0 (HBM) or 1 (shared memory when it fits)");
    return for_members_or_self(ctx, [&](mcp_ctx* m) -> int {
        m->opt_smem_scratch = mode;
        invalidate_plans(m);
        return 0;
    });
}

int mcp_set_accumulator_mode(mcp_ctx* ctx, int mode) {
    if (!ctx) return fail(nullptr, MCP_ERR_ARG, "null context");
    if (mode < -1 || mode > 1) return fail(ctx, MCP_ERR_ARG, "accumulator mode must be -1 (automatic), 0 (shared memory) or 1 (global memory)");
    return for_members_or_self(ctx, [&](mcp_ctx* m) -> int {
        m->opt_acc_mode = mode;
        invalidate_plans(m);
        return 0;
    });
}

int mcp_set_ring_mode(mcp_ctx* ctx, int mode) {
    if (!ctx) return fail(nullptr, MCP_ERR_ARG, "null context");
    if (mode < -1 || mode > 1) return fail(ctx, MCP_ERR_ARG, "ring mode must be -1 (automatic), 0 (off) or 1 (on where supported)");
    return for_members_or_self(ctx, [&](mcp_ctx* m) -> int {
        if (m->opt_ring != mode) invalidate_plans(m);
        m->opt_ring = mode;
        return 0;
    });
}

int mcp_set_large_alphabet_mode(mcp_ctx* ctx, int mode) {
    if (!ctx) return fail(nullptr, MCP_ERR_ARG, "null context");
    if (mode < -1 || mode > 1) return fail(ctx, MCP_ERR_ARG, "large-alphabet mode must be -1 (automatic), 0 (runtime-K kernel) or 1 (tensor-core kernel)");
    return for_members_or_self(ctx, [&](mcp_ctx* m) -> int {
        m->opt_mma = mode;
        invalidate_plans(m);
        return 0;
    });
}

int mcp_set_cherry_mode(mcp_ctx* ctx, int mode) {
    if (!ctx) return fail(nullptr, MCP_ERR_ARG, "null context");
    if (mode < -1 || mode > 1) return fail(ctx, MCP_ERR_ARG, "cherry mode must be -1 (automatic), 0 (stored) or 1 (recomputed)");
    return for_members_or_self(ctx, [&](mcp_ctx* m) -> int {
        m->opt_cherry = mode;
        invalidate_plans(m);
        return 0;
    });
}

int mcp_set_tile_order(mcp_ctx* ctx, int mode) {
    if (!ctx) return fail(nullptr, MCP_ERR_ARG, "null context");
    if (mode < -1 || mode > 1) return fail(ctx, MCP_ERR_ARG, "tile order must be -1 (automatic), 0 (static ranges) or 1 (atomic tickets)");
    return for_members_or_self(ctx, [&](mcp_ctx* m) -> int {
        m->opt_dynamic = mode;
        return 0;
    });
}

int mcp_set_level_mode(mcp_ctx* ctx, int mode) {
    if (!ctx) return fail(nullptr, MCP_ERR_ARG, "null context");
    if (mode < -1 || mode > 1) return fail(ctx, MCP_ERR_ARG, "level mode must be -1 (automatic), 0 (off) or 1 (whenever the tree fits)");
    return for_members_or_self(ctx, [&](mcp_ctx* m) -> int {
        m->opt_levels = mode;
        invalidate_plans(m);
        return 0;
    });
}

int mcp_set_timing(mcp_ctx* ctx, int on) {
    if (!ctx) return fail(nullptr, MCP_ERR_ARG, "null context");
    return for_members_or_self(ctx, [&](mcp_ctx* m) { m->opt_timing = on ? 1 : 0; return 0; });
}

int mcp_set_columns_per_thread(mcp_ctx* ctx, int cpt) {
    if (!ctx) return fail(nullptr, MCP_ERR_ARG, "null context");
    if (cpt != 0 && cpt != 1 && cpt != 2) return fail(ctx, MCP_ERR_ARG, "columns per thread must be 0 (automatic), 1 or 2");
    return for_members_or_self(ctx, [&](mcp_ctx* m) -> int {
        m->opt_cpt = cpt;
        invalidate_plans(m);
        return 0;
    });
}

static int alignment_from_codes_any(mcp_ctx* ctx, const uint8_t* codes, int K, int64_t S, const int32_t* leaf_nums,
                                    int n_leaves, mcp_alignment** out) {
    if (ctx->members.empty()) return make_alignment(ctx, codes, (size_t)S, K, S, leaf_nums, n_leaves, out);
    const int G = (int)ctx->members.size();
    std::unique_ptr<mcp_alignment> al(new mcp_alignment());
    al->K = K;
    al->S = S;
    al->n_leaves = n_leaves;
    al->owner = ctx;
    al->leaf_nums.assign(leaf_nums, leaf_nums + n_leaves);
    al->id = ctx->next_aln_id++;
    for (int g = 0; g < G; ++g) {
        long long lo, hi;
        shard_range(S, G, g, &lo, &hi);
        mcp_alignment* sh = nullptr;
        int e = make_alignment(ctx->members[g], codes + lo, (size_t)S, K, hi - lo, leaf_nums, n_leaves, &sh);
        if (e) {
            ctx->error = ctx->members[g]->error;
            for (int j = 0; j < g; ++j) destroy_alignment_one(ctx->members[j], al->shards[j]);
            return e;
        }
        al->shards.push_back(sh);
        al->shard_lo.push_back(lo);
    }
    *out = al.release();
    return 0;
}

int mcp_alignment_from_codes(mcp_ctx* ctx, const uint8_t* codes, int K, int64_t S, const int32_t* leaf_nums,
                             int n_leaves, mcp_alignment** out) {
    if (!ctx) return fail(nullptr, MCP_ERR_ARG, "null context");
    if (!out || !codes || !leaf_nums) return fail(ctx, MCP_ERR_ARG, "mcp_alignment_from_codes: null argument");
    if (K < 1 || K > 254 || S < 0 || n_leaves < 1) return fail(ctx, MCP_ERR_ARG, "mcp_alignment_from_codes: bad K/S/n_leaves");
    return alignment_from_codes_any(ctx, codes, K, S, leaf_nums, n_leaves, out);
}

int mcp_alignment_from_dense(mcp_ctx* ctx, const double* x, int K, int64_t S, int NN, const int32_t* leaf_nums,
                             int n_leaves, mcp_alignment** out) {
    if (!ctx) return fail(nullptr, MCP_ERR_ARG, "null context");
    if (!out || !x || !leaf_nums) return fail(ctx, MCP_ERR_ARG, "mcp_alignment_from_dense: null argument");
    if (K < 1 || K > 254 || S < 0 || n_leaves < 1 || NN < n_leaves) return fail(ctx, MCP_ERR_ARG, "mcp_alignment_from_dense: bad sizes");
    std::vector<unsigned char> codes;
    int e = dense_to_codes(ctx, x, K, S, NN, leaf_nums, n_leaves, codes);
    if (e) return e;
    static const unsigned char none = 0;
    return alignment_from_codes_any(ctx, codes.empty() ? &none : codes.data(), K, S, leaf_nums, n_leaves, out);
}

int mcp_alignment_update_codes(mcp_ctx* ctx, mcp_alignment* aln, const uint8_t* codes) {
    if (!ctx) return fail(nullptr, MCP_ERR_ARG, "null context");
    if (!aln || !codes) return fail(ctx, MCP_ERR_ARG, "mcp_alignment_update_codes: null argument");
    if (aln->owner != ctx) return fail(ctx, MCP_ERR_ARG, "mcp_alignment_update_codes: the alignment belongs to another context");
    if (ctx->members.empty()) return update_codes_one(ctx, aln, codes, (size_t)aln->S);
    for (size_t g = 0; g < aln->shards.size(); ++g) {
        int e = update_codes_one(ctx->members[g], aln->shards[g], codes + aln->shard_lo[g], (size_t)aln->S);
        if (e) { ctx->error = ctx->members[g]->error; return e; }
    }
    return 0;
}

int mcp_alignment_destroy(mcp_ctx* ctx, mcp_alignment* aln) {
    if (!aln) return 0;
    if (!aln->shards.empty()) {
        for (size_t g = 0; g < aln->shards.size(); ++g)
            destroy_alignment_one(ctx && g < ctx->members.size() ? ctx->members[g] : nullptr, aln->shards[g]);
        delete aln;
        return 0;
    }
    destroy_alignment_one(ctx, aln);
    return 0;
}

int mcp_eval(mcp_ctx* ctx, const mcp_alignment* aln, int NN, const int32_t* postorder_num, const int32_t* parent_num,
             const double* blv, const double* U, const double* D, const double* Uinv, double mu, const double* rates,
             int R, const double* pi, int want_grad, double* ll_out, double* grad_out) {
    BatchArgs a{1, &aln, &NN, &postorder_num, &parent_num, &blv, &U, &D, &Uinv, &mu, &rates, R, &pi, want_grad};
    double* g = grad_out;
    return eval_any(ctx, a, ll_out, &g);
}

int mcp_eval_posterior(mcp_ctx* ctx, const mcp_alignment* aln, int NN, const int32_t* postorder_num,
                       const int32_t* parent_num, const double* blv, const double* U, const double* D, const double* Uinv,
                       double mu, const double* rates, int R, const double* pi, int prior_kind, const double* prior_params,
                       double* lp_out, double* grad_out) {
    BatchArgs a{1, &aln, &NN, &postorder_num, &parent_num, &blv, &U, &D, &Uinv, &mu, &rates, R, &pi, grad_out ? 1 : 0};
    a.prior_kind = prior_kind;
    a.prior_params = prior_params;
    double* g = grad_out;
    return eval_any(ctx, a, lp_out, &g);
}

int mcp_eval_rate_gradient(mcp_ctx* ctx, const mcp_alignment* aln, int NN, const int32_t* postorder_num,
                           const int32_t* parent_num, const double* blv, const double* U, const double* D, const double* Uinv,
                           double mu, const double* rates, int R, const double* pi, double* ll_out, double* grad_out,
                           double* rate_grad_out) {
    if (!ctx) return fail(nullptr, MCP_ERR_ARG, "null context");
    if (!rates || !blv || !rate_grad_out || R < 1 || NN < 2) return fail(ctx, MCP_ERR_ARG, "mcp_eval_rate_gradient: bad argument");
    for (int r = 0; r < R; ++r)
        if (!(rates[r] > 0.0)) return fail(ctx, MCP_ERR_ARG, "mcp_eval_rate_gradient: rate category %d is not positive", r);
    // Rate categories are not mixed (logL = sum_r logL_r, VectorizedFunctions.jl:76-87) and category r sees every
    // branch as t_b * rate_r, so  d logL / d rate_r = (1 / rate_r) * sum_b t_b * d logL_r / d t_b:  one evaluation
    // per category with R = 1 -- the same columns as one R-category evaluation, on one cached plan -- gives
    // logL, the branch gradient and the rate gradient together.
    std::vector<double> g((size_t)NN - 1), gsum((size_t)NN - 1, 0.0);
    double ll_total = 0.0;
    for (int r = 0; r < R; ++r) {
        const double* rate_r = rates + r;
        double ll_r = 0.0;
        double* gp = g.data();
        BatchArgs a{1, &aln, &NN, &postorder_num, &parent_num, &blv, &U, &D, &Uinv, &mu, &rate_r, 1, &pi, 1};
        int e = eval_any(ctx, a, &ll_r, &gp);
        if (e) return e;
        ll_total += ll_r;
        double dot = 0.0;
        for (int b = 0; b < NN - 1; ++b) {
            gsum[b] += g[b];
            dot += blv[b] * g[b];
        }
        rate_grad_out[r] = dot / rates[r];
    }
    if (ll_out) *ll_out = ll_total;
    if (grad_out) std::memcpy(grad_out, gsum.data(), sizeof(double) * (NN - 1));
    return 0;
}

// --------------------------------------------------------------------------------------------
// gradient with respect to the parameters of the substitution model
// --------------------------------------------------------------------------------------------
namespace {

// Contraction of the moment matrices with d P_{b,r} / d theta.  With A = mu U diag(D) Uinv, x_i = mu D_i t_b rate_r,
//   P_{b,r} = exp(A t_b rate_r) = U diag(exp x) Uinv,
//   d P_{b,r} / d theta = U (Phi o (Uinv (dA/dtheta) U)) Uinv * t_b rate_r,     Phi_ij = (e^{x_i} - e^{x_j}) / (x_i - x_j), Phi_ii = e^{x_i}
// (Daleckii-Krein), hence   sum_{s,k} M[s][k] dP[s][k] = t rate sum_ij Phi_ij G_ij Mt_ij   with   Mt = U^T M Uinv^T,  G = Uinv dA U,
// and the sum over branches and categories needs one K x K matrix  H = sum_{b,r} t_b rate_r Phi^{b,r} o Mt^{b,r}  whatever the
// number of parameters:  d logL / d theta_p = sum_ij G^p_ij H_ij + sum_s (d pi_s / d theta_p) W[s].
// The same Mt gives the branch gradient (grad_check[b] = sum_r rate_r sum_i mu D_i e^{x_i} Mt_ii), returned for cross-checks,
// and the gradient with respect to the rate-category multipliers (rate_grad[r] = sum_b t_b sum_i mu D_i e^{x_i} Mt^{b,r}_ii).
void model_gradient_contract(int K, int R, int NB, const double* blv, const double* U, const double* D, const double* Uinv, double mu,
                             const double* rates, const double* M, const double* W, int n_par, const double* dA, const double* dpi,
                             double* par_grad, double* grad_check, double* rate_grad) {
    std::vector<double> H((size_t)K * K, 0.0), T1((size_t)K * K), Mt((size_t)K * K), x(K), ex(K);
    if (rate_grad)
        for (int r = 0; r < R; ++r) rate_grad[r] = 0.0;
    for (int b = 0; b < NB; ++b) {
        double gb = 0.0;
        for (int r = 0; r < R; ++r) {
            const double* Mb = M + ((size_t)b * R + r) * K * K;     // Mb[s * K + k]
            // T1[i][k] = sum_s U[s][i] Mb[s][k];   Mt[i][j] = sum_k T1[i][k] Uinv[j][k]     (col-major: U[s + K i], Uinv[j + K k])
            for (int i = 0; i < K; ++i)
                for (int k = 0; k < K; ++k) {
                    double acc = 0.0;
                    for (int s = 0; s < K; ++s) acc += U[s + (size_t)K * i] * Mb[(size_t)s * K + k];
                    T1[(size_t)i * K + k] = acc;
                }
            for (int i = 0; i < K; ++i)
                for (int j = 0; j < K; ++j) {
                    double acc = 0.0;
                    for (int k = 0; k < K; ++k) acc += T1[(size_t)i * K + k] * Uinv[j + (size_t)K * k];
                    Mt[(size_t)i * K + j] = acc;
                }
            const double tau = blv[b] * rates[r];
            for (int i = 0; i < K; ++i) {
                x[i] = mu * D[i] * tau;
                ex[i] = std::exp(x[i]);
                gb += rates[r] * mu * D[i] * ex[i] * Mt[(size_t)i * K + i];
                // category r sees the branch as t_b * rate_r: the same diagonal term, weighted by t_b instead of rate_r
                if (rate_grad) rate_grad[r] += blv[b] * mu * D[i] * ex[i] * Mt[(size_t)i * K + i];
            }
            for (int i = 0; i < K; ++i)
                for (int j = 0; j < K; ++j) {
                    const double d = x[i] - x[j];
                    // the larger exponent is the base, so the divided difference never overflows or cancels
                    const double phi = d == 0.0 ? ex[i] : d > 0.0 ? ex[i] * (-std::expm1(-d) / d) : ex[j] * (std::expm1(d) / d);
                    H[(size_t)i * K + j] += tau * phi * Mt[(size_t)i * K + j];
                }
        }
        if (grad_check) grad_check[b] = gb;
    }
    for (int p = 0; p < n_par; ++p) {
        const double* dAp = dA + (size_t)p * K * K;     // col-major K x K
        double g = 0.0;
        // G[i][j] = sum_{a,c} Uinv[i][a] dA[a][c] U[c][j]
        for (int i = 0; i < K; ++i)
            for (int c = 0; c < K; ++c) {
                double acc = 0.0;
                for (int a = 0; a < K; ++a) acc += Uinv[i + (size_t)K * a] * dAp[a + (size_t)K * c];
                T1[(size_t)i * K + c] = acc;
            }
        for (int i = 0; i < K; ++i)
            for (int j = 0; j < K; ++j) {
                double acc = 0.0;
                for (int c = 0; c < K; ++c) acc += T1[(size_t)i * K + c] * U[c + (size_t)K * j];
                g += acc * H[(size_t)i * K + j];
            }
        if (dpi && W)
            for (int s = 0; s < K; ++s) g += dpi[(size_t)p * K + s] * W[s];
        par_grad[p] = g;
    }
}

}  // namespace

int mcp_model_gradient_contract(int K, int R, int n_branches, const double* blv, const double* U, const double* D,
                                const double* Uinv, double mu, const double* rates, const double* moments,
                                const double* root_w, int n_par, const double* dA, const double* dpi, double* par_grad_out,
                                double* grad_check_out, double* rate_grad_out) {
    if (K < 1 || R < 1 || n_branches < 0 || n_par < 0 || !blv || !U || !D || !Uinv || !rates || !moments || (n_par > 0 && (!dA || !par_grad_out)))
        return fail(nullptr, MCP_ERR_ARG, "mcp_model_gradient_contract: bad argument");
    model_gradient_contract(K, R, n_branches, blv, U, D, Uinv, mu, rates, moments, root_w, n_par, dA, dpi, par_grad_out, grad_check_out,
                            rate_grad_out);
    return 0;
}

int mcp_eval_model_gradient(mcp_ctx* ctx, const mcp_alignment* aln, int NN, const int32_t* postorder_num,
                            const int32_t* parent_num, const double* blv, const double* U, const double* D, const double* Uinv,
                            double mu, const double* rates, int R, const double* pi, int n_par, const double* dA,
                            const double* dpi, double* ll_out, double* grad_out, double* par_grad_out, double* rate_grad_out,
                            double* moments_out) {
    if (!ctx) return fail(nullptr, MCP_ERR_ARG, "null context");
    if (!aln || NN < 2 || n_par < 0 || (n_par > 0 && (!dA || !par_grad_out)))
        return fail(ctx, MCP_ERR_ARG, "mcp_eval_model_gradient: bad argument");
    const int K = aln->K;
    const size_t n_m = (size_t)(NN - 1) * R * K * K;      // moments of the real branches; the root vector follows
    std::vector<double> msum(n_m + K, 0.0), grad((size_t)NN - 1, 0.0);
    double ll = 0.0;
    BatchArgs a{1, &aln, &NN, &postorder_num, &parent_num, &blv, &U, &D, &Uinv, &mu, &rates, R, &pi, 1};
    a.model_grad = 1;
    // rows of the device array beyond NN-1 belong to the root and to virtual branches (binarisation): dropped here
    auto take = [&](const mcp_ctx* m, double* dst, bool add) {
        const double* h = (const double*)m->h_mg.p;
        const size_t n_br = (size_t)m->last_plan->trees[0].n_br;
        for (size_t i = 0; i < n_m; ++i) dst[i] = add ? dst[i] + h[i] : h[i];
        for (int k = 0; k < K; ++k) dst[n_m + k] = add ? dst[n_m + k] + h[n_br * R * K * K + k] : h[n_br * R * K * K + k];
    };
    if (ctx->members.empty()) {
        double* gp = grad.data();
        int e = eval_impl(ctx, a, nullptr, &ll, &gp);
        if (e) return e;
        take(ctx, msum.data(), false);
    } else {
        // every device evaluates its site shard (one host thread per device); logL, the branch gradient and the moments
        // are sums over sites, added on the host in device order
        const int G = (int)ctx->members.size();
        int K0 = 0, e;
        if ((e = validate_batch(ctx, a, &K0))) return e;
        if ((int)aln->shards.size() != G) return fail(ctx, MCP_ERR_ARG, "the alignment was not created on this multi-device context");
        std::vector<double> lls(G, 0.0);
        std::vector<std::vector<double>> grads(G, std::vector<double>((size_t)NN - 1, 0.0));
        e = for_each_member(ctx, [&](int g) -> int {
            const mcp_alignment* sh = aln->shards[g];
            BatchArgs ag = a;
            ag.alns = &sh;
            double* gp = grads[g].data();
            return eval_impl(ctx->members[g], ag, nullptr, &lls[g], &gp);
        });
        if (e) return e;
        for (int g = 0; g < G; ++g) {
            ll += lls[g];
            for (int b = 0; b < NN - 1; ++b) grad[b] += grads[g][b];
            take(ctx->members[g], msum.data(), g > 0);
        }
    }
    if (ll_out) *ll_out = ll;
    if (grad_out) std::memcpy(grad_out, grad.data(), sizeof(double) * (NN - 1));
    if (moments_out) std::memcpy(moments_out, msum.data(), sizeof(double) * (n_m + K));
    if (n_par > 0 || rate_grad_out)
        model_gradient_contract(K, R, NN - 1, blv, U, D, Uinv, mu, rates, msum.data(), msum.data() + n_m, n_par, dA, dpi, par_grad_out, nullptr,
                                rate_grad_out);
    return 0;
}

int mcp_eval_device(mcp_ctx* ctx, const mcp_alignment* aln, int NN, const int32_t* postorder_num,
                    const int32_t* parent_num, const double* blv, const double* U, const double* D, const double* Uinv,
                    double mu, const double* rates, int R, const double* pi, int want_grad, double* d_out) {
    if (!ctx) return fail(nullptr, MCP_ERR_ARG, "null context");
    if (!d_out) return fail(ctx, MCP_ERR_ARG, "mcp_eval_device: null device output pointer");
    if (!ctx->members.empty()) return fail(ctx, MCP_ERR_UNSUPPORTED, "mcp_eval_device: not available on a multi-device context (use mcp_eval)");
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
    return eval_any(ctx, a, ll_out, grad_out);
}

int mcp_eval_streamed(mcp_ctx* ctx, const uint8_t* codes, int K, int64_t S, const int32_t* leaf_nums, int n_leaves,
                      int NN, const int32_t* postorder_num, const int32_t* parent_num, const double* blv, const double* U,
                      const double* D, const double* Uinv, double mu, const double* rates, int R, const double* pi,
                      int want_grad, double* ll_out, double* grad_out) {
    return eval_streamed(ctx, codes, K, S, leaf_nums, n_leaves, NN, postorder_num, parent_num, blv, U, D, Uinv, mu, rates, R,
                         pi, want_grad, ll_out, grad_out);
}

int mcp_stream_blocks(const mcp_ctx* ctx, int member, int64_t* lo_hi, int cap) {
    if (!ctx) return fail(nullptr, MCP_ERR_ARG, "null context");
    const mcp_ctx* m = ctx->members.empty() ? ctx : (member >= 0 && member < (int)ctx->members.size() ? ctx->members[member] : nullptr);
    if (!m || !m->stream_set) return 0;
    int n = 0;
    for (auto& b : m->stream_set->blocks) {
        if (lo_hi && n < cap) { lo_hi[2 * n] = b.lo; lo_hi[2 * n + 1] = b.hi; }
        ++n;
    }
    return n;
}

int mcp_stream_timeline(const mcp_ctx* ctx, int member, double* ms, int cap_blocks) {
    if (!ctx || !ms) return fail(nullptr, MCP_ERR_ARG, "mcp_stream_timeline: null argument");
    const mcp_ctx* m = ctx->members.empty() ? ctx : (member >= 0 && member < (int)ctx->members.size() ? ctx->members[member] : nullptr);
    if (!m || !m->stream_set || m->stream_set->blocks.empty()) return 0;
    cudaSetDevice(m->device);
    const auto& blocks = m->stream_set->blocks;
    cudaEvent_t origin = blocks[0].ev[0];
    int n = 0;
    for (const auto& b : blocks) {
        if (n >= cap_blocks) break;
        for (int i = 0; i < 5; ++i) {
            float t = 0.f;
            if (cudaEventElapsedTime(&t, origin, b.ev[i]) != cudaSuccess) { cudaGetLastError(); t = -1.f; }
            ms[5 * n + i] = t;
        }
        ++n;
    }
    return n;
}

int mcp_host_register(void* p, size_t bytes) {
    if (!p || !bytes) return fail(nullptr, MCP_ERR_ARG, "mcp_host_register: null argument");
    cudaError_t e = cudaHostRegister(p, bytes, cudaHostRegisterPortable);
    if (e != cudaSuccess) {
        cudaGetLastError();
        return fail(nullptr, MCP_ERR_CUDA, "cudaHostRegister failed: %s", cudaGetErrorString(e));
    }
    return 0;
}

int mcp_host_unregister(void* p) {
    if (!p) return 0;
    cudaError_t e = cudaHostUnregister(p);
    if (e != cudaSuccess) {
        cudaGetLastError();
        return fail(nullptr, MCP_ERR_CUDA, "cudaHostUnregister failed: %s", cudaGetErrorString(e));
    }
    return 0;
}

int mcp_wave_columns(mcp_ctx* ctx, int K, int n_nodes, int want_grad, int64_t* columns) {
    if (!ctx || !columns) return fail(ctx, MCP_ERR_ARG, "mcp_wave_columns: null argument");
    if (!k_supported(K)) return fail(ctx, MCP_ERR_UNSUPPORTED, "no kernel compiled for K = %d states", K);
    mcp_ctx* m = ctx->members.empty() ? ctx : ctx->members[0];
    int e = wave_columns_impl(m, K, n_nodes, want_grad, columns);
    if (e && m != ctx) ctx->error = m->error;
    return e;
}

static void read_stats(const mcp_ctx* ctx, mcp_stats* out) {
    *out = ctx->stats;
    // event times are read lazily: they exist once the stream has passed the last event
    float ms = 0.f;
    if (!ctx->opt_timing) return;
    if (ctx->ev[3] && cudaEventQuery(ctx->ev[3]) == cudaSuccess) {
        if (cudaEventElapsedTime(&ms, ctx->ev[1], ctx->ev[2]) == cudaSuccess) out->walk_ms = ms;
        if (cudaEventElapsedTime(&ms, ctx->ev[0], ctx->ev[3]) == cudaSuccess) out->device_ms = ms;
    } else {
        cudaGetLastError();
    }
}

int mcp_get_stats(const mcp_ctx* ctx, mcp_stats* out) {
    if (!ctx || !out) return fail(nullptr, MCP_ERR_ARG, "mcp_get_stats: null argument");
    if (ctx->members.empty()) {
        read_stats(ctx, out);
        return 0;
    }
    // multi-device: times are the slowest member's, bytes and launches are summed, the shape is member 0's
    mcp_stats acc{};
    for (size_t g = 0; g < ctx->members.size(); ++g) {
        mcp_stats s;
        cudaSetDevice(ctx->members[g]->device);
        read_stats(ctx->members[g], &s);
        if (g == 0) acc = s;
        else {
            acc.walk_ms = std::max(acc.walk_ms, s.walk_ms);
            acc.device_ms = std::max(acc.device_ms, s.device_ms);
            acc.h2d_bytes += s.h2d_bytes;
            acc.d2h_bytes += s.d2h_bytes;
            acc.kernel_launches += s.kernel_launches;
            acc.tiles += s.tiles;
            acc.scratch_bytes += s.scratch_bytes;
        }
    }
    *out = acc;
    return 0;
}

int mcp_get_stats_member(const mcp_ctx* ctx, int member, mcp_stats* out) {
    if (!ctx || !out) return fail(nullptr, MCP_ERR_ARG, "mcp_get_stats_member: null argument");
    if (ctx->members.empty()) {
        if (member != 0) return fail(nullptr, MCP_ERR_ARG, "mcp_get_stats_member: member out of range");
        read_stats(ctx, out);
        return 0;
    }
    if (member < 0 || member >= (int)ctx->members.size()) return fail(nullptr, MCP_ERR_ARG, "mcp_get_stats_member: member out of range");
    cudaSetDevice(ctx->members[member]->device);
    read_stats(ctx->members[member], out);
    return 0;
}

int mcp_timer_start(mcp_ctx* ctx) {
    if (!ctx) return fail(nullptr, MCP_ERR_ARG, "null context");
    return for_members_or_self(ctx, [](mcp_ctx* m) -> int {
        CUDA_TRY(m, cudaSetDevice(m->device));
        if (!m->ev_t0) {
            CUDA_TRY(m, cudaEventCreate(&m->ev_t0));
            CUDA_TRY(m, cudaEventCreate(&m->ev_t1));
        }
        CUDA_TRY(m, cudaStreamSynchronize(m->copy_stream));
        CUDA_TRY(m, cudaEventRecord(m->ev_t0, m->stream));
        return 0;
    });
}

int mcp_timer_stop(mcp_ctx* ctx, double* ms_out) {
    if (!ctx || !ms_out) return fail(ctx, MCP_ERR_ARG, "mcp_timer_stop: null argument");
    double worst = 0.0;
    int e = for_members_or_self(ctx, [&](mcp_ctx* m) -> int {
        if (!m->ev_t0) return fail(m, MCP_ERR_ARG, "mcp_timer_stop without mcp_timer_start");
        CUDA_TRY(m, cudaSetDevice(m->device));
        CUDA_TRY(m, cudaStreamSynchronize(m->copy_stream));
        CUDA_TRY(m, cudaEventRecord(m->ev_t1, m->stream));
        CUDA_TRY(m, cudaEventSynchronize(m->ev_t1));
        float ms = 0.f;
        CUDA_TRY(m, cudaEventElapsedTime(&ms, m->ev_t0, m->ev_t1));
        worst = std::max(worst, (double)ms);
        return 0;
    });
    if (e) return e;
    *ms_out = worst;
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
    mcp::Schedule sc;
    const bool by_levels = (want_grad & 2) != 0;   // bit 1 of want_grad selects the level-ordered program
    const bool cherries = (want_grad & 4) != 0;    // bit 2: cherries recomputed in the gradient pass (K <= 6 walk kernels)
    want_grad &= 1;
    std::string err = mcp::build_schedule(NN, postorder_num, parent_num, leaf_row, want_grad != 0, sc, by_levels, cherries && !by_levels);
    if (!err.empty()) return fail(nullptr, MCP_ERR_ARG, "%s", err.c_str());
    info[0] = (int32_t)sc.post.size();
    info[1] = (int32_t)sc.pre.size();
    info[2] = sc.n_slots;
    info[3] = sc.n_stack;
    info[4] = sc.n_dnodes;
    info[5] = (int32_t)(sc.post_levels.empty() ? 0 : sc.post_levels.size() - 1);
    info[6] = (int32_t)(sc.pre_levels.empty() ? 0 : sc.pre_levels.size() - 1);
    info[7] = sc.n_cherries;
    if ((int)sc.post.size() > cap_post || (int)sc.pre.size() > cap_pre)
        return fail(nullptr, MCP_ERR_ARG, "mcp_schedule_dump: output arrays too small");
    if (post_ops) std::memcpy(post_ops, sc.post.data(), sc.post.size() * 32);
    if (pre_ops) std::memcpy(pre_ops, sc.pre.data(), sc.pre.size() * 32);
    return 0;
}

int mcp_schedule_fetch_list(int NN, const int32_t* postorder_num, const int32_t* parent_num, const int32_t* leaf_row,
                            int cherries, uint16_t* slots, int cap, int32_t* n_out) {
    if (!postorder_num || !parent_num || !leaf_row || !n_out) return fail(nullptr, MCP_ERR_ARG, "mcp_schedule_fetch_list: null argument");
    mcp::Schedule sc;
    std::string err = mcp::build_schedule(NN, postorder_num, parent_num, leaf_row, true, sc, false, cherries != 0);
    if (!err.empty()) return fail(nullptr, MCP_ERR_ARG, "%s", err.c_str());
    *n_out = (int32_t)sc.pre_fetch.size();
    if ((int)sc.pre_fetch.size() > cap) return fail(nullptr, MCP_ERR_ARG, "mcp_schedule_fetch_list: output array too small");
    if (slots && !sc.pre_fetch.empty()) std::memcpy(slots, sc.pre_fetch.data(), sc.pre_fetch.size() * sizeof(uint16_t));
    return 0;
}

}  // extern "C"
