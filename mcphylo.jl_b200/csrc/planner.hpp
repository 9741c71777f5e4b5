// planner.hpp — from (alignments, topologies, K, R, want_grad) to an evaluation plan: schedules
// (schedule.hpp), launch shape (tile width, columns per thread, persistent grid), tile and
// accumulator-row assignment, and the topology block uploaded to the device.  Host only.
// Part of libmcphylo_b200.so; included by mcphylo_b200.cu only.
#pragma once
#include <algorithm>
#include <cmath>
#include <cstring>
#include <mutex>

#include "host_state.hpp"

namespace {

using mcpdev::KernelTable;
using mcpdev::LaunchCfg;

bool k_templated(int K) { return K >= 2 && K <= 6; }
bool k_supported(int K) { return K >= 2 && K <= KMAX_GENERIC; }

const KernelTable* kernels_for(int K) {
    switch (K) {
        case 2: return mcpdev::kernels_k2();
        case 3: return mcpdev::kernels_k3();
        case 4: return mcpdev::kernels_k4();
        case 5: return mcpdev::kernels_k5();
        case 6: return mcpdev::kernels_k6();
        default: return k_supported(K) ? mcpdev::kernels_generic() : nullptr;
    }
}

size_t walk_smem_bytes(int K, int max_br, int shared_acc, int block, int cpt) {
    const size_t acc = shared_acc ? (((size_t)max_br * 8 + 15) & ~(size_t)15) + walk_part_bytes(CH) : 0;
    return acc + (size_t)3 * CH * 32 + (size_t)2 * CH * 2 * 2 * K * 8 + (size_t)2 * CH * 32 +
           (size_t)2 * CH * 2 * 2 * K * (K + 1) * 8 + (size_t)2 * CH * 2 * 2 * block * cpt;   // = WalkSmem<K>::total
}
// the same kernel with the operand ring (shared accumulator, chunks of WALK_RING_CH ops, ring behind the staging buffers)
size_t walk_smem_bytes_ring(int K, int max_br, int block, int cpt) {
    const int TS = block * cpt, warps = block / 32;
    switch (K) {
        case 2: return WalkSmem<2, WALK_RING_CH>::ring_offset(max_br, 1, TS) + WalkSmem<2, WALK_RING_CH>::ring_bytes(warps, WALK_RING_DEPTH, cpt);
        case 4: return WalkSmem<4, WALK_RING_CH>::ring_offset(max_br, 1, TS) + WalkSmem<4, WALK_RING_CH>::ring_bytes(warps, WALK_RING_DEPTH, cpt);
        default: return 0;
    }
}
// runtime-K kernel: the per-branch gradient accumulator and, for a model-gradient evaluation (mg_K = state count),
// per warp 2 x 32 x K doubles behind it (kernel_generic.cuh: s_mq / s_ml, offset by the accumulator of max_br doubles)
size_t generic_smem_bytes(int max_br, int want_grad, int mg_K = 0, int block = 0) {
    if (mg_K > 0) return (((size_t)max_br * sizeof(double) + 15) & ~(size_t)15) + (size_t)(block / 32) * 64 * mg_K * sizeof(double);
    return want_grad ? (size_t)max_br * sizeof(double) : 0;
}
int mma_kp(int K) { return (K + 7) & ~7; }   // state count padded to the 8-wide MMA blocks
size_t mma_smem_bytes(int K, int max_br, int want_grad) {
    switch (mma_kp(K)) {
        case 8: return MmaSmem<8>::total(max_br, want_grad);
        case 16: return MmaSmem<16>::total(max_br, want_grad);
        case 24: return MmaSmem<24>::total(max_br, want_grad);
        default: return MmaSmem<32>::total(max_br, want_grad);
    }
}

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
    // model-gradient evaluation (mcp_eval_model_gradient): one tree, runtime-K kernel for every K, moment matrices
    // accumulated next to the branch gradient
    int model_grad = 0;
};

// Resident CTAs per SM of the walk kernel for a launch shape.  The occupancy calculator costs a few
// microseconds per query and the planner asks for several shapes, so answers are remembered per process.
int walk_occupancy(mcp_ctx* ctx, const KernelTable* kt, int K, int block, int cpt, size_t smem, bool sscr, bool accg,
                   bool levels, int* out, bool mma = false, bool ring = false) {
    struct Key { int device, K, block, cpt; size_t smem; bool sscr, accg, levels, mma, ring; int occ; };
    static std::mutex mu;
    static std::vector<Key> memo;
    {
        std::lock_guard<std::mutex> lock(mu);
        for (const Key& k : memo)
            if (k.device == ctx->device && k.K == K && k.block == block && k.cpt == cpt && k.smem == smem && k.sscr == sscr &&
                k.accg == accg && k.levels == levels && k.mma == mma && k.ring == ring) {
                *out = k.occ;
                return 0;
            }
    }
    LaunchCfg c;
    c.device = ctx->device;
    c.K = K;
    c.block = block;
    c.cpt = cpt;
    c.smem = smem;
    c.smem_scratch = sscr;
    c.acc_global = accg;
    c.mma = mma;
    c.ring = ring;
    int occ = 0;
    cudaError_t e = levels ? kt->occupancy_levels(c, &occ) : kt->occupancy_walk(c, &occ);
    if (e != cudaSuccess) {
        cudaGetLastError();
        // a shape that cannot be configured (too much shared memory) simply does not fit
        if (e == cudaErrorInvalidValue) occ = 0;
        else return fail(ctx, MCP_ERR_CUDA, "occupancy query failed: %s", cudaGetErrorString(e));
    }
    std::lock_guard<std::mutex> lock(mu);
    memo.push_back({ctx->device, K, block, cpt, smem, sscr, accg, levels, mma, ring, occ});
    *out = occ;
    return 0;
}

struct WalkShape {
    int block = 256, cpt = 1, occ = 0;
};

long long tiles_for(const BatchArgs& a, int tile_w) {
    long long n = 0;
    for (int t = 0; t < a.T; ++t) n += ((a.alns[t]->S + tile_w - 1) / tile_w) * a.R;
    return n;
}

// Launch shape of the depth-first walk.  Measured on cfg4 shards of 15 k .. 250 k sites for every tile width
// in {64 .. 256} x {1, 2} columns per thread (profiles/r2_shape_sweep.json):
//   - the widest tile (256 threads) wins at every size, also when it leaves SMs without a CTA: the cost of a
//     tile is dominated by per-CTA work that does not shrink with the tile (chunk staging, barriers), and a
//     15 k-site input costs one tile time (1.5 ms for a 1999-node tree) whatever the width -- 64-wide tiles
//     take 2x longer;
//   - a partly filled last round of tiles costs nothing measurable (3.3 rounds run at the per-site cost of
//     26 rounds: CTAs drift apart and the stragglers speed up as the SMs empty), so no width is ever chosen
//     to "fill the last round";
//   - two columns per thread win as soon as one column per thread would need more than two CTAs per SM --
//     for K = 2 always (1.28x, round 1), for K = 3, 4 only on large trees: cfg4's 1999-node tree gains 2-6 %,
//     cfg3's 399-node tree LOSES 5 % (1.52 vs 1.60 ms, profiles/r2_shape_sweep_cfg3.json; its smaller
//     accumulator leaves room for 24 resident warps at one column per thread).
// Hence: 256 threads, narrowed only for inputs of fewer than 256 columns; two columns per thread when the input
// has more than 2 x SMs tiles of 256 columns and (K = 2, or K <= 4 and a tree of at least 1000 nodes).
int choose_walk_shape(mcp_ctx* ctx, const KernelTable* kt, const BatchArgs& a, int K, int max_br, bool acc_global,
                      WalkShape* out) {
    const int R = a.R;
    long long total_cols = 0, widest = 0;
    for (int t = 0; t < a.T; ++t) {
        total_cols += a.alns[t]->S * R;
        widest = std::max<long long>(widest, a.alns[t]->S);
    }
    const bool templated = k_templated(K) && !a.model_grad;
    const int shared_acc = a.want_grad && !acc_global ? 1 : 0;
    if (!templated && ctx->opt_mma != 0 && !a.model_grad) {   // large alphabets: fixed shape, 8 warps x 16 columns (kernel_mma.cuh)
        int e0, o = 0;
        if ((e0 = walk_occupancy(ctx, kt, K, MMA_WARPS * 32, 1, mma_smem_bytes(K, max_br, a.want_grad), false, false, false, &o, true))) return e0;
        *out = {MMA_WARPS * 32, 1, o};
        return 0;
    }
    int block = ctx->opt_block;
    if (block <= 0) {
        block = templated ? 256 : 128;      // runtime-K kernel: at most 128 threads per CTA
        while (block > 32 && widest <= block / 2) block >>= 1;
    }
    if (!templated && block > 128) block = 128;
    int cpt = ctx->opt_cpt;
    const bool cpt2_ok = templated && K <= 4 && !acc_global;
    int max_nn = 0;
    for (int t = 0; t < a.T; ++t) max_nn = std::max(max_nn, (int)a.NN[t]);
    // Round 2, kernels with the operand ring and D = P L stored (K = 4, gradient evaluations): two columns per thread
    // win for EVERY tree size once the input is more than ~5 tiles of 256 columns per SM (20 .. 500 taxa: -5 .. -13 %,
    // cfg3 1.40 -> 1.33 ms); below that the 444 resident one-column CTAs cover the input in fewer rounds
    // (profiles/r2_ab_columns_per_thread.json).
    const bool ring_k4 = K == 4 && a.want_grad && ctx->opt_ring != 0 && !acc_global;
    if (cpt <= 0) {
        const long long tiles = tiles_for(a, block);
        cpt = cpt2_ok && ((tiles > 2LL * ctx->sm_count && (K == 2 || max_nn >= 1000)) || (ring_k4 && tiles > 5LL * ctx->sm_count)) ? 2 : 1;
    }
    if (!cpt2_ok) cpt = 1;
    const size_t smem = templated ? walk_smem_bytes(K, max_br, shared_acc, block, cpt)
                                  : generic_smem_bytes(max_br, a.want_grad, a.model_grad ? K : 0, block);
    int e, occ = 0;
    if ((e = walk_occupancy(ctx, kt, K, block, cpt, smem, false, acc_global, false, &occ))) return e;
    *out = {block, cpt, occ};
    return 0;
}

bool plan_matches(const Plan& pl, const BatchArgs& a, int K) {
    if (!pl.valid || (int)pl.sig.size() != a.T || pl.want_grad != a.want_grad || pl.K != K || pl.R != a.R ||
        pl.model_grad != (a.model_grad != 0))
        return false;
    for (int t = 0; t < a.T; ++t) {
        const auto& s = pl.sig[t];
        if (s.aln_id != a.alns[t]->id || s.NN != a.NN[t] ||
            std::memcmp(s.po.data(), a.po[t], sizeof(int32_t) * a.NN[t]) != 0 ||
            std::memcmp(s.pa.data(), a.pa[t], sizeof(int32_t) * a.NN[t]) != 0)
            return false;
    }
    return true;
}

void invalidate_plans(mcp_ctx* ctx, unsigned long long aln_id = 0) {
    for (auto& pl : ctx->plans) {
        if (aln_id) {
            bool uses = false;
            for (const auto& s : pl->sig) uses = uses || s.aln_id == aln_id;
            if (!uses) continue;
        }
        pl->valid = false;
        pl->uploaded = false;
        pl->sig.clear();
    }
}

// Fills `pl` for the batch.  On failure the plan is left invalid.
int build_plan(mcp_ctx* ctx, const BatchArgs& a, int K, Plan& pl) {
    const int T = a.T, R = a.R;
    pl.valid = false;
    pl.uploaded = false;
    pl.sig.clear();
    // model-gradient evaluations run the runtime-K kernel whatever K is
    const bool templ = k_templated(K) && !a.model_grad;
    const KernelTable* kt = a.model_grad ? mcpdev::kernels_generic() : kernels_for(K);
    if (!kt) return fail(ctx, MCP_ERR_UNSUPPORTED, "no kernel compiled for K = %d states", K);
    long long total_cols = 0;
    int max_nn = 0;
    for (int t = 0; t < T; ++t) {
        total_cols += a.alns[t]->S * R;
        max_nn = std::max(max_nn, (int)a.NN[t]);
    }
    // Very large trees: the per-branch accumulator no longer fits in shared memory next to the staging
    // buffers; those kernels exist with one column per thread only.
    const bool acc_global = a.want_grad && templ && walk_acc_global(max_nn, ctx->opt_acc_mode);
    // Small inputs (a few one-warp tiles per SM): level-parallel kernel, a tile is 32 columns wide
    // and is worked on by all 8 warps of a 256-thread CTA.
    bool level_mode = templ && ctx->opt_levels != 0 &&
                      (ctx->opt_levels == 1 || (ctx->opt_block == 0 && total_cols <= 32LL * 4 * ctx->sm_count));

    long long n_ops = 0, out_off = 0, dyn_off = 0, btab_off = 0, n_lvl_ints = 0;
    int n_slots = 1, n_stack = 1, max_br = 1, max_rows = 1;
    std::vector<int32_t> leaf_row;
    auto build_all = [&](bool by_levels) -> int {
        pl.scheds.assign(T, mcp::Schedule());
        pl.trees.assign(T, TreeDev());
        n_ops = out_off = dyn_off = btab_off = n_lvl_ints = 0;
        n_slots = 1; n_stack = 1; max_br = 1; max_rows = 1;
        for (int t = 0; t < T; ++t) {
            const mcp_alignment* al = a.alns[t];
            const int NN = a.NN[t];
            if (NN < 2) return fail(ctx, MCP_ERR_ARG, "tree %d: NN must be >= 2", t);
            leaf_row.assign(NN, -1);
            for (int i = 0; i < al->n_leaves; ++i) {
                int num = al->leaf_nums[i];
                if (num >= 1 && num <= NN) leaf_row[num - 1] = i;
            }
            std::string err = mcp::build_schedule(NN, a.po[t], a.pa[t], leaf_row.data(), a.want_grad != 0, pl.scheds[t], by_levels,
                                                  !by_levels && templ && ctx->opt_cherry != 0);
            if (!err.empty()) return fail(ctx, MCP_ERR_ARG, "tree %d: %s", t, err.c_str());
            const mcp::Schedule& sc = pl.scheds[t];
            TreeDev& td = pl.trees[t];
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
            td.n_rows = al->n_leaves;
            td.lvl_off = (int)n_lvl_ints;
            td.n_post_lvl = sc.post_levels.empty() ? 0 : (int)sc.post_levels.size() - 1;
            td.n_pre_lvl = sc.pre_levels.empty() ? 0 : (int)sc.pre_levels.size() - 1;
            n_lvl_ints += (long long)sc.post_levels.size() + (long long)sc.pre_levels.size();
            max_rows = std::max(max_rows, al->n_leaves);
            n_slots = std::max(n_slots, sc.n_slots);
            n_stack = std::max(n_stack, sc.n_stack);
            max_br = std::max(max_br, sc.n_dnodes);
        }
        return 0;
    };
    int e;
    if (level_mode) {
        if ((e = build_all(true))) return e;
        const size_t need = LevelSmem::total(max_br, a.want_grad ? 1 : 0, max_rows, n_slots, 0, K);
        if (need > 160 * 1024) level_mode = false;     // tree too large for the shared-memory path
    }
    if (!level_mode && (e = build_all(false))) return e;

    // launch shape
    const bool mma = !level_mode && !k_templated(K) && ctx->opt_mma != 0 && !a.model_grad;
    int block, cpt, occ = 0;
    size_t smem;
    if (level_mode) {
        block = 256;
        cpt = 1;
        smem = LevelSmem::total(max_br, a.want_grad ? 1 : 0, max_rows, n_slots, 0, K);   // pre vectors reuse the post slots
        if ((e = walk_occupancy(ctx, kt, K, block, cpt, smem, false, false, true, &occ))) return e;
    } else {
        WalkShape ws;
        if ((e = choose_walk_shape(ctx, kt, a, K, max_br, acc_global, &ws))) return e;
        block = ws.block;
        cpt = ws.cpt;
        occ = ws.occ;
        smem = templ ? walk_smem_bytes(K, max_br, a.want_grad && !acc_global ? 1 : 0, block, cpt)
               : mma ? mma_smem_bytes(K, max_br, a.want_grad)
                     : generic_smem_bytes(max_br, a.want_grad, a.model_grad ? K : 0, block);
    }
    const int tile_w = level_mode ? 32 : mma ? MMA_TILE : block * cpt;
    const int kdim = mma ? mma_kp(K) : K;       // doubles per column of a stored partial
    int tile_cursor = 0;
    for (int t = 0; t < T; ++t) {
        TreeDev& td = pl.trees[t];
        td.tiles_per_rate = (int)((td.S + (long long)tile_w - 1) / (long long)tile_w);
        td.tile_begin = tile_cursor;
        const long long nt = (long long)td.tiles_per_rate * R;
        if (tile_cursor + nt > 0x7fffffffLL) return fail(ctx, MCP_ERR_ARG, "too many column tiles");
        tile_cursor += (int)nt;
    }
    pl.level_mode = level_mode;
    pl.mma = mma;
    pl.model_grad = a.model_grad != 0;
    pl.max_rows = max_rows;
    pl.want_grad = a.want_grad;
    pl.cpt = cpt;
    pl.K = K;
    pl.R = R;
    pl.n_tiles = tile_cursor;
    pl.block = block;
    pl.n_slots = n_slots;
    pl.n_stack = a.want_grad && !level_mode ? n_stack : 0;
    pl.max_br = max_br;
    pl.total_out = out_off;
    pl.total_dyn = dyn_off;
    pl.total_btab = btab_off;
    pl.acc_global = acc_global && !level_mode;
    pl.smem_bytes = smem;
    // Small problems: keep the partials scratch in shared memory (latency path).  Automatic when the
    // whole input is a handful of tiles per SM and the scratch of one CTA fits next to the staging
    // buffers.
    pl.smem_scratch = false;
    {
        const size_t scr_bytes = (size_t)(pl.n_slots + pl.n_stack) * tile_w * kdim * 8;
        const bool fits = !level_mode && !acc_global && templ && cpt == 1 && pl.smem_bytes + scr_bytes <= 96 * 1024;
        pl.smem_scratch = fits && (ctx->opt_smem_scratch == 1 || (ctx->opt_smem_scratch < 0 && pl.n_tiles <= 4 * ctx->sm_count));
        if (pl.smem_scratch) {
            pl.smem_bytes += scr_bytes;
            if ((e = walk_occupancy(ctx, kt, K, block, cpt, pl.smem_bytes, true, false, false, &occ))) return e;
        }
    }
    if (pl.smem_bytes > 200 * 1024)
        return fail(ctx, MCP_ERR_UNSUPPORTED, "tree with %d nodes exceeds the shared-memory gradient accumulator", max_br);
    if (occ < 1) return fail(ctx, MCP_ERR_CUDA, "walk kernel does not fit on an SM (block %d, smem %zu)", block, pl.smem_bytes);
    // Operand ring for the gradient pass: taken when that variant keeps as many CTAs per SM as the plain
    // kernel (the persistent grid and the accumulator rows below are shared by both).
    pl.ring = false;
    pl.smem_ring = 0;
    // Automatic wherever the kernel exists (K = 2, 4): cfg4 -14 % at 1 M sites, -22 % on a 125 k-site shard, cfg3 -4 %;
    // K = 2 (cfg2's 50-taxon and cfg5's 100-taxon tree on millions of sites) -14 % / -19 %, where the shorter chunks
    // also make room for a third CTA per SM.
    const bool ring_wanted = ctx->opt_ring != 0;
    if (a.want_grad && ring_wanted && templ && !level_mode && !acc_global && !pl.smem_scratch && walk_ring_supported(K)) {
        bool lists = true;
        for (int t = 0; t < T; ++t) lists = lists && pl.scheds[t].n_slots < 65535;
        const size_t sr = walk_smem_bytes_ring(K, max_br, block, cpt);
        int occ_r = 0;
        if (lists && sr <= 220 * 1024) {
            if ((e = walk_occupancy(ctx, kt, K, block, cpt, sr, false, false, false, &occ_r, false, true))) return e;
            // The persistent grid is sized for the ring kernel; the plain kernel -- the launch-time fallback for
            // a decomposition without a null eigenvalue or a batch with several models -- runs the same grid
            // (and the same accumulator rows) even where it would fit fewer or more CTAs per SM itself.
            if (occ_r >= occ) {
                pl.ring = true;
                pl.smem_ring = sr;
                occ = occ_r;
            }
        }
    }
    if (ctx->opt_ctas_per_sm > 0) occ = std::min(occ, ctx->opt_ctas_per_sm);
    pl.grid = (int)std::min<long long>((long long)pl.n_tiles, (long long)occ * ctx->sm_count);
    if (pl.grid < 1) pl.grid = 1;
    {   // very large trees: fewer persistent CTAs rather than a scratch allocation that cannot succeed
        size_t free_b = 0, total_b = 0;
        const double per_cta = level_mode ? 0.0 : (double)(pl.n_slots + pl.n_stack) * tile_w * kdim * 8.0;
        if (per_cta * pl.grid <= (double)ctx->d_scratch.cap) {
            // fits the scratch already held: nothing to allocate, no need to ask the driver
            // (cudaMemGetInfo costs milliseconds on a GPU with many live allocations)
        } else if (cudaMemGetInfo(&free_b, &total_b) == cudaSuccess && per_cta > 0) {
            const double budget = 0.6 * ((double)free_b + (double)ctx->d_scratch.cap);
            if (per_cta * pl.grid > budget) pl.grid = (int)std::max(1.0, std::floor(budget / per_cta));
        } else {
            cudaGetLastError();
        }
    }

    // accumulator rows: one per (CTA, tree) pair in CTA order (also tree order)
    std::vector<int32_t> row_base(pl.grid, 0);
    {
        const int q = pl.n_tiles / pl.grid, rem = pl.n_tiles % pl.grid;
        int row = 0, ti = 0;
        for (int t = 0; t < T; ++t) pl.trees[t].row_lo = pl.trees[t].row_hi = 0;
        std::vector<char> seen(T, 0);
        for (int c = 0; c < pl.grid; ++c) {
            int t0 = c * q + std::min(c, rem), t1 = t0 + q + (c < rem ? 1 : 0);
            row_base[c] = row;
            int tile = t0;
            while (ti < T - 1 && tile >= pl.trees[ti].tile_begin + R * pl.trees[ti].tiles_per_rate) ++ti;
            int tj = ti;
            while (tile < t1) {
                int tend = std::min(t1, pl.trees[tj].tile_begin + R * pl.trees[tj].tiles_per_rate);
                if (!seen[tj]) { pl.trees[tj].row_lo = row; seen[tj] = 1; }
                ++row;
                pl.trees[tj].row_hi = row;
                tile = tend;
                ++tj;
            }
        }
        pl.n_rows = row;
    }
    pl.row_stride = (max_br + 3) & ~3;
    if ((!level_mode && (double)(pl.n_slots + n_stack + 1) * tile_w * kdim * 8.0 >= 4.0e9) || (double)max_br * R * bt_size(K) * 8.0 >= 4.0e9)
        return fail(ctx, MCP_ERR_UNSUPPORTED, "tree too large for 32-bit scratch offsets (%d nodes)", max_br);
    pl.scratch_per_cta = level_mode ? 4 : (long long)(pl.n_slots + pl.n_stack) * tile_w * kdim;

    // topology block: [TreeDev x T][ops][row_base][level offsets]
    pl.off_trees = 0;
    pl.off_ops = (sizeof(TreeDev) * T + 31) & ~(size_t)31;
    pl.off_rowbase = pl.off_ops + (size_t)n_ops * 32;
    pl.off_levels = pl.off_rowbase + sizeof(int32_t) * pl.grid;
    pl.off_fetch = (pl.off_levels + sizeof(int32_t) * (size_t)std::max<long long>(n_lvl_ints, 1) + 15) & ~(size_t)15;
    // fetch lists of the operand ring: per tree its entries + WALK_RING_DEPTH + 2 end marks 0xffff (the ring reads ahead)
    size_t n_fetch_total = 0;
    for (int t = 0; t < T; ++t) {
        pl.trees[t].fetch_off = (int)n_fetch_total;
        pl.trees[t].n_fetch = pl.ring ? (int)pl.scheds[t].pre_fetch.size() : 0;
        n_fetch_total += (size_t)pl.trees[t].n_fetch + WALK_RING_DEPTH + 2;
    }
    pl.topo_bytes = pl.off_fetch + sizeof(uint16_t) * n_fetch_total;
    if ((e = ensure_pin(ctx, pl.h_topo, pl.topo_bytes))) return e;
    if ((e = ensure_dev(ctx, pl.d_topo, pl.topo_bytes))) return e;
    char* h = (char*)pl.h_topo.p;
    std::memcpy(h + pl.off_trees, pl.trees.data(), sizeof(TreeDev) * T);
    char* ho = h + pl.off_ops;
    for (int t = 0; t < T; ++t) {
        const mcp::Schedule& sc = pl.scheds[t];
        std::memcpy(ho, sc.post.data(), sc.post.size() * 32);
        ho += sc.post.size() * 32;
        std::memcpy(ho, sc.pre.data(), sc.pre.size() * 32);
        ho += sc.pre.size() * 32;
    }
    std::memcpy(h + pl.off_rowbase, row_base.data(), sizeof(int32_t) * pl.grid);
    {
        int32_t* hl = (int32_t*)(h + pl.off_levels);
        for (int t = 0; t < T; ++t) {
            const mcp::Schedule& sc = pl.scheds[t];
            for (int32_t v : sc.post_levels) *hl++ = v;
            for (int32_t v : sc.pre_levels) *hl++ = v;
        }
    }
    {
        uint16_t* hf = (uint16_t*)(h + pl.off_fetch);
        std::memset(hf, 0xff, sizeof(uint16_t) * n_fetch_total);
        for (int t = 0; t < T; ++t)
            if (pl.trees[t].n_fetch) std::memcpy(hf + pl.trees[t].fetch_off, pl.scheds[t].pre_fetch.data(), sizeof(uint16_t) * pl.trees[t].n_fetch);
    }
    for (int t = 0; t < T; ++t) {
        Plan::TreeSig sg;
        sg.aln_id = a.alns[t]->id;
        sg.NN = a.NN[t];
        sg.po.assign(a.po[t], a.po[t] + a.NN[t]);
        sg.pa.assign(a.pa[t], a.pa[t] + a.NN[t]);
        pl.sig.push_back(std::move(sg));
    }
    pl.valid = true;
    return 0;
}

// The plan for this batch: a cached one, or a freshly built one in the least recently used slot.
int get_plan(mcp_ctx* ctx, const BatchArgs& a, int K, Plan** out, bool* rebuilt) {
    ++ctx->clock;
    for (auto& pl : ctx->plans)
        if (plan_matches(*pl, a, K)) {
            pl->stamp = ctx->clock;
            *out = pl.get();
            *rebuilt = false;
            return 0;
        }
    Plan* slot = nullptr;
    for (auto& pl : ctx->plans)
        if (!pl->valid) { slot = pl.get(); break; }
    if (!slot && (int)ctx->plans.size() < MCP_PLAN_SLOTS) {
        ctx->plans.emplace_back(new Plan());
        slot = ctx->plans.back().get();
    }
    if (!slot) {
        // Recycle the least recently used plan: its topology block on the device is overwritten by a copy
        // on the evaluation stream, i.e. after every kernel that still reads it.
        for (auto& pl : ctx->plans)
            if (!slot || pl->stamp < slot->stamp) slot = pl.get();
    }
    if (slot->h_topo.p && ctx->pending_async) {   // its pinned topology copy may still be the source of an upload
        CUDA_TRY(ctx, cudaStreamSynchronize(ctx->stream));
        ctx->pending_async = false;
    }
    int e = build_plan(ctx, a, K, *slot);
    if (e) {
        slot->valid = false;
        slot->sig.clear();
        return e;
    }
    slot->stamp = ctx->clock;
    *out = slot;
    *rebuilt = true;
    return 0;
}

}  // namespace
