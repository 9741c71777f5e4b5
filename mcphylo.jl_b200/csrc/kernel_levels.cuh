// kernel_levels.cuh — kernel 2a: level-parallel small-tree kernel (tables + walk + final reduction in one launch).
// Part of libmcphylo_b200.so; instantiated per state count K in walk_k*.cu (walk_inst.cuh).
#pragma once

namespace {

// --------------------------------------------------------------------------------------------
// kernel 2c: small-tree latency path.  A tile is ONE warp wide (32 columns) but is worked on by all
// W warps of the CTA: the level-ordered program (schedule.hpp, by_levels) lists ops of equal height
// (post pass) / depth (gradient pass) together, the warps split each level's ops, and a
// __syncthreads separates levels.  All partials and pre vectors of the tile live in shared memory.
// The critical path is the tree height instead of the node count; used when the whole input is only
// a few tiles per SM (MCMC-sized problems), where the depth-first walk runs at single-warp latency.
// --------------------------------------------------------------------------------------------
constexpr int LEVEL_FIN_GROUP = 8;          // CTAs per group of the two-level final reduction
constexpr int LEVEL_FIN_MIN_GRID = 192;     // fewer CTAs: the single last CTA walks the rows (at most 6 round trips)
constexpr int LEVEL_FIN_MAX_GROUPS = 62;    // group counters behind done_counter[0] (64 words); also the extra rows the host allocates

template <int K, bool DYN_MODEL>
// 3 CTAs per SM (<= 85 registers): an MCMC-sized input is one tile per CTA, and all of them should be resident at once
__global__ void __launch_bounds__(256, 3) felsenstein_walk_levels(const __grid_constant__ LevelParams lp) {
    const WalkParams& p = lp.w;
    // per-evaluation parameters: in device memory, or inline in the kernel arguments (device_layout.cuh: LevelParams)
    const double* const dynp = p.dyn != nullptr ? p.dyn : lp.dyn_inline;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    __shared__ long long s_e[8];
    __shared__ double s_l[8];

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
    // pre[i] takes over the slot of post[i]: a family's op reads the children's post partials into registers
    // before it writes their pre vectors, and nothing reads post[i] after its mother's op -- half the slots
    double* const s_pre = s_post;
    // the tree's op program and level offsets, copied once per (CTA, tree): read from global memory they put
    // two dependent L2 round trips (level bounds, then the descriptors) on the critical path of EVERY level
    int4* const s_ops = reinterpret_cast<int4*>(reinterpret_cast<unsigned char*>(s_tab) + LevelSmem::tab_bytes(p.max_br, K) +
                                                LevelSmem::slots_bytes(p.n_slots, p.n_stack, K));
    int* const s_lvl = reinterpret_cast<int*>(s_ops + 4 * (size_t)p.max_br);
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
        const ModelT<K, DYN_MODEL> mdl{p, DYN_MODEL ? (int)dynp[tr.dyn_off + dyn_slot(tr.NN, K, R)] * MODEL_SLOT : 0};
        // post program, then pre program (contiguous in the topology block); likewise the level offsets
        cp_async_wait_all();                                   // (a tree without a tile for this CTA leaves its program copy pending)
        __syncthreads();                                       // the previous tree's program is no longer in use
#ifdef MCP_LEVEL_SYNC_STAGING    // A/B builds: every thread loads and stores its own words (round 2's first half)
        for (int i = tid; i < 2 * (tr.n_post + tr.n_pre); i += NT) s_ops[i] = __ldg(p.ops + 2 * tr.post_off + i);
#else                            // asynchronous copies, completed together with the first tile's codes below
        for (int i = tid; i < 2 * (tr.n_post + tr.n_pre); i += NT) cp_async16(s_ops + i, p.ops + 2 * tr.post_off + i);
        cp_async_commit();
#endif
        for (int i = tid; i < tr.n_post_lvl + tr.n_pre_lvl + 2; i += NT) s_lvl[i] = __ldg(p.levels + tr.lvl_off + i);
        const int4* const post_ops = s_ops;
        const int4* const pre_ops = s_ops + 2 * tr.n_post;
        const int* const post_lvl = s_lvl;
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
#ifndef MCP_LEVEL_SYNC_STAGING
            // This tile's state codes, all leaves: one 16-byte asynchronous copy per (leaf row, half tile), requested BEFORE
            // the branch table is built so that the one memory round trip runs under that arithmetic.  (Byte loads by
            // every thread were 6-7 dependent round trips at cfg2's 50 leaves: 12 % of the kernel's stall samples.)
            // Rows are 1024-byte-padded and tiles start at multiples of 32 sites: 16-byte aligned, never past the row.
            for (int i = tid; i < tr.n_rows * 2; i += NT)
                cp_async16(s_code + i * 16, tr.codes + (long long)(i >> 1) * tr.code_stride + site0 + (i & 1) * 16);
            cp_async_commit();
#endif
            if (r != built_rate) {
                // branch table of (tree, rate r), built here instead of by a separate kernel:
                // e = exp(t * D mu rate), P = U diag(e) Uinv, dP = U diag(D mu rate e) Uinv, plus the
                // row-sum columns (same operation order as build_branch_tables)
                const double* const blv = dynp + tr.dyn_off;
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
                    const double t = blv[br];
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
#ifdef MCP_LEVEL_SYNC_STAGING
            for (int i = tid; i < tr.n_rows * 32; i += NT) {   // this tile's state codes, all leaves
                const int rw = i >> 5, l = i & 31;
                s_code[i] = (site0 + l < tr.S) ? __ldg(tr.codes + (long long)rw * tr.code_stride + site0 + l) : (unsigned char)K;
            }
#endif
            if (tid < 32) s_exp[tid] = 0;
            cp_async_wait_all();
            __syncthreads();

            // columns behind the last site read whatever the row holds there (inside its padded stride): masked here
            auto leaf_code = [&](int src) -> int { return (src >= 0 && valid) ? min((int)s_code[src * 32 + lane], K) : K; };
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
                const int lo = post_lvl[lv], hi = post_lvl[lv + 1];
                for (int i = lo + warp; i < hi; i += W) {
                    const int4 o0 = post_ops[2 * i], o1 = post_ops[2 * i + 1];
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
                    const int lo = pre_lvl[lv], hi = pre_lvl[lv + 1];
                    for (int i = lo + warp; i < hi; i += W) {
                        const int4 o0 = pre_ops[2 * i], o1 = pre_ops[2 * i + 1];
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
    //
    // One tree and a few hundred CTAs (an MCMC-sized evaluation: one tile and one row per CTA): the walk of the row
    // list by a single CTA is 10 dependent L2 round trips at cfg2 (313 rows), so there the reduction is a two-level tree -- the last CTA to finish in each group of LEVEL_FIN_GROUP consecutive CTAs
    // adds the group's rows (all of them requested at once: one round trip) into a group row behind the CTA rows,
    // and the last GROUP to finish adds the group rows.  Fixed grouping and fixed order of additions: reproducible.
    // Measured (profiles/r2e_ab_small_tree.jsonl): cfg2 kernel 39.6 -> 37.6 us L2-warm, 46.7 -> 43.2 us L2-cold; the second
    // counter round trip costs ~3 us, so a 32-CTA launch (cfg1) LOSES 3 us -- hence only above LEVEL_FIN_MIN_GRID CTAs.
    // done_counter[0] counts CTAs (one level) or finished groups (two levels), done_counter[1 + g] the CTAs of group g.
    __threadfence();
    __syncthreads();
    const int n_groups = ((int)gridDim.x + LEVEL_FIN_GROUP - 1) / LEVEL_FIN_GROUP;
#ifdef MCP_LEVEL_FIN_SINGLE     // A/B builds (tools/build_variant.py): the one-level reduction of round 2's first half
    const bool two_level = false;
#else
    const bool two_level = p.T == 1 && gridDim.x > LEVEL_FIN_MIN_GRID && n_groups <= LEVEL_FIN_MAX_GROUPS;
#endif
    bool last;
    if (two_level) {
        const int g = (int)blockIdx.x / LEVEL_FIN_GROUP, g_lo = g * LEVEL_FIN_GROUP, g_hi = min(g_lo + LEVEL_FIN_GROUP, (int)gridDim.x);
        if (tid == 0) s_ticket = atomicAdd(p.done_counter + 1 + g, 1u);
        __syncthreads();
        if (s_ticket != (unsigned)(g_hi - g_lo - 1)) return;
        __threadfence();
        const int nb = p.trees[0].NN - 1;
        if (p.want_grad) {
            for (int j = tid; j < nb; j += NT) {
                double v[LEVEL_FIN_GROUP];
#pragma unroll
                for (int b = 0; b < LEVEL_FIN_GROUP; ++b)
                    v[b] = g_lo + b < g_hi ? __ldcg(p.rows + (long long)(g_lo + b) * p.row_stride + j) : 0.0;
                double acc = v[0];
#pragma unroll
                for (int b = 1; b < LEVEL_FIN_GROUP; ++b)
                    if (g_lo + b < g_hi) acc += v[b];
                p.rows[(long long)((int)gridDim.x + g) * p.row_stride + j] = acc;
            }
        }
        if (warp == 0) {
            static_assert(LEVEL_FIN_GROUP == 8, "the shuffle tree below adds groups of 8 lanes");
            long long es = 0;
            double ls = 0.0;
            if (lane < LEVEL_FIN_GROUP && g_lo + lane < g_hi) {
                es = __ldcg(&p.rows_ll[g_lo + lane].esum);
                ls = __ldcg(&p.rows_ll[g_lo + lane].logsum);
            }
            for (int off = 4; off > 0; off >>= 1) {
                es += __shfl_xor_sync(0xffffffffu, es, off);
                ls += __shfl_xor_sync(0xffffffffu, ls, off);
            }
            if (lane == 0) {
                p.rows_ll[(int)gridDim.x + g].esum = es;
                p.rows_ll[(int)gridDim.x + g].logsum = ls;
            }
        }
        __threadfence();
        __syncthreads();
        if (tid == 0) s_ticket = atomicAdd(p.done_counter, 1u);
        __syncthreads();
        last = s_ticket == (unsigned)(n_groups - 1);
    } else {
        if (tid == 0) s_ticket = atomicAdd(p.done_counter, 1u);
        __syncthreads();
        last = s_ticket == gridDim.x - 1;
    }
    if (last) {
        __threadfence();
        // All 8 warps take part: warp w sums rows row_lo + w, row_lo + w + W, ... (lanes across the
        // branches, so a warp reads one contiguous run per row), the per-warp partial sums meet in
        // shared memory and are added in warp order.  Fixed order, hence reproducible; ~10x less
        // latency than one thread per output walking all rows (rows = CTAs of the launch).
        double* const s_fin = s_tab;               // W x NN doubles; the tile buffers are free now
        // block reduction of the branch-length prior: 2 x 256 doubles in the (idle) slot region
        double* const s_prior = reinterpret_cast<double*>(reinterpret_cast<unsigned char*>(s_tab) + LevelSmem::tab_bytes(p.max_br, K));
        for (int t = 0; t < p.T; ++t) {
            const TreeDev tr = p.trees[t];
            // rows to add: the tree's own CTA rows, or the group rows behind them
            const int fin_lo = two_level ? (int)gridDim.x : tr.row_lo, fin_hi = two_level ? (int)gridDim.x + n_groups : tr.row_hi;
            double* o = p.out + tr.out_off;
            const double* d = dynp + tr.dyn_off;
            const double* hdr = d + dyn_prior(tr.NN, K, R);
            const bool prior = hdr[0] != 0.0;
            PriorSums ps{0.0, 0.0};
            if (prior) ps = prior_block_sums(d + dyn_blv(tr.NN), hdr + 4, tr.NN - 1, tid, NT, s_prior);
            {
                long long es = 0;
                double ls = 0.0;
                for (int rw = fin_lo + tid; rw < fin_hi; rw += NT) {
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
                    // 4 rows are requested before the first is added (same order of additions): one L2 round
                    // trip per 4 rows instead of one per row -- the walk of the row list used to be 40 % of a
                    // cfg2 evaluation
                    for (int rw0 = fin_lo + warp; rw0 < fin_hi; rw0 += 4 * W) {
                        double v[4][4];
#pragma unroll
                        for (int b = 0; b < 4; ++b) {
                            const int rw = rw0 + b * W;
                            const double* rp = p.rows + (long long)rw * p.row_stride + j0;
#pragma unroll
                            for (int u = 0; u < 4; ++u) v[b][u] = (rw < fin_hi && j0 + 32 * u < nb) ? __ldcg(rp + 32 * u) : 0.0;
                        }
#pragma unroll
                        for (int b = 0; b < 4; ++b)
#pragma unroll
                            for (int u = 0; u < 4; ++u)
                                if (rw0 + b * W < fin_hi) acc[u] += v[b][u];
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
        // ready for the next launch
        if (tid == 0) *p.done_counter = 0;
        if (two_level && tid < n_groups) p.done_counter[1 + tid] = 0;
    }
}

}  // namespace
