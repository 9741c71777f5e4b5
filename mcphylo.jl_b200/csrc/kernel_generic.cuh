// kernel_generic.cuh — kernel 2b: column-per-thread walk on dense branch tables -- the runtime-K fallback for large
// alphabets and, with the moment matrices, the kernel of mcp_eval_model_gradient (compile-time K for K <= 6).
// Part of libmcphylo_b200.so; included by walk_generic.cu only (one translation unit).
#pragma once

namespace {

// --------------------------------------------------------------------------------------------
// kernel 2b: generic state count (6 < K <= KMAX_GENERIC), runtime K.  Same op program, same scratch
// layout and accumulator rows as the templated kernel, but dense-table arithmetic straight from the
// branch table (P / dP columns are stored for every branch) and per-thread vectors in local memory.
// Correctness path for large alphabets (e.g. 20-state protein models); not tuned.
//
// Model-gradient evaluations (mcp_eval_model_gradient, any K; mg != nullptr, one tree): next to logL and the
// branch gradient the gradient pass accumulates, for every branch b and rate category r, the K x K moment matrix
//     M[b][r][s][k] = sum over the columns of category r of  q_b[s] * L_b[k] / den,
// q_b = outer partial at the top of branch b (pre[mother] times the siblings' Down), L_b = the stored partial
// below it, den = the column likelihood -- i.e. d logL / d P_{b,r}[s][k] -- and the root vector
//     W[s] = sum over all columns of  L_root[s] / (pi . L_root)                 = d logL / d pi[s] at the root.
// Every derivative of logL with respect to a parameter of the substitution model is a contraction of these
// with d P_{b,r} / d theta (host side, model_gradient_contract in mcphylo_b200.cu).  Per op and child a warp
// leaves its 32 columns' q / den and L in shared memory and forms the 32-column sum of outer products
// cooperatively (K * K outputs spread over the lanes, no shuffles), then adds it to M with one atomic per output.
// mg layout: [device branch][rate][s * K + k], then W[K].
// --------------------------------------------------------------------------------------------
// Every CTA of the launch adds to the same few addresses at about the same time (all walk the ops in the same order), so
// the matrices are kept in mg_rep replicas (CTA c uses replica c % mg_rep) that the host folds with one small kernel.
// KT > 0 fixes the state count at compile time (model-gradient evaluations of K <= 6: the per-thread vectors then live
// in registers and every loop over states unrolls -- the runtime-K instantiation keeps them in local memory and is
// 4x slower at K = 4); KT == 0 is the runtime-K kernel for 6 < K <= KMAX_GENERIC.
template <int KT>
__global__ void __launch_bounds__(128) felsenstein_walk_generic(const WalkParams p, const int K_rt, double* __restrict__ const mg_base,
                                                                const int mg_rep, const long long mg_stride) {
    constexpr int KA = KT > 0 ? KT : KMAX_GENERIC;      // capacity of the per-thread vectors
    const int K = KT > 0 ? KT : K_rt;
    double* const mg = mg_base ? mg_base + (long long)(blockIdx.x % (unsigned)mg_rep) * mg_stride : nullptr;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    __shared__ long long s_e[8];
    __shared__ double s_l[8];
    double* const s_acc = reinterpret_cast<double*>(smem_raw);
    // model gradient: per warp [32 lanes][K] q / den and [32 lanes][K] L, behind the gradient accumulator
    double* const s_mq = reinterpret_cast<double*>(smem_raw + (((size_t)p.max_br * 8 + 15) & ~(size_t)15)) +
                         (size_t)(threadIdx.x >> 5) * 64 * K;
    double* const s_ml = s_mq + 32 * K;

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
            // Op descriptors and the state codes of leaf children are requested ONE OP AHEAD: a leaf child costs two dependent
            // loads (code, then the table column it selects) and the code comes from DRAM -- ncu on the first version
            // showed the warps waiting on exactly these (long scoreboard 5.9 per issue, a third of it on min(code, K)).
            // raw_code returns the byte as loaded; min(., K) is applied where the code is used, one op later.
            auto raw_code = [&](int src) -> int {
                return (valid && src >= 0) ? (int)__ldg(codes + (long long)src * tr.code_stride) : K;
            };
            int4 n0 = make_int4(0, 0, 0, 0), n1 = n0;
            int ca_n = K, cb_n = K;
            auto request = [&](const int4* ops, int i) {
                n0 = __ldg(ops + 2 * i);
                n1 = __ldg(ops + 2 * i + 1);
                ca_n = (n1.y & 3) == mcp::OPK_LEAF ? raw_code(n0.x) : K;
                cb_n = ((n1.y >> 2) & 3) == mcp::OPK_LEAF ? raw_code(n0.z) : K;
            };

            double cur[KA], Da[KA], Db[KA], L[KA];
            for (int k = 0; k < K; ++k) cur[k] = 1.0;
            int e_col = 0;
            if (tr.n_post > 0) request(post_ops, 0);
            for (int i = 0; i < tr.n_post; ++i) {
                const int4 o0 = n0, o1 = n1;
                const int ca = min(ca_n, K), cb = min(cb_n, K);
                if (i + 1 < tr.n_post) request(post_ops, i + 1);
                const int flags = o1.y, ka = flags & 3, kb = (flags >> 2) & 3;
                const double* ta = tab_r + o0.y * br_stride;
                const double* tb = tab_r + o0.w * br_stride;
                if (ka == mcp::OPK_LEAF) {
                    const double* col = ta + ca * K;
                    for (int s = 0; s < K; ++s) Da[s] = __ldg(col + s);
                } else if (ka == mcp::OPK_REG) {
                    down(ta, cur, Da);
                } else {
                    for (int k = 0; k < K; ++k) L[k] = __ldcg(slots + o0.x * slot_stride + k);
                    down(ta, L, Da);
                }
                if (kb == mcp::OPK_LEAF) {
                    const double* col = tb + cb * K;
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
                if (mg) {   // W[s] += L_root[s] / (pi . L_root)
                    double* const w_dst = mg + (long long)tr.n_br * R * K * K;
                    const double w = valid ? 1.0 / rootv : 0.0;
                    for (int k = 0; k < K; ++k) {
                        double v = valid ? cur[k] * w : 0.0;
                        for (int off = 16; off > 0; off >>= 1) v += __shfl_xor_sync(0xffffffffu, v, off);
                        if (lane == 0) atomicAdd(w_dst + k, v);
                    }
                }
            }

            if (p.want_grad) {
                double pm[KA], Ya[KA], Yb[KA];
                // M[br][r] += sum over this warp's 32 columns of (qv * w) (x) Lv; Lv == nullptr: leaf with state `code`
                auto moments = [&](const double* qv, const double* Lv, int code, double w, int br) {
                    for (int k = 0; k < K; ++k) {
                        s_mq[lane * K + k] = valid ? qv[k] * w : 0.0;
                        s_ml[lane * K + k] = Lv ? Lv[k] : ((code >= K || code == k) ? 1.0 : 0.0);
                    }
                    __syncwarp();
                    double* const dst = mg + ((long long)br * R + r) * K * K;
                    for (int idx = lane; idx < K * K; idx += 32) {
                        const int s = idx / K, k = idx - s * K;
                        double acc = 0.0;
                        for (int l = 0; l < 32; ++l) acc = fma(s_mq[l * K + s], s_ml[l * K + k], acc);
#ifdef MCP_MG_NO_ATOMIC      // measurement builds only: the cost of the atomics (results are wrong)
                        if (acc == 123.456) atomicAdd(dst + idx, acc);
#else
                        atomicAdd(dst + idx, acc);
#endif
                    }
                    __syncwarp();
                };
                double La[KA], Lb[KA];
                if (tr.n_pre > 0) request(pre_ops, 0);
                for (int i = 0; i < tr.n_pre; ++i) {
                    const int4 o0 = n0, o1 = n1;
                    const int ca = min(ca_n, K), cb = min(cb_n, K);
                    if (i + 1 < tr.n_pre) request(pre_ops, i + 1);
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
                        for (int k = 0; k < K; ++k) La[k] = __ldcg(slots + o0.x * slot_stride + k);
                        down(ta, La, Da);
                        down(ta + KK1, La, Ya);
                    } else {
                        const double* col = ta + ca * K;
                        for (int s = 0; s < K; ++s) { Da[s] = __ldg(col + s); Ya[s] = __ldg(col + KK1 + s); }
                    }
                    if (bi) {
                        for (int k = 0; k < K; ++k) Lb[k] = __ldcg(slots + o0.z * slot_stride + k);
                        down(tb, Lb, Db);
                        down(tb + KK1, Lb, Yb);
                    } else {
                        const double* col = tb + cb * K;
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
                    if (mg) {   // Ya / Yb hold qa / qb here
                        moments(Ya, ai ? La : nullptr, ca, inv, a_br);
                        moments(Yb, bi ? Lb : nullptr, cb, inv, b_br);
                    }

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

}  // namespace
