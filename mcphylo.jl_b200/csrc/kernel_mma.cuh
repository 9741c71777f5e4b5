// kernel_mma.cuh — kernel 2c: tile-cooperative walk for large alphabets (6 < K <= 32, e.g. 20-state
// protein models) on the FP64 tensor path (mma.sync.m8n8k4.f64; tcgen05 has no fp64).
// Part of libmcphylo_b200.so; compiled in walk_generic.cu.
//
// The contraction of the reference's comb_sum_product_loop!
// (/root/reference/src/Likelihood/VectorizedFunctions.jl:187-213), out[s] = sum_j P[s][j] L[j] per column, is a
// real dense product at K = 20: (K x K) by (K x columns).  One WARP owns 16 columns and walks the same
// depth-first op program as the K <= 6 kernel (schedule.hpp); a CTA of 8 warps is a tile of 128 columns and
// shares the per-op transition tables, staged in shared memory one op ahead with cp.async.
//
// Register layout of every K-vector (partials, products, pre vectors) -- ONE layout for everything, chosen
// so that the accumulator fragment of one product IS the A fragment of the next, no transposition:
//   v[m][n][h], lane = 4 g + t:  column 8 m + g (m < 2),  state 8 n + 2 t + h (n < KP / 8, h < 2)
// i.e. the four lanes of a quad share a column and hold KP / 4 of its states each (KP = K rounded up to 8;
// padded states are identically zero).  With columns as the M dimension of D = A B:
//   P L      A[g][t] = v[m][kb >> 1][kb & 1]   (the lane's own value: state j = 8 (kb >> 1) + 2 t + (kb & 1))
//            B[t][g] = T[j][8 n + g]           (table row j, parent state 8 n + g)        -> D = out[m][n][.]
//   P^T q    same A,  B[t][g] = T[8 n + g][s]  (s = the lane's state as above)            -> D = out[m][n][.]
// The k index of each 8-state group runs over the states in the order (0, 2, 4, 6), (1, 3, 5, 7) -- a
// permutation of a contraction index, applied to A and B alike.
// Shared-memory table T (per child and per P / dP): rows = child state j (KP rows, zero beyond K) plus one
// row for the all-ones leaf (row KP = row sums), columns = parent state s (zero beyond K), row stride KP + 2.
// Gradient evaluations form each product P L ONCE: a post op ends by applying the branch above its node, so what is
// carried in registers and what is stored is D = P L (same table, same operand, hence the same bits as applying it in
// the mother's op); the gradient pass uses a stored child as it is and gets dP L = Q' (P L) from one more product with
// Q' = mu * rate * Q, a table built once per tile -- 2 instead of 3 tensor-core products per internal child there
// (K = 20: 3.70 -> 3.42 ms, K = 12: 1.54 -> 1.39 ms).  logL-only evaluations keep the products in the mother's op,
// where the two children's products are independent of each other (the chained form cost them 9-19 %).
// Partials in HBM scratch are stored fragment-major, [slot][warp][m][n][h][lane]: every access of a warp is a
// run of 32 consecutive doubles.  Gradient sums: one fixed-order butterfly per op, per-warp sums parked in
// shared memory and folded by one thread per branch after the next barrier -- no atomics, bit-reproducible.
#pragma once

namespace {

__device__ __forceinline__ void dmma884(double (&d)[2], double a, double b) {
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
                 : "+d"(d[0]), "+d"(d[1]) : "d"(a), "d"(b));
}
__device__ __forceinline__ void cp_async8(void* smem_dst, const void* gmem_src) {
    unsigned d = (unsigned)__cvta_generic_to_shared(smem_dst);
    asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" :: "r"(d), "l"(gmem_src) : "memory");
}

// out = T L (TRANSPOSED = false) or T^T q (true); T = staged table, rows = child state, columns = parent state;
// g = lane >> 2, t = lane & 3
template <int KP, bool TRANSPOSED>
__device__ __forceinline__ void mma_product(const double* T, const double (&in)[MMA_MB][KP / 8][2], double (&out)[MMA_MB][KP / 8][2],
                                            int g, int t) {
    constexpr int NB = KP / 8, KB = KP / 4, STRIDE = MmaSmem<KP>::STRIDE;
#pragma unroll
    for (int m = 0; m < MMA_MB; ++m)
#pragma unroll
        for (int n = 0; n < NB; ++n) { out[m][n][0] = 0.0; out[m][n][1] = 0.0; }
#pragma unroll
    for (int n = 0; n < NB; ++n)
#pragma unroll
        for (int kb = 0; kb < KB; ++kb) {
            const int j = 8 * (kb >> 1) + 2 * t + (kb & 1);                 // this lane's state in k-block kb
            const double b = TRANSPOSED ? T[(8 * n + g) * STRIDE + j] : T[j * STRIDE + 8 * n + g];
#pragma unroll
            for (int m = 0; m < MMA_MB; ++m) dmma884(out[m][n], in[m][kb >> 1][kb & 1], b);
        }
}

// DST: stored / carried post results are D = P L (gradient evaluations) or the partials L themselves (logL only); two
// instantiations, because with both post ops in one kernel the logL-only path lost 5-17 % to the other's registers.
template <int KP, bool DST>
__global__ void __launch_bounds__(MMA_WARPS * 32, KP <= 16 ? 2 : 1) felsenstein_walk_mma(const __grid_constant__ WalkParams p, const int K) {
    constexpr int NB = KP / 8, KB = KP / 4, MB = MMA_MB, STRIDE = MmaSmem<KP>::STRIDE, TAB = MmaSmem<KP>::TAB;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    __shared__ long long s_e[MMA_WARPS];
    __shared__ double s_l[MMA_WARPS];

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, g = lane >> 2, t = lane & 3;
    const int q = p.n_tiles / gridDim.x, rem = p.n_tiles - q * gridDim.x;
    int tile = blockIdx.x * q + min((int)blockIdx.x, rem);
    const int tile_end = tile + q + ((int)blockIdx.x < rem ? 1 : 0);
    if (tile >= tile_end) return;

    double* const s_acc = reinterpret_cast<double*>(smem_raw);
    unsigned char* sp = smem_raw + MmaSmem<KP>::acc_bytes(p.max_br, p.want_grad);
    double* const s_part = reinterpret_cast<double*>(sp);                       // [2][2 children][8 warps]
    int* const s_pbr = reinterpret_cast<int*>(s_part + 2 * 2 * MMA_WARPS);      // [2][2]
    sp += (MmaSmem<KP>::part_bytes() + 15) & ~(size_t)15;
    int4* const sdesc = reinterpret_cast<int4*>(sp);                            // ring of 3 descriptors
    sp += MmaSmem<KP>::desc_bytes();
    unsigned char* const scode = sp;                                            // [2 buffers][2 children][128]
    sp += MmaSmem<KP>::code_bytes();
    double* const stab = reinterpret_cast<double*>(sp);                         // [2 buffers][4 tables][TAB]
    double* const stq = stab + 2 * 4 * TAB;                                     // Q' of the tile's (tree, rate)
    // this lane's element 0 of the warp's two prefetched child vectors (gradient pass), laid out like the scratch
    double* const spf = stab + (2 * 4 + 1) * TAB + (size_t)warp * 2 * (MMA_WCOLS * KP) + lane;
    constexpr bool dst = DST;                                                   // stored / carried vectors are D = P L

    // zero the tables once: the padding (states >= K) is never written again
    for (int i = tid; i < (2 * 4 + 1) * TAB; i += blockDim.x) stab[i] = 0.0;

    const int R = p.R;
    const int BT = bt_size(K), KK1 = K * (K + 1);
    const long long slot_stride = (long long)MMA_TILE * KP;                     // doubles per slot of this CTA
    double* const scr = p.scratch + (long long)blockIdx.x * p.scratch_per_cta + (long long)warp * (MMA_WCOLS * KP) + lane;
    double* const stk = scr + (long long)p.n_slots * slot_stride;
    int row = p.cta_row_base[blockIdx.x];

    auto ld_vec = [&](const double* base, double (&v)[MB][NB][2]) {
#pragma unroll
        for (int m = 0; m < MB; ++m)
#pragma unroll
            for (int n = 0; n < NB; ++n)
#pragma unroll
                for (int h = 0; h < 2; ++h) v[m][n][h] = __ldcg(base + ((m * NB + n) * 2 + h) * 32);
    };
    // the same vector from / into the warp's prefetch buffer in shared memory
    auto lds_vec = [&](const double* base, double (&v)[MB][NB][2]) {
#pragma unroll
        for (int m = 0; m < MB; ++m)
#pragma unroll
            for (int n = 0; n < NB; ++n)
#pragma unroll
                for (int h = 0; h < 2; ++h) v[m][n][h] = base[((m * NB + n) * 2 + h) * 32];
    };
    auto prefetch_vec = [&](double* dst, const double* src) {
#pragma unroll
        for (int i = 0; i < MB * NB * 2; ++i) cp_async8(dst + i * 32, src + i * 32);
    };
    auto st_vec = [&](double* base, const double (&v)[MB][NB][2]) {
#pragma unroll
        for (int m = 0; m < MB; ++m)
#pragma unroll
            for (int n = 0; n < NB; ++n)
#pragma unroll
                for (int h = 0; h < 2; ++h) __stcg(base + ((m * NB + n) * 2 + h) * 32, v[m][n][h]);
    };
    // sum over the states of each column (spread over the quad), result in all four lanes
    auto quad_sum = [&](double v) -> double {
        v += __shfl_xor_sync(0xffffffffu, v, 1);
        v += __shfl_xor_sync(0xffffffffu, v, 2);
        return v;
    };
    // multiply each column by the exact power of two that brings its maximum into [1, 2); returns the exponents
    auto rescale = [&](double (&v)[MB][NB][2], int (&ex)[MB]) {
#pragma unroll
        for (int m = 0; m < MB; ++m) {
            unsigned mx = 0;
#pragma unroll
            for (int n = 0; n < NB; ++n)
#pragma unroll
                for (int h = 0; h < 2; ++h) mx = max(mx, (unsigned)__double2hiint(v[m][n][h]) & 0x7fffffffu);
            mx = max(mx, __shfl_xor_sync(0xffffffffu, mx, 1));
            mx = max(mx, __shfl_xor_sync(0xffffffffu, mx, 2));
            const int e = (int)(mx >> 20);
            ex[m] = 0;
            if (e != 0 && e != 0x7ff) {
                const double sc = __hiloint2double((2046 - e) << 20, 0);
#pragma unroll
                for (int n = 0; n < NB; ++n) { v[m][n][0] *= sc; v[m][n][1] *= sc; }
                ex[m] = e - 1023;
            }
        }
    };

    int ti = 0;
    while (ti < p.T - 1 && tile >= p.trees[ti].tile_begin + R * p.trees[ti].tiles_per_rate) ++ti;

    while (tile < tile_end) {
        const TreeDev tr = p.trees[ti];
        const int tree_tile_end = min(tile_end, tr.tile_begin + R * tr.tiles_per_rate);
        if (p.want_grad)
            for (int i = tid; i < tr.n_br; i += blockDim.x) s_acc[i] = 0.0;
        long long e_total = 0;
        double logsum = 0.0;
        const double* const pi = p.dyn + tr.dyn_off + dyn_pi(tr.NN, K, R);
        double piv[NB][2];                                                      // pi in the vector layout
#pragma unroll
        for (int n = 0; n < NB; ++n)
#pragma unroll
            for (int h = 0; h < 2; ++h) piv[n][h] = (8 * n + 2 * t + h < K) ? __ldg(pi + 8 * n + 2 * t + h) : 0.0;
        const int4* const post_ops = p.ops + 2 * tr.post_off;
        const int4* const pre_ops = p.ops + 2 * tr.pre_off;

        for (; tile < tree_tile_end; ++tile) {
            const int local = tile - tr.tile_begin;
            const int stile = local / R, r = local - stile * R;
            const long long site0 = (long long)stile * MMA_TILE;
            bool valid[MB];
#pragma unroll
            for (int m = 0; m < MB; ++m) valid[m] = site0 + warp * MMA_WCOLS + 8 * m + g < tr.S;
            const unsigned char* const codes0 = tr.codes + site0;
            const double* const btab_r = p.btab + tr.btab_off + (long long)r * BT;      // (branch 0, rate r)
            const long long br_stride = (long long)R * BT;

            if (p.want_grad) {
                // Q' = mu rate U diag(D) Uinv in table layout (row = child state j, column = parent state s): dP = Q' P
                __syncthreads();                                    // the previous tile is done with the table
                const double* const dd = p.dyn + tr.dyn_off;
                const double* const U = dd + dyn_U(tr.NN);
                const double* const Dg = dd + dyn_D(tr.NN, K);
                const double* const Ui = dd + dyn_Uinv(tr.NN, K);
                const double scale = __ldg(dd + dyn_mu(tr.NN, K)) * __ldg(dd + dyn_rates(tr.NN, K) + r);
                for (int i = tid; i < K * K; i += blockDim.x) {
                    const int jj = i / K, s = i - jj * K;
                    double acc = 0.0;
                    for (int k = 0; k < K; ++k) acc += (__ldg(U + s + K * k) * __ldg(Dg + k)) * __ldg(Ui + k + K * jj);
                    stq[jj * STRIDE + s] = acc * scale;
                }
            }                                                       // (visible after the barriers of the prologues)
            // ---- staging, one op ahead: descriptor of op j + 2, tables and codes of op j + 1 ----
            auto stage_desc = [&](const int4* ops, int n_ops, int j) {
                if (j < n_ops && tid < 2) cp_async16(sdesc + (j % 3) * 2 + tid, ops + 2 * j + tid);
            };
            auto stage_tables = [&](int n_ops, int j, bool pre) {
                if (j >= n_ops) return;
                const int4 o0 = sdesc[(j % 3) * 2], o1 = sdesc[(j % 3) * 2 + 1];
                const int fl = o1.y;
                double* tb = stab + (size_t)(j & 1) * 4 * TAB;
                unsigned char* cb = scode + (size_t)(j & 1) * 2 * MMA_TILE;
                const int per = (K + 1) * K;                        // doubles of one global table (P or dP columns)
                if (dst && !pre && !(fl & mcp::POST_ROOT)) {
                    // post pass: the P table of the op's OWN branch (word 6 of the op: its device node), table slot 1
                    const double* gt = btab_r + o1.z * br_stride + 2 * K;
                    for (int i = tid; i < per; i += blockDim.x) {
                        const int jj = i / K, s = i - jj * K;
                        cp_async8(tb + (size_t)TAB + (jj == K ? KP : jj) * STRIDE + s, gt + i);
                    }
                }
                for (int ch = 0; ch < 2; ++ch) {
                    const int kind = ch ? ((fl >> 2) & 3) : (fl & 3);
                    const int src = ch ? o0.z : o0.x, br = ch ? o0.w : o0.y;
                    const double* gt = btab_r + br * br_stride + 2 * K;         // P columns, then dP columns
                    // a leaf child needs its table rows (P; gradient pass: and dP), an internal child only P, and
                    // only in the gradient pass (for P^T q): in the post pass it arrives with its branch applied
                    const int n_tab = kind == mcp::OPK_LEAF ? (pre ? 2 : 1) : (pre || !dst ? 1 : 0);
                    for (int i = tid; i < n_tab * per; i += blockDim.x) {
                        const int which = i >= per, e = i - which * per, jj = e / K, s = e - jj * K;
                        cp_async8(tb + (size_t)(ch * 2 + which) * TAB + (jj == K ? KP : jj) * STRIDE + s, gt + which * KK1 + e);
                    }
                    if (kind == mcp::OPK_LEAF && tid < MMA_TILE / 16) {
                        unsigned char* dst = cb + ch * MMA_TILE + tid * 16;
                        if (src >= 0) cp_async16(dst, codes0 + (long long)src * tr.code_stride + tid * 16);
                        else *reinterpret_cast<uint4*>(dst) = make_uint4(0xffffffffu, 0xffffffffu, 0xffffffffu, 0xffffffffu);
                    }
                }
            };
            auto prologue = [&](const int4* ops, int n_ops, bool pre) {
                __syncthreads();                                    // previous pass / tile done with the buffers
                stage_desc(ops, n_ops, 0);
                stage_desc(ops, n_ops, 1);
                cp_async_commit();
                cp_async_wait_all();
                __syncthreads();
                stage_tables(n_ops, 0, pre);
                cp_async_commit();
            };
            auto op_boundary = [&](const int4* ops, int n_ops, int j, bool pre) {
                cp_async_wait_all();
                __syncthreads();                                    // tables of op j, descriptors j and j + 1 visible
                stage_tables(n_ops, j + 1, pre);
                stage_desc(ops, n_ops, j + 2);
                cp_async_commit();
            };
            // leaf child: table column of the column's code (row KP = all-ones leaf)
            auto leaf_vec = [&](const double* T, const unsigned char* cb, double (&v)[MB][NB][2]) {
#pragma unroll
                for (int m = 0; m < MB; ++m) {
                    const int code = cb[warp * MMA_WCOLS + 8 * m + g];
                    const double* rowp = T + (code >= K ? KP : code) * STRIDE + 2 * t;
#pragma unroll
                    for (int n = 0; n < NB; ++n) { v[m][n][0] = rowp[8 * n]; v[m][n][1] = rowp[8 * n + 1]; }
                }
            };

            // ------------------------------ post pass ------------------------------
            double cur[MB][NB][2];
            int e_col[MB];
#pragma unroll
            for (int m = 0; m < MB; ++m) {
                e_col[m] = 0;
#pragma unroll
                for (int n = 0; n < NB; ++n) { cur[m][n][0] = 0.0; cur[m][n][1] = 0.0; }
            }
            {
                const int n_post = tr.n_post;
                prologue(post_ops, n_post, false);
                for (int j = 0; j < n_post; ++j) {
                    op_boundary(post_ops, n_post, j, false);
                    const int4 o0 = sdesc[(j % 3) * 2], o1 = sdesc[(j % 3) * 2 + 1];
                    const int flags = o1.y, ka = flags & 3, kb_ = (flags >> 2) & 3;
                    const double* tb = stab + (size_t)(j & 1) * 4 * TAB;
                    const unsigned char* cb = scode + (size_t)(j & 1) * 2 * MMA_TILE;
                    double Da[MB][NB][2], Db[MB][NB][2];
                    // canonical operand kinds (schedule.hpp): (LEAF, LEAF), (REG, LEAF), (MEM, REG); an internal child's
                    // vector -- `cur` (REG) or stored (MEM) -- already is D = P L
                    if constexpr (dst) {
                        if (ka == mcp::OPK_LEAF) {
                            leaf_vec(tb, cb, Da);
                        } else if (ka == mcp::OPK_MEM) {
                            ld_vec(scr + (long long)o0.x * slot_stride, Da);
                        }
                        if (kb_ == mcp::OPK_LEAF) leaf_vec(tb + 2 * TAB, cb + MMA_TILE, Db);
#pragma unroll
                        for (int m = 0; m < MB; ++m)
#pragma unroll
                            for (int n = 0; n < NB; ++n)
#pragma unroll
                                for (int h = 0; h < 2; ++h) {
                                    const double a = ka == mcp::OPK_REG ? cur[m][n][h] : Da[m][n][h];
                                    const double b = kb_ == mcp::OPK_LEAF ? Db[m][n][h] : cur[m][n][h];
                                    cur[m][n][h] = a * b;
                                }
                    } else {                                                        // logL only: L stored, products here
                        if (ka == mcp::OPK_LEAF) {
                            leaf_vec(tb, cb, Da);
                        } else if (ka == mcp::OPK_REG) {
                            mma_product<KP, false>(tb, cur, Da, g, t);
                        } else {
                            double L[MB][NB][2];
                            ld_vec(scr + (long long)o0.x * slot_stride, L);
                            mma_product<KP, false>(tb, L, Da, g, t);
                        }
                        if (kb_ == mcp::OPK_LEAF) leaf_vec(tb + 2 * TAB, cb + MMA_TILE, Db);
                        else mma_product<KP, false>(tb + 2 * TAB, cur, Db, g, t);   // REG
#pragma unroll
                        for (int m = 0; m < MB; ++m)
#pragma unroll
                            for (int n = 0; n < NB; ++n) { cur[m][n][0] = Da[m][n][0] * Db[m][n][0]; cur[m][n][1] = Da[m][n][1] * Db[m][n][1]; }
                    }
                    int ex[MB];
                    rescale(cur, ex);
#pragma unroll
                    for (int m = 0; m < MB; ++m) e_col[m] += ex[m];
                    if (dst && !(flags & mcp::POST_ROOT)) {                         // the branch above this node
                        mma_product<KP, false>(tb + TAB, cur, Da, g, t);
#pragma unroll
                        for (int m = 0; m < MB; ++m)
#pragma unroll
                            for (int n = 0; n < NB; ++n) { cur[m][n][0] = Da[m][n][0]; cur[m][n][1] = Da[m][n][1]; }
                    }
                    if (flags & mcp::POST_STORE) st_vec(scr + (long long)o1.x * slot_stride, cur);
                }
            }
#pragma unroll
            for (int m = 0; m < MB; ++m) {
                double rootv = 0.0;
#pragma unroll
                for (int n = 0; n < NB; ++n) rootv = fma(piv[n][0], cur[m][n][0], fma(piv[n][1], cur[m][n][1], rootv));
                rootv = quad_sum(rootv);
                if (valid[m] && t == 0) {
                    logsum += log(rootv);
                    e_total += e_col[m];
                }
            }

            // ------------------------------ gradient pass ------------------------------
            if (p.want_grad) {
                const int n_pre = tr.n_pre;
                prologue(pre_ops, n_pre, true);
#pragma unroll
                for (int m = 0; m < MB; ++m)
#pragma unroll
                    for (int n = 0; n < NB; ++n) { cur[m][n][0] = piv[n][0]; cur[m][n][1] = piv[n][1]; }   // the root's pre vector is pi
                auto fold_parked = [&](int buf) {
                    if (tid < 2) {
                        const double* pp = s_part + (buf * 2 + tid) * MMA_WARPS;
                        double sum = pp[0];
                        for (int w = 1; w < MMA_WARPS; ++w) sum += pp[w];
                        s_acc[s_pbr[buf * 2 + tid]] += sum;
                    }
                };
                for (int j = 0; j < n_pre; ++j) {
                    op_boundary(pre_ops, n_pre, j, true);
                    if (j > 0) fold_parked((j - 1) & 1);                        // complete since the barrier above
                    const int4 o0 = sdesc[(j % 3) * 2], o1 = sdesc[(j % 3) * 2 + 1];
                    const int flags = o1.y;
                    const bool ai = (flags & 3) == mcp::OPK_MEM, bi = ((flags >> 2) & 3) == mcp::OPK_MEM;
                    const double* tb = stab + (size_t)(j & 1) * 4 * TAB;
                    const unsigned char* cb = scode + (size_t)(j & 1) * 2 * MMA_TILE;
                    if (((flags >> 8) & 3) == mcp::PREM_STACK) ld_vec(stk + (long long)o1.x * slot_stride, cur);
                    // family {mother; a, b}: D = P L, Y = dP L
                    double Da[MB][NB][2], Ya[MB][NB][2], Db[MB][NB][2], Yb[MB][NB][2];
                    // The stored child vectors of family j were requested during family j - 1 (below) and are complete
                    // since the cp.async wait of the op boundary; the first family reads its own.
                    if (ai) {                                                   // stored D = P L; dP L = Q' D
                        if (j > 0) lds_vec(spf, Da); else ld_vec(scr + (long long)o0.x * slot_stride, Da);
                        mma_product<KP, false>(stq, Da, Ya, g, t);
                    } else {
                        leaf_vec(tb, cb, Da);
                        leaf_vec(tb + TAB, cb, Ya);
                    }
                    if (bi) {
                        if (j > 0) lds_vec(spf + MMA_WCOLS * KP, Db); else ld_vec(scr + (long long)o0.z * slot_stride, Db);
                        mma_product<KP, false>(stq, Db, Yb, g, t);
                    } else {
                        leaf_vec(tb + 2 * TAB, cb + MMA_TILE, Db);
                        leaf_vec(tb + 3 * TAB, cb + MMA_TILE, Yb);
                    }
                    if (j + 1 < n_pre) {
                        // request the stored children of the NEXT family (its descriptor is resident): post results, written
                        // a whole pass ago.  pre[mother] is not requested ahead -- this very op may be the one that pushes it.
                        const int4 n0 = sdesc[((j + 1) % 3) * 2], n1 = sdesc[((j + 1) % 3) * 2 + 1];
                        __syncwarp();                                            // every lane has read its part of the buffer
                        if ((n1.y & 3) == mcp::OPK_MEM) prefetch_vec(spf, scr + (long long)n0.x * slot_stride);
                        if (((n1.y >> 2) & 3) == mcp::OPK_MEM) prefetch_vec(spf + MMA_WCOLS * KP, scr + (long long)n0.z * slot_stride);
                    }
                    // qa = pre_m * Db, qb = pre_m * Da (kept in Db / Da), den = sum pre_m Da Db, numerators q . Y
                    double ga = 0.0, gb = 0.0;
#pragma unroll
                    for (int m = 0; m < MB; ++m) {
                        double den = 0.0, na = 0.0, nb = 0.0;
#pragma unroll
                        for (int n = 0; n < NB; ++n)
#pragma unroll
                            for (int h = 0; h < 2; ++h) {
                                const double qa = cur[m][n][h] * Db[m][n][h], qb = cur[m][n][h] * Da[m][n][h];
                                den = fma(qa, Da[m][n][h], den);
                                na = fma(qa, Ya[m][n][h], na);
                                nb = fma(qb, Yb[m][n][h], nb);
                                Db[m][n][h] = qa;
                                Da[m][n][h] = qb;
                            }
                        den = quad_sum(den);
                        na = quad_sum(na);
                        nb = quad_sum(nb);
                        if (valid[m] && t == 0) {
                            const double inv = 1.0 / den;
                            ga = fma(na, inv, ga);
                            gb = fma(nb, inv, gb);
                        }
                    }
                    // pre vectors of internal children: P_a^T qa (qa sits in Db), P_b^T qb (qb sits in Da)
                    const int a_out = (flags >> 10) & 3, b_out = (flags >> 12) & 3;
                    if (b_out != mcp::OUT_NONE) {
                        double pb[MB][NB][2];
                        int ex[MB];
                        mma_product<KP, true>(tb + 2 * TAB, Da, pb, g, t);
                        rescale(pb, ex);
                        st_vec(stk + (long long)o1.w * slot_stride, pb);       // canonical: b is pushed
                    }
                    if (a_out != mcp::OUT_NONE) {
                        int ex[MB];
                        mma_product<KP, true>(tb, Db, cur, g, t);                // canonical: a is kept in registers
                        rescale(cur, ex);
                    }
                    const double red = warp_pair_reduce(ga, gb, lane);         // lane 0: sum of ga, lane 16: sum of gb
                    if ((lane & 15) == 0) {
                        const int term = (j & 1) * 2 + (lane >> 4);
                        s_part[term * MMA_WARPS + warp] = red;
                        if (warp == 0) s_pbr[term] = (lane >> 4) ? o0.w : o0.y;
                    }
                }
                __syncthreads();                                               // the last op's sums are parked
                if (n_pre > 0) fold_parked((n_pre - 1) & 1);
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
            for (int w = 0; w < MMA_WARPS; ++w) { es += s_e[w]; ls += s_l[w]; }
            p.rows_ll[row].esum = es;
            p.rows_ll[row].logsum = ls;
        }
        if (p.want_grad) {
            double* dst = p.rows + (long long)row * p.row_stride;
            for (int i = tid; i < tr.n_br; i += blockDim.x) dst[i] = s_acc[i];
        }
        __syncthreads();
        ++row;
        ++ti;
    }
}

}  // namespace
