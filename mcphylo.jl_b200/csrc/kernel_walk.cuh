// kernel_walk.cuh — kernel 2: the persistent fused walk (post pass + gradient pass), one thread per column.
// Part of libmcphylo_b200.so; instantiated per state count K in walk_k*.cu (walk_inst.cuh).
#pragma once

namespace {

// --------------------------------------------------------------------------------------------
// kernel 2: the fused walk.  grid = persistent CTAs, each takes a contiguous range of column
// tiles; a tile = blockDim.x columns of one rate category of one tree.
//
// Per-op inputs that are uniform over the tile or byte-sized per column are staged in shared
// memory CHN ops at a time with cp.async (16, ring kernels 8), one chunk ahead of the compute:
//   sdesc  3 x CHN op descriptors (ring of 3: descriptors must be resident one chunk before the
//          data they describe can be requested)
//   se     2 x CHN x 2 x 2K doubles: (em1, de) eigen-coefficient vectors (gradient pass: of the internal
//          children; post pass with DST: of the op's own branch)
//   stab   2 x CHN x 2 x 2 leaf tables: P (and dP) columns of LEAF children
//   scode  2 x CHN x 2 x 2 x TS bytes: the state codes of LEAF children for the tile's columns
// so the only global accesses on the per-op critical path are the thread's own partials -- and with an
// operand ring (RD > 0) those of the gradient pass arrive through shared memory as well.  One
// __syncthreads per chunk.
// --------------------------------------------------------------------------------------------
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
// NE = active eigen-components (device_math.cuh): K - 1 when the host found (and moved last) a null
// eigenvalue, the case for every rate matrix; K otherwise.
// RD = depth of the per-warp OPERAND RING of the gradient pass (0 = none): the stored child partials of the
// families ahead are fetched into shared memory RD entries before their use, in the order of the tree's fetch
// list (schedule.hpp: pre_fetch), so that the warp never waits a DRAM round trip on its own partials.  Two
// flavours (smem_layout.cuh: MCP_RING_LDGSTS): every thread copies its own 16-byte pieces with cp.async
// (completion by cp.async.wait_group; the default), or one elected lane per warp issues a bulk copy
// (cp.async.bulk, completion on an mbarrier).  profiles/r2_walk_notes.md has the measurements.
template <int K, int CPT, bool DYN_MODEL, bool SSCR, int NE, bool ACCG, int RD = 0>
#ifndef MCP_WALK_MAXT
#define MCP_WALK_MAXT 256
#endif
#ifndef MCP_EARLY_LOADS
#define MCP_EARLY_LOADS (CPT == 1)   // stored operands of op j+1 are requested during op j (measured: +3 % at one column per thread; with two the pinned destination registers cost spills)
#endif
__global__ void __launch_bounds__(MCP_WALK_MAXT, K * CPT <= 4 ? MCP_WALK_MIN_BLOCKS : MCP_WALK_MIN_BLOCKS2) felsenstein_walk(const __grid_constant__ WalkParams p) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    constexpr int CHN = RD ? WALK_RING_CH : CH;       // ops staged per chunk
    static_assert(RD == 0 || (!SSCR && !ACCG && !DYN_MODEL), "the operand ring exists for the HBM-scratch, shared-accumulator, single-model kernels");
    __shared__ long long s_e[8];
    __shared__ double s_l[8];

    const int tid = threadIdx.x, TW = blockDim.x, lane = tid & 31, warp = tid >> 5;
    const int TS = TW * CPT;                          // sites per tile; column c of a thread = site0 + c*TW + tid
    // Tiles: a static contiguous range per CTA, or (streamed evaluation, one tree) an atomic ticket per tile
    // so that CTAs whose data arrives early simply do more of the work.
    __shared__ int s_ticket;
    const bool dynamic_tiles = p.ticket != nullptr;
    auto next_ticket = [&]() -> int {
        __syncthreads();                                  // everybody has read the previous ticket
        if (tid == 0) s_ticket = (int)min(atomicAdd(p.ticket, 1u), 0x7fffffffu);
        __syncthreads();
        return s_ticket;
    };
    const int q = p.n_tiles / gridDim.x, rem = p.n_tiles - q * gridDim.x;
    int tile = dynamic_tiles ? next_ticket() : blockIdx.x * q + min((int)blockIdx.x, rem);
    const int tile_end = dynamic_tiles ? p.n_tiles : tile + q + ((int)blockIdx.x < rem ? 1 : 0);
    if (tile >= tile_end) {
        if (dynamic_tiles) {   // this CTA got no tile: its accumulator row must still read as zero
            const int row0 = p.cta_row_base[blockIdx.x];
            if (tid == 0) { p.rows_ll[row0].esum = 0; p.rows_ll[row0].logsum = 0.0; }
            if (p.want_grad)
                for (int i = tid; i < p.max_br; i += TW) p.rows[(long long)row0 * p.row_stride + i] = 0.0;
        }
        return;
    }

    double* const s_acc = reinterpret_cast<double*>(smem_raw);
    constexpr bool GL2 = ACCG;
    int4* const sdesc = reinterpret_cast<int4*>(smem_raw + WalkSmem<K, CHN>::acc_bytes(p.max_br, p.want_grad && !ACCG));
    // parked per-warp branch sums + their branch ids sit right below the descriptors (see walk_part_bytes(CHN))
    double* const s_part = reinterpret_cast<double*>(reinterpret_cast<unsigned char*>(sdesc) - walk_part_bytes(CHN));
    int* const s_pbr = reinterpret_cast<int*>(s_part + walk_part_doubles(CHN));
    // folds the parked sums of one chunk (buffer `buf`, `n` ops) into s_acc; call after the barrier that
    // follows the chunk, by all threads (2 * n of them do the work; each term has its own branch)
    // The once-per-chunk jobs of a few threads -- this fold, the derivation of the op records -- are given to
    // DIFFERENT warps (the last one, the one before it): every warp waits at the chunk barrier for the slowest.
    const int fold_t = tid - ((TW >> 5) - 1) * 32;                       // lane of the folding warp, negative elsewhere
    const int rec_t = tid - ((TW >> 5) > 1 ? (TW >> 5) - 2 : 0) * 32;    // lane of the record-deriving warp
    auto fold_parked = [&](int buf, int n) {
        if (fold_t >= 0 && fold_t < 2 * n) {
            const double* pp = s_part + (buf * CHN * 2 + fold_t) * 8;
            double sum = pp[0];
            for (int w = 1; w < (TW >> 5); ++w) sum += pp[w];
            s_acc[s_pbr[buf * CHN * 2 + fold_t]] += sum;
        }
    };
    double* const se = reinterpret_cast<double*>(reinterpret_cast<unsigned char*>(sdesc) + WalkSmem<K, CHN>::desc_bytes());
    OpRec* const srec = reinterpret_cast<OpRec*>(reinterpret_cast<unsigned char*>(se) + WalkSmem<K, CHN>::e_bytes());
    double* const stab = reinterpret_cast<double*>(reinterpret_cast<unsigned char*>(srec) + WalkSmem<K, CHN>::rec_bytes());
    unsigned char* const scode = reinterpret_cast<unsigned char*>(stab) + WalkSmem<K, CHN>::tab_bytes();
    constexpr int KK1 = K * (K + 1);                  // doubles of one leaf table (P or dP columns)

    // per-thread base of the CTA-private scratch, laid out [slot][column c][thread][state] -- with an operand ring
    // [slot][warp][column c][lane][state], so that a warp's share of a slot is ONE contiguous run, fetched by one
    // bulk copy; all slot / LIFO offsets in the records are byte offsets from here
    constexpr unsigned WB = 32 * K * 8;               // bytes of one warp's share of one column vector of a slot
    // RLS: cp.async flavour of the ring -- a warp's share of a column is split into 16-byte pieces [piece][lane]
    // (device_math.cuh: piece layout), every thread copies its own pieces
    constexpr bool RLS = RD > 0 && WALK_RING_LDGSTS;
    // DST: what a stored (or register-carried) post result is.  false: the node's partial L, and every op applies the
    // children's branches (P L in the mother's post op and AGAIN in the family's gradient op).  true (K >= 3): D = P L --
    // a post op ends by applying the branch above its node, so the product is formed once; the gradient pass uses the
    // stored D as it is and takes the numerator's eigen-coordinates from it (dP = U diag(c) Uinv P).  Same bits in the
    // post pass either way; 15 of ~130 FP64 operations per node and column less at K = 4 (cfg4 -1.4 %, cfg3 -3.8 %).
    // At K = 2 the products are too small to matter and the longer dependency chain of the post op costs 2.6 %.
    constexpr bool DST = K >= 3;
    unsigned char* const scr = SSCR
        ? scode + WalkSmem<K, CHN>::code_bytes(TS, 2) + (size_t)tid * K * 8
        : keep_ptr(reinterpret_cast<unsigned char*>(p.scratch + (long long)blockIdx.x * p.scratch_per_cta) +
                   (RLS ? (long long)warp * (CPT * WB) + lane * 16
                        : RD > 0 ? ((long long)(warp * (CPT * 32) + lane) * K) * 8 : ((long long)tid * K) * 8));
    const unsigned col_bytes = RD > 0 ? WB : (unsigned)TW * K * 8;  // distance between a thread's columns within a slot
    const unsigned slot_bytes = (unsigned)TW * K * 8 * CPT;
    const unsigned stack_base = (unsigned)p.n_slots * slot_bytes;
    // ---- operand ring (RD > 0): this warp's stages and their mbarriers, as shared-window addresses ----
    constexpr unsigned RBAR = RD * CPT * WB;          // this warp's block: RD stages of CPT vectors, then RD mbarriers
    // Ring state is kept in ordinary (per-lane) registers on purpose -- it is warp-uniform, but values ptxas takes for
    // uniform are moved to uniform registers right behind their loads (a stall on the load) and compete with U / Uinv
    // there; `lane0` is a zero the compiler cannot see through and takes for lane-dependent.
    //   ring_l   this LANE's vector in stage 0, column 0 (lane 0: the warp's block itself); stage s, column c at
    //            + (s * CPT + c) * WB
    //   rbar_w   the warp's mbarrier of stage 0; stage s at + 8 s
    //   r_cnt    entries consumed since the kernel started: stage = r_cnt % RD, mbarrier parity = (r_cnt / RD) & 1
    unsigned ring_l = 0, rbar_w = 0, r_cnt = 0, lane0 = 0;
    if constexpr (RD > 0) {
        static_assert((RD & (RD - 1)) == 0, "ring depth must be a power of two");
        asm volatile("and.b32 %0, %1, 0;" : "=r"(lane0) : "r"(tid));
        const size_t roff = WalkSmem<K, CHN>::ring_offset(p.max_br, 1, TS);
        const unsigned base = (unsigned)__cvta_generic_to_shared(smem_raw + roff) + (unsigned)warp * (RBAR + RD * 8);
        if constexpr (!RLS) {
            if (lane == 0) {
#pragma unroll
                for (int st = 0; st < RD; ++st) mbar_init(base + RBAR + st * 8, 1);
            }
            fence_mbar_init();                        // ordered before the first use by the barriers of the first prologue
        }
        ring_l = keep_u32(base + lane * (RLS ? 16 : K * 8));
        rbar_w = keep_u32(base + RBAR + lane0);
        r_cnt = RLS ? 0u : lane0;
    }
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

        for (; tile < tree_tile_end; tile = dynamic_tiles ? next_ticket() : tile + 1) {
            const int local = tile - tr.tile_begin;
            // Site-major tile order: the R rate categories of one site range are consecutive tiles, so a CTA (or
            // its neighbours) re-reads that range's codes from L2 while they are hot, and the tickets of a
            // streamed evaluation follow the arrival of the sites.
            const int stile = local / R, r = local - stile * R;
            const long long site0 = (long long)stile * TS;
            if (p.ready_flags != nullptr && tid == 0) {
                // wait until the copy stream has marked the last site of this tile as landed (sites arrive in
                // order); the other threads wait at the barrier that opens the post pass
                const long long last = min(site0 + TS, tr.S) - 1;
                const unsigned* flag = p.ready_flags + (last >> p.ready_shift);
                unsigned seen, spins = 0;
                for (;;) {
                    asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(seen) : "l"(flag) : "memory");
                    if (seen == p.ready_epoch) break;
                    __nanosleep(200);
                    if (++spins > (1u << 22)) { atomicExch(p.error_flag, 1u); break; }   // seconds: give up, the host reports it
                }
            }
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
                const int base = c * CHN, cnt = min(CHN, n_ops - base);
                int4* dst = sdesc + (c % 3) * (CHN * 2);
                for (int i = tid; i < cnt * 2; i += TW) cp_async16(dst + i, ops + 2 * base + i);
            };
            // Copies e vectors / leaf codes / leaf tables of chunk c and derives the per-op records (byte
            // offsets), once per CTA instead of once per warp.  `pre` selects the pre-program field meaning.
            // Work split: one (op, child) pair per warp at a time, the lanes share the pair's 16-byte pieces --
            // the operand kind is warp-uniform and every kind fetches exactly what it needs:
            //   LEAF    its code row segment + the branch's P (gradient pass: and dP) columns
            //   CHERRY  (gradient pass) the code rows and P columns of BOTH leaves below it + its own (em1, de)
            //   REG/MEM (em1, de) of the branch
            auto stage_data = [&](int n_ops, int c, bool pre) {
                const int base = c * CHN, cnt = min(CHN, n_ops - base);
                const int4* d = sdesc + (c % 3) * (CHN * 2);
                double* eb = se + (c & 1) * (CHN * 2 * 2 * K);
                const int crows = pre ? 2 : 1;               // code rows per child slot
                unsigned char* cb = scode + (size_t)(c & 1) * (CHN * 2 * 2 * TS);
                OpRec* rb = srec + (c & 1) * CHN;
                double* tb = stab + (size_t)(c & 1) * (CHN * 2 * 2 * KK1);
                const int pieces = TS / 16;                  // 16-byte pieces of one code row segment
                const int tpieces = KK1 / 2;                 // 16-byte pieces of one leaf table (K (K + 1) is even)
                auto copy_codes = [&](unsigned char* dstrow, int row) {
                    for (int piece = lane; piece < pieces; piece += 32) {
                        unsigned char* dstp = dstrow + piece * 16;
                        if (row >= 0) cp_async16(dstp, codes0 + (long long)row * tr.code_stride + piece * 16);
                        else *reinterpret_cast<uint4*>(dstp) = make_uint4(0x01010101u * K, 0x01010101u * K, 0x01010101u * K, 0x01010101u * K);
                    }
                };
                auto copy_table = [&](double* dst, const double* src, int n_tab) {   // n_tab tables of KK1 doubles
                    for (int tp = lane; tp < n_tab * tpieces; tp += 32) {
                        if constexpr ((K * 8) % 16 == 0) {
                            cp_async16(dst + tp * 2, src + tp * 2);
                        } else {
                            dst[tp * 2] = __ldg(src + tp * 2);
                            dst[tp * 2 + 1] = __ldg(src + tp * 2 + 1);
                        }
                    }
                };
                for (int jc = warp; jc < cnt * 2; jc += (TW >> 5)) {
                    const int j = jc >> 1, ch = jc & 1;
                    const int4 o0 = d[2 * j], o1 = d[2 * j + 1];
                    const int fl = o1.y;
                    const int kind = ch ? ((fl >> 2) & 3) : (fl & 3);
                    const int src = ch ? o0.z : o0.x, br = ch ? o0.w : o0.y;
                    const double* bsrc = reinterpret_cast<const double*>(btab_b + (unsigned)br * br_bytes);
                    unsigned char* crow = cb + (size_t)(j * 2 + ch) * crows * TS;
                    double* tdst = tb + (size_t)(j * 2 + ch) * 2 * KK1;
                    if (kind == mcp::OPK_LEAF) {
                        copy_codes(crow, src);
                        copy_table(tdst, bsrc + 2 * K, pre ? 2 : 1);
                    }
                    if (DST && !pre) {
                        // post pass: internal children arrive with their own branch already applied (see the post op);
                        // what an op needs is the (em1, de) vector of ITS OWN branch (word 6 of the op: its device
                        // node), staged in the slot of child a -- the root has no branch
                        if (ch == 0 && !(fl & mcp::POST_ROOT) && lane < K)
                            cp_async16(reinterpret_cast<unsigned char*>(eb + (j * 2) * 2 * K) + lane * 16,
                                       btab_b + (unsigned)o1.z * br_bytes + lane * 16);
                    } else if (kind != mcp::OPK_LEAF) {
                        if (pre && kind == mcp::OPK_CHERRY) {    // child a only: leaves in a_src, their branches in a_dst
                            const int bx = o1.z & 0xffff, by = (o1.z >> 16) & 0xffff;
                            copy_codes(crow, src & 0xffff);
                            copy_codes(crow + TS, (src >> 16) & 0xffff);
                            copy_table(tdst, reinterpret_cast<const double*>(btab_b + (unsigned)bx * br_bytes) + 2 * K, 1);
                            copy_table(tdst + KK1, reinterpret_cast<const double*>(btab_b + (unsigned)by * br_bytes) + 2 * K, 1);
                        }
                        // (em1, de): 2K doubles = K 16-byte pieces; entries are 16-byte aligned (bt_size is even)
                        if (lane < K)
                            cp_async16(reinterpret_cast<unsigned char*>(eb + (j * 2 + ch) * 2 * K) + lane * 16,
                                       reinterpret_cast<const unsigned char*>(bsrc) + lane * 16);
                    }
                }
                if (rec_t >= 0 && rec_t < cnt) {
                    const int j = rec_t;
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
                if (n_ops > CHN) stage_desc(ops, n_ops, 1);
                cp_async_commit();
                cp_async_wait_all();
                __syncthreads();
                stage_data(n_ops, 0, pre);
                cp_async_commit();
            };
            // cp.async flavour of the ring: its refills are cp.async groups of their own, committed between the staging
            // groups.  ring_k counts the refill groups committed since the last staging group: the staging group (and
            // everything older) is complete once at most min(ring_k, RD) of the newest groups are pending -- the chunk
            // boundary does not wait for operands that were requested for the ops to come.
            int ring_k = 0;
            auto chunk_boundary = [&](const int4* ops, int n_ops, int c, int n_chunks, bool pre) {
                if constexpr (RLS) {
                    static_assert(!RLS || RD == 2, "the boundary wait below is written out for a ring of depth 2");
                    if (ring_k >= 2) cp_async_wait_group<2>();
                    else if (ring_k == 1) cp_async_wait_group<1>();
                    else cp_async_wait_group<0>();
                } else {
                    cp_async_wait_all();
                }
                __syncthreads();                              // chunk c data + descriptors c, c+1 visible
                if (c + 1 < n_chunks) stage_data(n_ops, c + 1, pre);
                if (c + 2 < n_chunks) stage_desc(ops, n_ops, c + 2);
                cp_async_commit();
                ring_k = 0;
            };
            auto ld_cols = [&](unsigned off, double (&v)[CPT][K]) {
#pragma unroll
                for (int c = 0; c < CPT; ++c) {
                    if constexpr (SSCR) {
                        const double* src = reinterpret_cast<const double*>(scr + off + c * col_bytes);
#pragma unroll
                        for (int k = 0; k < K; ++k) v[c][k] = src[k];
                    } else if constexpr (RLS) {
                        ld_partial_pc<K>(scr + off + c * col_bytes, v[c]);
                    } else {
                        ld_partial<K>(reinterpret_cast<const double*>(scr + off + c * col_bytes), v[c]);
                    }
                }
            };
            // keep_old: v untouched when !pred; otherwise v is undefined then (its old contents are dead)
            auto ld_cols_if = [&](bool pred, unsigned off, double (&v)[CPT][K], auto keep_old) {
#pragma unroll
                for (int c = 0; c < CPT; ++c) {
                    if constexpr (SSCR) {
                        const double* src = reinterpret_cast<const double*>(scr + off + c * col_bytes);
                        if (pred) {
#pragma unroll
                            for (int k = 0; k < K; ++k) v[c][k] = src[k];
                        }
                    } else if constexpr (RLS) {
                        ld_partial_pc_if<K, decltype(keep_old)::value>(pred, scr + off + c * col_bytes, v[c]);
                    } else {
                        ld_partial_if<K, decltype(keep_old)::value>(pred, reinterpret_cast<const double*>(scr + off + c * col_bytes), v[c]);
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
                    } else if constexpr (RLS) {
                        st_partial_pc<K>(scr + off + c * col_bytes, v[c]);
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
                const int n_post = tr.n_post, n_chunks = (n_post + CHN - 1) / CHN;
                prologue(post_ops, n_post, false);
                for (int c = 0; c < n_chunks; ++c) {
                    chunk_boundary(post_ops, n_post, c, n_chunks, false);
                    const OpRec* rb = srec + (c & 1) * CHN;
                    const double* eb = se + (c & 1) * (CHN * 2 * 2 * K);
                    const unsigned char* cb = scode + (size_t)(c & 1) * (CHN * 2 * 2 * TS) + tid;
                    const double* tb = stab + (size_t)(c & 1) * (CHN * 2 * 2 * KK1);
                    const int cnt = min(CHN, n_post - c * CHN);
                    double Lm[CPT][K];                                              // the op's stored operand
                    for (int j = 0; j < cnt; ++j) {
                        const uint4 rh = *reinterpret_cast<const uint4*>(rb + j);   // flags, xa, xb, y0
                        const int flags = (int)rh.x, ka = flags & 3;
                        // The stored operand of the NEXT op is requested as soon as Lm is free, so that its
                        // L2 / HBM latency overlaps the rest of this op (the first op of a chunk requests
                        // its own: the next chunk's records only become visible at the chunk barrier).
                        auto request_next = [&]() {
                            if (MCP_EARLY_LOADS) {
                                const uint2 rn = *reinterpret_cast<const uint2*>(rb + (j + 1 < cnt ? j + 1 : j));   // flags, xa
                                ld_cols_if(j + 1 < cnt && ((int)rn.x & 3) == mcp::OPK_MEM, rn.y, Lm, std::false_type{});
                            }
                        };
                        // Canonical operand kinds (schedule.hpp): (LEAF, LEAF), (REG, LEAF), (MEM, REG).  An internal
                        // child's vector -- carried in `cur` (REG) or stored (MEM) -- is D = P L, the child's partial with
                        // its own branch already applied (below), so the op itself is a product and one
                        // matrix-vector product for the branch above this node.
                        double Da[CPT][K], Db[CPT][K];
                        auto leaf_cols = [&](int ch, double (&D)[CPT][K]) {
#pragma unroll
                            for (int cc = 0; cc < CPT; ++cc) {
                                const int code = min((int)cb[(j * 2 + ch) * TS + cc * TW], K);
                                const double* t = tb + (j * 2 + ch) * 2 * KK1 + code * K;
#pragma unroll
                                for (int k = 0; k < K; ++k) D[cc][k] = t[k];
                            }
                        };
                        if constexpr (DST) {
                            if (ka == mcp::OPK_MEM) {
                                ld_cols_if(!MCP_EARLY_LOADS || j == 0, rh.y, Lm, std::true_type{});
#pragma unroll
                                for (int cc = 0; cc < CPT; ++cc)
#pragma unroll
                                    for (int k = 0; k < K; ++k) cur[cc][k] = Lm[cc][k] * cur[cc][k];
                                request_next();
                            } else if (ka == mcp::OPK_REG) {
                                request_next();
                                leaf_cols(1, Db);
#pragma unroll
                                for (int cc = 0; cc < CPT; ++cc)
#pragma unroll
                                    for (int k = 0; k < K; ++k) cur[cc][k] = cur[cc][k] * Db[cc][k];
                            } else {
                                request_next();
                                leaf_cols(0, Da);
                                leaf_cols(1, Db);
#pragma unroll
                                for (int cc = 0; cc < CPT; ++cc)
#pragma unroll
                                    for (int k = 0; k < K; ++k) cur[cc][k] = Da[cc][k] * Db[cc][k];
                            }
#pragma unroll
                            for (int cc = 0; cc < CPT; ++cc) e_col[cc] += rescale_pow2<K>(cur[cc]);
                            // The branch above this node is applied HERE, D = P L = L + U (em1 * (Uinv L)) (same operands,
                            // same order, hence the same bits as applying it in the mother's op): the mother's op and the
                            // gradient pass read D and never repeat this product.  The root has no branch.
                            if (!(flags & mcp::POST_ROOT)) {
                                double e[K], z[CPT][K];
#pragma unroll
                                for (int k = 0; k < K; ++k) e[k] = eb[(j * 2) * 2 * K + k];
                                eig_project<K, CPT, false, NE>(mdl, cur, e, nullptr, z, z);
                                eig_expand<K, CPT, NE>(mdl, z, cur, cur);
                            }
                        } else {
                            auto internal_cols = [&](int ch, const double (&L)[CPT][K], double (&D)[CPT][K]) {
                                double e[K], z[CPT][K];
#pragma unroll
                                for (int k = 0; k < K; ++k) e[k] = eb[(j * 2 + ch) * 2 * K + k];
                                eig_project<K, CPT, false, NE>(mdl, L, e, nullptr, z, z);
                                eig_expand<K, CPT, NE>(mdl, z, L, D);
                            };
                            if (ka == mcp::OPK_MEM) {
                                ld_cols_if(!MCP_EARLY_LOADS || j == 0, rh.y, Lm, std::true_type{});
                                if (MCP_EARLY_LOADS) {
                                    internal_cols(0, Lm, Da);
                                    request_next();
                                    internal_cols(1, cur, Db);
                                } else {   // latency of the stored operand overlaps the product on the register operand
                                    internal_cols(1, cur, Db);
                                    internal_cols(0, Lm, Da);
                                }
                            } else if (ka == mcp::OPK_REG) {
                                request_next();
                                leaf_cols(1, Db);
                                internal_cols(0, cur, Da);
                            } else {
                                request_next();
                                leaf_cols(0, Da);
                                leaf_cols(1, Db);
                            }
#pragma unroll
                            for (int cc = 0; cc < CPT; ++cc) {
#pragma unroll
                                for (int k = 0; k < K; ++k) cur[cc][k] = Da[cc][k] * Db[cc][k];
                                e_col[cc] += rescale_pow2<K>(cur[cc]);
                            }
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
                const int n_pre = tr.n_pre, n_chunks = (n_pre + CHN - 1) / CHN;
                // Operand ring: the post pass wrote its partials through the generic proxy, the ring reads them
                // through the async proxy -- every thread fences its own stores, the barriers of the prologue
                // order the fences before the first bulk copy.
                if constexpr (RD > 0 && !RLS) fence_proxy_async();
                prologue(pre_ops, n_pre, true);
                // ring state of this pass: r_fi indexes the fetch-list entry the next refill requests (RD entries
                // ahead of the one consumed next; the list ends with 0xffff entries), r_nxt holds that entry
                unsigned r_fi = 0, r_nxt = 0;
                // bulk flavour: lane 0 (its scratch base and ring address are the warp's) requests post slot `slot` into
                // stage `st`; cp.async flavour: every thread copies its own pieces of the slot
                auto ring_issue = [&](unsigned st, unsigned slot) {
                    if constexpr (RLS) {
#pragma unroll
                        for (int cc = 0; cc < CPT; ++cc)
#pragma unroll
                            for (int h = 0; h < K / 2; ++h)
                                cp_async16_s(ring_l + (st * CPT + cc) * WB + h * PIECE_STRIDE,
                                             scr + slot * slot_bytes + cc * col_bytes + h * PIECE_STRIDE);
                    } else {
                        const unsigned bar = rbar_w + st * 8;
                        mbar_expect_tx(bar, CPT * WB);
                        bulk_g2s(ring_l + st * (CPT * WB), scr + slot * slot_bytes, CPT * WB, bar);
                    }
                };
                // all lanes: wait for the entry at the head of the ring and read this thread's vectors
                auto ring_consume = [&](double (&v)[CPT][K]) {
                    const unsigned st = r_cnt & (RD - 1);
                    if constexpr (RLS) {
                        cp_async_wait_group<RD - 1>();   // at most the RD - 1 newer refills (or a staging group) stay pending
#pragma unroll
                        for (int cc = 0; cc < CPT; ++cc) lds_partial_pc<K>(ring_l + (st * CPT + cc) * WB, v[cc]);
                    } else {
                        mbar_wait(rbar_w + st * 8, (r_cnt / RD) & 1u);
#pragma unroll
                        for (int cc = 0; cc < CPT; ++cc) lds_partial<K>(ring_l + (st * CPT + cc) * WB, v[cc]);
                    }
                };
                // all lanes, once the vectors just consumed have been used: the freed stage takes the entry RD ahead
                auto ring_refill = [&]() {
                    if constexpr (RLS) {
                        if (r_nxt != 0xffffu) ring_issue(r_cnt & (RD - 1), r_nxt);
                        cp_async_commit();               // a group per refill, empty or not: the waits count groups
                        ++ring_k;
                        ++r_cnt;
                        r_nxt = __ldg(p.fetch + ++r_fi) + lane0;   // (+ lane0: no move to a uniform register right behind the load)
                    } else {
                        __syncwarp();
                        if (lane == 0 && r_nxt != 0xffffu) ring_issue(r_cnt & (RD - 1), r_nxt);
                        ++r_cnt;
                        r_nxt = __ldg(p.fetch + ++r_fi) + lane0;   // (+ lane0: stays in a vector register until lane 0 uses it)
                    }
                };
                if constexpr (RD > 0) {
                    r_fi = (unsigned)tr.fetch_off + (RLS ? 0u : lane0);
                    if (RLS || lane == 0) {
                        for (int i = 0; i < RD; ++i) {
                            const unsigned slot = __ldg(p.fetch + r_fi + i);
                            if (slot != 0xffffu) ring_issue((r_cnt + i) & (RD - 1), slot);
                            if constexpr (RLS) { cp_async_commit(); ++ring_k; }
                        }
                    }
                    r_fi += RD;
                    r_nxt = __ldg(p.fetch + r_fi) + lane0;
                }
                // the first family is the root's (PREM_ROOT, and only that one): its pre vector is pi
#pragma unroll
                for (int cc = 0; cc < CPT; ++cc)
#pragma unroll
                    for (int k = 0; k < K; ++k) cur[cc][k] = mdl.pi(k);
                for (int c = 0; c < n_chunks; ++c) {
                    chunk_boundary(pre_ops, n_pre, c, n_chunks, true);
                    if constexpr (!GL2) {
                        if (c > 0) fold_parked((c - 1) & 1, CHN);   // the chunk before: complete since the barrier above
                    }
                    const OpRec* rb = srec + (c & 1) * CHN;
                    const double* eb = se + (c & 1) * (CHN * 2 * 2 * K);
                    const unsigned char* cb = scode + (size_t)(c & 1) * (CHN * 2 * 2 * TS) + tid;
                    const double* tb = stab + (size_t)(c & 1) * (CHN * 2 * 2 * KK1);
                    const int cnt = min(CHN, n_pre - c * CHN);
                    double La[CPT][K], Lb[CPT][K];                                  // the family's stored child partials
                    // All stored operands of a family -- pre[mother] popped from the LIFO (into `cur`), the
                    // internal children's partials -- are requested together: by the first op of a chunk for
                    // itself, otherwise by the op before it, ahead of its warp reduction and atomics, where
                    // nothing but `cur` is live any more (profiles/r1_walk_notes.md: 19 % of all stall samples
                    // sat on the first use of these loads).
                    auto request = [&](const uint4& r, bool on) {
                        const int fl = (int)r.x;
                        ld_cols_if(on && ((fl >> 8) & 3) == mcp::PREM_STACK, r.w, cur, std::true_type{});
                        ld_cols_if(on && (fl & 3) == mcp::OPK_MEM, r.y, La, std::false_type{});
                        ld_cols_if(on && ((fl >> 2) & 3) == mcp::OPK_MEM, r.z, Lb, std::false_type{});
                    };
                    if (MCP_EARLY_LOADS && RD == 0) request(*reinterpret_cast<const uint4*>(rb), true);
                    for (int j = 0; j < cnt; ++j) {
                        const uint4 rh = *reinterpret_cast<const uint4*>(rb + j);   // flags, xa, xb, y0
                        const int flags = (int)rh.x;
                        const bool ai = (flags & 3) != mcp::OPK_LEAF, bi = ((flags >> 2) & 3) == mcp::OPK_MEM;
                        // Canonical family (schedule.hpp): a is the child whose pre vector stays in
                        // registers (internal, OUT_KEEP) or a leaf; b is pushed (internal, OUT_PUSH) or a
                        // leaf; b internal implies a internal.  pre[mother] lives in `cur`: it is either
                        // already there (PREM_REG: kept by the op just before; PREM_ROOT: pi, set before the
                        // pass) or popped from the LIFO.
                        if constexpr (RD > 0) {
                            // ring kernels: only pre[mother] is loaded directly (a recent push, served by L2) -- by the
                            // op before this one as soon as its own pre[mother] was dead, by the first op of a chunk itself
                            if (j == 0) ld_cols_if(((flags >> 8) & 3) == mcp::PREM_STACK, rh.w, cur, std::true_type{});
                        } else {
                            if (!MCP_EARLY_LOADS) request(rh, true);
                        }
                        // D = P L and Y: leaf child Y = dP L (table column); internal child Y = de * (Uinv L),
                        // the eigen-coordinates of dP L (the numerator is then formed in eigen-space)
                        double Da[CPT][K], Ya[CPT][K], Db[CPT][K], Yb[CPT][K];
                        auto leaf_cols = [&](int ch, double (&D)[CPT][K], double (&Y)[CPT][K]) {
#pragma unroll
                            for (int cc = 0; cc < CPT; ++cc) {
                                const int code = min((int)cb[(j * 2 + ch) * 2 * TS + cc * TW], K);
                                const double* t = tb + (j * 2 + ch) * 2 * KK1 + code * K;
#pragma unroll
                                for (int k = 0; k < K; ++k) { D[cc][k] = t[k]; Y[cc][k] = t[KK1 + k]; }
                            }
                        };
                        auto internal_cols = [&](int ch, const double (&L)[CPT][K], double (&D)[CPT][K], double (&Y)[CPT][K]) {
                            double e[K], z[CPT][K];
#pragma unroll
                            for (int k = 0; k < K; ++k) e[k] = eb[(j * 2 + ch) * 2 * K + k];
                            eig_project<K, CPT, true, NE>(mdl, L, e, eb + (j * 2 + ch) * 2 * K + K, z, Y);
                            eig_expand<K, CPT, NE>(mdl, z, L, D);
                        };
                        if ((flags & 3) == mcp::OPK_CHERRY) {
                            // a is a cherry: its post result is rebuilt from the two leaves below it -- the same
                            // products and the same power-of-two rescaling as in the post pass, hence the same bits
#pragma unroll
                            for (int cc = 0; cc < CPT; ++cc) {
                                const int cx = min((int)cb[(j * 2) * 2 * TS + cc * TW], K);
                                const int cy = min((int)cb[((j * 2) * 2 + 1) * TS + cc * TW], K);
                                const double* tx = tb + (j * 2) * 2 * KK1 + cx * K;
                                const double* ty = tb + (j * 2) * 2 * KK1 + KK1 + cy * K;
#pragma unroll
                                for (int k = 0; k < K; ++k) La[cc][k] = tx[k] * ty[k];
                                rescale_pow2<K>(La[cc]);
                            }
                        }
                        // A STORED child arrives as D = P L (the post pass applied the child's branch before storing),
                        // nothing is recomputed; the eigen-coordinates of dP L follow from dP = U diag(c) Uinv P,
                        // c = D mu rate:  Y = c * (Uinv D).
                        auto stored_cols = [&](const double (&D)[CPT][K], double (&Y)[CPT][K]) {
                            eig_project_rate<K, CPT, NE>(mdl, D, r, Y);
                        };
                        const bool am = (flags & 3) == mcp::OPK_MEM;
                        if constexpr (!DST) {            // stored children are partials L: apply their branches again
                            if constexpr (RD > 0) {
                                if (am) ring_consume(La);
                                if (ai) internal_cols(0, La, Da, Ya); else leaf_cols(0, Da, Ya);
                                if (am) ring_refill();
                                if (bi) {
                                    ring_consume(Lb);
                                    internal_cols(1, Lb, Db, Yb);
                                    ring_refill();
                                } else {
                                    leaf_cols(1, Db, Yb);
                                }
                            } else {
                                if (ai) internal_cols(0, La, Da, Ya); else leaf_cols(0, Da, Ya);
                                if (bi) internal_cols(1, Lb, Db, Yb); else leaf_cols(1, Db, Yb);
                            }
                        } else if constexpr (RD > 0) {
                            if (am) {
                                ring_consume(Da);
                                stored_cols(Da, Ya);
                                ring_refill();
                            } else if (ai) {
                                internal_cols(0, La, Da, Ya);          // cherry, rebuilt above
                            } else {
                                leaf_cols(0, Da, Ya);
                            }
                            if (bi) {
                                ring_consume(Db);
                                stored_cols(Db, Yb);
                                ring_refill();
                            } else {
                                leaf_cols(1, Db, Yb);
                            }
                        } else {
                            if (am) {
#pragma unroll
                                for (int cc = 0; cc < CPT; ++cc)
#pragma unroll
                                    for (int k = 0; k < K; ++k) Da[cc][k] = La[cc][k];
                                stored_cols(Da, Ya);
                            } else if (ai) {
                                internal_cols(0, La, Da, Ya);          // cherry, rebuilt above
                            } else {
                                leaf_cols(0, Da, Ya);
                            }
                            if (bi) {
#pragma unroll
                                for (int cc = 0; cc < CPT; ++cc)
#pragma unroll
                                    for (int k = 0; k < K; ++k) Db[cc][k] = Lb[cc][k];
                                stored_cols(Db, Yb);
                            } else {
                                leaf_cols(1, Db, Yb);
                            }
                        }
                        double qa[CPT][K], qb[CPT][K];
                        double na[CPT], nb[CPT], inv[CPT];
#pragma unroll
                        for (int cc = 0; cc < CPT; ++cc) {
                            double den = 0.0;
                            na[cc] = 0.0;
                            nb[cc] = 0.0;
#pragma unroll
                            for (int k = 0; k < K; ++k) {
                                qa[cc][k] = cur[cc][k] * Db[cc][k];
                                qb[cc][k] = cur[cc][k] * Da[cc][k];
                                den = fma(qa[cc][k], Da[cc][k], den);
                            }
                            inv[cc] = fast_rcp(den) * vmask[cc];
                        }
                        if constexpr (RD > 0) {
                            // a is a leaf: nothing is kept, pre[mother] is dead from here on and the next family's
                            // mother comes off the LIFO -- request it now, a whole op half ahead of its use
                            if (!ai && j + 1 < cnt) {
                                const uint4 rn = *reinterpret_cast<const uint4*>(rb + j + 1);
                                ld_cols_if((((int)rn.x >> 8) & 3) == mcp::PREM_STACK, rn.w, cur, std::true_type{});
                            }
                        }
                        // numerators q . (dP L) and the children's pre vectors P^T q (internal children only)
                        auto num_direct = [&](const double (&q)[CPT][K], const double (&Y)[CPT][K], double (&n)[CPT]) {
#pragma unroll
                            for (int cc = 0; cc < CPT; ++cc)
#pragma unroll
                                for (int k = 0; k < K; ++k) n[cc] = fma(q[cc][k], Y[cc][k], n[cc]);
                        };
                        auto pre_child = [&](int ch, const double (&q)[CPT][K], const double (&Y)[CPT][K], double (&n)[CPT],
                                             double (&out)[CPT][K]) {
                            double e[K];
#pragma unroll
                            for (int k = 0; k < K; ++k) e[k] = eb[(j * 2 + ch) * 2 * K + k];
                            eig_transposed_num<K, CPT, NE>(mdl, q, e, Y, n, out);
#pragma unroll
                            for (int cc = 0; cc < CPT; ++cc) rescale_pow2<K>(out[cc]);
                        };
                        if (bi) {
                            double pb[CPT][K];
                            pre_child(1, qb, Yb, nb, pb);
                            st_cols(rb[j].y2, pb);
                        } else {
                            num_direct(qb, Yb, nb);
                        }
                        if (ai) pre_child(0, qa, Ya, na, cur);   // cur (pre[mother]) is dead: qa, qb hold all that is left of it
                        else num_direct(qa, Ya, na);
                        if (MCP_EARLY_LOADS && RD == 0) request(*reinterpret_cast<const uint4*>(rb + (j + 1 < cnt ? j + 1 : j)), j + 1 < cnt);
                        double ga = 0.0, gb = 0.0;
#pragma unroll
                        for (int cc = 0; cc < CPT; ++cc) {
                            ga = fma(na[cc], inv[cc], ga);
                            gb = fma(nb[cc], inv[cc], gb);
                        }
                        const double red = warp_pair_reduce(ga, gb, lane);   // lane 0: sum of ga, lane 16: sum of gb
                        if constexpr (GL2) {
                            if ((lane & 15) == 0) atomicAdd(grow + ((lane >> 4) ? rb[j].b_br : rb[j].a_br), red);
                        } else {
                            if ((lane & 15) == 0) {
                                const int term = ((c & 1) * CHN + j) * 2 + (lane >> 4);
                                s_part[term * 8 + warp] = red;
                                if (warp == 0) s_pbr[term] = (lane >> 4) ? rb[j].b_br : rb[j].a_br;
                            }
                        }
                    }
                }
                if constexpr (!GL2) {
                    __syncthreads();                                   // the last chunk's sums are parked
                    fold_parked((n_chunks - 1) & 1, n_pre - (n_chunks - 1) * CHN);
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

}  // namespace
