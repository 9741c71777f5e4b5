// device_math.cuh — per-thread arithmetic: partial-vector loads/stores, eigen-space products, power-of-two rescaling, reductions.
// Part of libmcphylo_b200.so; included by mcphylo_b200.cu only (one translation unit).
#pragma once

namespace {

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

// Predicated load INTO the existing registers (KEEP_OLD: v keeps its value when !pred).  A plain `if (pred)
// ld_partial(...)` makes the compiler load into temporaries and move them over under the predicate
// (asm outputs are always fresh values): K moves per load in the innermost loop.
// KEEP_OLD = false: v is undefined when !pred -- tells the compiler the previous contents are dead, so
// the registers are free between the last use of v and this load.
template <int K, bool KEEP_OLD>
__device__ __forceinline__ void ld_partial_if(bool pred, const double* p, double (&v)[K]) {
    if constexpr (!KEEP_OLD) {
#pragma unroll
        for (int k = 0; k < K; ++k) asm volatile("" : "=d"(v[k]));
    }
    if constexpr (K == 4) {
        asm volatile("{\n\t.reg .pred q;\n\tsetp.ne.u32 q, %5, 0;\n\t@q ld.global.cg.v4.f64 {%0,%1,%2,%3}, [%4];\n\t}"
                     : "+d"(v[0]), "+d"(v[1]), "+d"(v[2]), "+d"(v[3]) : "l"(p), "r"((unsigned)pred) : "memory");
    } else if constexpr (K == 2) {
        asm volatile("{\n\t.reg .pred q;\n\tsetp.ne.u32 q, %3, 0;\n\t@q ld.global.cg.v2.f64 {%0,%1}, [%2];\n\t}"
                     : "+d"(v[0]), "+d"(v[1]) : "l"(p), "r"((unsigned)pred) : "memory");
    } else {
#pragma unroll
        for (int k = 0; k < K; ++k)
            asm volatile("{\n\t.reg .pred q;\n\tsetp.ne.u32 q, %2, 0;\n\t@q ld.global.cg.f64 %0, [%1];\n\t}"
                         : "+d"(v[k]) : "l"(p + k), "r"((unsigned)pred) : "memory");
    }
}

// ---- piece layout (cp.async flavour of the operand ring): a K-vector is split into 16-byte pieces, piece h of lane l
// at + h * 512 + l * 16 inside the warp's share of a column, so that every warp access (global: 512 contiguous bytes,
// shared: 32 x 16 bytes) is dense and conflict-free and a thread can move its own vector with 16-byte cp.async copies ----
constexpr unsigned PIECE_STRIDE = 32 * 16;
template <int K>
__device__ __forceinline__ void ld_partial_pc(const unsigned char* p, double (&v)[K]) {
    static_assert(K % 2 == 0, "piece layout needs an even state count");
#pragma unroll
    for (int h = 0; h < K / 2; ++h)
        asm volatile("ld.global.cg.v2.f64 {%0,%1}, [%2];" : "=d"(v[2 * h]), "=d"(v[2 * h + 1]) : "l"(p + h * PIECE_STRIDE) : "memory");
}
template <int K>
__device__ __forceinline__ void st_partial_pc(unsigned char* p, const double (&v)[K]) {
#pragma unroll
    for (int h = 0; h < K / 2; ++h)
        asm volatile("st.global.cg.v2.f64 [%0], {%1,%2};" :: "l"(p + h * PIECE_STRIDE), "d"(v[2 * h]), "d"(v[2 * h + 1]) : "memory");
}
template <int K, bool KEEP_OLD>
__device__ __forceinline__ void ld_partial_pc_if(bool pred, const unsigned char* p, double (&v)[K]) {
    if constexpr (!KEEP_OLD) {
#pragma unroll
        for (int k = 0; k < K; ++k) asm volatile("" : "=d"(v[k]));
    }
#pragma unroll
    for (int h = 0; h < K / 2; ++h)
        asm volatile("{\n\t.reg .pred q;\n\tsetp.ne.u32 q, %3, 0;\n\t@q ld.global.cg.v2.f64 {%0,%1}, [%2];\n\t}"
                     : "+d"(v[2 * h]), "+d"(v[2 * h + 1]) : "l"(p + h * PIECE_STRIDE), "r"((unsigned)pred) : "memory");
}
template <int K>
__device__ __forceinline__ void lds_partial_pc(unsigned addr, double (&v)[K]) {
#pragma unroll
    for (int h = 0; h < K / 2; ++h)
        asm volatile("ld.shared.v2.f64 {%0,%1}, [%2];" : "=d"(v[2 * h]), "=d"(v[2 * h + 1]) : "r"(addr + h * PIECE_STRIDE) : "memory");
}
__device__ __forceinline__ void cp_async16_s(unsigned smem_dst, const void* gmem_src) {
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" :: "r"(smem_dst), "l"(gmem_src) : "memory");
}
template <int N>
__device__ __forceinline__ void cp_async_wait_group() { asm volatile("cp.async.wait_group %0;" :: "n"(N) : "memory"); }

// Identity moves the compiler cannot see through: a value passed through them is kept in a register
// (or spilled as one word) instead of being RE-COMPUTED at every use.  ptxas otherwise rematerialises
// the per-thread scratch base (blockIdx * scratch_per_cta + tid * K * 8, ~13 instructions) in front of
// every partial load/store of the walk.
__device__ __forceinline__ unsigned char* keep_ptr(unsigned char* p) {
    asm volatile("mov.u64 %0, %0;" : "+l"(p));
    return p;
}
__device__ __forceinline__ unsigned keep_u32(unsigned v) {
    asm volatile("mov.u32 %0, %0;" : "+r"(v));
    return v;
}
__device__ __forceinline__ double keep_f64(double v) {
    asm volatile("mov.f64 %0, %0;" : "+d"(v));
    return v;
}

__device__ __forceinline__ void cp_async16(void* smem_dst, const void* gmem_src) {
    unsigned d = (unsigned)__cvta_generic_to_shared(smem_dst);
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" :: "r"(d), "l"(gmem_src) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_group 0;" ::: "memory"); }

// ---- operand ring primitives (kernel_walk.cuh, RD > 0): mbarrier + bulk asynchronous copy, shared-window addresses ----
__device__ __forceinline__ void mbar_init(unsigned bar, unsigned count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" :: "r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void fence_mbar_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
// orders this thread's generic-proxy accesses (here: its st.global of partials) before later async-proxy accesses
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async;" ::: "memory"); }
__device__ __forceinline__ void mbar_expect_tx(unsigned bar, unsigned bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" :: "r"(bar), "r"(bytes) : "memory");
}
// All 32 lanes wait together.  The loop is written in C with a warp vote as its condition: a branch loop inside the
// asm block hid the control flow from the compiler, which then gave up keeping U / Uinv in uniform registers anywhere
// in the kernel (+35 % instructions in the post pass); a per-lane condition still made it reload them after every wait.
__device__ __forceinline__ unsigned mbar_try(unsigned bar, unsigned parity) {
    unsigned ok;
    asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                 : "=r"(ok) : "r"(bar), "r"(parity) : "memory");
    return ok;
}
__device__ __forceinline__ void mbar_wait(unsigned bar, unsigned parity) {
    while (!__all_sync(0xffffffffu, mbar_try(bar, parity))) {}
}
// global -> shared bulk copy of `bytes` (multiple of 16, both addresses 16-byte aligned), completion counted on `bar`
__device__ __forceinline__ void bulk_g2s(unsigned dst, const void* src, unsigned bytes, unsigned bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 :: "r"(dst), "l"(src), "r"(bytes), "r"(bar) : "memory");
}
template <int K>
__device__ __forceinline__ void lds_partial(unsigned addr, double (&v)[K]) {
    if constexpr (K % 2 == 0) {
#pragma unroll
        for (int k = 0; k < K; k += 2)
            asm volatile("ld.shared.v2.f64 {%0,%1}, [%2];" : "=d"(v[k]), "=d"(v[k + 1]) : "r"(addr + k * 8) : "memory");
    } else {
#pragma unroll
        for (int k = 0; k < K; ++k) asm volatile("ld.shared.f64 %0, [%1];" : "=d"(v[k]) : "r"(addr + k * 8) : "memory");
    }
}

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

// NE = number of ACTIVE eigen-components: the host moves a null eigenvalue (every rate matrix has
// one: its rows sum to zero) to the last position, where em1 = expm1(0) = 0 and de = 0 contribute
// nothing, and the walk kernel is instantiated with NE = K - 1 (a quarter of the matrix-vector work
// and of the U / Uinv constants at K = 4, half at K = 2).
//
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
template <int K, int C, bool WD, int NE = K, class M>
__device__ __forceinline__ void eig_project(const M& m, const double (&L)[C][K], const double (&em1)[K], const double* de,
                                            double (&z)[C][K], double (&zd)[C][K]) {
#pragma unroll
    for (int i = 0; i < NE; ++i) {
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
// yd[c] = c_r * (Uinv D[c]),  c_r = D mu rate_r:  the eigen-coordinates of dP L when D = P L is what is at hand
// (dP = U diag(c_r) Uinv P); rate category `r` is uniform over the tile
template <int K, int C, int NE = K, class M>
__device__ __forceinline__ void eig_project_rate(const M& m, const double (&D)[C][K], int r, double (&yd)[C][K]) {
#pragma unroll
    for (int i = 0; i < NE; ++i) {
        double w[C];
#pragma unroll
        for (int c = 0; c < C; ++c) w[c] = m.Ui(i, 0) * D[c][0];
#pragma unroll
        for (int j = 1; j < K; ++j)
#pragma unroll
            for (int c = 0; c < C; ++c) w[c] = fma(m.Ui(i, j), D[c][j], w[c]);
        const double cr = m.c(r, i);
#pragma unroll
        for (int c = 0; c < C; ++c) yd[c][i] = cr * w[c];
    }
}
// out[c] = base[c] + U z[c]
template <int K, int C, int NE = K, class M>
__device__ __forceinline__ void eig_expand(const M& m, const double (&z)[C][K], const double (&base)[C][K], double (&out)[C][K]) {
#pragma unroll
    for (int s = 0; s < K; ++s) {
#pragma unroll
        for (int c = 0; c < C; ++c) out[c][s] = fma(m.U(s, 0), z[c][0], base[c][s]);
#pragma unroll
        for (int i = 1; i < NE; ++i)
#pragma unroll
            for (int c = 0; c < C; ++c) out[c][s] = fma(m.U(s, i), z[c][i], out[c][s]);
    }
}
// out[c] = U z[c]
template <int K, int C, int NE = K, class M>
__device__ __forceinline__ void eig_expand0(const M& m, const double (&z)[C][K], double (&out)[C][K]) {
#pragma unroll
    for (int s = 0; s < K; ++s) {
#pragma unroll
        for (int c = 0; c < C; ++c) out[c][s] = m.U(s, 0) * z[c][0];
#pragma unroll
        for (int i = 1; i < NE; ++i)
#pragma unroll
            for (int c = 0; c < C; ++c) out[c][s] = fma(m.U(s, i), z[c][i], out[c][s]);
    }
}
// out[c] = P^T q[c] = q[c] + Uinv^T (em1 * (U^T q[c]))
template <int K, int C, int NE = K, class M>
__device__ __forceinline__ void eig_transposed(const M& m, const double (&q)[C][K], const double (&em1)[K], double (&out)[C][K]) {
    double z[C][K];
#pragma unroll
    for (int i = 0; i < NE; ++i) {
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
        for (int i = 1; i < NE; ++i)
#pragma unroll
            for (int c = 0; c < C; ++c) out[c][j] = fma(m.Ui(i, j), z[c][i], out[c][j]);
    }
}

// As eig_transposed, and also the branch-gradient numerator q . (dP L) formed in eigen-space:
//   q^T U diag(de) Uinv L = sum_i (U^T q)_i * yd_i,   yd = de * (Uinv L)  (from eig_project<WD>),
// which reuses w = U^T q and saves the expansion U yd (K*K FMAs per internal child).
template <int K, int C, int NE = K, class M>
__device__ __forceinline__ void eig_transposed_num(const M& m, const double (&q)[C][K], const double (&em1)[K],
                                                   const double (&yd)[C][K], double (&num)[C], double (&out)[C][K]) {
    double z[C][K];
#pragma unroll
    for (int i = 0; i < NE; ++i) {
        double w[C];
#pragma unroll
        for (int c = 0; c < C; ++c) w[c] = m.U(0, i) * q[c][0];
#pragma unroll
        for (int s = 1; s < K; ++s)
#pragma unroll
            for (int c = 0; c < C; ++c) w[c] = fma(m.U(s, i), q[c][s], w[c]);
#pragma unroll
        for (int c = 0; c < C; ++c) {
            z[c][i] = em1[i] * w[c];
            num[c] = fma(w[c], yd[c][i], num[c]);
        }
    }
#pragma unroll
    for (int j = 0; j < K; ++j) {
#pragma unroll
        for (int c = 0; c < C; ++c) out[c][j] = fma(m.Ui(0, j), z[c][0], q[c][j]);
#pragma unroll
        for (int i = 1; i < NE; ++i)
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

}  // namespace
