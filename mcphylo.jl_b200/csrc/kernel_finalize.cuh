// kernel_finalize.cuh — kernel 3: fixed-order reduction of the accumulator rows.
// Part of libmcphylo_b200.so; included by mcphylo_b200.cu only (one translation unit).
#pragma once

namespace {

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
            // 8 rows in flight per step (same order of additions): the row list is walked at one L2 round trip
            // per 8 rows instead of one per row
            for (int rw0 = tr.row_lo + g; rw0 < tr.row_hi; rw0 += 8 * FIN_G) {
                double t[8];
#pragma unroll
                for (int b = 0; b < 8; ++b) {
                    const int rw = rw0 + b * FIN_G;
                    t[b] = rw < tr.row_hi ? __ldcg(rows + (long long)rw * row_stride + (j - 1)) : 0.0;
                }
#pragma unroll
                for (int b = 0; b < 8; ++b)
                    if (rw0 + b * FIN_G < tr.row_hi) v += t[b];
            }
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

}  // namespace

namespace {

// --------------------------------------------------------------------------------------------
// Fixed-order sum of n result vectors of `len` doubles, `stride` doubles apart:  out[j] = sum_i in[i][j].
// Used for the site blocks of a streamed alignment and, on a multi-device context in PEER mode, for
// the per-device parts the other GPUs have written into device 0's gather buffer over NVLink (`out` may
// be pinned host memory: the result then needs no separate copy).
// --------------------------------------------------------------------------------------------
__global__ void sum_rows(const double* __restrict__ in, long long stride, int n, long long len, double* __restrict__ out) {
    const long long j = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= len) return;
    double v = in[j];
    for (int i = 1; i < n; ++i) v += in[(long long)i * stride + j];
    out[j] = v;
}

// --------------------------------------------------------------------------------------------
// Per-evaluation parameters and topology blocks reach the device through THIS kernel, not through the copy
// engine: it reads the pinned host staging buffer directly (unified addressing) and writes the device copy.
// A cudaMemcpyAsync on the evaluation stream shares the host-to-device copy engine with the bulk alignment
// transfers of mcp_eval_streamed, and the engine does not serve streams in issue order: measured on B200, the
// 16 KB parameter copy of the first site block waited for four later bulk transfers (11 ms) before its
// kernel could start (profiles/r2_e2e_timeline.json).  n16 = number of 16-byte words.
// --------------------------------------------------------------------------------------------
__global__ void stage_from_host(const uint4* __restrict__ h_src, uint4* __restrict__ d_dst, long long n16) {
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n16; i += (long long)gridDim.x * blockDim.x)
        d_dst[i] = h_src[i];
}

}  // namespace
