// epilogue_prior.cuh — branch-length prior folded into the final reduction.
// Part of libmcphylo_b200.so; included by mcphylo_b200.cu only (one translation unit).
#pragma once

namespace {

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

}  // namespace
