// kernel_tables.cuh — kernel 1: branch tables.
// Part of libmcphylo_b200.so; included by mcphylo_b200.cu only (one translation unit).
#pragma once

namespace {

// --------------------------------------------------------------------------------------------
// kernel 1: branch tables for every (tree, branch, rate)
//   e   = exp(mu t D r)
//   P   = U diag(e) Uinv                             VectorizedFunctions.jl:116-168
//   dP  = U diag(D r mu e) Uinv                      VectorizedFunctions.jl:89-113, 139-152
// (P, dP columns are only read for LEAF children; same operation order as the reference.)
// --------------------------------------------------------------------------------------------
constexpr int KMAX_TABLE = 32;

__global__ void build_branch_tables(const TreeDev* __restrict__ trees, const double* __restrict__ dyn,
                                    double* __restrict__ btab, int K, int R) {
    const TreeDev tr = trees[blockIdx.y];
    const int idx = blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= tr.n_br * R) return;
    const int br = idx / R, r = idx - br * R;
    const double* d = dyn + tr.dyn_off;
    const double* U = d + dyn_U(tr.NN);
    const double* D = d + dyn_D(tr.NN, K);
    const double* Uinv = d + dyn_Uinv(tr.NN, K);
    const double mu = d[dyn_mu(tr.NN, K)];
    const double rate = d[dyn_rates(tr.NN, K) + r];
    double* ev = btab + tr.btab_off + ((long long)br * R + r) * bt_size(K);
    double* P = ev + 2 * K;
    double* dP = P + K * (K + 1);
    if (br >= tr.NN - 1) {  // root row (unused) and virtual branches: identity, zero derivative
        for (int i = 0; i < K; ++i) { ev[i] = 0.0; ev[K + i] = 0.0; }   // expm1(0) and zero derivative: identity branch
        for (int n = 0; n <= K; ++n)
            for (int m = 0; m < K; ++m) {
                P[n * K + m] = (n == K || n == m) ? 1.0 : 0.0;
                dP[n * K + m] = 0.0;
            }
        return;
    }
    const double t = d[dyn_blv(tr.NN) + br];
    double em1[KMAX_TABLE], de[KMAX_TABLE];
    for (int i = 0; i < K; ++i) {
        const double x = mu * t * D[i] * rate;
        em1[i] = expm1(x);
        ev[i] = em1[i];
        de[i] = D[i] * rate * mu * exp(x);
        ev[K + i] = de[i];
    }
    for (int m = 0; m < K; ++m) {
        double rs = 0.0, drs = 0.0;
        for (int n = 0; n < K; ++n) {
            double c = 0.0, dc = 0.0;
            for (int k = 0; k < K; ++k) {
                const double u = U[m + K * k], ui = Uinv[k + K * n];
                c += (u * em1[k]) * ui;      // P - I, formed without cancellation against the identity
                dc += (u * de[k]) * ui;
            }
            c += (m == n) ? 1.0 : 0.0;
            P[n * K + m] = c;
            dP[n * K + m] = dc;
            rs += c;    // what P * (all-ones leaf) gives: sum_s1 1 * P[s, s1]
            drs += dc;
        }
        P[K * K + m] = rs;
        dP[K * K + m] = drs;
    }
}

}  // namespace
