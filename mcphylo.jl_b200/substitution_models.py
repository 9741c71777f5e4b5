"""Substitution models: eigendecomposition (U, D, Uinv, mu) of the rate matrix Q.

Host-side part of the hot path, restating /root/reference/src/Likelihood/SubstitutionModels.jl:
  Restriction :13-24   JC :36-50   GTR :62-74   freeK :83-102   setmatrix :106-113
The device builds P(t) = U diag(exp(mu t D r)) Uinv from these (csrc/kernel_tables.cuh,
kernel build_transition_tables), so only K x K work happens here.

`eigen` follows what Julia's LinearAlgebra.eigen does for a real matrix: the symmetric
solver (ascending eigenvalues) when Q is symmetric, otherwise the general solver with
eigenvalues sorted ascending by (real, imag).  Eigenvector scaling/sign is LAPACK's and
cancels in U diag(.) Uinv.
"""
from __future__ import annotations

from typing import Sequence, Tuple

import numpy as np

ModelOut = Tuple[np.ndarray, np.ndarray, np.ndarray, float]


def _eigen(Q: np.ndarray) -> Tuple[np.ndarray, np.ndarray]:
    Q = np.asarray(Q, dtype=np.float64)
    if np.array_equal(Q, Q.T):
        D, U = np.linalg.eigh(Q)
        return D, U
    D, U = np.linalg.eig(Q)
    order = np.lexsort((D.imag, D.real))
    D = D[order]
    U = U[:, order]
    if np.any(D.imag != 0.0) or np.iscomplexobj(U) and np.any(U.imag != 0.0):
        # The reference's Float64-typed transition arrays cannot hold complex values
        # (VectorizedFunctions.jl:155-168 allocates Array{Float64,4}); same restriction here.
        raise ValueError("rate matrix has complex eigenvalues; not supported on this path")
    return np.ascontiguousarray(D.real), np.ascontiguousarray(U.real)


def Restriction(base_freq: Sequence[float], SubstitutionRates: Sequence[float] = ()) -> ModelOut:
    """Two-state restriction-site model, closed form (SubstitutionModels.jl:13-24)."""
    pi = np.asarray(base_freq, dtype=np.float64)
    assert pi.shape == (2,), "Restriction needs exactly two base frequencies"
    D = np.array([-1.0, 0.0])
    U = np.array([[pi[1], pi[1]], [pi[1] - 1.0, pi[1]]])
    Uinv = np.linalg.inv(U)
    mu = 1.0 / (2.0 * pi[0] * pi[1])
    return U, D, Uinv, float(mu)


def JC(base_freq: Sequence[float], SubstitutionRates: Sequence[float] = ()) -> ModelOut:
    """Jukes-Cantor for K = len(base_freq) states (SubstitutionModels.jl:36-50)."""
    K = len(base_freq)
    off = 1.0 / K
    diag = off * (K - 1)
    Q = np.full((K, K), off)
    np.fill_diagonal(Q, -diag)
    D, U = _eigen(Q)
    Uinv = np.linalg.inv(U)
    mu = 1.0 / diag          # `sum(diag)` of a scalar in the reference
    return U, D, Uinv, float(mu)


def setmatrix(vec_vals: Sequence[float]) -> np.ndarray:
    """Symmetric matrix with zero diagonal whose upper triangle is filled column by column
    (Julia comprehension order, SubstitutionModels.jl:106-113)."""
    v = np.asarray(vec_vals, dtype=np.float64)
    n = v.size
    s = int(round((np.sqrt(8 * n + 1) + 1) / 2))
    if s * (s - 1) // 2 != n:
        raise ValueError("setmatrix: length of vector is not triangular")
    Q = np.zeros((s, s))
    k = 0
    for j in range(s):          # column-major walk: j outer, i inner
        for i in range(s):
            if i < j:
                Q[i, j] = v[k]
                k += 1
    return Q + Q.T


def GTR(base_freq: Sequence[float], SubstitutionRates: Sequence[float]) -> ModelOut:
    """General time-reversible model (SubstitutionModels.jl:62-74)."""
    pi = np.asarray(base_freq, dtype=np.float64)
    K = pi.size
    Q = setmatrix(SubstitutionRates)
    Q = Q * pi[None, :]                     # every row multiplied elementwise by pi
    dia = Q.sum(axis=1)
    Q[np.arange(K), np.arange(K)] = -dia
    D, U = _eigen(Q)
    Uinv = np.linalg.inv(U)
    return U, D, Uinv, float(1.0 / dia.sum())


def freeK(base_freq: Sequence[float], SubstitutionRates: Sequence[float]) -> ModelOut:
    """Unconstrained K-state model; rates fill Q off-diagonals as Q[j, i] for i outer, j
    inner (SubstitutionModels.jl:83-102)."""
    r = np.asarray(SubstitutionRates, dtype=np.float64).ravel()
    K = int(np.ceil(np.sqrt(r.size)))
    Q = np.zeros((K, K))
    c = 0
    for i in range(K):
        for j in range(K):
            if i != j:
                Q[j, i] = r[c]
                c += 1
    dia = Q.sum(axis=1)
    Q[np.arange(K), np.arange(K)] = -dia
    D, U = _eigen(Q)
    Uinv = np.linalg.inv(U)
    return U, D, Uinv, float(1.0 / dia.sum())


def freeK_equilibrium(SubstitutionRates: Sequence[float]) -> np.ndarray:
    """Equilibrium frequencies the reference derives for the freeK constructor
    (/root/reference/src/distributions/Phylodist.jl:85-97)."""
    U, D, Uinv, _ = freeK([], SubstitutionRates)
    D = np.zeros_like(D)
    D[-1] = 1.0
    return np.real((U @ np.diag(D) @ Uinv)[0, :])


def transition_matrix(model_out: ModelOut, t: float, rate: float = 1.0) -> np.ndarray:
    """Host P(t) with the reference's operation order (VectorizedFunctions.jl:116-152);
    used by tests and by the synthetic-alignment simulator, never by the product path."""
    U, D, Uinv, mu = model_out
    return (U * np.exp(mu * t * D * rate)[None, :]) @ Uinv


# ---------------------------------------------------------------------------------------------
# Derivatives of the NORMALISED rate matrix A = mu * U diag(D) Uinv (so that P(t) = exp(A t rate))
# with respect to the model's own parameters -- what mcp_eval_model_gradient takes as `dA` / `dpi`
# (include/mcphylo_b200.h).  The reference has no such derivative; these follow
# the functional forms above literally, base frequencies treated as independent coordinates (a
# caller on the simplex projects afterwards).
# ---------------------------------------------------------------------------------------------
def normalised_rate_matrix(model_out: ModelOut) -> np.ndarray:
    U, D, Uinv, mu = model_out
    return mu * ((U * np.asarray(D)[None, :]) @ Uinv)


def _gtr_A(pi: np.ndarray, sr: np.ndarray):
    S = setmatrix(sr)
    Q = S * pi[None, :]
    dia = Q.sum(axis=1)
    Q[np.arange(pi.size), np.arange(pi.size)] = -dia
    return S, Q, 1.0 / dia.sum()


def model_derivatives(model, base_freq: Sequence[float], SubstitutionRates: Sequence[float] = ()):
    """(names, dA, dpi): dA[:, :, p] = d A / d theta_p (K x K), dpi[:, p] = d pi / d theta_p as seen by the root term,
    for theta = (base_freq[0..K-1], SubstitutionRates[0..]) -- analytic for Restriction, JC, GTR and freeK, a
    Richardson-extrapolated central difference of the model function otherwise."""
    pi = np.atleast_1d(np.asarray(base_freq, dtype=np.float64))
    sr = np.atleast_1d(np.asarray(SubstitutionRates, dtype=np.float64)).ravel()
    K = pi.size
    names = [f"base_freq[{i}]" for i in range(K)] + [f"substitution_rates[{i}]" for i in range(sr.size)]
    n_par = K + sr.size
    dA = np.zeros((K, K, n_par))
    dpi = np.zeros((K, n_par))
    dpi[np.arange(K), np.arange(K)] = 1.0
    if model is JC:
        pass                                            # Q has no parameters
    elif model is Restriction:
        # the closed form above gives Q = [[-pi2, pi2], [1 - pi2, -(1 - pi2)]] (a function of pi2 only) and mu = 1 / (2 pi1 pi2)
        p1, p2 = pi
        Q = np.array([[-p2, p2], [1.0 - p2, -(1.0 - p2)]])
        mu = 1.0 / (2.0 * p1 * p2)
        dA[:, :, 0] = -(mu / p1) * Q
        dA[:, :, 1] = -(mu / p2) * Q + mu * np.array([[-1.0, 1.0], [-1.0, 1.0]])
    elif model is GTR:
        S, Q, mu = _gtr_A(pi, sr)
        for m in range(K):                              # d / d pi_m: column m of Q is S[:, m] pi_m, the diagonal carries -row sums
            dQ = np.zeros((K, K))
            dQ[:, m] = S[:, m]
            dQ[np.arange(K), np.arange(K)] -= S[:, m]
            dmu = -mu * mu * S[:, m].sum()
            dA[:, :, m] = mu * dQ + dmu * Q
        k = 0
        for j in range(K):                              # setmatrix order: column by column, i < j
            for i in range(j):
                dQ = np.zeros((K, K))
                dQ[i, j] = pi[j]
                dQ[j, i] = pi[i]
                dQ[i, i] = -pi[j]
                dQ[j, j] = -pi[i]
                dmu = -mu * mu * (pi[i] + pi[j])
                dA[:, :, K + k] = mu * dQ + dmu * Q
                k += 1
    elif model is freeK:
        Kq = int(np.ceil(np.sqrt(sr.size)))
        assert Kq == K, "freeK: number of rates and of base frequencies disagree"
        Q = np.zeros((K, K))
        c = 0
        where = []
        for i in range(K):
            for j in range(K):
                if i != j:
                    Q[j, i] = sr[c]
                    where.append((j, i))
                    c += 1
        dia = Q.sum(axis=1)
        Q[np.arange(K), np.arange(K)] = -dia
        mu = 1.0 / dia.sum()
        for c, (j, i) in enumerate(where):
            dQ = np.zeros((K, K))
            dQ[j, i] = 1.0
            dQ[j, j] = -1.0
            dA[:, :, K + c] = mu * dQ - mu * mu * Q
    else:
        theta = np.concatenate([pi, sr])

        def A_of(th):
            return normalised_rate_matrix(model(th[:K].copy(), th[K:].copy()))

        for p in range(n_par):
            h = 1e-3 * max(abs(theta[p]), 1e-2)

            def cd(step):
                tp, tm = theta.copy(), theta.copy()
                tp[p] += step
                tm[p] -= step
                return (A_of(tp) - A_of(tm)) / (2.0 * step)

            d1, d2, d3 = cd(h), cd(h / 2.0), cd(h / 4.0)
            r1, r2 = (4.0 * d2 - d1) / 3.0, (4.0 * d3 - d2) / 3.0      # O(h^4), then O(h^6)
            dA[:, :, p] = (16.0 * r2 - r1) / 15.0
    return names, dA, dpi
