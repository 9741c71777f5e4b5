"""Extended-precision arbiter for the PhyloDist hot path — TEST INFRASTRUCTURE ONLY (see
oracle/__init__.py for who may import this package).

`felsenstein_extended` evaluates the SAME quantity as the reference
(/root/reference/src/Likelihood/LikelihoodCalculator_Node.jl:3-114, semantics of SURVEY.md section 8:
rate categories not mixed, gradient entry num = branch above node num) in numpy long double (x87
80-bit, eps 1.1e-19), with P(t) = exp(Q mu t r) from a scaling-and-squaring Taylor series instead of
the reference's U diag(exp(.)) Uinv product.  It is NOT the parity oracle (that is
felsenstein_oracle.c, which restates the reference's fp64 arithmetic loop for loop); it exists to
decide who is right when the CUDA path and the fp64 oracle disagree beyond the tolerance: the
reference formula forms small off-diagonal entries of P by cancellation ("Conditioning"),
so for short branches x slow rate categories its own error reaches 1e-8..1e-5 on single gradient
components, while the CUDA path (expm1 form) stays at 1e-13 of this arbiter.

Q is taken from the caller's decomposition, Q = U diag(D) Uinv evaluated in long double, so all three
implementations share the same generator.  Dense numpy over sites: seconds for 40 taxa x 2500 sites.
"""
from __future__ import annotations

import numpy as np

LD = np.longdouble


def expm_extended(A: np.ndarray) -> np.ndarray:
    """exp(A) in long double: scaling and squaring around a 40-term Taylor series."""
    A = np.asarray(A, dtype=LD)
    norm = float(np.abs(A).sum(axis=1).max())
    n = max(0, int(np.ceil(np.log2(norm))) + 8) if norm > 0 else 0
    B = A / LD(2) ** n
    E = np.eye(A.shape[0], dtype=LD)
    T = np.eye(A.shape[0], dtype=LD)
    for k in range(1, 40):
        T = T @ B / LD(k)
        E = E + T
    for _ in range(n):
        E = E @ E
    return E


def felsenstein_extended(codes, leaf_nums, K, postorder_num, parent_num, blv, U, D, Uinv, mu, rates, pi):
    """(logL, grad[NN-1]) as long doubles.  `codes` (n_leaves, S) uint8 with code K = gap; the tree as
    the C ABI takes it (1-based post-order numbers, parent numbers indexed by num-1, blv by num-1)."""
    po = np.asarray(postorder_num, dtype=np.int64)
    pa = np.asarray(parent_num, dtype=np.int64)
    NN, S = po.size, codes.shape[1]
    Q = np.asarray(U, LD) @ np.diag(np.asarray(D, LD)) @ np.asarray(Uinv, LD)
    pos = {int(num) - 1: i for i, num in enumerate(po)}
    children = [[] for _ in range(NN)]
    for n in range(NN - 1):
        children[pa[n] - 1].append(n)
    for c in children:                      # stored child order = order of appearance in post-order
        c.sort(key=lambda n: pos[n])
    row = {int(n) - 1: i for i, n in enumerate(leaf_nums)}
    pi = np.asarray(pi, LD)
    ll = LD(0)
    grad = np.zeros(NN - 1, dtype=LD)
    for r in np.asarray(rates, dtype=np.float64):
        P, dP = {}, {}
        for b in range(NN - 1):
            P[b] = expm_extended(Q * LD(mu) * LD(blv[b]) * LD(r))
            dP[b] = (Q * LD(mu) * LD(r)) @ P[b]
        L, Down = {}, {}
        scal = np.zeros(S, dtype=LD)
        for num in po:
            n = int(num) - 1
            if not children[n]:
                c = np.minimum(codes[row[n]].astype(int), K)
                M = np.ones((K, S), dtype=LD)
                hot = np.nonzero(c < K)[0]
                M[:, hot] = 0
                M[c[hot], hot] = 1
                L[n] = M
                continue
            acc = np.ones((K, S), dtype=LD)
            for ch in children[n]:
                Down[ch] = P[ch] @ L[ch]
                acc = acc * Down[ch]
            if n != NN - 1:
                m = acc.max(axis=0)
                scal += np.log(m)
                acc = acc / m
            L[n] = acc
        ll += (scal + np.log((pi[:, None] * L[NN - 1]).sum(axis=0))).sum()
        pre = {NN - 1: np.repeat(pi[:, None], S, axis=1)}
        for num in po[::-1]:
            n = int(num) - 1
            if n == NN - 1:
                continue
            m = pa[n] - 1
            q = pre[m].copy()
            for sib in children[m]:
                if sib != n:
                    q = q * Down[sib]
            numer = (q * (dP[n] @ L[n])).sum(axis=0)
            pn = P[n].T @ q
            grad[n] += (numer / (L[n] * pn).sum(axis=0)).sum()
            if children[n]:
                pn = pn / pn.max(axis=0)
            pre[n] = pn
    return ll, grad
