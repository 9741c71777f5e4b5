"""Discretised-Gamma rate categories (Yang 1994), restating
/root/reference/src/Likelihood/Rates.jl:11-58 including its parametrisation quirks
(the chi-square has (2a)/(2b) degrees of freedom, and the median boundaries are not
divided by 2b).  Golden vectors: /root/reference/test/likelihood/rates.jl:3-18.

User-side helper that produces the `rates` vector handed to PhyloDist; the likelihood
itself never calls it.
"""
from __future__ import annotations

import numpy as np
from scipy.special import gammainc
from scipy.stats import chi2


def mean_boundaries(alpha: float, beta: float, k: int) -> np.ndarray:
    df = (2.0 * alpha) / (2.0 * beta)
    b = np.zeros(k, dtype=np.float64)
    for i in range(1, k):
        b[i - 1] = chi2.ppf(i / k, df) / (2.0 * beta)
    return b


def median_boundaries(alpha: float, beta: float, k: int) -> np.ndarray:
    df = (2.0 * alpha) / (2.0 * beta)
    return np.array([chi2.ppf(((i - 1) * 2 + 1) / (2 * k), df) for i in range(1, k + 1)],
                    dtype=np.float64)


def discrete_gamma_rates(alpha: float, beta: float, k: int, method: str = "mean") -> np.ndarray:
    factor = alpha / beta * k
    if method == "median":
        m = median_boundaries(alpha, beta, k)
        return m * (factor / m.sum())
    m = mean_boundaries(alpha, beta, k)
    for i in range(k - 1):
        m[i] = gammainc(alpha + 1.0, m[i] * beta)
    m[k - 1] = 1.0
    for i in range(k - 1, 0, -1):
        m[i] -= m[i - 1]
        m[i] *= factor
    m[0] *= factor
    return m


def discrete_gamma_rates_dalpha(alpha: float, k: int, method: str = "mean", rel_step: float = 1e-6) -> np.ndarray:
    """d discrete_gamma_rates(alpha, alpha, k) / d alpha (the usual one-parameter Gamma: shape = rate, mean 1),
    by a central difference -- the chain-rule factor that turns the library's d logL / d rates
    (mcp_eval_rate_gradient) into d logL / d alpha.  Not in the reference, which samples alpha gradient-free."""
    h = rel_step * max(abs(alpha), 1e-3)
    return (discrete_gamma_rates(alpha + h, alpha + h, k, method) - discrete_gamma_rates(alpha - h, alpha - h, k, method)) / (2.0 * h)
