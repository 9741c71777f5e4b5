"""ctypes front end of oracle/felsenstein_oracle.c (test infrastructure, see its header)."""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SRC = os.path.join(_HERE, "felsenstein_oracle.c")
_LIB = os.path.join(_HERE, "liboracle.so")
_lib = None

_dp = np.ctypeslib.ndpointer(dtype=np.float64, flags="C_CONTIGUOUS")
_ip = np.ctypeslib.ndpointer(dtype=np.int32, flags="C_CONTIGUOUS")
_up = np.ctypeslib.ndpointer(dtype=np.uint8, flags="C_CONTIGUOUS")


def build(force: bool = False) -> str:
    """Compile liboracle.so next to the source if missing or stale."""
    if force or not os.path.exists(_LIB) or os.path.getmtime(_LIB) < os.path.getmtime(_SRC):
        cmd = ["/usr/bin/gcc", "-O3", "-march=x86-64-v3", "-fopenmp", "-fPIC", "-std=c11", "-shared",
               "-o", _LIB, _SRC, "-lm"]
        subprocess.run(cmd, check=True, cwd=_HERE)
    return _LIB


def _load():
    global _lib
    if _lib is None:
        build()
        lib = C.CDLL(_LIB)
        lib.oracle_num_threads.restype = C.c_int
        lib.oracle_transition.restype = C.c_int
        lib.oracle_transition.argtypes = [C.c_int, C.c_int, C.c_int, _dp, _dp, _dp, C.c_double,
                                          _dp, _dp, _dp, C.c_void_p]
        lib.oracle_felsenstein.restype = C.c_int
        lib.oracle_felsenstein.argtypes = [C.c_int, C.c_int64, C.c_int, C.c_int, _dp, _ip, _ip, _dp,
                                           _dp, _dp, _dp, C.c_double, _dp, _dp, C.c_int, C.c_int,
                                           C.POINTER(C.c_double), C.c_void_p]
        lib.oracle_codes_to_dense.restype = C.c_int
        lib.oracle_codes_to_dense.argtypes = [C.c_int, C.c_int64, C.c_int, _up, _ip, C.c_int, _dp]
        _lib = lib
    return _lib


def num_threads() -> int:
    return int(_load().oracle_num_threads())


def _f(a):
    # column-major (Fortran) matrices are passed as their flat memory
    return np.ascontiguousarray(np.asarray(a, dtype=np.float64).ravel(order="F"))


def transition(U, D, Uinv, mu, rates, blv, want_dP=False):
    """P (K,K,R,NB) [and dP] as Fortran-ordered arrays, P[s_parent, s_child, r, b]."""
    lib = _load()
    K = len(D)
    rates = np.ascontiguousarray(rates, dtype=np.float64)
    blv = np.ascontiguousarray(blv, dtype=np.float64)
    R, NB = rates.size, blv.size
    P = np.zeros(K * K * R * NB, dtype=np.float64)
    dP = np.zeros(K * K * R * NB, dtype=np.float64) if want_dP else None
    rc = lib.oracle_transition(K, R, NB, _f(U), _f(D), _f(Uinv), float(mu), rates, blv, P,
                               dP.ctypes.data if want_dP else None)
    if rc:
        raise RuntimeError(f"oracle_transition failed ({rc})")
    P = P.reshape((K, K, R, NB), order="F")
    if want_dP:
        return P, dP.reshape((K, K, R, NB), order="F")
    return P


def codes_to_dense(codes, leaf_nums, K, NN):
    lib = _load()
    codes = np.ascontiguousarray(codes, dtype=np.uint8)
    leaf_nums = np.ascontiguousarray(leaf_nums, dtype=np.int32)
    n_leaves, S = codes.shape
    x = np.zeros(K * S * NN, dtype=np.float64)
    rc = lib.oracle_codes_to_dense(K, S, NN, codes, leaf_nums, n_leaves, x)
    if rc:
        raise RuntimeError(f"oracle_codes_to_dense failed ({rc})")
    return x.reshape((K, S, NN), order="F")


def felsenstein(x, postorder_num, parent_num, blv, U, D, Uinv, mu, rates, pi, want_grad=True,
                nthreads=0):
    """(ll, grad) with grad indexed by num-1 (None when want_grad is False).
    x is the dense (K,S,NN) array (Fortran order, as datafortree returns it)."""
    lib = _load()
    x = np.asarray(x, dtype=np.float64)
    K, S, NN = x.shape
    xf = np.ascontiguousarray(x.ravel(order="F"))
    rates = np.ascontiguousarray(rates, dtype=np.float64)
    pi = np.ascontiguousarray(pi, dtype=np.float64)
    blv = np.ascontiguousarray(blv, dtype=np.float64)
    po = np.ascontiguousarray(postorder_num, dtype=np.int32)
    pa = np.ascontiguousarray(parent_num, dtype=np.int32)
    assert po.size == NN and pa.size == NN and blv.size == NN - 1 and pi.size == K
    ll = C.c_double(0.0)
    grad = np.zeros(max(NN - 1, 1), dtype=np.float64)
    rc = lib.oracle_felsenstein(K, S, rates.size, NN, xf, po, pa, blv, _f(U), _f(D), _f(Uinv),
                                float(mu), rates, pi, int(bool(want_grad)), int(nthreads),
                                C.byref(ll), grad.ctypes.data if want_grad else None)
    if rc:
        raise RuntimeError(f"oracle_felsenstein failed ({rc})")
    return ll.value, (grad[:NN - 1] if want_grad else None)


# ---------------------------------------------------------------------------------------------
# Branch-length prior of the tree node.  Test infrastructure like the rest
# of this file.  Restates internal_logpdf, /root/reference/src/Likelihood/Prior.jl:1-37, as the
# same scalar loop; the reference obtains the gradient by reverse-mode AD of that function
# (Prior.jl:39-48), which is reproduced here to machine precision by complex-step differentiation
# of the restated loop (no hand-derived formula on the checking side).
# Pinned by test/distributions/treedists.jl:61-69 (tests/test_prior_host.py, test_oracle_prior).
# ---------------------------------------------------------------------------------------------
def compound_dirichlet_logpdf(alpha, a, beta, c, b_lens, int_leave_map):
    import cmath
    import math

    blen_int = blen_leave = blen_int_log = blen_leave_log = 0.0
    nterm = 0.0
    for i in range(len(int_leave_map)):
        if int_leave_map[i] == 1:
            blen_int += b_lens[i]
            blen_int_log += cmath.log(b_lens[i])
        else:
            blen_leave += b_lens[i]
            blen_leave_log += cmath.log(b_lens[i])
            nterm += 1
    t_l = blen_int + blen_leave
    n_int = nterm - 3.0
    first = (alpha * math.log(beta)) - math.log(math.gamma(alpha)) - (t_l * beta)
    second = -math.log(math.gamma(a)) - math.log(math.gamma(c)) + math.log(math.gamma(a + c))
    third = blen_leave_log * (a - 1.0) + blen_int_log * (a * c - 1.0)
    fourth = (alpha - a * nterm - a * c * n_int) * cmath.log(t_l)
    return first + second + third + fourth


def compound_dirichlet_gradlogpdf(alpha, a, beta, c, b_lens, int_leave_map):
    b = [complex(v) for v in b_lens]
    val = compound_dirichlet_logpdf(alpha, a, beta, c, b, int_leave_map).real
    h = 1e-30
    grad = np.zeros(len(b))
    for j in range(len(b)):
        bj = list(b)
        bj[j] = b[j] + 1j * h
        grad[j] = compound_dirichlet_logpdf(alpha, a, beta, c, bj, int_leave_map).imag / h
    return val, grad


def exponential_bl_gradlogpdf(scale, b_lens):
    import math

    b = np.asarray(b_lens, dtype=np.float64)           # Prior.jl:50-57, 80-83
    return float(np.sum(-math.log(scale) - b / scale)), np.full(b.size, -1.0 / scale)
