"""ctypes bindings for oracle/liboracle.so -- TEST INFRASTRUCTURE ONLY.

Importable from tests/, __graft_entry__.smoke() and bench.py's CPU-baseline legs;
never from the product path (ipx_b200/).
"""

import ctypes as C
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
LIB = os.path.join(HERE, "liboracle.so")

_dp = C.POINTER(C.c_double)
_ip = C.POINTER(C.c_int64)
oint = C.c_int64


def build():
    subprocess.check_call(["make", "-s", "-C", HERE, "oracle"])


def _d(a):
    return None if a is None else a.ctypes.data_as(_dp)


def _i(a):
    return None if a is None else a.ctypes.data_as(_ip)


def _f64(a):
    return None if a is None else np.ascontiguousarray(a, dtype=np.float64)


def _i64(a):
    return None if a is None else np.ascontiguousarray(a, dtype=np.int64)


class _Split(C.Structure):
    _fields_ = [("dim", oint), ("Lp", _ip), ("Li", _ip), ("Lx", _dp), ("Up", _ip), ("Ui", _ip),
                ("Ux", _dp), ("ncolN", oint), ("Np", _ip), ("Ni", _ip), ("Nx", _dp),
                ("num_free", oint), ("free_positions", _ip)]


class _Operator(C.Structure):
    _fields_ = [("kind", C.c_int), ("m", oint), ("n", oint), ("Ap", _ip), ("Ai", _ip),
                ("Ax", _dp), ("W", _dp), ("split", C.POINTER(_Split)), ("work", _dp)]


_lib = None


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(LIB):
            build()
        _lib = C.CDLL(LIB)
        _lib.orc_triangular_solve.restype = oint
        _lib.orc_pcr_solve.restype = oint
        _lib.orc_cr_solve.restype = oint
        _lib.orc_kktdiag_solve.restype = oint
    return _lib


class Csc:
    """Keeps int64/f64 contiguous copies alive."""

    def __init__(self, Ap, Ai, Ax):
        self.Ap, self.Ai, self.Ax = _i64(Ap), _i64(Ai), _f64(Ax)

    @property
    def args(self):
        return _i(self.Ap), _i(self.Ai), _d(self.Ax)


def normal_apply(m, n, A, W, rhs, want_dot=True):
    W, rhs, lhs, dot = _f64(W), _f64(rhs), np.empty(m), C.c_double(np.nan)
    lib().orc_normal_apply(oint(m), oint(n), *A.args, _d(W), _d(rhs), _d(lhs),
                           C.byref(dot) if want_dot else None)
    return lhs, dot.value


def diag_build(m, n, A, W):
    W, diag = _f64(W), np.empty(m)
    lib().orc_diag_build(oint(m), oint(n), *A.args, _d(W), _d(diag))
    return diag


def diag_apply(diag, rhs):
    diag, rhs = _f64(diag), _f64(rhs)
    lhs, dot = np.empty(len(rhs)), C.c_double()
    lib().orc_diag_apply(oint(len(rhs)), _d(diag), _d(rhs), _d(lhs), C.byref(dot))
    return lhs, dot.value


def add_normal_product(nrow, ncol, A, D, rhs, lhs):
    D, rhs, lhs = _f64(D), _f64(rhs), _f64(lhs).copy()
    lib().orc_add_normal_product(oint(nrow), oint(ncol), *A.args, _d(D), _d(rhs), _d(lhs))
    return lhs


def multiply_add(nrow, ncol, A, rhs, alpha, lhs, trans):
    """ipx::MultiplyAdd (reference src/sparse_matrix.cc:194-209); returns the updated lhs."""
    rhs, lhs = _f64(rhs), _f64(lhs).copy()
    lib().orc_multiply_add(oint(nrow), oint(ncol), *A.args, _d(rhs), C.c_double(alpha), _d(lhs),
                           C.c_char(trans.encode()))
    return lhs


def triangular_solve(dim, A, x, trans, uplo, unitdiag):
    x = _f64(x).copy()
    nz = lib().orc_triangular_solve(oint(dim), *A.args, _d(x), C.c_char(trans.encode()),
                                    C.c_char(uplo.encode()), C.c_int(unitdiag))
    return x, nz


class SplitOperator:
    """Prepared factors of C = I + inv(B) N N' inv(B')."""

    def __init__(self, dim, L, U, N, ncolN, free_positions):
        self.dim, self.L, self.U, self.N = dim, L, U, N
        self.free = _i64(free_positions)
        self.work = np.zeros(max(dim, 1))
        self.s = _Split(oint(dim), *L.args, *U.args, oint(ncolN), *N.args, oint(len(self.free)),
                        _i(self.free))

    def apply(self, rhs, want_dot=True):
        rhs, lhs, dot = _f64(rhs), np.empty(self.dim), C.c_double(np.nan)
        lib().orc_split_apply(C.byref(self.s), _d(rhs), _d(lhs), _d(self.work),
                              C.byref(dot) if want_dot else None)
        return lhs, dot.value

    def operator(self):
        op = _Operator()
        op.kind = 1
        op.split = C.pointer(self.s)
        op.work = _d(self.work)
        return op


def normal_operator(m, n, A, W):
    op = _Operator()
    op.kind, op.m, op.n = 0, m, n
    op.Ap, op.Ai, op.Ax = A.args
    op._W = _f64(W)
    op.W = _d(op._W)
    return op


def pcr_solve(op, m, diag, rhs, tol, resscale, maxiter, lhs0=None, hist_cap=0):
    diag, rhs, resscale = _f64(diag), _f64(rhs), _f64(resscale)
    lhs = np.zeros(m) if lhs0 is None else _f64(lhs0).copy()
    it = oint(0)
    hist = np.full(max(hist_cap, 1), np.nan)
    err = lib().orc_pcr_solve(C.byref(op), oint(m), _d(diag), _d(rhs), C.c_double(tol),
                              _d(resscale), oint(maxiter), _d(lhs), C.byref(it),
                              _d(hist) if hist_cap else None, oint(hist_cap))
    return lhs, {"errflag": int(err), "iter": it.value, "hist": hist[:min(hist_cap, it.value + 1)]}


def cr_solve(op, m, rhs, tol, resscale, maxiter, lhs0=None, hist_cap=0):
    rhs, resscale = _f64(rhs), _f64(resscale)
    lhs = np.zeros(m) if lhs0 is None else _f64(lhs0).copy()
    it = oint(0)
    hist = np.full(max(hist_cap, 1), np.nan)
    err = lib().orc_cr_solve(C.byref(op), oint(m), _d(rhs), C.c_double(tol), _d(resscale),
                             oint(maxiter), _d(lhs), C.byref(it),
                             _d(hist) if hist_cap else None, oint(hist_cap))
    return lhs, {"errflag": int(err), "iter": it.value, "hist": hist[:min(hist_cap, it.value + 1)]}


def kktdiag_weights(m, n, xl=None, xu=None, zl=None, zu=None, mu=0.0):
    W, resscale = np.empty(n + m), np.empty(m)
    have = xl is not None
    a = [_f64(v) for v in (xl, xu, zl, zu)]
    lib().orc_kktdiag_weights(oint(m), oint(n), C.c_int(1 if have else 0), *[_d(v) for v in a],
                              C.c_double(mu), _d(W), _d(resscale))
    return W, resscale


def kktdiag_solve(m, n, A, W, diag, resscale, a, b, tol, maxiter):
    W, diag, resscale, a, b = (_f64(v) for v in (W, diag, resscale, a, b))
    x, y, it = np.empty(n + m), np.empty(m), oint(0)
    err = lib().orc_kktdiag_solve(oint(m), oint(n), *A.args, _d(W), _d(diag), _d(resscale), _d(a),
                                  _d(b), C.c_double(tol), oint(maxiter), _d(x), _d(y),
                                  C.byref(it))
    return x, y, {"errflag": int(err), "iter": it.value}


# ---- Maxvolume column sweeps (numpy restatement; reference src/maxvolume.cc) ----

def column_dots(AIp, AIi, AIx, x):
    """AI[:,j]'x for every column, each summed in storage order like DotColumn
    (reference src/sparse_matrix.h:136-143)."""
    AIp, AIi = _i64(AIp), _i64(AIi)
    prod = _f64(x)[AIi] * _f64(AIx)
    ncol = len(AIp) - 1
    length = np.diff(AIp)
    out = np.zeros(ncol)
    for k in range(int(length.max()) if ncol else 0):
        sel = np.nonzero(length > k)[0]
        out[sel] = out[sel] + prod[AIp[sel] + k]
    return out


def find_largest(weights):
    """FindLargest (reference src/maxvolume.cc:170-200): (jmax2, jmax) of an ascending scan
    with strict comparisons, both starting at column 0 with weight 0."""
    w = np.abs(_f64(weights))
    jmax = jmax2 = 0
    wmax = wmax2 = 0.0
    for j in range(len(w)):
        if w[j] > wmax:
            wmax2, jmax2 = wmax, jmax
            wmax, jmax = w[j], j
        elif w[j] > wmax2:
            wmax2, jmax2 = w[j], j
    return jmax2, jmax


def maxvol_weights(AIp, AIi, AIx, colscale, work):
    """Column weights at the top of Maxvolume::Driver (reference src/maxvolume.cc:220-231)."""
    cs = _f64(colscale)
    return np.where(cs != 0.0, column_dots(AIp, AIi, AIx, work) * cs, 0.0)


def maxvol_update(AIp, AIi, AIx, colscale, colweights, btran, alpha, jb, colscale_jb,
                  colweight_jb, jn):
    """Weight update after the exchange of jb and jn (reference src/maxvolume.cc:291-308) with
    the tableau row of jb over the nonbasic columns from Basis::TableauRow's dense branch
    (src/basis.cc:266-279); jb is still basic when the row is formed, so its entry is 0."""
    cs, cw = _f64(colscale).copy(), _f64(colweights).copy()
    row = np.where(cs != 0.0, column_dots(AIp, AIi, AIx, btran), 0.0)
    row[jb] = 0.0
    cs[jb] = colscale_jb
    cs[jn] = 0.0
    cw = cw + (alpha * row) * cs
    cw[jb] = colweight_jb
    cw[jn] = 0.0
    return cs, cw
