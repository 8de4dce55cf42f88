"""ctypes bindings of the TEST HARNESS (oracle/ipx_harness.cc) over a built IPX.

Two libraries export the same harness symbols next to the unchanged reference C API
(reference include/ipx_c.h:13-62):

* ``oracle/_ref/libipx_ref.so``          - the reference's own CPU code (oracle / CPU baseline),
  harness linked in
* ``oracle/_ref/libipx_gpu_harness.so``  - the harness alone, linked AGAINST the product
  ``ipx_b200/_build/libipx_gpu.so`` (the drop-in build; it contains no test code)

so a test drives IPX's own classes on both arms through the same calls. Test infrastructure:
only tests/, tools/ and bench.py's reference legs import this module.
"""

import ctypes as C
import os

import numpy as np

from ipx_b200.ipxc import (Info, LpSolver, Parameters, _d, _f64, _i, _i64,  # noqa: F401
                           declare_c_api, ipxint)

_dp = C.POINTER(C.c_double)
_ip = C.POINTER(C.c_int64)

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF_LIB = os.path.join(REPO, "oracle", "_ref", "libipx_ref.so")
GPU_LIB = os.path.join(REPO, "oracle", "_ref", "libipx_gpu_harness.so")


INFO_OUT_KEYS = ("errflag kktiter1 kktiter2 time_cr1 time_cr1_AAt time_cr1_pre time_cr2 "
                 "time_cr2_NNt time_cr2_B time_cr2_Bt time_kkt_factorize time_kkt_solve "
                 "updates_ipm primal_dropped dual_dropped time_maxvol").split()


class IpxLibrary:
    """A loaded IPX build (reference or GPU drop-in)."""

    def __init__(self, path):
        if not os.path.exists(path):
            raise FileNotFoundError(f"{path} is not built (run python -c 'import "
                                    "__graft_entry__ as g; g.build()')")
        self.path = path
        self.lib = C.CDLL(path, mode=os.RTLD_LOCAL | os.RTLD_NOW)
        L = self.lib
        L.ipxh_create.restype = C.c_void_p
        L.ipxh_normal_apply_timed.restype = C.c_double
        for name in ("ipxh_diag_factorize", "ipxh_kktdiag_factorize", "ipxh_kktdiag_solve",
                     "ipxh_kktdiag_iter", "ipxh_basis_load", "ipxh_basis_from_weights",
                     "ipxh_kktbasis_factorize", "ipxh_kktbasis_solve",
                     "ipxh_triangular_solve", "ipxh_maxvolume"):
            getattr(L, name).restype = ipxint
        declare_c_api(L)  # the public API, from the same handle (the drop-in build is a dependency)

    def default_parameters(self):
        return self.lib.ipx_default_parameters()

    def model(self, lp, **params):
        return IpxModel(self, lp, **params)

    def lp_solver(self):
        return LpSolver(self)

    def triangular_solve(self, dim, Ap, Ai, Ax, x, trans, uplo, unitdiag):
        """ipx::TriangularSolve (reference src/sparse_matrix.cc:224-301)."""
        Ap, Ai, Ax, x = _i64(Ap), _i64(Ai), _f64(Ax), _f64(x).copy()
        nz = self.lib.ipxh_triangular_solve(
            ipxint(dim), _i(Ap), _i(Ai), _d(Ax), _d(x), C.c_char(trans.encode()),
            C.c_char(uplo.encode()), C.c_int(unitdiag))
        return x, nz

    def add_normal_product(self, nrow, ncol, Ap, Ai, Ax, D, rhs, lhs):
        """ipx::AddNormalProduct (reference src/sparse_matrix.cc:211-222)."""
        Ap, Ai, Ax, D, rhs, lhs = _i64(Ap), _i64(Ai), _f64(Ax), _f64(D), _f64(rhs), _f64(lhs).copy()
        self.lib.ipxh_add_normal_product(ipxint(nrow), ipxint(ncol), _i(Ap), _i(Ai), _d(Ax),
                                         _d(D), _d(rhs), _d(lhs))
        return lhs


class IpxModel:
    """Control + UserModel + Presolver + Model and the hot-path objects on it."""

    def __init__(self, ipxlib, lp, **params):
        self.ipxlib = ipxlib
        self.lib = ipxlib.lib
        p = ipxlib.default_parameters()
        p.display = 0
        p.dualize = 0
        for k, v in params.items():
            setattr(p, k, v)
        self._keep = [_i64(lp.Ap), _i64(lp.Ai), _f64(lp.Ax), _f64(lp.rhs), _f64(lp.obj),
                      _f64(lp.lb), _f64(lp.ub)]
        Ap, Ai, Ax, rhs, obj, lb, ub = self._keep
        err = ipxint(0)
        self.h = C.c_void_p(self.lib.ipxh_create(
            ipxint(lp.m), ipxint(lp.n), _i(Ap), _i(Ai), _d(Ax), _d(rhs),
            C.c_char_p(lp.constr_type), _d(obj), _d(lb), _d(ub), C.byref(p), C.byref(err)))
        if not self.h:
            raise RuntimeError(f"ipxh_create failed, errflag {err.value}")
        m, n, nnz = ipxint(), ipxint(), ipxint()
        self.lib.ipxh_dims(self.h, C.byref(m), C.byref(n), C.byref(nnz))
        self.m, self.n, self.nnz_AI = m.value, n.value, nnz.value

    def close(self):
        if self.h:
            self.lib.ipxh_free(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def AI(self):
        Ap = np.empty(self.n + self.m + 1, np.int64)
        Ai = np.empty(self.nnz_AI, np.int64)
        Ax = np.empty(self.nnz_AI, np.float64)
        self.lib.ipxh_get_AI(self.h, _i(Ap), _i(Ai), _d(Ax))
        return Ap, Ai, Ax

    def model_vectors(self):
        b = np.empty(self.m)
        c, lb, ub = (np.empty(self.n + self.m) for _ in range(3))
        self.lib.ipxh_get_model_vectors(self.h, _d(b), _d(c), _d(lb), _d(ub))
        return b, c, lb, ub

    # NormalMatrix
    def multiply_add_AI(self, rhs, alpha, lhs, trans):
        """ipx::MultiplyAdd(model.AI(), rhs, alpha, lhs, trans); returns the updated lhs."""
        rhs, lhs = _f64(rhs), _f64(lhs).copy()
        self.lib.ipxh_multiply_add_AI(self.h, _d(rhs), C.c_double(alpha), _d(lhs),
                                      C.c_char(trans.encode()))
        return lhs

    def normal_prepare(self, W):
        W = _f64(W)
        self.lib.ipxh_normal_prepare(self.h, _d(W))

    def normal_apply(self, rhs, want_dot=True):
        rhs, lhs, dot = _f64(rhs), np.empty(self.m), C.c_double(np.nan)
        self.lib.ipxh_normal_apply(self.h, _d(rhs), _d(lhs), C.byref(dot) if want_dot else None)
        return lhs, dot.value

    def normal_apply_timed(self, rhs, reps):
        rhs, lhs = _f64(rhs), np.empty(self.m)
        t = self.lib.ipxh_normal_apply_timed(self.h, _d(rhs), _d(lhs), ipxint(reps))
        return lhs, t

    # DiagonalPrecond
    def diag_factorize(self, W, precond_dense_cols=1):
        W = _f64(W)
        return self.lib.ipxh_diag_factorize(self.h, _d(W), ipxint(precond_dense_cols))

    def diag_apply(self, rhs, want_dot=True):
        rhs, lhs, dot = _f64(rhs), np.empty(self.m), C.c_double(np.nan)
        self.lib.ipxh_diag_apply(self.h, _d(rhs), _d(lhs), C.byref(dot) if want_dot else None)
        return lhs, dot.value

    # ConjugateResiduals
    def pcr_solve(self, rhs, tol, resscale, maxiter, lhs0=None):
        rhs, resscale = _f64(rhs), _f64(resscale)
        lhs = np.zeros(self.m) if lhs0 is None else _f64(lhs0).copy()
        out = np.zeros(3)
        self.lib.ipxh_pcr_solve(self.h, _d(rhs), C.c_double(tol), _d(resscale), ipxint(maxiter),
                                _d(lhs), _d(out))
        return lhs, {"errflag": int(out[0]), "iter": int(out[1]), "time": out[2]}

    def cr_solve_normal(self, rhs, tol, resscale, maxiter, lhs0=None):
        rhs, resscale = _f64(rhs), _f64(resscale)
        lhs = np.zeros(self.m) if lhs0 is None else _f64(lhs0).copy()
        out = np.zeros(3)
        self.lib.ipxh_cr_solve_normal(self.h, _d(rhs), C.c_double(tol), _d(resscale),
                                      ipxint(maxiter), _d(lhs), _d(out))
        return lhs, {"errflag": int(out[0]), "iter": int(out[1]), "time": out[2]}

    # Iterate / KKTSolverDiag
    def iterate_set(self, x, xl, xu, y, zl, zu):
        a = [_f64(v) for v in (x, xl, xu, y, zl, zu)]
        self.lib.ipxh_iterate_set(self.h, *[_d(v) for v in a])

    def kktdiag_maxiter(self, maxiter):
        self.lib.ipxh_kktdiag_maxiter(self.h, ipxint(maxiter))

    def kktdiag_factorize(self, use_iterate):
        return self.lib.ipxh_kktdiag_factorize(self.h, ipxint(1 if use_iterate else 0))

    def kktdiag_solve(self, a, b, tol):
        a, b = _f64(a), _f64(b)
        x, y, out = np.empty(self.n + self.m), np.empty(self.m), np.zeros(16)
        err = self.lib.ipxh_kktdiag_solve(self.h, _d(a), _d(b), C.c_double(tol), _d(x), _d(y),
                                          _d(out))
        return x, y, dict(zip(INFO_OUT_KEYS, out.tolist()), err=err)

    # Basis / SplittedNormalMatrix / KKTSolverBasis
    def basis_load(self, basic_status):
        s = np.ascontiguousarray(basic_status, dtype=np.int32)
        return self.lib.ipxh_basis_load(self.h, s.ctypes.data_as(C.POINTER(C.c_int)))

    def basis_from_weights(self, colweights):
        w = _f64(colweights)
        return self.lib.ipxh_basis_from_weights(self.h, _d(w))

    def basis_get(self):
        basis = np.empty(self.m, np.int64)
        status = np.empty(self.n + self.m, np.int32)
        self.lib.ipxh_basis_get(self.h, _i(basis), status.ctypes.data_as(C.POINTER(C.c_int)))
        return basis, status

    def basis_free_variable(self, j):
        self.lib.ipxh_basis_free_variable(self.h, ipxint(j))

    def basis_fix_variable(self, j):
        self.lib.ipxh_basis_fix_variable(self.h, ipxint(j))

    def basis_lu(self):
        lnz, unz = ipxint(), ipxint()
        self.lib.ipxh_basis_lu_sizes(self.h, C.byref(lnz), C.byref(unz))
        m = self.m
        Lp, Up = np.empty(m + 1, np.int64), np.empty(m + 1, np.int64)
        Li, Lx = np.empty(lnz.value, np.int64), np.empty(lnz.value)
        Ui, Ux = np.empty(unz.value, np.int64), np.empty(unz.value)
        rowperm, colperm = np.empty(m, np.int64), np.empty(m, np.int64)
        self.lib.ipxh_basis_lu(self.h, _i(Lp), _i(Li), _d(Lx), _i(Up), _i(Ui), _d(Ux),
                               _i(rowperm), _i(colperm))
        return (Lp, Li, Lx), (Up, Ui, Ux), rowperm, colperm

    def basis_solve_dense(self, rhs, trans):
        rhs, lhs = _f64(rhs), np.empty(self.m)
        self.lib.ipxh_basis_solve_dense(self.h, _d(rhs), _d(lhs), C.c_char(trans.encode()))
        return lhs

    def split_prepare(self, colscale):
        cs = _f64(colscale)
        self.lib.ipxh_split_prepare(self.h, _d(cs))

    def split_colperm(self):
        cp = np.empty(self.m, np.int64)
        self.lib.ipxh_split_colperm(self.h, _i(cp))
        return cp

    def split_apply(self, rhs, want_dot=True):
        rhs, lhs, dot = _f64(rhs), np.empty(self.m), C.c_double(np.nan)
        self.lib.ipxh_split_apply(self.h, _d(rhs), _d(lhs), C.byref(dot) if want_dot else None)
        return lhs, dot.value

    def split_apply_timed(self, rhs, reps):
        rhs, lhs, times = _f64(rhs), np.empty(self.m), np.zeros(3)
        self.lib.ipxh_split_apply_timed(self.h, _d(rhs), _d(lhs), ipxint(reps), _d(times))
        return lhs, dict(B=times[0], Bt=times[1], NNt=times[2])

    def cr_solve_split(self, rhs, tol, maxiter, lhs0=None):
        rhs = _f64(rhs)
        lhs = np.zeros(self.m) if lhs0 is None else _f64(lhs0).copy()
        out = np.zeros(3)
        self.lib.ipxh_cr_solve_split(self.h, _d(rhs), C.c_double(tol), ipxint(maxiter), _d(lhs),
                                     _d(out))
        return lhs, {"errflag": int(out[0]), "iter": int(out[1]), "time": out[2]}

    def kktbasis_maxiter(self, maxiter):
        self.lib.ipxh_kktbasis_maxiter(self.h, ipxint(maxiter))

    def kktbasis_factorize(self):
        out = np.zeros(16)
        err = self.lib.ipxh_kktbasis_factorize(self.h, _d(out))
        return dict(zip(INFO_OUT_KEYS, out.tolist()), err=err)

    def maxvolume(self, colscale, heuristic=True):
        """Maxvolume::RunHeuristic / RunSequential on the current basis (src/maxvolume.h)."""
        cs = None if colscale is None else _f64(colscale)
        out = np.zeros(7)
        err = self.lib.ipxh_maxvolume(self.h, _d(cs) if cs is not None else None,
                                      ipxint(1 if heuristic else 0), _d(out))
        keys = ("errflag", "updates", "skipped", "passes", "slices", "volinc", "time")
        return dict(zip(keys, out.tolist()), err=err)

    def kktbasis_solve(self, a, b, tol):
        a, b = _f64(a), _f64(b)
        x, y, out = np.empty(self.n + self.m), np.empty(self.m), np.zeros(16)
        err = self.lib.ipxh_kktbasis_solve(self.h, _d(a), _d(b), C.c_double(tol), _d(x), _d(y),
                                           _d(out))
        return x, y, dict(zip(INFO_OUT_KEYS, out.tolist()), err=err)
