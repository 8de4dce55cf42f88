"""ctypes bindings for a built IPX shared library.

Two libraries export the same symbols (ipx_b200/host/ipx_harness.cc plus the
unchanged reference C API, reference include/ipx_c.h:13-62):

* ``oracle/_ref/libipx_ref.so``     - the reference's own CPU code (oracle / CPU baseline)
* ``ipx_b200/_build/libipx_gpu.so`` - the same with the hot-path TUs replaced by the
  GPU drop-ins

so a test drives both through the same calls. This module is harness code, not
part of the product path.
"""

import ctypes as C
import os

import numpy as np

ipxint = C.c_int64
_dp = C.POINTER(C.c_double)
_ip = C.POINTER(C.c_int64)

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF_LIB = os.path.join(REPO, "oracle", "_ref", "libipx_ref.so")
GPU_LIB = os.path.join(REPO, "ipx_b200", "_build", "libipx_gpu.so")


class Parameters(C.Structure):
    """struct ipx_parameters (reference include/ipx_parameters.h:6-50)."""
    _fields_ = [
        ("display", ipxint), ("logfile", C.c_char_p), ("print_interval", C.c_double),
        ("time_limit", C.c_double), ("dualize", ipxint), ("scale", ipxint),
        ("ipm_maxiter", ipxint), ("ipm_feasibility_tol", C.c_double),
        ("ipm_optimality_tol", C.c_double), ("ipm_drop_primal", C.c_double),
        ("ipm_drop_dual", C.c_double), ("kkt_tol", C.c_double),
        ("precond_dense_cols", ipxint), ("crash_basis", ipxint),
        ("dependency_tol", C.c_double), ("volume_tol", C.c_double),
        ("rows_per_slice", ipxint), ("maxskip_updates", ipxint), ("lu_kernel", ipxint),
        ("lu_pivottol", C.c_double), ("crossover", ipxint), ("crossover_start", C.c_double),
        ("pfeasibility_tol", C.c_double), ("dfeasibility_tol", C.c_double),
        ("debug", ipxint), ("switchiter", ipxint), ("stop_at_switch", ipxint),
        ("update_heuristic", ipxint), ("maxpasses", ipxint),
    ]


_INFO_INT = ("status status_ipm status_crossover errflag num_var num_constr num_entries "
             "num_rows_solver num_cols_solver num_entries_solver dualized dense_cols "
             "dependent_rows dependent_cols rows_inconsistent cols_inconsistent "
             "primal_dropped dual_dropped").split()
_INFO_DBL1 = ("abs_presidual abs_dresidual rel_presidual rel_dresidual pobjval dobjval "
              "rel_objgap complementarity normx normy normz objval primal_infeas "
              "dual_infeas").split()
_INFO_INT2 = ("iter kktiter1 kktiter2 basis_repairs updates_start updates_ipm "
              "updates_crossover").split()
_INFO_DBL2 = ("time_total time_ipm1 time_ipm2 time_starting_basis time_crossover "
              "time_kkt_factorize time_kkt_solve time_maxvol time_cr1 time_cr1_AAt "
              "time_cr1_pre time_cr2 time_cr2_NNt time_cr2_B time_cr2_Bt ftran_sparse "
              "btran_sparse time_ftran time_btran time_lu_invert time_lu_update mean_fill "
              "max_fill time_symb_invert").split()
_INFO_INT3 = "maxvol_updates maxvol_skipped maxvol_passes tbl_nnz".split()
_INFO_DBL3 = "tbl_max frobnorm_squared lambdamax volume_increase".split()


class Info(C.Structure):
    """struct ipx_info (reference include/ipx_info.h:6-100)."""
    _fields_ = ([(k, ipxint) for k in _INFO_INT] + [(k, C.c_double) for k in _INFO_DBL1] +
                [(k, ipxint) for k in _INFO_INT2] + [(k, C.c_double) for k in _INFO_DBL2] +
                [(k, ipxint) for k in _INFO_INT3] + [(k, C.c_double) for k in _INFO_DBL3])

    def asdict(self):
        return {k: getattr(self, k) for k, _ in self._fields_}


INFO_OUT_KEYS = ("errflag kktiter1 kktiter2 time_cr1 time_cr1_AAt time_cr1_pre time_cr2 "
                 "time_cr2_NNt time_cr2_B time_cr2_Bt time_kkt_factorize time_kkt_solve "
                 "updates_ipm primal_dropped dual_dropped time_maxvol").split()


def _d(a):
    return None if a is None else a.ctypes.data_as(_dp)


def _i(a):
    return None if a is None else a.ctypes.data_as(_ip)


def _f64(a):
    return None if a is None else np.ascontiguousarray(a, dtype=np.float64)


def _i64(a):
    return None if a is None else np.ascontiguousarray(a, dtype=np.int64)


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
                     "ipxh_triangular_solve", "ipx_load_model", "ipx_solve",
                     "ipx_get_interior_solution", "ipx_get_basic_solution"):
            getattr(L, name).restype = ipxint
        L.ipx_default_parameters.restype = Parameters
        L.ipx_get_parameters.restype = Parameters
        L.ipx_get_info.restype = Info

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

    def kktbasis_solve(self, a, b, tol):
        a, b = _f64(a), _f64(b)
        x, y, out = np.empty(self.n + self.m), np.empty(self.m), np.zeros(16)
        err = self.lib.ipxh_kktbasis_solve(self.h, _d(a), _d(b), C.c_double(tol), _d(x), _d(y),
                                           _d(out))
        return x, y, dict(zip(INFO_OUT_KEYS, out.tolist()), err=err)


class LpSolver:
    """ipx::LpSolver through the unchanged C API (reference include/ipx_c.h)."""

    def __init__(self, ipxlib):
        self.lib = ipxlib.lib
        self.h = C.c_void_p()
        self.lib.ipx_new(C.byref(self.h))
        self.num_var = self.num_constr = 0

    def close(self):
        if self.h:
            self.lib.ipx_free(C.byref(self.h))

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def set_parameters(self, **params):
        p = self.lib.ipx_get_parameters(self.h)
        for k, v in params.items():
            setattr(p, k, v)
        self.lib.ipx_set_parameters(self.h, p)

    def load_model(self, lp):
        self._keep = [_i64(lp.Ap), _i64(lp.Ai), _f64(lp.Ax), _f64(lp.rhs), _f64(lp.obj),
                      _f64(lp.lb), _f64(lp.ub)]
        Ap, Ai, Ax, rhs, obj, lb, ub = self._keep
        self.num_var, self.num_constr = lp.n, lp.m
        return self.lib.ipx_load_model(self.h, ipxint(lp.n), _d(obj), _d(lb), _d(ub), ipxint(lp.m),
                                       _i(Ap), _i(Ai), _d(Ax), _d(rhs), C.c_char_p(lp.constr_type))

    def solve(self):
        return self.lib.ipx_solve(self.h)

    def info(self):
        return self.lib.ipx_get_info(self.h).asdict()

    def interior_solution(self):
        n, m = self.num_var, self.num_constr
        x, xl, xu, zl, zu = (np.empty(n) for _ in range(5))
        slack, y = np.empty(m), np.empty(m)
        err = self.lib.ipx_get_interior_solution(self.h, _d(x), _d(xl), _d(xu), _d(slack), _d(y),
                                                 _d(zl), _d(zu))
        return err, dict(x=x, xl=xl, xu=xu, slack=slack, y=y, zl=zl, zu=zu)
