"""ctypes binding of IPX's public C API (reference include/ipx_c.h:13-62) for a built IPX
shared library - here the drop-in build ipx_b200/_build/libipx_gpu.so, whose KKT-solve path runs
on the device. This is the call a Python user of IPX makes; nothing else of IPX is bound here.
"""

import ctypes as C
import os

import numpy as np

ipxint = C.c_int64
_dp = C.POINTER(C.c_double)
_ip = C.POINTER(C.c_int64)

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
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


def _d(a):
    return None if a is None else a.ctypes.data_as(_dp)


def _i(a):
    return None if a is None else a.ctypes.data_as(_ip)


def _f64(a):
    return None if a is None else np.ascontiguousarray(a, dtype=np.float64)


def _i64(a):
    return None if a is None else np.ascontiguousarray(a, dtype=np.int64)


def declare_c_api(lib):
    """Return types of the ipx_c.h functions on a loaded library handle."""
    for name in ("ipx_load_model", "ipx_solve", "ipx_get_interior_solution",
                 "ipx_get_basic_solution"):
        getattr(lib, name).restype = ipxint
    lib.ipx_default_parameters.restype = Parameters
    lib.ipx_get_parameters.restype = Parameters
    lib.ipx_get_info.restype = Info


class IpxC:
    """A loaded IPX build, public API only."""

    def __init__(self, path=GPU_LIB):
        if not os.path.exists(path):
            raise FileNotFoundError(f"{path} is not built (run python -c 'import "
                                    "__graft_entry__ as g; g.build()')")
        self.path = path
        self.lib = C.CDLL(path, mode=os.RTLD_LOCAL | os.RTLD_NOW)
        declare_c_api(self.lib)

    def lp_solver(self):
        return LpSolver(self)


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
