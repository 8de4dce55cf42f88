"""ctypes binding of the C ABI in include/ipxgpu.h (libipxgpu.so).

This is the Python view of the same boundary the C++ drop-ins
(ipx_b200/host/*_gpu.cc) call. There is no fallback: if the CUDA library is
missing or no device is present, construction raises.
"""

import ctypes as C
import os

import numpy as np

PKG = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(PKG, "_build", "libipxgpu.so")

_dp = C.POINTER(C.c_double)
_ip = C.POINTER(C.c_int64)
i64 = C.c_int64


class Options(C.Structure):
    _fields_ = [("device", C.c_int32), ("rank", C.c_int32), ("nranks", C.c_int32),
                ("col_begin", i64), ("col_end", i64), ("panel_cols", i64), ("stream", C.c_void_p)]


class MaxvolTop(C.Structure):
    _fields_ = [("jmax", i64), ("jmax2", i64), ("wmax", C.c_double), ("wmax2", C.c_double)]

    def asdict(self):
        return {k: getattr(self, k) for k, _ in self._fields_}


class CrResult(C.Structure):
    _fields_ = [("errflag", i64), ("iter", i64), ("time", C.c_double), ("time_op", C.c_double),
                ("time_pre", C.c_double), ("time_B", C.c_double), ("time_Bt", C.c_double),
                ("time_NNt", C.c_double), ("resnorm", C.c_double)]

    def asdict(self):
        return {k: getattr(self, k) for k, _ in self._fields_}


INTERRUPT_FN = C.CFUNCTYPE(i64, C.c_void_p)

EXPORTS = ("ipxgpu_default_options ipxgpu_last_error ipxgpu_device_count ipxgpu_warmup ipxgpu_create "
           "ipxgpu_destroy ipxgpu_get_layout ipxgpu_get_tiling ipxgpu_synchronize ipxgpu_partition_columns ipxgpu_comm_unique_id "
           "ipxgpu_comm_init ipxgpu_normal_prepare ipxgpu_normal_prepare_dev ipxgpu_normal_apply "
           "ipxgpu_normal_apply_dev ipxgpu_diag_factorize ipxgpu_diag_get ipxgpu_diag_set "
           "ipxgpu_diag_factorize_masked ipxgpu_smw_load ipxgpu_smw_clear ipxgpu_create_group "
           "ipxgpu_diag_apply ipxgpu_pcr_solve ipxgpu_pcr_solve_dev ipxgpu_cr_solve ipxgpu_kktdiag_factorize "
           "ipxgpu_kktdiag_solve ipxgpu_lu_load ipxgpu_tri_solve ipxgpu_split_prepare "
           "ipxgpu_split_apply ipxgpu_kktbasis_prepare ipxgpu_basis_solve ipxgpu_kktbasis_solve ipxgpu_time_normal_apply ipxgpu_launch_count ipxgpu_band_selftest ipxgpu_peer_export ipxgpu_peer_import "
           "ipxgpu_maxvol_weights ipxgpu_maxvol_skip ipxgpu_maxvol_update ipxgpu_maxvol_get "
           "ipxgpu_maxvol_release ipxgpu_set_option ipxgpu_time_tri_solve ipxgpu_tri_trace "
           "ipxgpu_multiply_add").split()

_lib = None


class IpxGpuError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__(f"ipxgpu error {code}: {msg}")
        self.code = code


def load():
    """Loads libipxgpu.so; raises if it has not been built."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise FileNotFoundError(
                f"{LIB_PATH} missing: build it with `python -m ipx_b200.build` "
                "(there is no CPU fallback)")
        lib = C.CDLL(LIB_PATH, mode=os.RTLD_LOCAL | os.RTLD_NOW)
        lib.ipxgpu_last_error.restype = C.c_char_p
        _lib = lib
    return _lib


def _check(rc):
    if rc != 0:
        raise IpxGpuError(rc, load().ipxgpu_last_error().decode())


def band_selftest(m, n, AIp, AIi, AIx, x, force=False):
    """Host-only check of the banded sweep layout (ipxgpu_band_selftest); no device needed."""
    import numpy as np
    AIp = np.ascontiguousarray(AIp, dtype=np.int64)
    AIi = np.ascontiguousarray(AIi, dtype=np.int64)
    AIx = np.ascontiguousarray(AIx, dtype=np.float64)
    x = np.ascontiguousarray(x, dtype=np.float64)
    out = (C.c_double * 6)()
    _check(load().ipxgpu_band_selftest(
        i64(m), i64(n), AIp.ctypes.data_as(C.POINTER(i64)), AIi.ctypes.data_as(C.POINTER(i64)),
        AIx.ctypes.data_as(C.POINTER(C.c_double)), x.ctypes.data_as(C.POINTER(C.c_double)),
        C.c_int32(1 if force else 0), out))
    keys = "planned err pad".split()
    return {"sweep1": dict(zip(keys, list(out)[:3])), "sweep2": dict(zip(keys, list(out)[3:]))}


def _d(a):
    return None if a is None else a.ctypes.data_as(_dp)


def _i(a):
    return None if a is None else a.ctypes.data_as(_ip)


def _f64(a):
    return None if a is None else np.ascontiguousarray(a, dtype=np.float64)


def _i64(a):
    return None if a is None else np.ascontiguousarray(a, dtype=np.int64)


def device_count():
    n = C.c_int(0)
    _check(load().ipxgpu_device_count(C.byref(n)))
    return n.value


def partition_columns(n, AIp, nranks):
    """Column shard bounds balanced by nonzeros (host arithmetic only)."""
    AIp = _i64(AIp)
    bounds = np.zeros(nranks + 1, np.int64)
    _check(load().ipxgpu_partition_columns(i64(n), _i(AIp), C.c_int32(nranks), _i(bounds)))
    return bounds


def comm_unique_id():
    buf = C.create_string_buffer(128)
    _check(load().ipxgpu_comm_unique_id(buf))
    return buf.raw


class Context:
    """Device-resident AI = [A I] (or one column shard of it)."""

    def __init__(self, m, n, AIp, AIi, AIx, device=-1, rank=0, nranks=1, col_begin=-1,
                 col_end=-1, panel_cols=0, stream=None):
        self.lib = load()
        self.m, self.n = int(m), int(n)
        AIp, AIi, AIx = _i64(AIp), _i64(AIi), _f64(AIx)
        opt = Options()
        self.lib.ipxgpu_default_options(C.byref(opt))
        opt.device, opt.rank, opt.nranks = device, rank, nranks
        opt.col_begin, opt.col_end, opt.panel_cols = col_begin, col_end, panel_cols
        opt.stream = stream
        self.h = C.c_void_p()
        _check(self.lib.ipxgpu_create(C.byref(self.h), i64(m), i64(n), _i(AIp), _i(AIi), _d(AIx),
                                      C.byref(opt)))

    def close(self):
        if getattr(self, "h", None):
            self.lib.ipxgpu_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def layout(self):
        out = (i64 * 8)()
        _check(self.lib.ipxgpu_get_layout(self.h, out))
        keys = "m n nnz_local col_begin col_end num_panels csc_tiles csr_tiles".split()
        return dict(zip(keys, list(out)))

    def tiling(self):
        out = (i64 * 16)()
        _check(self.lib.ipxgpu_get_tiling(self.h, out))
        keys = "enabled VB SB NVB NSB K nparts nitems".split()
        return {"sweep1": dict(zip(keys, list(out)[:8])), "sweep2": dict(zip(keys, list(out)[8:]))}

    def synchronize(self):
        _check(self.lib.ipxgpu_synchronize(self.h))

    def launch_count(self):
        n = i64(0)
        _check(self.lib.ipxgpu_launch_count(self.h, C.byref(n)))
        return n.value

    def comm_init(self, uid):
        _check(self.lib.ipxgpu_comm_init(self.h, C.c_char_p(uid)))

    def peer_export(self):
        """IPC handle (64 bytes) of this rank's exchange buffer."""
        buf = C.create_string_buffer(64)
        _check(self.lib.ipxgpu_peer_export(self.h, buf))
        return buf.raw

    def peer_import(self, handles):
        """handles: the nranks 64-byte handles concatenated in rank order."""
        _check(self.lib.ipxgpu_peer_import(self.h, C.c_char_p(bytes(handles))))

    # ---- NormalMatrix ----
    def normal_prepare(self, W):
        W = _f64(W)
        _check(self.lib.ipxgpu_normal_prepare(self.h, _d(W)))

    def normal_prepare_dev(self, W_ptr):
        _check(self.lib.ipxgpu_normal_prepare_dev(self.h, C.c_void_p(W_ptr)))

    def normal_apply(self, rhs, want_dot=True):
        rhs, lhs, dot = _f64(rhs), np.empty(self.m), C.c_double(np.nan)
        _check(self.lib.ipxgpu_normal_apply(self.h, _d(rhs), _d(lhs),
                                            C.byref(dot) if want_dot else None))
        return lhs, dot.value

    def normal_apply_dev(self, rhs_ptr, lhs_ptr):
        _check(self.lib.ipxgpu_normal_apply_dev(self.h, C.c_void_p(rhs_ptr), C.c_void_p(lhs_ptr)))

    # ---- DiagonalPrecond ----
    def diag_factorize(self, W, use_prepared=False):
        W = _f64(W)
        _check(self.lib.ipxgpu_diag_factorize(self.h, _d(W), C.c_int(1 if use_prepared else 0)))

    def diag_get(self):
        d = np.empty(self.m)
        _check(self.lib.ipxgpu_diag_get(self.h, _d(d)))
        return d

    def diag_set(self, diag):
        diag = _f64(diag)
        _check(self.lib.ipxgpu_diag_set(self.h, _d(diag)))

    def diag_apply(self, rhs):
        rhs, lhs, dot = _f64(rhs), np.empty(self.m), C.c_double(np.nan)
        _check(self.lib.ipxgpu_diag_apply(self.h, _d(rhs), _d(lhs), C.byref(dot)))
        return lhs, dot.value

    def diag_factorize_masked(self, W, dense_cols, use_prepared=False):
        W, cols = _f64(W), _i64(dense_cols)
        _check(self.lib.ipxgpu_diag_factorize_masked(self.h, _d(W), C.c_int(1 if use_prepared else 0),
                                                     C.c_int64(len(cols)), _i(cols)))

    def smw_load(self, Adp, Adi, Adx, L):
        """Ad (CSC of the dense columns) and the lower Cholesky factor L (nd x nd, column-major)
        of the Schur complement; switches the preconditioner to its SMW form."""
        Adp, Adi, Adx = _i64(Adp), _i64(Adi), _f64(Adx)
        L = np.asfortranarray(L, dtype=np.float64)
        nd = len(Adp) - 1
        assert L.shape == (nd, nd)
        _check(self.lib.ipxgpu_smw_load(self.h, C.c_int64(nd), _i(Adp), _i(Adi), _d(Adx),
                                        L.ctypes.data_as(C.POINTER(C.c_double))))

    def smw_clear(self):
        _check(self.lib.ipxgpu_smw_clear(self.h))

    # ---- ConjugateResiduals ----
    def _cr(self, fn, pre_args, rhs, tol, resscale, maxiter, lhs0, hist_cap, interrupt):
        rhs, resscale = _f64(rhs), _f64(resscale)
        lhs = np.zeros(self.m) if lhs0 is None else _f64(lhs0).copy()
        res = CrResult()
        hist = np.full(max(hist_cap, 1), np.nan)
        cb = INTERRUPT_FN(interrupt) if interrupt else C.cast(None, INTERRUPT_FN)
        _check(fn(self.h, *pre_args, _d(rhs), C.c_double(tol), _d(resscale), i64(maxiter), _d(lhs),
                  C.byref(res), cb, None, _d(hist) if hist_cap else None, i64(hist_cap)))
        out = res.asdict()
        out["hist"] = hist[:min(hist_cap, out["iter"] + 1)]
        return lhs, out

    def pcr_solve(self, rhs, tol, resscale, maxiter, lhs0=None, hist_cap=0, interrupt=None):
        return self._cr(self.lib.ipxgpu_pcr_solve, (), rhs, tol, resscale, maxiter, lhs0, hist_cap,
                        interrupt)

    def pcr_solve_dev(self, rhs_ptr, tol, resscale_ptr, maxiter, lhs_ptr, zero_start=True):
        res = CrResult()
        _check(self.lib.ipxgpu_pcr_solve_dev(self.h, C.c_void_p(rhs_ptr), C.c_double(tol),
                                             C.c_void_p(resscale_ptr), i64(maxiter),
                                             C.c_void_p(lhs_ptr), C.c_int(1 if zero_start else 0),
                                             C.byref(res)))
        return res.asdict()

    def cr_solve(self, op, rhs, tol, resscale, maxiter, lhs0=None, hist_cap=0, interrupt=None):
        return self._cr(self.lib.ipxgpu_cr_solve, (C.c_int(op),), rhs, tol, resscale, maxiter,
                        lhs0, hist_cap, interrupt)

    # ---- KKTSolverDiag ----
    def kktdiag_factorize(self, xl=None, xu=None, zl=None, zu=None, mu=0.0, want_W=False):
        a = [_f64(v) for v in (xl, xu, zl, zu)]
        W = np.empty(self.n + self.m) if want_W else None
        rs = np.empty(self.m) if want_W else None
        _check(self.lib.ipxgpu_kktdiag_factorize(self.h, *[_d(v) for v in a], C.c_double(mu),
                                                 _d(W), _d(rs)))
        return W, rs

    def kktdiag_solve(self, a, b, tol, maxiter):
        a, b = _f64(a), _f64(b)
        x, y, res = np.empty(self.n + self.m), np.empty(self.m), CrResult()
        _check(self.lib.ipxgpu_kktdiag_solve(self.h, _d(a), _d(b), C.c_double(tol), i64(maxiter),
                                             _d(x), _d(y), C.byref(res),
                                             C.cast(None, INTERRUPT_FN), None))
        return x, y, res.asdict()

    # ---- triangular solves / split operator ----
    def lu_load(self, L, U):
        Lp, Li, Lx = _i64(L[0]), _i64(L[1]), _f64(L[2])
        Up, Ui, Ux = _i64(U[0]), _i64(U[1]), _f64(U[2])
        lev = (i64 * 4)()
        _check(self.lib.ipxgpu_lu_load(self.h, i64(self.m), _i(Lp), _i(Li), _d(Lx), _i(Up), _i(Ui),
                                       _d(Ux), lev))
        return list(lev)

    def tri_solve(self, which, x):
        x = _f64(x).copy()
        _check(self.lib.ipxgpu_tri_solve(self.h, C.c_int(which), _d(x)))
        return x

    def split_prepare(self, nonbasic_scale, rowperm_inv, free_positions):
        s, r, f = _f64(nonbasic_scale), _i64(rowperm_inv), _i64(free_positions)
        _check(self.lib.ipxgpu_split_prepare(self.h, _d(s), _i(r), i64(len(f)), _i(f)))

    def split_apply(self, rhs):
        rhs, lhs, dot = _f64(rhs), np.empty(self.m), C.c_double(np.nan)
        _check(self.lib.ipxgpu_split_apply(self.h, _d(rhs), _d(lhs), C.byref(dot)))
        return lhs, dot.value

    # ---- KKTSolverBasis ----
    def kktbasis_prepare(self, basic_var, colperm, basic_scale):
        v, p, s = _i64(basic_var), _i64(colperm), _f64(basic_scale)
        _check(self.lib.ipxgpu_kktbasis_prepare(self.h, _i(v), _i(p), _d(s)))

    def basis_solve(self, rhs, trans):
        rhs, lhs = _f64(rhs), np.empty(self.m)
        _check(self.lib.ipxgpu_basis_solve(self.h, C.c_char(trans.encode()), _d(rhs), _d(lhs)))
        return lhs

    def kktbasis_solve(self, a, b, tol, maxiter):
        a, b = _f64(a), _f64(b)
        x, y, res = np.empty(self.n + self.m), np.empty(self.m), CrResult()
        _check(self.lib.ipxgpu_kktbasis_solve(self.h, _d(a), _d(b), C.c_double(tol), i64(maxiter),
                                              _d(x), _d(y), C.byref(res),
                                              C.cast(None, INTERRUPT_FN), None))
        return x, y, res.asdict()

    def set_option(self, name, value):
        _check(self.lib.ipxgpu_set_option(self.h, name.encode(), i64(value)))

    # ---- products with AI outside the KKT solve ----
    def multiply_add(self, rhs, alpha, lhs, trans):
        """lhs + alpha*AI*rhs ('N') or lhs + alpha*AI'*rhs ('T'), summed in the reference's order."""
        rhs, lhs = _f64(rhs), _f64(lhs).copy()
        _check(self.lib.ipxgpu_multiply_add(self.h, _d(rhs), C.c_double(alpha), _d(lhs),
                                            C.c_char(trans.encode())))
        return lhs

    # ---- Maxvolume column sweeps ----
    def maxvol_weights(self, colscale, work):
        cs, w, top = _f64(colscale), _f64(work), MaxvolTop()
        _check(self.lib.ipxgpu_maxvol_weights(self.h, _d(cs), _d(w), C.byref(top)))
        return top.asdict()

    def maxvol_skip(self, j, search=True):
        top = MaxvolTop()
        _check(self.lib.ipxgpu_maxvol_skip(self.h, i64(j), C.byref(top) if search else None))
        return top.asdict() if search else None

    def maxvol_update(self, btran, alpha, jb, colscale_jb, colweight_jb, jn):
        b, top = _f64(btran), MaxvolTop()
        _check(self.lib.ipxgpu_maxvol_update(self.h, _d(b), C.c_double(alpha), i64(jb),
                                             C.c_double(colscale_jb), C.c_double(colweight_jb),
                                             i64(jn), C.byref(top)))
        return top.asdict()

    def maxvol_get(self):
        cs, cw = np.empty(self.n + self.m), np.empty(self.n + self.m)
        _check(self.lib.ipxgpu_maxvol_get(self.h, _d(cs), _d(cw)))
        return cs, cw

    def maxvol_release(self):
        _check(self.lib.ipxgpu_maxvol_release(self.h))

    # ---- measurement ----
    def time_tri_solve(self, which, x, reps=10):
        x, out = _f64(x), C.c_double(0.0)
        _check(self.lib.ipxgpu_time_tri_solve(self.h, C.c_int(which), C.c_int(reps), _d(x),
                                              C.byref(out)))
        return out.value

    def tri_trace(self):
        out = np.zeros(2 * self.m, np.uint64)
        _check(self.lib.ipxgpu_tri_trace(self.h, out.ctypes.data_as(C.POINTER(C.c_uint64))))
        return out.reshape(self.m, 2)

    def time_normal_apply(self, reps, flush_l2=True):
        out = (C.c_double * 3)()
        _check(self.lib.ipxgpu_time_normal_apply(self.h, C.c_int(reps), C.c_int(1 if flush_l2 else 0),
                                                 out))
        return {"apply_ms": out[0], "sweep1_ms": out[1], "sweep2_ms": out[2]}
