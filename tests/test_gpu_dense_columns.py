"""Dense columns: the Sherman-Morrison-Woodbury form of the diagonal preconditioner on the
device (reference src/diagonal_precond.cc:48-102 Factorize, :133-149 _Apply; columns marked by
Model::FindDenseColumns, src/model.cc:34-56), against the reference's CPU classes and against
a dense numpy restatement, through the C ABI and through the drop-in build of IPX."""

import numpy as np
import pytest
import scipy.linalg
import scipy.sparse as sp

from conftest import rel_err
from ipx_b200 import capi, lpgen

pytestmark = pytest.mark.gpu


def _dense_lp(ndense=3, m=400, n=3000, seed=21):
    return lpgen.dense_column_lp(m, n, 5, ndense, seed)


def _smw_host(lp, W):
    """E, Ad, the Cholesky factor of S and the dense preconditioner matrix inv(P)."""
    m, n = lp.m, lp.n
    A = sp.csc_matrix((lp.Ax, lp.Ai, lp.Ap), shape=(m, n))
    dense = np.array(lp.extra["dense_cols"])
    mask = np.ones(n, bool)
    mask[dense] = False
    E = W[n:] + np.asarray(A[:, mask].multiply(A[:, mask]) @ W[:n][mask]).ravel()
    Ad = A[:, dense].tocsc()
    Ad.sort_indices()
    S = np.diag(1.0 / W[dense]) + (Ad.T @ sp.diags(1.0 / E) @ Ad).toarray()
    L = np.linalg.cholesky(S)
    return E, Ad, L, dense


@pytest.mark.parametrize("regime", ["mid", "wide"])
@pytest.mark.parametrize("ndense", [1, 3, 40])
def test_smw_apply_matches_dense_algebra(regime, ndense):
    """C ABI: masked diagonal + SMW apply against numpy in the reference's operation order."""
    lp = _dense_lp(ndense)
    m, n = lp.m, lp.n
    W = lpgen.weights(n + m, regime, 5)
    # a dense column whose weight dominates its rows: E must come out without cancellation
    W[lp.extra["dense_cols"][0]] *= 1e12
    E, Ad, L, dense = _smw_host(lp, W)
    ctx = capi.Context(m, n, *lp.solver_form())
    ctx.diag_factorize_masked(W, dense)
    assert rel_err(ctx.diag_get(), E) <= 1e-13
    assert (ctx.diag_get() > 0).all()
    ctx.smw_load(Ad.indptr, Ad.indices, Ad.data, L)
    x = np.random.default_rng(2).standard_normal(m)
    lhs, dot = ctx.diag_apply(x)
    z = scipy.linalg.cho_solve((L, True), Ad.T @ (x / E))
    want = (x - Ad @ z) / E
    assert rel_err(lhs, want) <= 1e-11
    assert abs(dot - want @ x) <= 1e-11 * np.abs(want * x).sum()
    # back to the pure diagonal
    ctx.smw_clear()
    lhs2, _ = ctx.diag_apply(x)
    assert rel_err(lhs2, x / E) <= 1e-15
    ctx.close()


def test_pcr_with_dense_columns_c_abi():
    """The device CR loop with the SMW preconditioner converges to the solution of the normal
    equations; iteration count against the same loop in numpy."""
    lp = _dense_lp(3)
    m, n = lp.m, lp.n
    W = lpgen.weights(n + m, "mid", 6)
    W[lp.extra["dense_cols"]] *= 1e3
    E, Ad, L, dense = _smw_host(lp, W)
    AIp, AIi, AIx = lp.solver_form()
    AI = sp.csc_matrix((AIx, AIi, AIp), shape=(m, n + m))
    Cmat = (AI @ sp.diags(W) @ AI.T).toarray()
    ctx = capi.Context(m, n, AIp, AIi, AIx)
    ctx.normal_prepare(W)
    ctx.diag_factorize_masked(None, dense, use_prepared=True)
    ctx.smw_load(Ad.indptr, Ad.indices, Ad.data, L)
    rhs = np.random.default_rng(3).standard_normal(m)
    y, info = ctx.pcr_solve(rhs, 1e-10, None, -1)
    assert info["errflag"] == 0
    assert np.abs(Cmat @ y - rhs).max() <= 1e-9
    # with the dense columns inside the preconditioner the solve needs far fewer iterations
    ctx.diag_factorize(W)
    y2, info2 = ctx.pcr_solve(rhs, 1e-10, None, -1)
    assert info["iter"] < info2["iter"]
    ctx.close()


@pytest.fixture(scope="module")
def pair(reflib, gpulib):
    lp = _dense_lp(4, m=500, n=4000, seed=23)
    a, b = reflib.model(lp), gpulib.model(lp)
    yield lp, a, b
    a.close()
    b.close()


@pytest.mark.parametrize("regime", ["mid", "wide"])
def test_diagonal_precond_dense_columns_vs_reference(pair, regime):
    """DiagonalPrecond::Factorize / Apply of the drop-in build against the reference build."""
    lp, ref, gpu = pair
    m, n = ref.m, ref.n
    W = lpgen.weights(n + m, regime, 7)
    W[lp.extra["dense_cols"][1]] *= 1e10
    x = np.random.default_rng(8).standard_normal(m)
    assert ref.diag_factorize(W, 1) == 0 and gpu.diag_factorize(W, 1) == 0
    l0, d0 = ref.diag_apply(x)
    l1, d1 = gpu.diag_apply(x)
    assert rel_err(l1, l0) <= 1e-11
    assert abs(d1 - d0) <= 1e-11 * np.abs(x * l0).sum()
    # precond_dense_cols = 0: the plain diagonal over all columns on both arms
    assert ref.diag_factorize(W, 0) == 0 and gpu.diag_factorize(W, 0) == 0
    l0, _ = ref.diag_apply(x)
    l1, _ = gpu.diag_apply(x)
    assert rel_err(l1, l0) <= 1e-12


def test_pcr_dense_columns_vs_reference(pair):
    lp, ref, gpu = pair
    m, n = ref.m, ref.n
    W = lpgen.weights(n + m, "mid", 9)
    rhs = np.random.default_rng(10).standard_normal(m)
    resscale = 1.0 / np.sqrt(W[n:])
    for mdl in (ref, gpu):
        mdl.normal_prepare(W)
        assert mdl.diag_factorize(W, 1) == 0
    y0, i0 = ref.pcr_solve(rhs, 1e-8, resscale, -1)
    y1, i1 = gpu.pcr_solve(rhs, 1e-8, resscale, -1)
    assert i0["errflag"] == i1["errflag"] == 0
    assert abs(i0["iter"] - i1["iter"]) <= 1
    assert rel_err(y1, y0) <= 1e-6


def test_lp_solver_with_dense_columns(reflib, gpulib):
    """End to end through ipx_c.h with the default precond_dense_cols = 1."""
    lp = lpgen.dense_column_lp(300, 2000, 4, 3, 7)
    infos = []
    for lib in (reflib, gpulib):
        s = lib.lp_solver()
        s.set_parameters(display=0, dualize=0)
        assert s.load_model(lp) == 0
        status = s.solve()
        infos.append((status, s.info()))
        s.close()
    (st0, i0), (st1, i1) = infos
    assert st0 == st1 == 1000
    assert i0["dense_cols"] == i1["dense_cols"] == 3
    assert i0["status_ipm"] == i1["status_ipm"] == 1
    assert abs(i0["iter"] - i1["iter"]) <= 1
    assert abs(i0["objval"] - i1["objval"]) <= 1e-9 * max(1.0, abs(i0["objval"]))
    assert abs(i1["objval"] - lp.optimum) <= 1e-7 * max(1.0, abs(lp.optimum))
