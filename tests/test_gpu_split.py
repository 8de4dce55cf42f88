"""GPU parity of the basis path: level-scheduled triangular solves and the
basis-preconditioned operator against the CPU oracle."""

import numpy as np
import pytest
import scipy.sparse as sp

from conftest import rel_err
from ipx_b200 import lpgen

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def capi():
    from ipx_b200 import capi as c
    c.load()
    assert c.device_count() >= 1
    return c


def random_factors(dim, seed, band=None, density=3.0):
    """Random unit-lower L (strict part) and upper U (diagonal last) in CSC."""
    rng = np.random.default_rng(seed)

    def tri(lower):
        cols_i, cols_x = [], []
        for j in range(dim):
            k = int(rng.poisson(density))
            if lower:
                lo, hi = j + 1, dim if band is None else min(dim, j + 1 + band)
            else:
                lo, hi = 0 if band is None else max(0, j - band), j
            k = min(k, max(hi - lo, 0))
            rows = np.sort(rng.choice(np.arange(lo, hi), size=k, replace=False)) if k else \
                np.array([], dtype=np.int64)
            vals = rng.uniform(-0.9, 0.9, k)
            if not lower:
                rows = np.append(rows, j)
                vals = np.append(vals, rng.uniform(1.0, 3.0) * rng.choice([-1.0, 1.0]))
            cols_i.append(rows.astype(np.int64))
            cols_x.append(vals)
        p = np.zeros(dim + 1, np.int64)
        p[1:] = np.cumsum([len(c) for c in cols_i])
        return p, np.concatenate(cols_i).astype(np.int64), np.concatenate(cols_x)

    return tri(True), tri(False)


@pytest.mark.parametrize("dim,band,seed", [(1, None, 1), (50, None, 2), (3000, None, 3),
                                           (20000, 40, 4), (6000, 2, 5)])
def test_triangular_solves_bit_exact(capi, oracle, dim, band, seed):
    L, U = random_factors(dim, seed, band)
    lp = lpgen.random_sparse_lp(dim, max(2 * dim, 4), min(3, dim), 50 + seed)
    ctx = capi.Context(lp.m, lp.n, *lp.solver_form())
    levels = ctx.lu_load(L, U)
    assert all(1 <= l <= dim for l in levels)
    Lo, Uo = oracle.Csc(*L), oracle.Csc(*U)
    x = np.random.default_rng(seed).standard_normal(dim)
    x[::7] = 0.0  # exercise the reference's skip-on-zero branches
    expect = {
        0: oracle.triangular_solve(dim, Lo, x, "n", "l", 1)[0],
        1: oracle.triangular_solve(dim, Uo, x, "n", "u", 0)[0],
        2: oracle.triangular_solve(dim, Uo, x, "t", "u", 0)[0],
        3: oracle.triangular_solve(dim, Lo, x, "t", "l", 1)[0],
    }
    for which in range(4):
        got = ctx.tri_solve(which, x)
        assert np.array_equal(got, expect[which]), f"solve {which}"
    fwd = oracle.triangular_solve(dim, Uo, expect[0], "n", "u", 0)[0]
    bwd = oracle.triangular_solve(dim, Lo, expect[2], "t", "l", 1)[0]
    assert np.array_equal(ctx.tri_solve(4, x), fwd)
    assert np.array_equal(ctx.tri_solve(5, x), bwd)
    ctx.close()


def _split_setup(lp, seed):
    """Synthetic basis data with the reference's conventions
    (src/splitted_normal_matrix.cc:18-66)."""
    rng = np.random.default_rng(seed)
    m, n = lp.m, lp.n
    L, U = random_factors(m, seed, band=30)
    rowperm = rng.permutation(m).astype(np.int64)
    rowperm_inv = np.empty(m, np.int64)
    rowperm_inv[rowperm] = np.arange(m)
    # m basic columns among n+m, the rest nonbasic except a few fixed
    basic = np.zeros(n + m, bool)
    basic[rng.choice(n + m, size=m, replace=False)] = True
    fixed = (~basic) & (rng.random(n + m) < 0.03)
    colscale = np.exp(rng.uniform(-5, -3, n + m))  # keeps C = I + small well conditioned
    nonbasic_scale = np.where(basic | fixed, 0.0, colscale)
    free_positions = np.sort(rng.choice(m, size=max(1, m // 50), replace=False)).astype(np.int64)
    # N = AI[:, nonbasic] with permuted rows and scaled columns
    AIp, AIi, AIx = lp.solver_form()
    AI = sp.csc_matrix((AIx, AIi, AIp), shape=(m, n + m))
    nb = np.nonzero(~(basic | fixed))[0]
    N = AI[:, nb].tocsc()
    N = sp.csc_matrix((N.data * np.repeat(colscale[nb], np.diff(N.indptr)),
                       rowperm_inv[N.indices], N.indptr), shape=N.shape)
    return L, U, rowperm_inv, nonbasic_scale, free_positions, N


@pytest.mark.parametrize("shape", [(400, 3000, 4), (6000, 50000, 8)])
def test_split_apply_and_cr(capi, oracle, shape):
    m, n, k = shape
    lp = lpgen.random_sparse_lp(m, n, k, 61)
    L, U, rinv, nbscale, freepos, N = _split_setup(lp, 62)
    ctx = capi.Context(lp.m, lp.n, *lp.solver_form())
    ctx.lu_load(L, U)
    ctx.split_prepare(nbscale, rinv, freepos)
    S = oracle.SplitOperator(m, oracle.Csc(*L), oracle.Csc(*U),
                             oracle.Csc(N.indptr, N.indices, N.data), N.shape[1], freepos)
    x = np.random.default_rng(63).standard_normal(m)
    y, dot = ctx.split_apply(x)
    y0, dot0 = S.apply(x)
    assert rel_err(y, y0) <= 1e-12
    assert abs(dot - dot0) <= 1e-12 * np.abs(x * y0).sum()
    assert np.all(y[freepos] == 0.0)
    # unpreconditioned CR on the split operator (reference kkt_solver_basis.cc:150)
    rhs = np.random.default_rng(64).standard_normal(m)
    rhs[freepos] = 0.0
    z, info = ctx.cr_solve(1, rhs, 1e-6, None, 60, hist_cap=128)
    z0, info0 = oracle.cr_solve(S.operator(), m, rhs, 1e-6, None, 60, hist_cap=128)
    assert info["errflag"] == info0["errflag"]
    kk = min(len(info["hist"]), len(info0["hist"]), 5)
    assert np.allclose(info["hist"][:kk], info0["hist"][:kk], rtol=1e-9, atol=0.0)
    if info0["errflag"] == 0:
        assert abs(info["iter"] - info0["iter"]) <= 1
        assert rel_err(z, z0) <= 1e-5
    else:
        assert info["iter"] == info0["iter"]
    assert info["time_B"] > 0 and info["time_Bt"] > 0 and info["time_NNt"] > 0
    ctx.close()
