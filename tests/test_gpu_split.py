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
    # L, U, U': the rows are summed in the reference's order whatever the options say. L': in the
    # reference's order with the option below (bit-identical), bottom up by default - the same
    # sums in the opposite order, so the bar is what one ulp in the right-hand side does to the
    # reference's own solve.
    x_ulp = x + np.spacing(np.abs(x)) * np.random.default_rng(seed + 1).choice([-1.0, 1.0], dim)
    lt_ulp = oracle.triangular_solve(dim, Lo, x_ulp, "t", "l", 1)[0]
    lt_bar = 100 * np.abs(lt_ulp - expect[3]).max() + 1e-14 * np.abs(expect[3]).max()
    for which in range(3):
        got = ctx.tri_solve(which, x)
        assert np.array_equal(got, expect[which]), f"solve {which}"
    got = ctx.tri_solve(3, x)
    assert np.abs(got - expect[3]).max() <= lt_bar
    assert np.array_equal(ctx.tri_solve(3, x), got)  # deterministic
    fwd = oracle.triangular_solve(dim, Uo, expect[0], "n", "u", 0)[0]
    bwd = oracle.triangular_solve(dim, Lo, expect[2], "t", "l", 1)[0]
    assert np.array_equal(ctx.tri_solve(4, x), fwd)
    assert np.abs(ctx.tri_solve(5, x) - bwd).max() <= 100 * np.abs(
        oracle.triangular_solve(dim, Lo, oracle.triangular_solve(dim, Uo, x_ulp, "t", "u", 0)[0],
                                "t", "l", 1)[0] - bwd).max() + 1e-14 * np.abs(bwd).max()
    ctx.set_option("tri_reference_order", 1)
    assert np.array_equal(ctx.tri_solve(3, x), expect[3]), "solve 3, reference order"
    assert np.array_equal(ctx.tri_solve(5, x), bwd)
    with pytest.raises(capi.IpxGpuError):
        ctx.set_option("no_such_option", 1)
    ctx.close()


def test_triangular_solves_long_rows(capi, oracle):
    """Rows with more than 2048 entries (linking rows of a block-angular basis): summed as 32
    interleaved partial sums by default - a rounding-level difference, held to what one ulp in the
    right-hand side does to the oracle's own solve - and entry by entry, bit for bit, with the
    option tri_reference_order."""
    dim = 6000
    rng = np.random.default_rng(77)
    Lm = sp.random(dim, dim, density=2e-3, random_state=rng.integers(1 << 30), format="lil",
                   data_rvs=lambda k: rng.uniform(-0.9, 0.9, k))
    Lm[:, 0] = rng.uniform(-0.01, 0.01, (dim, 1))      # long row 0 of the L' solve
    Lm[dim - 1, :] = rng.uniform(-0.01, 0.01, (1, dim))  # long last row of the L solve
    Lm = sp.tril(Lm.tocsc(), -1).tocsc()
    Um = sp.random(dim, dim, density=2e-3, random_state=rng.integers(1 << 30), format="lil",
                   data_rvs=lambda k: rng.uniform(-0.9, 0.9, k))
    Um[0, :] = rng.uniform(-0.01, 0.01, (1, dim))        # long row 0 of the U solve
    Um[:, dim - 1] = rng.uniform(-0.01, 0.01, (dim, 1))  # long last row of the U' solve
    Um = (sp.triu(Um.tocsc(), 1) + sp.diags(rng.uniform(1.0, 3.0, dim))).tocsc()
    for M in (Lm, Um):
        M.sort_indices()
    L = (Lm.indptr.astype(np.int64), Lm.indices.astype(np.int64), Lm.data.copy())
    U = (Um.indptr.astype(np.int64), Um.indices.astype(np.int64), Um.data.copy())
    lp = lpgen.random_sparse_lp(dim, 2 * dim, 3, 78)
    ctx = capi.Context(lp.m, lp.n, *lp.solver_form())
    ctx.lu_load(L, U)
    Lo, Uo = oracle.Csc(*L), oracle.Csc(*U)
    x = rng.standard_normal(dim)
    x_ulp = x + np.spacing(np.abs(x)) * rng.choice([-1.0, 1.0], dim)
    systems = [(Lo, "n", "l", 1), (Uo, "n", "u", 0), (Uo, "t", "u", 0), (Lo, "t", "l", 1)]
    want = [oracle.triangular_solve(dim, A, x, t, ul, unit)[0] for A, t, ul, unit in systems]
    moved = [oracle.triangular_solve(dim, A, x_ulp, t, ul, unit)[0] for A, t, ul, unit in systems]
    for which in range(4):
        got = ctx.tri_solve(which, x)
        bar = 100 * np.abs(moved[which] - want[which]).max() + 1e-13 * np.abs(want[which]).max()
        assert np.abs(got - want[which]).max() <= bar, which
        assert np.array_equal(ctx.tri_solve(which, x), got)  # deterministic
    ctx.set_option("tri_reference_order", 1)
    for which in range(4):
        assert np.array_equal(ctx.tri_solve(which, x), want[which]), which
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


def _basis_setup(reflib, capi, lp, seed, num_free):
    """A real basis of `lp` (crash basis of the reference's Basis object with the host LU
    provider) loaded into a device context the way SplittedNormalMatrix::Prepare and
    KKTSolverBasis::_Factorize do (src/splitted_normal_matrix.cc:18-66,
    src/kkt_solver_basis.cc:59-65)."""
    rng = np.random.default_rng(seed)
    mdl = reflib.model(lp, dualize=0)
    m, n = mdl.m, mdl.n
    colscale = np.exp(rng.uniform(-2, 2, n + m))
    mdl.basis_from_weights(colscale)
    basis, status = mdl.basis_get()          # status: 0 basic, -1 nonbasic (src/basis.h:57-66)
    (Lp, Li, Lx), (Up, Ui, Ux), rowperm, colperm = mdl.basis_lu()
    rowperm_inv = np.empty(m, np.int64)
    rowperm_inv[rowperm] = np.arange(m)
    basic_var = basis[colperm]
    free_positions = np.sort(rng.choice(m, size=num_free, replace=False)).astype(np.int64)
    is_free = np.zeros(m, bool)
    is_free[free_positions] = True
    basic_scale = np.where(is_free, 1.0, colscale[basic_var])
    Ux_scaled = Ux * np.repeat(basic_scale, np.diff(Up))
    nonbasic = status < 0
    nonbasic_scale = np.where(nonbasic, colscale, 0.0)
    AIp, AIi, AIx = mdl.AI()
    ctx = capi.Context(m, n, AIp, AIi, AIx)
    ctx.lu_load((Lp, Li, Lx), (Up, Ui, Ux_scaled))
    ctx.split_prepare(nonbasic_scale, rowperm_inv, free_positions)
    ctx.kktbasis_prepare(basic_var, colperm, basic_scale)
    AI = sp.csc_matrix((AIx, AIi, AIp), shape=(m, n + m))
    return mdl, ctx, AI, basis, basic_var, colperm, colscale, nonbasic, is_free


@pytest.mark.parametrize("shape", [(60, 300, 4), (3000, 20000, 6)])
def test_basis_solve_dense_on_device(capi, reflib, shape):
    """ipxgpu_basis_solve = Basis::SolveDense (src/basis.cc:168-170) on the loaded factors."""
    m, n, k = shape
    lp = lpgen.random_sparse_lp(m, n, k, 71)
    mdl, ctx, AI, basis, *_ = _basis_setup(reflib, capi, lp, 72, num_free=max(1, m // 40))
    B = AI[:, basis].tocsc()
    rhs = np.random.default_rng(73).standard_normal(m)
    for trans in ("N", "T"):
        want = mdl.basis_solve_dense(rhs, trans)
        got = ctx.basis_solve(rhs, trans)
        assert rel_err(got, want) <= 1e-10, trans
        resid = (B @ got if trans == "N" else B.T @ got) - rhs
        assert np.abs(resid).max() <= 1e-9 * (1.0 + np.abs(got).max() * np.abs(B).max())
    ctx.close()
    mdl.close()


def test_kktbasis_solve_against_dense_kkt(capi, reflib):
    """The whole KKTSolverBasis::_Solve on the device against a dense solve of the KKT system
    it reduces (src/kkt_solver.h:20-39): [G AI'; AI 0] (x, y) = (a, b) with G = diag(1/s^2),
    G = 0 for free basic variables and x = 0 for fixed nonbasic ones."""
    m, n = 50, 160
    lp = lpgen.random_sparse_lp(m, n, 4, 81)
    mdl, ctx, AI, basis, basic_var, colperm, colscale, nonbasic, is_free = \
        _basis_setup(reflib, capi, lp, 82, num_free=3)
    rng = np.random.default_rng(83)
    # a few nonbasic variables are "fixed" (scale 0, src/kkt_solver_basis.cc:381-383): they are
    # exactly the nonbasic columns whose scale was masked; re-prepare with them masked
    nb = np.nonzero(nonbasic)[0]
    fixed = rng.choice(nb, size=5, replace=False)
    g = 1.0 / colscale**2
    g[basic_var[is_free]] = 0.0
    keep = np.ones(n + m, bool)
    keep[fixed] = False
    nonbasic_scale = np.where(nonbasic & keep, colscale, 0.0)
    rowperm_inv = np.empty(m, np.int64)
    rowperm_inv[mdl.basis_lu()[2]] = np.arange(m)
    ctx.split_prepare(nonbasic_scale, rowperm_inv, np.nonzero(is_free)[0].astype(np.int64))
    ctx.kktbasis_prepare(basic_var, colperm, np.where(is_free, 1.0, colscale[basic_var]))
    a, b = rng.standard_normal(n + m), rng.standard_normal(m)
    x, y, info = ctx.kktbasis_solve(a, b, 1e-13, 2000)
    assert info["errflag"] == 0, info
    A = AI.toarray()[:, keep]
    K = np.block([[np.diag(g[keep]), A.T], [A, np.zeros((m, m))]])
    sol = np.linalg.solve(K, np.concatenate([a[keep], b]))
    x0 = np.zeros(n + m)
    x0[keep] = sol[:keep.sum()]
    y0 = sol[keep.sum():]
    assert np.all(x[fixed] == 0.0)
    assert rel_err(x, x0) <= 1e-8 and rel_err(y, y0) <= 1e-8
    # the second block row holds to rounding (x_B = inverse(B)(b - N x_N), :179-193)
    assert np.abs(AI @ x - b).max() <= 1e-10 * (1.0 + np.abs(x).max() * np.abs(AI).max())
    assert info["time_B"] > 0 and info["time_NNt"] > 0
    ctx.close()
    mdl.close()
