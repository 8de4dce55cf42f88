"""Pins the CPU oracle (oracle/ipx_oracle.c): bit-for-bit against the golden
vectors generated from the compiled reference (tests/golden/*.npz) and, where
oracle/_ref is present, against the reference's classes on fresh seeded inputs."""

import glob
import os

import numpy as np
import pytest

from ipx_b200 import lpgen

GOLDEN = sorted(glob.glob(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden",
                                       "*.npz")))


def test_golden_fixtures_present():
    assert len(GOLDEN) >= 3


@pytest.fixture(scope="module", params=GOLDEN, ids=[os.path.basename(g)[:-4] for g in GOLDEN])
def gold(request):
    return dict(np.load(request.param))


def _A(oracle, g):
    return oracle.Csc(g["AIp"], g["AIi"], g["AIx"])


def test_normal_apply_matches_golden(oracle, gold):
    g = gold
    m, n = int(g["m"]), int(g["n"])
    y, dot = oracle.normal_apply(m, n, _A(oracle, g), g["W"], g["x"])
    assert np.array_equal(y, g["normal_y"]) and dot == float(g["normal_dot"])
    y, dot = oracle.normal_apply(m, n, _A(oracle, g), None, g["x"])
    assert np.array_equal(y, g["normal_y_nullW"]) and dot == float(g["normal_dot_nullW"])


def test_diagonal_matches_golden(oracle, gold):
    g = gold
    m, n = int(g["m"]), int(g["n"])
    diag = oracle.diag_build(m, n, _A(oracle, g), g["W"])
    lhs, dot = oracle.diag_apply(diag, g["x"])
    assert np.array_equal(lhs, g["diag_lhs"]) and dot == float(g["diag_dot"])


def test_conjugate_residuals_match_golden(oracle, gold):
    g = gold
    m, n = int(g["m"]), int(g["n"])
    A = _A(oracle, g)
    diag = oracle.diag_build(m, n, A, g["W"])
    op = oracle.normal_operator(m, n, A, g["W"])
    y, info = oracle.pcr_solve(op, m, diag, g["cr_rhs"], 1e-8, g["cr_resscale"], -1)
    assert info["iter"] == int(g["pcr_iter"]) and info["errflag"] == int(g["pcr_errflag"])
    assert np.array_equal(y, g["pcr_y"])
    y, info = oracle.cr_solve(op, m, g["cr_rhs"], 1e-6, None, 1000)
    assert info["iter"] == int(g["cr_iter"]) and info["errflag"] == int(g["cr_errflag"])
    assert np.array_equal(y, g["cr_y"])


def test_kktdiag_matches_golden(oracle, gold):
    g = gold
    m, n = int(g["m"]), int(g["n"])
    A = _A(oracle, g)
    # mu of the reference's Iterate: complementarity over barrier terms
    xl, xu, zl, zu = g["it_xl"], g["it_xu"], g["it_zl"], g["it_zu"]
    fin_l, fin_u = np.isfinite(xl), np.isfinite(xu)
    mu = (np.sum(xl[fin_l] * zl[fin_l]) + np.sum(xu[fin_u] * zu[fin_u])) / (fin_l.sum() + fin_u.sum())
    W, resscale = oracle.kktdiag_weights(m, n, xl, xu, zl, zu, mu)
    diag = oracle.diag_build(m, n, A, W)
    x, y, info = oracle.kktdiag_solve(m, n, A, W, diag, resscale, g["kkt_a"], g["kkt_b"], 1e-8, -1)
    assert info["errflag"] == int(g["kktdiag_err"])
    assert info["iter"] == int(g["kktdiag_iter"])
    # regval only matters where g == 0 (free variables); everything else is exact
    assert np.allclose(y, g["kktdiag_y"], rtol=1e-12, atol=1e-14)
    assert np.allclose(x, g["kktdiag_x"], rtol=1e-10, atol=1e-12)


def test_triangular_solves_match_golden(oracle, gold):
    g = gold
    m = int(g["m"])
    L = oracle.Csc(g["Lp"], g["Li"], g["Lx"])
    U = oracle.Csc(g["Up"], g["Ui"], g["Ux"])
    x = g["x"]
    assert np.array_equal(oracle.triangular_solve(m, L, x, "n", "l", 1)[0], g["tri_L_n"])
    assert np.array_equal(oracle.triangular_solve(m, U, x, "n", "u", 0)[0], g["tri_U_n"])
    assert np.array_equal(oracle.triangular_solve(m, U, x, "t", "u", 0)[0], g["tri_U_t"])
    assert np.array_equal(oracle.triangular_solve(m, L, x, "t", "l", 1)[0], g["tri_L_t"])


def split_from_golden(oracle, g):
    """Rebuilds what SplittedNormalMatrix::Prepare builds
    (reference src/splitted_normal_matrix.cc:26-64) from the exported factors."""
    import scipy.sparse as sp
    m, n = int(g["m"]), int(g["n"])
    status, basis, colperm = g["basis_status"], g["basis"], g["colperm"]
    colscale = g["colscale"]
    rowperm_inv = np.empty(m, np.int64)
    rowperm_inv[g["rowperm"]] = np.arange(m)
    Up, Ui, Ux = g["Up"], g["Ui"], g["Ux"].copy()
    free_positions = []
    for k in range(m):
        j = basis[colperm[k]]
        if status[j] == 0:      # BASIC
            Ux[Up[k]:Up[k + 1]] *= colscale[j]
        elif status[j] == 1:    # BASIC_FREE
            free_positions.append(k)
    nb = np.nonzero(status == -1)[0]
    AI = sp.csc_matrix((g["AIx"], g["AIi"], g["AIp"]), shape=(m, n + m))
    N = AI[:, nb].tocsc()
    N.sort_indices()
    Nx = N.data * np.repeat(colscale[nb], np.diff(N.indptr))
    Ni = rowperm_inv[N.indices]
    S = oracle.SplitOperator(m, oracle.Csc(g["Lp"], g["Li"], g["Lx"]), oracle.Csc(Up, Ui, Ux),
                             oracle.Csc(N.indptr, Ni, Nx), len(nb),
                             np.array(free_positions, dtype=np.int64))
    nonbasic_scale = np.where(status == -1, colscale, 0.0)
    return S, rowperm_inv, nonbasic_scale, np.array(free_positions, dtype=np.int64), (Up, Ui, Ux)


def test_split_operator_matches_golden(oracle, gold):
    g = gold
    m = int(g["m"])
    S, *_ = split_from_golden(oracle, g)
    y, dot = S.apply(g["x"])
    assert np.array_equal(y, g["split_y"]) and dot == float(g["split_dot"])
    z, info = oracle.cr_solve(S.operator(), m, g["cr_rhs"], 1e-8, None, 400)
    assert info["iter"] == int(g["split_cr_iter"])
    assert info["errflag"] == int(g["split_cr_errflag"])
    assert np.array_equal(z, g["split_cr_y"])


def test_golden_lp_objectives(gold):
    g = gold
    assert int(g["lp_status"]) == 1000 and int(g["lp_status_ipm"]) == 1
    if int(g["m"]) == 9:  # afiro: Netlib optimum
        assert abs(float(g["lp_objval"]) - (-464.753142857143)) < 1e-9


# ---- against the compiled reference on fresh inputs (where it is built) ----

@pytest.mark.parametrize("shape", [(50, 300, 3, 1), (700, 9000, 7, 2)])
def test_oracle_equals_reference_bitwise(oracle, reflib, shape):
    m, n, k, seed = shape
    lp = lpgen.random_sparse_lp(m, n, k, seed)
    mdl = reflib.model(lp)
    A = oracle.Csc(*mdl.AI())
    rng = np.random.default_rng(seed)
    x = rng.standard_normal(m)
    for regime in ("ones", "mid", "wide", None):
        W = None if regime is None else lpgen.weights(n + m, regime, seed + 10)
        mdl.normal_prepare(W)
        y0, d0 = mdl.normal_apply(x)
        y1, d1 = oracle.normal_apply(m, n, A, W, x)
        assert np.array_equal(y0, y1) and d0 == d1
        mdl.diag_factorize(W)
        l0, e0 = mdl.diag_apply(x)
        l1, e1 = oracle.diag_apply(oracle.diag_build(m, n, A, W), x)
        assert np.array_equal(l0, l1) and e0 == e1
    W = lpgen.weights(n + m, "mid", seed + 20)
    mdl.normal_prepare(W)
    mdl.diag_factorize(W)
    diag = oracle.diag_build(m, n, A, W)
    rhs = rng.standard_normal(m)
    resscale = 1.0 / np.sqrt(W[n:])
    for tol in (1e-3, 1e-9):
        y0, i0 = mdl.pcr_solve(rhs, tol, resscale, -1)
        y1, i1 = oracle.pcr_solve(oracle.normal_operator(m, n, A, W), m, diag, rhs, tol, resscale, -1)
        assert i0["iter"] == i1["iter"] and i0["errflag"] == i1["errflag"]
        assert np.array_equal(y0, y1)
    # MultiplyAdd on AI (src/sparse_matrix.cc:194-209), both directions
    for alpha in (-1.0, 0.37):
        xx, ll = rng.standard_normal(n + m), rng.standard_normal(m)
        assert (mdl.multiply_add_AI(xx, alpha, ll, "N").tobytes()
                == oracle.multiply_add(m, n + m, A, xx, alpha, ll, "N").tobytes())
        yy, ll = rng.standard_normal(m), rng.standard_normal(n + m)
        assert (mdl.multiply_add_AI(yy, alpha, ll, "T").tobytes()
                == oracle.multiply_add(m, n + m, A, yy, alpha, ll, "T").tobytes())
    # iteration limit and nonzero start
    y0, i0 = mdl.pcr_solve(rhs, 0.0, None, 4, lhs0=0.5 * rhs)
    y1, i1 = oracle.pcr_solve(oracle.normal_operator(m, n, A, W), m, diag, rhs, 0.0, None, 4,
                              lhs0=0.5 * rhs)
    assert i0["errflag"] == i1["errflag"] == 201 and np.array_equal(y0, y1)
    mdl.close()


def test_dropin_multiply_add_without_a_context_is_the_reference(reflib, gpulib):
    """ipx::MultiplyAdd of the drop-in build (ipx_b200/host/multiply_add_gpu.cc) for a matrix
    that is not resident on a device - here: a model nobody has factorized yet, no GPU needed -
    is the reference's own function under its compile-time name: same bits."""
    lp = lpgen.random_sparse_lp(120, 900, 6, 77)
    ref, gpu = reflib.model(lp), gpulib.model(lp)
    m, n = ref.m, ref.n
    rng = np.random.default_rng(78)
    x, lm = rng.standard_normal(n + m), rng.standard_normal(m)
    y, ln = rng.standard_normal(m), rng.standard_normal(n + m)
    assert gpu.multiply_add_AI(x, -1.0, lm, "N").tobytes() == ref.multiply_add_AI(x, -1.0, lm, "N").tobytes()
    assert gpu.multiply_add_AI(y, 0.5, ln, "T").tobytes() == ref.multiply_add_AI(y, 0.5, ln, "T").tobytes()
    ref.close()
    gpu.close()
