"""GPU parity: the C ABI (libipxgpu.so) against the CPU oracle on seeded inputs.

Tolerances: operator applies 1e-12 norm-wise relative (BASELINE.json
north_star); CR iteration counts equal to the oracle's (+-1 allowed where a
residual sits on the tolerance); triangular solves bit-exact.
"""

import numpy as np
import pytest

from conftest import rel_err
from ipx_b200 import lpgen

pytestmark = pytest.mark.gpu

APPLY_TOL = 1e-12


@pytest.fixture(scope="module")
def capi():
    from ipx_b200 import capi as c
    c.load()
    assert c.device_count() >= 1, "no CUDA device"
    return c


def _case(kind):
    if kind == "afiro":
        return lpgen.afiro_lp()
    if kind == "random_small":
        return lpgen.random_sparse_lp(300, 2000, 5, 21)
    if kind == "random_mid":
        return lpgen.random_sparse_lp(5000, 60000, 10, 22)
    if kind == "transport":  # long rows (> one tile) and 2-entry columns
        return lpgen.transportation_lp(20, 3000, 23)
    if kind == "ragged":
        return _ragged_lp()
    if kind == "tall":  # more rows than the planner gives one block: sweep 1 in several parts
        return lpgen.random_sparse_lp(60000, 30000, 10, 24)
    raise ValueError(kind)


def _ragged_lp():
    """Empty columns, empty rows, one dense column, one very long row."""
    rng = np.random.default_rng(77)
    m, n = 700, 5000
    cols = []
    for j in range(n):
        if j % 17 == 0:
            rows = np.array([], dtype=np.int64)          # empty column
        elif j == 5:
            rows = np.arange(0, m - 50, dtype=np.int64)  # dense column
        else:
            k = int(rng.integers(1, 9))
            rows = np.sort(rng.choice(m - 50, size=k, replace=False)).astype(np.int64)
            if j % 2 == 0:
                rows = np.unique(np.append(rows, 3))     # row 3 is very long
        cols.append(rows)
    Ap = np.zeros(n + 1, np.int64)
    Ap[1:] = np.cumsum([len(c) for c in cols])
    Ai = np.concatenate(cols)
    Ax = rng.uniform(0.5, 4.0, len(Ai)) * rng.choice([-1.0, 1.0], len(Ai))
    return lpgen.LP(m, n, Ap, Ai, Ax, np.zeros(m), b"=" * m, np.zeros(n), np.zeros(n),
                    np.full(n, np.inf), name="ragged")


CASES = ["afiro", "random_small", "random_mid", "transport", "ragged"]


@pytest.fixture(scope="module", params=CASES)
def problem(request, capi, oracle):
    lp = _case(request.param)
    AIp, AIi, AIx = lp.solver_form()
    ctx = capi.Context(lp.m, lp.n, AIp, AIi, AIx)
    A = oracle.Csc(AIp, AIi, AIx)
    yield lp, ctx, A
    ctx.close()


@pytest.mark.parametrize("regime", ["ones", "mid", "wide", "null"])
def test_normal_apply(problem, oracle, regime):
    lp, ctx, A = problem
    m, n = lp.m, lp.n
    W = None if regime == "null" else lpgen.weights(n + m, regime, 5)
    x = np.random.default_rng(9).standard_normal(m)
    ctx.normal_prepare(W)
    y, dot = ctx.normal_apply(x)
    y0, dot0 = oracle.normal_apply(m, n, A, W, x)
    assert rel_err(y, y0) <= APPLY_TOL
    # The dot is a sum of products of the compared vectors: same tolerance
    # relative to sum |x_i y_i|.
    assert abs(dot - dot0) <= APPLY_TOL * np.abs(x * y0).sum()
    # deterministic: a second apply returns the same bits
    y2, dot2 = ctx.normal_apply(x)
    assert np.array_equal(y, y2) and dot == dot2


@pytest.mark.parametrize("regime", ["ones", "wide", "null"])
def test_diag_build_and_apply(problem, oracle, regime):
    lp, ctx, A = problem
    m, n = lp.m, lp.n
    W = None if regime == "null" else lpgen.weights(n + m, regime, 6)
    ctx.diag_factorize(W)
    d = ctx.diag_get()
    d0 = oracle.diag_build(m, n, A, W)
    assert rel_err(d, d0) <= APPLY_TOL
    if np.all(d0 > 0):
        x = np.random.default_rng(10).standard_normal(m)
        ctx.diag_set(d0)
        l, dot = ctx.diag_apply(x)
        l0, dot0 = oracle.diag_apply(d0, x)
        assert np.array_equal(l, l0)
        assert abs(dot - dot0) <= 1e-13 * np.abs(l0 * x).sum()


@pytest.mark.parametrize("tol", [1e-2, 1e-8])
def test_pcr_matches_oracle(problem, oracle, tol):
    lp, ctx, A = problem
    m, n = lp.m, lp.n
    W = lpgen.weights(n + m, "mid", 7)
    rng = np.random.default_rng(11)
    rhs = rng.standard_normal(m)
    resscale = 1.0 / np.sqrt(W[n:])
    diag = oracle.diag_build(m, n, A, W)
    ctx.normal_prepare(W)
    ctx.diag_factorize(None, use_prepared=True)
    y, info = ctx.pcr_solve(rhs, tol, resscale, -1, hist_cap=4096)
    y0, info0 = oracle.pcr_solve(oracle.normal_operator(m, n, A, W), m, diag, rhs, tol, resscale,
                                 -1, hist_cap=4096)
    assert info["errflag"] == info0["errflag"]
    # CR amplifies rounding on ill-conditioned systems (the ragged case has a dense column), so
    # the bar is the oracle's own sensitivity: the same solve with a right-hand side that differs
    # by one ulp per entry. Where the oracle agrees with itself to rounding, so must the device
    # (1e-9 on the residual norms and on the iterate); where it does not, the device may be a
    # few times as far from the oracle as the oracle is from itself. One pass more or fewer is
    # allowed where a residual norm lies within rounding of the tolerance.
    ulp = np.spacing(np.abs(rhs)) * rng.choice([-1.0, 1.0], m)
    y1, info1 = oracle.pcr_solve(oracle.normal_operator(m, n, A, W), m, diag, rhs + ulp, tol,
                                 resscale, -1, hist_cap=4096)
    h, h0, h1 = info["hist"], info0["hist"], info1["hist"]
    k = min(len(h), len(h0), len(h1))
    self_dev = np.maximum.accumulate(np.abs(h1[:k] - h0[:k]) / h0[:k])
    assert np.all(np.abs(h[:k] - h0[:k]) <= np.maximum(1e-9, 50 * self_dev) * h0[:k])
    assert abs(info["iter"] - info0["iter"]) <= 1
    if info["iter"] == info0["iter"]:
        self_err = rel_err(y1, y0) if info1["iter"] == info0["iter"] else 1.0
        assert rel_err(y, y0) <= max(1e-9, 50 * self_err)
    # the returned iterate solves the system to the requested accuracy
    Cy, _ = oracle.normal_apply(m, n, A, W, y)
    if info["errflag"] == 0:
        assert np.abs(resscale * (rhs - Cy)).max() <= tol * (1 + 1e-6) + 1e-12


def test_pcr_nonzero_start_and_iter_limit(problem, oracle):
    lp, ctx, A = problem
    m, n = lp.m, lp.n
    W = lpgen.weights(n + m, "mid", 8)
    rng = np.random.default_rng(12)
    rhs, y_init = rng.standard_normal(m), 0.1 * rng.standard_normal(m)
    diag = oracle.diag_build(m, n, A, W)
    ctx.normal_prepare(W)
    ctx.diag_factorize(None, use_prepared=True)
    y, info = ctx.pcr_solve(rhs, 1e-30, None, 3, lhs0=y_init)
    y0, info0 = oracle.pcr_solve(oracle.normal_operator(m, n, A, W), m, diag, rhs, 1e-30, None, 3,
                                 lhs0=y_init)
    assert info0["errflag"] == 201 and info["errflag"] == 201
    assert info["iter"] == info0["iter"] == 3
    assert rel_err(y, y0) <= 1e-10


def test_cr_unpreconditioned(problem, oracle):
    lp, ctx, A = problem
    m, n = lp.m, lp.n
    W = lpgen.weights(n + m, "mid", 13)
    rhs = np.random.default_rng(14).standard_normal(m)
    ctx.normal_prepare(W)
    y, info = ctx.cr_solve(0, rhs, 1e-6, None, 200, hist_cap=256)
    y0, info0 = oracle.cr_solve(oracle.normal_operator(m, n, A, W), m, rhs, 1e-6, None, 200,
                                hist_cap=256)
    assert info["errflag"] == info0["errflag"]
    k = min(len(info["hist"]), len(info0["hist"]), 5)
    assert np.allclose(info["hist"][:k], info0["hist"][:k], rtol=1e-9, atol=0.0)
    if info0["errflag"] == 0:
        # converged: same count up to a residual sitting on the tolerance
        assert abs(info["iter"] - info0["iter"]) <= max(1, info0["iter"] // 20)
        assert rel_err(y, y0) <= 1e-4
    else:
        assert info["iter"] == info0["iter"] == 200


def test_pcr_not_posdef_flag(capi, oracle):
    """Negative weights make v'Cv <= 0: errflag 202 like the reference."""
    lp = lpgen.random_sparse_lp(200, 1000, 4, 31)
    AIp, AIi, AIx = lp.solver_form()
    ctx = capi.Context(lp.m, lp.n, AIp, AIi, AIx)
    A = oracle.Csc(AIp, AIi, AIx)
    W = -np.ones(lp.n + lp.m)
    rhs = np.random.default_rng(1).standard_normal(lp.m)
    ctx.normal_prepare(W)
    ctx.diag_set(np.ones(lp.m))
    _, info = ctx.pcr_solve(rhs, 1e-8, None, 50)
    _, info0 = oracle.pcr_solve(oracle.normal_operator(lp.m, lp.n, A, W), lp.m, np.ones(lp.m), rhs,
                                1e-8, None, 50)
    assert info0["errflag"] == 202 and info["errflag"] == 202
    assert info["iter"] == info0["iter"]
    ctx.close()


def test_kktdiag_factorize_and_solve(problem, oracle):
    lp, ctx, A = problem
    m, n = lp.m, lp.n
    rng = np.random.default_rng(15)
    nm = n + m
    xl = rng.uniform(0.1, 10.0, nm)
    xu = rng.uniform(0.1, 10.0, nm)
    zl = rng.uniform(0.01, 5.0, nm)
    zu = rng.uniform(0.01, 5.0, nm)
    # free variables: g == 0 -> W = 1/regval
    free = rng.random(nm) < 0.05
    zl[free] = 0.0
    zu[free] = 0.0
    xl[free] = np.inf
    xu[free] = np.inf
    mu = 0.37
    W, resscale = ctx.kktdiag_factorize(xl, xu, zl, zu, mu, want_W=True)
    W0, resscale0 = oracle.kktdiag_weights(m, n, xl, xu, zl, zu, mu)
    assert np.array_equal(W, W0)
    assert np.array_equal(resscale, resscale0)
    diag0 = oracle.diag_build(m, n, A, W0)
    assert rel_err(ctx.diag_get(), diag0) <= APPLY_TOL
    a, b = rng.standard_normal(nm), rng.standard_normal(m)
    x, y, info = ctx.kktdiag_solve(a, b, 1e-8, -1)
    x0, y0, info0 = oracle.kktdiag_solve(m, n, A, W0, diag0, resscale0, a, b, 1e-8, -1)
    assert info["errflag"] == info0["errflag"] == 0
    assert abs(info["iter"] - info0["iter"]) <= 1
    assert rel_err(y, y0) <= 1e-6
    assert rel_err(x, x0) <= 1e-6
    # KKT residual: AI x = b exactly by construction of the recovery
    import scipy.sparse as sp
    AIp, AIi, AIx = lp.solver_form()
    AI = sp.csc_matrix((AIx, AIi, AIp), shape=(m, nm))
    assert np.abs(AI @ x - b).max() <= 1e-9 * (1 + np.abs(b).max() + np.abs(x).max())


def test_kktdiag_identity_weights(problem, oracle):
    lp, ctx, A = problem
    m, n = lp.m, lp.n
    W, resscale = ctx.kktdiag_factorize(want_W=True)
    assert np.all(W == 1.0) and np.all(resscale == 1.0)


def test_launch_counter_moves(problem):
    lp, ctx, A = problem
    before = ctx.launch_count()
    ctx.normal_prepare(None)
    ctx.normal_apply(np.ones(lp.m))
    assert ctx.launch_count() > before


@pytest.mark.parametrize("kind", ["random_mid", "transport", "ragged", "tall"])
def test_band_sweeps_forced(capi, oracle, kind, monkeypatch):
    """The banded shared-memory sweeps, forced on small shapes (IPXGPU_SWEEP=band), give the
    same operator as the oracle."""
    monkeypatch.setenv("IPXGPU_SWEEP", "band")
    lp = _case(kind)
    m, n = lp.m, lp.n
    AIp, AIi, AIx = lp.solver_form()
    ctx = capi.Context(m, n, AIp, AIi, AIx)
    monkeypatch.delenv("IPXGPU_SWEEP")
    A = oracle.Csc(AIp, AIi, AIx)
    tiling = ctx.tiling()
    if kind == "random_mid":
        assert tiling["sweep1"]["enabled"] == 1 and tiling["sweep2"]["enabled"] == 1
    x = np.random.default_rng(19).standard_normal(m)
    for regime in ("mid", "wide", "null"):
        W = None if regime == "null" else lpgen.weights(n + m, regime, 5)
        ctx.normal_prepare(W)
        y, dot = ctx.normal_apply(x)
        y0, dot0 = oracle.normal_apply(m, n, A, W, x)
        assert rel_err(y, y0) <= APPLY_TOL
        assert abs(dot - dot0) <= APPLY_TOL * np.abs(x * y0).sum()
        y2, dot2 = ctx.normal_apply(x)
        assert np.array_equal(y, y2) and dot == dot2
    W = lpgen.weights(n + m, "mid", 7)
    ctx.normal_prepare(W)
    ctx.diag_factorize(None, use_prepared=True)
    rhs = np.random.default_rng(20).standard_normal(m)
    z, info = ctx.pcr_solve(rhs, 1e-8, None, -1)
    Cz, _ = oracle.normal_apply(m, n, A, W, z)
    if info["errflag"] == 0:
        assert np.abs(rhs - Cz).max() <= 1e-8 * (1 + 1e-6) + 1e-12
    ctx.close()


def test_band_sweeps_auto(capi, oracle):
    """A shape large enough that the planner picks the banded sweeps on its own:
    apply parity in every weight regime, run-to-run determinism, the generic
    sweeps (IPXGPU_SWEEP=generic) as a second opinion, and a PCR solve."""
    import os
    lp = lpgen.random_sparse_lp(20000, 400000, 10, 31)
    m, n = lp.m, lp.n
    AIp, AIi, AIx = lp.solver_form()
    ctx = capi.Context(m, n, AIp, AIi, AIx)
    tiling = ctx.tiling()
    assert tiling["sweep1"]["enabled"] == 1 and tiling["sweep2"]["enabled"] == 1, tiling
    os.environ["IPXGPU_SWEEP"] = "generic"
    try:
        ctx_gen = capi.Context(m, n, AIp, AIi, AIx)
    finally:
        del os.environ["IPXGPU_SWEEP"]
    assert ctx_gen.tiling()["sweep1"]["enabled"] == 0
    A = oracle.Csc(AIp, AIi, AIx)
    x = np.random.default_rng(41).standard_normal(m)
    for regime in ("ones", "mid", "wide", "null"):
        W = None if regime == "null" else lpgen.weights(n + m, regime, 6)
        ctx.normal_prepare(W)
        ctx_gen.normal_prepare(W)
        y, dot = ctx.normal_apply(x)
        y0, dot0 = oracle.normal_apply(m, n, A, W, x)
        assert rel_err(y, y0) <= APPLY_TOL
        assert abs(dot - dot0) <= APPLY_TOL * np.abs(x * y0).sum()
        y2, dot2 = ctx.normal_apply(x)
        assert np.array_equal(y, y2) and dot == dot2
        yg, dotg = ctx_gen.normal_apply(x)
        assert rel_err(yg, y0) <= APPLY_TOL
    W = lpgen.weights(n + m, "mid", 8)
    ctx.normal_prepare(W)
    ctx.diag_factorize(None, use_prepared=True)
    rhs = np.random.default_rng(42).standard_normal(m)
    z, info = ctx.pcr_solve(rhs, 1e-8, None, -1)
    Cz, _ = oracle.normal_apply(m, n, A, W, z)
    assert info["errflag"] == 0
    assert np.abs(rhs - Cz).max() <= 1e-8 * (1 + 1e-6) + 1e-12
    ctx.close()
    ctx_gen.close()


def _with_dense_row_and_column(lp, rng):
    """Adds one dense column (every 3rd row) and one dense row (every 2nd column)."""
    m, n = lp.m, lp.n
    cols = [lp.Ai[lp.Ap[j]:lp.Ap[j + 1]] for j in range(n)]
    dense_row = m - 1
    Ap = np.zeros(n + 2, np.int64)
    out_i, out_x = [], []
    for j in range(n):
        rows = cols[j]
        if j % 2 == 0 and dense_row not in rows:
            rows = np.append(rows, dense_row)
        out_i.append(rows)
        Ap[j + 1] = Ap[j] + len(rows)
    dense_col = np.arange(0, m, 3, dtype=np.int64)
    out_i.append(dense_col)
    Ap[n + 1] = Ap[n] + len(dense_col)
    Ai = np.concatenate(out_i)
    Ax = rng.uniform(0.5, 4.0, len(Ai)) * rng.choice([-1.0, 1.0], len(Ai))
    n2 = n + 1
    return lpgen.LP(m, n2, Ap, Ai, Ax, np.zeros(m), b"=" * m, np.zeros(n2), np.zeros(n2),
                    np.full(n2, np.inf), name="dense_row_and_column")


def test_band_sweeps_spill_dense_segments(capi, oracle):
    """A well-spread matrix with one dense row and one dense column: the banded sweeps keep
    everything else and leave the two dense segments to the generic kernel (tiling 'enabled'
    = 2); transportation rows (config 4 in small) take the same route in sweep 2."""
    rng = np.random.default_rng(91)
    base = lpgen.random_sparse_lp(20000, 400000, 10, 33)
    cases = [_with_dense_row_and_column(base, rng), lpgen.transportation_lp(400, 5000, 34)]
    for lp in cases:
        m, n = lp.m, lp.n
        AIp, AIi, AIx = lp.solver_form()
        ctx = capi.Context(m, n, AIp, AIi, AIx)
        tiling = ctx.tiling()
        assert tiling["sweep2"]["enabled"] == 2, (lp.name, tiling)
        if lp.name == "dense_row_and_column":
            assert tiling["sweep1"]["enabled"] == 2, tiling
        A = oracle.Csc(AIp, AIi, AIx)
        x = rng.standard_normal(m)
        for regime in ("ones", "wide", "null"):
            W = None if regime == "null" else lpgen.weights(n + m, regime, 6)
            ctx.normal_prepare(W)
            y, dot = ctx.normal_apply(x)
            y0, dot0 = oracle.normal_apply(m, n, A, W, x)
            assert rel_err(y, y0) <= APPLY_TOL, (lp.name, regime)
            assert abs(dot - dot0) <= APPLY_TOL * np.abs(x * y0).sum()
            y2, dot2 = ctx.normal_apply(x)
            assert np.array_equal(y, y2) and dot == dot2
        W = lpgen.weights(n + m, "mid", 8)
        ctx.normal_prepare(W)
        ctx.diag_factorize(None, use_prepared=True)
        d0 = oracle.diag_build(m, n, A, W)
        rhs = rng.standard_normal(m)
        zg, info = ctx.pcr_solve(rhs, 1e-30, None, 15)
        zo, info0 = oracle.pcr_solve(oracle.normal_operator(m, n, A, W), m, d0, rhs, 1e-30, None, 15)
        assert info["iter"] == info0["iter"] and info["errflag"] == info0["errflag"]
        assert rel_err(zg, zo) <= 1e-8
        ctx.close()


def test_degenerate_shapes(capi, oracle):
    """No structural columns (a dualized LP without constraints, reference
    check/solver.cc:153-185) and no rows."""
    # n = 0: AI = I
    m = 3
    AIp, AIi, AIx = np.arange(m + 1), np.arange(m), np.ones(m)
    ctx = capi.Context(m, 0, AIp, AIi, AIx)
    W = np.array([0.5, 2.0, 4.0])
    x = np.array([1.0, -2.0, 3.0])
    ctx.normal_prepare(W)
    y, dot = ctx.normal_apply(x)
    assert np.array_equal(y, W * x) and dot == float(x @ (W * x))
    ctx.diag_factorize(W)
    assert np.array_equal(ctx.diag_get(), W)
    z, info = ctx.pcr_solve(x, 1e-12, None, -1)
    assert info["errflag"] == 0 and np.allclose(z, x / W, rtol=1e-14)
    Wk, rs = ctx.kktdiag_factorize(want_W=True)
    xx, yy, info = ctx.kktdiag_solve(np.array([1.0, 2.0, 3.0]), np.array([0.5, 0.5, 0.5]), 1e-10, -1)
    A = oracle.Csc(AIp, AIi, AIx)
    x0, y0, info0 = oracle.kktdiag_solve(m, 0, A, Wk, np.ones(m), rs, np.array([1.0, 2.0, 3.0]),
                                         np.array([0.5, 0.5, 0.5]), 1e-10, -1)
    assert info["errflag"] == info0["errflag"] == 0
    assert np.allclose(xx, x0, rtol=1e-12) and np.allclose(yy, y0, rtol=1e-12)
    ctx.close()
    # m = 0
    n = 4
    ctx = capi.Context(0, n, np.zeros(n + 1, np.int64), np.zeros(0, np.int64), np.zeros(0))
    ctx.normal_prepare(np.ones(n))
    y, dot = ctx.normal_apply(np.zeros(0))
    assert y.size == 0
    ctx.close()


def test_interrupt_callback(capi):
    lp = lpgen.random_sparse_lp(2000, 20000, 8, 41)
    AIp, AIi, AIx = lp.solver_form()
    ctx = capi.Context(lp.m, lp.n, AIp, AIi, AIx)
    W = lpgen.weights(lp.n + lp.m, "wide", 3)
    ctx.normal_prepare(W)
    ctx.diag_factorize(None, use_prepared=True)
    calls = []

    def interrupt(_):
        calls.append(1)
        return 999 if len(calls) >= 3 else 0

    _, info = ctx.pcr_solve(np.ones(lp.m), 1e-300, None, 100000, interrupt=interrupt)
    assert info["errflag"] == 999
    assert len(calls) == 3
    ctx.close()


@pytest.mark.parametrize("exchange", ["auto", "two", "pull"])
def test_sharded_solve_over_nvlink_peer_memory(exchange):
    """Two ranks (one per GPU, torchrun): column shards, the persistent CR kernel sums the
    ranks' partial products over NVLink peer memory - one-hop records (auto at 2 ranks), the
    two-hop reduce-scatter + all-gather of records used from 4 ranks on, and the flag + P2P-load
    exchange. Needs two GPUs on the box; tools/check_sharded.py compares with the single-GPU
    solve and checks that all ranks hold identical iterates."""
    import os
    import subprocess
    import sys
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs")
    repo = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    env = dict(os.environ, IPXGPU_XCHG=exchange)
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1",
                        "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
                        "--master-port", "29517", os.path.join(repo, "tools", "check_sharded.py")],
                       capture_output=True, text=True, timeout=600, env=env)
    assert r.returncode == 0 and "SHARDED PARITY PASS" in r.stdout, r.stdout[-2000:] + r.stderr[-2000:]


def test_model_without_structural_columns(capi):
    """n = 0: AI is the identity, C = diag(W) (reference src/normal_matrix.cc:65-66 with an empty
    column loop); apply, diagonal and a preconditioned CR solve."""
    m = 50
    AIp = np.arange(m + 1, dtype=np.int64)
    AIi = np.arange(m, dtype=np.int64)
    AIx = np.ones(m)
    ctx = capi.Context(m, 0, AIp, AIi, AIx)
    W = lpgen.weights(m, "mid", 3)
    x = np.random.default_rng(4).standard_normal(m)
    ctx.normal_prepare(W)
    y, dot = ctx.normal_apply(x)
    assert np.array_equal(y, W * x)
    assert abs(dot - x @ (W * x)) <= 1e-13 * np.abs(x * W * x).sum()
    ctx.diag_factorize(W)
    assert np.array_equal(ctx.diag_get(), W)
    z, info = ctx.pcr_solve(x, 1e-12, None, -1)
    assert info["errflag"] == 0 and info["iter"] <= 2
    assert np.abs(W * z - x).max() <= 1e-12 * np.abs(x).max()
    ctx.close()
