"""Host-side logic of the banded sweeps (ipx_b200/csrc/band_sweep.cuh, band_plan.cuh):
planner and row-stream builder, checked by walking the streams on the host exactly as
the kernel does (ipxgpu_band_selftest) - no device needed. The operator it restates is
NormalMatrix::_Apply (reference src/normal_matrix.cc:65-75)."""

import numpy as np
import pytest

from ipx_b200 import lpgen


@pytest.fixture(scope="module")
def capi():
    from ipx_b200 import capi as c
    c.load()
    return c


def _structural(lp):
    return lp.Ap, lp.Ai, lp.Ax


def test_planner_accepts_the_benchmark_like_shape(capi):
    """A scaled-down BASELINE configs[1] shape: both sweeps planned, streams exact up to
    rounding, little padding."""
    lp = lpgen.random_sparse_lp(20000, 400000, 10, 31)
    x = np.random.default_rng(1).standard_normal(lp.m)
    r = capi.band_selftest(lp.m, lp.n, *_structural(lp), x)
    for sweep in ("sweep1", "sweep2"):
        assert r[sweep]["planned"] == 1.0, r
        assert r[sweep]["pad"] < 0.25, r
    # t = A'x: sums of <= 10 products; y = A t: sums of ~200 products of size ~|t|
    assert r["sweep1"]["err"] <= 1e-13 * 10 * 4 * np.abs(x).max()
    assert r["sweep2"]["err"] <= 1e-11


@pytest.mark.parametrize("shape", [(300, 2000, 5), (5000, 60000, 10), (64, 64, 3)])
def test_forced_layout_is_exact_on_small_shapes(capi, shape):
    m, n, k = shape
    lp = lpgen.random_sparse_lp(m, n, k, 5)
    x = np.random.default_rng(2).standard_normal(lp.m)
    r = capi.band_selftest(lp.m, lp.n, *_structural(lp), x, force=True)
    for sweep in ("sweep1", "sweep2"):
        assert r[sweep]["planned"] == 1.0, r
    assert r["sweep1"]["err"] <= 1e-12 and r["sweep2"]["err"] <= 1e-10, r


def test_long_segments_are_left_to_the_generic_sweep(capi):
    """Transportation source rows hold thousands of entries inside one band: sweep 2 leaves
    them out of the streams (planned == 2: the generic kernel sweeps those rows) and carries
    the sink rows; sweep 1 (2 entries per column) takes everything."""
    lp = lpgen.transportation_lp(400, 5000, 23)
    x = np.random.default_rng(3).standard_normal(lp.m)
    r = capi.band_selftest(lp.m, lp.n, *_structural(lp), x)
    assert r["sweep2"]["planned"] == 2.0 and r["sweep2"]["err"] <= 1e-10, r
    assert r["sweep2"]["pad"] < 0.25, r
    assert r["sweep1"]["planned"] == 1.0 and r["sweep1"]["err"] <= 1e-12


def test_dense_row_and_column_are_spilled(capi):
    """One dense row and one dense column in a well-spread matrix: both sweeps stay banded
    and list one spilled segment each (planned == 2); the rest of the streams is exact."""
    base = lpgen.random_sparse_lp(20000, 400000, 10, 33)
    m, n = base.m, base.n
    rng = np.random.default_rng(7)
    # dense row m-1 in every other column (appended last keeps the rows sorted)
    has = base.Ai.reshape(n, 10)[:, -1] == m - 1
    add = (np.arange(n) % 2 == 0) & ~has
    counts = 10 + add.astype(np.int64)
    Ap = np.zeros(n + 2, np.int64)
    Ap[1:n + 1] = np.cumsum(counts)
    dense_col = np.arange(0, m, 3, dtype=np.int64)
    Ap[n + 1] = Ap[n] + len(dense_col)
    Ai = np.empty(Ap[n + 1], np.int64)
    Ax = rng.uniform(0.5, 4.0, Ap[n + 1])
    pos = Ap[:n, None] + np.arange(10)[None, :]
    Ai[pos.reshape(-1)] = base.Ai
    Ai[Ap[1:n + 1][add] - 1] = m - 1
    Ai[Ap[n]:] = dense_col
    x = rng.standard_normal(m)
    r = capi.band_selftest(m, n + 1, Ap, Ai, Ax, x)
    for sweep in ("sweep1", "sweep2"):
        assert r[sweep]["planned"] == 2.0, r
        assert r[sweep]["pad"] < 0.25, r
    assert r["sweep1"]["err"] <= 1e-12 and r["sweep2"]["err"] <= 1e-10, r


def test_small_problems_are_not_planned_without_force(capi):
    lp = lpgen.random_sparse_lp(300, 2000, 5, 21)
    x = np.ones(lp.m)
    r = capi.band_selftest(lp.m, lp.n, *_structural(lp), x, force=False)
    # either refused by the cost model or accepted with a correct stream; never wrong
    for sweep in ("sweep1", "sweep2"):
        if r[sweep]["planned"]:
            assert r[sweep]["err"] <= 1e-10


def test_threaded_build_is_deterministic(capi):
    """The layout is built on several host threads (segment blocks and tiles are independent):
    two builds of the same matrix must agree exactly (same padding, same stream contents as seen
    through the host walk)."""
    lp = lpgen.random_sparse_lp(20000, 400000, 10, 35)
    x = np.random.default_rng(4).standard_normal(lp.m)
    a = capi.band_selftest(lp.m, lp.n, *_structural(lp), x)
    b = capi.band_selftest(lp.m, lp.n, *_structural(lp), x)
    assert a == b
    assert a["sweep1"]["planned"] == 1.0 and a["sweep2"]["planned"] == 1.0
