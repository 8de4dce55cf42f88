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
    """Transportation rows hold thousands of entries inside one band: sweep 2 must be
    refused (the segmented generic kernel handles it), sweep 1 (2 entries per column) not."""
    lp = lpgen.transportation_lp(20, 3000, 23)
    x = np.random.default_rng(3).standard_normal(lp.m)
    r = capi.band_selftest(lp.m, lp.n, *_structural(lp), x, force=True)
    assert r["sweep2"]["planned"] == 0.0
    assert r["sweep1"]["planned"] == 1.0 and r["sweep1"]["err"] <= 1e-12


def test_small_problems_are_not_planned_without_force(capi):
    lp = lpgen.random_sparse_lp(300, 2000, 5, 21)
    x = np.ones(lp.m)
    r = capi.band_selftest(lp.m, lp.n, *_structural(lp), x, force=False)
    # either refused by the cost model or accepted with a correct stream; never wrong
    for sweep in ("sweep1", "sweep2"):
        if r[sweep]["planned"]:
            assert r[sweep]["err"] <= 1e-10
