"""ipx::MultiplyAdd on the resident AI (the residual products of
Iterate::ComputeResiduals, reference src/iterate.cc:543-551 -> src/sparse_matrix.cc:194-209):
bit-identical to the C restatement of the reference's loops through the C ABI, and bit-identical
to the compiled reference through the drop-in build, where ipx::MultiplyAdd goes to the device as
soon as the model has a context and to the reference's own function before."""

import numpy as np
import pytest

from ipx_b200 import lpgen

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def capi():
    from ipx_b200 import capi as c
    c.load()
    return c


def _lps():
    return [lpgen.random_sparse_lp(300, 2000, 4, 31, name="random_small"),
            lpgen.transportation_lp(40, 300, 32, name="transport"),
            lpgen.random_sparse_lp(4000, 50000, 10, 33, name="random_mid"),
            lpgen.dense_column_lp(600, 4000, 5, 3, 34, name="dense_cols"),
            lpgen.random_sparse_lp(50, 400, 10, 35, name="rows_longer_than_a_warp")]


def _vectors(rng, size):
    v = rng.standard_normal(size) * np.exp(rng.uniform(-8, 8, size))
    v[rng.random(size) < 0.1] = 0.0
    v[rng.random(size) < 0.02] = -0.0
    return v


@pytest.mark.parametrize("lp", _lps(), ids=lambda lp: lp.name)
def test_multiply_add_bit_identical_to_restatement(capi, oracle, lp):
    m, n = lp.m, lp.n
    AIp, AIi, AIx = lp.solver_form()
    A = oracle.Csc(AIp, AIi, AIx)
    ctx = capi.Context(m, n, AIp, AIi, AIx)
    rng = np.random.default_rng(51)
    for alpha in (-1.0, 1.0, 0.37):
        x, lhs_m = _vectors(rng, n + m), _vectors(rng, m)
        got = ctx.multiply_add(x, alpha, lhs_m, "N")
        want = oracle.multiply_add(m, n + m, A, x, alpha, lhs_m, "N")
        assert got.tobytes() == want.tobytes()
        y, lhs_n = _vectors(rng, m), _vectors(rng, n + m)
        got = ctx.multiply_add(y, alpha, lhs_n, "T")
        want = oracle.multiply_add(m, n + m, A, y, alpha, lhs_n, "T")
        assert got.tobytes() == want.tobytes()
    # several column panels: the row sums continue from panel to panel in column order
    ctx2 = capi.Context(m, n, AIp, AIi, AIx, panel_cols=max(1, n // 3))
    x, lhs_m = _vectors(rng, n + m), _vectors(rng, m)
    assert (ctx2.multiply_add(x, -1.0, lhs_m, "N").tobytes()
            == oracle.multiply_add(m, n + m, A, x, -1.0, lhs_m, "N").tobytes())
    y, lhs_n = _vectors(rng, m), _vectors(rng, n + m)
    assert (ctx2.multiply_add(y, -1.0, lhs_n, "T").tobytes()
            == oracle.multiply_add(m, n + m, A, y, -1.0, lhs_n, "T").tobytes())


def test_multiply_add_without_structural_columns(capi, oracle):
    m = 7
    AIp, AIi, AIx = np.arange(m + 1), np.arange(m), np.ones(m)   # AI = I
    A = oracle.Csc(AIp, AIi, AIx)
    ctx = capi.Context(m, 0, AIp, AIi, AIx)
    rng = np.random.default_rng(52)
    x, l = rng.standard_normal(m), rng.standard_normal(m)
    assert (ctx.multiply_add(x, -1.0, l, "N").tobytes()
            == oracle.multiply_add(m, m, A, x, -1.0, l, "N").tobytes())
    assert (ctx.multiply_add(x, 2.5, l, "T").tobytes()
            == oracle.multiply_add(m, m, A, x, 2.5, l, "T").tobytes())


@pytest.mark.parametrize("lp", _lps()[:4], ids=lambda lp: lp.name)
def test_dropin_multiply_add_matches_compiled_reference(reflib, gpulib, lp):
    """ipx::MultiplyAdd(model.AI(), ...) of the drop-in build: the reference's own function
    while the model has no device context, the device once it has one - the same bits as the
    compiled reference either way (the presolved, scaled AI of both builds is the same)."""
    ref, gpu = reflib.model(lp), gpulib.model(lp)
    m, n = ref.m, ref.n
    rng = np.random.default_rng(53)
    x, lhs_m = _vectors(rng, n + m), _vectors(rng, m)
    y, lhs_n = _vectors(rng, m), _vectors(rng, n + m)
    want_n = ref.multiply_add_AI(x, -1.0, lhs_m, "N")
    want_t = ref.multiply_add_AI(y, -1.0, lhs_n, "T")
    # no context yet
    assert gpu.multiply_add_AI(x, -1.0, lhs_m, "N").tobytes() == want_n.tobytes()
    assert gpu.multiply_add_AI(y, -1.0, lhs_n, "T").tobytes() == want_t.tobytes()
    # NormalMatrix::Prepare creates the model's context: from here on the device
    gpu.normal_prepare(np.ones(n + m))
    assert gpu.multiply_add_AI(x, -1.0, lhs_m, "N").tobytes() == want_n.tobytes()
    assert gpu.multiply_add_AI(y, -1.0, lhs_n, "T").tobytes() == want_t.tobytes()
