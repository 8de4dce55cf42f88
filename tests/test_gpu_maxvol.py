"""Maxvolume on the device: the column sweeps through the C ABI against
the numpy restatement of reference src/maxvolume.cc:170-320, and Maxvolume::RunHeuristic of the
drop-in build against the compiled reference on the same basis."""

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
            lpgen.dense_column_lp(600, 4000, 5, 3, 34, name="dense_cols")]


@pytest.mark.parametrize("lp", _lps(), ids=lambda lp: lp.name)
def test_column_sweeps_match_restatement(capi, oracle, lp):
    m, n = lp.m, lp.n
    AIp, AIi, AIx = lp.solver_form()
    ctx = capi.Context(m, n, AIp, AIi, AIx)
    rng = np.random.default_rng(41)
    colscale = np.exp(rng.uniform(-6, 6, n + m))
    basic = rng.choice(n + m, m, replace=False)
    colscale[basic] = 0.0                              # basic columns carry no factor
    work = rng.standard_normal(m) * (rng.random(m) < 0.3)
    # Sums of two terms do not depend on their order: identical bits. Longer columns are summed
    # in storage order in most tiles of the sweep and by 2..32 lanes in tiles with few columns.
    tol = 0.0 if np.diff(AIp).max() <= 2 else 1e-14

    def same(got, want):
        err = np.abs(got - want).max() / max(np.abs(want).max(), 1e-300)
        assert err <= tol, err
        return True

    top = ctx.maxvol_weights(colscale, work)
    cw0 = oracle.maxvol_weights(AIp, AIi, AIx, colscale, work)
    cs, cw = ctx.maxvol_get()
    assert np.array_equal(cs, colscale) and same(cw, cw0)
    assert (top["jmax2"], top["jmax"]) == oracle.find_largest(cw)
    assert top["wmax"] == abs(cw[top["jmax"]]) and top["wmax2"] == abs(cw[top["jmax2"]])

    for step in range(4):
        # a column given up (both on the device and in the restatement), then an exchange
        js = top["jmax"]
        top = ctx.maxvol_skip(js)
        colscale[js] = 0.0
        cw0 = cw.copy()
        cw0[js] = 0.0
        cs, cw = ctx.maxvol_get()
        assert np.array_equal(cs, colscale) and np.array_equal(cw, cw0)
        assert (top["jmax2"], top["jmax"]) == oracle.find_largest(cw)

        jn = top["jmax"]
        jb = int(basic[step]) if step % 2 == 0 else int(n + rng.integers(m))
        if colscale[jb] != 0.0 or jb == jn:
            jb = int(basic[step])
        btran = rng.standard_normal(m) * (rng.random(m) < 0.5)
        alpha, cs_jb, cw_jb = rng.standard_normal(), np.exp(rng.uniform(-3, 3)), rng.random()
        top = ctx.maxvol_update(btran, alpha, jb, cs_jb, cw_jb, jn)
        colscale, cw0 = oracle.maxvol_update(AIp, AIi, AIx, colscale, cw, btran, alpha, jb, cs_jb,
                                             cw_jb, jn)
        cs, cw = ctx.maxvol_get()
        assert np.array_equal(cs, colscale) and same(cw, cw0)
        assert (top["jmax2"], top["jmax"]) == oracle.find_largest(cw)

    # the next slice of a run keeps the resident factors
    work2 = rng.standard_normal(m)
    top = ctx.maxvol_weights(None, work2)
    cs, cw = ctx.maxvol_get()
    assert np.array_equal(cs, colscale)
    assert same(cw, oracle.maxvol_weights(AIp, AIi, AIx, colscale, work2))
    ctx.maxvol_release()
    with pytest.raises(capi.IpxGpuError):
        ctx.maxvol_skip(0)
    ctx.close()


def test_find_largest_ties_and_empty(capi, oracle):
    """Equal weights: the first column wins, the second one follows; nothing positive: column 0
    with weight 0 (FindLargest's initial state, src/maxvolume.cc:172-177)."""
    lp = lpgen.random_sparse_lp(50, 400, 3, 35)
    m, n = lp.m, lp.n
    AIp, AIi, AIx = lp.solver_form()
    ctx = capi.Context(m, n, AIp, AIi, AIx)
    # slack columns: weight = work[i] * colscale[n+i], so the weights can be set at will
    colscale = np.zeros(n + m)
    colscale[n:] = 1.0
    for work in (np.zeros(m), np.r_[np.zeros(7), 3.0, np.zeros(m - 8)],
                 np.r_[0.0, 2.0, -2.0, 1.0, 2.0, np.zeros(m - 5)],
                 np.r_[-5.0, np.zeros(m - 1)], np.full(m, -1.5)):
        top = ctx.maxvol_weights(colscale, work)
        _, cw = ctx.maxvol_get()
        assert (top["jmax2"], top["jmax"]) == oracle.find_largest(cw), work[:8]
    ctx.close()


@pytest.mark.parametrize("case", ["random", "transport", "block"])
def test_run_heuristic_matches_reference(reflib, gpulib, case):
    """Maxvolume::RunHeuristic of the drop-in build (device weights, host pivoting) and of the
    compiled reference on the same crash basis and scaling factors: same exchanges."""
    if case == "random":
        lp = lpgen.random_sparse_lp(400, 3000, 4, 51)
    elif case == "transport":
        lp = lpgen.transportation_lp(30, 200, 52)
    else:
        lp = lpgen.block_angular_lp(1200, 9000, 4, 53)
    out = []
    for lib in (reflib, gpulib):
        mdl = lib.model(lp)
        m, n = mdl.m, mdl.n
        rng = np.random.default_rng(54)
        colscale = np.exp(rng.uniform(-4, 4, n + m))
        assert mdl.basis_from_weights(colscale) == 0
        res = mdl.maxvolume(colscale, heuristic=True)
        basis, status = mdl.basis_get()
        out.append((res, basis.copy(), status.copy()))
    (r0, b0, s0), (r1, b1, s1) = out
    assert r0["err"] == r1["err"] == 0
    assert r0["updates"] > 0
    assert (r1["updates"], r1["skipped"], r1["slices"]) == (r0["updates"], r0["skipped"], r0["slices"])
    assert np.array_equal(b0, b1) and np.array_equal(s0, s1)
    assert abs(r1["volinc"] - r0["volinc"]) <= 1e-9 * max(1.0, abs(r0["volinc"]))


def test_run_sequential_matches_reference(reflib, gpulib):
    lp = lpgen.random_sparse_lp(200, 1200, 4, 55)
    out = []
    for lib in (reflib, gpulib):
        mdl = lib.model(lp)
        colscale = np.exp(np.random.default_rng(56).uniform(-4, 4, mdl.n + mdl.m))
        assert mdl.basis_from_weights(colscale) == 0
        res = mdl.maxvolume(colscale, heuristic=False)
        out.append((res, mdl.basis_get()[0].copy()))
    (r0, b0), (r1, b1) = out
    assert r0["err"] == r1["err"] == 0 and r0["updates"] > 0
    for key in ("updates", "skipped", "passes", "volinc"):
        assert r0[key] == r1[key]
    assert np.array_equal(b0, b1)
