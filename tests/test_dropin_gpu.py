"""Drop-in parity: IPX built with the GPU translation units
(ipx_b200/_build/libipx_gpu.so) against the reference's own CPU build
(oracle/_ref/libipx_ref.so), driven through identical calls on IPX's own
classes and the unchanged ipx_c.h API.

Bars (BASELINE.json north_star): operator applies within 1e-12 relative;
end-to-end objective within 1e-9 relative, same status, IPM iterations +-1.
"""

import os
import subprocess

import numpy as np
import pytest

from conftest import rel_err
from ipx_b200 import lpgen

pytestmark = pytest.mark.gpu

BUILD = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "ipx_b200",
                     "_build")


def _lps():
    return {
        "afiro": lpgen.afiro_lp(),
        "random": lpgen.random_sparse_lp(800, 8000, 8, 301),
        "transport": lpgen.transportation_lp(12, 40, 302),
    }


@pytest.fixture(scope="module", params=["afiro", "random", "transport"])
def pair(request, reflib, gpulib):
    lp = _lps()[request.param]
    a, b = reflib.model(lp), gpulib.model(lp)
    yield lp, a, b
    a.close()
    b.close()


@pytest.mark.parametrize("regime", ["ones", "mid", "wide", "null"])
def test_normal_matrix_apply(pair, regime):
    lp, ref, gpu = pair
    m, n = ref.m, ref.n
    W = None if regime == "null" else lpgen.weights(n + m, regime, 3)
    x = np.random.default_rng(4).standard_normal(m)
    ref.normal_prepare(W)
    gpu.normal_prepare(W)
    y0, d0 = ref.normal_apply(x)
    y1, d1 = gpu.normal_apply(x)
    assert rel_err(y1, y0) <= 1e-12
    assert abs(d1 - d0) <= 1e-12 * np.abs(x * y0).sum()
    y2, _ = gpu.normal_apply(x, want_dot=False)
    assert np.array_equal(y1, y2)


def test_diagonal_precond(pair):
    lp, ref, gpu = pair
    m, n = ref.m, ref.n
    W = lpgen.weights(n + m, "mid", 5)
    x = np.random.default_rng(6).standard_normal(m)
    assert ref.diag_factorize(W) == 0 and gpu.diag_factorize(W) == 0
    l0, d0 = ref.diag_apply(x)
    l1, d1 = gpu.diag_apply(x)
    assert rel_err(l1, l0) <= 1e-12
    assert abs(d1 - d0) <= 1e-12 * np.abs(x * l0).sum()


def test_conjugate_residuals(pair):
    lp, ref, gpu = pair
    m, n = ref.m, ref.n
    W = lpgen.weights(n + m, "mid", 7)
    rhs = np.random.default_rng(8).standard_normal(m)
    resscale = 1.0 / np.sqrt(W[n:])
    for mdl in (ref, gpu):
        mdl.normal_prepare(W)
        mdl.diag_factorize(W)
    y0, i0 = ref.pcr_solve(rhs, 1e-8, resscale, -1)
    y1, i1 = gpu.pcr_solve(rhs, 1e-8, resscale, -1)
    assert i0["errflag"] == i1["errflag"] == 0
    assert abs(i0["iter"] - i1["iter"]) <= 1
    assert rel_err(y1, y0) <= 1e-6
    z0, j0 = ref.cr_solve_normal(rhs, 1e-6, None, 500)
    z1, j1 = gpu.cr_solve_normal(rhs, 1e-6, None, 500)
    assert j0["errflag"] == j1["errflag"]
    if j0["errflag"] == 0:
        assert abs(j0["iter"] - j1["iter"]) <= max(1, j0["iter"] // 20)
        assert rel_err(z1, z0) <= 1e-4


def _iterate(m, n, lb, ub, seed):
    """A strictly interior iterate consistent with the model's bounds."""
    rng = np.random.default_rng(seed)
    nm = n + m
    x = rng.uniform(0.5, 1.5, nm)
    y = rng.standard_normal(m)
    has_lb, has_ub = np.isfinite(lb), np.isfinite(ub)
    xl = np.where(has_lb, rng.uniform(0.1, 3.0, nm), np.inf)
    xu = np.where(has_ub, rng.uniform(0.1, 3.0, nm), np.inf)
    zl = np.where(has_lb, rng.uniform(0.05, 2.0, nm), 0.0)
    zu = np.where(has_ub, rng.uniform(0.05, 2.0, nm), 0.0)
    return x, xl, xu, y, zl, zu


def test_kkt_solver_diag(pair):
    lp, ref, gpu = pair
    m, n = ref.m, ref.n
    _, _, lb, ub = ref.model_vectors()
    it = _iterate(m, n, lb, ub, 9)
    rng = np.random.default_rng(10)
    a, b = rng.standard_normal(n + m), rng.standard_normal(m)
    out = []
    for mdl in (ref, gpu):
        mdl.iterate_set(*it)
        assert mdl.kktdiag_factorize(True) == 0
        out.append(mdl.kktdiag_solve(a, b, 1e-8))
    (x0, y0, i0), (x1, y1, i1) = out
    assert i0["err"] == i1["err"] == 0
    assert abs(i0["kktiter1"] - i1["kktiter1"]) <= 1
    assert rel_err(y1, y0) <= 1e-6
    assert rel_err(x1, x0) <= 1e-6
    assert i1["time_cr1"] > 0 and i1["time_cr1_AAt"] > 0
    # identity weights: Factorize(nullptr)
    out = []
    for mdl in (ref, gpu):
        assert mdl.kktdiag_factorize(False) == 0
        out.append(mdl.kktdiag_solve(a, b, 1e-6))
    assert abs(out[0][2]["kktiter1"] - out[1][2]["kktiter1"]) <= 1
    assert rel_err(out[1][1], out[0][1]) <= 1e-5


def test_two_operator_instances_on_one_model(pair):
    """The reference's classes each own their weights; the device context of a model holds one
    set. The harness' stand-alone NormalMatrix / DiagonalPrecond and the members of its
    KKTSolverDiag are two instances each on the same Model: whoever is used after the other
    one primed the context must get ITS state back (gpu_bridge.h: ClaimState / EnsurePrimed)."""
    lp, ref, gpu = pair
    m, n = ref.m, ref.n
    _, _, lb, ub = ref.model_vectors()
    it = _iterate(m, n, lb, ub, 21)
    rng = np.random.default_rng(22)
    W1 = lpgen.weights(n + m, "mid", 23)
    x = rng.standard_normal(m)
    a, b = rng.standard_normal(n + m), rng.standard_normal(m)
    out = []
    for mdl in (ref, gpu):
        mdl.normal_prepare(W1)                   # stand-alone instances: weights W1
        assert mdl.diag_factorize(W1) == 0
        mdl.iterate_set(*it)
        assert mdl.kktdiag_factorize(True) == 0  # the solver's members: the iterate's weights
        y_a, _ = mdl.normal_apply(x)             # must still be A*W1*A'
        l_a, _ = mdl.diag_apply(x)
        sol = mdl.kktdiag_solve(a, b, 1e-8)      # and the solver must get its own state back
        y_b, _ = mdl.normal_apply(x)
        out.append((y_a, l_a, sol, y_b))
    (ya0, la0, s0, yb0), (ya1, la1, s1, yb1) = out
    assert rel_err(ya1, ya0) <= 1e-12 and rel_err(yb1, yb0) <= 1e-12
    assert rel_err(la1, la0) <= 1e-12
    assert s0[2]["err"] == s1[2]["err"] == 0
    assert abs(s0[2]["kktiter1"] - s1[2]["kktiter1"]) <= 1
    assert rel_err(s1[1], s0[1]) <= 1e-6 and rel_err(s1[0], s0[0]) <= 1e-6


def test_splitted_normal_matrix_and_basis_cr(pair):
    lp, ref, gpu = pair
    m, n = ref.m, ref.n
    rng = np.random.default_rng(11)
    # The basis takes the heavy columns, as in the IPM (colweights = scaling
    # factors), which keeps C = I + inv(B) N N' inv(B') well conditioned.
    colscale = np.exp(rng.uniform(-3, 3, n + m))
    colweights = colscale
    for mdl in (ref, gpu):
        mdl.basis_from_weights(colweights)
    b0, s0 = ref.basis_get()
    b1, s1 = gpu.basis_get()
    assert np.array_equal(b0, b1) and np.array_equal(s0, s1)  # same host LU provider
    for mdl in (ref, gpu):
        mdl.split_prepare(colscale)
    assert np.array_equal(ref.split_colperm(), gpu.split_colperm())
    x = rng.standard_normal(m)
    y0, d0 = ref.split_apply(x)
    y1, d1 = gpu.split_apply(x)
    assert rel_err(y1, y0) <= 1e-12
    assert abs(d1 - d0) <= 1e-12 * np.abs(x * y0).sum()
    rhs = rng.standard_normal(m)
    z0, i0 = ref.cr_solve_split(rhs, 1e-8, 400)
    z1, i1 = gpu.cr_solve_split(rhs, 1e-8, 400)
    assert i0["errflag"] == i1["errflag"]
    if i0["errflag"] == 0:
        assert abs(i0["iter"] - i1["iter"]) <= max(1, i0["iter"] // 20)
        assert rel_err(z1, z0) <= 1e-5
    else:
        assert i0["iter"] == i1["iter"] == 400


def test_kkt_solver_basis(pair):
    lp, ref, gpu = pair
    m, n = ref.m, ref.n
    _, _, lb, ub = ref.model_vectors()
    it = _iterate(m, n, lb, ub, 12)
    rng = np.random.default_rng(13)
    a, b = rng.standard_normal(n + m), rng.standard_normal(m)
    colweights = 1.0 / (it[4] / it[1] + it[5] / it[2] + 1e-30)
    out = []
    for mdl in (ref, gpu):
        mdl.iterate_set(*it)
        mdl.basis_from_weights(colweights)
        f = mdl.kktbasis_factorize()
        assert f["err"] == 0
        out.append(mdl.kktbasis_solve(a, b, 1e-8))
    (x0, y0, i0), (x1, y1, i1) = out
    assert i0["err"] == i1["err"] == 0
    # several hundred CR iterations on the random LP: the count moves with the rounding of the
    # right-hand side (the device forms inverse(B)(rhs - b) in one solve)
    assert abs(i0["kktiter2"] - i1["kktiter2"]) <= max(1, int(i0["kktiter2"]) // 100)
    assert rel_err(y1, y0) <= 1e-6
    assert rel_err(x1, x0) <= 1e-6


def test_kkt_solver_basis_with_free_variables(pair):
    """BASIC_FREE positions (src/kkt_solver_basis.cc:88-98, :131-137, :166-174) on both arms."""
    lp, ref, gpu = pair
    m, n = ref.m, ref.n
    _, _, lb, ub = ref.model_vectors()
    it = _iterate(m, n, lb, ub, 14)
    rng = np.random.default_rng(15)
    a, b = rng.standard_normal(n + m), rng.standard_normal(m)
    colweights = 1.0 / (it[4] / it[1] + it[5] / it[2] + 1e-30)
    out = []
    for mdl in (ref, gpu):
        mdl.iterate_set(*it)
        mdl.basis_from_weights(colweights)
        basis, _ = mdl.basis_get()
        for p in (0, m // 2):
            mdl.basis_free_variable(int(basis[p]))
        f = mdl.kktbasis_factorize()
        assert f["err"] == 0
        out.append(mdl.kktbasis_solve(a, b, 1e-9))
    (x0, y0, i0), (x1, y1, i1) = out
    assert i0["err"] == i1["err"] == 0
    assert abs(i0["kktiter2"] - i1["kktiter2"]) <= max(1, int(i0["kktiter2"]) // 100)
    assert rel_err(y1, y0) <= 1e-6
    assert rel_err(x1, x0) <= 1e-6


LP_CASES = {
    "afiro": (lpgen.afiro_lp, {}),
    "random_500x5000": (lambda: lpgen.random_sparse_lp(500, 5000, 10, 7), {"dualize": 0}),
    "random_basis_phase": (lambda: lpgen.random_sparse_lp(300, 1500, 6, 17),
                           {"dualize": 0, "switchiter": 3}),
    "transport_15x40": (lambda: lpgen.transportation_lp(15, 40, 1004), {"dualize": 0}),
    "blockangular": (lambda: lpgen.block_angular_lp(600, 4000, 6, 1003, block_rows=50),
                     {"dualize": 0}),
}


@pytest.mark.parametrize("case", sorted(LP_CASES))
def test_lp_solver_end_to_end(case, reflib, gpulib):
    make, params = LP_CASES[case]
    lp = make()
    infos = []
    for lib in (reflib, gpulib):
        s = lib.lp_solver()
        s.set_parameters(display=0, **params)
        assert s.load_model(lp) == 0
        status = s.solve()
        info = s.info()
        infos.append((status, info))
        s.close()
    (st0, i0), (st1, i1) = infos
    assert st0 == st1 == 1000
    assert i0["status_ipm"] == i1["status_ipm"] == 1
    assert i0["status_crossover"] == i1["status_crossover"]
    assert abs(i0["iter"] - i1["iter"]) <= 1
    scale = max(1.0, abs(i0["objval"]))
    assert abs(i0["objval"] - i1["objval"]) <= 1e-9 * scale
    assert abs(i0["pobjval"] - i1["pobjval"]) <= 1e-6 * scale
    if np.isfinite(lp.optimum):
        assert abs(i1["objval"] - lp.optimum) <= 1e-7 * max(1.0, abs(lp.optimum))
    # the path's counters are filled like the reference's
    assert i1["kktiter1"] > 0
    assert (i1["kktiter2"] > 0) == (i0["kktiter2"] > 0)
    assert i1["time_cr1"] > 0


def test_reference_example_and_check_suite_link_unmodified():
    """example/afiro.cc and check/*.cc of the reference, compiled unmodified and
    linked against the GPU build, run on the device."""
    afiro = os.path.join(BUILD, "afiro_gpu")
    check = os.path.join(BUILD, "ipx_check_gpu")
    if not (os.path.exists(afiro) and os.path.exists(check)):
        pytest.fail("drop-in executables missing: run __graft_entry__.build() where the "
                    "reference tree is available")
    out = subprocess.run([afiro], capture_output=True, text=True, timeout=300)
    assert out.returncode == 0, out.stdout + out.stderr
    assert "-4.64753143e+02" in out.stdout
    assert "Status interior point solve:                        optimal" in out.stdout
    out = subprocess.run([check], capture_output=True, text=True, timeout=600)
    assert out.returncode == 0, out.stdout[-2000:] + out.stderr[-2000:]
    assert "All tests passed" in out.stdout
