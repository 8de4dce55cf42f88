"""Parity at BASELINE.json's FULL sizes (configs 2, 4 and 5) on one GPU.

The oracle's apply takes 0.1-3 s at these sizes, so the apply, the diagonal build
and the first 20 PCR iterations are compared with it directly (apply: 1e-12
norm-wise relative, BASELINE.json north_star); on top of that the size-independent
properties of the operator are asserted: symmetry x'Cz = z'Cx, linearity,
positive definiteness, run-to-run determinism, and the residual bound of a solve
to tolerance. The checks live in tools/fullsize.py (the same script fills
DESIGN.md's full-size table).
"""

import os
import sys

import pytest

pytestmark = pytest.mark.gpu

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(REPO, "tools"))


@pytest.fixture(scope="module")
def capi():
    from ipx_b200 import capi as c
    c.load()
    assert c.device_count() >= 1, "no CUDA device"
    return c


@pytest.mark.parametrize("config", ["C2", "C4", "C5"])
def test_full_size_config(config, capi, oracle):
    import fullsize
    lp = fullsize.make_lp(config)
    out = fullsize.run_operator_checks(config, lp, capi, oracle, pcr_iters=20, reps=3,
                                       log=lambda s: None)
    assert out["matvecs_per_s"] > out["oracle_matvecs_per_s"]


def test_basis_path_quarter_scale_config3(reflib, gpulib):
    """Config 3 (block-angular, KKTSolverBasis path) at a quarter of its size - 50,000 rows x
    500,000 columns: reference build against the drop-in build on the same basis. Split
    operator apply <= 1e-12 (or the operator's own one-ulp sensitivity), CR on it, and a full KKTSolverBasis Factorize + Solve
    (right-hand side sweeps, SolveDense steps, CR, recovery on the device)."""
    import fullsize
    lp = fullsize.make_lp("C3", 0.25)
    out = fullsize.run_basis_checks("C3", lp, reflib, gpulib, log=lambda s: None,
                                    kkt_maxiter=100, volume_tol=1e6)
    assert out["split_apply_rel_err"] <= max(1e-12, 4.0 * out["split_apply_ulp_sensitivity"])
    assert out["kkt_solve_gpu"]["time_cr2"] < out["kkt_solve_ref"]["time_cr2"]
