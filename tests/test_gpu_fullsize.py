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
