"""Several GPUs behind LpSolver: IPXGPU_NGPUS = G makes KKTSolverDiag of the drop-in build run
on a one-process group of G column-sharded contexts (ipxgpu_create_group, gpu_bridge.cc). The
environment variable is read when the library is loaded, so every arm is a process of its own
(tools/solve_lp.py). Needs 2 GPUs; skipped otherwise."""

import json
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _gpus():
    from ipx_b200 import capi
    return capi.device_count()


def _solve(lp, ngpus, tmp_path, *extra):
    out = tmp_path / f"g{ngpus}.json"
    env = dict(os.environ, IPXGPU_NGPUS=str(ngpus))
    r = subprocess.run([sys.executable, os.path.join(REPO, "tools", "solve_lp.py"), lp, "--impl",
                        "gpu", "--per-iter", "--out", str(out), *extra],
                       env=env, capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    with open(out) as f:
        return json.load(f)["results"]["gpu"]


def _band(lp):
    """The reference's own runs on this LP: generator order, other column orders, one-ulp copies
    (tests/golden/make_group_golden.py)."""
    path = os.path.join(REPO, "tests", "golden", "e2e_group_bands.json")
    if not os.path.exists(path):
        pytest.skip("tests/golden/e2e_group_bands.json missing (tests/golden/make_group_golden.py)")
    with open(path) as f:
        return json.load(f)[lp]["runs"]


@pytest.mark.parametrize("lp,extra", [
    ("random:20000:200000:8", ("--stop-at-switch", "-1", "--crossover", "0")),
    ("random:500:5000:10", ()),          # both phases and crossover: the basis phase is single-GPU
    ("transport:15:40", ()),
])
def test_lp_solver_on_two_gpus_matches_one(lp, extra, tmp_path):
    if _gpus() < 2:
        pytest.skip("needs 2 GPUs")
    ref = _band(lp)
    one = _solve(lp, 1, tmp_path, *extra)
    two = _solve(lp, 2, tmp_path, *extra)
    # The bar is the reference, not the other arm: IPM iterations within +-1 of what the
    # reference's own runs span (the same LP with its columns in another order, or with one-ulp
    # changes, takes 13 or 14 iterations at 500 x 5000: a Newton solve there ends within rounding
    # of its tolerance, and which side it lands on decides the path), same status, objective to
    # 1e-9 of the reference's.
    iters = [r["iter"] for r in ref]
    for arm in (one, two):
        assert arm["status"] == ref[0]["status"] and arm["status_ipm"] == ref[0]["status_ipm"]
        assert arm["status_crossover"] == ref[0]["status_crossover"]
        assert min(iters) - 1 <= arm["iter"] <= max(iters) + 1, (arm["iter"], iters)
        if ref[0]["status"] == 1000:
            assert abs(arm["objval"] - ref[0]["objval"]) <= 1e-9 * max(1.0, abs(ref[0]["objval"]))
    # the sharded solve sums the ranks' partial products in rank order: rounding-level
    # differences only, so the early iterations agree line by line - with each other and with
    # the reference's CR counts
    for k, (a, b) in enumerate(zip(one["per_iter"][:4], two["per_iter"][:4])):
        assert a["kktiter"] == b["kktiter"] and a["mu"] == b["mu"], (a, b)
        assert a["kktiter"] == ref[0]["kktiter_per_iter"][k], (k, a, ref[0]["kktiter_per_iter"])
    assert two["kktiter1"] > 0 and two["time_cr1"] > 0
