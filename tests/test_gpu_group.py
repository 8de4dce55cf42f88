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


@pytest.mark.parametrize("lp,extra", [
    ("random:20000:200000:8", ("--stop-at-switch", "-1", "--crossover", "0")),
    ("random:500:5000:10", ()),          # both phases and crossover: the basis phase is single-GPU
    ("transport:15:40", ()),
])
def test_lp_solver_on_two_gpus_matches_one(lp, extra, tmp_path):
    if _gpus() < 2:
        pytest.skip("needs 2 GPUs")
    one = _solve(lp, 1, tmp_path, *extra)
    two = _solve(lp, 2, tmp_path, *extra)
    assert one["status"] == two["status"] and one["status_ipm"] == two["status_ipm"]
    assert one["status_crossover"] == two["status_crossover"]
    assert abs(one["iter"] - two["iter"]) <= 1
    # the sharded solve sums the ranks' partial products in rank order: rounding-level
    # differences only, so the early iterations agree line by line
    for a, b in zip(one["per_iter"][:4], two["per_iter"][:4]):
        assert a["kktiter"] == b["kktiter"] and a["mu"] == b["mu"], (a, b)
    if one["status"] == 1000:
        assert abs(one["objval"] - two["objval"]) <= 1e-9 * max(1.0, abs(one["objval"]))
    assert two["kktiter1"] > 0 and two["time_cr1"] > 0
