"""BASELINE.json configs at their FULL sizes on one GPU: parity against the oracle
(the oracle's apply is 0.1-3 s at these sizes), size-independent properties of the
operator, and timings of every stage (layout build, apply, diagonal build, PCR).

    python tools/fullsize.py C2 C4 C5 [--basis C3] [--out gpurun_out/fullsize.json]

Used by tests/test_gpu_fullsize.py (same checks, asserted) and to fill the
full-size table of DESIGN.md. The oracle is the checker only.
"""

import argparse
import json
import os
import sys
import time

import numpy as np

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if REPO not in sys.path:
    sys.path.insert(0, REPO)

from ipx_b200 import lpgen  # noqa: E402

APPLY_TOL = 1e-12


def make_lp(name, scale=1.0):
    """SURVEY.md section 8d recipes; `scale` < 1 shrinks the shape (CPU smoke of this script)."""
    s = lambda v: max(8, int(v * scale))
    if name == "C2":
        return lpgen.random_sparse_lp(s(100_000), s(1_000_000), 10, 1002)
    if name == "C3":
        return lpgen.block_angular_lp(s(200_000), s(2_000_000), 10, 1003)
    if name == "C4":
        return lpgen.transportation_lp(s(2000), s(5000), 1004)
    if name == "C5":
        return lpgen.random_sparse_lp(s(1_000_000), s(20_000_000), 5, 1005)
    raise ValueError(name)


def algorithmic_bytes(m, n, nnz):
    """bytes of one A*D^2*A' apply in the dual-layout design."""
    return 2 * nnz * 12 + 4 * (n + 1) + 4 * (m + 1) + 8 * (n + m) + 8 * m + 8 * m


def rel_err(a, b):
    d = np.abs(b).max()
    return float(np.abs(a - b).max() / d) if d > 0 else float(np.abs(a - b).max())


def run_operator_checks(name, lp, capi, oracle, pcr_iters=20, reps=10, log=print):
    """Returns a dict of measurements; raises AssertionError on a parity failure."""
    m, n = lp.m, lp.n
    out = {"config": name, "m": m, "n": n, "nnz": int(lp.nnz)}
    AIp, AIi, AIx = lp.solver_form()
    t0 = time.time()
    ctx = capi.Context(m, n, AIp, AIi, AIx)
    out["context_create_s"] = time.time() - t0
    tiling = ctx.tiling()
    out["tiling"] = tiling
    out["path"] = ("band" if tiling["sweep1"]["enabled"] and tiling["sweep2"]["enabled"]
                   else "generic")
    A = oracle.Csc(AIp, AIi, AIx)
    rng = np.random.default_rng(5)
    x = rng.standard_normal(m)
    z = rng.standard_normal(m)

    # parity of the apply in every weight regime, W == NULL included
    errs = {}
    for regime in ("ones", "mid", "wide", "null"):
        W = None if regime == "null" else lpgen.weights(n + m, regime, 6)
        ctx.normal_prepare(W)
        y, dot = ctx.normal_apply(x)
        t0 = time.time()
        y0, dot0 = oracle.normal_apply(m, n, A, W, x)
        out["oracle_apply_s"] = time.time() - t0
        errs[regime] = rel_err(y, y0)
        assert errs[regime] <= APPLY_TOL, (name, regime, errs[regime])
        assert abs(dot - dot0) <= APPLY_TOL * np.abs(x * y0).sum(), (name, regime, dot, dot0)
        y2, dot2 = ctx.normal_apply(x)
        assert np.array_equal(y, y2) and dot == dot2, "apply is not run-to-run deterministic"
    out["apply_rel_err"] = errs

    # size-independent properties (W = mid): symmetry, linearity, positive definiteness
    W = lpgen.weights(n + m, "mid", 8)
    ctx.normal_prepare(W)
    Cx, xCx = ctx.normal_apply(x)
    Cz, zCz = ctx.normal_apply(z)
    sym = abs(z @ Cx - x @ Cz) / (np.abs(z * Cx).sum() + 1e-300)
    Cxz, _ = ctx.normal_apply(2.0 * x - 3.0 * z)
    lin = rel_err(Cxz, 2.0 * Cx - 3.0 * Cz)
    out["symmetry_rel"] = float(sym)
    out["linearity_rel"] = lin
    assert sym <= 1e-12 and lin <= 1e-12, (sym, lin)
    assert xCx > 0 and zCz > 0

    # diagonal build against the oracle
    ctx.diag_factorize(None, use_prepared=True)
    d = ctx.diag_get()
    d0 = oracle.diag_build(m, n, A, W)
    out["diag_rel_err"] = rel_err(d, d0)
    assert out["diag_rel_err"] <= APPLY_TOL

    # fixed-count PCR against the oracle (first iterations: identical histories)
    rhs = rng.standard_normal(m)
    zg, info = ctx.pcr_solve(rhs, 1e-30, None, pcr_iters)
    op = oracle.normal_operator(m, n, A, W)
    t0 = time.time()
    zo, info0 = oracle.pcr_solve(op, m, d0, rhs, 1e-30, None, pcr_iters)
    out["oracle_pcr_s"] = time.time() - t0
    out["pcr_rel_err"] = rel_err(zg, zo)
    out["pcr_errflag"] = (int(info["errflag"]), int(info0["errflag"]))
    out["pcr_iter"] = (int(info["iter"]), int(info0["iter"]))
    assert info["errflag"] == info0["errflag"] and info["iter"] == info0["iter"], (info, info0)
    assert out["pcr_rel_err"] <= 1e-8, out["pcr_rel_err"]

    # solve to a tolerance: the residual bound is the size-independent check
    t0 = time.time()
    zs, infos = ctx.pcr_solve(rhs, 1e-8, None, -1)
    out["pcr_solve_s"] = time.time() - t0
    out["pcr_solve_iter"] = int(infos["iter"])
    Czs, _ = ctx.normal_apply(zs)
    out["pcr_solve_resid"] = float(np.abs(rhs - Czs).max())
    assert infos["errflag"] == 0 and out["pcr_solve_resid"] <= 1e-8 * (1 + 1e-6) + 1e-12

    # timings
    t = ctx.time_normal_apply(reps, True)
    out["apply_us_isolated_l2_flushed"] = 1e3 * t["apply_ms"]
    out["sweep1_us"] = 1e3 * t["sweep1_ms"]
    out["sweep2_us"] = 1e3 * t["sweep2_ms"]
    nbytes = algorithmic_bytes(m, n, lp.nnz)
    out["algorithmic_bytes"] = nbytes
    out["apply_gbs_isolated"] = nbytes / (t["apply_ms"] * 1e-3) / 1e9
    # in-loop: a fixed number of CR iterations, device resident
    ctx.pcr_solve(rhs, 1e-30, None, 5)
    t0 = time.time()
    _, inf = ctx.pcr_solve(rhs, 1e-30, None, 50)
    wall = time.time() - t0
    out["pcr50_wall_s"] = wall
    out["pcr50_time_s"] = float(inf["time"])
    out["pcr50_time_AAt_s"] = float(inf["time_op"]) if "time_op" in inf else None
    out["cr_iter_us"] = 1e6 * wall / max(1, int(inf["iter"]))
    out["matvecs_per_s"] = (int(inf["iter"]) + 1) / wall
    out["oracle_matvecs_per_s"] = 1.0 / max(out["oracle_apply_s"], 1e-9)
    ctx.close()
    log(json.dumps(out))
    return out


def run_basis_checks(name, lp, reflib, gpulib, log=print, kkt_maxiter=200, volume_tol=2.0):
    """KKTSolverBasis path (config 3): reference build against the drop-in build on the same
    basis (same host LU provider), split operator apply, CR, and a full KKT solve."""
    out = {"config": name + "-basis", "m": lp.m, "n": lp.n, "nnz": int(lp.nnz)}
    t0 = time.time()
    # volume_tol: Maxvolume's threshold (reference src/maxvolume.cc:209); a large value keeps the
    # host-side basis updates of KKTSolverBasis::Factorize out of a full-size timing run
    params = dict(dualize=0, volume_tol=volume_tol)
    ref, gpu = reflib.model(lp, **params), gpulib.model(lp, **params)
    out["model_s"] = time.time() - t0
    m, n = ref.m, ref.n
    rng = np.random.default_rng(11)
    # Scaling factors as in round 1's runs (the same basis, so the numbers compare), and an
    # interior iterate that HAS these scaling factors (reference src/iterate.cc:183-198:
    # 1/sqrt(zl/xl + zu/xu)), so that KKTSolverBasis::Factorize below works on the basis these
    # weights chose and Maxvolume has little left to improve.
    nm = n + m
    colscale = np.exp(rng.uniform(-3, 3, nm))
    _, _, lb, ub = ref.model_vectors()
    has_lb, has_ub = np.isfinite(lb), np.isfinite(ub)
    share = np.where(has_lb & has_ub, 0.5, 1.0)
    it = (rng.uniform(0.5, 1.5, nm), np.where(has_lb, colscale, np.inf),
          np.where(has_ub, colscale, np.inf), rng.standard_normal(m),
          np.where(has_lb, share / colscale, 0.0), np.where(has_ub, share / colscale, 0.0))
    for key, mdl in (("ref", ref), ("gpu", gpu)):
        t0 = time.time()
        mdl.basis_from_weights(colscale)
        out[f"basis_{key}_s"] = time.time() - t0
    b0, s0 = ref.basis_get()
    b1, s1 = gpu.basis_get()
    assert np.array_equal(b0, b1) and np.array_equal(s0, s1)
    for key, mdl in (("ref", ref), ("gpu", gpu)):
        t0 = time.time()
        mdl.split_prepare(colscale)
        out[f"split_prepare_{key}_s"] = time.time() - t0
    x = rng.standard_normal(m)
    y0, d0 = ref.split_apply(x)
    y1, d1 = gpu.split_apply(x)
    out["split_apply_rel_err"] = rel_err(y1, y0)
    # How far one-ulp perturbations of the input move the REFERENCE's own result: C = I +
    # inverse(B) N N' inverse(B') amplifies rounding by the conditioning of the basis (a crash
    # basis here; |Cx| / |x| reaches 1e14 at 50,000 rows). The triangular solves are bit-identical
    # on both arms, so the arms differ by the summation order inside N N' only, amplified the
    # same way. The 1e-12 bar applies wherever the operator itself is that well determined.
    ulp = np.where(rng.random(m) < 0.5, -1.0, 1.0) * 2.0 ** -52
    y0p, _ = ref.split_apply(x * (1.0 + ulp))
    out["split_apply_ulp_sensitivity"] = rel_err(y0p, y0)
    out["split_apply_growth"] = float(np.abs(y0).max() / np.abs(x).max())
    tol = max(APPLY_TOL, 4.0 * out["split_apply_ulp_sensitivity"])
    assert out["split_apply_rel_err"] <= tol, (out["split_apply_rel_err"], tol)
    assert abs(d1 - d0) <= tol * np.abs(x * y0).sum()
    for key, mdl, reps in (("ref", ref, 3), ("gpu", gpu, 20)):
        t0 = time.time()
        _, parts = mdl.split_apply_timed(x, reps)
        out[f"split_apply_{key}_ms"] = 1e3 * (time.time() - t0) / reps
        out[f"split_apply_{key}_parts_ms"] = {k: 1e3 * v for k, v in parts.items()}
    rhs = rng.standard_normal(m)
    t0 = time.time()
    z0, i0 = ref.cr_solve_split(rhs, 1e-8, 200)
    out["cr_ref_s"] = time.time() - t0
    t0 = time.time()
    z1, i1 = gpu.cr_solve_split(rhs, 1e-8, 200)
    out["cr_gpu_s"] = time.time() - t0
    out["cr_iter"] = (int(i0["iter"]), int(i1["iter"]))
    out["cr_errflag"] = (int(i0["errflag"]), int(i1["errflag"]))
    assert i0["errflag"] == i1["errflag"]
    if i0["errflag"] == 0:
        assert abs(i0["iter"] - i1["iter"]) <= max(1, i0["iter"] // 20)
        out["cr_rel_err"] = rel_err(z1, z0)
        assert out["cr_rel_err"] <= 1e-5
    # KKTSolverBasis: Factorize from an interior iterate (scaling factors, Maxvolume, LU,
    # Prepare) and one full Solve per arm - right-hand side sweeps, SolveDense steps, CR, recovery
    # (reference src/kkt_solver_basis.cc:20-194; on the device in the drop-in build).
    a, b = rng.standard_normal(nm), rng.standard_normal(m)
    sol = {}
    for key, mdl in (("ref", ref), ("gpu", gpu)):
        mdl.iterate_set(*it)
        mdl.kktbasis_maxiter(kkt_maxiter)
        t0 = time.time()
        f = mdl.kktbasis_factorize()
        out[f"kkt_factorize_{key}_s"] = time.time() - t0
        assert f["err"] == 0, f
        t0 = time.time()
        x, y, info = mdl.kktbasis_solve(a, b, 1e-6)
        out[f"kkt_solve_{key}_s"] = time.time() - t0
        out[f"kkt_solve_{key}"] = {k: info[k] for k in ("err", "kktiter2", "time_cr2", "time_cr2_NNt",
                                                        "time_cr2_B", "time_cr2_Bt")}
        sol[key] = (x, y, info)
    # The same solve cut off after a few CR iterations, while the two arms' Krylov iterates
    # have not parted ways yet: what the device does around the CR loop (right-hand side
    # sweeps, the Basis::SolveDense steps, the recovery sweeps) against the reference's.
    short = {}
    for key, mdl in (("ref", ref), ("gpu", gpu)):
        mdl.kktbasis_maxiter(5)
        short[key] = mdl.kktbasis_solve(a, b, 1e-6)
        mdl.kktbasis_maxiter(kkt_maxiter)
    out["kkt5_x_rel_err"] = rel_err(short["gpu"][0], short["ref"][0])
    out["kkt5_y_rel_err"] = rel_err(short["gpu"][1], short["ref"][1])
    out["kkt5_iter"] = (int(short["ref"][2]["kktiter2"]), int(short["gpu"][2]["kktiter2"]))
    (x0, y0, i0), (x1, y1, i1) = sol["ref"], sol["gpu"]
    assert i0["err"] == i1["err"], (i0, i1)
    if i0["err"] == 0:
        assert abs(i0["kktiter2"] - i1["kktiter2"]) <= max(1, int(i0["kktiter2"]) // 20)
        out["kkt_x_rel_err"] = rel_err(x1, x0)
        out["kkt_y_rel_err"] = rel_err(y1, y0)
        assert out["kkt_x_rel_err"] <= 1e-4 and out["kkt_y_rel_err"] <= 1e-4
    else:
        # iteration limit on both arms
        assert i0["kktiter2"] == i1["kktiter2"]
        out["kkt_x_rel_err"] = rel_err(x1, x0)
        out["kkt_y_rel_err"] = rel_err(y1, y0)
    # What the KKT solver promises (reference src/kkt_solver.h:20-39), measured on both arms:
    # the second block row AI x = b holds to rounding, the first block row a - G x - AI'y is
    # what the CR tolerance (or its iteration limit) leaves. CR iterates of an ill-conditioned
    # system part ways after a few dozen iterations, the residuals they reach do not.
    import scipy.sparse as sp
    AIp, AIi, AIx = ref.AI()
    AI = sp.csc_matrix((AIx, AIi, AIp), shape=(m, nm))
    gdiag = 1.0 / colscale ** 2
    for key, (xs, ys, _) in sol.items():
        out[f"kkt_res_primal_{key}"] = float(np.abs(AI @ xs - b).max() / (1.0 + np.abs(b).max()))
        out[f"kkt_res_dual_{key}"] = float(np.abs(a - gdiag * xs - AI.T @ ys).max() /
                                           (1.0 + np.abs(a).max()))
    assert out["kkt_res_primal_gpu"] <= max(1e-9, 10.0 * out["kkt_res_primal_ref"]), out
    assert out["kkt_res_dual_gpu"] <= max(1e-9, 10.0 * out["kkt_res_dual_ref"]), out
    ref.close()
    gpu.close()
    log(json.dumps(out))
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("configs", nargs="*", default=[])
    ap.add_argument("--basis", nargs="*", default=[])
    ap.add_argument("--scale", type=float, default=1.0)
    ap.add_argument("--volume-tol", type=float, default=2.0)
    ap.add_argument("--kkt-maxiter", type=int, default=200)
    ap.add_argument("--out", default="")
    args = ap.parse_args()
    if not args.configs and not args.basis:
        args.configs = ["C2"]
    from ipx_b200 import capi
    from oracle import ipxlib
    from oracle import pyoracle
    capi.load()
    pyoracle.lib()
    results = []
    for name in args.configs:
        t0 = time.time()
        lp = make_lp(name, args.scale)
        print(f"# {name}: generated {lp.m} x {lp.n}, {lp.nnz} nnz in {time.time() - t0:.1f} s",
              flush=True)
        results.append(run_operator_checks(name, lp, capi, pyoracle))
        del lp
    for name in args.basis:
        lp = make_lp(name, args.scale)
        results.append(run_basis_checks(name, lp, ipxlib.IpxLibrary(ipxlib.REF_LIB),
                                        ipxlib.IpxLibrary(ipxlib.GPU_LIB),
                                        kkt_maxiter=args.kkt_maxiter, volume_tol=args.volume_tol))
    if args.out:
        with open(args.out, "w") as f:
            json.dump(results, f, indent=1)


if __name__ == "__main__":
    main()
