"""End-to-end LP solves through the unchanged ipx_c.h API with the reference CPU
build and with the GPU drop-in build; prints the counters and phase times of
ipx_info side by side (solver level)."""
import argparse
import json
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from ipx_b200 import e2e, lpgen
from oracle import ipxlib  # noqa: E402



def make(name):
    kind, *dims = name.split(":")
    d = [int(v) for v in dims]
    if kind == "random":
        return lpgen.random_sparse_lp(d[0], d[1], d[2], 1002)
    if kind == "block":
        return lpgen.block_angular_lp(d[0], d[1], d[2], 1003)
    if kind == "transport":
        return lpgen.transportation_lp(d[0], d[1], 1004)
    if kind == "afiro":
        return lpgen.afiro_lp()
    raise ValueError(name)


ap = argparse.ArgumentParser()
ap.add_argument("lp", help="random:m:n:k | block:m:n:k | transport:S:T | afiro")
ap.add_argument("--impl", default="both", choices=["both", "ref", "gpu"])
ap.add_argument("--crossover", type=int, default=1)
ap.add_argument("--switchiter", type=int, default=-1)
ap.add_argument("--maxiter", type=int, default=300)
ap.add_argument("--stop-at-switch", type=int, default=0,
                help="-1: stop after the diagonal-preconditioned phase (reference debug parameter)")
ap.add_argument("--debug", type=int, default=0, help="ipx debug level (1: per-iteration kktiter/step sizes in the log)")
ap.add_argument("--out", default=None)
ap.add_argument("--per-iter", action="store_true",
                help="log each arm with debug >= 1 and print the CR iterations of every IPM "
                     "iteration side by side (where the arms part ways, and by how much)")
args = ap.parse_args()
if args.per_iter:
    args.debug = max(args.debug, 1)


# Libraries first, as an application would link them: the drop-in build starts creating its
# CUDA context when it is loaded (IPXGPU_EAGER_INIT=0 turns that off).
libs = {impl: ipxlib.IpxLibrary(path)
        for impl, path in (("ref", ipxlib.REF_LIB), ("gpu", ipxlib.GPU_LIB))
        if args.impl in ("both", impl)}
lp = make(args.lp)
print(f"LP {lp.name}: m={lp.m} n={lp.n} nnz={lp.nnz} optimum={lp.optimum}", flush=True)
results = {}
for impl, path in (("ref", ipxlib.REF_LIB), ("gpu", ipxlib.GPU_LIB)):
    if args.impl not in ("both", impl):
        continue
    results[impl] = e2e.solve(libs[impl], lp, per_iter=args.per_iter,
                              display=int(os.environ.get("IPX_DISPLAY", "0")), dualize=0,
                              crossover=args.crossover, switchiter=args.switchiter,
                              ipm_maxiter=args.maxiter, stop_at_switch=args.stop_at_switch,
                              debug=args.debug)
    print(impl, json.dumps({k: v for k, v in results[impl].items() if k != "per_iter"}), flush=True)
if len(results) == 2:
    r, g = results["ref"], results["gpu"]
    print("objective diff (rel):", abs(r["objval"] - g["objval"]) / max(1.0, abs(r["objval"])))
    print("speed-up total:", r["time_total"] / g["time_total"],
          " CR1:", r["time_cr1"] / max(g["time_cr1"], 1e-12),
          " CR2:", r["time_cr2"] / max(g["time_cr2"], 1e-12) if g["time_cr2"] > 0 else None)
if args.per_iter and len(results) == 2:
    # KKTSolver::iter() restarts at every Factorize: the column is the CR iterations of the
    # iteration's two Newton solves (predictor + corrector).
    def increments(rows):
        return [r["kktiter"] for r in rows]
    ra, ga = results["ref"]["per_iter"], results["gpu"]["per_iter"]
    di, dg = increments(ra), increments(ga)
    print(" iter |      mu ref      mu gpu | CR ref CR gpu | pres ref  pres gpu")
    for k in range(max(len(ra), len(ga))):
        a = ra[k] if k < len(ra) else None
        g = ga[k] if k < len(ga) else None
        print(" %4d | %11s %11s | %6s %6s | %9s %9s" % (
            k, "%.3e" % a["mu"] if a else "-", "%.3e" % g["mu"] if g else "-",
            di[k] if a else "-", dg[k] if g else "-",
            "%.2e" % a["pres"] if a else "-", "%.2e" % g["pres"] if g else "-"))
if args.out:
    with open(args.out, "w") as f:
        json.dump({"lp": args.lp, "m": lp.m, "n": lp.n, "nnz": lp.nnz, "results": results}, f,
                  indent=1)
