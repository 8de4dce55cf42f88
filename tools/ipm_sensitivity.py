"""How far do one-ulp perturbations of the LP data move the REFERENCE's own IPM run? Runs the
reference build (oracle/_ref) on an LP and on copies whose right-hand side and objective are
multiplied entrywise by (1 +- 2^-52), and prints the per-iteration CR counts side by side.
The drop-in build differs from the reference by summation order only, i.e. by perturbations of
this size inside every operator apply; this tool measures what such perturbations do to the
iteration path of the unmodified CPU code.

    python tools/ipm_sensitivity.py random:50000:500000:10 --copies 4 [--gpu]
"""
import argparse
import json
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from ipx_b200 import e2e, lpgen
from oracle import ipxlib  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("lp")
ap.add_argument("--copies", type=int, default=4)
ap.add_argument("--mode", default="ulp", choices=["ulp", "permute"],
                help="ulp: rhs and obj times (1 +- 2^-52); permute: the SAME LP with its columns in "
                     "a random order - every sum over columns inside the reference (A*W*A' applies, "
                     "right-hand sides, recoveries) is then taken in another order, which is the "
                     "kind of difference a device arm has in every operator apply")
ap.add_argument("--seed", type=int, default=5)
ap.add_argument("--gpu", action="store_true", help="also run the drop-in build on the unperturbed LP")
ap.add_argument("--out", default=None)
ap.add_argument("--golden", default=None, help="write the compact fixture the tests read")
args = ap.parse_args()
kind, *dims = args.lp.split(":")
d = [int(v) for v in dims]
lp = (lpgen.random_sparse_lp(d[0], d[1], d[2], 1002) if kind == "random"
      else lpgen.transportation_lp(d[0], d[1], 1004))
params = dict(dualize=0, crossover=0, stop_at_switch=-1)
ref = ipxlib.IpxLibrary(ipxlib.REF_LIB)
runs = {}
runs["ref"] = e2e.solve(ref, lp, per_iter=True, **params)
rng = np.random.default_rng(args.seed)
def permuted(lp, perm):
    """The same LP with column perm[k] at position k."""
    cnt = np.diff(lp.Ap)[perm]
    Ap = np.concatenate([[0], np.cumsum(cnt)]).astype(np.int64)
    src = (np.repeat(lp.Ap[:-1][perm], cnt) + np.arange(int(Ap[-1])) - np.repeat(Ap[:-1], cnt))
    q = lpgen.LP(**{**lp.__dict__})
    q.Ap, q.Ai, q.Ax = Ap, lp.Ai[src], lp.Ax[src]
    q.obj, q.lb, q.ub = lp.obj[perm], lp.lb[perm], lp.ub[perm]
    return q


for c in range(args.copies):
    if args.mode == "permute":
        q = permuted(lp, rng.permutation(lp.n))
        runs[f"ref+perm{c}"] = e2e.solve(ref, q, per_iter=True, **params)
        continue
    q = lpgen.LP(**{**lp.__dict__})
    ulp = 2.0 ** -52
    q.rhs = lp.rhs * (1.0 + ulp * rng.choice(np.array([-1.0, 0.0, 1.0]), lp.m))
    q.obj = lp.obj * (1.0 + ulp * rng.choice(np.array([-1.0, 0.0, 1.0]), lp.n))
    runs[f"ref+ulp{c}"] = e2e.solve(ref, q, per_iter=True, **params)
if args.gpu:
    runs["gpu"] = e2e.solve(ipxlib.IpxLibrary(ipxlib.GPU_LIB), lp, per_iter=True, **params)
names = list(runs)
print("LP", lp.name)
print("%-10s" % "", " ".join("%10s" % n for n in names))
print("%-10s" % "IPM iter", " ".join("%10d" % runs[n]["iter"] for n in names))
print("%-10s" % "CR iter", " ".join("%10d" % runs[n]["kktiter1"] for n in names))
print("%-10s" % "pobjval", " ".join("%10.6g" % runs[n]["pobjval"] for n in names))
depth = max(len(runs[n]["per_iter"]) for n in names)
for k in range(depth):
    cells = []
    for n in names:
        t = runs[n]["per_iter"]
        cells.append("%4d %5.0e" % (t[k]["kktiter"], t[k]["mu"]) if k < len(t) else "         -")
    print("%-10s" % f"it {k}", " ".join("%10s" % c for c in cells))
if args.out:
    with open(args.out, "w") as f:
        json.dump({"lp": lp.name, "runs": runs}, f, indent=1)
if args.golden:
    # compact form for tests/golden: the reference runs only, and the number of leading
    # iterations on which all of them print the same table row
    refs = [runs[n] for n in names if n.startswith("ref")]
    stable = 0
    key = lambda row: row["kktiter"]  # the printed residuals differ in the 3rd digit by iteration 2
    while all(len(r["per_iter"]) > stable for r in refs) and all(
            key(r["per_iter"][stable]) == key(refs[0]["per_iter"][stable]) for r in refs):
        stable += 1
    with open(args.golden, "w") as f:
        json.dump({"lp": lp.name, "perturbation": ("the same LP with its columns in random orders; first run "
                                    "in the generator's order" if args.mode == "permute" else
                                    "rhs and obj entrywise times (1 + e * 2^-52), e in {-1, 0, 1}; "
                                    "first run unperturbed"), "stable_iterations": stable,
                   "stable_iterations_meaning": "leading IPM iterations on which all runs need the "
                   "same number of CR iterations",
                   "runs": [{k: r[k] for k in ("status", "status_ipm", "iter", "kktiter1", "pobjval",
                                                "dobjval", "per_iter")} for r in refs]}, f, indent=1)
