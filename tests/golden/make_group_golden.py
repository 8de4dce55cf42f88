"""Pins the band the REFERENCE's own end-to-end runs span on the LPs of tests/test_gpu_group.py:
the LP as generated, the same LP with its columns in other orders (every sum over columns inside
the reference is then taken in another order - the kind of difference a column-sharded device arm
has) and copies whose rhs / objective differ by one ulp. From oracle/_ref/libipx_ref.so (the
unmodified reference compiled by oracle/Makefile); a few seconds to a few minutes of one core.

    python tests/golden/make_group_golden.py

random:500:5000:10 (both phases, crossover): 24 reference runs take 13 or 14 IPM iterations; the
CR count of iteration 7 is 66 in 22 of them and 65 in two, after which the paths differ - a
device arm that lands on the other side of that threshold is as far from the generator-order
reference run as the reference is from itself.
"""

import json
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))

from ipx_b200 import e2e, lpgen  # noqa: E402
from oracle import ipxlib  # noqa: E402

# (lp spec as tools/solve_lp.py takes it, IPX parameters, column orders, ulp copies)
CASES = [
    ("random:20000:200000:8", dict(dualize=0, crossover=0, stop_at_switch=-1), 3, 2),
    ("random:500:5000:10", dict(dualize=0), 15, 8),
    ("transport:15:40", dict(dualize=0), 6, 4),
]


def make_lp(spec):
    kind, *dims = spec.split(":")
    d = [int(v) for v in dims]
    if kind == "random":
        return lpgen.random_sparse_lp(d[0], d[1], d[2], 1002)
    return lpgen.transportation_lp(d[0], d[1], 1004)


def permuted(lp, perm):
    """The same LP with column perm[k] at position k."""
    cnt = np.diff(lp.Ap)[perm]
    Ap = np.concatenate([[0], np.cumsum(cnt)]).astype(np.int64)
    src = np.repeat(lp.Ap[:-1][perm], cnt) + np.arange(int(Ap[-1])) - np.repeat(Ap[:-1], cnt)
    q = lpgen.LP(**{**lp.__dict__})
    q.Ap, q.Ai, q.Ax = Ap, lp.Ai[src], lp.Ax[src]
    q.obj, q.lb, q.ub = lp.obj[perm], lp.lb[perm], lp.ub[perm]
    return q


if __name__ == "__main__":
    ref = ipxlib.IpxLibrary(ipxlib.REF_LIB)
    out = {}
    for spec, params, norders, nulp in CASES:
        lp = make_lp(spec)
        rng = np.random.default_rng(5)
        runs = []

        def run(q, what):
            r = e2e.solve(ref, q, per_iter=True, **params)
            runs.append({"what": what, "status": r["status"], "status_ipm": r["status_ipm"],
                         "status_crossover": r["status_crossover"], "iter": r["iter"],
                         "kktiter1": r["kktiter1"], "kktiter2": r["kktiter2"],
                         "objval": r["objval"], "pobjval": r["pobjval"],
                         "kktiter_per_iter": [row["kktiter"] for row in r["per_iter"]]})
            print(spec, what, r["status"], r["iter"], r["kktiter1"], r["kktiter2"], flush=True)

        run(lp, "generator order")
        for c in range(norders):
            run(permuted(lp, rng.permutation(lp.n)), f"column order {c}")
        ulp = 2.0 ** -52
        for c in range(nulp):
            q = lpgen.LP(**{**lp.__dict__})
            q.rhs = lp.rhs * (1.0 + ulp * rng.choice(np.array([-1.0, 0.0, 1.0]), lp.m))
            q.obj = lp.obj * (1.0 + ulp * rng.choice(np.array([-1.0, 0.0, 1.0]), lp.n))
            run(q, f"ulp copy {c}")
        out[spec] = {"params": params, "runs": runs}
    with open(os.path.join(HERE, "e2e_group_bands.json"), "w") as f:
        json.dump(out, f, indent=1)
