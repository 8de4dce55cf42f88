"""Profiling driver: a few device-resident A*W*A' applies (and optionally a short
PCR solve) on the BASELINE configs[1] shape. Used under ncu; prints timings when
run plainly."""
import argparse
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from ipx_b200 import capi, lpgen  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--rows", type=int, default=100_000)
ap.add_argument("--cols", type=int, default=1_000_000)
ap.add_argument("--k", type=int, default=10)
ap.add_argument("--reps", type=int, default=5)
ap.add_argument("--pcr", type=int, default=0, help="also run a PCR solve of this many iterations")
ap.add_argument("--transport", action="store_true")
ap.add_argument("--madd", type=int, default=0,
                help="also run ipx::MultiplyAdd on the resident AI this many times per direction "
                     "(ipxgpu_multiply_add, host vectors) and check it against the oracle")
args = ap.parse_args()

lp = lpgen.transportation_lp(2000, 5000, 1004) if args.transport else \
    lpgen.random_sparse_lp(args.rows, args.cols, args.k, 1002)
m, n = lp.m, lp.n
import time
_t0 = time.time()
ctx = capi.Context(m, n, *lp.solver_form())
print(f"context creation {time.time() - _t0:.2f} s")
W = lpgen.weights(n + m, "mid", 1003)
ctx.normal_prepare(W)
ctx.diag_factorize(None, use_prepared=True)
print(ctx.layout())
print(ctx.tiling())
print(ctx.time_normal_apply(args.reps, flush_l2=True))
if args.pcr:
    rhs = np.random.default_rng(3).standard_normal(m)
    y, info = ctx.pcr_solve(rhs, 0.0, 1.0 / np.sqrt(W[n:]), args.pcr)
    print({k: v for k, v in info.items() if k != "hist"})
if args.madd:
    from oracle import pyoracle
    rng = np.random.default_rng(4)
    x, lm = rng.standard_normal(n + m), rng.standard_normal(m)
    y, ln = rng.standard_normal(m), rng.standard_normal(n + m)
    A = pyoracle.Csc(*lp.solver_form())
    for trans, rhs, lhs in (("N", x, lm), ("T", y, ln)):
        best = 1e30
        for _ in range(args.madd):
            t0 = time.perf_counter()
            got = ctx.multiply_add(rhs, -1.0, lhs, trans)
            best = min(best, time.perf_counter() - t0)
        t0 = time.perf_counter()
        want = pyoracle.multiply_add(m, n + m, A, rhs, -1.0, lhs, trans)
        cpu = time.perf_counter() - t0
        print(f"multiply_add '{trans}': {1e3 * best:.2f} ms per call through host vectors "
              f"(pageable copies included), C restatement on one core {1e3 * cpu:.1f} ms, "
              f"bit-identical: {got.tobytes() == want.tobytes()}")
ctx.close()
