"""Triangular solves of a basis factorization: device (ipxgpu_tri_solve) against the oracle,
bit for bit, with timings. The factors come from an .npz written by --make (reference build,
CPU): C3 block-angular LP at a given scale, basis from random weights.

    python tools/tri_bench.py --make .scratch/lu_c3_q.npz --scale 0.25     (CPU, here)
    python tools/tri_bench.py .scratch/lu_c3_q.npz                          (GPU box)
"""
import argparse
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

ap = argparse.ArgumentParser()
ap.add_argument("npz", nargs="?")
ap.add_argument("--make")
ap.add_argument("--scale", type=float, default=0.25)
ap.add_argument("--reps", type=int, default=5)
ap.add_argument("--lt-reference", "--reference-order", dest="lt_reference", action="store_true",
                help="option tri_reference_order: the reference's summation order in every row "
                     "(all four solves bit-identical, slower)")
args = ap.parse_args()

if args.make:
    from tools.fullsize import make_lp
    from oracle import ipxlib
    ref = ipxlib.IpxLibrary(ipxlib.REF_LIB)
    lp = make_lp("C3", args.scale)
    mdl = ref.model(lp, dualize=0)
    cs = np.exp(np.random.default_rng(11).uniform(-3, 3, mdl.n + mdl.m))
    mdl.basis_from_weights(cs)
    (Lp, Li, Lx), (Up, Ui, Ux), _, _ = mdl.basis_lu()
    np.savez(args.make, Lp=Lp, Li=Li.astype(np.int32), Lx=Lx, Up=Up, Ui=Ui.astype(np.int32), Ux=Ux)
    print("wrote", args.make, "m", mdl.m, "nnz", Lp[-1], Up[-1])
    sys.exit(0)

from ipx_b200 import capi
from oracle import pyoracle as O
capi.load()
d = np.load(args.npz)
L = (d["Lp"], d["Li"].astype(np.int64), d["Lx"])
U = (d["Up"], d["Ui"].astype(np.int64), d["Ux"])
m = len(L[0]) - 1
# a context needs a matrix: AI = [I] (no structural columns)
AIp = np.arange(m + 1, dtype=np.int64)
ctx = capi.Context(m, 0, AIp, np.arange(m, dtype=np.int64), np.ones(m))
t0 = time.time()
levels = ctx.lu_load(L, U)
if args.lt_reference:
    ctx.set_option("tri_reference_order", 1)
print(f"m={m} nnz(L)={L[0][-1]} nnz(U)={U[0][-1]} lu_load {time.time() - t0:.2f} s levels {levels}")
x0 = np.random.default_rng(1).standard_normal(m)
for which, (fac, trans, uplo, unit) in enumerate(
        [(L, "n", "l", 1), (U, "n", "u", 0), (U, "t", "u", 0), (L, "t", "l", 1)]):
    ctx.tri_solve(which, x0)  # warm-up
    t0 = time.time()
    for _ in range(args.reps):
        xg = ctx.tri_solve(which, x0)
    tg = (time.time() - t0) / args.reps
    Ao = O.Csc(*fac)
    t0 = time.time()
    xo = O.triangular_solve(m, Ao, x0, trans, uplo, unit)[0]
    tc = time.time() - t0
    same = np.array_equal(xg, xo)
    err = np.abs(xg - xo).max() / max(np.abs(xo).max(), 1e-300)
    td = ctx.time_tri_solve(which, x0, args.reps)
    print(f"system {which} ({uplo}{trans}): device {td:7.3f} ms  through host buffers {1e3 * tg:7.2f} ms  cpu {1e3 * tc:8.2f} ms  "
          f"bit-identical {same}  rel err {err:.2e}", flush=True)
    assert same if args.lt_reference else err <= 1e-9
ctx.close()
