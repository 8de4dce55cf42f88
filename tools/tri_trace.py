"""Timeline of one sync-free triangular solve (tuning tool): every row stamps %globaltimer when
its warp starts it and when its value is final (context option "tri_trace"). Prints where the
solve's time goes: the critical path through the dependency DAG, the latency of its hops (finish
of a row minus finish of the dependency it waited for last), split by whether producer and
consumer ran in the same CTA.

    python tools/tri_trace.py .scratch/lu_c3_q.npz --which 0          (GPU box)
"""
import argparse
import os
import sys

import numpy as np
import scipy.sparse as sp

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from ipx_b200 import capi  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("npz")
ap.add_argument("--which", type=int, default=0)
args = ap.parse_args()
capi.load()
d = np.load(args.npz)
L = (d["Lp"], d["Li"].astype(np.int64), d["Lx"])
U = (d["Up"], d["Ui"].astype(np.int64), d["Ux"])
m = len(L[0]) - 1
ctx = capi.Context(m, 0, np.arange(m + 1, dtype=np.int64), np.arange(m, dtype=np.int64), np.ones(m))
ctx.lu_load(L, U)
x0 = np.random.default_rng(1).standard_normal(m)
ctx.tri_solve(args.which, x0)
ctx.set_option("tri_trace", 1)
ctx.tri_solve(args.which, x0)
tr = ctx.tri_trace().astype(np.int64)
t0 = tr[:, 0].min()
start, fin = (tr[:, 0] - t0) * 1e-3, (tr[:, 1] - t0) * 1e-3   # us
print(f"system {args.which}: first start 0, last finish {fin.max():.1f} us")
# gather form: row i depends on the columns of G's row i
Lm = sp.csc_matrix((L[2], L[1], L[0]), shape=(m, m))
Um = sp.csc_matrix((U[2], U[1], U[0]), shape=(m, m))
G = {0: sp.tril(Lm, -1).tocsr(), 1: sp.triu(Um, 1).tocsr(), 2: sp.triu(Um, 1).T.tocsr(),
     3: sp.tril(Lm, -1).T.tocsr()}[args.which]
ptr, idx = G.indptr, G.indices
has = np.diff(ptr) > 0
# the dependency that finished last, per row
fdep = np.full(m, -1.0)
last = np.full(m, -1, np.int64)
rows = np.repeat(np.arange(m), np.diff(ptr))
order = np.lexsort((fin[idx], rows))
lastpos = ptr[1:][has] - 1
last[has] = idx[order][lastpos]
fdep[has] = fin[last[has]]
hop = fin - np.maximum(fdep, start)          # after both the warp got to the row and the dep
wait = np.maximum(fdep - start, 0.0)
print(f"rows {m}, with dependencies {has.sum()}; rows that waited for a dependency "
      f"{(fdep > start).sum()}")
w = fdep > start
print("hop latency of rows that waited (us): median %.2f  mean %.2f  p90 %.2f" %
      (np.median(hop[w]), hop[w].mean(), np.percentile(hop[w], 90)))
print("row time of rows that did not wait (us): median %.2f  mean %.2f  p90 %.2f" %
      (np.median(hop[~w]), hop[~w].mean(), np.percentile(hop[~w], 90)))
# same CTA or not (round-robin schedule, IPXGPU_TRI_AFFINITY=0): position in the level order
level = np.zeros(m, np.int64)
asc = args.which in (0, 2)
for i in (range(m) if asc else range(m - 1, -1, -1)):
    if ptr[i + 1] > ptr[i]:
        level[i] = level[idx[ptr[i]:ptr[i + 1]]].max() + 1
pos = np.empty(m, np.int64)
pos[np.argsort(level, kind="stable")] = np.arange(m)
ncta = min(148, (m + 11) // 12)
cta = (pos % (ncta * 12)) // 12
near = w & (last >= 0) & (cta[np.maximum(last, 0)] == cta)
far = w & ~near
for name, sel in (("same CTA", near), ("other CTA", far)):
    if sel.any():
        print("rows whose last dependency ran in the %s: %d, hop median %.2f mean %.2f p10 %.2f p90 %.2f us"
              % (name, sel.sum(), np.median(hop[sel]), hop[sel].mean(), np.percentile(hop[sel], 10),
                 np.percentile(hop[sel], 90)))
# critical path: walk back from the last row along `last`
i = int(np.argmax(fin))
path = []
while i >= 0:
    path.append(i)
    i = int(last[i]) if (has[i] and fdep[i] > start[i]) else -1
path = path[::-1]
print(f"critical path: {len(path)} rows, from {fin[path[0]]:.1f} to {fin[path[-1]]:.1f} us; "
      f"first row started at {start[path[0]]:.1f} us")
hops = np.diff(fin[path])
nnz_row = np.diff(ptr)[path[1:]]
print("hops on the critical path (us): median %.2f mean %.2f p90 %.2f max %.2f; entries per row "
      "median %d max %d" % (np.median(hops), hops.mean(), np.percentile(hops, 90), hops.max(),
                            np.median(nnz_row), nnz_row.max()))
for lo, hi in [(0, 0.2), (0.2, 0.5), (0.5, 1.0), (1.0, 2.0), (2.0, 5.0), (5.0, 1e9)]:
    sel = (hops >= lo) & (hops < hi)
    print(f"  hops in [{lo}, {hi}) us: {sel.sum():5d}  total {hops[sel].sum():8.1f} us")
ctx.close()
