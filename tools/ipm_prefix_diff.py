"""How far apart are the two arms after k IPM iterations? Runs the reference build and the
drop-in build with ipm_maxiter = k (k = 1, 2, ...) on the same LP and compares the interior
iterates (x, y, zl, ...) the solvers return. Rounding-level agreement after k iterations and a
visible difference after k+1 localise where the paths part ways."""
import argparse
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from ipx_b200 import lpgen
from oracle import ipxlib  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("lp")
ap.add_argument("--kmax", type=int, default=3)
args = ap.parse_args()
kind, *dims = args.lp.split(":")
d = [int(v) for v in dims]
lp = lpgen.random_sparse_lp(d[0], d[1], d[2], 1002)
libs = {"ref": ipxlib.IpxLibrary(ipxlib.REF_LIB), "gpu": ipxlib.IpxLibrary(ipxlib.GPU_LIB)}
for k in range(1, args.kmax + 1):
    sol = {}
    for arm, lib in libs.items():
        s = lib.lp_solver()
        s.set_parameters(display=0, dualize=0, crossover=0, stop_at_switch=-1, ipm_maxiter=k)
        assert s.load_model(lp) == 0
        s.solve()
        info = s.info()
        err, it = s.interior_solution()
        sol[arm] = (info, it)
        s.close()
    (i0, a), (i1, b) = sol["ref"], sol["gpu"]
    diffs = {key: float(np.abs(a[key] - b[key]).max() / max(1e-300, np.abs(a[key]).max()))
             for key in ("x", "y", "zl", "slack")}
    print(f"after {k} iteration(s): kktiter1 ref/gpu {i0['kktiter1']}/{i1['kktiter1']}  "
          f"rel. differences {diffs}", flush=True)
