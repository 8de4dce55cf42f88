"""Pins the REFERENCE's end-to-end results at BASELINE.json's full sizes (configs[1] diagonal
phase, configs[3] full IPM solve): status, objective, IPM and CR iteration counts and the
per-iteration table, from oracle/_ref/libipx_ref.so (the unmodified reference compiled by
oracle/Makefile). The reference is single-threaded and deterministic, so the GPU tests compare
the drop-in build against these files instead of re-running 70 s / 230 s of CPU work.

    python tests/golden/make_e2e_golden.py [C2_diag_phase] [C4_full_ipm]
"""

import json
import os
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))

from ipx_b200 import e2e
from oracle import ipxlib  # noqa: E402

if __name__ == "__main__":
    names = sys.argv[1:] or sorted(e2e.CONFIGS)
    ref = ipxlib.IpxLibrary(ipxlib.REF_LIB)
    for name in names:
        spec, params = e2e.CONFIGS[name]
        lp = e2e.make_lp(spec)
        res = e2e.solve(ref, lp, per_iter=True, **params)
        res.update(config=name, lp=lp.name, m=lp.m, n=lp.n, nnz=lp.nnz, params=params,
                   optimum=None if lp.optimum != lp.optimum else lp.optimum)
        with open(os.path.join(HERE, f"e2e_{name}.json"), "w") as f:
            json.dump(res, f, indent=1)
        print(name, {k: res[k] for k in ("status", "status_ipm", "iter", "kktiter1", "kktiter2",
                                         "pobjval", "time_total")}, flush=True)
