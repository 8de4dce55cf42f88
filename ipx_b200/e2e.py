"""End-to-end LP solves through the unchanged ipx_c.h API of an IPX build (the drop-in build
ipx_b200/_build/libipx_gpu.so, or the reference's own CPU build when a test passes it in), with the
per-iteration table of IPM::PrintOutput (reference src/ipm.cc:659-678) parsed from the log."""

import os
import tempfile
import time

INFO_KEYS = ("status status_ipm status_crossover iter kktiter1 kktiter2 objval pobjval dobjval "
             "time_total time_ipm1 time_ipm2 time_starting_basis time_crossover "
             "time_kkt_factorize time_kkt_solve time_maxvol time_cr1 time_cr1_AAt time_cr1_pre "
             "time_cr2 time_cr2_NNt time_cr2_B time_cr2_Bt time_lu_invert updates_ipm mean_fill "
             "max_fill dense_cols").split()

# BASELINE.json configs as end-to-end solves: (generator arguments, parameters)
CONFIGS = {
    # configs[1]: the diagonal-preconditioned IPM phase (IPX stops where it would switch to
    # basis preconditioning: stop_at_switch = -1, reference src/lp_solver.cc:434-441)
    "C2_diag_phase": (("random", 100_000, 1_000_000, 10, 1002),
                      dict(dualize=0, crossover=0, stop_at_switch=-1)),
    # configs[3]: transportation LP, full IPM solve (both phases) and crossover, so that the
    # objective compared is that of a vertex (exact up to rounding)
    "C4_full_ipm": (("transport", 2000, 5000, 1004), dict(dualize=0, crossover=1)),
}


def make_lp(spec):
    from ipx_b200 import lpgen
    kind, *a = spec
    if kind == "random":
        return lpgen.random_sparse_lp(*a)
    if kind == "transport":
        return lpgen.transportation_lp(*a)
    if kind == "block":
        return lpgen.block_angular_lp(*a)
    raise ValueError(kind)


def parse_log(path):
    """Rows of the IPM's iteration table written with debug >= 1: iteration, residuals,
    objectives, mu, step sizes, basis changes and the CR iterations of the iteration's two
    Newton solves (KKTSolver::iter() restarts at every Factorize)."""
    rows = []
    with open(path) as f:
        for line in f:
            t = line.replace("*", " ").split()
            if len(t) >= 12 and t[0].isdigit() and t[6].endswith("s"):
                try:
                    rows.append({"iter": int(t[0]), "pres": float(t[1]), "dres": float(t[2]),
                                 "pobj": float(t[3]), "dobj": float(t[4]), "mu": float(t[5]),
                                 "step_p": float(t[7]), "step_d": float(t[8]),
                                 "pivots": int(t[9]), "kktiter": int(t[10])})
                except ValueError:
                    pass
    return rows


def solve(lib, lp, per_iter=False, display=0, **params):
    """Solves lp with the IPX build `lib` (anything with lp_solver(): ipxc.IpxC, or the test harness's loader); returns the ipx_info fields, the
    wall time and, with per_iter, the iteration table."""
    s = lib.lp_solver()
    logpath = logbytes = None
    if per_iter:
        fd, logpath = tempfile.mkstemp(prefix="ipx_", suffix=".log")
        os.close(fd)
        logbytes = logpath.encode()  # IPX keeps the pointer: must outlive the solve
        params = dict(params, debug=max(1, int(params.get("debug", 0))), logfile=logbytes)
    s.set_parameters(display=display, **params)
    assert s.load_model(lp) == 0
    t0 = time.perf_counter()
    s.solve()
    wall = time.perf_counter() - t0
    info = s.info()
    s.close()
    out = {k: info[k] for k in INFO_KEYS if k in info}
    out["wall"] = wall
    if per_iter:
        out["per_iter"] = parse_log(logpath)
        os.remove(logpath)
    return out
