"""Generates the golden vectors under tests/golden/ from the REFERENCE ITSELF
(oracle/_ref/libipx_ref.so, compiled from /root/reference by oracle/Makefile).

The reference's own test-suite holds no numeric known-answer for the KKT-solve
path, so these fixtures are the pinned outputs of its
classes - NormalMatrix, DiagonalPrecond, ConjugateResiduals, KKTSolverDiag,
Basis/SplittedNormalMatrix, TriangularSolve and LpSolver - on small seeded
inputs. Run where /root/reference exists:

    python tests/golden/make_golden.py
"""

import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))

from ipx_b200 import lpgen
from oracle import ipxlib  # noqa: E402


def cases():
    return {
        "afiro": lpgen.afiro_lp(),
        "random_60x400": lpgen.random_sparse_lp(60, 400, 4, 901),
        "transport_6x9": lpgen.transportation_lp(6, 9, 902),
    }


def interior_iterate(m, n, lb, ub, seed):
    rng = np.random.default_rng(seed)
    nm = n + m
    has_lb, has_ub = np.isfinite(lb), np.isfinite(ub)
    x = rng.uniform(0.5, 1.5, nm)
    y = rng.standard_normal(m)
    xl = np.where(has_lb, rng.uniform(0.1, 3.0, nm), np.inf)
    xu = np.where(has_ub, rng.uniform(0.1, 3.0, nm), np.inf)
    zl = np.where(has_lb, rng.uniform(0.05, 2.0, nm), 0.0)
    zu = np.where(has_ub, rng.uniform(0.05, 2.0, nm), 0.0)
    return x, xl, xu, y, zl, zu


def generate(name, lp, ref):
    out = {}
    mdl = ref.model(lp)
    m, n = mdl.m, mdl.n
    AIp, AIi, AIx = mdl.AI()
    b, c, lb, ub = mdl.model_vectors()
    out.update(m=m, n=n, AIp=AIp, AIi=AIi, AIx=AIx, lb=lb, ub=ub)
    rng = np.random.default_rng(1234)
    W = lpgen.weights(n + m, "mid", 77)
    x = rng.standard_normal(m)
    out.update(W=W, x=x)
    # NormalMatrix
    mdl.normal_prepare(W)
    out["normal_y"], out["normal_dot"] = mdl.normal_apply(x)
    mdl.normal_prepare(None)
    out["normal_y_nullW"], out["normal_dot_nullW"] = mdl.normal_apply(x)
    # DiagonalPrecond
    mdl.diag_factorize(W)
    out["diag_lhs"], out["diag_dot"] = mdl.diag_apply(x)
    # ConjugateResiduals (preconditioned, with resscale) and unpreconditioned
    mdl.normal_prepare(W)
    rhs = rng.standard_normal(m)
    resscale = 1.0 / np.sqrt(W[n:])
    y, info = mdl.pcr_solve(rhs, 1e-8, resscale, -1)
    out.update(cr_rhs=rhs, cr_resscale=resscale, pcr_y=y, pcr_iter=info["iter"],
               pcr_errflag=info["errflag"])
    y, info = mdl.cr_solve_normal(rhs, 1e-6, None, 1000)
    out.update(cr_y=y, cr_iter=info["iter"], cr_errflag=info["errflag"])
    # KKTSolverDiag
    it = interior_iterate(m, n, lb, ub, 55)
    mdl.iterate_set(*it)
    assert mdl.kktdiag_factorize(True) == 0
    a, bb = rng.standard_normal(n + m), rng.standard_normal(m)
    xs, ys, info = mdl.kktdiag_solve(a, bb, 1e-8)
    out.update(it_xl=it[1], it_xu=it[2], it_zl=it[4], it_zu=it[5], it_x=it[0], it_y=it[3],
               kkt_a=a, kkt_b=bb, kktdiag_x=xs, kktdiag_y=ys, kktdiag_iter=info["kktiter1"],
               kktdiag_err=info["err"])
    # Basis, LU factors, SplittedNormalMatrix
    colscale = np.exp(np.random.default_rng(66).uniform(-3, 3, n + m))
    mdl.basis_from_weights(colscale)
    basis, status = mdl.basis_get()
    L, U, rowperm, colperm = mdl.basis_lu()
    mdl.split_prepare(colscale)
    ysplit, dsplit = mdl.split_apply(x)
    z, info = mdl.cr_solve_split(rhs, 1e-8, 400)
    out.update(colscale=colscale, basis=basis, basis_status=status, Lp=L[0], Li=L[1], Lx=L[2],
               Up=U[0], Ui=U[1], Ux=U[2], rowperm=rowperm, colperm=colperm, split_y=ysplit,
               split_dot=dsplit, split_cr_y=z, split_cr_iter=info["iter"],
               split_cr_errflag=info["errflag"])
    # TriangularSolve, all four variants, on the unscaled factors
    out["tri_L_n"] = ref.triangular_solve(m, *L, x, "n", "l", 1)[0]
    out["tri_U_n"] = ref.triangular_solve(m, *U, x, "n", "u", 0)[0]
    out["tri_U_t"] = ref.triangular_solve(m, *U, x, "t", "u", 0)[0]
    out["tri_L_t"] = ref.triangular_solve(m, *L, x, "t", "l", 1)[0]
    mdl.close()
    # LpSolver end to end
    s = ref.lp_solver()
    s.set_parameters(display=0, dualize=0)
    assert s.load_model(lp) == 0
    status = s.solve()
    info = s.info()
    out.update(lp_status=status, lp_status_ipm=info["status_ipm"],
               lp_status_crossover=info["status_crossover"], lp_iter=info["iter"],
               lp_kktiter1=info["kktiter1"], lp_kktiter2=info["kktiter2"],
               lp_objval=info["objval"], lp_pobjval=info["pobjval"])
    s.close()
    np.savez_compressed(os.path.join(HERE, name + ".npz"), **out)
    print(name, "m", m, "n", n, "pcr_iter", out["pcr_iter"], "lp_iter", out["lp_iter"], "obj",
          out["lp_objval"])


if __name__ == "__main__":
    ref = ipxlib.IpxLibrary(ipxlib.REF_LIB)
    for name, lp in cases().items():
        generate(name, lp, ref)
