"""The host LU provider that stands in for the un-vendored BASICLU
(ipx_b200/host/sparse_lu.cc, lu_provider.cc), checked through the reference's
own Basis object: B[rowperm, colperm] = (L+I)U with the documented structure
(reference src/lu_factorization.h:22-59)."""

import numpy as np
import pytest
import scipy.sparse as sp

from ipx_b200 import lpgen


def _check_factors(mdl, lp):
    m, n = mdl.m, mdl.n
    AIp, AIi, AIx = mdl.AI()
    AI = sp.csc_matrix((AIx, AIi, AIp), shape=(m, n + m))
    basis, status = mdl.basis_get()
    (Lp, Li, Lx), (Up, Ui, Ux), rowperm, colperm = mdl.basis_lu()
    assert sorted(rowperm) == list(range(m)) and sorted(colperm) == list(range(m))
    L = sp.csc_matrix((Lx, Li, Lp), shape=(m, m))
    U = sp.csc_matrix((Ux, Ui, Up), shape=(m, m))
    # structure: L strictly lower, U upper with the diagonal as last entry, sorted
    for k in range(m):
        li = Li[Lp[k]:Lp[k + 1]]
        ui = Ui[Up[k]:Up[k + 1]]
        assert np.all(li > k) and np.all(np.diff(li) > 0)
        assert len(ui) >= 1 and ui[-1] == k and np.all(np.diff(ui) > 0)
    B = AI[:, basis].tocsc()
    Bperm = B[rowperm, :][:, colperm]
    LU = (L + sp.identity(m)) @ U
    err = abs(Bperm - LU).max()
    scale = abs(B).max()
    assert err <= 1e-10 * max(1.0, scale), err
    # SolveDense agrees with the factors
    rhs = np.random.default_rng(3).standard_normal(m)
    x = mdl.basis_solve_dense(rhs, "N")
    assert np.abs(B @ x - rhs).max() <= 1e-8 * (1 + np.abs(x).max() * scale)
    xt = mdl.basis_solve_dense(rhs, "T")
    assert np.abs(B.T @ xt - rhs).max() <= 1e-8 * (1 + np.abs(xt).max() * scale)
    return Lp[-1] + Up[-1], B.nnz


@pytest.mark.parametrize("case", ["afiro", "random", "blockangular", "transport"])
def test_lu_contract(reflib, case):
    lp = {
        "afiro": lambda: lpgen.afiro_lp(),
        "random": lambda: lpgen.random_sparse_lp(300, 2500, 5, 41),
        "blockangular": lambda: lpgen.block_angular_lp(800, 6000, 5, 42, block_rows=40),
        "transport": lambda: lpgen.transportation_lp(20, 30, 43),
    }[case]()
    mdl = reflib.model(lp)
    rng = np.random.default_rng(44)
    # slack basis first (identity), then a weighted crash basis
    _check_factors(mdl, lp)
    mdl.basis_from_weights(np.exp(rng.uniform(-3, 3, mdl.n + mdl.m)))
    lu_nnz, b_nnz = _check_factors(mdl, lp)
    assert lu_nnz >= b_nnz - mdl.m  # sanity: factors hold at least the off-diagonal pattern
    mdl.close()


def test_singular_basis_is_repaired(reflib):
    """Dependent columns are replaced by unit columns (slack variables enter)."""
    lp = lpgen.random_sparse_lp(50, 200, 4, 45)
    mdl = reflib.model(lp)
    m, n = mdl.m, mdl.n
    status = np.full(n + m, -1, np.int32)
    status[n:] = 0
    # make basis columns 0 and 1 structurally identical duplicates impossible; instead use an
    # all-structural basis whose first two columns are linearly dependent by construction
    status[n:n + 2] = -1
    status[0] = 0
    status[1] = 0
    err = mdl.basis_load(status)
    assert err in (0, 301)  # 301 = IPX_ERROR_basis_singular is not an error state for Basis
    _check_factors(mdl, lp)
    mdl.close()
