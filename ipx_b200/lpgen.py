"""Synthetic sparse LPs of the shapes named in BASELINE.json.

All matrices are CSC with int64 indices sorted per column, no duplicate rows in
a column, and |a_ij| in [0.5, 8) so that the reference's equilibration is a
no-op (reference src/presolver.cc:909-924) and, with dualize=0, the solver-form
matrix is AI = [A I] (src/presolver.cc:147-153).

The LPs are primal/dual feasible by construction with a known optimal value
(complementary x0, (y0, z0)), so the end-to-end objective can be checked
without any solver.
"""

from dataclasses import dataclass, field

import numpy as np


@dataclass
class LP:
    """User-form LP: min obj'x s.t. A x {=,<,>} rhs, lb <= x <= ub."""
    m: int
    n: int
    Ap: np.ndarray
    Ai: np.ndarray
    Ax: np.ndarray
    rhs: np.ndarray
    constr_type: bytes
    obj: np.ndarray
    lb: np.ndarray
    ub: np.ndarray
    optimum: float = float("nan")
    name: str = ""
    extra: dict = field(default_factory=dict)

    @property
    def nnz(self):
        return int(self.Ap[-1])

    def solver_form(self):
        """Returns (AIp, AIi, AIx) of AI = [A I] (m x (n+m)), int64 CSC."""
        m, n = self.m, self.n
        AIp = np.concatenate([self.Ap, self.Ap[-1] + np.arange(1, m + 1, dtype=np.int64)])
        AIi = np.concatenate([self.Ai, np.arange(m, dtype=np.int64)])
        AIx = np.concatenate([self.Ax, np.ones(m)])
        return AIp, AIi, AIx


def _distinct_rows(rng, n, k, lo, hi):
    """k distinct sorted draws from [lo_j, hi_j) for each of n columns."""
    span = hi - lo
    rows = lo[:, None] + (rng.random((n, k)) * span[:, None]).astype(np.int64)
    rows.sort(axis=1)
    for _ in range(100):
        dup = np.nonzero((rows[:, 1:] == rows[:, :-1]).any(axis=1))[0]
        if dup.size == 0:
            break
        rows[dup] = lo[dup, None] + (rng.random((dup.size, k)) * span[dup, None]).astype(np.int64)
        rows[dup] = np.sort(rows[dup], axis=1)
    else:
        raise RuntimeError("could not draw distinct rows")
    return rows


def _values(rng, size):
    return rng.uniform(0.5, 4.0, size) * rng.choice(np.array([-1.0, 1.0]), size)


def _feasible_rhs_obj(rng, m, n, Ap, Ai, Ax):
    """x0 >= 0 with a third of the entries at 0; complementary dual (y0, z0)."""
    x0 = np.where(rng.random(n) < 1.0 / 3.0, 0.0, rng.uniform(0.5, 1.5, n))
    y0 = rng.standard_normal(m)
    col = np.repeat(np.arange(n, dtype=np.int64), np.diff(Ap))
    rhs = np.bincount(Ai, weights=Ax * x0[col], minlength=m)
    aty = np.bincount(col, weights=Ax * y0[Ai], minlength=n)
    z0 = np.where(x0 > 0.0, 0.0, rng.uniform(0.5, 1.5, n))
    obj = aty + z0
    return x0, rhs, obj


def random_sparse_lp(m, n, k, seed, name=""):
    """Config 2 / 5 shape: k distinct uniformly random rows per column."""
    rng = np.random.default_rng(seed)
    rows = _distinct_rows(rng, n, k, np.zeros(n, np.int64), np.full(n, m, np.int64))
    Ap = np.arange(0, (n + 1) * k, k, dtype=np.int64)
    Ai = rows.reshape(-1)
    Ax = _values(rng, n * k)
    x0, rhs, obj = _feasible_rhs_obj(rng, m, n, Ap, Ai, Ax)
    return LP(m, n, Ap, Ai, Ax, rhs, b"=" * m, obj, np.zeros(n), np.full(n, np.inf),
              float(obj @ x0), name or f"random_{m}x{n}_k{k}")


def dense_column_lp(m, n, k, ndense, seed, dense_frac=0.5, name=""):
    """random_sparse_lp plus `ndense` dense columns (a fraction `dense_frac` of the rows each),
    spread over the column range. Model::FindDenseColumns (reference src/model.cc:34-56) marks
    a column dense when its count exceeds max(40, 10 x the next smaller count), so these
    columns take the Sherman-Morrison-Woodbury branch of the diagonal preconditioner
    (src/diagonal_precond.cc:48-102) under the default precond_dense_cols = 1."""
    rng = np.random.default_rng(seed)
    kd = max(41, 10 * k + 1, int(m * dense_frac))
    assert kd <= m
    where = np.sort(rng.choice(n, ndense, replace=False))
    is_dense = np.zeros(n, bool)
    is_dense[where] = True
    counts = np.where(is_dense, kd, k).astype(np.int64)
    Ap = np.concatenate([[0], np.cumsum(counts)]).astype(np.int64)
    Ai = np.empty(int(Ap[-1]), np.int64)
    sparse_rows = _distinct_rows(rng, n - ndense, k, np.zeros(n - ndense, np.int64),
                                 np.full(n - ndense, m, np.int64))
    starts = Ap[:-1][~is_dense]
    Ai[(starts[:, None] + np.arange(k)[None, :]).reshape(-1)] = sparse_rows.reshape(-1)
    for j in where:
        Ai[Ap[j]:Ap[j + 1]] = np.sort(rng.choice(m, kd, replace=False))
    Ax = _values(rng, int(Ap[-1]))
    x0, rhs, obj = _feasible_rhs_obj(rng, m, n, Ap, Ai, Ax)
    return LP(m, n, Ap, Ai, Ax, rhs, b"=" * m, obj, np.zeros(n), np.full(n, np.inf),
              float(obj @ x0), name or f"densecols_{m}x{n}_k{k}_d{ndense}",
              {"dense_cols": [int(j) for j in where], "dense_count": int(kd)})


def block_angular_lp(m, n, k, seed, block_rows=200, link_frac=0.005, name=""):
    """Config 3 shape: diagonal blocks of `block_rows` rows plus linking rows.

    Column j draws k-1 rows inside its block and, with probability 0.1, its last
    row among the linking rows (else inside the block as well), which keeps the
    LU factors of IPM bases sparse.
    """
    rng = np.random.default_rng(seed)
    nlink = max(1, int(m * link_frac))
    mb = m - nlink
    K = max(1, mb // block_rows)
    bounds = np.linspace(0, mb, K + 1).astype(np.int64)
    blk = (np.arange(n, dtype=np.int64) * K) // n
    lo, hi = bounds[blk], bounds[blk + 1]
    rows = _distinct_rows(rng, n, k, lo, hi)
    linked = rng.random(n) < 0.1
    rows[linked, -1] = mb + rng.integers(0, nlink, int(linked.sum()))
    Ap = np.arange(0, (n + 1) * k, k, dtype=np.int64)
    Ai = rows.reshape(-1)
    Ax = _values(rng, n * k)
    x0, rhs, obj = _feasible_rhs_obj(rng, m, n, Ap, Ai, Ax)
    return LP(m, n, Ap, Ai, Ax, rhs, b"=" * m, obj, np.zeros(n), np.full(n, np.inf),
              float(obj @ x0), name or f"blockangular_{m}x{n}_k{k}",
              {"blocks": int(K), "linking_rows": int(nlink)})


def transportation_lp(S, T, seed, name=""):
    """Config 4: S sources x T sinks; column (i,j) = e_i + e_{S+j}.

    Supply rows '<', demand rows '='; total supply = 1.2 * total demand.
    """
    rng = np.random.default_rng(seed)
    m, n = S + T, S * T
    src = np.repeat(np.arange(S, dtype=np.int64), T)
    dst = S + np.tile(np.arange(T, dtype=np.int64), S)
    Ai = np.stack([src, dst], axis=1).reshape(-1)
    Ap = np.arange(0, 2 * n + 1, 2, dtype=np.int64)
    Ax = np.ones(2 * n)
    d = rng.uniform(50.0, 150.0, T)
    s = rng.uniform(0.8, 1.2, S)
    s *= 1.2 * d.sum() / s.sum()
    rhs = np.concatenate([s, d])
    obj = rng.uniform(1.0, 2.0, n)
    ctype = b"<" * S + b"=" * T
    return LP(m, n, Ap, Ai, Ax, rhs, ctype, obj, np.zeros(n), np.full(n, np.inf),
              float("nan"), name or f"transport_{S}x{T}")


def afiro_lp():
    """Config 1: the presolved 12-variable x 9-constraint afiro of the
    reference's example (data as in reference example/afiro.cc:12-46;
    optimum -464.753142857143)."""
    inf = np.inf
    obj = np.array([-0.2194, 0.0, 0.0, 0.0, 0.0, 0.0, 0.0, 0.0, -0.32, -0.5564, 0.6, -0.48])
    ub = np.array([80.0, 283.303, 283.303, 312.813, 349.187, inf, inf, inf, 57.201,
                   500.0, 500.501, 357.501])
    Ap = np.array([0, 2, 6, 10, 14, 18, 20, 22, 24, 26, 28, 30, 32], dtype=np.int64)
    Ai = np.array([0, 5, 1, 6, 7, 8, 2, 6, 7, 8, 3, 6, 7, 8, 4, 6, 7, 8, 1, 2, 2, 3,
                   2, 4, 0, 6, 0, 5, 2, 5, 5, 7], dtype=np.int64)
    Ax = np.array([-1.0, 0.301, 1.0, -1.0, 0.301, 1.06, 1.0, -1.0, 0.313, 1.06,
                   1.0, -1.0, 0.313, 0.96, 1.0, -1.0, 0.326, 0.86, -1.0, 0.99078,
                   1.00922, -1.0, 1.01802, -1.0, 1.4, 1.0, 0.109, -1.0, -0.419111, 1.0,
                   1.4, -1.0])
    rhs = np.array([0.0, 80.0, 0.0, 0.0, 0.0, 0.0, 0.0, 44.0, 300.0])
    ctype = b"<<=<<=<<<"
    m, n = 9, 12
    return LP(m, n, Ap, Ai, Ax, rhs, ctype, obj, np.zeros(n), ub, -464.753142857143, "afiro")


def weights(n_plus_m, regime, seed):
    """Weight regimes of SURVEY.md §8d operator-parity vectors."""
    rng = np.random.default_rng(seed)
    if regime == "ones":
        return np.ones(n_plus_m)
    if regime == "mid":
        return np.exp(rng.uniform(-2.0, 2.0, n_plus_m))
    if regime == "wide":
        return np.exp(rng.uniform(-15.0, 15.0, n_plus_m))
    raise ValueError(regime)
