"""Column sharding of A*W*A'  on CPU: the partition rule of
the C ABI and the sum-of-shards identity, with world_size 2 over gloo."""

import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from ipx_b200 import lpgen


def test_partition_balances_nonzeros():
    from ipx_b200 import capi
    rng = np.random.default_rng(5)
    n = 5000
    counts = rng.integers(0, 40, n)
    counts[100] = 5000  # one heavy column
    Ap = np.zeros(n + 1, np.int64)
    Ap[1:] = np.cumsum(counts)
    for nranks in (1, 2, 3, 8):
        b = capi.partition_columns(n, Ap, nranks)
        assert b[0] == 0 and b[-1] == n and np.all(np.diff(b) >= 0)
        share = np.diff(Ap[b])
        assert share.sum() == Ap[-1]
        # no shard exceeds the ideal share by more than the heaviest column
        assert share.max() <= Ap[-1] / nranks + counts.max()
    # empty matrix and more ranks than columns
    assert list(capi.partition_columns(0, np.zeros(1, np.int64), 3)) == [0, 0, 0, 0]
    b = capi.partition_columns(2, np.array([0, 1, 2]), 4)
    assert b[0] == 0 and b[-1] == 2 and np.all(np.diff(b) >= 0)


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, out_dir):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from ipx_b200 import capi
    from oracle import pyoracle as O
    lp = lpgen.random_sparse_lp(400, 6000, 6, 808)
    m, n = lp.m, lp.n
    AIp, AIi, AIx = lp.solver_form()
    W = lpgen.weights(n + m, "mid", 809)
    x = np.random.default_rng(810).standard_normal(m)
    b = capi.partition_columns(n, AIp, world)
    c0, c1 = int(b[rank]), int(b[rank + 1])
    # this rank's shard: columns [c0, c1) of A, its slice of W; the slack term
    # only on rank 0 (ipx_b200/csrc: OpRowGather.Ws)
    p0, p1 = int(AIp[c0]), int(AIp[c1])
    nl = c1 - c0
    Sp = np.concatenate([AIp[c0:c1 + 1] - p0, (p1 - p0) + np.arange(1, m + 1)])
    Si = np.concatenate([AIi[p0:p1], np.arange(m)])
    Sx = np.concatenate([AIx[p0:p1], np.ones(m)])
    Wl = np.concatenate([W[c0:c1], W[n:] if rank == 0 else np.zeros(m)])
    A = O.Csc(Sp, Si, Sx)
    y, dot = O.normal_apply(m, nl, A, Wl, x)
    diag = O.diag_build(m, nl, A, Wl)
    buf = torch.from_numpy(np.concatenate([y, [dot], diag]))
    dist.all_reduce(buf)  # the one exchange step: sum of the (m+1)-vector
    if rank == 0:
        full = O.Csc(AIp, AIi, AIx)
        y0, dot0 = O.normal_apply(m, n, full, W, x)
        d0 = O.diag_build(m, n, full, W)
        got = buf.numpy()
        np.save(os.path.join(out_dir, "err.npy"), np.array([
            np.abs(got[:m] - y0).max() / np.abs(y0).max(),
            abs(got[m] - dot0) / np.abs(x * y0).sum(),
            np.abs(got[m + 1:] - d0).max() / np.abs(d0).max()]))
    dist.barrier()
    dist.destroy_process_group()


def test_sum_of_shards_equals_full_apply_gloo(tmp_path):
    world = 2
    port = _free_port()
    mp.spawn(_worker, args=(world, port, str(tmp_path)), nprocs=world, join=True)
    err = np.load(tmp_path / "err.npy")
    assert err.max() <= 1e-13, err
