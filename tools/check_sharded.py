"""Multi-GPU parity of the sharded CR solve (run under torchrun, one rank per GPU):
every rank holds a column shard; the persistent kernel sums the ranks' partial products
over NVLink peer memory. Rank 0 also solves on the whole matrix alone and compares.

  python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 \
      --master-port 29511 tools/check_sharded.py
"""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from ipx_b200 import capi, lpgen  # noqa: E402

rank = int(os.environ.get("RANK", "0"))
world = int(os.environ.get("WORLD_SIZE", "1"))
local = int(os.environ.get("LOCAL_RANK", "0"))
if rank != 0:
    os.environ.pop("IPXGPU_FUSED_TRACE", None)  # tuning trace: rank 0 only
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)

m, ncols = (100000, 1000000) if os.environ.get("CHECK_SHARDED_BIG") else (20000, 400000)
lp = lpgen.random_sparse_lp(m, ncols * world, 10, 77)
n = lp.n
AIp, AIi, AIx = lp.solver_form()
W = lpgen.weights(n + m, "mid", 78)
rhs = np.random.default_rng(79).standard_normal(m)
resscale = 1.0 / np.sqrt(W[n:])

ctx = capi.Context(m, n, AIp, AIi, AIx, device=local, rank=rank, nranks=world)
uid = torch.zeros(128, dtype=torch.uint8, device=dev)
if rank == 0:
    uid.copy_(torch.frombuffer(bytearray(capi.comm_unique_id()), dtype=torch.uint8))
dist.broadcast(uid, 0)
ctx.comm_init(bytes(uid.cpu().numpy().tobytes()))
mine = torch.frombuffer(bytearray(ctx.peer_export()), dtype=torch.uint8).to(dev)
allh = [torch.zeros(64, dtype=torch.uint8, device=dev) for _ in range(world)]
dist.all_gather(allh, mine)
ctx.peer_import(b"".join(bytes(h.cpu().numpy().tobytes()) for h in allh))

ctx.normal_prepare(W)
ctx.diag_factorize(None, use_prepared=True)
results = []
for it in range(3):  # several solves: the exchange generations carry over
    y, info = ctx.pcr_solve(rhs * (it + 1), 1e-8, resscale, -1)
    results.append((y, info))
ok = True
# every rank must hold the same iterate, bit for bit
for y, info in results:
    t = torch.from_numpy(y).to(dev)
    ref = t.clone()
    dist.broadcast(ref, 0)
    same = bool(torch.equal(t, ref))
    flag = torch.tensor([1 if same else 0], device=dev)
    dist.all_reduce(flag, op=dist.ReduceOp.MIN)
    ok &= bool(flag.item())
if rank == 0:
    solo = capi.Context(m, n, AIp, AIi, AIx, device=local)
    solo.normal_prepare(W)
    solo.diag_factorize(None, use_prepared=True)
    for it, (y, info) in enumerate(results):
        y1, info1 = solo.pcr_solve(rhs * (it + 1), 1e-8, resscale, -1)
        err = np.abs(y - y1).max() / np.abs(y1).max()
        print(f"solve {it}: sharded iter {info['iter']} errflag {info['errflag']} | single-GPU iter "
              f"{info1['iter']} errflag {info1['errflag']} | rel diff {err:.2e}", flush=True)
        ok &= info["errflag"] == 0 and abs(info["iter"] - info1["iter"]) <= 1 and err <= 1e-6
    solo.close()
    print("SHARDED PARITY", "PASS" if ok else "FAIL", f"(ranks identical: {ok})", flush=True)
ctx.close()
dist.barrier()
dist.destroy_process_group()
sys.exit(0 if ok else 1)
