#!/usr/bin/env python
"""Benchmark of the IPX KKT-solve hot path (BASELINE.json metric: CR matvecs/s and
HBM GB/s for A*D^2*A').

A step is one diagonally preconditioned Conjugate Residuals solve of exactly
ITERS iterations (tol = 0, maxiter = ITERS), i.e. ITERS+1 applications of
A*D^2*A' plus the preconditioner and vector updates of the CR loop
(reference src/conjugate_residuals.cc:90-213 called from
src/kkt_solver_diag.cc:95-99) on the synthetic LP of BASELINE.json configs[1]
(100k rows x 1M columns, 10 nonzeros per column).

  value  matvecs/s with all vectors resident in HBM (ipxgpu_pcr_solve_dev)
  e2e    the same through the host-buffer C-ABI call ipxgpu_pcr_solve, which
         copies rhs/resscale host->device and the solution device->host inside
         the timed region.

N > 1 (torchrun, one rank per GPU): weak scaling - every rank holds a 1M-column
shard of a 100k x (N*1M) LP, the ranks' partial products are summed once per CR
iteration (in-kernel record exchange over NVLink, or one NCCL allreduce of the
(m+1)-vector); the unit is one 1M-column SHARD application ("shard_matvec/s"):
value = N * (global applies/s), and "global_applies_per_s" carries the other view.

Every line also carries, measured outside the timed region:
  parity         one sharded apply and a 20-iteration PCR solve checked against the
                 CPU oracle on the GLOBAL LP, and a bit-identity check of the ranks'
                 iterates; a mismatch fails the run
  north_star_c5  BASELINE.json configs[4] (1M x 20M, 100M nnz) column-sharded over the
                 same N ranks, strong scaling: CR matvec/s, apply / exchange times
  e2e_ipm        (N = 1) BASELINE's second metric: the diagonal-preconditioned IPM
                 phase of configs[1] through the unchanged ipx_c.h API of the drop-in
                 build (the reference arm reports the same for the reference build)

--impl reference times the reference's own CPU code (oracle/_ref, built from
/root/reference) on the same LP, metric and unit with a bounded sample per step.
"""

import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

REPO = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, REPO)

ITERS = 50            # CR iterations per step (GPU arm)
REF_ITERS = 3         # CR iterations per step (CPU reference arm; ~0.1 s per apply)
M_ROWS = 100_000
N_COLS = 1_000_000
NNZ_PER_COL = 10
SEED = 1002
NCU_TRAFFIC_BYTES_PER_APPLY = 296_715_000  # see roofline.traffic below


def algorithmic_bytes(m, n, nnzA):
    """two-sweep dual layout, int32 indices."""
    return 2 * nnzA * (8 + 4) + 4 * (n + 1) + 4 * (m + 1) + 8 * (n + m) + 8 * m + 8 * m


def measured_peak():
    path = os.path.join(REPO, "MEASURED_PEAKS.json")
    try:
        with open(path) as f:
            return float(json.load(f)["hbm_gbs"]), "measured"
    except Exception:
        return 6650.0, "fallback"


class ClockSampler:
    """Samples nvidia-smi SM clocks and throttle reasons during the timed region."""

    QUERY = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,"
             "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
             "clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index = index
        self.samples = []
        self.stop_flag = threading.Event()
        self.thread = threading.Thread(target=self._run, daemon=True)

    def _run(self):
        while not self.stop_flag.is_set():
            try:
                out = subprocess.run(
                    ["nvidia-smi", f"--id={self.index}", f"--query-gpu={self.QUERY}",
                     "--format=csv,noheader,nounits"], capture_output=True, text=True, timeout=5)
                if out.returncode == 0 and out.stdout.strip():
                    self.samples.append([s.strip() for s in out.stdout.strip().split(",")])
            except Exception:
                pass
            self.stop_flag.wait(0.1)

    def start(self):
        self.thread.start()

    def stop(self):
        self.stop_flag.set()
        self.thread.join(timeout=10)
        sm, smax, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for s in self.samples:
            try:
                sm.append(float(s[0]))
                smax.append(float(s[1]))
            except Exception:
                continue
            for name, v in zip(names, s[2:6]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": float(np.median(sm)) if sm else None,
                "sm_max_mhz": max(smax) if smax else None,
                "reasons": sorted(reasons), "samples": len(sm)}


STRONG = False        # --workload c5: one fixed LP (config 5) column-sharded over the ranks


def set_workload(name):
    """c2 (default): BASELINE.json configs[1], weak scaling (one 1M-column shard per GPU).
    c5: configs[4], 1M x 20M with 100M nonzeros, strong scaling (nnz-balanced column shards)."""
    global M_ROWS, N_COLS, NNZ_PER_COL, SEED, STRONG
    if name == "c5":
        M_ROWS, N_COLS, NNZ_PER_COL, SEED, STRONG = 1_000_000, 20_000_000, 5, 1005, True


def make_lp(world):
    from ipx_b200 import lpgen
    ncols = N_COLS if STRONG else N_COLS * world
    return lpgen.random_sparse_lp(M_ROWS, ncols, NNZ_PER_COL, SEED)


def workload_name(world):
    ncols = N_COLS if STRONG else N_COLS * world
    return (f"synthetic sparse LP {M_ROWS} rows x {ncols} cols, {NNZ_PER_COL} nnz/col, "
            f"KKTSolverDiag diagonal-preconditioned CR")


def cpu_reference_rate(lp, W, rhs, resscale, iters, reps):
    """CR applies/s of the reference's CPU code (NormalMatrix + DiagonalPrecond +
    ConjugateResiduals from oracle/_ref), or of the oracle port if that build is
    absent. One core: the reference has no threading."""
    from oracle import ipxlib
    if os.path.exists(ipxlib.REF_LIB):
        ref = ipxlib.IpxLibrary(ipxlib.REF_LIB)
        mdl = ref.model(lp)
        mdl.normal_prepare(W)
        mdl.diag_factorize(W)

        def step():
            t0 = time.perf_counter()
            _, info = mdl.pcr_solve(rhs, 0.0, resscale, iters)
            assert info["iter"] == iters
            return time.perf_counter() - t0
        kind = "reference"
    else:
        from oracle import pyoracle as O
        AIp, AIi, AIx = lp.solver_form()
        A = O.Csc(AIp, AIi, AIx)
        diag = O.diag_build(lp.m, lp.n, A, W)
        op = O.normal_operator(lp.m, lp.n, A, W)

        def step():
            t0 = time.perf_counter()
            _, info = O.pcr_solve(op, lp.m, diag, rhs, 0.0, resscale, iters)
            assert info["iter"] == iters
            return time.perf_counter() - t0
        kind = "port"
    return step, kind


def run_reference(args):
    """--impl reference: the reference's CPU path on this box's host cores."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    from ipx_b200 import lpgen
    lp = make_lp(1)
    W = lpgen.weights(lp.n + lp.m, "mid", SEED + 1)
    rhs = np.random.default_rng(SEED + 2).standard_normal(lp.m)
    resscale = 1.0 / np.sqrt(W[lp.n:])
    step, kind = cpu_reference_rate(lp, W, rhs, resscale, REF_ITERS, 1)
    for _ in range(args.warmup):
        step()
    t = sum(step() for _ in range(args.steps))
    applies = (REF_ITERS + 1) * args.steps
    value = applies / t
    line = {
        "impl": "reference", "metric": "cr_matvecs_per_sec", "value": value, "unit": "matvec/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": 1e3 * t / args.steps, "higher_is_better": True,
        "scaling": "strong" if STRONG else "weak",
        "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": workload_name(1), "cr_iterations_per_step": REF_ITERS},
        "cpu_baseline": {"value": value, "unit": "matvec/s", "cores": 1, "kind": kind,
                         "sample": f"{args.steps} steps x {REF_ITERS + 1} applies, 1 thread "
                                   f"(the reference is single-threaded) of {os.cpu_count()} cores"},
        "e2e": {"value": value, "unit": "matvec/s", "h2d_bytes_per_step": 0,
                "d2h_bytes_per_step": 0},
    }
    if not args.no_ipm and not STRONG:
        from oracle import ipxlib
        if os.path.exists(ipxlib.REF_LIB):
            line["e2e_ipm"] = ipm_diag_phase(ipxlib.REF_LIB, lp)
    print(json.dumps(line), flush=True)


def ipm_diag_phase(lib_path, lp):
    """End-to-end IPM solve through ipx_c.h (ipx_b200/ipxc.py binds nothing but that API): the
    diagonal-preconditioned phase of the LP (stop_at_switch = -1: IPX stops where it would
    switch to basis preconditioning, no crossover). Returns the ipx_info fields BASELINE.json's
    second metric needs."""
    from ipx_b200 import e2e, ipxc
    res = e2e.solve(ipxc.IpxC(lib_path), lp, dualize=0, crossover=0, stop_at_switch=-1)
    keys = ("status status_ipm iter kktiter1 time_total time_ipm1 time_kkt_factorize "
            "time_kkt_solve time_cr1 time_cr1_AAt time_cr1_pre pobjval dobjval").split()
    out = {k: res[k] for k in keys}
    out["wall_s"] = res["wall"]
    out["workload"] = (f"{lp.name}: diagonal-preconditioned IPM phase (KKTSolverDiag, "
                       "stop_at_switch = -1, no crossover), ipx_c.h API")
    return out


def setup_peers(ctx, world, rank, dev):
    """NCCL communicator and (unless IPXGPU_PEER=0) the NVLink peer exchange buffers."""
    import torch
    import torch.distributed as dist
    from ipx_b200 import capi
    uid = torch.zeros(128, dtype=torch.uint8, device=dev)
    if rank == 0:
        uid.copy_(torch.frombuffer(bytearray(capi.comm_unique_id()), dtype=torch.uint8))
    dist.broadcast(uid, 0)
    ctx.comm_init(bytes(uid.cpu().numpy().tobytes()))
    if os.environ.get("IPXGPU_PEER", "1") != "0":
        # NVLink peer exchange: the CR solve stays one persistent kernel per rank
        mine = torch.frombuffer(bytearray(ctx.peer_export()), dtype=torch.uint8).to(dev)
        allh = [torch.zeros(64, dtype=torch.uint8, device=dev) for _ in range(world)]
        dist.all_gather(allh, mine)
        ctx.peer_import(b"".join(bytes(h.cpu().numpy().tobytes()) for h in allh))


def ranks_identical(y, world, dev):
    """True if every rank holds the same vector, bit for bit."""
    import torch
    import torch.distributed as dist
    if world == 1:
        return True
    t = torch.from_numpy(np.ascontiguousarray(y)).to(dev)
    ref = t.clone()
    dist.broadcast(ref, 0)
    flag = torch.tensor([1 if torch.equal(t.view(torch.int64), ref.view(torch.int64)) else 0],
                        device=dev)
    dist.all_reduce(flag, op=dist.ReduceOp.MIN)
    return bool(flag.item())


PARITY_ITERS = 20
APPLY_TOL = 1e-12      # BASELINE.json north_star: operator apply within 1e-12 relative
PCR_TOL = 1e-9         # iterate after PARITY_ITERS iterations (rounding carried by the recurrences)


def parity_block(ctx, lp, W, rhs, resscale, world, rank, dev):
    """One (sharded) apply and a PARITY_ITERS-iteration PCR solve against the CPU oracle on the
    GLOBAL LP (rank 0), and bit-identity of the ranks' results. Outside the timed region."""
    m, n = lp.m, lp.n
    x = np.random.default_rng(SEED + 3).standard_normal(m)
    y, dot = ctx.normal_apply(x)
    z, info = ctx.pcr_solve(rhs, 0.0, resscale, PARITY_ITERS)
    same = ranks_identical(y, world, dev) and ranks_identical(z, world, dev)
    out = None
    if rank == 0:
        from oracle import pyoracle as O
        AIp, AIi, AIx = lp.solver_form()
        A = O.Csc(AIp, AIi, AIx)
        t0 = time.perf_counter()
        y0, dot0 = O.normal_apply(m, n, A, W, x)
        diag0 = O.diag_build(m, n, A, W)
        z0, info0 = O.pcr_solve(O.normal_operator(m, n, A, W), m, diag0, rhs, 0.0, resscale,
                                PARITY_ITERS)
        apply_err = float(np.abs(y - y0).max() / np.abs(y0).max())
        dot_err = float(abs(dot - dot0) / np.abs(x * y0).sum())
        pcr_err = float(np.abs(z - z0).max() / np.abs(z0).max())
        ok = (apply_err <= APPLY_TOL and dot_err <= APPLY_TOL and pcr_err <= PCR_TOL and
              info["iter"] == info0["iter"] == PARITY_ITERS and
              info["errflag"] == info0["errflag"] and same)
        out = {"ok": bool(ok), "checker": "oracle/ipx_oracle.c on the global LP (1 host core)",
               "apply_rel_err": apply_err, "dot_rel_err": dot_err, "apply_tol": APPLY_TOL,
               "pcr_iters": PARITY_ITERS, "pcr_rel_err": pcr_err, "pcr_tol": PCR_TOL,
               "errflag": [int(info["errflag"]), int(info0["errflag"])],
               "ranks_bit_identical": bool(same), "oracle_s": time.perf_counter() - t0}
    return out


def c5_block(args, world, rank, local_rank, dev, stream):
    """BASELINE.json configs[4] (the north star's multi-GPU config): ONE 1M x 20M LP with 100M
    nonzeros, nnz-balanced column shards over the N ranks (strong scaling), 50 PCR iterations
    per step, same timing rules as the headline."""
    import torch
    import torch.distributed as dist
    from ipx_b200 import capi, lpgen
    m, ncols, k, seed = 1_000_000, 20_000_000, 5, 1005
    t0 = time.perf_counter()
    lp = lpgen.random_sparse_lp(m, ncols, k, seed)
    n = lp.n
    AIp, AIi, AIx = lp.solver_form()
    W = lpgen.weights(n + m, "mid", seed + 1)
    rhs = np.random.default_rng(seed + 2).standard_normal(m)
    resscale = 1.0 / np.sqrt(W[n:])
    t_gen = time.perf_counter() - t0
    t0 = time.perf_counter()
    ctx = capi.Context(m, n, AIp, AIi, AIx, device=local_rank, rank=rank, nranks=world,
                       stream=stream.cuda_stream)
    if world > 1:
        setup_peers(ctx, world, rank, dev)
    ctx.normal_prepare(W)
    ctx.diag_factorize(None, use_prepared=True)
    t_ctx = time.perf_counter() - t0
    layout = ctx.layout()
    d_rhs = torch.from_numpy(rhs).to(dev)
    d_res = torch.from_numpy(resscale).to(dev)
    d_y = torch.zeros(m, dtype=torch.float64, device=dev)

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def step():
        info = ctx.pcr_solve_dev(d_rhs.data_ptr(), 0.0, d_res.data_ptr(), ITERS, d_y.data_ptr())
        assert info["iter"] == ITERS and info["errflag"] == 201, info
        return info
    steps = max(3, min(args.steps, 5))
    for _ in range(3):
        step()
    barrier()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ev0.record(stream)
    t_op = t_pre = 0.0
    for _ in range(steps):
        info = step()
        t_op += info["time_op"]
        t_pre += info["time_pre"]
    ev1.record(stream)
    barrier()
    t_dev = 1e-3 * ev0.elapsed_time(ev1)
    if world > 1:
        tt = torch.tensor([t_dev, t_op, t_pre], dtype=torch.float64, device=dev)
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        t_dev, t_op, t_pre = tt.tolist()
    # parity on this config as well: the sharded solve's iterate is the same on every rank and
    # agrees with the oracle after PARITY_ITERS iterations (rank 0; ~15 s of CPU)
    par = parity_block(ctx, lp, W, rhs, resscale, world, rank, dev)
    tiling = ctx.tiling()
    ctx.close()
    applies = (ITERS + 1) * steps
    ncols_local = layout["col_end"] - layout["col_begin"]
    bytes_apply = algorithmic_bytes(m, ncols_local, layout["nnz_local"])
    peak, _ = measured_peak()
    t_apply = t_op / applies
    return {
        "workload": f"synthetic sparse LP {m} rows x {ncols} cols, {k} nnz/col (100M nnz), "
                    f"column-sharded over {world} GPU(s), {ITERS} PCR iterations per step",
        "scaling": "strong", "cr_matvecs_per_sec": applies / t_dev, "steps": steps,
        "ms_per_step": 1e3 * t_dev / steps,
        "apply_us": 1e6 * t_apply,
        "apply_includes": "sweep 1 + sweep 2 + the ranks' exchange (device timers)",
        "other_us_per_iter": 1e6 * (t_dev - t_op) / applies,
        "precond_us_per_iter": 1e6 * t_pre / applies,
        "shard_algorithmic_bytes_per_apply": bytes_apply,
        "shard_hbm_frac": bytes_apply / t_apply / 1e9 / peak,
        "banded": bool(tiling["sweep1"]["enabled"] and tiling["sweep2"]["enabled"]),
        "exchange": exchange_name(world, bool(tiling["sweep1"]["enabled"] and
                                              tiling["sweep2"]["enabled"])),
        "parity": par, "lp_generate_s": t_gen, "context_s": t_ctx,
    }


def exchange_name(world, banded):
    """What sums the ranks' partial products, from the same switches the library reads."""
    if world == 1:
        return "none"
    how = os.environ.get("IPXGPU_XCHG", "auto")
    if os.environ.get("IPXGPU_PEER", "1") == "0" or how == "nccl":
        return "ncclAllReduce(m+1 f64) per CR iteration"
    where = "in-kernel" if banded else "stand-alone kernel (xchg_records_kernel)"
    if how == "pull" and banded:
        return (f"{where} sum of the ranks' partial products over NVLink peer memory "
                "(P2P loads, per-slice flags) once per CR iteration")
    if how == "two" or (how != "one" and world >= 4):
        return (f"{where} reduce-scatter + all-gather of self-validating 16-byte records "
                "pushed over NVLink peer memory, once per CR iteration")
    return (f"{where} all-to-all of self-validating 16-byte records pushed over "
            "NVLink peer memory, summed in rank order, once per CR iteration")


def run_gpu(args):
    import torch
    import torch.distributed as dist
    from ipx_b200 import capi, lpgen

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if world != args.gpus and world > 1:
        raise SystemExit(f"--gpus {args.gpus} but WORLD_SIZE={world}")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)

    lp = make_lp(world)
    m, n = lp.m, lp.n
    AIp, AIi, AIx = lp.solver_form()
    W = lpgen.weights(n + m, "mid", SEED + 1)
    rhs = np.random.default_rng(SEED + 2).standard_normal(m)
    resscale = 1.0 / np.sqrt(W[n:])

    # This rank's column shard: a contiguous 1M-column slice.
    # (--workload c5: nnz-balanced shards of the one LP, chosen by the library.)
    c0, c1 = (-1, -1) if STRONG else (N_COLS * rank, N_COLS * (rank + 1))
    stream = torch.cuda.Stream(device=dev)  # the library launches on this stream
    torch.cuda.set_stream(stream)
    ctx = capi.Context(m, n, AIp, AIi, AIx, device=local_rank, rank=rank, nranks=world,
                       col_begin=c0 if world > 1 else -1, col_end=c1 if world > 1 else -1,
                       stream=stream.cuda_stream)
    if world > 1:
        setup_peers(ctx, world, rank, dev)
    ctx.normal_prepare(W)
    ctx.diag_factorize(None, use_prepared=True)
    layout = ctx.layout()

    # Device-resident vectors (torch is the allocator) and pinned host buffers.
    d_rhs = torch.from_numpy(rhs).to(dev)
    d_res = torch.from_numpy(resscale).to(dev)
    d_y = torch.zeros(m, dtype=torch.float64, device=dev)
    h_rhs = torch.from_numpy(rhs).pin_memory()
    h_res = torch.from_numpy(resscale).pin_memory()

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def step_dev():
        info = ctx.pcr_solve_dev(d_rhs.data_ptr(), 0.0, d_res.data_ptr(), ITERS, d_y.data_ptr())
        assert info["iter"] == ITERS and info["errflag"] == 201, info
        return info

    def step_host():
        y, info = ctx.pcr_solve(h_rhs.numpy(), 0.0, h_res.numpy(), ITERS)
        assert info["iter"] == ITERS and info["errflag"] == 201, info
        return y

    applies_per_step = ITERS + 1

    # ---- device-resident arm ----
    for _ in range(args.warmup):
        step_dev()
    sampler = ClockSampler(local_rank) if rank == 0 else None
    barrier()
    if sampler:
        sampler.start()
    l0 = ctx.launch_count()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ev0.record(stream)
    t_op = 0.0
    for _ in range(args.steps):
        info = step_dev()
        t_op += info["time_op"]
    ev1.record(stream)
    barrier()
    t_dev = 1e-3 * ev0.elapsed_time(ev1)   # CUDA events on the launching stream
    launches = ctx.launch_count() - l0

    # ---- host-buffer arm (e2e) ----
    for _ in range(min(args.warmup, 3)):
        step_host()
    barrier()
    ev0.record(stream)
    for _ in range(args.steps):
        step_host()
    ev1.record(stream)
    barrier()
    t_e2e = 1e-3 * ev0.elapsed_time(ev1)

    # The timed regions last tens of milliseconds, one nvidia-smi query ~0.1 s: keep the same
    # loop running (untimed) for another 1.5 s so that the clock samples are taken under this load.
    t_end = time.perf_counter() + 1.5
    while time.perf_counter() < t_end:
        step_dev()
    torch.cuda.synchronize()
    clocks = sampler.stop() if sampler else None
    if clocks is not None:
        clocks["window"] = ("timed device and host-buffer arms plus 1.5 s of the same device loop "
                            "right after them")
    barrier()

    if world > 1:
        tt = torch.tensor([t_dev, t_e2e, t_op], dtype=torch.float64, device=dev)
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        t_dev, t_e2e, t_op = tt.tolist()

    total_applies = applies_per_step * args.steps
    # weak scaling: every rank applies its own 1M-column shard; strong: one apply of the LP
    units = 1 if STRONG else world
    value = units * total_applies / t_dev
    e2e = units * total_applies / t_e2e

    # Roofline of the A*D^2*A' apply (sweep 1 + sweep 2) from the device timers
    # inside the timed CR loops: algorithmic bytes of this rank's shard per apply.
    nnz_local = layout["nnz_local"]
    ncols_local = layout["col_end"] - layout["col_begin"]
    bytes_apply = algorithmic_bytes(m, ncols_local, nnz_local)
    t_apply = t_op / total_applies
    peak, peak_src = measured_peak()
    achieved = bytes_apply / t_apply / 1e9
    iso = ctx.time_normal_apply(20, flush_l2=True) if world == 1 else None

    tiling = ctx.tiling()
    banded = bool(tiling["sweep1"]["enabled"] and tiling["sweep2"]["enabled"])
    if not banded:
        kernel_name = ("seg_sweep_kernel<OpColDotScale> + seg_sweep_kernel<OpRowGather> (generic "
                       "sweeps: the banded layout does not fit this shape)")
    elif world == 1 or os.environ.get("IPXGPU_PEER", "1") != "0":
        kernel_name = ("pcr_fused_kernel (persistent CR solve): banded sweep 1 + sweep 2 + combine "
                       "stages of one A*D^2*A' apply")
    else:
        kernel_name = "band_sweep_kernel (sweep 1) + band_sweep_kernel (sweep 2) + band_combine_kernel"
    collective = exchange_name(world, banded)
    line = None
    if rank == 0:
        line = {
            "metric": "cr_matvecs_per_sec", "value": value, "unit": "matvec/s",
            "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": 1e3 * t_dev / args.steps, "higher_is_better": True,
            "scaling": "strong" if STRONG else "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": workload_name(world), "cr_iterations_per_step": ITERS,
                       "rows": m, "cols": n,
                       "nnz": int(lp.nnz), "cols_per_gpu": int(ncols_local),
                       "l2": f"inputs larger than L2 ({24 * nnz_local / 1e6:.0f} MB of matrix data "
                             "per apply and GPU)",
                       "collective": collective},
            "e2e": {"value": e2e, "unit": "matvec/s",
                    "h2d_bytes_per_step": 16 * m, "d2h_bytes_per_step": 8 * m},
            "gpu_launches": int(launches),
            "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s",
                         "frac": achieved / peak,
                         # dram__bytes_read+write of one pcr_fused_kernel launch (a CR solve of
                         # 6 iterations = 7 applies: 2052.27 MB read + 24.74 MB written) / 7, from
                         # the ncu --set full capture profiles/r02_ncu_full_band_fused.csv;
                         # only valid for the 1-GPU C2 workload.
                         "traffic": NCU_TRAFFIC_BYTES_PER_APPLY if world == 1 and not STRONG else None,
                         "traffic_source": "profiles/r02_ncu_full_band_fused.csv",
                         "peak_source": peak_src,
                         "kernel": kernel_name,
                         "algorithmic_bytes_per_apply": bytes_apply,
                         "apply_us_in_loop": 1e6 * t_apply,
                         "apply_us_isolated_l2_flushed": 1e3 * iso["apply_ms"] if iso else None,
                         "sweep1_us": 1e3 * iso["sweep1_ms"] if iso else None,
                         "sweep2_us": 1e3 * iso["sweep2_ms"] if iso else None},
            "clocks": clocks,
        }
        if world > 1 and not STRONG:
            # weak scaling: the unit is one 1M-column shard application
            line["unit"] = line["e2e"]["unit"] = "shard_matvec/s"
            line["global_applies_per_s"] = value / world
    par = None if args.no_parity else parity_block(ctx, lp, W, rhs, resscale, world, rank, dev)
    if rank == 0:
        line["parity"] = par
    ctx.close()
    del ctx

    c5 = None
    if not STRONG and not args.no_c5:
        c5 = c5_block(args, world, rank, local_rank, dev, stream)
    if rank == 0:
        line["north_star_c5"] = c5

    if rank == 0 and world == 1 and not args.no_cpu_baseline and not STRONG:
        # Bounded CPU sample of the same workload: ~20 applies on one core.
        lp1 = lp
        step, kind = cpu_reference_rate(lp1, W, rhs, resscale, 9, 1)
        step()  # warm-up
        t = step() + step()
        line["cpu_baseline"] = {
            "value": 20 / t, "unit": "matvec/s", "cores": 1, "kind": kind,
            "sample": "2 CR solves x 10 applies of the same 100k x 1M LP on one host core "
                      f"(the reference is single-threaded; box has {os.cpu_count()} cores)"}
    if rank == 0 and world == 1 and not args.no_ipm and not STRONG:
        from ipx_b200 import ipxc
        line["e2e_ipm"] = ipm_diag_phase(ipxc.GPU_LIB, lp)
    failed = False
    if rank == 0:
        print(json.dumps(line), flush=True)
        for name, par_k in (("parity", line.get("parity")),
                            ("north_star_c5.parity", (line.get("north_star_c5") or {}).get("parity"))):
            if par_k is not None and not par_k["ok"]:
                sys.stderr.write(f"bench.py: {name} check FAILED: {json.dumps(par_k)}\n")
                failed = True
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    if failed:
        raise SystemExit(1)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-parity", action="store_true", help="skip the oracle parity block")
    ap.add_argument("--no-c5", action="store_true", help="skip the north_star_c5 block")
    ap.add_argument("--no-ipm", action="store_true", help="skip the end-to-end IPM block")
    ap.add_argument("--workload", default="c2", choices=["c2", "c5"],
                    help="c2: BASELINE.json configs[1] (the headline metric; weak scaling); "
                         "c5: configs[4], 1M x 20M, 100M nnz (strong scaling; not the headline)")
    args = ap.parse_args()
    set_workload(args.workload)
    args.warmup = max(args.warmup, 3) if args.impl == "ours" else args.warmup
    if args.impl == "reference":
        run_reference(args)
    else:
        run_gpu(args)


if __name__ == "__main__":
    main()
