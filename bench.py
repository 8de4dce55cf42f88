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
shard of a 100k x (N*1M) LP, one NCCL allreduce of the (m+1)-vector per CR
iteration; a "matvec" unit is one 1M-column shard application, so
value = N * (global applies/s).

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
NCU_TRAFFIC_BYTES_PER_APPLY = 297_520_000  # see roofline.traffic below


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
            f"KKTSolverDiag diagonal-preconditioned CR, {ITERS} iterations per step")


def cpu_reference_rate(lp, W, rhs, resscale, iters, reps):
    """CR applies/s of the reference's CPU code (NormalMatrix + DiagonalPrecond +
    ConjugateResiduals from oracle/_ref), or of the oracle port if that build is
    absent. One core: the reference has no threading."""
    from ipx_b200 import ipxlib
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
        "config": {"workload": workload_name(1).replace(f"{ITERS} iterations",
                                                        f"{REF_ITERS} iterations")},
        "cpu_baseline": {"value": value, "unit": "matvec/s", "cores": 1, "kind": kind,
                         "sample": f"{args.steps} steps x {REF_ITERS + 1} applies, 1 thread "
                                   f"(the reference is single-threaded) of {os.cpu_count()} cores"},
        "e2e": {"value": value, "unit": "matvec/s", "h2d_bytes_per_step": 0,
                "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line), flush=True)


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
    if world == 1:
        collective = "none"
    elif os.environ.get("IPXGPU_PEER", "1") != "0" and (
            banded or os.environ.get("IPXGPU_XCHG", "auto") not in ("pull", "nccl")):
        # (without the banded layouts the same record exchange runs as a kernel of its own
        # in the launch-per-stage CR loop)
        how = os.environ.get("IPXGPU_XCHG", "auto")
        if how == "pull":
            collective = ("in-kernel sum of the ranks' partial products over NVLink peer memory "
                          "(P2P loads, per-slice flags) once per CR iteration")
        elif how == "two" or (how != "one" and world >= 4):
            collective = ("in-kernel reduce-scatter + all-gather of self-validating 16-byte records "
                          "pushed over NVLink peer memory, once per CR iteration")
        else:
            collective = ("in-kernel all-to-all of self-validating 16-byte records pushed over "
                          "NVLink peer memory, summed in rank order, once per CR iteration")
    else:
        collective = "ncclAllReduce(m+1 f64) per CR iteration"
    line = None
    if rank == 0:
        line = {
            "metric": "cr_matvecs_per_sec", "value": value, "unit": "matvec/s",
            "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": 1e3 * t_dev / args.steps, "higher_is_better": True,
            "scaling": "strong" if STRONG else "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": workload_name(world), "rows": m, "cols": n,
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
                         # 6 iterations = 7 applies: 2051.9 MB read + 30.7 MB written) / 7, from
                         # the ncu --set full capture profiles/r01b_ncu_full_band_fused.csv;
                         # only valid for the 1-GPU C2 workload.
                         "traffic": NCU_TRAFFIC_BYTES_PER_APPLY if world == 1 and not STRONG else None,
                         "traffic_source": "profiles/r01b_ncu_full_band_fused.csv",
                         "peak_source": peak_src,
                         "kernel": kernel_name,
                         "algorithmic_bytes_per_apply": bytes_apply,
                         "apply_us_in_loop": 1e6 * t_apply,
                         "apply_us_isolated_l2_flushed": 1e3 * iso["apply_ms"] if iso else None,
                         "sweep1_us": 1e3 * iso["sweep1_ms"] if iso else None,
                         "sweep2_us": 1e3 * iso["sweep2_ms"] if iso else None},
            "clocks": clocks,
        }
    ctx.close()

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
    if rank == 0:
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
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
