#!/usr/bin/env python
"""bench.py -- fp64 DOF-updates/s of the matrix-free IP-DG heat operator y = A x.

Workload (BASELINE.json `configs[4]`): uniform 2-D mesh of Q_p cells,
two phases (circle of radius 0.4, k1 = 1, k2 = 2), mesh-aligned interface with
flux + penalty + Robin terms (lambda = 1), Dirichlet boundary, pen = 1e3 * nx.
Default p = 1, 5000 x 5000 cells per GPU = 100 M dofs per GPU (weak scaling: rank r
owns the r-th slab of 5000 cell columns of a (5000 N) x 5000 mesh).  N > 1: the slab
vectors live in CUDA-IPC peer memory and every rank's apply kernel reads its neighbours'
edge columns over NVLink itself (`--halo peer`, the default: one launch per application,
ordered by epoch flags that the same kernel publishes, waits for and releases); `--halo nccl` exchanges the columns with NCCL send/recv on a side stream,
overlapped with the interior columns.

One "step" = one application of the operator.  `value` = dofs x steps / seconds with
x resident in HBM; `e2e` = the same through stefan_ipdg_apply_host (host buffers,
pinned; H2D and D2H inside the timed region).  `--impl reference` times the CPU
path the reference would take (assemble COO -> sparse -> A*x), as restated in
oracle/csrc/ip_oracle.c, on a bounded sample with all host threads.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "fp64 DOF-updates/sec for IP-DG heat operator"
UNIT = "DOF-updates/s"
ALGO_BYTES_PER_DOF = 16  # read x once + write y once


def measured_hbm_peak():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons while the timed region runs."""

    def __init__(self, gpu_index):
        self.gpu = gpu_index
        self.samples = []
        self.reasons = set()
        self.sm_max = None
        self._stop = threading.Event()
        self._t = None

    def _run(self):
        q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,"
             "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
             "clocks_event_reasons.sw_power_cap")
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        while not self._stop.is_set():
            try:
                out = subprocess.run(["nvidia-smi", "-i", str(self.gpu), f"--query-gpu={q}",
                                      "--format=csv,noheader,nounits"], capture_output=True,
                                     text=True, timeout=5).stdout.strip().split(",")
                self.samples.append(float(out[0]))
                self.sm_max = float(out[1])
                for n, v in zip(names, out[2:]):
                    if v.strip().lower().startswith("active"):
                        self.reasons.add(n)
            except Exception:
                pass
            self._stop.wait(0.1)

    def __enter__(self):
        self._t = threading.Thread(target=self._run, daemon=True)
        self._t.start()
        return self

    def __exit__(self, *a):
        self._stop.set()
        self._t.join(timeout=5)

    def summary(self):
        s = sorted(self.samples)
        med = s[len(s) // 2] if s else None
        return {"sm_mhz": med, "sm_max_mhz": self.sm_max, "reasons": sorted(self.reasons)}


def circle_phase(nx_global, ny, i0, nxl):
    import numpy as np
    i, j = np.divmod(np.arange(nxl * ny, dtype=np.int64), ny)
    cx, cy = (i0 + i + 0.5) / nx_global, (j + 0.5) / ny
    return np.where((cx - 0.5) ** 2 + (cy - 0.5) ** 2 < 0.16, -1, 1).astype(np.int8)


def column_phase(nx_global, ny, i):
    import numpy as np
    if i < 0 or i >= nx_global:
        return None
    j = np.arange(ny)
    cx, cy = (i + 0.5) / nx_global, (j + 0.5) / ny
    return np.where((cx - 0.5) ** 2 + (cy - 0.5) ** 2 < 0.16, -1, 1).astype(np.int8)


def run_ours(args):
    import numpy as np
    import torch
    import torch.distributed as dist
    from stefanproblem_b200 import cutcelldg as cc

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if args.gpus != world:
        if world == 1 and args.gpus > 1:
            raise SystemExit("launch with torch.distributed.run --nproc-per-node N for --gpus N")
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))

    p, nxl, ny = args.order, args.nx, args.ny
    nf = (p + 1) ** 2
    nxg = nxl * world
    from stefanproblem_b200.distributed import DistributedIPOperator, PeerSlab, apply_peer, apply_peer_fused
    dop = DistributedIPOperator(nxg, ny, (1.0, 1.0), p, p + 1, 1.0, 2.0, 1e3 * nxg, 1.0, 0.5,
                                lambda i: column_phase(nxg, ny, i), rank, world)
    op = dop.op
    ndofs_local = op.ndofs
    gen = torch.Generator(device="cuda")
    gen.manual_seed(1234 + rank)
    x = torch.rand(ndofs_local, dtype=torch.float64, device="cuda", generator=gen) * 2 - 1
    y = torch.empty_like(x)
    x_lo, x_hi = dop.halo.x_lo, dop.halo.x_hi
    peer = world > 1 and args.halo == "peer"
    slab = None
    if peer:
        # collective; if CUDA IPC is not available to some rank (container policy), every rank
        # falls back to the NCCL halos together and the line says so
        failed = 0
        try:
            slab = PeerSlab(ndofs_local, rank, world)
        except Exception as exc:  # noqa: BLE001
            failed = 1
            sys.stderr.write(f"rank {rank}: peer memory unavailable ({exc}); falling back to --halo nccl\n")
        flag = torch.tensor([failed], dtype=torch.int32, device="cuda")
        dist.all_reduce(flag, op=dist.ReduceOp.MAX)
        if int(flag.item()):
            if slab is not None:
                slab.close()
            slab, peer = None, False
        else:
            slab.x.copy_(x)
            x = slab.x

    def step():
        if peer:
            # ONE launch: the kernel publishes my vector's epoch, its warps wait for the
            # neighbours' epoch before they read their edge columns over NVLink, its last CTA
            # releases them
            apply_peer_fused(dop, slab, y)
        else:
            # halo trace exchange (NCCL, side stream) overlapped with the interior columns
            dop.apply(x, y)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(args.warmup):
        step()
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    with ClockSampler(local) as clk:
        barrier()
        e0.record()
        for _ in range(args.steps):
            step()
        e1.record()
        barrier()
    ms = e0.elapsed_time(e1)
    if world > 1:
        t = torch.tensor([ms], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t.item())
    ms_per_step = ms / args.steps
    total_dofs = ndofs_local * world
    value = total_dofs / (ms_per_step * 1e-3)

    # kernel-only duration for the roofline (no halo exchange), same stream, CUDA events
    barrier()
    k0, k1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    k0.record()
    for _ in range(args.steps):
        if peer:
            apply_peer(dop, slab, y, publish=False)
        else:
            op.apply(x, out=y, x_lo=x_lo, x_hi=x_hi)
    k1.record()
    torch.cuda.synchronize()
    kernel_ms = k0.elapsed_time(k1) / args.steps
    peak, peak_src = measured_hbm_peak()
    achieved = ALGO_BYTES_PER_DOF * ndofs_local / (kernel_ms * 1e-3) / 1e9

    # end to end through the C ABI with host buffers (rank-local slab, no halo: N = 1 only)
    e2e = None
    if world == 1:
        xh = torch.empty(ndofs_local, dtype=torch.float64).pin_memory()
        xh.copy_(x.cpu())
        yh = torch.empty(ndofs_local, dtype=torch.float64).pin_memory()
        n_e2e = max(2, min(args.steps, 10))
        op.apply_host(xh, yh)
        t0 = time.perf_counter()
        for _ in range(n_e2e):
            op.apply_host(xh, yh)
        dt = (time.perf_counter() - t0) / n_e2e
        e2e = {"value": ndofs_local / dt, "unit": UNIT, "h2d_bytes_per_step": ndofs_local * 8,
               "d2h_bytes_per_step": ndofs_local * 8, "steps": n_e2e,
               "api": "stefan_ipdg_apply_host (pinned host buffers, 32 pipelined slabs)"}
        del xh, yh

    if world > 1 and peer:
        # end to end at N GPUs: every rank moves its slab over its own PCIe link (pinned host
        # buffers, H2D of x and D2H of y inside the timed region) around the one fused launch
        xh = torch.empty(ndofs_local, dtype=torch.float64).pin_memory()
        xh.copy_(x.cpu())
        yh = torch.empty(ndofs_local, dtype=torch.float64).pin_memory()
        n_e2e = max(2, min(args.steps, 10))
        ybuf = [y, torch.empty_like(y)]
        d2h, done, copied = torch.cuda.Stream(), [torch.cuda.Event(), torch.cuda.Event()], [None, None]
        work = torch.cuda.Stream()  # not the legacy default stream, which serialises with other streams' copies

        def e2e_step(k):
            # the result of step k leaves on a second stream while step k+1's input arrives
            # (PCIe is full duplex); two result buffers
            b = k & 1
            slab.wait_released()            # the neighbours are done with my previous vector
            slab.x.copy_(xh, non_blocking=True)
            if copied[b] is not None:
                torch.cuda.current_stream().wait_event(copied[b])   # ybuf[b] has left the device
            apply_peer_fused(dop, slab, ybuf[b])
            done[b].record()
            with torch.cuda.stream(d2h):
                d2h.wait_event(done[b])
                yh.copy_(ybuf[b], non_blocking=True)
                copied[b] = torch.cuda.Event()
                copied[b].record(d2h)

        barrier()
        with torch.cuda.stream(work):
            e2e_step(0)
            barrier()
            t0 = time.perf_counter()
            for k in range(n_e2e):
                e2e_step(k + 1)
            barrier()
        dt = torch.tensor([(time.perf_counter() - t0) / n_e2e], dtype=torch.float64, device="cuda")
        dist.all_reduce(dt, op=dist.ReduceOp.MAX)
        e2e = {"value": total_dofs / float(dt.item()), "unit": UNIT, "h2d_bytes_per_step": total_dofs * 8,
               "d2h_bytes_per_step": total_dofs * 8, "steps": n_e2e,
               "api": "PeerSlab + stefan_ipdg_apply_peer per rank (pinned host buffers, each rank over its own PCIe link)"}
        del xh, yh

    cpu_baseline = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        cpu_baseline = cpu_port_baseline(p, seconds=6.0)

    if rank == 0:
        # N > 1, nccl halos: interior + two edge columns; peer halos: the one apply kernel
        launches = args.steps * (3 if (world > 1 and not peer) else 1)
        if peer:
            assert int(slab.status.item()) == 0, "a peer-flag wait timed out"
        partition = (f"{world} column slabs; neighbours' edge columns read over NVLink from CUDA-IPC peer memory "
                     "inside the apply kernel, which also publishes / waits for / releases the epoch flags: one launch per step") if peer else (
                     f"{world} column slabs; 1 halo cell column per neighbour over NCCL send/recv, "
                     "overlapped with the interior columns")
        if world == 1:
            partition = "1 slab (no halo)"
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms_per_step, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {
                "workload": f"IP-DG p={p} matrix-free apply, two-phase circle + Robin interface, "
                            f"{nxl}x{ny} cells per GPU ({ndofs_local / 1e6:.1f} M dofs per GPU)",
                "order": p, "cells_per_gpu": [nxl, ny], "global_cells": [nxg, ny],
                "global_dofs": total_dofs, "partition": partition,
                "l2": "inputs larger than L2 (x and y are 800 MB each per GPU at the default size)",
            },
            "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s",
                         "frac": achieved / peak, "traffic": args.traffic_bytes,
                         "peak_source": peak_src, "kernel_ms": kernel_ms,
                         "algorithmic_bytes_per_dof": ALGO_BYTES_PER_DOF},
            "clocks": clk.summary(),
            "gpu_launches": launches,
        }
        if e2e is not None:
            line["e2e"] = e2e
        if cpu_baseline is not None:
            line["cpu_baseline"] = cpu_baseline
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.barrier()
        if slab is not None:
            slab.close()
        dist.destroy_process_group()


_REF_CACHE = {}


def _cpu_port(order, n):
    """Assemble once per process (the reference's COO -> sparse path, restated in C)."""
    import numpy as np
    from oracle.cport import CIPOracle
    key = (order, n)
    if key not in _REF_CACHE:
        phase = circle_phase(n, n, 0, n)
        t0 = time.perf_counter()
        C = CIPOracle(order, order + 1, n, n, 1.0 / n, 1.0 / n, 1.0, 2.0, 1e3 * n, 1.0, "current",
                      255, phase)
        t_asm = time.perf_counter() - t0
        _REF_CACHE[key] = (C, np.random.default_rng(0).uniform(-1, 1, C.nrows), t_asm)
    return _REF_CACHE[key]


def _spmv_rate(C, x, seconds, threads=0):
    import numpy as np
    y = np.empty_like(x)
    C.spmv(x, y, threads)
    reps, t0 = 0, time.perf_counter()
    while time.perf_counter() - t0 < seconds:
        C.spmv(x, y, threads)
        reps += 1
    return C.nrows / ((time.perf_counter() - t0) / reps), reps


def cpu_port_baseline(order, seconds=6.0, n=512, threads=0):
    """The reference's CPU path (COO assembly -> sparse -> A*x) as restated in C, on a
    bounded sample: n x n cells, same physics as the GPU workload, all host threads."""
    from oracle.cport import CIPOracle
    C, x, t_asm = _cpu_port(order, n)
    rate, reps = _spmv_rate(C, x, seconds, threads)
    cores = CIPOracle.max_threads() if threads <= 0 else threads
    return {"value": rate, "unit": UNIT, "cores": cores, "kind": "port",
            "sample": f"{n}x{n} cells p={order} ({C.nrows / 1e6:.2f} M dofs): CSR SpMV of the "
                      f"oracle-assembled matrix, {reps} applications; assembly (COO + compress) "
                      f"took {t_asm:.1f} s once and is not in the rate",
            "assembly_seconds": t_asm, "nnz": C.nnz}


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    C, x, _ = _cpu_port(args.order, 512)
    for _ in range(args.warmup):
        _spmv_rate(C, x, 0.2)
    # one step = a bounded sample (about 1 s) of applications with all host threads
    nsteps = max(1, min(args.steps, 10))
    t0 = time.perf_counter()
    rates = [_spmv_rate(C, x, 1.0)[0] for _ in range(nsteps)]
    wall = time.perf_counter() - t0
    value = sum(rates) / len(rates)
    b = cpu_port_baseline(args.order, seconds=1.0, n=512)
    b["value"] = value
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * wall / nsteps,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64",
        "data": "synthetic",
        "config": {"workload": f"IP-DG p={args.order} two-phase circle + Robin: CPU path of the reference "
                               "(COO assembly -> sparse -> A*x) restated in C (the reference is Julia; "
                               "no Julia runtime in this image)", "order": args.order,
                   "timed_steps": nsteps},
        "cpu_baseline": b,
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=50)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--order", type=int, default=1)
    ap.add_argument("--nx", type=int, default=0, help="cell columns per GPU")
    ap.add_argument("--ny", type=int, default=0)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--halo", default="peer", choices=["peer", "nccl"], help="N > 1: how the halo columns travel")
    ap.add_argument("--traffic-bytes", type=float, default=None,
                    help="dram bytes per launch from the committed ncu capture (profiles/)")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3)
    n = {1: 5000, 2: 3334, 3: 2500}[args.order]
    args.nx = args.nx or n
    args.ny = args.ny or n
    if args.traffic_bytes is None and args.order == 1 and (args.nx, args.ny) == (5000, 5000):
        args.traffic_bytes = PROFILED_TRAFFIC_BYTES
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


# dram__bytes_read.sum + dram__bytes_write.sum of the dominant kernel for the default
# workload, from profiles/ (updated whenever the kernel is re-profiled)
PROFILED_TRAFFIC_BYTES = 819.482624e6 + 758.8672e6  # profiles/r1_ipdg_p1_march_ncu.csv

if __name__ == "__main__":
    main()
