#!/usr/bin/env python
"""bench.py -- fp64 DOF-updates/s of the matrix-free IP-DG heat operator y = A x.

Workload (BASELINE.json `configs[4]`): uniform 2-D mesh of Q_p cells,
two phases (circle of radius 0.4, k1 = 1, k2 = 2), mesh-aligned interface with
flux + penalty + Robin terms (lambda = 1), Dirichlet boundary, pen = 1e3 * nx.
Parity of this workload: oracle extrapolation -- the reference's
face-level functions on the faces between unlike phases, driven by the oracle's own loop.

`--gpus N` is STRONG scaling (default): the global mesh is fixed (p = 1: 5000 x 5000 cells =
100 M dofs) and rank r owns the r-th slab of nx/N cell columns.  The slab vectors live in
CUDA-IPC peer memory and every rank's apply kernel reads its neighbours' edge columns over
NVLink itself (one launch per application, ordered by epoch flags that the same kernel
publishes and waits for); the K timed steps are one CUDA-graph replay per rank.  The line also
carries `weak` (5000 x 5000 cells PER GPU, the round-1 measurement), `parity_max_rel` (fused
peer halos + march kernel against NCCL halos + the plain one-thread-per-cell kernel, max over
ranks, asserted <= 1e-13) and `step_loop` (the explicit time loop: fused step kernel + interface
sums + 4-double NCCL all-reduce per step).  `--halo nccl` exchanges the columns with NCCL
send/recv on a side stream, overlapped with the interior columns.

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
STEP_BYTES_PER_DOF = 24  # fused explicit step: u and b read, unew written
PARITY_NOTE = "oracle extrapolation (SURVEY 8(d)(ii)): reference face-level functions on unlike-phase faces, oracle's own driver loop"


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
            self._stop.wait(0.05)

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


class SlabProblem:
    """One rank's part of a global nxg x ny mesh: operator, slab vector(s), result vector."""

    def __init__(self, order, nxg, ny, rank, world, want_peer, nbuf=1, bounds=None):
        import torch
        import torch.distributed as dist
        from stefanproblem_b200.distributed import DistributedIPOperator, PeerSlab
        self.rank, self.world, self.order = rank, world, order
        self.dop = DistributedIPOperator(nxg, ny, (1.0, 1.0), order, order + 1, 1.0, 2.0, 1e3 * nxg, 1.0, 0.5,
                                         lambda i: column_phase(nxg, ny, i), rank, world, bounds=bounds)
        self.op = self.dop.op
        self.ndofs = self.op.ndofs
        gen = torch.Generator(device="cuda")
        gen.manual_seed(1234 + rank)
        self.x = torch.rand(self.ndofs, dtype=torch.float64, device="cuda", generator=gen) * 2 - 1
        self.y = torch.empty_like(self.x)
        self.slab = None
        self.peer = False
        if world > 1 and want_peer:
            # collective; if CUDA IPC is not available to some rank (container policy), every rank
            # falls back to the NCCL halos together and the line says so
            failed = 0
            try:
                self.slab = PeerSlab(self.ndofs, rank, world, nbuf=nbuf)
            except Exception as exc:  # noqa: BLE001
                failed = 1
                sys.stderr.write(f"rank {rank}: peer memory unavailable ({exc}); falling back to --halo nccl\n")
            flag = torch.tensor([failed], dtype=torch.int32, device="cuda")
            dist.all_reduce(flag, op=dist.ReduceOp.MAX)
            if int(flag.item()):
                if self.slab is not None:
                    self.slab.close()
                self.slab = None
            else:
                self.slab.x.copy_(self.x)
                self.x = self.slab.x
                self.peer = True

    chain_ok = True    # --sync ready: gate on the neighbours' ready flag of the same epoch even in the timed loop
    push = True        # --no-push: poll the neighbours' flags over NVLink instead of local mailboxes

    def step(self, device_epoch=False, chained=False):
        from stefanproblem_b200.distributed import apply_peer_fused
        chained = chained and SlabProblem.chain_ok
        if self.peer:
            # ONE launch: the kernel publishes my epoch, its edge tiles wait for the neighbours' flags before
            # they read their edge columns over NVLink, its last CTA releases them.  chained (the timed loop:
            # x does not change between its applications): the gate is the neighbours' done value of the
            # previous epoch, posted into this rank's own memory
            apply_peer_fused(self.dop, self.slab, self.y, device_epoch=device_epoch, chained=chained,
                             push=SlabProblem.push)
        elif self.world > 1:
            # halo trace exchange (NCCL, side stream) overlapped with the interior columns
            self.dop.apply(self.x, self.y)
        else:
            self.op.apply(self.x, out=self.y)

    def close(self):
        if self.slab is not None:
            self.slab.close()
            self.slab = None

    def kernel_seconds(self, reps=30):
        """This rank's apply kernel alone (no flags, local halo buffers), CUDA events, plain launches."""
        import torch
        h = self.dop.halo
        for _ in range(3):
            self.op.apply(self.x, out=self.y, x_lo=h.x_lo, x_hi=h.x_hi)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize()
        e0.record()
        for _ in range(reps):
            self.op.apply(self.x, out=self.y, x_lo=h.x_lo, x_hi=h.x_hi)
        e1.record()
        torch.cuda.synchronize()
        return e0.elapsed_time(e1) * 1e-3 / reps


def barrier(world):
    import torch
    import torch.distributed as dist
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()


def max_over_ranks(v, world):
    import torch
    import torch.distributed as dist
    if world == 1:
        return float(v)
    t = torch.tensor([v], dtype=torch.float64, device="cuda")
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


def timed_steps(prob, steps, warmup, use_graph, clk_gpu=None):
    """W untimed steps, then exactly `steps` steps between barrier + synchronize, CUDA events, max over ranks.
    use_graph: the K steps are ONE CUDA-graph replay (the peer epochs live on the device)."""
    import torch
    world = prob.world
    graph = None
    dev_epoch = bool(use_graph and prob.peer)
    stream = torch.cuda.Stream()
    if dev_epoch:
        torch.cuda.synchronize()
        prob.slab.to_device_epochs()
        torch.cuda.synchronize()
    with torch.cuda.stream(stream):
        for _ in range(warmup):
            prob.step(device_epoch=dev_epoch, chained=True)
        stream.synchronize()
        if use_graph and (prob.peer or world == 1):
            try:
                graph = torch.cuda.CUDAGraph()
                with torch.cuda.graph(graph, stream=stream):
                    for _ in range(steps):
                        prob.step(device_epoch=dev_epoch, chained=True)
            except Exception as exc:  # noqa: BLE001
                sys.stderr.write(f"rank {prob.rank}: CUDA-graph capture failed ({exc}); timing plain launches\n")
                graph = None
    # capture must have worked on every rank (a replay launches `steps` kernels, the epochs must agree)
    ok = max_over_ranks(0.0 if graph is not None else 1.0, world) == 0.0
    if not ok:
        graph = None
        if dev_epoch:   # back to host-side epochs: resynchronise them with the device counter
            torch.cuda.synchronize()
            prob.slab.to_host_epochs()
            dev_epoch = False
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    sampler = ClockSampler(clk_gpu) if clk_gpu is not None else None
    if sampler:
        sampler.__enter__()
    with torch.cuda.stream(stream):
        if graph is not None:   # one untimed replay: uploads the graph
            graph.replay()
        barrier(world)
        e0.record(stream)
        if graph is not None:
            graph.replay()
        else:
            for _ in range(steps):
                prob.step(device_epoch=dev_epoch, chained=True)
        e1.record(stream)
        barrier(world)
    if sampler:
        # keep sampling for a moment if the region was shorter than one nvidia-smi call
        t_end = time.time() + 0.3
        while not sampler.samples and time.time() < t_end:
            time.sleep(0.02)
        sampler.__exit__()
    ms = max_over_ranks(e0.elapsed_time(e1), world)
    if dev_epoch:
        prob.slab.to_host_epochs()
    return ms / steps, graph is not None, (sampler.summary() if sampler else None)


def parity_check(prob):
    """Fused peer halos + march kernel against NCCL halos + the plain one-thread-per-cell kernel on this rank's slab
    (N = 1: march against plain).  Returns (max over ranks of the relative difference, checksum of y)."""
    import torch
    import torch.distributed as dist
    world = prob.world
    prob.step()
    y1 = prob.y.clone()
    if world > 1:
        h = prob.dop.halo
        h.exchange(prob.x)
        y2 = prob.op.apply(prob.x, x_lo=h.x_lo, x_hi=h.x_hi, algo=prob.op.SIMPLE)
    else:
        y2 = prob.op.apply(prob.x, algo=prob.op.SIMPLE)
    rel = float((y1 - y2).norm() / y2.norm())
    chk = torch.stack([y1.sum(), y2.sum()])
    if world > 1:
        dist.all_reduce(chk, op=dist.ReduceOp.SUM)
        if prob.peer:
            prob.slab.check()   # raises on every rank if a flag wait timed out anywhere
    return max_over_ranks(rel, world), float(chk[0].item()), float(chk[1].item())


def time_step_loop(order, nxg, ny, rank, world, steps, warmup, bounds=None):
    """The explicit time loop that drives the front: per step ONE fused kernel (unew = u + dt Minv (b - A u), the
    neighbours' edge columns of u read over NVLink), then the interface sums of unew and their 4-double all-reduce."""
    import torch
    import torch.distributed as dist
    from stefanproblem_b200.distributed import interface_sums_peer, step_peer_fused
    prob = SlabProblem(order, nxg, ny, rank, world, want_peer=True, nbuf=2, bounds=bounds)
    if world > 1 and not prob.peer:
        prob.close()
        return None
    b = torch.rand_like(prob.x)
    dt = 1e-13 * (5000.0 / nxg) ** 2   # (stable explicit step: dt * lambda_max(Minv A) < 2 with pen = 1e3 nx)
    if world > 1:
        bufs = prob.slab.bufs
    else:
        bufs = [prob.x, torch.empty_like(prob.x)]
    sums = torch.zeros(4, dtype=torch.float64, device="cuda")

    def one(k):
        src, dst = k & 1, (k + 1) & 1
        if world > 1:
            step_peer_fused(prob.dop, prob.slab, src, dst, b, dt)
            interface_sums_peer(prob.dop, prob.slab, dst, sums)
            dist.all_reduce(sums, op=dist.ReduceOp.SUM)
        else:
            prob.op.step(bufs[src], b, dt, out=bufs[dst])
            prob.op.interface_sums(bufs[dst], out=sums)

    barrier(world)          # chained halo gate: every rank's initial vector is final before the first step
    for k in range(warmup):
        one(k)
    barrier(world)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for k in range(warmup, warmup + steps):
        one(k)
    e1.record()
    barrier(world)
    ms = max_over_ranks(e0.elapsed_time(e1), world) / steps
    if world > 1:
        prob.slab.check()
    ndofs = prob.ndofs
    prob.close()
    tot = ndofs
    if world > 1:
        t = torch.tensor([ndofs], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.SUM)
        tot = float(t.item())
    return {"ms_per_step": ms, "value": tot / (ms * 1e-3), "unit": "DOF-updates/s (time steps)",
            "per_step": "stefan_ipdg_step (fused apply + mass inverse + update, peer halos) + stefan_ipdg_interface_sums "
                        "+ 4-double NCCL all-reduce; plain launches",
            "algorithmic_bytes_per_dof": STEP_BYTES_PER_DOF,
            "interface_sums_last": [float(v) for v in sums.cpu()]}


def run_ours(args):
    import torch
    import torch.distributed as dist

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if args.gpus != world:
        if world == 1 and args.gpus > 1:
            raise SystemExit("launch with torch.distributed.run --nproc-per-node N for --gpus N")
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))

    p, n_x, ny = args.order, args.nx, args.ny
    strong = args.scaling == "strong"
    nxg = n_x if strong else n_x * world
    want_peer = args.halo == "peer"
    prob = SlabProblem(p, nxg, ny, rank, world, want_peer)
    # optional cost-balanced slabs (--balance): the slabs' per-column costs differ (exception cells of the interface
    # and of the domain boundary are not spread evenly over the columns), and every step waits for the slowest
    # rank.  One measurement of every rank's kernel, one re-partition.
    bounds, per_rank_ms = None, None
    if world > 1 and strong and args.balance:
        from stefanproblem_b200.distributed import balanced_bounds
        tw = torch.tensor([prob.kernel_seconds(), float(prob.dop.nxl)], dtype=torch.float64, device="cuda")
        allv = [torch.zeros_like(tw) for _ in range(world)]
        dist.all_gather(allv, tw)
        secs, wid = [float(v[0]) for v in allv], [int(v[1]) for v in allv]
        per_rank_ms = [s_ * 1e3 for s_ in secs]
        bounds = balanced_bounds(nxg, secs, wid, fixed_seconds=8e-6)
        prob.close()
        del prob
        prob = SlabProblem(p, nxg, ny, rank, world, want_peer, bounds=bounds)
    peer = prob.peer
    ndofs_local = prob.ndofs
    t = torch.tensor([ndofs_local], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.SUM)
    total_dofs = int(t.item())

    # ---- parity before anything is timed (driver-visible: the line carries the number) ----
    parity_rel, chk_fused, chk_plain = parity_check(prob)
    if parity_rel > 1e-13:
        raise SystemExit(f"rank {rank}: parity check failed: relative difference {parity_rel:.3e} > 1e-13")

    # ---- the timed region ----
    use_graph = not args.no_graph
    ms_per_step, graphed, clocks = timed_steps(prob, args.steps, args.warmup, use_graph, clk_gpu=local)
    value = total_dofs / (ms_per_step * 1e-3)

    # kernel-only duration of this rank's slab for the roofline (no flags, static x), CUDA events on the launching stream
    barrier(world)
    x_lo, x_hi = prob.dop.halo.x_lo, prob.dop.halo.x_hi
    if world > 1:
        prob.dop.halo.exchange(prob.x)
    for _ in range(3):
        prob.op.apply(prob.x, out=prob.y, x_lo=x_lo, x_hi=x_hi)
    k0, k1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    k0.record()
    for _ in range(args.steps):
        prob.op.apply(prob.x, out=prob.y, x_lo=x_lo, x_hi=x_hi)
    k1.record()
    torch.cuda.synchronize()
    kernel_ms = k0.elapsed_time(k1) / args.steps
    peak, peak_src = measured_hbm_peak()
    achieved = ALGO_BYTES_PER_DOF * ndofs_local / (kernel_ms * 1e-3) / 1e9

    # ---- end to end through the C ABI with host buffers ----
    e2e = None
    if world == 1:
        xh = torch.empty(ndofs_local, dtype=torch.float64).pin_memory()
        xh.copy_(prob.x.cpu())
        yh = torch.empty(ndofs_local, dtype=torch.float64).pin_memory()
        n_e2e = max(2, min(args.steps, 10))
        prob.op.apply_host(xh, yh)
        t0 = time.perf_counter()
        for _ in range(n_e2e):
            prob.op.apply_host(xh, yh)
        dt = (time.perf_counter() - t0) / n_e2e
        e2e = {"value": ndofs_local / dt, "unit": UNIT, "h2d_bytes_per_step": ndofs_local * 8,
               "d2h_bytes_per_step": ndofs_local * 8, "steps": n_e2e,
               "api": "stefan_ipdg_apply_host (pinned host buffers, 32 pipelined slabs)",
               "bound": "PCIe (about 45 GB/s each way)"}
        del xh, yh
    elif peer:
        from stefanproblem_b200.distributed import apply_peer_fused
        # end to end at N GPUs: every rank moves its slab over its own PCIe link (pinned host
        # buffers, H2D of x and D2H of y inside the timed region) around the one fused launch
        slab, dop = prob.slab, prob.dop
        xh = torch.empty(ndofs_local, dtype=torch.float64).pin_memory()
        xh.copy_(prob.x.cpu())
        yh = torch.empty(ndofs_local, dtype=torch.float64).pin_memory()
        n_e2e = max(2, min(args.steps, 10))
        ybuf = [prob.y, torch.empty_like(prob.y)]
        d2h, done, copied = torch.cuda.Stream(), [torch.cuda.Event(), torch.cuda.Event()], [None, None]
        work = torch.cuda.Stream()  # not the legacy default stream, which serialises with other streams' copies

        def e2e_step(k):
            # the result of step k leaves on a second stream while step k+1's input arrives
            # (PCIe is full duplex); two result buffers
            bb = k & 1
            slab.begin_write()              # the neighbours are done with my previous vector
            slab.x.copy_(xh, non_blocking=True)
            if copied[bb] is not None:
                torch.cuda.current_stream().wait_event(copied[bb])   # ybuf[b] has left the device
            apply_peer_fused(dop, slab, ybuf[bb])
            done[bb].record()
            with torch.cuda.stream(d2h):
                d2h.wait_event(done[bb])
                yh.copy_(ybuf[bb], non_blocking=True)
                copied[bb] = torch.cuda.Event()
                copied[bb].record(d2h)

        barrier(world)
        with torch.cuda.stream(work):
            e2e_step(0)
            barrier(world)
            t0 = time.perf_counter()
            for k in range(n_e2e):
                e2e_step(k + 1)
            barrier(world)
        dt = max_over_ranks((time.perf_counter() - t0) / n_e2e, world)
        slab.check()
        e2e = {"value": total_dofs / dt, "unit": UNIT, "h2d_bytes_per_step": total_dofs * 8,
               "d2h_bytes_per_step": total_dofs * 8, "steps": n_e2e,
               "api": "PeerSlab + stefan_ipdg_apply_peer per rank (pinned host buffers, each rank over its own PCIe link)",
               "bound": "host memory system: every rank's slab crosses PCIe twice per step (all GPUs of a box share the "
                        "host's memory controllers), so this number does not scale with N"}
        del xh, yh
    prob.close()

    # ---- the other scaling mode and the time loop, as extra fields (not the headline) ----
    other = None
    if world > 1 and not args.no_extras:
        nxg2 = n_x * world if strong else n_x
        p2 = SlabProblem(p, nxg2, ny, rank, world, want_peer)
        t2 = torch.tensor([p2.ndofs], dtype=torch.float64, device="cuda")
        dist.all_reduce(t2, op=dist.ReduceOp.SUM)
        ms2, g2, _ = timed_steps(p2, max(10, args.steps // 2), args.warmup, use_graph)
        other = {"scaling": "weak" if strong else "strong", "global_cells": [nxg2, ny], "global_dofs": int(t2.item()),
                 "ms_per_step": ms2, "value": float(t2.item()) / (ms2 * 1e-3), "unit": UNIT, "cuda_graph": g2}
        p2.close()
    step_loop = None
    if not args.no_extras:
        step_loop = time_step_loop(p, nxg, ny, rank, world, max(10, args.steps // 2), args.warmup, bounds=bounds)

    cpu_baseline = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        cpu_baseline = cpu_port_baseline(p, seconds=6.0)

    if rank == 0:
        # N > 1, nccl halos: interior + two edge columns; peer halos: the one apply kernel
        launches = args.steps * (3 if (world > 1 and not peer) else 1)
        if peer:
            partition = (f"{world} column slabs of the fixed global mesh (strong scaling)" if strong else
                         f"{world} column slabs, {n_x} columns per GPU (weak scaling)")
            partition += ("; neighbours' edge columns read over NVLink from CUDA-IPC peer memory inside the apply kernel, "
                          "which also publishes / waits for the epoch flags: one launch per step"
                          + (", the K steps are one CUDA-graph replay per rank" if graphed else ""))
        elif world > 1:
            partition = (f"{world} column slabs ({'strong' if strong else 'weak'} scaling); 1 halo cell column per "
                         "neighbour over NCCL send/recv, overlapped with the interior columns")
        else:
            partition = "1 slab (no halo)" + ("; the K steps are one CUDA-graph replay" if graphed else "")
        nxl = prob.dop.nxl
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms_per_step, "higher_is_better": True,
            "scaling": args.scaling, "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {
                "workload": f"IP-DG p={p} matrix-free apply, two-phase circle + Robin interface, global mesh "
                            f"{nxg}x{ny} cells ({total_dofs / 1e6:.1f} M dofs), {nxl}x{ny} cells on rank 0",
                "order": p, "cells_rank0": [nxl, ny], "global_cells": [nxg, ny],
                "global_dofs": total_dofs, "partition": partition, "cuda_graph": graphed,
                "parity_of_workload": PARITY_NOTE,
                "l2": ("inputs larger than L2: x and y of a rank are "
                       f"{ndofs_local * 8 / 1e6:.0f} MB each (L2: 126 MB)"),
            },
            "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s",
                         "frac": achieved / peak, "traffic": args.traffic_bytes,
                         "peak_source": peak_src, "kernel_ms": kernel_ms,
                         "algorithmic_bytes_per_dof": ALGO_BYTES_PER_DOF,
                         "kernel": f"ipdg_apply_march<{p}, 0> on rank 0's slab ({ndofs_local / 1e6:.1f} M dofs), plain launches"},
            "clocks": clocks,
            "gpu_launches": launches,
            "slab_bounds": bounds, "per_rank_kernel_ms_equal_slabs": per_rank_ms,
            "parity_max_rel": parity_rel,
            "parity": {"max_rel": parity_rel, "tolerance": 1e-13,
                       "what": ("fused peer halos + march kernel vs NCCL halos + plain per-cell kernel, every rank's slab"
                                if world > 1 else "march kernel vs plain per-cell kernel"),
                       "checksum_fused": chk_fused, "checksum_plain": chk_plain},
        }
        if other is not None:
            line[other["scaling"]] = other
        if step_loop is not None:
            line["step_loop"] = step_loop
        if e2e is not None:
            line["e2e"] = e2e
        if cpu_baseline is not None:
            line["cpu_baseline"] = cpu_baseline
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


_REF_CACHE = {}


def reference_sample_cells():
    """Cells per side of the CPU sample: the largest of 1024 / 768 / 512 whose COO -> CSR path (about 13 KB per
    p = 1 cell at its peak) fits a quarter of the host's free memory; 1024^2 cells = 4.2 M dofs, CSR about 1 GB
    (well outside the caches)."""
    try:
        import psutil
        avail = psutil.virtual_memory().available
    except Exception:  # noqa: BLE001
        avail = 16 << 30
    for n in (1024, 768, 512):
        if n * n * 13e3 < 0.25 * avail:
            return n
    return 512


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


def _spmv_time(C, x, reps, threads=0):
    import numpy as np
    y = np.empty_like(x)
    t0 = time.perf_counter()
    for _ in range(reps):
        C.spmv(x, y, threads)
    return (time.perf_counter() - t0) / reps


def cpu_port_baseline(order, seconds=6.0, n=None, threads=0):
    """The reference's CPU path (COO assembly -> sparse -> A*x) as restated in C, on a
    bounded sample: n x n cells, same physics as the GPU workload, all host threads."""
    from oracle.cport import CIPOracle
    n = n or (reference_sample_cells() if order == 1 else 512)
    C, x, t_asm = _cpu_port(order, n)
    t1 = _spmv_time(C, x, 2, threads)
    reps = max(3, int(seconds / max(t1, 1e-6)))
    t = _spmv_time(C, x, reps, threads)
    cores = CIPOracle.max_threads() if threads <= 0 else threads
    return {"value": C.nrows / t, "unit": UNIT, "cores": cores, "kind": "port",
            "sample": f"{n}x{n} cells p={order} ({C.nrows / 1e6:.2f} M dofs, CSR {C.nnz * 12 / 1e9:.2f} GB): CSR SpMV of the "
                      f"oracle-assembled matrix, {reps} applications of {t * 1e3:.2f} ms; assembly (COO + compress) "
                      f"took {t_asm:.1f} s once and is not in the rate",
            "assembly_seconds": t_asm, "nnz": C.nnz, "ms_per_apply": t * 1e3}


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    n = reference_sample_cells() if args.order == 1 else 512
    C, x, t_asm = _cpu_port(args.order, n)
    for _ in range(args.warmup):
        _spmv_time(C, x, 1)
    # one step = one application (CSR SpMV with all host threads) of the bounded sample; K steps timed
    steps = args.steps
    t = _spmv_time(C, x, steps)
    value = C.nrows / t
    from oracle.cport import CIPOracle
    b = {"value": value, "unit": UNIT, "cores": CIPOracle.max_threads(), "kind": "port",
         "sample": f"{n}x{n} cells p={args.order} ({C.nrows / 1e6:.2f} M dofs, CSR {C.nnz * 12 / 1e9:.2f} GB): {steps} CSR SpMV "
                   f"applications of the oracle-assembled matrix with all host threads; assembly {t_asm:.1f} s, not timed",
         "assembly_seconds": t_asm, "nnz": C.nnz}
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus,
        "steps": steps, "warmup": args.warmup, "ms_per_step": 1e3 * t,
        "higher_is_better": True, "scaling": args.scaling, "vs_baseline": None, "dtype": "f64",
        "data": "synthetic",
        "config": {"workload": f"IP-DG p={args.order} two-phase circle + Robin: CPU path of the reference "
                               "(COO assembly -> sparse -> A*x) restated in C (the reference is Julia; "
                               f"no Julia runtime in this image); bounded sample {n}x{n} cells, one step = one A*x",
                   "order": args.order, "sample_cells": [n, n], "sample_dofs": C.nrows,
                   "parity_of_workload": PARITY_NOTE},
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
    ap.add_argument("--scaling", default="strong", choices=["strong", "weak"],
                    help="strong: --nx is the GLOBAL column count; weak: columns per GPU")
    ap.add_argument("--nx", type=int, default=0, help="cell columns (global if strong, per GPU if weak)")
    ap.add_argument("--ny", type=int, default=0)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-graph", action="store_true", help="time plain launches instead of one CUDA-graph replay")
    ap.add_argument("--no-extras", action="store_true", help="skip the weak-scaling and time-loop fields")
    ap.add_argument("--sync", default="chained", choices=["chained", "ready"],
                    help="halo gate of the timed loop: the neighbours' done value of the previous epoch (x is static), "
                         "or their ready flag of the same epoch")
    ap.add_argument("--balance", action="store_true",
                    help="N > 1, strong scaling: re-partition the columns by one measurement of every rank's kernel "
                         "(measured at 8 GPUs: the ranks' kernels differ by 1.4 us of 41 and the re-cut tiles cost more "
                         "than the balance gains -- 46.2 vs 44.5 us per step -- so equal shares are the default)")
    ap.add_argument("--no-push", action="store_true", help="poll the neighbours' flags over NVLink (no mailboxes)")
    ap.add_argument("--halo", default="peer", choices=["peer", "nccl"], help="N > 1: how the halo columns travel")
    ap.add_argument("--traffic-bytes", type=float, default=None,
                    help="dram bytes per launch from the committed ncu capture (profiles/)")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3)
    args.steps = max(args.steps, 1)
    n = {1: 5000, 2: 3334, 3: 2500}[args.order]
    args.nx = args.nx or n
    args.ny = args.ny or n
    if args.traffic_bytes is None and args.order == 1 and (args.nx, args.ny) == (5000, 5000) and args.gpus == 1:
        args.traffic_bytes = PROFILED_TRAFFIC_BYTES
    if args.impl == "reference":
        run_reference(args)
    else:
        SlabProblem.chain_ok = args.sync == "chained"
        SlabProblem.push = not args.no_push
        run_ours(args)


# dram__bytes_read.sum + dram__bytes_write.sum of the dominant kernel for the default
# workload, from profiles/ (updated whenever the kernel is re-profiled)
PROFILED_TRAFFIC_BYTES = 812.808704e6 + 755.800832e6  # profiles/r2_ipdg_p1_march_ncu.csv

if __name__ == "__main__":
    main()
