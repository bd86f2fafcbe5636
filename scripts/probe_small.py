"""Small-mesh regime of the march kernels (round 2): the 10 M-dof LDG meshes and the IP-DG column slabs of the
strong-scaling run (global 100 M-dof mesh cut into N slabs), timed on ONE GPU with local halo buffers.

  python scripts/probe_small.py time            # CUDA-event timings, plain launches and CUDA-graph replays
  python scripts/probe_small.py cta ldg ORDER   # per-CTA phase times (development hook STEFAN_DEV_CTA_TIMES)
  python scripts/probe_small.py cta ip ORDER N  # ... of the middle slab of an N-slab cut
"""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, ".")

IP_N = {1: 5000, 2: 3334, 3: 2500}
LDG_N = {1: 913, 2: 609, 3: 457}


def circle_phase(nx_glob, ny, c_lo, c_hi):
    """phase of the global circle mesh for columns [c_lo, c_hi) (clipped columns -> None)"""
    i, j = np.meshgrid(np.arange(c_lo, c_hi), np.arange(ny), indexing="ij")
    cx, cy = (i + 0.5) / nx_glob, (j + 0.5) / ny
    return np.where((cx - 0.5) ** 2 + (cy - 0.5) ** 2 < 0.16, -1, 1).astype(np.int8).reshape(-1)


def ip_slab(order, nslabs, rank=None):
    from stefanproblem_b200 import cutcelldg as cc
    n = IP_N[order]
    rank = nslabs // 2 if rank is None else rank
    c0, c1 = n * rank // nslabs, n * (rank + 1) // nslabs
    basis = cc.LagrangeTensorProductBasis(2, order)
    ph = circle_phase(n, n, c0, c1)
    plo = circle_phase(n, n, c0 - 1, c0) if c0 > 0 else None
    phi = circle_phase(n, n, c1, c1 + 1) if c1 < n else None
    mesh = cc.UniformMesh([c0 / n, 0], [(c1 - c0) / n, 1], [c1 - c0, n], basis, phase=ph)
    op = cc.IPOperator(mesh, order, order + 1, 1.0, 2.0, 1e3 * n, 1.0, 0.5, "current", 255, phase_lo=plo, phase_hi=phi)
    nf = (order + 1) ** 2
    x = torch.rand(op.ndofs, dtype=torch.float64, device="cuda") * 2 - 1
    xlo = torch.rand(n * nf, dtype=torch.float64, device="cuda") if plo is not None else None
    xhi = torch.rand(n * nf, dtype=torch.float64, device="cuda") if phi is not None else None
    return op, x, xlo, xhi


def ldg_op(order, n=None):
    from stefanproblem_b200 import cutcelldg as cc
    n = n or LDG_N[order]
    basis = cc.LagrangeTensorProductBasis(2, order)
    mesh = cc.UniformMesh([0, 0], [1, 1], [n, n], basis, phase=circle_phase(n, n, 0, n))
    op = cc.LDGOperator(mesh, order, order + 1, 1.0, 2.0, 0.0, 0.0, 1.0, (np.cos(np.pi / 4), np.sin(np.pi / 4)))
    x = torch.rand(op.ndofs, dtype=torch.float64, device="cuda") * 2 - 1
    return op, x


def time_fn(fn, reps):
    for _ in range(5):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    best = 1e30
    for _ in range(3):
        e0.record()
        for _ in range(reps):
            fn()
        e1.record()
        torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1) / reps)
    return best


def time_graph(fn, per_graph=20, reps=10):
    s = torch.cuda.Stream()
    with torch.cuda.stream(s):
        for _ in range(3):
            fn()
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g, stream=s):
        for _ in range(per_graph):
            fn()
    return time_fn(g.replay, reps) / per_graph


def report(tag, ndofs, ms, ms_g):
    g = ndofs / ms / 1e6
    gg = ndofs / ms_g / 1e6
    print(f"{tag}: {ms * 1e3:8.2f} us  {g:6.1f} GDOF/s ({16 * g / 65.51:5.1f}% of 6551)   graph: {ms_g * 1e3:8.2f} us  "
          f"{gg:6.1f} GDOF/s ({16 * gg / 65.51:5.1f}%)", flush=True)


def cmd_time(orders=(1, 2, 3)):
    for order in orders:
        for ns in (1, 2, 4, 8):
            op, x, xlo, xhi = ip_slab(order, ns)
            y = torch.empty_like(x)
            fn = lambda: op.apply(x, out=y, x_lo=xlo, x_hi=xhi)  # noqa: E731
            report(f"IP p={order} slab 1/{ns} ({op.mesh.nx}x{op.mesh.ny})", op.ndofs, time_fn(fn, 100), time_graph(fn))
            del op, x, y, xlo, xhi
    for order in orders:
        op, x = ldg_op(order)
        y = torch.empty_like(x)
        fn = lambda: op.apply(x, out=y)  # noqa: E731
        report(f"LDG p={order} {op.mesh.nx}^2", op.ndofs, time_fn(fn, 200), time_graph(fn))


def cmd_step(orders=(1, 2, 3)):
    """fused explicit step (24 B per dof: u, b read, unew written) against apply + euler_update (48 B)"""
    for order in orders:
        for ns in (1, 8):
            op, x, xlo, xhi = ip_slab(order, ns)
            y = torch.empty_like(x)
            b = torch.rand_like(x)
            fused = lambda: op.step(x, b, 1e-9, out=y, u_lo=xlo, u_hi=xhi)  # noqa: E731

            def unfused():
                op.apply(x, out=y, x_lo=xlo, x_hi=xhi)
                op.euler_update(x, y, b, 1e-9)
            for name, fn, bpd in (("fused step", fused, 24), ("apply + euler_update", unfused, 48)):
                ms = time_fn(fn, 50)
                g = op.ndofs / ms / 1e6
                print(f"IP p={order} slab 1/{ns} {name:22s}: {ms * 1e3:8.2f} us  {g:6.1f} GDOF/s  {bpd * g:6.0f} GB/s "
                      f"({bpd * g / 65.51:5.1f}% of 6551 at {bpd} B/dof)", flush=True)
            del op, x, y, b, xlo, xhi


def cta_report(path, tag):
    t = np.loadtxt(path)
    t0 = t[:, 3].min()
    st, pro, loop, end = ((t[:, k] - t0) / 1e3 for k in (3, 4, 5, 6))
    ncol = t[:, 2] - t[:, 1]
    ok = t[:, 4] > 0
    print(f"{tag}: {len(t)} CTAs, columns/tile {ncol.min():.0f}..{ncol.max():.0f} (mean {ncol.mean():.1f}); kernel span {end.max():.1f} us")
    q = lambda v: f"min {v.min():6.1f} med {np.median(v):6.1f} max {v.max():6.1f}"  # noqa: E731
    print(f"  start offset us   : {q(st)}")
    print(f"  prologue us       : {q((pro - st)[ok])}")
    print(f"  loop us           : {q((loop - pro)[ok])}   per column: {q(((loop - pro) / ncol)[ok])}")
    print(f"  tail us           : {q(end - loop)}")
    print(f"  CTA total us      : {q(end - st)};  end of last loop {loop.max():.1f}, of last CTA {end.max():.1f}")


def cmd_cta(kind, order, nslabs=1):
    path = f"gpurun_out/cta_{kind}_p{order}_n{nslabs}.txt"
    os.environ["STEFAN_DEV_CTA_TIMES"] = path
    if kind == "ldg":
        op, x = ldg_op(order)
        y = torch.empty_like(x)
        for _ in range(4):
            op.apply(x, out=y)
        tag = f"LDG p={order} {op.mesh.nx}^2"
    else:
        op, x, xlo, xhi = ip_slab(order, nslabs)
        y = torch.empty_like(x)
        for _ in range(4):
            op.apply(x, out=y, x_lo=xlo, x_hi=xhi)
        tag = f"IP p={order} slab 1/{nslabs} ({op.mesh.nx}x{op.mesh.ny})"
    torch.cuda.synchronize()
    cta_report(path, tag)


if __name__ == "__main__":
    os.makedirs("gpurun_out", exist_ok=True)
    if sys.argv[1] == "time":
        cmd_time(tuple(int(a) for a in sys.argv[2:]) or (1, 2, 3))
    elif sys.argv[1] == "step":
        cmd_step(tuple(int(a) for a in sys.argv[2:]) or (1, 2, 3))
    elif sys.argv[1] == "cta":
        cmd_cta(sys.argv[2], int(sys.argv[3]), int(sys.argv[4]) if len(sys.argv) > 4 else 1)
