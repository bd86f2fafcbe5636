"""One-process sweep of the host-side tiling knobs of the LDG march kernel at the BASELINE 10 M-dof meshes
(STEFAN_LDG_VARIANT: CTA shape, STEFAN_MIN_TILE_COLS: shortest tile; both read when the operator / its tiling is made)."""
import os
import sys
import numpy as np
import torch
sys.path.insert(0, ".")
from stefanproblem_b200 import cutcelldg as cc


def time_op(order, n, reps=100):
    basis = cc.LagrangeTensorProductBasis(2, order)
    c = (np.arange(n) + 0.5) / n - 0.5
    phase = np.where(c[:, None] ** 2 + c[None, :] ** 2 < 0.16, -1, 1).astype(np.int8).reshape(-1)
    mesh = cc.UniformMesh([0, 0], [1, 1], [n, n], basis, phase=phase)
    op = cc.LDGOperator(mesh, order, order + 1, 1.0, 2.0, 0.0, 0.0, 1.0, (np.cos(np.pi / 4), np.sin(np.pi / 4)))
    x = torch.rand(op.ndofs, dtype=torch.float64, device="cuda") * 2 - 1
    y = torch.empty_like(x)
    for _ in range(5):
        op.apply(x, out=y, algo=2)
    best = 1e9
    for _ in range(3):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize()
        e0.record()
        for _ in range(reps):
            op.apply(x, out=y, algo=2)
        e1.record()
        torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1) / reps)
    return op.ndofs / best / 1e6, best * 1e3


for order, n in ((1, 913), (2, 609), (3, 457)):
    for variant in ("0", "1"):
        row = []
        for cols in ("2", "3", "4", "6", "8", "12", "16", "24"):
            os.environ["STEFAN_LDG_VARIANT"] = variant
            os.environ["STEFAN_MIN_TILE_COLS"] = cols
            g, us = time_op(order, n)
            row.append(f"{cols}:{g:.0f}")
        print(f"LDG p={order} {n}^2 variant {variant}  min_tile_cols:GDOF/s  " + "  ".join(row), flush=True)
