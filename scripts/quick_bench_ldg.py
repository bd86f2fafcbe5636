"""Development timing of the LDG apply kernel (CUDA events)."""
import sys
import numpy as np
import torch
sys.path.insert(0, ".")
from stefanproblem_b200 import cutcelldg as cc

def run(order, n, reps=20):
    basis = cc.LagrangeTensorProductBasis(2, order)
    mesh = cc.UniformMesh([0, 0], [1, 1], [n, n], basis)
    op = cc.LDGOperator(mesh, order, order + 1, 1.0, 1.0, 0.0, 0.0, 1.0, (np.cos(np.pi / 4), np.sin(np.pi / 4)))
    x = torch.rand(op.ndofs, dtype=torch.float64, device="cuda") * 2 - 1
    y = torch.empty_like(x)
    for _ in range(3):
        op.apply(x, out=y)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        op.apply(x, out=y)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / reps
    g = op.ndofs / ms / 1e6
    print(f"LDG p={order} {n}x{n} ({op.ndofs/1e6:.1f} M dofs): {ms:.3f} ms  {g:.1f} GDOF/s  {16*g:.0f} GB/s ({16*g/6551*100:.1f}% of 6551)", flush=True)

if __name__ == "__main__":
    for order, n in ((1, 913), (2, 609), (3, 457), (1, 2887), (2, 1925)):
        run(order, n)
