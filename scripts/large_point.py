"""The SURVEY §8(d) "x4 cells" point: IP apply at ~400 M dofs (two-phase circle + Robin), CUDA events, plus the
march kernel against the simple kernel on the same vector (size-independent parity check at the largest size)."""
import sys
import numpy as np
import torch
sys.path.insert(0, ".")
from stefanproblem_b200 import cutcelldg as cc


def run(order, n, reps=20):
    basis = cc.LagrangeTensorProductBasis(2, order)
    i, j = np.divmod(np.arange(n * n), n)
    phase = np.where(((i + 0.5) / n - 0.5) ** 2 + ((j + 0.5) / n - 0.5) ** 2 < 0.16, -1, 1).astype(np.int8)
    del i, j
    mesh = cc.UniformMesh([0, 0], [1, 1], [n, n], basis, phase=phase)
    op = cc.IPOperator(mesh, order, order + 1, 1.0, 2.0, 1e3 * n, 1.0, 0.5, "current", 255)
    x = torch.rand(op.ndofs, dtype=torch.float64, device="cuda") * 2 - 1
    y, y1 = torch.empty_like(x), torch.empty_like(x)
    for _ in range(3):
        op.apply(x, out=y, algo=0)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    e0.record()
    for _ in range(reps):
        op.apply(x, out=y, algo=0)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / reps
    op.apply(x, out=y1, algo=1)
    rel = ((y - y1).norm() / y1.norm()).item()
    print(f"IP p={order} {n}x{n} two-phase circle, {op.ndofs / 1e6:.1f} M dofs: {ms:.3f} ms  {op.ndofs / ms / 1e6:.1f} GDOF/s  "
          f"{16 * op.ndofs / ms / 1e6 / 6551 * 100:.1f}% of 6551 GB/s   march vs simple kernel rel. diff {rel:.2e}", flush=True)


if __name__ == "__main__":
    for order, n in ((1, 10000), (2, 6668), (3, 5000)):
        run(order, n)
