"""Per-CTA durations of the IP-DG march kernel (development hook STEFAN_DEV_CTA_TIMES of ipdg.cu):
usage: python scripts/cta_times.py ORDER TWO_PHASE(0|1)   -- on a GPU box."""
import os, sys, numpy as np, torch
sys.path.insert(0, ".")
order = int(sys.argv[1]); tp = int(sys.argv[2])
os.environ["STEFAN_DEV_CTA_TIMES"] = f"gpurun_out/cta_p{order}_tp{tp}.txt"
from stefanproblem_b200 import cutcelldg as cc
n = {1: 5000, 2: 3334, 3: 2500}[order]
basis = cc.LagrangeTensorProductBasis(2, order)
phase = None
if tp:
    i, j = np.divmod(np.arange(n * n), n)
    phase = np.where(((i + .5) / n - .5) ** 2 + ((j + .5) / n - .5) ** 2 < .16, -1, 1).astype(np.int8)
op = cc.IPOperator(cc.UniformMesh([0, 0], [1, 1], [n, n], basis, phase=phase), order, order + 1, 1.0, 2.0, 1e3 * n, 1.0, 0.5, "current", 255)
x = torch.rand(op.ndofs, dtype=torch.float64, device="cuda"); y = torch.empty_like(x)
for _ in range(4): op.apply(x, out=y)
t = np.loadtxt(os.environ["STEFAN_DEV_CTA_TIMES"])
t0 = t[:, 3].min(); dur = (t[:, 4] - t[:, 3]) / 1e3; end = (t[:, 4] - t0) / 1e3; start = (t[:, 3] - t0) / 1e3
print(f"p={order} two_phase={tp}: {len(t)} CTAs; duration us: min {dur.min():.0f} median {np.median(dur):.0f} max {dur.max():.0f}; "
      f"last end {end.max():.0f} us; start spread {start.max():.0f} us")
ncol = t[:, 2] - t[:, 1]
for g in np.unique(t[:, 0]).astype(int)[:40]:
    m = t[:, 0] == g
    print(f"  group {g:2d}: {m.sum():2d} tiles, columns/tile {ncol[m].mean():6.1f}, us/tile median {np.median(dur[m]):6.1f} max {dur[m].max():6.1f}, us/column {np.median(dur[m]/ncol[m]):.3f}")
