import sys, time, torch, numpy as np
sys.path.insert(0, ".")
from stefanproblem_b200 import cutcelldg as cc
n = 5000
basis = cc.LagrangeTensorProductBasis(2, 1)
i, j = np.divmod(np.arange(n * n), n)
phase = np.where(((i + .5) / n - .5) ** 2 + ((j + .5) / n - .5) ** 2 < .16, -1, 1).astype(np.int8)
op = cc.IPOperator(cc.UniformMesh([0, 0], [1, 1], [n, n], basis, phase=phase), 1, 2, 1.0, 2.0, 1e3 * n, 1.0, 0.5, "current", 255)
xh = torch.rand(op.ndofs, dtype=torch.float64).pin_memory(); yh = torch.empty_like(xh).pin_memory()
for ns in (4, 8, 16, 32):
    op.apply_host(xh, yh, nslabs=ns)
    t0 = time.perf_counter()
    for _ in range(5): op.apply_host(xh, yh, nslabs=ns)
    dt = (time.perf_counter() - t0) / 5
    print(f"nslabs={ns}: {dt*1e3:.2f} ms  {op.ndofs/dt/1e9:.2f} GDOF/s  ({1.6e9/dt/1e9:.1f} GB/s PCIe both ways)", flush=True)
# plain copies for reference
xd = torch.empty(op.ndofs, dtype=torch.float64, device="cuda")
torch.cuda.synchronize(); t0 = time.perf_counter()
for _ in range(5): xd.copy_(xh, non_blocking=True)
torch.cuda.synchronize(); dt = (time.perf_counter() - t0) / 5
print(f"H2D alone: {0.8e9/dt/1e9:.1f} GB/s")
torch.cuda.synchronize(); t0 = time.perf_counter()
for _ in range(5): yh.copy_(xd, non_blocking=True)
torch.cuda.synchronize(); dt = (time.perf_counter() - t0) / 5
print(f"D2H alone: {0.8e9/dt/1e9:.1f} GB/s")
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
xd2 = torch.empty_like(xd)
torch.cuda.synchronize(); t0 = time.perf_counter()
for _ in range(5):
    with torch.cuda.stream(s1): xd.copy_(xh, non_blocking=True)
    with torch.cuda.stream(s2): yh.copy_(xd2, non_blocking=True)
torch.cuda.synchronize(); dt = (time.perf_counter() - t0) / 5
print(f"H2D + D2H concurrently: {0.8e9/dt/1e9:.1f} GB/s each")
