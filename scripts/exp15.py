import sys, time, os, numpy as np, torch
sys.path.insert(0, ".")
from stefanproblem_b200 import cutcelldg as cc
for order, n in ((2, 33), (1, 128), (1, 1024)):
    basis = cc.LagrangeTensorProductBasis(2, order)
    mesh = cc.UniformMesh([0, 0], [1, 1], [n, n], basis)
    op = cc.IPOperator(mesh, order, order + 1, 1.0, 1.0, 1e3 * n, 0.0, 0.0, "current", 255)
    b = torch.rand(op.ndofs, dtype=torch.float64, device="cuda")
    for mode in ("graph", "no graph"):
        if mode == "no graph": os.environ["STEFAN_CG_NO_GRAPH"] = "1"
        else: os.environ.pop("STEFAN_CG_NO_GRAPH", None)
        op.solve(b, tol=1e-10, maxiter=200)
        torch.cuda.synchronize(); t0 = time.perf_counter()
        x, info = op.solve(b, tol=1e-30, maxiter=500)
        torch.cuda.synchronize(); dt = time.perf_counter() - t0
        print(f"p={order} {n}x{n} ({op.ndofs} dofs) {mode}: {info['iterations']} iterations in {dt*1e3:.1f} ms = {dt/info['iterations']*1e6:.1f} us/iteration", flush=True)
