"""Development timing of the IP-DG apply kernels (CUDA events, inputs > L2)."""
import sys
import numpy as np
import torch
sys.path.insert(0, ".")
from stefanproblem_b200 import cutcelldg as cc

PATTERN = "circle"


def run(order, nx, ny, two_phase, algo, reps=20):
    basis = cc.LagrangeTensorProductBasis(2, order)
    phase = None
    if two_phase:
        i, j = np.divmod(np.arange(nx * ny), ny)
        cx, cy = (i + 0.5) / nx, (j + 0.5) / ny
        inside = {"circle": (cx - 0.5) ** 2 + (cy - 0.5) ** 2 < 0.16, "vsplit": cx > 0.5, "hsplit": cy > 0.5,
                  "allneg": cx > -1, "hstripes": (j // 300) % 2 == 1}[PATTERN]
        phase = np.where(inside, -1, 1).astype(np.int8)
    mesh = cc.UniformMesh([0, 0], [1, 1], [nx, ny], basis, phase=phase)
    op = cc.IPOperator(mesh, order, order + 1, 1.0, 2.0, 1e3 * nx, 1.0, 0.5, "current", 255)
    x = torch.rand(op.ndofs, dtype=torch.float64, device="cuda") * 2 - 1
    y = torch.empty_like(x)
    for _ in range(3):
        op.apply(x, out=y, algo=algo)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        op.apply(x, out=y, algo=algo)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / reps
    gdofs = op.ndofs / ms / 1e6
    print(f"p={order} {nx}x{ny} two_phase={two_phase}{'/' + PATTERN if two_phase else ''} algo={algo}: {ms:.3f} ms  {gdofs:.1f} GDOF/s  "
          f"{16 * gdofs:.0f} GB/s  ({16 * gdofs / 6551 * 100:.1f}% of 6551)", flush=True)

def run_ldg(order, n, two_phase, reps=20, algo=0):
    basis = cc.LagrangeTensorProductBasis(2, order)
    phase = None
    if two_phase:
        i, j = np.divmod(np.arange(n * n), n)
        cx, cy = (i + 0.5) / n, (j + 0.5) / n
        phase = np.where((cx - 0.5) ** 2 + (cy - 0.5) ** 2 < 0.16, -1, 1).astype(np.int8)
    mesh = cc.UniformMesh([0, 0], [1, 1], [n, n], basis, phase=phase)
    op = cc.LDGOperator(mesh, order, order + 1, 1.0, 2.0, 0.0, 0.0, 1.0, (np.cos(np.pi / 4), np.sin(np.pi / 4)))
    x = torch.rand(op.ndofs, dtype=torch.float64, device="cuda") * 2 - 1
    y = torch.empty_like(x)
    for _ in range(3):
        op.apply(x, out=y, algo=algo)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        op.apply(x, out=y, algo=algo)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / reps
    gdofs = op.ndofs / ms / 1e6
    print(f"LDG p={order} {n}x{n} two_phase={two_phase} algo={algo}: {ms:.4f} ms  {gdofs:.1f} GDOF/s  "
          f"{16 * gdofs:.0f} GB/s  ({16 * gdofs / 6551 * 100:.1f}% of 6551)", flush=True)


def run_blocks(order, n, reps=10):
    """Element-block path on an n x n uniform two-phase mesh: assembly and block apply."""
    from stefanproblem_b200 import blocks as B
    nq, nf = order + 1, (order + 1) ** 2
    i, j = np.divmod(np.arange(n * n), n)
    phase = np.where(((i + 0.5) / n - 0.5) ** 2 + ((j + 0.5) / n - 0.5) ** 2 < 0.16, -1, 1).astype(np.int8)
    data = B.uniform_mesh_block_data(n, n, 1.0 / n, 1.0 / n, nq, phase, 1.0, 2.0, 1e3 * n, 1.0)
    data = {k: torch.from_numpy(v).cuda() for k, v in data.items()}
    sys_ = B.BlockSystem(n * n, order, nq * nq, 4, nq)
    sys_.assemble(data)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(3):
        sys_.assemble(data)
    e1.record()
    torch.cuda.synchronize()
    ms_a = e0.elapsed_time(e1) / 3
    x = torch.rand(sys_.ndofs, dtype=torch.float64, device="cuda")
    y = torch.empty_like(x)
    for _ in range(3):
        sys_.apply(x, out=y)
    e0.record()
    for _ in range(reps):
        sys_.apply(x, out=y)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / reps
    bpd = 16 + 8 * 5 * nf
    g = sys_.ndofs / ms / 1e6
    print(f"blocks p={order} {n}x{n}: assemble {ms_a:.2f} ms ({n * n * 5 * nf * nf / ms_a / 1e6:.1f} G block entries/s, "
          f"{n * n * 5 * nf * nf * 8 / ms_a / 1e6:.0f} GB/s written); apply {ms:.3f} ms  {g:.1f} GDOF/s  "
          f"{bpd * g:.0f} GB/s  ({bpd * g / 6551 * 100:.1f}% of 6551 at {bpd} B/dof)", flush=True)


def run_fd(n, reps=20):
    """FD Laplace apply / fused explicit step on n nodes, one interface in the middle."""
    from stefanproblem_b200 import finite_difference as FD
    phaseid = np.ones(n, dtype=np.int8)
    phaseid[n // 2:] = 2
    op = FD.FDOperator(phaseid, 1.0 / (n - 1))
    u = torch.rand(n, dtype=torch.float64, device="cuda")
    y = torch.empty_like(u)
    for name, fn in (("apply", lambda: op.apply(u, out=y)), ("step", lambda: op.step(u, 1e-9, 1.0, 2.0, out=y))):
        for _ in range(3):
            fn()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(reps):
            fn()
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / reps
        g = n / ms / 1e6
        print(f"FD {name} n={n}: {ms:.3f} ms  {g:.1f} Gnode/s  {17 * g:.0f} GB/s  ({17 * g / 6551 * 100:.1f}% of 6551; "
              f"17 B per node: u read, y written, 1 B row kind)", flush=True)


if __name__ == "__main__":
    import argparse
    ap = argparse.ArgumentParser()
    ap.add_argument("--order", type=int, default=0)
    ap.add_argument("--n", type=int, default=0)
    ap.add_argument("--nx", type=int, default=0)
    ap.add_argument("--ny", type=int, default=0)
    ap.add_argument("--algo", type=int, default=0)
    ap.add_argument("--sweep", type=int, default=0)
    ap.add_argument("--reps", type=int, default=20)
    ap.add_argument("--two-phase", type=int, default=1)
    ap.add_argument("--ldg", type=int, default=0)
    ap.add_argument("--fd", type=int, default=0)
    ap.add_argument("--blocks", type=int, default=0)
    ap.add_argument("--pattern", default="circle", help="two-phase pattern: circle|vsplit|hsplit|allneg|hstripes")
    args = ap.parse_args()
    PATTERN = args.pattern
    if args.blocks:
        for o in ([args.order] if args.order else [1, 2, 3]):
            run_blocks(o, args.n or {1: 2000, 2: 1000, 3: 600}[o], args.reps)
        sys.exit(0)
    if args.fd:
        run_fd(args.n or (1 << 27), args.reps)
        sys.exit(0)
    if args.ldg:
        orders = [args.order] if args.order else [1, 2, 3]
        for o in orders:
            run_ldg(o, args.n or {1: 913, 2: 609, 3: 457}[o], bool(args.two_phase), args.reps, args.algo)
        sys.exit(0)
    if args.order:
        n = args.n or {1: 5000, 2: 3334, 3: 2500}[args.order]
        run(args.order, args.nx or n, args.ny or n, bool(args.two_phase), args.algo, args.reps)
        sys.exit(0)
    a = torch.empty(1 << 28, dtype=torch.float64, device="cuda"); b = torch.empty_like(a)
    for _ in range(3): b.copy_(a)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(10): b.copy_(a)
    e1.record(); torch.cuda.synchronize()
    print("copy GB/s:", 2 * a.numel() * 8 * 10 / e0.elapsed_time(e1) / 1e6, flush=True)
    del a, b
    for order, n in ((1, 5000), (2, 3334), (3, 2500)):
        for tp in (False, True):
            for algo in (0, 1):
                run(order, n, n, tp, algo)
