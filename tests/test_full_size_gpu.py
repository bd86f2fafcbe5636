"""Parity at the BASELINE.json sizes through size-independent properties.

The oracle cannot assemble 100 M dofs, so at full size the production kernels are checked
against what holds at any size: the independent one-thread-per-cell kernel of the same
operator (bit-level agreement up to summation order), linearity, symmetry of the IP
operator (`issymmetric`, test_uncut_IP_convergence.jl:88), and slabs composing to the
whole.  Workloads as in bench.py / SURVEY.md 8(d): two-phase circle + Robin, random x."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _circle(n):
    c = (np.arange(n) + 0.5) / n - 0.5
    return np.where(c[:, None] ** 2 + c[None, :] ** 2 < 0.16, -1, 1).astype(np.int8).reshape(-1)


def _uniform(n, seed):
    import torch
    g = torch.Generator(device="cuda")
    g.manual_seed(seed)
    return torch.rand(n, dtype=torch.float64, device="cuda", generator=g) * 2 - 1


def _rel(a, b):
    return float((a - b).norm() / b.norm())


@pytest.mark.parametrize("order,n", [(1, 5000), (2, 3334), (3, 2500)])
def test_ip_apply_at_100M_dofs(cuda, order, n):
    import torch
    from stefanproblem_b200 import cutcelldg as cc
    basis = cc.LagrangeTensorProductBasis(2, order)
    mesh = cc.UniformMesh([0, 0], [1, 1], [n, n], basis, phase=_circle(n))
    op = cc.IPOperator(mesh, order, order + 1, 1.0, 2.0, 1e3 * n, 1.0, 0.5, "current", 255)
    assert op.ndofs >= 100_000_000
    x, z = _uniform(op.ndofs, 1), _uniform(op.ndofs, 2)
    ax = op.apply(x)
    assert _rel(ax, op.apply(x, algo=1)) < 1e-15          # march kernel == plain kernel
    az = op.apply(z)
    assert _rel(op.apply(2.0 * x + z), 2.0 * ax + az) < 1e-13          # linearity
    assert abs(float(torch.dot(z, ax) - torch.dot(x, az))) <= 1e-12 * float(z.norm() * ax.norm())   # symmetry
    assert bool(torch.isfinite(ax).all())


@pytest.mark.parametrize("order,n", [(1, 913), (2, 609), (3, 457)])
def test_ldg_apply_at_10M_dofs(cuda, order, n):
    import torch
    from stefanproblem_b200 import cutcelldg as cc
    basis = cc.LagrangeTensorProductBasis(2, order)
    mesh = cc.UniformMesh([0, 0], [1, 1], [n, n], basis, phase=_circle(n))
    op = cc.LDGOperator(mesh, order, order + 1, 1.0, 2.0, 0.0, 0.0, 1.0, (np.cos(np.pi / 4), np.sin(np.pi / 4)))
    assert op.ndofs >= 10_000_000
    x, z = _uniform(op.ndofs, 3), _uniform(op.ndofs, 4)
    ax = op.apply(x)
    assert _rel(ax, op.apply(x, algo=1)) < 1e-15 and _rel(op.apply(x, algo=2), op.apply(x, algo=1)) < 1e-15
    assert _rel(op.apply(2.0 * x + z), 2.0 * ax + op.apply(z)) < 1e-13
    assert bool(torch.isfinite(ax).all())


def test_fd_apply_at_2_to_27_nodes(cuda):
    import torch
    from stefanproblem_b200 import finite_difference as FD
    n = 1 << 27
    i = np.arange(n)
    ph = np.where((i > n // 5) & (i < (n // 2) + 2), 1, 2).astype(np.int8)   # second interface 2 nodes past the cut
    h = 1.0 / (n - 1)
    u = _uniform(n, 5)
    full = FD.FDOperator(ph, h)
    y = full.apply(u)
    cut = n // 2
    lo = FD.FDOperator(ph[:cut], h, phase_hi=ph[cut:cut + 3])
    hi = FD.FDOperator(ph[cut:], h, phase_lo=ph[cut - 3:cut])
    assert torch.equal(lo.apply(u[:cut].clone(), u_hi=u[cut:cut + 3].clone()), y[:cut])     # slabs compose, bit-identical
    assert torch.equal(hi.apply(u[cut:].clone(), u_lo=u[cut - 3:cut].clone()), y[cut:])
    # interior same-phase rows are the central stencil
    j = torch.arange(10_000_000, 10_000_100, device="cuda")
    assert ph[10_000_000 - 4:10_000_104].min() == ph[10_000_000 - 4:10_000_104].max()
    central = (u[j - 1] - 2.0 * u[j] + u[j + 1]) / (h * h)
    assert float((y[j] - central).abs().max()) <= 1e-13 * float(central.abs().max())
    w = _uniform(n, 6)
    assert _rel(full.apply(2.0 * u + w), 2.0 * y + full.apply(w)) < 1e-13


# ---- substitutes for compute-sanitizer (closed on this GPU pool, profiles/r2_sanitize_closed.log) -------------------
def _mid_ops(order):
    import numpy as np
    from stefanproblem_b200 import cutcelldg as cc
    n = {1: 700, 2: 470, 3: 350}[order]
    basis = cc.LagrangeTensorProductBasis(2, order)
    i, j = np.divmod(np.arange(n * (n + 37)), n + 37)
    phase = np.where(((i + 0.5) / n - 0.5) ** 2 + ((j + 0.5) / (n + 37) - 0.5) ** 2 < 0.16, -1, 1).astype(np.int8)
    mesh = cc.UniformMesh([0, 0], [1, 1], [n, n + 37], basis, phase=phase)
    ip = cc.IPOperator(mesh, order, order + 1, 1.0, 2.0, 1e3 * n, 1.0, 0.5, "current", 255)
    ldg = cc.LDGOperator(mesh, order, order + 1, 1.0, 2.0, 0.5, 0.2, 1.3, (np.cos(0.4), np.sin(0.4)), interfacepenalty=0.9)
    return ip, ldg


@pytest.mark.parametrize("order", [1, 2, 3])
def test_march_kernels_are_deterministic_under_repetition(cuda, order):
    """Race proxy (racecheck is closed on this pool): the ring slots, mbarriers, the slot release, the tail queue and
    the staged stores are all per-warp or atomics-ordered, so every repetition must give bit-identical results -- the one
    race this design ever had (round 1: a ring slot refilled before a queued LDS had completed) showed up as results that
    changed from launch to launch.  60 back-to-back launches each (programmatic dependent launch overlaps them)."""
    import torch
    ip, ldg = _mid_ops(order)
    gen = torch.Generator(device="cuda").manual_seed(7)
    for op, fns in ((ip, ("apply", "step")), (ldg, ("apply",))):
        x = torch.rand(op.ndofs, dtype=torch.float64, device="cuda", generator=gen) * 2 - 1
        b = torch.rand(op.ndofs, dtype=torch.float64, device="cuda", generator=gen)
        for fn in fns:
            run = (lambda out: op.apply(x, out=out)) if fn == "apply" else (lambda out: op.step(x, b, 1e-9, out=out))
            first = run(torch.empty_like(x)).clone()
            outs = [torch.full_like(x, float("nan")) for _ in range(4)]
            for rep in range(60):
                run(outs[rep % 4])
                if rep % 4 == 3:
                    for o in outs:
                        assert torch.equal(o, first), (fn, rep)


@pytest.mark.parametrize("order", [1, 2, 3])
def test_march_kernels_stay_inside_their_buffers(cuda, order):
    """Memcheck proxy: x and y live inside larger allocations whose guard bands hold NaN (x) and a bit pattern (y).
    A read past the vector would put NaN into the result (TMA copies of the last chunk, the tail's neighbour loads),
    a write past it would change the pattern."""
    import torch
    ip, ldg = _mid_ops(order)
    G = 4096  # doubles of guard either side (32 KB: more than a column chunk)
    gen = torch.Generator(device="cuda").manual_seed(9)
    for op in (ip, ldg):
        n = op.ndofs
        bigx = torch.full((n + 2 * G,), float("nan"), dtype=torch.float64, device="cuda")
        bigy = torch.full((n + 2 * G,), -7.25, dtype=torch.float64, device="cuda")
        x, y = bigx[G:G + n], bigy[G:G + n]
        x.copy_(torch.rand(n, dtype=torch.float64, device="cuda", generator=gen) * 2 - 1)
        want = op.apply(x.clone())
        op.apply(x, out=y)
        assert torch.equal(y, want) and bool(torch.isfinite(y).all())
        assert bool((bigy[:G] == -7.25).all()) and bool((bigy[G + n:] == -7.25).all())
        if op is ip:
            bb = torch.rand(n, dtype=torch.float64, device="cuda", generator=gen)
            want = op.step(x.clone(), bb, 1e-9)
            bigy.fill_(-7.25)
            op.step(x, bb, 1e-9, out=y)
            assert torch.equal(y, want) and bool((bigy[:G] == -7.25).all()) and bool((bigy[G + n:] == -7.25).all())
