"""GPU parity of the IP-DG apply kernels against the oracle (`A_assembled @ x`).

Tolerance: north star <= 1e-12 relative per operator application; asserted here at
1e-13 (observed ~5e-16)."""
import numpy as np
import pytest

from oracle import fast
from oracle.mesh import UniformMesh2D

pytestmark = pytest.mark.gpu

TOL = 1e-13
ALL, REFSYS = 255, 1 | 2 | 4 | 32 | 64


def _case(order, nx, ny, two_phase, terms, variant, lam=1.5):
    import torch
    from stefanproblem_b200 import cutcelldg as cc
    omesh = UniformMesh2D([0, 0], [1.0, 0.8], [nx, ny], (order + 1) ** 2)
    if two_phase:
        omesh.cellsign[:] = fast.circle_phase(omesh, (0.5, 0.4), 0.3)
    k1, k2, pen = 1.0, 2.0, 1e3 * nx
    coupled = bool(terms & 8)
    A = fast.ip_matrix(omesh, order, order + 1, k1, k2, pen, lam if (terms & 128) else 0.0,
                       variant, aligned_interface=coupled)
    basis = cc.LagrangeTensorProductBasis(2, order)
    mesh = cc.UniformMesh([0, 0], [1.0, 0.8], [nx, ny], basis, phase=omesh.cellsign)
    op = cc.IPOperator(mesh, order, order + 1, k1, k2, pen, lam, 0.5, variant, terms)
    x = np.random.default_rng(0).uniform(-1, 1, A.shape[0])
    return A, op, x, torch.from_numpy(x).cuda()


@pytest.mark.parametrize("order", [1, 2, 3])
@pytest.mark.parametrize("algo", [0, 1])
@pytest.mark.parametrize("nx,ny", [(3, 3), (7, 6), (17, 33), (40, 95)])
@pytest.mark.parametrize("two_phase,terms,variant", [
    (False, REFSYS, "current"), (False, REFSYS, "legacy_boundary"),
    (True, ALL, "current"), (True, REFSYS, "current")])
def test_apply_matches_oracle(cuda, order, algo, nx, ny, two_phase, terms, variant):
    A, op, x, xd = _case(order, nx, ny, two_phase, terms, variant)
    y = op.apply(xd, algo=algo).cpu().numpy()
    yr = A @ x
    assert np.linalg.norm(y - yr) / np.linalg.norm(yr) < TOL
    assert np.max(np.abs(y - yr)) / np.max(np.abs(yr)) < TOL


@pytest.mark.parametrize("order", [1, 2, 3])
@pytest.mark.parametrize("nx,ny", [(1, 1), (1, 7), (2, 1), (5, 29), (4, 30), (3, 31), (2, 61), (33, 2)])
def test_degenerate_mesh_shapes(cuda, order, nx, ny):
    """Single cells, single rows / columns, strips of exactly 29 / 30 / 31 cells."""
    A, op, x, xd = _case(order, nx, ny, nx * ny > 8, ALL, "current")
    yr = A @ x
    for algo in (0, 1):
        y = op.apply(xd, algo=algo).cpu().numpy()
        assert np.linalg.norm(y - yr) <= TOL * np.linalg.norm(yr)


@pytest.mark.parametrize("order", [1, 2, 3])
def test_march_equals_simple_large(cuda, order):
    """Two independent kernels agree on a mesh with many strips / segments."""
    import torch
    from stefanproblem_b200 import cutcelldg as cc
    nx, ny = 700, 611
    basis = cc.LagrangeTensorProductBasis(2, order)
    omesh = UniformMesh2D([0, 0], [1, 1], [nx, ny], basis.nf)
    phase = fast.circle_phase(omesh)
    mesh = cc.UniformMesh([0, 0], [1, 1], [nx, ny], basis, phase=phase)
    op = cc.IPOperator(mesh, order, order + 1, 1.0, 2.0, 1e3 * nx, 1.0, 0.5, "current", ALL)
    x = torch.from_numpy(np.random.default_rng(1).uniform(-1, 1, op.ndofs)).cuda()
    y0 = op.apply(x, algo=0)
    y1 = op.apply(x, algo=1)
    assert torch.equal(y0, y1) or (y0 - y1).norm() / y1.norm() < 1e-15
    assert (y0 - y1).abs().max() / y1.abs().max() < 1e-14
    # symmetry of the operator: x' A z == z' A x
    z = torch.from_numpy(np.random.default_rng(2).uniform(-1, 1, op.ndofs)).cuda()
    a = torch.dot(z, y0)
    b = torch.dot(x, op.apply(z))
    assert abs(a - b) / abs(a) < 1e-11


@pytest.mark.parametrize("order", [1, 2, 3])
@pytest.mark.parametrize("variant", ["current", "legacy_boundary"])
def test_rhs_matches_oracle(cuda, order, variant):
    """assemble_interior_penalty_rhs! + Robin source through the reference-shaped API."""
    from stefanproblem_b200 import cutcelldg as cc
    from stefanproblem_b200 import interior_penalty as IP
    nx, ny, nq = 9, 7, order + 1
    omesh = UniformMesh2D([0.1, -0.2], [1.0, 0.8], [nx, ny], (order + 1) ** 2)
    omesh.cellsign[:] = fast.circle_phase(omesh, (0.6, 0.2), 0.3)
    k1, k2, pen, lam, Tm = 1.0, 2.0, 1e3 * nx, 1.5, 0.5
    f1 = lambda v: np.sin(3 * v[0]) * np.cos(2 * v[1]) + 1  # noqa: E731
    f2 = lambda v: 0.5 * v[0] - v[1] ** 2  # noqa: E731
    g = lambda v: v[0] + 2 * v[1] * v[1]  # noqa: E731
    fq = np.where((omesh.cellsign > 0)[:, None],
                  fast.sample_cell_quadrature(omesh, nq, lambda x, y: f1((x, y))),
                  fast.sample_cell_quadrature(omesh, nq, lambda x, y: f2((x, y))))
    ref = fast.ip_rhs(omesh, order, nq, k1, k2, pen, fq,
                      fast.sample_face_quadrature(omesh, nq, lambda x, y: g((x, y))), lam, Tm, variant)
    basis = cc.LagrangeTensorProductBasis(2, order)
    mesh = cc.UniformMesh([0.1, -0.2], [1.0, 0.8], [nx, ny], basis, phase=omesh.cellsign)
    cq, fqd = cc.CellQuadratures(mesh, nq), cc.FaceQuadratures(mesh, nq)
    rhs = cc.SystemRHS()
    IP.assemble_boundary_source(rhs, g, basis, fqd, k1, k2, pen, mesh, variant)
    IP.assemble_two_phase_source(rhs, f1, f2, basis, cq, mesh)
    IP.assemble_robin_source(rhs, basis, fqd, lam, Tm, mesh)
    b = IP.rhs_vector(rhs, mesh, 1).cpu().numpy()
    assert np.linalg.norm(b - ref) / np.linalg.norm(ref) < 1e-13


@pytest.mark.parametrize("order", [1, 2, 3])
@pytest.mark.parametrize("nx,ny", [(1, 1), (1, 4), (3, 1), (6, 5)])
def test_coo_seam_matches_oracle(cuda, order, nx, ny):
    """The triplets handed to sparse() give the oracle's matrix (and a steady solve with
    them reproduces the oracle's solution)."""
    import scipy.sparse as sp
    from stefanproblem_b200 import cutcelldg as cc
    omesh = UniformMesh2D([0, 0], [1.0, 0.8], [nx, ny], (order + 1) ** 2)
    if nx * ny > 4:
        omesh.cellsign[:] = fast.circle_phase(omesh, (0.5, 0.4), 0.3)
    k1, k2, pen, lam = 1.0, 2.0, 1e3 * nx, 1.5
    A = fast.ip_matrix(omesh, order, order + 1, k1, k2, pen, lam)
    basis = cc.LagrangeTensorProductBasis(2, order)
    mesh = cc.UniformMesh([0, 0], [1.0, 0.8], [nx, ny], basis, phase=omesh.cellsign)
    op = cc.IPOperator(mesh, order, order + 1, k1, k2, pen, lam, 0.5, "current", ALL)
    r, c, v = (t.cpu().numpy() for t in op.to_coo(index_base=1))
    assert r.min() == 1 and r.max() == A.shape[0]
    B = sp.coo_matrix((v, (r - 1, c - 1)), shape=A.shape).tocsr()
    assert abs(A - B).max() < 1e-13 * abs(A).max()


@pytest.mark.parametrize("order", [1, 2, 3])
def test_column_ranges_compose(cuda, order):
    """apply_columns over a partition of the columns == one full apply (what the slab
    overlap in stefanproblem_b200.distributed relies on), and apply_host == apply."""
    import torch
    from stefanproblem_b200 import cutcelldg as cc
    nx, ny = 67, 95
    basis = cc.LagrangeTensorProductBasis(2, order)
    omesh = UniformMesh2D([0, 0], [1, 1], [nx, ny], basis.nf)
    mesh = cc.UniformMesh([0, 0], [1, 1], [nx, ny], basis, phase=fast.circle_phase(omesh))
    op = cc.IPOperator(mesh, order, order + 1, 1.0, 2.0, 1e3 * nx, 1.0, 0.5, "current", ALL)
    x = torch.from_numpy(np.random.default_rng(5).uniform(-1, 1, op.ndofs)).cuda()
    full = op.apply(x)
    out = torch.full_like(x, float("nan"))
    for b, e in ((1, nx - 1), (0, 1), (nx - 1, nx)):
        op.apply_columns(x, out, b, e)
    assert torch.equal(out, full)
    yh = op.apply_host(x.cpu().pin_memory(), nslabs=5)
    assert torch.equal(yh, full.cpu())


@pytest.mark.parametrize("order", [1, 2, 3])
@pytest.mark.parametrize("algo", [0, 1])
@pytest.mark.parametrize("nx,ny,cut", [(64, 95, 29), (37, 64, 18), (9, 31, 4)])
def test_two_slabs_equal_one(cuda, order, algo, nx, ny, cut):
    """Column-slab partition: two slab handles fed each other's edge column reproduce the
    one-GPU result (every production kernel reads the halo columns its own way)."""
    import torch
    from stefanproblem_b200 import cutcelldg as cc
    basis = cc.LagrangeTensorProductBasis(2, order)
    omesh = UniformMesh2D([0, 0], [1, 1], [nx, ny], basis.nf)
    P = fast.circle_phase(omesh).reshape(nx, ny)
    pen = 1e3 * nx
    full = cc.IPOperator(cc.UniformMesh([0, 0], [1, 1], [nx, ny], basis, phase=P.ravel()),
                         order, order + 1, 1.0, 2.0, pen, 1.0, 0.5, "current", ALL)
    lo = cc.IPOperator(cc.UniformMesh([0, 0], [cut / nx, 1], [cut, ny], basis, phase=P[:cut].ravel()),
                       order, order + 1, 1.0, 2.0, pen, 1.0, 0.5, "current", ALL, phase_hi=P[cut])
    hi = cc.IPOperator(cc.UniformMesh([cut / nx, 0], [1 - cut / nx, 1], [nx - cut, ny], basis, phase=P[cut:].ravel()),
                       order, order + 1, 1.0, 2.0, pen, 1.0, 0.5, "current", ALL, phase_lo=P[cut - 1])
    x = torch.from_numpy(np.random.default_rng(11).uniform(-1, 1, full.ndofs)).cuda()
    col = ny * basis.nf
    yf = full.apply(x, algo=1)
    # clones: fresh (aligned) allocations, as separate GPUs would hold them
    ylo = lo.apply(x[:cut * col].clone(), x_hi=x[cut * col:(cut + 1) * col].clone(), algo=algo)
    yhi = hi.apply(x[cut * col:].clone(), x_lo=x[(cut - 1) * col:cut * col].clone(), algo=algo)
    # the slab meshes compute their cell width from their own extent: last-bit differences in h
    assert (torch.cat([ylo, yhi]) - yf).norm() <= 1e-13 * yf.norm()


@pytest.mark.parametrize("order", [1, 2, 3])
def test_rhs_of_two_slabs_equals_one(cuda, order):
    """Slab handles produce their part of the right-hand side (boundary terms only where the slab
    touches the domain boundary, Robin source on the faces they share with the neighbour)."""
    import torch
    from stefanproblem_b200 import cutcelldg as cc
    nx, ny, cut, nq = 11, 9, 4, order + 1
    basis = cc.LagrangeTensorProductBasis(2, order)
    omesh = UniformMesh2D([0, 0], [1, 1], [nx, ny], basis.nf)
    P = fast.circle_phase(omesh).reshape(nx, ny)
    P[cut - 1, 3:6] = 1   # an interface that runs along the cut
    P[cut, 3:6] = -1
    pen, lam, Tm = 1e3 * nx, 1.3, 0.4
    mk = lambda m, ph, **kw: cc.IPOperator(m, order, nq, 1.0, 2.0, pen, lam, Tm, "current", ALL, **kw)  # noqa: E731
    full = mk(cc.UniformMesh([0, 0], [1, 1], [nx, ny], basis, phase=P.ravel()), P)
    lo = mk(cc.UniformMesh([0, 0], [cut / nx, 1], [cut, ny], basis, phase=P[:cut].ravel()), P, phase_hi=P[cut])
    hi = mk(cc.UniformMesh([cut / nx, 0], [1 - cut / nx, 1], [nx - cut, ny], basis, phase=P[cut:].ravel()), P,
            phase_lo=P[cut - 1])
    rng = np.random.default_rng(4)
    f = rng.uniform(-1, 1, (nx, ny, nq * nq))
    gb, gr, gt, gl = (rng.uniform(-1, 1, (n, nq)) for n in (nx, ny, nx, ny))
    dev = lambda a: torch.from_numpy(np.ascontiguousarray(a)).cuda()  # noqa: E731
    bfull = full.rhs(dev(f), dev(np.concatenate([gb.ravel(), gr.ravel(), gt.ravel(), gl.ravel()])))
    z = np.zeros((ny, nq))
    blo = lo.rhs(dev(f[:cut]), dev(np.concatenate([gb[:cut].ravel(), z.ravel(), gt[:cut].ravel(), gl.ravel()])))
    bhi = hi.rhs(dev(f[cut:]), dev(np.concatenate([gb[cut:].ravel(), gr.ravel(), gt[cut:].ravel(), z.ravel()])))
    assert (torch.cat([blo, bhi]) - bfull).norm() <= 1e-13 * bfull.norm()


@pytest.mark.parametrize("order", [1, 2, 3])
def test_interface_sums_match_oracle(cuda, order):
    import torch
    from oracle import timeloop
    from stefanproblem_b200 import cutcelldg as cc
    nx, ny = 12, 9
    omesh = UniformMesh2D([0, 0], [1.0, 0.9], [nx, ny], (order + 1) ** 2)
    omesh.cellsign[:] = fast.circle_phase(omesh, (0.45, 0.5), 0.3)
    k1, k2, lam, Tm = 1.0, 2.0, 1.7, 0.4
    basis = cc.LagrangeTensorProductBasis(2, order)
    mesh = cc.UniformMesh([0, 0], [1.0, 0.9], [nx, ny], basis, phase=omesh.cellsign)
    op = cc.IPOperator(mesh, order, order + 1, k1, k2, 1e3 * nx, lam, Tm, "current", ALL)
    u = np.random.default_rng(7).uniform(-1, 1, op.ndofs)
    ref = timeloop.ip_interface_sums(omesh, order, order + 1, k1, k2, lam, Tm, u)
    got = op.interface_sums(torch.from_numpy(u).cuda()).cpu().numpy()
    assert np.all(np.abs(got - ref) <= 1e-12 * np.abs(ref).max())
    # two column slabs: every face is counted once (owned by its +1 cell)
    cut = 5
    P = omesh.cellsign.reshape(nx, ny)
    col = ny * basis.nf
    lo = cc.IPOperator(cc.UniformMesh([0, 0], [cut / nx, 0.9], [cut, ny], basis, phase=P[:cut].ravel()),
                       order, order + 1, k1, k2, 1e3 * nx, lam, Tm, "current", ALL, phase_hi=P[cut])
    hi = cc.IPOperator(cc.UniformMesh([cut / nx, 0], [1 - cut / nx, 0.9], [nx - cut, ny], basis, phase=P[cut:].ravel()),
                       order, order + 1, k1, k2, 1e3 * nx, lam, Tm, "current", ALL, phase_lo=P[cut - 1])
    ud = torch.from_numpy(u).cuda()
    s = (lo.interface_sums(ud[:cut * col].contiguous(), u_hi=ud[cut * col:(cut + 1) * col].contiguous())
         + hi.interface_sums(ud[cut * col:].contiguous(), u_lo=ud[(cut - 1) * col:cut * col].contiguous()))
    assert np.all(np.abs(s.cpu().numpy() - ref) <= 1e-11 * np.abs(ref).max())


@pytest.mark.parametrize("order", [1, 2])
def test_hundred_explicit_euler_steps(cuda, order):
    """u <- u + dt Minv (b - A u), 100 steps, vs the oracle loop: <= 1e-10 (north star)."""
    import torch
    from oracle import timeloop
    from stefanproblem_b200 import cutcelldg as cc
    nx, ny, nq = 10, 8, order + 1
    omesh = UniformMesh2D([0, 0], [1.0, 1.0], [nx, ny], (order + 1) ** 2)
    omesh.cellsign[:] = fast.circle_phase(omesh)
    k1, k2, pen, lam, Tm = 1.0, 2.0, 10.0 * nx, 1.0, 0.5
    A = fast.ip_matrix(omesh, order, nq, k1, k2, pen, lam)
    f = fast.sample_cell_quadrature(omesh, nq, lambda x, y: 1.0 + np.sin(3 * x) * y)
    g = fast.sample_face_quadrature(omesh, nq, lambda x, y: x + y)
    b = fast.ip_rhs(omesh, order, nq, k1, k2, pen, f, g, lam, Tm)
    Minv = timeloop.ip_block_mass_inverse(omesh, order, nq)
    lam_max = np.abs(np.linalg.eigvals((np.kron(np.eye(nx * ny), Minv) @ A.toarray()))).max()
    dt = 1.0 / lam_max
    u0 = np.random.default_rng(3).uniform(-1, 1, A.shape[0])
    ref = timeloop.ip_euler_run(A, b, Minv, u0, dt, 100)
    basis = cc.LagrangeTensorProductBasis(2, order)
    mesh = cc.UniformMesh([0, 0], [1.0, 1.0], [nx, ny], basis, phase=omesh.cellsign)
    op = cc.IPOperator(mesh, order, nq, k1, k2, pen, lam, Tm, "current", ALL)
    u = torch.from_numpy(u0).cuda()
    bd = torch.from_numpy(b).cuda()
    y = torch.empty_like(u)
    for _ in range(100):
        op.apply(u, out=y)
        op.euler_update(u, y, bd, dt)
    assert np.linalg.norm(u.cpu().numpy() - ref) <= 1e-10 * np.linalg.norm(ref)


@pytest.mark.parametrize("order", [1, 2, 3])
def test_hundred_fused_steps(cuda, order):
    """stefan_ipdg_step (apply + mass inverse + update fused into the march kernel's store epilogue, K9):
    100 steps vs the oracle loop <= 1e-10 (north star), on one slab and on two slabs that read each other's
    edge column of the CURRENT iterate."""
    import torch
    from oracle import timeloop
    from stefanproblem_b200 import cutcelldg as cc
    nx, ny, nq = 10, 8, order + 1
    omesh = UniformMesh2D([0, 0], [1.0, 1.0], [nx, ny], (order + 1) ** 2)
    omesh.cellsign[:] = fast.circle_phase(omesh)
    k1, k2, pen, lam, Tm = 1.0, 2.0, 10.0 * nx, 1.0, 0.5
    A = fast.ip_matrix(omesh, order, nq, k1, k2, pen, lam)
    f = fast.sample_cell_quadrature(omesh, nq, lambda x, y: 1.0 + np.sin(3 * x) * y)
    g = fast.sample_face_quadrature(omesh, nq, lambda x, y: x + y)
    b = fast.ip_rhs(omesh, order, nq, k1, k2, pen, f, g, lam, Tm)
    Minv = timeloop.ip_block_mass_inverse(omesh, order, nq)
    lam_max = np.abs(np.linalg.eigvals((np.kron(np.eye(nx * ny), Minv) @ A.toarray()))).max()
    dt = 1.0 / lam_max
    u0 = np.random.default_rng(3).uniform(-1, 1, A.shape[0])
    ref = timeloop.ip_euler_run(A, b, Minv, u0, dt, 100)
    basis = cc.LagrangeTensorProductBasis(2, order)
    P = omesh.cellsign.reshape(nx, ny)
    mk = lambda m, **kw: cc.IPOperator(m, order, nq, k1, k2, pen, lam, Tm, "current", ALL, **kw)  # noqa: E731
    op = mk(cc.UniformMesh([0, 0], [1.0, 1.0], [nx, ny], basis, phase=omesh.cellsign))
    u, v = torch.from_numpy(u0).cuda(), torch.empty(A.shape[0], dtype=torch.float64, device="cuda")
    bd = torch.from_numpy(b).cuda()
    for _ in range(100):
        op.step(u, bd, dt, out=v)
        u, v = v, u
    assert np.linalg.norm(u.cpu().numpy() - ref) <= 1e-10 * np.linalg.norm(ref)
    # two slabs
    cut, col = 4, ny * basis.nf
    lo = mk(cc.UniformMesh([0, 0], [cut / nx, 1], [cut, ny], basis, phase=P[:cut].ravel()), phase_hi=P[cut])
    hi = mk(cc.UniformMesh([cut / nx, 0], [1 - cut / nx, 1], [nx - cut, ny], basis, phase=P[cut:].ravel()),
            phase_lo=P[cut - 1])
    u, v = torch.from_numpy(u0).cuda(), torch.empty(A.shape[0], dtype=torch.float64, device="cuda")
    for _ in range(100):
        lo.step(u[:cut * col], bd[:cut * col], dt, out=v[:cut * col], u_hi=u[cut * col:(cut + 1) * col])
        hi.step(u[cut * col:], bd[cut * col:], dt, out=v[cut * col:], u_lo=u[(cut - 1) * col:cut * col])
        u, v = v, u
    assert np.linalg.norm(u.cpu().numpy() - ref) <= 1e-10 * np.linalg.norm(ref)


@pytest.mark.parametrize("order,nx,ny", [(1, 300, 217), (2, 131, 170), (3, 97, 125)])
def test_fused_step_equals_apply_plus_update(cuda, order, nx, ny):
    """Many strips / tiles / mixed columns: the fused step against apply (plain kernel) + euler_update."""
    import torch
    from stefanproblem_b200 import cutcelldg as cc
    basis = cc.LagrangeTensorProductBasis(2, order)
    omesh = UniformMesh2D([0, 0], [1, 1], [nx, ny], basis.nf)
    mesh = cc.UniformMesh([0, 0], [1, 1], [nx, ny], basis, phase=fast.circle_phase(omesh))
    op = cc.IPOperator(mesh, order, order + 1, 1.0, 2.0, 1e3 * nx, 1.0, 0.5, "current", ALL)
    gen = torch.Generator(device="cuda").manual_seed(5)
    u = torch.rand(op.ndofs, dtype=torch.float64, device="cuda", generator=gen) * 2 - 1
    b = torch.rand(op.ndofs, dtype=torch.float64, device="cuda", generator=gen) * 2 - 1
    dt = 1e-7
    for bb in (b, None):
        got = op.step(u, bb, dt)
        want = u.clone()
        op.euler_update(want, op.apply(u, algo=op.SIMPLE), bb, dt)
        assert (got - want).norm() <= 1e-14 * want.norm()
        assert (got - u).norm() > 0


@pytest.mark.parametrize("order,n", [(1, 12), (2, 9), (3, 6)])
@pytest.mark.parametrize("two_phase", [False, True])
@pytest.mark.parametrize("precond", ["block_jacobi", "none"])
def test_cg_solve_and_l2_error(cuda, order, n, two_phase, precond):
    """`matrix \\ rhs` on the device (preconditioned CG over the matrix-free apply) against the
    oracle's sparse direct solve, and mesh_L2_error on the device against the oracle's norm;
    problem of uncut_mesh_tests/test_uncut_IP_convergence.jl:8-16 (trig solution, pen = 1e3/h)."""
    import scipy.sparse.linalg as spla
    import torch
    from oracle.basis import LagrangeTensorProductBasis as OB
    from oracle.mesh import CellQuadratures
    from oracle.norms import integral_norm_on_mesh, mesh_L2_error
    from stefanproblem_b200 import cutcelldg as cc
    nq = order + 1
    ex = lambda v: np.cos(2 * np.pi * v[0]) * np.sin(2 * np.pi * v[1])  # noqa: E731
    src = lambda v: 8 * np.pi ** 2 * np.cos(2 * np.pi * v[0]) * np.sin(2 * np.pi * v[1])  # noqa: E731
    omesh = UniformMesh2D([0, 0], [1, 1], [n, n], (order + 1) ** 2)
    k1, k2, lam, Tm = 1.0, (2.0 if two_phase else 1.0), (1.0 if two_phase else 0.0), 0.5
    if two_phase:
        omesh.cellsign[:] = fast.circle_phase(omesh)
    pen = 1e3 * n
    A = fast.ip_matrix(omesh, order, nq, k1, k2, pen, lam)
    b = fast.ip_rhs(omesh, order, nq, k1, k2, pen,
                    fast.sample_cell_quadrature(omesh, nq, lambda x, y: src((x, y))),
                    fast.sample_face_quadrature(omesh, nq, lambda x, y: ex((x, y))), lam, Tm)
    xref = spla.spsolve(A.tocsc(), b)
    basis = cc.LagrangeTensorProductBasis(2, order)
    mesh = cc.UniformMesh([0, 0], [1, 1], [n, n], basis, phase=omesh.cellsign)
    op = cc.IPOperator(mesh, order, nq, k1, k2, pen, lam, Tm, "current", ALL)
    x, info = op.solve(torch.from_numpy(b).cuda(), tol=1e-13, maxiter=20000, preconditioner=precond)
    assert info["converged"], info
    assert np.linalg.norm(x.cpu().numpy() - xref) <= 1e-8 * np.linalg.norm(xref)
    if precond == "block_jacobi":
        _, plain = op.solve(torch.from_numpy(b).cuda(), tol=1e-13, maxiter=20000, preconditioner="none")
        assert info["iterations"] <= plain["iterations"]
    # error norms of the GPU solution against the exact solution
    uex_qp = torch.from_numpy(fast.sample_cell_quadrature(omesh, nq, lambda xx, yy: ex((xx, yy)))).cuda()
    err, ref = op.l2_error(x, uex_qp)
    oerr = mesh_L2_error(xref[None, :], ex, OB(2, order), CellQuadratures(nq), omesh)[0]
    oref = integral_norm_on_mesh(ex, CellQuadratures(nq), omesh, 1)[0]
    assert abs(err - oerr) <= 1e-7 * oerr + 1e-12 and abs(ref - oref) <= 1e-12 * oref


def test_reference_convergence_study_end_to_end_on_gpu(cuda):
    """examples/uncut_ip_convergence.py = test_uncut_IP_convergence.jl through the Python twin
    (GPU operator, GPU rhs, device CG, device L2 error).  The reference's acceptance
    criterion for this study is the convergence rate (`rate > 1.95`-style @tests, SURVEY.md
    4.1); the errors themselves are checked against the oracle's direct solve above."""
    import importlib.util
    import os
    path = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "examples", "uncut_ip_convergence.py")
    spec = importlib.util.spec_from_file_location("uncut_ip_convergence", path)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    nelmts, err, rate = mod.main(1, powers=(1, 2, 3, 4, 5))
    assert np.all(np.diff(err) < 0) and rate[-1] > 1.95 and rate[-2] > 1.9
    nelmts, err2, rate2 = mod.main(2, powers=(1, 2, 3, 4))
    assert rate2[-1] > 2.9
    # the oracle's numbers for the first meshes (same variant, direct solve)
    from oracle.basis import LagrangeTensorProductBasis as OB
    from oracle.mesh import CellQuadratures
    from oracle.norms import mesh_L2_error
    import scipy.sparse.linalg as spla
    for n, e in zip(nelmts[:2], err[:2]):
        om = UniformMesh2D([0, 0], [1, 1], [n, n], 4)
        A = fast.ip_matrix(om, 1, 2, 1.0, 1.0, 1e3 * n)
        b = fast.ip_rhs(om, 1, 2, 1.0, 1.0, 1e3 * n,
                        fast.sample_cell_quadrature(om, 2, lambda x, y: mod.source_term((x, y), 1.0)),
                        fast.sample_face_quadrature(om, 2, lambda x, y: mod.exact_solution((x, y))))
        xo = spla.spsolve(A.tocsc(), b)
        eo = mesh_L2_error(xo[None, :], mod.exact_solution, OB(2, 1), CellQuadratures(2), om)[0]
        assert abs(e - eo) <= 1e-7 * eo


@pytest.mark.parametrize("order,nx,ny", [(1, 512, 512), (2, 384, 384), (3, 160, 192)])
@pytest.mark.parametrize("variant", ["current", "legacy_boundary"])
def test_apply_matches_c_oracle_many_tiles(cuda, order, nx, ny, variant):
    """The march kernel against the ORACLE (C restatement of the reference's COO -> CSR -> SpMV path,
    oracle/csrc/ip_oracle.c, itself checked against the literal NumPy oracle on the CPU) on meshes with many
    strips, many column tiles and many mixed columns -- not only against the repo's own plain kernel, which
    shares its tables (a shared-table bug that shows only with many tiles would slip through)."""
    import torch
    from oracle.cport import CIPOracle
    from stefanproblem_b200 import cutcelldg as cc
    basis = cc.LagrangeTensorProductBasis(2, order)
    omesh = UniformMesh2D([0, 0], [1, 1], [nx, ny], basis.nf)
    phase = fast.circle_phase(omesh)
    k1, k2, pen, lam = 1.0, 2.0, 1e3 * nx, 1.0
    C = CIPOracle(order, order + 1, nx, ny, 1.0 / nx, 1.0 / ny, k1, k2, pen, lam, variant, ALL, phase)
    x = np.random.default_rng(11).uniform(-1, 1, C.nrows)
    yr = C.spmv(x)
    mesh = cc.UniformMesh([0, 0], [1, 1], [nx, ny], basis, phase=phase)
    op = cc.IPOperator(mesh, order, order + 1, k1, k2, pen, lam, 0.5, variant, ALL)
    y = op.apply(torch.from_numpy(x).cuda()).cpu().numpy()
    assert np.linalg.norm(y - yr) <= 1e-13 * np.linalg.norm(yr)
    # per-cell: no single cell may be off either (a wrong tile edge would drown in the global norm)
    d = np.abs(y - yr).reshape(nx * ny, -1).max(axis=1)
    assert d.max() <= 1e-12 * np.abs(yr).max()


def test_ip_convergence_csv_reproduced_from_the_gpu(cuda):
    """StefanDG/InteriorPenalty/uncut_mesh_tests/IP_convergence.csv (all five rows, n = 3 ... 33, p = 1 and 2)
    from the GPU path: the legacy-boundary operator through the COO seam (stefan_ipdg_assemble_coo), the right-hand
    side from stefan_ipdg_rhs, the reference's own `matrix \\ rhs` step on the host (sparse direct solve of the
    GPU's triplets), mesh_L2_error on the device.  Tolerance 1e-10 up to n = 9; from n = 17 on the systems
    (pen = 1e3 / h) are conditioned badly enough that round-off-level differences in the triplets move the
    error norm by 1e-10 ... 1e-9 relative (observed 1.5e-10 at n = 17, p = 2; the CPU oracle itself is at
    5.6e-10 / 1.3e-9 at n = 33): 5e-9 there."""
    import json
    import os
    import scipy.sparse as sp
    import scipy.sparse.linalg as spla
    import torch
    from stefanproblem_b200 import cutcelldg as cc
    from stefanproblem_b200 import interior_penalty as IP
    gold = json.load(open(os.path.join(os.path.dirname(__file__), "golden", "reference_csv.json")))["ip_uncut"]
    ex = lambda v: np.cos(2 * np.pi * v[0]) * np.sin(2 * np.pi * v[1])  # noqa: E731
    src = lambda v: 8 * np.pi ** 2 * np.cos(2 * np.pi * v[0]) * np.sin(2 * np.pi * v[1])  # noqa: E731
    for n, lin, quad in gold:
        for order, want in ((1, lin), (2, quad)):
            nq = order + 1
            basis = cc.LagrangeTensorProductBasis(2, order)
            mesh = cc.UniformMesh([0.0, 0.0], [1.0, 1.0], [n, n], basis)
            pen = 1e3 / min(mesh.h)
            cq, fq = cc.CellQuadratures(mesh, nq), cc.FaceQuadratures(mesh, nq)
            A, b = cc.SystemMatrix(), cc.SystemRHS()
            IP.assemble_interior_penalty_linear_system(A, basis, cq, fq, None, 1.0, 1.0, pen, mesh, "legacy_boundary")
            IP.assemble_interior_penalty_rhs(b, src, ex, basis, cq, fq, 1.0, 1.0, pen, mesh, "legacy_boundary")
            op = cc.sparse_operator(A, mesh, 1)
            r, c, v = (t.cpu().numpy() for t in op.to_coo())
            M = sp.coo_matrix((v, (r, c)), shape=op.shape).tocsc()
            sol = spla.spsolve(M, cc.rhs_vector(b, mesh, 1).cpu().numpy())
            xq, yq = cc.cell_quadrature_points(mesh, nq)
            uex = torch.from_numpy(np.ascontiguousarray(ex(np.stack([xq, yq])))).cuda()
            err, _ = op.l2_error(torch.from_numpy(sol).cuda(), uex)
            tol = 1e-10 if n <= 9 else 5e-9
            assert abs(err - want) / want < tol, (n, order, err, want)


@pytest.mark.parametrize("order,preconditioner", [(1, "block_jacobi"), (2, "block_jacobi"), (3, "none")])
def test_cg_across_two_slabs_equals_single_solve(cuda, order, preconditioner):
    """stefan_ipdg_cg_solve_slab: two slab handles, each driven by its own host thread (standing in for
    the two processes of a two-GPU run), the halo exchange and the all-reduce callbacks implemented with
    copies between the slabs' vectors and a host barrier.  Same iteration as the one-handle solve: equal
    iteration counts (to the check granularity) and solutions equal to the solver tolerance."""
    import threading
    import torch
    from stefanproblem_b200 import cutcelldg as cc
    from stefanproblem_b200.distributed import _wrap_device
    nx, ny, cut, nq = 21, 17, 9, order + 1
    basis = cc.LagrangeTensorProductBasis(2, order)
    omesh = UniformMesh2D([0, 0], [1, 1], [nx, ny], basis.nf)
    P = fast.circle_phase(omesh).reshape(nx, ny)
    pen, lam, Tm = 10.0 * nx * (order + 1) ** 2, 1.0, 0.5
    mk = lambda m, **kw: cc.IPOperator(m, order, nq, 1.0, 2.0, pen, lam, Tm, "current", ALL, **kw)  # noqa: E731
    full = mk(cc.UniformMesh([0, 0], [1, 1], [nx, ny], basis, phase=P.ravel()))
    ops = [mk(cc.UniformMesh([0, 0], [cut / nx, 1], [cut, ny], basis, phase=P[:cut].ravel()), phase_hi=P[cut]),
           mk(cc.UniformMesh([cut / nx, 0], [1 - cut / nx, 1], [nx - cut, ny], basis, phase=P[cut:].ravel()),
              phase_lo=P[cut - 1])]
    col = ny * basis.nf
    b = torch.from_numpy(np.random.default_rng(3).uniform(-1, 1, full.ndofs)).cuda()
    xref, iref = full.solve(b, tol=1e-11, maxiter=4000, preconditioner=preconditioner, check_every=5)
    assert iref["converged"]
    bs = [b[:cut * col].clone(), b[cut * col:].clone()]
    # halos[r] is what rank r RECEIVES: rank 0 its x_hi (rank 1's first column), rank 1 its x_lo
    halos = [torch.zeros(col, dtype=torch.float64, device="cuda") for _ in range(2)]
    bar = threading.Barrier(2, timeout=120)
    shared, results, failures = {}, [None, None], []

    def run(r):
        other = 1 - r

        def exchange(v_ptr, st):
            ext = torch.cuda.ExternalStream(st)
            ext.synchronize()  # v is final
            shared["v", r] = v_ptr
            bar.wait()
            ov = _wrap_device(shared["v", other], ops[other].ndofs)
            with torch.cuda.stream(ext):
                halos[r].copy_(ov[:col] if r == 0 else ov[-col:])
            ext.synchronize()
            bar.wait()  # the neighbour has copied: v may change

        def allreduce(d_ptr, count, st):
            ext = torch.cuda.ExternalStream(st)
            t = _wrap_device(d_ptr, count)
            with torch.cuda.stream(ext):
                shared["s", r] = t.cpu()
            bar.wait()
            tot = shared["s", 0] + shared["s", 1]  # the same order on both sides
            with torch.cuda.stream(ext):
                t.copy_(tot.cuda())
            ext.synchronize()
            bar.wait()

        try:
            torch.cuda.set_device(0)
            results[r] = ops[r].solve_slab(bs[r], halos[1] if r == 1 else None, halos[0] if r == 0 else None,
                                           exchange, allreduce, tol=1e-11, maxiter=4000,
                                           preconditioner=preconditioner, check_every=5)
        except Exception as e:  # noqa: BLE001
            failures.append(e)
            bar.abort()

    threads = [threading.Thread(target=run, args=(r,)) for r in range(2)]
    for t in threads:
        t.start()
    for t in threads:
        t.join(300)
    assert not failures, failures
    (x0, i0), (x1, i1) = results
    assert i0["converged"] and i1["converged"] and i0["iterations"] == i1["iterations"]
    assert abs(i0["iterations"] - iref["iterations"]) <= 5
    assert abs(i0["residual_norm"] - i1["residual_norm"]) == 0.0
    x = torch.cat([x0, x1])
    assert float((x - xref).norm() / xref.norm()) <= 1e-8
    r = b - full.apply(x, algo=1)
    assert float(r.norm() / b.norm()) <= 1e-9
