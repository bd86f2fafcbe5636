"""GPU parity of the local-DG apply kernel against the oracle (`A_assembled @ x`,
oracle/ldg.py restating StefanDG/LocalDG).  Tolerance 1e-13 relative (north star 1e-12)."""
import numpy as np
import pytest

from oracle import fast
from oracle.mesh import UniformMesh2D

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("order", [1, 2, 3])
@pytest.mark.parametrize("algo", [0, 1, 2])
@pytest.mark.parametrize("nx,ny", [(3, 3), (7, 6), (33, 18), (21, 65)])
@pytest.mark.parametrize("two_phase", [False, True])
@pytest.mark.parametrize("theta,pens", [(45.0, (0.0, 0.0, 1.0)), (143.0, (0.7, 0.3, 1.1)), (0.0, (1.0, 1.0, 1.0))])
def test_apply_matches_oracle(cuda, order, algo, nx, ny, two_phase, theta, pens):
    import torch
    from stefanproblem_b200 import cutcelldg as cc
    from stefanproblem_b200 import local_dg as LocalDG
    omesh = UniformMesh2D([0, 0], [1.0, 0.8], [nx, ny], (order + 1) ** 2)
    if two_phase:
        omesh.cellsign[:] = fast.circle_phase(omesh, (0.5, 0.4), 0.3)
    V0 = np.array([np.cos(np.radians(theta)), np.sin(np.radians(theta))])
    k1, k2 = 1.0, 2.0
    A = fast.ldg_matrix(omesh, order, order + 1, k1, k2, pens[0], pens[1], pens[2], V0)
    basis = cc.LagrangeTensorProductBasis(2, order)
    mesh = cc.UniformMesh([0, 0], [1.0, 0.8], [nx, ny], basis, phase=omesh.cellsign)
    cq, fq = cc.CellQuadratures(mesh, order + 1), cc.FaceQuadratures(mesh, order + 1)
    sm = cc.SystemMatrix()
    LocalDG.assemble_LDG_linear_system(sm, basis, cq, fq, None, k1, k2, pens[0], 0.0, pens[1], pens[2], V0, mesh)
    op = cc.sparse_operator(sm, mesh, 3)
    x = np.random.default_rng(0).uniform(-1, 1, A.shape[0])
    y = op.apply(torch.from_numpy(x).cuda(), algo=algo).cpu().numpy()
    yr = A @ x
    assert np.linalg.norm(y - yr) / np.linalg.norm(yr) < 1e-13
    if algo == 0 and nx * ny <= 42:
        # COO seam: the triplets handed to sparse() give the oracle's matrix
        import scipy.sparse as sp
        r, c, v = (t.cpu().numpy() for t in op.to_coo(index_base=1))
        B = sp.coo_matrix((v, (r - 1, c - 1)), shape=A.shape).tocsr()
        assert abs(A - B).max() < 1e-13 * abs(A).max()


@pytest.mark.parametrize("order", [1, 2, 3])
@pytest.mark.parametrize("algo", [0, 1])
@pytest.mark.parametrize("nx,ny", [(7, 6), (33, 65)])
def test_aligned_interface_apply_matches_oracle(cuda, order, algo, nx, ny):
    """LDG with the mesh-aligned interface terms (interfacepenalty; stefan_ldg_params.aligned_interface) against the
    oracle, through the reference-named call surface; the COO seam gives the oracle's matrix."""
    import torch
    from stefanproblem_b200 import cutcelldg as cc
    from stefanproblem_b200 import local_dg as LocalDG
    omesh = UniformMesh2D([0, 0], [1.0, 0.8], [nx, ny], (order + 1) ** 2)
    omesh.cellsign[:] = fast.circle_phase(omesh, (0.5, 0.4), 0.3)
    V0 = np.array([np.cos(0.4), np.sin(0.4)])
    k1, k2, ai, aif, an, ap = 1.0, 2.0, 0.7, 0.9, 0.3, 1.1
    A = fast.ldg_matrix(omesh, order, order + 1, k1, k2, ai, an, ap, V0, interfacepenalty=aif)
    basis = cc.LagrangeTensorProductBasis(2, order)
    mesh = cc.UniformMesh([0, 0], [1.0, 0.8], [nx, ny], basis, phase=omesh.cellsign)
    cq, fq = cc.CellQuadratures(mesh, order + 1), cc.FaceQuadratures(mesh, order + 1)
    sm = cc.SystemMatrix()
    LocalDG.assemble_LDG_linear_system(sm, basis, cq, fq, None, k1, k2, ai, aif, an, ap, V0, mesh)
    LocalDG.assemble_aligned_interface_scalar_flux_operator(sm, basis, fq, V0, mesh)
    LocalDG.assemble_aligned_interface_vector_flux_operator(sm, basis, fq, k1, k2, aif, V0, mesh)
    op = cc.sparse_operator(sm, mesh, 3)
    x = np.random.default_rng(0).uniform(-1, 1, A.shape[0])
    y = op.apply(torch.from_numpy(x).cuda(), algo=algo).cpu().numpy()
    yr = A @ x
    assert np.linalg.norm(y - yr) / np.linalg.norm(yr) < 1e-13
    if algo == 0 and nx * ny <= 42:
        import scipy.sparse as sp
        r, c, v = (t.cpu().numpy() for t in op.to_coo())
        B = sp.coo_matrix((v, (r, c)), shape=A.shape).tocsr()
        assert abs(A - B).max() < 1e-13 * abs(A).max()


@pytest.mark.parametrize("order,nx,ny", [(1, 300, 200), (2, 200, 150), (3, 130, 100)])
def test_apply_matches_oracle_many_tiles(cuda, order, nx, ny):
    """The march kernel against the ORACLE (vectorised NumPy assembly oracle/fast.py, pinned to the literal
    restatement and through it to MD-LDG_convergence.csv) on meshes with several strips, one tile per CTA slot and
    a two-phase circle (mixed columns, exception cells in the tail queue) -- not only against the plain kernel."""
    import torch
    from stefanproblem_b200 import cutcelldg as cc
    omesh = UniformMesh2D([0, 0], [1.0, 1.0], [nx, ny], (order + 1) ** 2)
    omesh.cellsign[:] = fast.circle_phase(omesh)
    V0 = np.array([np.cos(0.4), np.sin(0.4)])
    A = fast.ldg_matrix(omesh, order, order + 1, 1.0, 2.0, 0.5, 0.2, 1.3, V0)
    basis = cc.LagrangeTensorProductBasis(2, order)
    mesh = cc.UniformMesh([0, 0], [1.0, 1.0], [nx, ny], basis, phase=omesh.cellsign)
    op = cc.LDGOperator(mesh, order, order + 1, 1.0, 2.0, 0.5, 0.2, 1.3, V0)
    x = np.random.default_rng(12).uniform(-1, 1, A.shape[0])
    yr = A @ x
    y = op.apply(torch.from_numpy(x).cuda()).cpu().numpy()
    assert np.linalg.norm(y - yr) <= 1e-13 * np.linalg.norm(yr)
    d = np.abs(y - yr).reshape(nx * ny, -1).max(axis=1)
    assert d.max() <= 1e-12 * np.abs(yr).max()


@pytest.mark.parametrize("order", [1, 2, 3])
@pytest.mark.parametrize("nx,ny", [(1, 1), (1, 7), (2, 1), (5, 29), (4, 30), (3, 31), (2, 61)])
def test_degenerate_mesh_shapes(cuda, order, nx, ny):
    """Single cells, single rows / columns, strips of exactly 29 / 30 / 31 cells."""
    import torch
    from stefanproblem_b200 import cutcelldg as cc
    omesh = UniformMesh2D([0, 0], [1.0, 0.8], [nx, ny], (order + 1) ** 2)
    if nx * ny > 8:
        omesh.cellsign[:] = fast.circle_phase(omesh, (0.5, 0.4), 0.3)
    V0 = np.array([np.cos(0.7), np.sin(0.7)])
    A = fast.ldg_matrix(omesh, order, order + 1, 1.0, 2.0, 0.4, 0.3, 1.1, V0)
    basis = cc.LagrangeTensorProductBasis(2, order)
    mesh = cc.UniformMesh([0, 0], [1.0, 0.8], [nx, ny], basis, phase=omesh.cellsign)
    op = cc.LDGOperator(mesh, order, order + 1, 1.0, 2.0, 0.4, 0.3, 1.1, V0)
    x = np.random.default_rng(0).uniform(-1, 1, A.shape[0])
    yr = A @ x
    for algo in (1, 2):
        y = op.apply(torch.from_numpy(x).cuda(), algo=algo).cpu().numpy()
        assert np.linalg.norm(y - yr) <= 1e-13 * np.linalg.norm(yr)


@pytest.mark.parametrize("order", [1, 2, 3])
def test_march_equals_simple_large(cuda, order):
    """The two independent kernels agree on a mesh with many strips / segments (odd sizes)."""
    import torch
    from stefanproblem_b200 import cutcelldg as cc
    nx, ny = 301, 263
    basis = cc.LagrangeTensorProductBasis(2, order)
    omesh = UniformMesh2D([0, 0], [1, 1], [nx, ny], basis.nf)
    mesh = cc.UniformMesh([0, 0], [1, 1], [nx, ny], basis, phase=fast.circle_phase(omesh))
    op = cc.LDGOperator(mesh, order, order + 1, 1.0, 2.0, 0.5, 0.2, 1.3, (np.cos(0.4), np.sin(0.4)))
    x = torch.from_numpy(np.random.default_rng(1).uniform(-1, 1, op.ndofs)).cuda()
    y0, y1 = op.apply(x, algo=2), op.apply(x, algo=1)
    assert (y0 - y1).norm() / y1.norm() < 1e-15 and (y0 - y1).abs().max() / y1.abs().max() < 1e-14


@pytest.mark.parametrize("order,nx,ny,cut", [(2, 20, 13, 8), (1, 40, 65, 17), (2, 37, 33, 11), (3, 12, 40, 5)])
def test_two_slabs_equal_one(cuda, order, nx, ny, cut):
    import torch
    from stefanproblem_b200 import cutcelldg as cc
    basis = cc.LagrangeTensorProductBasis(2, order)
    omesh = UniformMesh2D([0, 0], [1, 1], [nx, ny], basis.nf)
    phase = fast.circle_phase(omesh)
    V0 = (np.cos(0.4), np.sin(0.4))
    args = (order, order + 1, 1.0, 2.0, 0.5, 0.2, 1.3, V0)
    full = cc.LDGOperator(cc.UniformMesh([0, 0], [1, 1], [nx, ny], basis, phase=phase), *args)
    x = torch.from_numpy(np.random.default_rng(3).uniform(-1, 1, full.ndofs)).cuda()
    yf = full.apply(x, algo=1)
    col = 3 * ny * basis.nf
    P = phase.reshape(nx, ny)
    m_lo = cc.UniformMesh([0, 0], [cut / nx, 1], [cut, ny], basis, phase=P[:cut].ravel())
    m_hi = cc.UniformMesh([cut / nx, 0], [1 - cut / nx, 1], [nx - cut, ny], basis, phase=P[cut:].ravel())
    lo = cc.LDGOperator(m_lo, *args, phase_hi=P[cut])
    hi = cc.LDGOperator(m_hi, *args, phase_lo=P[cut - 1])
    # clones: fresh (aligned) allocations, as separate GPUs would hold them
    ylo = lo.apply(x[:cut * col].clone(), x_hi=x[cut * col:(cut + 1) * col].clone(), algo=2)
    yhi = hi.apply(x[cut * col:].clone(), x_lo=x[(cut - 1) * col:cut * col].clone(), algo=2)
    # the slab meshes compute their cell width from their own extent: last-bit differences in h
    assert (torch.cat([ylo, yhi]) - yf).norm() <= 1e-13 * yf.norm()


@pytest.mark.parametrize("order,nq", [(1, 2), (2, 5), (3, 4)])
def test_rhs_and_golden_solve(cuda, order, nq):
    """assemble_LDG_rhs! on the GPU vs the oracle; and the reference's golden convergence
    number (MD-LDG_convergence.csv, n = 5) reproduced from the GPU operator + GPU rhs."""
    import json
    import os
    import torch
    from oracle.basis import LagrangeTensorProductBasis as OB
    from oracle.mesh import CellQuadratures
    from oracle.norms import integral_norm_on_mesh, mesh_L2_error
    from stefanproblem_b200 import cutcelldg as cc
    from stefanproblem_b200 import local_dg as LocalDG
    n = 5
    V0 = np.array([np.cos(np.pi / 4), np.sin(np.pi / 4)])
    ex = lambda v: np.cos(2 * np.pi * v[0]) * np.sin(2 * np.pi * v[1])  # noqa: E731
    src = lambda v: 8 * np.pi ** 2 * np.cos(2 * np.pi * v[0]) * np.sin(2 * np.pi * v[1])  # noqa: E731
    omesh = UniformMesh2D([0, 0], [1, 1], [n, n], (order + 1) ** 2)
    ref_b = fast.ldg_rhs(omesh, order, nq, 0.0, 1.0, V0,
                         fast.sample_cell_quadrature(omesh, nq, lambda x, y: src((x, y))),
                         fast.sample_face_quadrature(omesh, nq, lambda x, y: ex((x, y))))
    basis = cc.LagrangeTensorProductBasis(2, order)
    mesh = cc.UniformMesh([0, 0], [1, 1], [n, n], basis)
    cq, fq = cc.CellQuadratures(mesh, nq), cc.FaceQuadratures(mesh, nq)
    sm, rhs = cc.SystemMatrix(), cc.SystemRHS()
    LocalDG.assemble_LDG_linear_system(sm, basis, cq, fq, None, 1.0, 1.0, 0.0, 0.0, 0.0, 1.0, V0, mesh)
    LocalDG.assemble_LDG_rhs(rhs, src, ex, basis, cq, fq, 0.0, 1.0, V0, mesh)
    op = LocalDG.sparse_operator(sm, mesh, 3)
    b = LocalDG.rhs_vector(rhs, mesh, 3).cpu().numpy()
    assert np.linalg.norm(b - ref_b) <= 1e-13 * np.linalg.norm(ref_b)
    if order == 3:
        return
    # recover the matrix from the GPU operator (small: 3 * 25 * NF columns) and solve
    N = op.ndofs
    eye = torch.eye(N, dtype=torch.float64, device="cuda")
    A = np.stack([(op @ eye[i].contiguous()).cpu().numpy() for i in range(N)], axis=1)
    sol = np.linalg.solve(A, b).reshape(-1, 3).T
    ob = OB(2, order)
    ocq = CellQuadratures(nq)
    errT = mesh_L2_error(sol[0:1], ex, ob, ocq, omesh)[0] / integral_norm_on_mesh(ex, ocq, omesh, 1)[0]
    gold = json.load(open(os.path.join(os.path.dirname(__file__), "golden", "reference_csv.json")))["ldg_uncut"][1]
    assert gold[0] == 5
    ref = gold[1] if order == 1 else gold[4]
    assert abs(errT - ref) / ref < 1e-10


def test_reference_ldg_convergence_csv_through_the_coo_seam(cuda):
    """examples/uncut_ldg_convergence.py: GPU-assembled triplets (COO seam) + GPU right-hand
    side, solved on the host the way the reference does, reproduce the rows of the reference's
    stored MD-LDG_convergence.csv (tests/golden/reference_csv.json) for nelmts = 3 .. 17."""
    import importlib.util
    import json
    import os
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    spec = importlib.util.spec_from_file_location("uncut_ldg_convergence", os.path.join(root, "examples", "uncut_ldg_convergence.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    table = mod.main(powers=(1, 2, 3, 4))
    gold = json.load(open(os.path.join(root, "tests", "golden", "reference_csv.json")))["ldg_uncut"]
    for got, ref in zip(table, gold):
        assert got[0] == ref[0]
        assert np.allclose(got[1:], ref[1:], rtol=1e-9, atol=0.0), (got, ref)


def test_device_schur_solve_reproduces_the_reference_csv(cuda):
    """`LDGOperator.solve` (Schur complement on T, BiCGStab over the device operator) in the
    reference's convergence study: the errors of nelmts = 3, 5, 9 agree with the stored
    MD-LDG_convergence.csv to the iteration's tolerance (1e-7 relative; solver tol 1e-13)."""
    import importlib.util
    import json
    import os
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    spec = importlib.util.spec_from_file_location("uncut_ldg_convergence", os.path.join(root, "examples", "uncut_ldg_convergence.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    table = mod.main(powers=(1, 2, 3), device_solve=True)
    gold = json.load(open(os.path.join(root, "tests", "golden", "reference_csv.json")))["ldg_uncut"]
    for got, ref in zip(table, gold):
        assert got[0] == ref[0]
        assert np.allclose(got[1:], ref[1:], rtol=1e-7, atol=0.0), (got, ref)
