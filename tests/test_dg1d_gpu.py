"""GPU parity of the DG1D kernels against the oracle (StefanRobinIC/DG1D restated in
oracle/dg1d.py), including the reference's own @test conditions evaluated on the
GPU-produced operator and right-hand side."""
import numpy as np
import pytest
import scipy.sparse.linalg as spla

from oracle import dg1d as o
from oracle.basis import LagrangeTensorProductBasis, interpolation_points, tensor_product_quadrature
from oracle.mesh import SystemMatrix, SystemRHS, rhs_vector, sparse_operator

pytestmark = pytest.mark.gpu
EPS = np.finfo(float).eps


def oracle_system(ne1, ne2, p, nq, q1, k1, q2, k2, lam, xi, TL, TR, Tm, pf):
    basis = LagrangeTensorProductBasis(1, p)
    quad = tensor_product_quadrature(1, nq)
    mesh = o.DGMesh1D(0.0, 1.0, xi, ne1, ne2, interpolation_points(basis))
    pen = pf / min(mesh.element_size())
    A, b = SystemMatrix(), SystemRHS()
    o.assemble_gradient_operator(A, basis, quad, k1, k2, mesh)
    o.assemble_flux_operator(A, basis, k1, k2, mesh)
    o.assemble_penalty_operator(A, basis, pen, mesh)
    o.assemble_interface_robin_operator(A, basis, lam, mesh)
    o.assemble_interface_robin_source(b, basis, lam, Tm, mesh)
    o.assemble_two_phase_source(b, q1, q2, basis, quad, mesh)
    o.assemble_boundary_flux_operator(A, basis, k1, k2, mesh)
    o.assemble_boundary_penalty_operator(A, basis, pen, mesh)
    o.assemble_boundary_rhs(b, k1, k2, TL, TR, basis, pen, mesh)
    n = mesh.number_of_nodes()
    return sparse_operator(A, n), rhs_vector(b, n), mesh, basis, quad, pen


def gpu_system(ne1, ne2, p, nq, q1, k1, q2, k2, lam, xi, TL, TR, Tm, pen):
    from stefanproblem_b200 import cutcelldg as cc
    from stefanproblem_b200 import dg1d as DG1D
    basis = cc.LagrangeTensorProductBasis(1, p)
    mesh = DG1D.DGMesh1D(0.0, 1.0, xi, ne1, ne2, cc.interpolation_points(basis))
    A, b = DG1D.SystemMatrix(), DG1D.SystemRHS()
    DG1D.assemble_gradient_operator(A, basis, nq, k1, k2, mesh)
    DG1D.assemble_flux_operator(A, basis, k1, k2, mesh)
    DG1D.assemble_penalty_operator(A, basis, pen, mesh)
    DG1D.assemble_interface_robin_operator(A, basis, lam, mesh)
    DG1D.assemble_interface_robin_source(b, basis, lam, Tm, mesh)
    DG1D.assemble_two_phase_source(b, q1, q2, basis, nq, mesh)
    DG1D.assemble_boundary_flux_operator(A, basis, k1, k2, mesh)
    DG1D.assemble_boundary_penalty_operator(A, basis, pen, mesh)
    DG1D.assemble_boundary_rhs(b, k1, k2, TL, TR, basis, pen, mesh)
    return DG1D.sparse_operator(A, mesh, 1), DG1D.rhs_vector(b, mesh, 1)


@pytest.mark.parametrize("p,nq", [(1, 2), (2, 3), (3, 4), (1, 3)])
@pytest.mark.parametrize("ne1,ne2", [(1, 1), (2, 5), (32, 32), (300, 211)])
def test_apply_and_rhs_match_oracle(cuda, p, nq, ne1, ne2):
    import torch
    args = (lambda x: 1.0 + 0 * np.asarray(x), 0.5, lambda x: 0.5 * np.cos(3 * np.asarray(x)), 1.0,
            1.3, 0.7, 0.2, 1.0, 0.3)
    M, r, mesh, basis, quad, pen = oracle_system(ne1, ne2, p, nq, lambda x: 1.0, args[1],
                                                 lambda x: 0.5 * np.cos(3 * x[0]), *args[3:], 1e2)
    op, b = gpu_system(ne1, ne2, p, nq, *args, pen)
    x = np.random.default_rng(0).uniform(-1, 1, M.shape[0])
    y = (op @ torch.from_numpy(x).cuda()).cpu().numpy()
    assert np.linalg.norm(y - M @ x) <= 1e-13 * np.linalg.norm(M @ x)
    assert np.linalg.norm(b.cpu().numpy() - r) <= 1e-13 * np.linalg.norm(r)
    # COO seam: the triplets handed to sparse() give the oracle's matrix
    import scipy.sparse as sp
    rr, cc_, vv = (t.cpu().numpy() for t in op.to_coo(index_base=1))
    B = sp.coo_matrix((vv, (rr - 1, cc_ - 1)), shape=M.shape).tocsr()
    assert abs(B - M).max() <= 1e-13 * abs(M).max()


def test_reference_robin_interface_test_on_gpu_operator(cuda):
    """StefanRobinIC/DG1D/test_robin_interface.jl:97-101, with the matrix recovered from
    the GPU operator column by column and the rhs from the GPU kernel."""
    import torch
    q1 = q2 = 0.0
    k1, k2, lam, xi, TL, TR, Tm = 1.0, 2.0, 5.0, 0.7, 0.0, 1.0, 0.8
    M, r, mesh, basis, quad, pen = oracle_system(1, 1, 1, 3, lambda x: q1, k1, lambda x: q2, k2, lam, xi, TL,
                                                 TR, Tm, 1e2)
    op, b = gpu_system(1, 1, 1, 3, lambda x: q1 + 0 * x, k1, lambda x: q2 + 0 * x, k2, lam, xi, TL, TR, Tm, pen)
    n = op.ndofs
    A = np.stack([(op @ torch.eye(n, dtype=torch.float64, device="cuda")[i].contiguous()).cpu().numpy()
                  for i in range(n)], axis=1)
    assert np.abs(A - A.T).max() <= 1e-12 * np.abs(A).max()  # issymmetric
    sol = np.linalg.solve(A, b.cpu().numpy())
    lg = (sol[1] - sol[0]) / (2 * mesh.jacobian(0))
    rg = (sol[3] - sol[2]) / (2 * mesh.jacobian(1))
    TI = 0.5 * (sol[1] + sol[2])
    assert abs(k1 * lg - k2 * rg - lam * (Tm - TI)) < 1e3 * EPS
    ex = o.Analytical1D(q1, k1, q2, k2, lam, xi, TL, TR, Tm)
    assert o.uniform_mesh_L2_error(sol, ex, basis, quad, mesh) < 1e3 * EPS
