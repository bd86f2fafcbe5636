"""GPU parity of the local-DG apply kernel against the oracle (`A_assembled @ x`,
oracle/ldg.py restating StefanDG/LocalDG).  Tolerance 1e-13 relative (north star 1e-12)."""
import numpy as np
import pytest

from oracle import fast
from oracle.mesh import UniformMesh2D

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("order", [1, 2, 3])
@pytest.mark.parametrize("nx,ny", [(3, 3), (7, 6), (33, 18)])
@pytest.mark.parametrize("two_phase", [False, True])
@pytest.mark.parametrize("theta,pens", [(45.0, (0.0, 0.0, 1.0)), (143.0, (0.7, 0.3, 1.1)), (0.0, (1.0, 1.0, 1.0))])
def test_apply_matches_oracle(cuda, order, nx, ny, two_phase, theta, pens):
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
    y = (op @ torch.from_numpy(x).cuda()).cpu().numpy()
    yr = A @ x
    assert np.linalg.norm(y - yr) / np.linalg.norm(yr) < 1e-13


def test_two_slabs_equal_one(cuda):
    import torch
    from stefanproblem_b200 import cutcelldg as cc
    order, nx, ny, cut = 2, 20, 13, 8
    basis = cc.LagrangeTensorProductBasis(2, order)
    omesh = UniformMesh2D([0, 0], [1, 1], [nx, ny], basis.nf)
    phase = fast.circle_phase(omesh)
    V0 = (np.cos(0.4), np.sin(0.4))
    args = (order, order + 1, 1.0, 2.0, 0.5, 0.2, 1.3, V0)
    full = cc.LDGOperator(cc.UniformMesh([0, 0], [1, 1], [nx, ny], basis, phase=phase), *args)
    x = torch.from_numpy(np.random.default_rng(3).uniform(-1, 1, full.ndofs)).cuda()
    yf = full.apply(x)
    col = 3 * ny * basis.nf
    P = phase.reshape(nx, ny)
    m_lo = cc.UniformMesh([0, 0], [cut / nx, 1], [cut, ny], basis, phase=P[:cut].ravel())
    m_hi = cc.UniformMesh([cut / nx, 0], [1 - cut / nx, 1], [nx - cut, ny], basis, phase=P[cut:].ravel())
    lo = cc.LDGOperator(m_lo, *args, phase_hi=P[cut])
    hi = cc.LDGOperator(m_hi, *args, phase_lo=P[cut - 1])
    ylo = lo.apply(x[:cut * col].contiguous(), x_hi=x[cut * col:(cut + 1) * col].contiguous())
    yhi = hi.apply(x[cut * col:].contiguous(), x_lo=x[(cut - 1) * col:cut * col].contiguous())
    # the slab meshes compute their cell width from their own extent: last-bit differences in h
    assert (torch.cat([ylo, yhi]) - yf).norm() <= 1e-14 * yf.norm()
