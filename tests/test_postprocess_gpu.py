"""point evaluation (IP_postprocess.jl:1-52), the front-velocity formulas
(IP_robin_interface_flux_velocity.jl:213-278) and interface_L2_error (useful_routines.jl:239-274) on the device against
the literal restatement oracle/postprocess.py; then the reference's velocity check on the Robin cylinder: the front
velocity from the GPU solution against the exact -lambda (Tm - T(R))."""
import numpy as np
import pytest

from oracle import postprocess as opp
from oracle.analytical import RobinCylindricalSolver
from oracle.basis import LagrangeTensorProductBasis as OB

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("order", [1, 2, 3])
def test_point_evaluation_and_velocity_formulas_match_oracle(cuda, order):
    import torch
    from stefanproblem_b200 import cutcelldg as cc
    from stefanproblem_b200 import postprocess as PP
    rng = np.random.default_rng(order)
    ncells, npts = 37, 500
    basis, obasis = cc.LagrangeTensorProductBasis(2, order), OB(2, order)
    nf = basis.nf
    u = rng.uniform(-1, 1, ncells * nf)
    ref = rng.uniform(-1.3, 1.3, (2, npts))          # (merged cells evaluate outside [-1, 1] too)
    ids = rng.integers(0, ncells, npts).astype(np.int32)
    jac = np.array([0.013, 0.021])
    ud = torch.from_numpy(u).cuda()
    v = PP.interpolate_at_reference_points(ud, basis, ref, ids).cpu().numpy()
    g = PP.gradient_at_reference_points(ud, basis, ref, ids, jac).cpu().numpy()
    vr = opp.interpolate_at_reference_points(u, obasis, ref, ids)
    gr = opp.gradient_at_reference_points(u, obasis, ref, ids, jac)
    assert np.abs(v - vr).max() <= 1e-13 * np.abs(vr).max()
    assert np.abs(g - gr).max() <= 1e-13 * np.abs(gr).max()
    u2 = rng.uniform(-1, 1, ncells * nf)
    u2d = torch.from_numpy(u2).cuda()
    v2 = PP.interpolate_at_reference_points(u2d, basis, ref, ids)
    g2 = PP.gradient_at_reference_points(u2d, basis, ref, ids, jac)
    ang = rng.uniform(0, 2 * np.pi, npts)
    normals = np.stack([np.cos(ang), np.sin(ang)])
    lam, Tm, k1, k2 = 1.7, 0.4, 1.0, 2.0
    got = PP.front_velocities(torch.from_numpy(v).cuda(), v2, torch.from_numpy(g).cuda(), g2, torch.from_numpy(normals).cuda(),
                              lam, Tm, k1, k2)
    want = opp.front_velocities(vr, opp.interpolate_at_reference_points(u2, obasis, ref, ids), gr,
                                opp.gradient_at_reference_points(u2, obasis, ref, ids, jac), normals, lam, Tm, k1, k2)
    for a, b in zip(got, want):
        assert np.abs(a.cpu().numpy() - b).max() <= 1e-12 * np.abs(b).max()


def test_robin_cylinder_front_velocity_and_interface_error(cuda):
    """The velocity evaluation of IP_robin_interface_flux_velocity.jl on the cut-cell Robin cylinder (p = 2, n = 17):
    solution from the block path, both sides' temperatures at the interface quadrature points of every cut cell,
    velocity = -lambda (Tm - T): relative error against the exact velocity < 1e-3 (the reference plots this error;
    its discretisation error at n = 17 is O(h^3)), and interface_L2_error on both sides against the oracle."""
    import torch
    from stefanproblem_b200 import blocks as B
    from stefanproblem_b200 import cutcelldg as cc
    from stefanproblem_b200 import postprocess as PP
    from stefanproblem_b200.cutcell import CircleCutMesh
    order, n, nq = 2, 17, 5
    k1, k2, q1, q2, lam, Tm, R, b, Tw = 1.0, 2.0, 2.0, 1.0, 1.0, 0.5, 0.7, 1.5, 1.0
    sol = RobinCylindricalSolver(q1, k1, q2, k2, lam, R, b, Tw, Tm)
    m = CircleCutMesh([0, 0], [1, 1], [n, n], [0, 0], R, nq, k1, k2, 1e3 * n, lam, merge_fraction=0.25)
    system = B.BlockSystem(m.nsol, order, m.nqv, m.nslots, m.nqf).assemble(m.data)
    rhs = system.rhs(m.sample_source(lambda x, y: q1 + 0 * x, lambda x, y: q2 + 0 * x), m.sample_boundary(sol.at), Tm)
    u, info = system.solve(rhs, tol=1e-13, maxiter=100000)
    assert info["converged"]
    # the interface points seen from the + side: own reference coordinates and the - side's (face_nbr_pts)
    sel = m.slot_is_interface & (m.sol_phase == 1)[:, None] & (m.data["slot_nbr"] >= 0)
    cs = np.argwhere(sel)
    pts1, pts2, ids1, ids2, nrm = [], [], [], [], []
    for c, s in cs:
        keep = m.data["face_wts"][c, s] != 0
        pts1.append(m.data["face_pts"][c, s][keep])
        pts2.append(m.data["face_nbr_pts"][c, s][keep])
        ids1 += [c] * int(keep.sum())
        ids2 += [int(m.data["slot_nbr"][c, s])] * int(keep.sum())
        nrm.append(m.data["face_normals"][c, s][keep])
    ref1, ref2, normals = np.concatenate(pts1).T, np.concatenate(pts2).T, np.concatenate(nrm).T
    basis = cc.LagrangeTensorProductBasis(2, order)
    jac = 0.5 * m.h
    v1 = PP.interpolate_at_reference_points(u, basis, ref1, np.array(ids1, dtype=np.int32))
    v2 = PP.interpolate_at_reference_points(u, basis, ref2, np.array(ids2, dtype=np.int32))
    g1 = PP.gradient_at_reference_points(u, basis, ref1, np.array(ids1, dtype=np.int32), jac)
    g2 = PP.gradient_at_reference_points(u, basis, ref2, np.array(ids2, dtype=np.int32), jac)
    vel1, vel2, mean, gvel = (t.cpu().numpy() for t in PP.front_velocities(v1, v2, g1, g2, normals, lam, Tm, k1, k2))
    exact = -lam * (Tm - sol(R))
    assert np.abs(mean - exact).max() / abs(exact) < 1e-5          # (observed 1.4e-6)
    assert np.abs(vel1 - exact).max() / abs(exact) < 1e-5 and np.abs(vel2 - exact).max() / abs(exact) < 1e-5
    # flux-jump form k2 dnT2 - k1 dnT1 (normals out of the + side): the Robin condition makes it the same velocity
    # (observed 8e-4 at this mesh; gradients converge one order slower than values)
    assert np.abs(gvel - exact).max() / abs(exact) < 5e-3
    for sign in (1, -1):
        ssel = m.slot_is_interface & (m.sol_phase == sign)[:, None] & (m.data["slot_nbr"] >= 0)
        uex = sol.at(m.face_xy[..., 0], m.face_xy[..., 1]) * (m.data["face_wts"] != 0)
        e, nn = system.interface_l2_error(ssel, u, uex)
        er, nr = opp.interface_L2_error(order, m.data, ssel, u.cpu().numpy(), uex)
        assert abs(e - er) <= 1e-10 * er + 1e-16 and abs(nn - nr) <= 1e-12 * nr
        assert e / nn < 1e-5


@pytest.mark.parametrize("order", [1, 2, 3])
def test_ldg_field_interpolation_flux_and_error_norms_match_oracle(cuda, order):
    """Vector-field interpolation and the LDG numerical interface flux (lagrange_levelset_LDG_interface_flux.jl:256-307)
    against the literal restatement; mesh_L2_error for a 3-dof-per-node field (useful_routines.jl:95-133) against
    oracle/norms.py on an LDG-shaped solution vector."""
    import torch
    from oracle.mesh import CellQuadratures, UniformMesh2D
    from oracle.norms import integral_norm_on_mesh, mesh_L2_error
    from stefanproblem_b200 import cutcelldg as cc
    from stefanproblem_b200 import postprocess as PP
    rng = np.random.default_rng(20 + order)
    nx, ny, npts = 6, 5, 300
    basis, obasis = cc.LagrangeTensorProductBasis(2, order), OB(2, order)
    nf, ncells = basis.nf, nx * ny
    u = rng.uniform(-1, 1, ncells * nf * 3)            # (T, qx, qy) interleaved per node
    ud = torch.from_numpy(u).cuda()
    ref = rng.uniform(-1, 1, (2, npts))
    ids = rng.integers(0, ncells, npts).astype(np.int32)
    G = u.reshape(-1, 3).T                              # the reference's [ndofs, nnodes] view
    q = PP.interpolate_field_at_reference_points(ud, 3, 1, 2, basis, ref, ids).cpu().numpy()
    qr = opp.interpolate_field_at_reference_points(G[1:3], 2, obasis, ref, ids)
    assert np.abs(q - qr).max() <= 1e-13 * np.abs(qr).max()
    T = PP.interpolate_field_at_reference_points(ud, 3, 0, 1, basis, ref, ids).cpu().numpy()
    assert np.abs(T - opp.interpolate_field_at_reference_points(G[0:1], 1, obasis, ref, ids)).max() <= 1e-13
    q2 = rng.uniform(-1, 1, (2, npts))
    ang = rng.uniform(0, 2 * np.pi, npts)
    normals = np.stack([np.cos(ang), np.sin(ang)])
    V0 = (np.cos(0.7), np.sin(0.7))
    got = PP.ldg_interface_flux(torch.from_numpy(np.ascontiguousarray(qr)).cuda(), torch.from_numpy(q2).cuda(),
                                torch.from_numpy(normals).cuda(), 1.0, 2.0, V0)
    want = opp.ldg_interface_flux(qr, q2, normals, 1.0, 2.0, V0)
    for a, b in zip(got, want):
        assert np.abs(a.cpu().numpy() - b).max() <= 1e-13 * np.abs(b).max()
    # error norms of the three components
    omesh = UniformMesh2D([0, 0], [1.0, 0.8], [nx, ny], nf)
    nq = order + 2
    cq = CellQuadratures(nq)
    ex = lambda v: np.array([np.sin(v[0]) * v[1], np.cos(v[0] + v[1]), v[0] * v[1] ** 2])  # noqa: E731
    err_ref = mesh_L2_error(G, ex, obasis, cq, omesh)
    nrm_ref = integral_norm_on_mesh(ex, cq, omesh, 3)
    mesh = cc.UniformMesh([0, 0], [1.0, 0.8], [nx, ny], basis)
    xq, yq = cc.cell_quadrature_points(mesh, nq)
    uex = np.stack([np.sin(xq) * yq, np.cos(xq + yq), xq * yq ** 2], axis=-1)
    err, nrm = PP.mesh_l2_error(2, order, nq, ncells, 0.25 * mesh.h[0] * mesh.h[1], 3, ud, uex)
    assert np.abs(err - err_ref).max() <= 1e-12 * err_ref.max() and np.abs(nrm - nrm_ref).max() <= 1e-12 * nrm_ref.max()


@pytest.mark.parametrize("order", [1, 2])
def test_dg1d_error_norm_matches_oracle(cuda, order):
    """uniform_mesh_L2_error / integral_norm_on_uniform_mesh of StefanRobinIC/useful_routines.jl:52-78 (two element
    sizes either side of the interface) on the device against oracle/dg1d.py."""
    import torch
    from oracle import dg1d as od
    from oracle.basis import interpolation_points, tensor_product_quadrature
    from stefanproblem_b200 import postprocess as PP
    ne1, ne2, xi, nq = 5, 7, 0.37, order + 2
    obasis = OB(1, order)
    quad = tensor_product_quadrature(1, nq)
    mesh = od.DGMesh1D(0.0, 1.0, xi, ne1, ne2, interpolation_points(obasis))
    rng = np.random.default_rng(order)
    ncells, n1 = ne1 + ne2, order + 1
    u = rng.uniform(-1, 1, ncells * n1)
    ex = lambda x: np.sin(3 * x) + x * x  # noqa: E731
    err_ref = od.uniform_mesh_L2_error(u, ex, obasis, quad, mesh)
    nrm_ref = od.integral_norm_on_uniform_mesh(ex, quad, mesh)
    g, _ = np.polynomial.legendre.leggauss(nq)
    dx = np.array([xi / ne1] * ne1 + [(1.0 - xi) / ne2] * ne2)
    left = np.concatenate([[0.0], np.cumsum(dx)[:-1]])
    xq = left[:, None] + 0.5 * dx[:, None] * (1.0 + g[None, :])
    err, nrm = PP.mesh_l2_error(1, order, nq, ncells, 0.0, 1, torch.from_numpy(u).cuda(), ex(xq)[..., None], cell_detjac=0.5 * dx)
    assert abs(err[0] - float(np.ravel(err_ref)[0])) <= 1e-12 * float(np.ravel(err_ref)[0])
    assert abs(nrm[0] - float(np.ravel(nrm_ref)[0])) <= 1e-12 * float(np.ravel(nrm_ref)[0])
