"""Time loops (oracle, test-only).  The reference NEVER time-steps: it
assembles a steady operator and solves once.  The explicit loops below are therefore
DEFINITIONS shared by the oracle and the kernels, not restatements; parity for them
is unpinned by construction.

* FD:  u <- u + dt * kappa(phase) * (K u) on the rows that carry a Laplace stencil
       (central / backward / forward, finite-difference.jl:34-53); all other rows
       (Dirichlet, empty phase-2 interface rows, continuity rows) are held.
* DG:  u <- u + dt * Minv (b - A u), M the block-diagonal DG mass matrix.
"""
import numpy as np

from . import fd


def fd_step(phaseid, stepsize, u, dt, kappa1, kappa2):
    K = fd.laplace_matrix(phaseid, stepsize, dirichlet=False)
    lap = np.diff(K.indptr) > 0
    kappa = np.where(np.asarray(phaseid) == 1, kappa1, kappa2)
    return np.where(lap, u + dt * kappa * (K @ u), u)


def fd_run(phaseid, stepsize, u0, dt, kappa1, kappa2, nsteps):
    u = np.array(u0, dtype=float)
    for _ in range(nsteps):
        u = fd_step(phaseid, stepsize, u, dt, kappa1, kappa2)
    return u


def ip_block_mass_inverse(mesh, order, numqp):
    """Inverse of the block-diagonal DG mass matrix of a uniform mesh, as one NF x NF block."""
    from .basis import LagrangeTensorProductBasis
    from .mesh import CellQuadratures
    basis = LagrangeTensorProductBasis(2, order)
    M = np.zeros((basis.nf, basis.nf))
    for p, w in CellQuadratures(numqp)[1, 0]:
        V = basis(p)
        M += np.outer(V, V) * mesh.determinant_jacobian() * w
    return np.linalg.inv(M)


def ip_euler_run(A, b, Minv_block, u0, dt, nsteps):
    """u <- u + dt Minv (b - A u), nsteps times."""
    u = np.array(u0, dtype=float)
    nf = Minv_block.shape[0]
    for _ in range(nsteps):
        r = (b - A @ u).reshape(-1, nf)
        u = u + dt * (r @ Minv_block.T).ravel()
    return u


def ip_interface_sums(mesh, order, numqp, k1, k2, lam, Tm, u):
    """Face-integrated interface quantities (see stefan_ipdg_interface_sums): values and
    gradients evaluated with the basis at the face quadrature points, exactly as
    IP_postprocess.jl:1-52 does at reference points."""
    from .basis import LagrangeTensorProductBasis, transform_gradient
    from .ip import aligned_interface_faces
    from .mesh import FaceQuadratures, reference_face_normals
    basis = LagrangeTensorProductBasis(2, order)
    fq = FaceQuadratures(numqp)
    jac = mesh.jacobian()
    fdj = mesh.face_determinant_jacobian()
    normals = reference_face_normals()
    s = np.zeros(4)
    nf = basis.nf
    for c1, f1, c2, f2 in aligned_interface_faces(mesh):
        u1, u2 = u[c1 * nf:(c1 + 1) * nf], u[c2 * nf:(c2 + 1) * nf]
        n = normals[f1]
        for (p1, w), (p2, _) in zip(fq[1, f1, c1], fq[-1, f2, c2]):
            T1, T2 = basis(p1) @ u1, basis(p2) @ u2
            d1 = (transform_gradient(basis.gradient(p1), jac) @ n) @ u1
            d2 = (transform_gradient(basis.gradient(p2), jac) @ n) @ u2
            wq = w * fdj[f1]
            s += np.array([1.0, lam * (Tm - 0.5 * (T1 + T2)), k1 * d1 - k2 * d2, 0.5 * (T1 + T2)]) * wq
    return s
