"""Vectorised global assembly on uniform meshes (oracle, test-only).

The element-level and face-level matrices are produced by the LITERAL functions in
`oracle/ip.py` / `oracle/ldg.py` (on a local one- or two-cell numbering); only the
reference's `for cellid = 1:ncells` driver loops are replaced by NumPy
broadcasting, which is legitimate on an uncut uniform mesh because every cell/face
of the same kind produces the same local matrix.  `tests/test_oracle_fast.py`
checks these against the literal loops on small meshes.
"""
import numpy as np
import scipy.sparse as sp

from . import ip, ldg
from .basis import LagrangeTensorProductBasis
from .mesh import (
    CellQuadratures,
    FaceQuadratures,
    SystemMatrix,
    SystemRHS,
    opposite_face,
    reference_face_normals,
    rhs_vector,
    sparse_operator,
)


def _local_dense(fn, nnodes, dpn):
    sm = SystemMatrix()
    fn(sm)
    return sparse_operator(sm, nnodes * dpn).toarray()


def _celldofs(cells, nf, dpn):
    cells = np.asarray(cells, dtype=np.int64)
    return cells[:, None] * (nf * dpn) + np.arange(nf * dpn)[None, :]


class _Coo:
    def __init__(self, n):
        self.n = n
        self.r, self.c, self.v = [], [], []

    def add(self, rowdofs, coldofs, block):
        """rowdofs: (m, R), coldofs: (m, C), block: (R, C) shared by all m."""
        if rowdofs.shape[0] == 0 or not np.any(block):
            return
        m = rowdofs.shape[0]
        R, C = block.shape
        self.r.append(np.broadcast_to(rowdofs[:, :, None], (m, R, C)).ravel())
        self.c.append(np.broadcast_to(coldofs[:, None, :], (m, R, C)).ravel())
        self.v.append(np.broadcast_to(block[None], (m, R, C)).ravel())

    def csr(self):
        A = sp.coo_matrix(
            (np.concatenate(self.v), (np.concatenate(self.r), np.concatenate(self.c))),
            shape=(self.n, self.n)).tocsr()
        A.sum_duplicates()
        A.eliminate_zeros()
        return A


def _neighbours(mesh):
    """nbr[f, cell] (or -1), vectorised `cell_connectivity`."""
    nx, ny = mesh.nx, mesh.ny
    cid = np.arange(mesh.ncells)
    i, j = np.divmod(cid, ny)
    nbr = np.full((4, mesh.ncells), -1, dtype=np.int64)
    nbr[0] = np.where(j > 0, cid - 1, -1)
    nbr[1] = np.where(i < nx - 1, cid + ny, -1)
    nbr[2] = np.where(j < ny - 1, cid + 1, -1)
    nbr[3] = np.where(i > 0, cid - ny, -1)
    return nbr


def _pair_scatter(coo, L, cellsA, cellsB, nf, dpn):
    n = nf * dpn
    dA = _celldofs(cellsA, nf, dpn)
    dB = _celldofs(cellsB, nf, dpn)
    coo.add(dA, dA, L[:n, :n])
    coo.add(dA, dB, L[:n, n:])
    coo.add(dB, dA, L[n:, :n])
    coo.add(dB, dB, L[n:, n:])


# ------------------------------------------------------------------------ IP
def ip_matrix(mesh, order, numqp, k1, k2, penalty, lam=0.0, variant="current",
              aligned_interface=True):
    """CSR of `assemble_interior_penalty_linear_system!` (+ aligned interface
    flux / penalty and, if lam != 0, `assemble_face_robin_operator!` on the
    unlike-phase faces)."""
    basis = LagrangeTensorProductBasis(2, order)
    nf = basis.nf
    cq, fq = CellQuadratures(numqp), FaceQuadratures(numqp)
    jac, detjac = mesh.jacobian(), mesh.determinant_jacobian()
    fdj = mesh.face_determinant_jacobian()
    normals = reference_face_normals()
    n1, n2 = np.arange(nf), np.arange(nf, 2 * nf)
    sign = mesh.cellsign.astype(np.int64)
    nbr = _neighbours(mesh)
    cid = np.arange(mesh.ncells)
    coo = _Coo(mesh.ncells * nf)

    for s, k in ((+1, k1), (-1, k2)):
        cells = cid[sign == s]
        # volume
        V = ip.gradient_operator(basis, cq[s, 0], k, jac, detjac)
        d = _celldofs(cells, nf, 1)
        coo.add(d, d, V)
        for f in range(4):
            q1, q2 = fq[s, f, 0], fq[s, opposite_face(f), 0]
            has = nbr[f] >= 0
            safe = np.where(has, nbr[f], 0)
            # interelement: cellid < nbr, same sign
            sel = has & (sign == s) & (sign[safe] == s) & (cid < nbr[f])
            if np.any(sel):
                def both(sm, f=f, q1=q1, q2=q2, k=k):
                    ip.assemble_face_interelement_flux_operator(
                        sm, basis, q1, q2, normals[f], k, jac, fdj[f], n1, n2)
                    ip.assemble_face_penalty_operator(
                        sm, basis, q1, q2, penalty, np.repeat(fdj[f], len(q1)), n1, n2)
                _pair_scatter(coo, _local_dense(both, 2 * nf, 1), cid[sel], nbr[f][sel], nf, 1)
            # boundary
            selb = (~has) & (sign == s)
            if np.any(selb):
                def bnd(sm, f=f, q1=q1, k=k):
                    ip.assemble_boundary_face_flux_operator(
                        sm, basis, q1, normals[f], k, jac, fdj[f], n1, variant)
                    M = penalty * ip.mass_operator(basis, q1, q1, np.repeat(fdj[f], len(q1)))
                    ip.assemble_couple_cell_matrix(sm, n1, n1, 1, ip.vec(M))
                d = _celldofs(cid[selb], nf, 1)
                coo.add(d, d, _local_dense(bnd, nf, 1))
    if aligned_interface:
        for f in range(4):
            has = nbr[f] >= 0
            safe = np.where(has, nbr[f], 0)
            sel = has & (sign == +1) & (sign[safe] == -1)
            if not np.any(sel):
                continue
            q1, q2 = fq[+1, f, 0], fq[-1, opposite_face(f), 0]
            nq = len(q1)
            nrm = np.repeat(normals[f].reshape(2, 1), nq, axis=1)
            sa = np.repeat(fdj[f], nq)

            def iface(sm):
                ip.assemble_face_flux_operator(sm, basis, q1, q2, nrm, k1, k2, jac, sa, n1, n2)
                ip.assemble_face_penalty_operator(sm, basis, q1, q2, penalty, sa, n1, n2)
                if lam != 0.0:
                    ip.assemble_face_robin_operator(sm, basis, q1, q2, lam, sa, n1, n2)
            _pair_scatter(coo, _local_dense(iface, 2 * nf, 1), cid[sel], nbr[f][sel], nf, 1)
    return coo.csr()


def _quad_arrays(quad):
    pts = np.array([p for p, _ in quad])
    w = np.array([w for _, w in quad])
    return pts, w


def ip_rhs(mesh, order, numqp, k1, k2, penalty, f_qp, g_qp, lam=0.0, Tm=0.0,
           variant="current"):
    """Vectorised `assemble_interior_penalty_rhs!` (+ aligned Robin source).

    f_qp: (ncells, numqp^2) source sampled at the cell quadrature points (order of
          `tensor_product_quadrature(2, numqp)`); g_qp: (4, ncells, numqp) boundary
          data sampled at the face quadrature points (only boundary faces read).
    """
    basis = LagrangeTensorProductBasis(2, order)
    nf = basis.nf
    cq, fq = CellQuadratures(numqp), FaceQuadratures(numqp)
    jac, detjac = mesh.jacobian(), mesh.determinant_jacobian()
    fdj = mesh.face_determinant_jacobian()
    normals = reference_face_normals()
    sign = mesh.cellsign.astype(np.int64)
    kcell = np.where(sign > 0, k1, k2).astype(float)
    nbr = _neighbours(mesh)
    b = np.zeros((mesh.ncells, nf))
    pts, w = _quad_arrays(cq[1, 0])
    V = np.array([basis(p) for p in pts])  # (nq2, nf)
    b += (f_qp * (w * detjac)[None, :]) @ V
    for f in range(4):
        pts, w = _quad_arrays(fq[1, f, 0])
        Vf = np.array([basis(p) for p in pts])
        Gn = np.array([(basis.gradient(p) / jac[None, :]) @ normals[f] for p in pts])
        isb = nbr[f] < 0
        g = np.where(isb[:, None], g_qp[f], 0.0)
        b += penalty * (g * (w * fdj[f])[None, :]) @ Vf
        if variant == "current":
            b -= kcell[:, None] * ((g * (w * fdj[f])[None, :]) @ Gn)
        if lam != 0.0:
            has = nbr[f] >= 0
            safe = np.where(has, nbr[f], 0)
            isi = has & (sign != sign[safe])
            rl = 0.5 * lam * Tm * ((w * fdj[f]) @ Vf)
            b += isi[:, None] * rl[None, :]
    return b.ravel()


# ----------------------------------------------------------------------- LDG
def ldg_matrix(mesh, order, numqp, k1, k2, interiorpenalty, negboundarypenalty,
               posboundarypenalty, V0, interfacepenalty=None):
    """CSR of `assemble_LDG_linear_system!` on an uncut mesh.  interfacepenalty=None: cells of unlike sign
    are not coupled, exactly as the reference's drivers do; a number: the face-level functions are also applied
    to the unlike-phase faces (ldg.assemble_aligned_interface_*_operator)."""
    basis = LagrangeTensorProductBasis(2, order)
    nf = basis.nf
    cq, fq = CellQuadratures(numqp), FaceQuadratures(numqp)
    jac, detjac = mesh.jacobian(), mesh.determinant_jacobian()
    fdj = mesh.face_determinant_jacobian()
    normals = reference_face_normals()
    n1, n2 = np.arange(nf), np.arange(nf, 2 * nf)
    sign = mesh.cellsign.astype(np.int64)
    nbr = _neighbours(mesh)
    cid = np.arange(mesh.ncells)
    coo = _Coo(mesh.ncells * nf * 3)
    V0 = np.asarray(V0, dtype=float)

    for s, k in ((+1, k1), (-1, k2)):
        cells = cid[sign == s]

        def cellterms(sm, k=k, s=s):
            q = cq[s, 0]
            ldg.assemble_couple_cell_matrix(
                sm, n1, n1, 3, ldg.vec(ldg.divergence_operator(basis, q, jac, detjac)),
                ldg.Q_DOFS, ldg.T_DOF)
            ldg.assemble_couple_cell_matrix(
                sm, n1, n1, 3, ldg.vec(ldg.mass_matrix(basis, q, 2, detjac)),
                ldg.Q_DOFS, ldg.Q_DOFS)
            ldg.assemble_couple_cell_matrix(
                sm, n1, n1, 3, ldg.vec(ldg.gradient_operator(basis, q, k, jac, detjac)),
                ldg.T_DOF, ldg.Q_DOFS)
        d = _celldofs(cells, nf, 3)
        coo.add(d, d, _local_dense(cellterms, nf, 3))
        for f in range(4):
            q1, q2 = fq[s, f, 0], fq[s, opposite_face(f), 0]
            nq = len(q1)
            nrm = np.repeat(normals[f].reshape(2, 1), nq, axis=1)
            sa = np.repeat(fdj[f], nq)
            has = nbr[f] >= 0
            safe = np.where(has, nbr[f], 0)
            sel = has & (sign == s) & (sign[safe] == s)
            if np.any(sel):
                def face(sm, k=k):
                    ldg.assemble_face_scalar_flux_operator(sm, basis, q1, q2, nrm, V0, sa, n1, n2)
                    ldg.assemble_face_vector_flux_operator(
                        sm, basis, q1, q2, nrm, k, k, interiorpenalty, V0, sa, n1, n2)
                L = _local_dense(face, 2 * nf, 3)
                n = 3 * nf
                dA = _celldofs(cid[sel], nf, 3)
                dB = _celldofs(nbr[f][sel], nf, 3)
                coo.add(dA, dA, L[:n, :n])
                coo.add(dA, dB, L[:n, n:])
            seli = has & (sign == s) & (sign[safe] == -s)
            if interfacepenalty is not None and np.any(seli):
                kb = k2 if s == 1 else k1
                q2i = fq[-s, opposite_face(f), 0]

                def iface(sm, k=k, kb=kb, q2i=q2i):
                    ldg.assemble_face_scalar_flux_operator(sm, basis, q1, q2i, nrm, V0, sa, n1, n2)
                    ldg.assemble_face_vector_flux_operator(
                        sm, basis, q1, q2i, nrm, k, kb, interfacepenalty, V0, sa, n1, n2)
                L = _local_dense(iface, 2 * nf, 3)
                n = 3 * nf
                dA = _celldofs(cid[seli], nf, 3)
                dB = _celldofs(nbr[f][seli], nf, 3)
                coo.add(dA, dA, L[:n, :n])
                coo.add(dA, dB, L[:n, n:])
            selb = (~has) & (sign == s)
            if np.any(selb):
                def bnd(sm, k=k):
                    ldg.assemble_boundary_face_vector_flux_operator(
                        sm, basis, q1, normals[f], k, negboundarypenalty,
                        posboundarypenalty, V0, fdj[f], n1)
                d = _celldofs(cid[selb], nf, 3)
                coo.add(d, d, _local_dense(bnd, nf, 3))
    return coo.csr()


def ldg_rhs(mesh, order, numqp, negboundarypenalty, posboundarypenalty, V0, f_qp, g_qp):
    """Vectorised `assemble_LDG_rhs!`; same sampling convention as `ip_rhs`."""
    basis = LagrangeTensorProductBasis(2, order)
    nf = basis.nf
    cq, fq = CellQuadratures(numqp), FaceQuadratures(numqp)
    detjac = mesh.determinant_jacobian()
    fdj = mesh.face_determinant_jacobian()
    normals = reference_face_normals()
    nbr = _neighbours(mesh)
    V0 = np.asarray(V0, dtype=float)
    b = np.zeros((mesh.ncells, nf, 3))
    pts, w = _quad_arrays(cq[1, 0])
    V = np.array([basis(p) for p in pts])
    b[:, :, 0] += (f_qp * (w * detjac)[None, :]) @ V
    for f in range(4):
        pts, w = _quad_arrays(fq[1, f, 0])
        Vf = np.array([basis(p) for p in pts])
        isb = nbr[f] < 0
        g = np.where(isb[:, None], g_qp[f], 0.0)
        gv = (g * (w * fdj[f])[None, :]) @ Vf  # sum_q V g sa w
        n = normals[f]
        b[:, :, 1] += n[0] * gv
        b[:, :, 2] += n[1] * gv
        alpha = negboundarypenalty if float(V0 @ n) < 0 else posboundarypenalty
        nnp = float(-n @ n)
        b[:, :, 0] += -alpha * nnp * gv
    return b.ravel()


# ------------------------------------------------------------ sampling helpers
def sample_cell_quadrature(mesh, numqp, func):
    """func(x, y) vectorised -> (ncells, numqp^2) at the cell quadrature points."""
    pts, _ = _quad_arrays(CellQuadratures(numqp)[1, 0])
    i, j = np.divmod(np.arange(mesh.ncells), mesh.ny)
    x = mesh.x0[0] + mesh.h[0] * (i[:, None] + 0.5 * (1.0 + pts[None, :, 0]))
    y = mesh.x0[1] + mesh.h[1] * (j[:, None] + 0.5 * (1.0 + pts[None, :, 1]))
    return func(x, y)


def sample_face_quadrature(mesh, numqp, func):
    """func(x, y) vectorised -> (4, ncells, numqp) at the face quadrature points."""
    fq = FaceQuadratures(numqp)
    i, j = np.divmod(np.arange(mesh.ncells), mesh.ny)
    out = np.zeros((4, mesh.ncells, numqp))
    for f in range(4):
        pts, _ = _quad_arrays(fq[1, f, 0])
        x = mesh.x0[0] + mesh.h[0] * (i[:, None] + 0.5 * (1.0 + pts[None, :, 0]))
        y = mesh.x0[1] + mesh.h[1] * (j[:, None] + 0.5 * (1.0 + pts[None, :, 1]))
        out[f] = func(x, y)
    return out


def nodal_coordinates(mesh, order):
    """(ncells*NF,) x and y of the DG nodes in the reference's dof order."""
    basis = LagrangeTensorProductBasis(2, order)
    rp = basis.interpolation_points()
    i, j = np.divmod(np.arange(mesh.ncells), mesh.ny)
    x = mesh.x0[0] + mesh.h[0] * (i[:, None] + 0.5 * (1.0 + rp[0][None, :]))
    y = mesh.x0[1] + mesh.h[1] * (j[:, None] + 0.5 * (1.0 + rp[1][None, :]))
    return x.ravel(), y.ravel()


def circle_phase(mesh, centre=(0.5, 0.5), radius=0.4):
    """sign = -1 for cells whose centre lies inside the circle."""
    cx, cy = mesh.cell_centres()
    inside = (cx - centre[0]) ** 2 + (cy - centre[1]) ** 2 < radius ** 2
    return np.where(inside, -1, 1).astype(np.int8)
