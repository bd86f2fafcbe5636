"""TEST INFRASTRUCTURE (oracle): the local-DG element-block path restated with the reference's own
cell- and face-level functions (oracle/ldg.py: divergence_operator LDG_div_operator.jl:1-11, mass_matrix,
gradient_operator LDG_gradient_operator.jl:1-11, assemble_face_scalar_flux_operator!
LDG_scalar_flux_operator.jl:25-107, assemble_face_vector_flux_operator! LDG_vector_flux_operator.jl:47-154,
assemble_boundary_face_vector_flux_operator! :391-439, the linear forms of LDG_source_term.jl) for ARBITRARY
per-cell / per-face quadrature data in the cell-centric layout of `stefan_blocks_data`
(include/stefan_b200.h) -- what the cut-cell drivers `assemble_interface_scalar_flux_operator!`
(LDG_scalar_flux_operator.jl:241-309) and `assemble_interface_vector_flux_operator!`
(LDG_vector_flux_operator.jl:318-386) feed those functions on a cut mesh.

Every face function is called exactly as the reference's drivers call it: from the cell's own side (side 1) with
the neighbour as side 2, which adds to the rows of the cell only -- the (1,1) contributions land in the cell's
diagonal block, the (1,2) contributions in the block of that slot.

Dofs of a solution cell: 3 per node, interleaved (T, q_x, q_y); block entry [3 a + i][3 b + j].
`slot_pen` is the penalty alpha of the slot (interior / interface penalty for neighbour slots, the negative- or
positive-side boundary penalty for boundary slots -- chosen by the producer of the data, as
LDG_vector_flux_operator.jl:391-439 does from sign(V0 . n)); `slot_lambda` is not used.

Pinned on uniform meshes through oracle/fast.ldg_matrix (tests/test_cutcell_cpu.py); parity unpinned against the
reference itself for cut-cell data (its geometry lives in un-vendored packages).  Not part of the product."""
import numpy as np

from . import ldg
from .basis import LagrangeTensorProductBasis
from .mesh import SystemMatrix, SystemRHS


def _dense(sysmatrix, n):
    A = np.zeros((n, n))
    r, c, v = sysmatrix.triplets()
    np.add.at(A, (r, c), v)
    return A


def assemble_blocks(order, data, V0):
    """Returns blocks [ncells][1 + nslots][3 NF][3 NF] (diag, then one per slot; zeros for empty slots)."""
    basis = LagrangeTensorProductBasis(2, order)
    nf = basis.nf
    n3 = 3 * nf
    ncells, nslots = data["slot_nbr"].shape
    out = np.zeros((ncells, 1 + nslots, n3, n3))
    own, other = np.arange(nf), nf + np.arange(nf)
    V0 = np.asarray(V0, float)
    for c in range(ncells):
        jac = data["cell_jac"][c]
        detjac = jac[0] * jac[1]
        k1 = data["cell_k"][c]
        quad = [(p, w) for p, w in zip(data["vol_pts"][c], data["vol_wts"][c]) if w != 0.0]
        sm = SystemMatrix()
        ldg.assemble_couple_cell_matrix(sm, own, own, 3, ldg.vec(ldg.divergence_operator(basis, quad, jac, detjac)),
                                        ldg.Q_DOFS, ldg.T_DOF)
        ldg.assemble_couple_cell_matrix(sm, own, own, 3, ldg.vec(ldg.mass_matrix(basis, quad, 2, detjac)),
                                        ldg.Q_DOFS, ldg.Q_DOFS)
        ldg.assemble_couple_cell_matrix(sm, own, own, 3, ldg.vec(ldg.gradient_operator(basis, quad, k1, jac, detjac)),
                                        ldg.T_DOF, ldg.Q_DOFS)
        out[c, 0] += _dense(sm, 2 * n3)[:n3, :n3]
        for s in range(nslots):
            nbr = int(data["slot_nbr"][c, s])
            if nbr < -1:
                continue
            keep = data["face_wts"][c, s] != 0.0
            w = data["face_wts"][c, s][keep]
            if len(w) == 0:
                continue
            q1 = list(zip(data["face_pts"][c, s][keep], w))
            normals = data["face_normals"][c, s][keep].T
            ones = np.ones(len(q1))
            alpha = data["slot_pen"][c, s]
            sm = SystemMatrix()
            if nbr == -1:
                # (constant normal of a mesh face; the scale area is folded into the weights: facedetjac = 1)
                M11 = -k1 * ldg.vec(ldg.vector_flux_operator(basis, q1, q1, normals, ones))
                ldg.assemble_couple_cell_matrix(sm, own, own, 3, M11, ldg.T_DOF, ldg.Q_DOFS)
                nnp = np.sum(normals * normals, axis=0)
                N11 = alpha * ldg.vec(ldg.mass_operator(basis, q1, q1, 1, nnp * ones))
                ldg.assemble_couple_cell_matrix(sm, own, own, 3, N11, ldg.T_DOF, ldg.T_DOF)
                out[c, 0] += _dense(sm, 2 * n3)[:n3, :n3]
                continue
            q2 = list(zip(data["face_nbr_pts"][c, s][keep], w))
            k2 = data["cell_k"][nbr]
            ldg.assemble_face_scalar_flux_operator(sm, basis, q1, q2, normals, V0, ones, own, other)
            ldg.assemble_face_vector_flux_operator(sm, basis, q1, q2, normals, k1, k2, alpha, V0, ones, own, other)
            A = _dense(sm, 2 * n3)
            assert not A[n3:].any()
            out[c, 0] += A[:n3, :n3]
            out[c, 1 + s] += A[:n3, n3:]
    return out


def dense_matrix(blocks, slot_nbr):
    ncells, nb, n3, _ = blocks.shape
    A = np.zeros((ncells * n3, ncells * n3))
    for c in range(ncells):
        A[c * n3:(c + 1) * n3, c * n3:(c + 1) * n3] += blocks[c, 0]
        for s in range(nb - 1):
            n = int(slot_nbr[c, s])
            if n >= 0:
                A[c * n3:(c + 1) * n3, n * n3:(n + 1) * n3] += blocks[c, 1 + s]
    return A


def assemble_rhs(order, data, f_qp, g_qp):
    """assemble_LDG_rhs! (LDG_assembly.jl:93-131) for block data: volume source `linear_form` into the T rows
    (LDG_source_term.jl:1-60), on boundary slots `boundary_face_scalar_flux` into the q rows (:64-198) and
    -alpha (-n . n) linear_form = +alpha sum V g w into the T rows (:203-321).
    f_qp [ncells][nqv]: source at the volume points; g_qp [ncells][nslots][nqf]: boundary data."""
    basis = LagrangeTensorProductBasis(2, order)
    nf = basis.nf
    ncells, nslots = data["slot_nbr"].shape
    ident = lambda p: p  # noqa: E731
    rhs = np.zeros((ncells, 3 * nf))
    own = np.arange(nf)
    for c in range(ncells):
        jac = data["cell_jac"][c]
        sr = SystemRHS()
        for p, w, fv in zip(data["vol_pts"][c], data["vol_wts"][c], f_qp[c]):
            if w != 0.0:
                ldg.assemble_cell_rhs(sr, own, 3, ldg.linear_form(lambda _p, fv=fv: fv, basis, [(p, w)], ident,
                                                                  jac[0] * jac[1]), ldg.T_DOF)
        for s in range(nslots):
            if int(data["slot_nbr"][c, s]) != -1:
                continue
            keep = data["face_wts"][c, s] != 0.0
            alpha = data["slot_pen"][c, s]
            for p, wq, n, gq in zip(data["face_pts"][c, s][keep], data["face_wts"][c, s][keep],
                                    data["face_normals"][c, s][keep], g_qp[c, s][keep]):
                g = lambda _p, gq=gq: gq  # noqa: E731
                q1 = [(p, wq)]
                ldg.assemble_cell_rhs(sr, own, 3, ldg.boundary_face_scalar_flux(g, basis, q1, n.reshape(2, 1), ident,
                                                                                np.ones(1)), ldg.Q_DOFS)
                nnp = float(np.dot(-n, n))
                ldg.assemble_cell_rhs(sr, own, 3, -alpha * ldg.linear_form(g, basis, q1, ident, nnp), ldg.T_DOF)
        for r, v in zip(sr.rows, sr.vals):
            np.add.at(rhs[c], r, v)
    return rhs.ravel()


def boundary_penalties(data, V0, negboundarypenalty, posboundarypenalty, interiorpenalty, interfacepenalty=None,
                       slot_is_interface=None):
    """slot_pen for local-DG block data: interior / interface penalty on neighbour slots, on boundary slots the
    penalty selected by sign(V0 . n) (LDG_vector_flux_operator.jl:391-439, LDG_source_term.jl:203-321)."""
    nbr = data["slot_nbr"]
    pen = np.zeros(nbr.shape)
    pen[nbr >= 0] = interiorpenalty
    if slot_is_interface is not None and interfacepenalty is not None:
        pen[(nbr >= 0) & slot_is_interface] = interfacepenalty
    n0 = data["face_normals"][:, :, 0, :]
    v0n = n0 @ np.asarray(V0, float)
    pen[nbr == -1] = np.where(v0n < 0, negboundarypenalty, posboundarypenalty)[nbr == -1]
    return pen
