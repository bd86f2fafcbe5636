"""Local-DG heat operator (3 dofs per node: T, q_x, q_y): literal restatement
(oracle, test-only).

Follows `/root/reference/StefanDG/LocalDG/`: LDG_utilities.jl, LDG_div_operator.jl,
LDG_mass_operator.jl, LDG_gradient_operator.jl, LDG_scalar_flux_operator.jl,
LDG_vector_flux_operator.jl, LDG_source_term.jl, LDG_assembly.jl, for meshes
without cut cells.  Dof components are 0-based here: 0 = T, 1 = q_x, 2 = q_y
(Julia: 1, 2, 3).

Quirk reproduced on purpose: `assemble_LDG_linear_system!` passes the
unit vector `V0` positionally as `beta` to the vector-flux drivers
(LDG_assembly.jl:44-53,65-74), so the vector flux uses `bn = n . V0`, while the
scalar flux uses `beta . n = 0.5 sign(V0 . n)`.  The stored golden CSV is only
reproduced with this literal behaviour.
"""
import numpy as np

from .basis import (
    gradient,
    interpolation_matrix,
    number_of_basis_functions,
    transform_gradient,
)
from .mesh import (
    assemble_cell_rhs,
    assemble_couple_cell_matrix,
    check_cellsign,
    opposite_face,
    reference_face_normals,
    vec,
)

T_DOF = [0]
Q_DOFS = [1, 2]


def edge_switch(v0, normals):
    """LDG_utilities.jl:1-4"""
    s = np.sign(np.sum(np.asarray(v0).reshape(2, 1) * normals, axis=0, keepdims=True))
    return 0.5 * (s * normals)


# --------------------------------------------------------------- cell terms
def divergence_operator(basis, quad, jacobian, detjac):
    """LDG_div_operator.jl:1-11"""
    nf = number_of_basis_functions(basis)
    matrix = np.zeros((2 * nf, nf))
    for p, w in quad:
        G = transform_gradient(gradient(basis, p), jacobian)
        DivG = vec(G.T)
        V = basis(p)
        matrix += np.outer(DivG, V) * detjac * w
    return matrix


def mass_matrix(basis, quad, ndofs, detjac):
    """`CutCellDG.mass_matrix(basis, quad, nd, detjac)` = sum w N'N detjac."""
    nf = number_of_basis_functions(basis)
    matrix = np.zeros((ndofs * nf, ndofs * nf))
    for p, w in quad:
        N = interpolation_matrix(basis(p), ndofs)
        matrix += (N.T @ N) * detjac * w
    return matrix


def gradient_operator(basis, quad, conductivity, jacobian, detjac):
    """LDG_gradient_operator.jl:1-11"""
    nf = number_of_basis_functions(basis)
    matrix = np.zeros((nf, 2 * nf))
    for p, w in quad:
        G = transform_gradient(gradient(basis, p), jacobian)
        V = interpolation_matrix(basis(p), 2)
        matrix += conductivity * (G @ V) * detjac * w
    return matrix


def _phases(mesh, cellid, k1, k2):
    s = mesh.cell_sign(cellid)
    check_cellsign(s)
    out = []
    if s == +1 or s == 0:
        out.append((+1, k1))
    if s == -1 or s == 0:
        out.append((-1, k2))
    return out


def assemble_divergence_operator(sysmatrix, basis, cellquads, mesh):
    """LDG_div_operator.jl:17-70"""
    jac = mesh.jacobian()
    detjac = mesh.determinant_jacobian()
    for cellid in range(mesh.number_of_cells()):
        for cellsign, _ in _phases(mesh, cellid, None, None):
            quad = cellquads[cellsign, cellid]
            cellmatrix = vec(divergence_operator(basis, quad, jac, detjac))
            nodeids = mesh.nodal_connectivity(cellsign, cellid)
            assemble_couple_cell_matrix(sysmatrix, nodeids, nodeids, 3, cellmatrix,
                                        Q_DOFS, T_DOF)


def assemble_mass_operator(sysmatrix, basis, cellquads, mesh):
    """LDG_mass_operator.jl:1-53"""
    detjac = mesh.determinant_jacobian()
    for cellid in range(mesh.number_of_cells()):
        for cellsign, _ in _phases(mesh, cellid, None, None):
            quad = cellquads[cellsign, cellid]
            cellmatrix = vec(mass_matrix(basis, quad, 2, detjac))
            nodeids = mesh.nodal_connectivity(cellsign, cellid)
            assemble_couple_cell_matrix(sysmatrix, nodeids, nodeids, 3, cellmatrix,
                                        Q_DOFS, Q_DOFS)


def assemble_gradient_operator(sysmatrix, basis, cellquads, k1, k2, mesh):
    """LDG_gradient_operator.jl:13-68"""
    jac = mesh.jacobian()
    detjac = mesh.determinant_jacobian()
    for cellid in range(mesh.number_of_cells()):
        for cellsign, k in _phases(mesh, cellid, k1, k2):
            quad = cellquads[cellsign, cellid]
            cellmatrix = vec(gradient_operator(basis, quad, k, jac, detjac))
            nodeids = mesh.nodal_connectivity(cellsign, cellid)
            assemble_couple_cell_matrix(sysmatrix, nodeids, nodeids, 3, cellmatrix,
                                        T_DOF, Q_DOFS)


# ------------------------------------------------------------- scalar flux
def scalar_flux_operator(basis, quad1, quad2, normals, scaleareas):
    """LDG_scalar_flux_operator.jl:1-23"""
    numqp = len(quad1)
    assert len(quad2) == numqp
    assert normals.shape == (2, numqp)
    assert len(scaleareas) == numqp
    nf = number_of_basis_functions(basis)
    matrix = np.zeros((2 * nf, nf))
    for idx in range(numqp):
        p1, w1 = quad1[idx]
        p2, w2 = quad2[idx]
        assert np.isclose(w1, w2)
        n = normals[:, idx]
        a = scaleareas[idx]
        VL = interpolation_matrix(basis(p1), 2)
        VR = basis(p2)
        matrix += np.outer(VL.T @ n, VR) * a * w1
    return matrix


def assemble_face_scalar_flux_operator(sysmatrix, basis, quad1, quad2, normals, V0,
                                       scaleareas, nodeids1, nodeids2):
    """LDG_scalar_flux_operator.jl:25-107"""
    M11 = -0.5 * vec(scalar_flux_operator(basis, quad1, quad1, normals, scaleareas))
    assemble_couple_cell_matrix(sysmatrix, nodeids1, nodeids1, 3, M11, Q_DOFS, T_DOF)
    M12 = -0.5 * vec(scalar_flux_operator(basis, quad1, quad2, normals, scaleareas))
    assemble_couple_cell_matrix(sysmatrix, nodeids1, nodeids2, 3, M12, Q_DOFS, T_DOF)

    beta = edge_switch(V0, normals)
    bp = np.sum(beta * normals, axis=0)
    bm = -bp
    N11 = -1.0 * vec(scalar_flux_operator(basis, quad1, quad1, normals, bp * scaleareas))
    assemble_couple_cell_matrix(sysmatrix, nodeids1, nodeids1, 3, N11, Q_DOFS, T_DOF)
    N12 = -1.0 * vec(scalar_flux_operator(basis, quad1, quad2, normals, bm * scaleareas))
    assemble_couple_cell_matrix(sysmatrix, nodeids1, nodeids2, 3, N12, Q_DOFS, T_DOF)


def _interior_faces_from_each_cell(mesh, facequads, k1, k2):
    """Loop structure shared by LDG_scalar_flux_operator.jl:140-237 and
    LDG_vector_flux_operator.jl:192-300: EVERY cell assembles its own rows for
    every face with a neighbour of the same phase (`nbrcellid != 0`)."""
    faceids = range(facequads.number_of_faces_per_cell())
    nbrfaceids = [opposite_face(f) for f in faceids]
    for cellid in range(mesh.number_of_cells()):
        for cellsign, k in _phases(mesh, cellid, k1, k2):
            nodeids1 = mesh.nodal_connectivity(cellsign, cellid)
            for faceid in faceids:
                quad1 = facequads[cellsign, faceid, cellid]
                nbrcellid = mesh.cell_connectivity(faceid, cellid)
                if (nbrcellid >= 0
                        and mesh.solution_cell_id(cellsign, nbrcellid)
                        != mesh.solution_cell_id(cellsign, cellid)):
                    nbrcellsign = mesh.cell_sign(nbrcellid)
                    if nbrcellsign == cellsign or nbrcellsign == 0:
                        quad2 = facequads[cellsign, nbrfaceids[faceid], nbrcellid]
                        nodeids2 = mesh.nodal_connectivity(cellsign, nbrcellid)
                        yield faceid, k, quad1, quad2, nodeids1, nodeids2


def _const_face_data(normal, facedetjac, numqp):
    normals = np.repeat(np.asarray(normal, dtype=float).reshape(2, 1), numqp, axis=1)
    scaleareas = np.repeat(facedetjac, numqp)
    return normals, scaleareas


def assemble_interelement_scalar_flux_operator(sysmatrix, basis, facequads, V0, mesh):
    """LDG_scalar_flux_operator.jl:111-237"""
    facedetjac = mesh.face_determinant_jacobian()
    normals_ref = reference_face_normals()
    for faceid, _, quad1, quad2, n1, n2 in _interior_faces_from_each_cell(
            mesh, facequads, None, None):
        normals, scaleareas = _const_face_data(normals_ref[faceid], facedetjac[faceid],
                                               len(quad1))
        assemble_face_scalar_flux_operator(sysmatrix, basis, quad1, quad2, normals, V0,
                                           scaleareas, n1, n2)


# ------------------------------------------------------------- vector flux
def vector_flux_operator(basis, quad1, quad2, normals, scaleareas):
    """LDG_vector_flux_operator.jl:1-24"""
    numqp = len(quad1)
    assert len(quad2) == numqp
    assert normals.shape == (2, numqp)
    assert len(scaleareas) == numqp
    nf = number_of_basis_functions(basis)
    matrix = np.zeros((nf, 2 * nf))
    for idx in range(numqp):
        p1, w1 = quad1[idx]
        p2, w2 = quad2[idx]
        assert np.isclose(w1, w2)
        n = normals[:, idx]
        a = scaleareas[idx]
        V = basis(p1)
        V2 = interpolation_matrix(basis(p2), 2)
        matrix += np.outer(V, n @ V2) * a * w1
    return matrix


def mass_operator(basis, quad1, quad2, ndofs, scaleareas):
    """LDG_vector_flux_operator.jl:26-45"""
    numqp = len(quad1)
    assert len(quad2) == numqp
    assert len(scaleareas) == numqp
    nf = number_of_basis_functions(basis)
    matrix = np.zeros((ndofs * nf, ndofs * nf))
    for idx in range(numqp):
        p1, w1 = quad1[idx]
        p2, _ = quad2[idx]
        V1 = interpolation_matrix(basis(p1), ndofs)
        V2 = interpolation_matrix(basis(p2), ndofs)
        matrix += (V1.T @ V2) * scaleareas[idx] * w1
    return matrix


def assemble_face_vector_flux_operator(sysmatrix, basis, quad1, quad2, normals,
                                       conductivity1, conductivity2, alpha, beta,
                                       scaleareas, nodeids1, nodeids2):
    """LDG_vector_flux_operator.jl:47-154"""
    M11 = -0.5 * conductivity1 * vec(
        vector_flux_operator(basis, quad1, quad1, normals, scaleareas))
    assemble_couple_cell_matrix(sysmatrix, nodeids1, nodeids1, 3, M11, T_DOF, Q_DOFS)
    M12 = -0.5 * conductivity2 * vector_flux_operator(basis, quad1, quad2, normals, scaleareas)
    assemble_couple_cell_matrix(sysmatrix, nodeids1, nodeids2, 3, M12, T_DOF, Q_DOFS)

    posnormals = normals
    negnormals = -normals
    pp = np.sum(posnormals * posnormals, axis=0)
    np_ = np.sum(negnormals * posnormals, axis=0)

    N11 = alpha * vec(mass_operator(basis, quad1, quad1, 1, pp * scaleareas))
    assemble_couple_cell_matrix(sysmatrix, nodeids1, nodeids1, 3, N11, T_DOF, T_DOF)
    N12 = alpha * vec(mass_operator(basis, quad1, quad2, 1, np_ * scaleareas))
    assemble_couple_cell_matrix(sysmatrix, nodeids1, nodeids2, 3, N12, T_DOF, T_DOF)

    bn = posnormals.T @ np.asarray(beta, dtype=float)
    L11 = conductivity1 * vec(
        vector_flux_operator(basis, quad1, quad1, posnormals, bn * scaleareas))
    assemble_couple_cell_matrix(sysmatrix, nodeids1, nodeids1, 3, L11, T_DOF, Q_DOFS)
    L12 = conductivity2 * vec(
        vector_flux_operator(basis, quad1, quad2, negnormals, bn * scaleareas))
    assemble_couple_cell_matrix(sysmatrix, nodeids1, nodeids2, 3, L12, T_DOF, Q_DOFS)


def assemble_interelement_vector_flux_operator(sysmatrix, basis, facequads, k1, k2, alpha,
                                               beta, mesh):
    """LDG_vector_flux_operator.jl:158-300"""
    facedetjac = mesh.face_determinant_jacobian()
    normals_ref = reference_face_normals()
    for faceid, k, quad1, quad2, n1, n2 in _interior_faces_from_each_cell(
            mesh, facequads, k1, k2):
        normals, scaleareas = _const_face_data(normals_ref[faceid], facedetjac[faceid],
                                               len(quad1))
        assemble_face_vector_flux_operator(sysmatrix, basis, quad1, quad2, normals, k, k,
                                           alpha, beta, scaleareas, n1, n2)


def assemble_boundary_face_vector_flux_operator(sysmatrix, basis, quad, normal,
                                                conductivity, negboundarypenalty,
                                                posboundarypenalty, V0, facedetjac,
                                                nodeids):
    """LDG_vector_flux_operator.jl:391-439"""
    numqp = len(quad)
    V0nk = float(np.dot(V0, normal))
    alpha = negboundarypenalty if V0nk < 0 else posboundarypenalty
    if numqp > 0:
        posnormals, scaleareas = _const_face_data(normal, facedetjac, numqp)
        M11 = -conductivity * vec(
            vector_flux_operator(basis, quad, quad, posnormals, scaleareas))
        assemble_couple_cell_matrix(sysmatrix, nodeids, nodeids, 3, M11, T_DOF, Q_DOFS)
        nnp = np.sum(posnormals * posnormals, axis=0)
        N11 = alpha * vec(mass_operator(basis, quad, quad, 1, nnp * scaleareas))
        assemble_couple_cell_matrix(sysmatrix, nodeids, nodeids, 3, N11, T_DOF, T_DOF)


def assemble_boundary_vector_flux_operator(sysmatrix, basis, facequads, k1, k2,
                                           negboundarypenalty, posboundarypenalty, V0, mesh):
    """LDG_vector_flux_operator.jl:441-535"""
    facedetjac = mesh.face_determinant_jacobian()
    faceids = range(facequads.number_of_faces_per_cell())
    normals = reference_face_normals()
    for cellid in range(mesh.number_of_cells()):
        for cellsign, k in _phases(mesh, cellid, k1, k2):
            nodeids = mesh.nodal_connectivity(cellsign, cellid)
            for faceid in faceids:
                quad = facequads[cellsign, faceid, cellid]
                if mesh.cell_connectivity(faceid, cellid) < 0:
                    assemble_boundary_face_vector_flux_operator(
                        sysmatrix, basis, quad, normals[faceid], k, negboundarypenalty,
                        posboundarypenalty, V0, facedetjac[faceid], nodeids)


# --------------------------------------------------------------------- rhs
def linear_form(rhsfunc, basis, quad, cellmap, detjac):
    """LDG_source_term.jl:1-11"""
    rhs = np.zeros(number_of_basis_functions(basis))
    for p, w in quad:
        rhs += basis(p) * rhsfunc(cellmap(p)) * detjac * w
    return rhs


def assemble_source(sysrhs, rhsfunc, basis, cellquads, mesh):
    """LDG_source_term.jl:14-60"""
    detjac = mesh.determinant_jacobian()
    for cellid in range(mesh.number_of_cells()):
        for cellsign, _ in _phases(mesh, cellid, None, None):
            cellmap = mesh.cell_map(cellsign, cellid)
            quad = cellquads[cellsign, cellid]
            rhs = linear_form(rhsfunc, basis, quad, cellmap, detjac)
            nodeids = mesh.nodal_connectivity(cellsign, cellid)
            assemble_cell_rhs(sysrhs, nodeids, 3, rhs, T_DOF)


def boundary_face_scalar_flux(rhsfunc, basis, quad, normals, cellmap, scaleareas):
    """LDG_source_term.jl:64-89"""
    numqp = len(quad)
    assert normals.shape == (2, numqp)
    assert len(scaleareas) == numqp
    rhs = np.zeros(2 * number_of_basis_functions(basis))
    for idx, (p, w) in enumerate(quad):
        V2 = interpolation_matrix(basis(p), 2)
        gd = rhsfunc(cellmap(p))
        rhs += (V2.T @ normals[:, idx]) * gd * scaleareas[idx] * w
    return rhs


def assemble_boundary_scalar_flux_source(sysrhs, rhsfunc, basis, facequads, mesh):
    """LDG_source_term.jl:91-198"""
    facedetjac = mesh.face_determinant_jacobian()
    normals_ref = reference_face_normals()
    faceids = range(facequads.number_of_faces_per_cell())
    for cellid in range(mesh.number_of_cells()):
        for cellsign, _ in _phases(mesh, cellid, None, None):
            nodeids = mesh.nodal_connectivity(cellsign, cellid)
            cellmap = mesh.cell_map(cellsign, cellid)
            for faceid in faceids:
                quad = facequads[cellsign, faceid, cellid]
                if mesh.cell_connectivity(faceid, cellid) < 0:
                    normals, scaleareas = _const_face_data(
                        normals_ref[faceid], facedetjac[faceid], len(quad))
                    rhs = boundary_face_scalar_flux(rhsfunc, basis, quad, normals,
                                                    cellmap, scaleareas)
                    assemble_cell_rhs(sysrhs, nodeids, 3, rhs, Q_DOFS)


def assemble_boundary_vector_flux_source(sysrhs, rhsfunc, basis, facequads,
                                         negboundarypenalty, posboundarypenalty, V0, mesh):
    """LDG_source_term.jl:203-321"""
    facedetjac = mesh.face_determinant_jacobian()
    faceids = range(facequads.number_of_faces_per_cell())
    normals = reference_face_normals()
    for cellid in range(mesh.number_of_cells()):
        for cellsign, _ in _phases(mesh, cellid, None, None):
            nodeids = mesh.nodal_connectivity(cellsign, cellid)
            cellmap = mesh.cell_map(cellsign, cellid)
            for faceid in faceids:
                quad = facequads[cellsign, faceid, cellid]
                if mesh.cell_connectivity(faceid, cellid) < 0:
                    normal = normals[faceid]
                    V0nk = float(np.dot(V0, normal))
                    alpha = negboundarypenalty if V0nk < 0 else posboundarypenalty
                    nnp = float(np.dot(-normal, normal))
                    rhs = -alpha * linear_form(rhsfunc, basis, quad, cellmap,
                                               nnp * facedetjac[faceid])
                    assemble_cell_rhs(sysrhs, nodeids, 3, rhs, T_DOF)


# --------------------------------------------- mesh-aligned interface (extension)
def _aligned_interface_faces_from_each_cell(mesh, facequads, k1, k2):
    """Faces between UNCUT cells of unlike sign, seen from each of the two cells (the loop structure of
    assemble_cell_interface_scalar_flux_operator!, LDG_scalar_flux_operator.jl:241-283, which applies the face
    function once per orientation).  The reference's drivers couple phases only inside cut cells; on a mesh-aligned
    interface the same FACE-LEVEL functions are applied to these faces (as for IP).
    Yields (faceid, k_own, k_nbr, quad1, quad2, nodeids1, nodeids2)."""
    for cellid in range(mesh.number_of_cells()):
        s1 = mesh.cell_sign(cellid)
        for faceid in range(facequads.number_of_faces_per_cell()):
            nbr = mesh.cell_connectivity(faceid, cellid)
            if nbr < 0:
                continue
            s2 = mesh.cell_sign(nbr)
            if s1 == s2 or s1 == 0 or s2 == 0:
                continue
            yield (faceid, k1 if s1 == 1 else k2, k1 if s2 == 1 else k2,
                   facequads[s1, faceid, cellid], facequads[s2, opposite_face(faceid), nbr],
                   mesh.nodal_connectivity(s1, cellid), mesh.nodal_connectivity(s2, nbr))


def assemble_aligned_interface_scalar_flux_operator(sysmatrix, basis, facequads, V0, mesh):
    """assemble_face_scalar_flux_operator! (LDG_scalar_flux_operator.jl:25-107) on unlike-phase faces, both
    orientations (as :261-282 does for the two sides of a cut cell's interface)."""
    facedetjac = mesh.face_determinant_jacobian()
    normals_ref = reference_face_normals()
    for faceid, _, _, quad1, quad2, n1, n2 in _aligned_interface_faces_from_each_cell(mesh, facequads, None, None):
        normals, scaleareas = _const_face_data(normals_ref[faceid], facedetjac[faceid], len(quad1))
        assemble_face_scalar_flux_operator(sysmatrix, basis, quad1, quad2, normals, V0, scaleareas, n1, n2)


def assemble_aligned_interface_vector_flux_operator(sysmatrix, basis, facequads, k1, k2, alpha, beta, mesh):
    """assemble_face_vector_flux_operator! (LDG_vector_flux_operator.jl:47-154) on unlike-phase faces with the own
    cell's conductivity first and the neighbour's second, alpha = interfacepenalty (as :318-355)."""
    facedetjac = mesh.face_determinant_jacobian()
    normals_ref = reference_face_normals()
    for faceid, ka, kb, quad1, quad2, n1, n2 in _aligned_interface_faces_from_each_cell(mesh, facequads, k1, k2):
        normals, scaleareas = _const_face_data(normals_ref[faceid], facedetjac[faceid], len(quad1))
        assemble_face_vector_flux_operator(sysmatrix, basis, quad1, quad2, normals, ka, kb, alpha, beta,
                                           scaleareas, n1, n2)


# ---------------------------------------------------------------- top level
def assemble_LDG_linear_system(sysmatrix, solverbasis, cellquads, facequads, interfacequads,
                               k1, k2, interiorpenalty, interfacepenalty,
                               negboundarypenalty, posboundarypenalty, V0, mesh):
    """LDG_assembly.jl:1-91 (`interfacequads` unused: no cut cells)."""
    assemble_divergence_operator(sysmatrix, solverbasis, cellquads, mesh)
    assemble_mass_operator(sysmatrix, solverbasis, cellquads, mesh)
    assemble_gradient_operator(sysmatrix, solverbasis, cellquads, k1, k2, mesh)
    assemble_interelement_scalar_flux_operator(sysmatrix, solverbasis, facequads, V0, mesh)
    # V0 passed positionally as `beta` (LDG_assembly.jl:51)
    assemble_interelement_vector_flux_operator(sysmatrix, solverbasis, facequads, k1, k2,
                                               interiorpenalty, V0, mesh)
    assemble_boundary_vector_flux_operator(sysmatrix, solverbasis, facequads, k1, k2,
                                           negboundarypenalty, posboundarypenalty, V0, mesh)


def assemble_LDG_rhs(sysrhs, sourceterm, boundaryfunc, solverbasis, cellquads, facequads,
                     negboundarypenalty, posboundarypenalty, V0, mesh):
    """LDG_assembly.jl:93-131"""
    assemble_boundary_scalar_flux_source(sysrhs, boundaryfunc, solverbasis, facequads, mesh)
    assemble_boundary_vector_flux_source(sysrhs, boundaryfunc, solverbasis, facequads,
                                         negboundarypenalty, posboundarypenalty, V0, mesh)
    assemble_source(sysrhs, sourceterm, solverbasis, cellquads, mesh)
