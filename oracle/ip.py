"""Interior-penalty DG heat operator: literal restatement (oracle, test-only).

Follows `/root/reference/StefanDG/InteriorPenalty/`:
  IP_gradient_operator.jl, IP_flux_operator.jl, IP_penalty_operator.jl,
  IP_robin_interface_operator.jl, IP_robin_interface_source.jl,
  IP_source_term.jl, IP_assembly.jl
for meshes without cut cells (`cell_sign` = +1 -> k1, -1 -> k2).

Two things the reference files do not contain:
  * `variant="legacy_boundary"`: the boundary treatment that reproduces the stored
    `uncut_mesh_tests/IP_convergence.csv`: boundary operator `-M'`
    (no symmetrising term) and boundary rhs = penalty term only.
  * mesh-aligned two-phase interface (`assemble_aligned_interface_*`): the
    reference only couples unlike phases inside cut cells.  Here the face between
    two uncut cells of unlike sign is handed to the reference's own FACE-level
    functions exactly as `assemble_cell_interface_flux_operator!`
    (IP_flux_operator.jl:261-294) does: side 1 = the +1 cell (k1), side 2 = the -1
    cell (k2), normal = outward normal of side 1, scale area = face det-jacobian.
"""
import numpy as np

from .basis import gradient, number_of_basis_functions, transform_gradient
from .mesh import (
    assemble_cell_matrix,
    assemble_cell_rhs,
    assemble_couple_cell_matrix,
    check_cellsign,
    opposite_face,
    reference_face_normals,
    vec,
)


# ---------------------------------------------------------------- volume term
def gradient_operator(basis, quad, conductivity, jacobian, detjac):
    """IP_gradient_operator.jl:1-9"""
    nf = number_of_basis_functions(basis)
    matrix = np.zeros((nf, nf))
    for p, w in quad:
        G = transform_gradient(gradient(basis, p), jacobian)
        matrix += conductivity * (G @ G.T) * detjac * w
    return matrix


def assemble_cut_cell_gradient_operator(sysmatrix, basis, cellquads, conductivity,
                                        mesh, cellsign, cellid):
    """IP_gradient_operator.jl:11-27"""
    jac = mesh.jacobian()
    detjac = mesh.determinant_jacobian()
    quad = cellquads[cellsign, cellid]
    cellmatrix = vec(gradient_operator(basis, quad, conductivity, jac, detjac))
    nodeids = mesh.nodal_connectivity(cellsign, cellid)
    assemble_cell_matrix(sysmatrix, nodeids, 1, cellmatrix)


def assemble_gradient_operator(sysmatrix, basis, cellquads, k1, k2, mesh):
    """IP_gradient_operator.jl:29-57"""
    for cellid in range(mesh.number_of_cells()):
        s = mesh.cell_sign(cellid)
        check_cellsign(s)
        if s == +1 or s == 0:
            assemble_cut_cell_gradient_operator(sysmatrix, basis, cellquads, k1, mesh, +1, cellid)
        if s == -1 or s == 0:
            assemble_cut_cell_gradient_operator(sysmatrix, basis, cellquads, k2, mesh, -1, cellid)


# ------------------------------------------------------------------ flux term
def flux_operator(basis, quad1, quad2, normals, conductivity, jacobian, scaleareas):
    """IP_flux_operator.jl:1-33"""
    numqp = len(quad1)
    assert len(quad2) == numqp
    assert normals.shape == (2, numqp)
    assert len(scaleareas) == numqp
    nf = number_of_basis_functions(basis)
    matrix = np.zeros((nf, nf))
    for idx in range(numqp):
        p1, w1 = quad1[idx]
        n = normals[:, idx]
        p2, w2 = quad2[idx]
        assert np.isclose(w1, w2)
        G = transform_gradient(gradient(basis, p1), jacobian)
        V = basis(p2)
        matrix += conductivity * np.outer(G @ n, V) * scaleareas[idx] * w1
    return matrix


def assemble_face_flux_operator(sysmatrix, basis, quad1, quad2, normals, conductivity1,
                                conductivity2, jacobian, scaleareas, nodeids1, nodeids2):
    """IP_flux_operator.jl:35-118"""
    M11 = -0.5 * flux_operator(basis, quad1, quad1, normals, conductivity1, jacobian, scaleareas)
    M12 = -0.5 * flux_operator(basis, quad1, quad2, normals, conductivity1, jacobian, scaleareas)
    M21 = -0.5 * flux_operator(basis, quad2, quad1, normals, conductivity2, jacobian, scaleareas)
    M22 = -0.5 * flux_operator(basis, quad2, quad2, normals, conductivity2, jacobian, scaleareas)
    assemble_couple_cell_matrix(sysmatrix, nodeids1, nodeids1, 1, vec(M11 + M11.T))
    assemble_couple_cell_matrix(sysmatrix, nodeids2, nodeids1, 1, vec(M21 - M12.T))
    assemble_couple_cell_matrix(sysmatrix, nodeids1, nodeids2, 1, vec(-M12 + M21.T))
    assemble_couple_cell_matrix(sysmatrix, nodeids2, nodeids2, 1, vec(-M22 - M22.T))


def assemble_face_interelement_flux_operator(sysmatrix, basis, quad1, quad2, normal,
                                             conductivity, jacobian, facedetjac,
                                             nodeids1, nodeids2):
    """IP_flux_operator.jl:122-152"""
    numqp = len(quad1)
    normals = np.repeat(np.asarray(normal, dtype=float).reshape(2, 1), numqp, axis=1)
    scaleareas = np.repeat(facedetjac, numqp)
    assemble_face_flux_operator(sysmatrix, basis, quad1, quad2, normals, conductivity,
                                conductivity, jacobian, scaleareas, nodeids1, nodeids2)


def assemble_cut_cell_interelement_flux_operator(sysmatrix, basis, facequads, normals,
                                                 conductivity, mesh, cellsign, cellid,
                                                 faceids, nbrfaceids, jacobian, facedetjac):
    """IP_flux_operator.jl:154-200 (0-based: boundary neighbour is -1, so the
    reference's `cellid < nbrcellid` also needs `nbrcellid >= 0` here)."""
    nodeids1 = mesh.nodal_connectivity(cellsign, cellid)
    for faceid in faceids:
        quad1 = facequads[cellsign, faceid, cellid]
        nbrcellid = mesh.cell_connectivity(faceid, cellid)
        if (nbrcellid >= 0 and cellid < nbrcellid
                and mesh.solution_cell_id(cellsign, nbrcellid)
                != mesh.solution_cell_id(cellsign, cellid)):
            nbrfaceid = nbrfaceids[faceid]
            nbrcellsign = mesh.cell_sign(nbrcellid)
            if nbrcellsign == cellsign or nbrcellsign == 0:
                quad2 = facequads[cellsign, nbrfaceid, nbrcellid]
                nodeids2 = mesh.nodal_connectivity(cellsign, nbrcellid)
                assemble_face_interelement_flux_operator(
                    sysmatrix, basis, quad1, quad2, normals[faceid], conductivity,
                    jacobian, facedetjac[faceid], nodeids1, nodeids2)


def assemble_interelement_flux_operator(sysmatrix, basis, facequads, k1, k2, mesh):
    """IP_flux_operator.jl:202-256"""
    jacobian = mesh.jacobian()
    facedetjac = mesh.face_determinant_jacobian()
    faceids = range(facequads.number_of_faces_per_cell())
    nbrfaceids = [opposite_face(f) for f in faceids]
    normals = reference_face_normals()
    for cellid in range(mesh.number_of_cells()):
        s = mesh.cell_sign(cellid)
        check_cellsign(s)
        if s == +1 or s == 0:
            assemble_cut_cell_interelement_flux_operator(
                sysmatrix, basis, facequads, normals, k1, mesh, +1, cellid, faceids,
                nbrfaceids, jacobian, facedetjac)
        if s == -1 or s == 0:
            assemble_cut_cell_interelement_flux_operator(
                sysmatrix, basis, facequads, normals, k2, mesh, -1, cellid, faceids,
                nbrfaceids, jacobian, facedetjac)


def assemble_boundary_face_flux_operator(sysmatrix, basis, quad, normal, conductivity,
                                         jacobian, facedetjac, nodeids, variant="current"):
    """IP_flux_operator.jl:330-357 (`current`); `legacy_boundary`: -M' only."""
    numqp = len(quad)
    normals = np.repeat(np.asarray(normal, dtype=float).reshape(2, 1), numqp, axis=1)
    scaleareas = np.repeat(facedetjac, numqp)
    M = flux_operator(basis, quad, quad, normals, conductivity, jacobian, scaleareas)
    if variant == "current":
        vM = -vec(M + M.T)
    elif variant == "legacy_boundary":
        vM = -vec(M.T)
    else:
        raise ValueError(variant)
    assemble_couple_cell_matrix(sysmatrix, nodeids, nodeids, 1, vM)


def assemble_boundary_flux_operator(sysmatrix, basis, facequads, k1, k2, mesh,
                                    variant="current"):
    """IP_flux_operator.jl:359-444"""
    jacobian = mesh.jacobian()
    facedetjac = mesh.face_determinant_jacobian()
    faceids = range(facequads.number_of_faces_per_cell())
    normals = reference_face_normals()
    for cellid in range(mesh.number_of_cells()):
        s = mesh.cell_sign(cellid)
        check_cellsign(s)
        for cellsign, k in ((+1, k1), (-1, k2)):
            if s == cellsign or s == 0:
                nodeids = mesh.nodal_connectivity(cellsign, cellid)
                for faceid in faceids:
                    quad = facequads[cellsign, faceid, cellid]
                    if mesh.cell_connectivity(faceid, cellid) < 0:
                        assemble_boundary_face_flux_operator(
                            sysmatrix, basis, quad, normals[faceid], k, jacobian,
                            facedetjac[faceid], nodeids, variant)


# --------------------------------------------------------------- penalty term
def mass_operator(basis, quad1, quad2, scaleareas):
    """IP_penalty_operator.jl:1-18"""
    numqp = len(quad1)
    assert len(quad2) == numqp
    assert len(scaleareas) == numqp
    nf = number_of_basis_functions(basis)
    matrix = np.zeros((nf, nf))
    for idx in range(numqp):
        p1, w1 = quad1[idx]
        p2, _ = quad2[idx]
        matrix += np.outer(basis(p1), basis(p2)) * scaleareas[idx] * w1
    return matrix


def assemble_face_penalty_operator(sysmatrix, basis, quad1, quad2, penalty, scaleareas,
                                   nodeids1, nodeids2):
    """IP_penalty_operator.jl:21-65"""
    M11 = vec(penalty * mass_operator(basis, quad1, quad1, scaleareas))
    M12 = vec(penalty * mass_operator(basis, quad1, quad2, scaleareas))
    M21 = vec(penalty * mass_operator(basis, quad2, quad1, scaleareas))
    M22 = vec(penalty * mass_operator(basis, quad2, quad2, scaleareas))
    assemble_couple_cell_matrix(sysmatrix, nodeids1, nodeids1, 1, M11)
    assemble_couple_cell_matrix(sysmatrix, nodeids1, nodeids2, 1, -M12)
    assemble_couple_cell_matrix(sysmatrix, nodeids2, nodeids1, 1, -M21)
    assemble_couple_cell_matrix(sysmatrix, nodeids2, nodeids2, 1, M22)


def assemble_interelement_penalty_operator(sysmatrix, basis, facequads, penalty, mesh):
    """IP_penalty_operator.jl:67-181"""
    facedetjac = mesh.face_determinant_jacobian()
    faceids = range(facequads.number_of_faces_per_cell())
    nbrfaceids = [opposite_face(f) for f in faceids]
    for cellid in range(mesh.number_of_cells()):
        s = mesh.cell_sign(cellid)
        check_cellsign(s)
        for cellsign in (+1, -1):
            if not (s == cellsign or s == 0):
                continue
            nodeids1 = mesh.nodal_connectivity(cellsign, cellid)
            for faceid in faceids:
                quad1 = facequads[cellsign, faceid, cellid]
                nbrcellid = mesh.cell_connectivity(faceid, cellid)
                if (nbrcellid >= 0 and cellid < nbrcellid
                        and mesh.solution_cell_id(cellsign, nbrcellid)
                        != mesh.solution_cell_id(cellsign, cellid)):
                    nbrcellsign = mesh.cell_sign(nbrcellid)
                    if nbrcellsign == cellsign or nbrcellsign == 0:
                        quad2 = facequads[cellsign, nbrfaceids[faceid], nbrcellid]
                        nodeids2 = mesh.nodal_connectivity(cellsign, nbrcellid)
                        scaleareas = np.repeat(facedetjac[faceid], len(quad1))
                        assemble_face_penalty_operator(
                            sysmatrix, basis, quad1, quad2, penalty, scaleareas,
                            nodeids1, nodeids2)


def assemble_boundary_penalty_operator(sysmatrix, basis, facequads, penalty, mesh):
    """IP_penalty_operator.jl:239-329"""
    facedetjac = mesh.face_determinant_jacobian()
    faceids = range(facequads.number_of_faces_per_cell())
    for cellid in range(mesh.number_of_cells()):
        s = mesh.cell_sign(cellid)
        check_cellsign(s)
        for cellsign in (+1, -1):
            if not (s == cellsign or s == 0):
                continue
            nodeids = mesh.nodal_connectivity(cellsign, cellid)
            for faceid in faceids:
                quad = facequads[cellsign, faceid, cellid]
                if mesh.cell_connectivity(faceid, cellid) < 0:
                    scaleareas = np.repeat(facedetjac[faceid], len(quad))
                    M = vec(penalty * mass_operator(basis, quad, quad, scaleareas))
                    assemble_couple_cell_matrix(sysmatrix, nodeids, nodeids, 1, M)


# ----------------------------------------------------------------- Robin term
def assemble_face_robin_operator(sysmatrix, basis, quad1, quad2, lam, scaleareas,
                                 nodeids1, nodeids2):
    """IP_robin_interface_operator.jl:1-45"""
    M11 = 0.25 * lam * vec(mass_operator(basis, quad1, quad1, scaleareas))
    M12 = 0.25 * lam * vec(mass_operator(basis, quad1, quad2, scaleareas))
    M21 = 0.25 * lam * vec(mass_operator(basis, quad2, quad1, scaleareas))
    M22 = 0.25 * lam * vec(mass_operator(basis, quad2, quad2, scaleareas))
    assemble_couple_cell_matrix(sysmatrix, nodeids1, nodeids1, 1, M11)
    assemble_couple_cell_matrix(sysmatrix, nodeids1, nodeids2, 1, M12)
    assemble_couple_cell_matrix(sysmatrix, nodeids2, nodeids1, 1, M21)
    assemble_couple_cell_matrix(sysmatrix, nodeids2, nodeids2, 1, M22)


def robin_interface_linear_form(basis, quad, scaleareas):
    """IP_robin_interface_source.jl:1-10"""
    assert len(quad) == len(scaleareas)
    rhs = np.zeros(number_of_basis_functions(basis))
    for idx, (p, w) in enumerate(quad):
        rhs += basis(p) * scaleareas[idx] * w
    return rhs


def assemble_face_robin_source(sysrhs, basis, quad, lam, Tm, scaleareas, nodeids):
    """IP_robin_interface_source.jl:12-25"""
    rhs = 0.5 * lam * Tm * robin_interface_linear_form(basis, quad, scaleareas)
    assemble_cell_rhs(sysrhs, nodeids, 1, rhs)


# ------------------------------------ mesh-aligned interface (see module doc)
def aligned_interface_faces(mesh):
    """Yield (cell1, face1, cell2, face2) with cell1 the +1 side, cell2 the -1 side."""
    for cellid in range(mesh.number_of_cells()):
        if mesh.cell_sign(cellid) != +1:
            continue
        for faceid in range(4):
            nbr = mesh.cell_connectivity(faceid, cellid)
            if nbr >= 0 and mesh.cell_sign(nbr) == -1:
                yield cellid, faceid, nbr, opposite_face(faceid)


def assemble_aligned_interface_flux_operator(sysmatrix, basis, facequads, k1, k2, mesh):
    jacobian = mesh.jacobian()
    facedetjac = mesh.face_determinant_jacobian()
    normals_ref = reference_face_normals()
    for c1, f1, c2, f2 in aligned_interface_faces(mesh):
        quad1 = facequads[+1, f1, c1]
        quad2 = facequads[-1, f2, c2]
        numqp = len(quad1)
        normals = np.repeat(normals_ref[f1].reshape(2, 1), numqp, axis=1)
        scaleareas = np.repeat(facedetjac[f1], numqp)
        assemble_face_flux_operator(
            sysmatrix, basis, quad1, quad2, normals, k1, k2, jacobian, scaleareas,
            mesh.nodal_connectivity(+1, c1), mesh.nodal_connectivity(-1, c2))


def assemble_aligned_interface_penalty_operator(sysmatrix, basis, facequads, penalty, mesh):
    facedetjac = mesh.face_determinant_jacobian()
    for c1, f1, c2, f2 in aligned_interface_faces(mesh):
        quad1 = facequads[+1, f1, c1]
        quad2 = facequads[-1, f2, c2]
        scaleareas = np.repeat(facedetjac[f1], len(quad1))
        assemble_face_penalty_operator(
            sysmatrix, basis, quad1, quad2, penalty, scaleareas,
            mesh.nodal_connectivity(+1, c1), mesh.nodal_connectivity(-1, c2))


def assemble_aligned_robin_operator(sysmatrix, basis, facequads, lam, mesh):
    facedetjac = mesh.face_determinant_jacobian()
    for c1, f1, c2, f2 in aligned_interface_faces(mesh):
        quad1 = facequads[+1, f1, c1]
        quad2 = facequads[-1, f2, c2]
        scaleareas = np.repeat(facedetjac[f1], len(quad1))
        assemble_face_robin_operator(
            sysmatrix, basis, quad1, quad2, lam, scaleareas,
            mesh.nodal_connectivity(+1, c1), mesh.nodal_connectivity(-1, c2))


def assemble_aligned_robin_source(sysrhs, basis, facequads, lam, Tm, mesh):
    """Face-level analogue of IP_robin_interface_source.jl:27-63."""
    facedetjac = mesh.face_determinant_jacobian()
    for c1, f1, c2, f2 in aligned_interface_faces(mesh):
        quad1 = facequads[+1, f1, c1]
        quad2 = facequads[-1, f2, c2]
        scaleareas = np.repeat(facedetjac[f1], len(quad1))
        assemble_face_robin_source(sysrhs, basis, quad1, lam, Tm, scaleareas,
                                   mesh.nodal_connectivity(+1, c1))
        assemble_face_robin_source(sysrhs, basis, quad2, lam, Tm, scaleareas,
                                   mesh.nodal_connectivity(-1, c2))


# ------------------------------------------------------------------ rhs terms
def linear_form(rhsfunc, basis, quad, cellmap, detjac):
    """IP_source_term.jl:97-107"""
    rhs = np.zeros(number_of_basis_functions(basis))
    for p, w in quad:
        rhs += basis(p) * rhsfunc(cellmap(p)) * detjac * w
    return rhs


def flux_linear_form(rhsfunc, basis, quad, normal, conductivity, cellmap, jacobian,
                     facedetjac):
    """IP_source_term.jl:109-129"""
    rhs = np.zeros(number_of_basis_functions(basis))
    for p, w in quad:
        G = transform_gradient(gradient(basis, p), jacobian)
        R = rhsfunc(cellmap(p))
        rhs += conductivity * (G @ normal) * R * facedetjac * w
    return rhs


def assemble_boundary_face_source(sysrhs, rhsfunc, basis, quad, normal, conductivity,
                                  penalty, cellmap, jacobian, facedetjac, nodeids,
                                  variant="current"):
    """IP_source_term.jl:131-159 (`current`: R1 - R2; `legacy_boundary`: R1)."""
    R1 = penalty * linear_form(rhsfunc, basis, quad, cellmap, facedetjac)
    if variant == "current":
        R2 = flux_linear_form(rhsfunc, basis, quad, normal, conductivity, cellmap,
                              jacobian, facedetjac)
        rhs = R1 - R2
    elif variant == "legacy_boundary":
        rhs = R1
    else:
        raise ValueError(variant)
    assemble_cell_rhs(sysrhs, nodeids, 1, rhs)


def assemble_boundary_source(sysrhs, rhsfunc, basis, facequads, k1, k2, penalty, mesh,
                             variant="current"):
    """IP_source_term.jl:161-258"""
    jacobian = mesh.jacobian()
    facedetjac = mesh.face_determinant_jacobian()
    faceids = range(facequads.number_of_faces_per_cell())
    normals = reference_face_normals()
    for cellid in range(mesh.number_of_cells()):
        s = mesh.cell_sign(cellid)
        check_cellsign(s)
        for cellsign, k in ((+1, k1), (-1, k2)):
            if not (s == cellsign or s == 0):
                continue
            nodeids = mesh.nodal_connectivity(cellsign, cellid)
            cellmap = mesh.cell_map(cellsign, cellid)
            for faceid in faceids:
                quad = facequads[cellsign, faceid, cellid]
                if mesh.cell_connectivity(faceid, cellid) < 0:
                    assemble_boundary_face_source(
                        sysrhs, rhsfunc, basis, quad, normals[faceid], k, penalty,
                        cellmap, jacobian, facedetjac[faceid], nodeids, variant)


def assemble_cell_source(sysrhs, rhsfunc, basis, cellquads, mesh, cellsign, cellid):
    """IP_source_term.jl:2-17"""
    detjac = mesh.determinant_jacobian()
    cellmap = mesh.cell_map(cellsign, cellid)
    quad = cellquads[cellsign, cellid]
    rhs = linear_form(rhsfunc, basis, quad, cellmap, detjac)
    assemble_cell_rhs(sysrhs, mesh.nodal_connectivity(cellsign, cellid), 1, rhs)


def assemble_source(sysrhs, rhsfunc, basis, cellquads, mesh):
    """IP_source_term.jl:19-49"""
    assemble_two_phase_source(sysrhs, rhsfunc, rhsfunc, basis, cellquads, mesh)


def assemble_two_phase_source(sysrhs, rhsfunc1, rhsfunc2, basis, cellquads, mesh):
    """IP_source_term.jl:54-91"""
    for cellid in range(mesh.number_of_cells()):
        s = mesh.cell_sign(cellid)
        check_cellsign(s)
        if s == +1 or s == 0:
            assemble_cell_source(sysrhs, rhsfunc1, basis, cellquads, mesh, +1, cellid)
        if s == -1 or s == 0:
            assemble_cell_source(sysrhs, rhsfunc2, basis, cellquads, mesh, -1, cellid)


# ----------------------------------------------------------------- top level
def assemble_interior_penalty_linear_system(sysmatrix, solverbasis, cellquads, facequads,
                                            interfacequads, k1, k2, penalty, mesh,
                                            variant="current", aligned_interface=False):
    """IP_assembly.jl:1-80.  `interfacequads` is unused on uncut meshes (no cell
    has sign 0); `aligned_interface=True` adds the mesh-aligned interface terms in
    the slot where the reference assembles the cut-cell interface terms."""
    assemble_gradient_operator(sysmatrix, solverbasis, cellquads, k1, k2, mesh)
    assemble_interelement_flux_operator(sysmatrix, solverbasis, facequads, k1, k2, mesh)
    assemble_interelement_penalty_operator(sysmatrix, solverbasis, facequads, penalty, mesh)
    if aligned_interface:
        assemble_aligned_interface_flux_operator(sysmatrix, solverbasis, facequads, k1, k2, mesh)
        assemble_aligned_interface_penalty_operator(sysmatrix, solverbasis, facequads, penalty, mesh)
    assemble_boundary_flux_operator(sysmatrix, solverbasis, facequads, k1, k2, mesh, variant)
    assemble_boundary_penalty_operator(sysmatrix, solverbasis, facequads, penalty, mesh)


def assemble_interior_penalty_rhs(sysrhs, sourceterm, boundaryfunc, solverbasis, cellquads,
                                  facequads, k1, k2, penalty, mesh, variant="current"):
    """IP_assembly.jl:82-115"""
    assemble_boundary_source(sysrhs, boundaryfunc, solverbasis, facequads, k1, k2,
                             penalty, mesh, variant)
    assemble_source(sysrhs, sourceterm, solverbasis, cellquads, mesh)
