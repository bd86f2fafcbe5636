"""1-D two-phase IP-DG with a Robin interface: literal restatement (oracle, test-only).

Follows `/root/reference/StefanRobinIC/DG1D/`: dgmesh_1d.jl,
assemble_gradient_operator.jl, assemble_flux_operator.jl,
assemble_penalty_operator.jl, assemble_interface_robin_operator.jl,
assemble_interface_robin_source.jl, assemble_boundary_conditions.jl,
assemble_source_term.jl, assembly.jl; and `StefanRobinIC/analytical_solution_1d.jl`,
`StefanRobinIC/useful_routines.jl:52-76`.
"""
import numpy as np

from .basis import gradient, number_of_basis_functions
from .mesh import (
    CellMap,
    assemble_cell_matrix,
    assemble_cell_rhs,
    assemble_couple_cell_matrix,
    vec,
)


class DGMesh1D:
    """dgmesh_1d.jl:1-51 (labels 1 / 2, consecutive node ids per cell)."""

    def __init__(self, xL, xR, interfacepoint, nelmts1, nelmts2, refpoints):
        self.xL, self.xR, self.interfacepoint = xL, xR, interfacepoint
        dx1 = (interfacepoint - xL) / nelmts1
        dx2 = (xR - interfacepoint) / nelmts2
        self.elementsize = [dx1, dx2]
        self.cellmaps = _cell_maps(xL, interfacepoint, nelmts1) + _cell_maps(
            interfacepoint, xR, nelmts2)
        refpoints = np.atleast_2d(refpoints)
        self.nodesperelement = refpoints.shape[1]
        self.numcells = nelmts1 + nelmts2
        self.numnodes = self.nodesperelement * self.numcells
        self.labels = np.array([1] * nelmts1 + [2] * nelmts2)
        self.nodalcoordinates = np.concatenate(
            [[cm(r)[0] for r in refpoints[0]] for cm in self.cellmaps])

    def element_size(self):
        return self.elementsize

    def cell_map(self, cellid):
        return self.cellmaps[cellid]

    def number_of_elements(self):
        return self.numcells

    number_of_cells = number_of_elements

    def number_of_nodes(self):
        return self.numnodes

    def nodal_connectivity(self, cellid):
        n = self.nodesperelement
        return np.arange(cellid * n, (cellid + 1) * n)

    def element_label(self, cellid):
        return int(self.labels[cellid])

    def jacobian(self, cellid):
        return float(self.cellmaps[cellid].jacobian()[0])


def _cell_maps(xL, xR, ne):
    """dgmesh_1d.jl:137-146 (`range(xL, stop=xR, length=ne+1)`)."""
    xr = np.linspace(xL, xR, ne + 1)
    return [CellMap(xr[i], xr[i + 1]) for i in range(ne)]


def _k(mesh, cellid, c1, c2):
    return c1 if mesh.element_label(cellid) == 1 else c2


# ---------------------------------------------------------------- operators
def gradient_operator(basis, quad, conductivity, jacobian):
    """assemble_gradient_operator.jl:1-9"""
    nf = number_of_basis_functions(basis)
    matrix = np.zeros((nf, nf))
    for p, w in quad:
        G = gradient(basis, p) / jacobian
        matrix += conductivity * (G @ G.T) * jacobian * w
    return matrix


def assemble_gradient_operator(sysmatrix, basis, quad, k1, k2, mesh):
    """assemble_gradient_operator.jl:11-48"""
    for cellid in range(mesh.number_of_elements()):
        label = mesh.element_label(cellid)
        if label in (1, 2):
            k = k1 if label == 1 else k2
            cellmatrix = vec(gradient_operator(basis, quad, k, mesh.jacobian(cellid)))
            assemble_cell_matrix(sysmatrix, mesh.nodal_connectivity(cellid), 1, cellmatrix)


def flux_operator(basis, qp1, qp2, jacobian):
    """assemble_flux_operator.jl:1-5: V(qp1) * (N'(qp2) / j)'"""
    V = basis(qp1)
    G = gradient(basis, qp2)[:, 0] / jacobian
    return np.outer(V, G)


def assemble_edge_flux_operator(sysmatrix, basis, qp1, qp2, normal, conductivity1,
                                conductivity2, jacobian1, jacobian2, nodeids1, nodeids2):
    """assemble_flux_operator.jl:7-59"""
    M11 = normal * conductivity1 * flux_operator(basis, qp1, qp1, jacobian1)
    M12 = normal * conductivity2 * flux_operator(basis, qp1, qp2, jacobian2)
    M21 = normal * conductivity1 * flux_operator(basis, qp2, qp1, jacobian1)
    M22 = normal * conductivity2 * flux_operator(basis, qp2, qp2, jacobian2)
    vM11 = -0.5 * vec(M11 + M11.T)
    vM12 = -0.5 * vec(M12 - M21.T)
    vM21 = -0.5 * vec(-M21 + M12.T)
    vM22 = -0.5 * vec(-M22 - M22.T)
    assemble_couple_cell_matrix(sysmatrix, nodeids1, nodeids1, 1, vM11)
    assemble_couple_cell_matrix(sysmatrix, nodeids2, nodeids1, 1, vM21)
    assemble_couple_cell_matrix(sysmatrix, nodeids1, nodeids2, 1, vM12)
    assemble_couple_cell_matrix(sysmatrix, nodeids2, nodeids2, 1, vM22)


def assemble_flux_operator(sysmatrix, basis, conductivity1, conductivity2, mesh):
    """assemble_flux_operator.jl:61-120: every interior edge, incl. the interface."""
    for cellid in range(mesh.number_of_elements() - 1):
        nbr = cellid + 1
        k1 = _k(mesh, cellid, conductivity1, conductivity2)
        k2 = _k(mesh, nbr, conductivity1, conductivity2)
        assemble_edge_flux_operator(
            sysmatrix, basis, +1.0, -1.0, 1.0, k1, k2, mesh.jacobian(cellid),
            mesh.jacobian(nbr), mesh.nodal_connectivity(cellid),
            mesh.nodal_connectivity(nbr))


def assemble_boundary_flux_operator(sysmatrix, basis, conductivity1, conductivity2, mesh):
    """assemble_flux_operator.jl:122-173"""
    for cellid, qp, normal in ((0, -1.0, -1.0), (mesh.number_of_elements() - 1, 1.0, 1.0)):
        k = _k(mesh, cellid, conductivity1, conductivity2)
        M11 = normal * k * flux_operator(basis, qp, qp, mesh.jacobian(cellid))
        M = -1.0 * vec(M11 + M11.T)
        nodeids = mesh.nodal_connectivity(cellid)
        assemble_couple_cell_matrix(sysmatrix, nodeids, nodeids, 1, M)


def mass_operator(basis, qp1, qp2):
    """assemble_penalty_operator.jl:1-5"""
    return np.outer(basis(qp1), basis(qp2))


def assemble_edge_penalty_operator(sysmatrix, basis, qp1, qp2, penalty, nodeids1, nodeids2):
    """assemble_penalty_operator.jl:7-50"""
    M11 = vec(penalty * mass_operator(basis, qp1, qp1))
    M12 = -vec(penalty * mass_operator(basis, qp1, qp2))
    M21 = -vec(penalty * mass_operator(basis, qp2, qp1))
    M22 = vec(penalty * mass_operator(basis, qp2, qp2))
    assemble_couple_cell_matrix(sysmatrix, nodeids1, nodeids1, 1, M11)
    assemble_couple_cell_matrix(sysmatrix, nodeids1, nodeids2, 1, M12)
    assemble_couple_cell_matrix(sysmatrix, nodeids2, nodeids1, 1, M21)
    assemble_couple_cell_matrix(sysmatrix, nodeids2, nodeids2, 1, M22)


def assemble_penalty_operator(sysmatrix, basis, penalty, mesh):
    """assemble_penalty_operator.jl:52-70"""
    for cellid in range(mesh.number_of_elements() - 1):
        assemble_edge_penalty_operator(
            sysmatrix, basis, +1.0, -1.0, penalty, mesh.nodal_connectivity(cellid),
            mesh.nodal_connectivity(cellid + 1))


def assemble_boundary_penalty_operator(sysmatrix, basis, penalty, mesh):
    """assemble_penalty_operator.jl:72-102"""
    for cellid, qp in ((0, -1.0), (mesh.number_of_elements() - 1, 1.0)):
        M = vec(penalty * mass_operator(basis, qp, qp))
        nodeids = mesh.nodal_connectivity(cellid)
        assemble_couple_cell_matrix(sysmatrix, nodeids, nodeids, 1, M)


def assemble_robin_edge_operator(sysmatrix, basis, qp1, qp2, lam, nodeids1, nodeids2):
    """assemble_interface_robin_operator.jl:1-44"""
    M11 = 0.25 * lam * vec(mass_operator(basis, qp1, qp1))
    M12 = 0.25 * lam * vec(mass_operator(basis, qp1, qp2))
    M21 = 0.25 * lam * vec(mass_operator(basis, qp2, qp1))
    M22 = 0.25 * lam * vec(mass_operator(basis, qp2, qp2))
    assemble_couple_cell_matrix(sysmatrix, nodeids1, nodeids1, 1, M11)
    assemble_couple_cell_matrix(sysmatrix, nodeids1, nodeids2, 1, M12)
    assemble_couple_cell_matrix(sysmatrix, nodeids2, nodeids1, 1, M21)
    assemble_couple_cell_matrix(sysmatrix, nodeids2, nodeids2, 1, M22)


def assemble_interface_robin_operator(sysmatrix, basis, lam, mesh):
    """assemble_interface_robin_operator.jl:46-69"""
    for cellid in range(mesh.number_of_elements() - 1):
        if mesh.element_label(cellid) != mesh.element_label(cellid + 1):
            assemble_robin_edge_operator(
                sysmatrix, basis, 1.0, -1.0, lam, mesh.nodal_connectivity(cellid),
                mesh.nodal_connectivity(cellid + 1))


# ---------------------------------------------------------------------- rhs
def assemble_interface_robin_source(sysrhs, basis, lam, Tm, mesh):
    """assemble_interface_robin_source.jl:1-43"""
    for cellid in range(mesh.number_of_elements() - 1):
        if mesh.element_label(cellid) != mesh.element_label(cellid + 1):
            rhs1 = 0.5 * lam * Tm * basis(+1.0)
            rhs2 = 0.5 * lam * Tm * basis(-1.0)
            assemble_cell_rhs(sysrhs, mesh.nodal_connectivity(cellid), 1, rhs1)
            assemble_cell_rhs(sysrhs, mesh.nodal_connectivity(cellid + 1), 1, rhs2)


def assemble_boundary_rhs(sysrhs, conductivity1, conductivity2, TL, TR, basis, penalty, mesh):
    """assemble_boundary_conditions.jl:1-44"""
    for cellid, qp, normal, Tb in ((0, -1.0, -1.0, TL),
                                   (mesh.number_of_elements() - 1, 1.0, 1.0, TR)):
        k = _k(mesh, cellid, conductivity1, conductivity2)
        jac = mesh.jacobian(cellid)
        R1 = penalty * Tb * basis(qp)
        R2 = -Tb * k * normal * gradient(basis, qp)[:, 0] / jac
        assemble_cell_rhs(sysrhs, mesh.nodal_connectivity(cellid), 1, R1 + R2)


def linear_form(rhsfunc, basis, quad, cellmap, detjac):
    """assemble_source_term.jl:1-11"""
    rhs = np.zeros(number_of_basis_functions(basis))
    for p, w in quad:
        rhs += basis(p) * rhsfunc(cellmap(p)) * detjac * w
    return rhs


def assemble_two_phase_source(sysrhs, rhsfunc1, rhsfunc2, basis, quad, mesh):
    """assemble_source_term.jl:13-61"""
    for cellid in range(mesh.number_of_cells()):
        label = mesh.element_label(cellid)
        cellmap = mesh.cell_map(cellid)
        if label == 1:
            f = rhsfunc1
        elif label == 2:
            f = rhsfunc2
        else:
            raise ValueError(f"Expected label = {{1,2}}, got label = {label}")
        rhs = linear_form(f, basis, quad, cellmap, mesh.jacobian(cellid))
        assemble_cell_rhs(sysrhs, mesh.nodal_connectivity(cellid), 1, rhs)


# ------------------------------------------------------- analytical solution
class Analytical1D:
    """StefanRobinIC/analytical_solution_1d.jl:3-66"""

    def __init__(self, q1, k1, q2, k2, lam, R, T1, T2, Tm):
        m = np.zeros((3, 3))
        m[0, 1:3] = 1.0
        m[1, 0] = k1 + lam * R
        m[1, 1] = -k2
        m[2, 0] = -R
        m[2, 1] = R
        m[2, 2] = 1.0
        r = np.zeros(3)
        r[0] = T2 + q2 / (2 * k2)
        r[1] = R * (q1 - q2) + lam * (Tm - T1) + lam * q1 / (2 * k1) * R ** 2
        r[2] = T1 + R ** 2 / 2 * (q2 / k2 - q1 / k1)
        self.q1, self.k1, self.q2, self.k2, self.R = q1, k1, q2, k2, R
        self.B1 = T1
        self.A1, self.A2, self.B2 = np.linalg.solve(m, r)

    def __call__(self, x):
        x = float(np.atleast_1d(x)[0])
        if x < self.R:
            return -self.q1 / (2 * self.k1) * x ** 2 + self.A1 * x + self.B1
        return -self.q2 / (2 * self.k2) * x ** 2 + self.A2 * x + self.B2


def uniform_mesh_L2_error(nodalsolution, exactsolution, basis, quad, mesh):
    """StefanRobinIC/useful_routines.jl:52-66"""
    err = 0.0
    for cellid in range(mesh.number_of_cells()):
        cellmap = mesh.cell_map(cellid)
        detjac = cellmap.determinant_jacobian()
        el = nodalsolution[mesh.nodal_connectivity(cellid)]
        for p, w in quad:
            err += (el @ basis(p) - exactsolution(cellmap(p))) ** 2 * detjac * w
    return np.sqrt(err)


def integral_norm_on_uniform_mesh(func, quad, mesh):
    """StefanRobinIC/useful_routines.jl:68-76"""
    val = 0.0
    for cellid in range(mesh.number_of_cells()):
        cellmap = mesh.cell_map(cellid)
        detjac = cellmap.determinant_jacobian()
        for p, w in quad:
            val += func(cellmap(p)) ** 2 * detjac * w
    return np.sqrt(val)
