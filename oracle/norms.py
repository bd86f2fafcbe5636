"""Error norms of `/root/reference/StefanDG/useful_routines.jl` (oracle, test-only)."""
import numpy as np


def mesh_L2_error(nodalsolutions, exactsolution, basis, cellquads, mesh):
    """useful_routines.jl:80-133.  `nodalsolutions` is ndofs x nnodes."""
    nodalsolutions = np.atleast_2d(nodalsolutions)
    ndofs = nodalsolutions.shape[0]
    err = np.zeros(ndofs)
    for cellid in range(mesh.number_of_cells()):
        s = mesh.cell_sign(cellid)
        for cellsign in (+1, -1):
            if not (s == cellsign or s == 0):
                continue
            cellmap = mesh.cell_map(cellsign, cellid)
            detjac = cellmap.determinant_jacobian()
            nodeids = mesh.nodal_connectivity(cellsign, cellid)
            elementsolution = nodalsolutions[:, nodeids]
            for p, w in cellquads[cellsign, cellid]:
                numsol = elementsolution @ basis(p)
                exsol = np.atleast_1d(exactsolution(cellmap(p)))
                err += (numsol - exsol) ** 2 * detjac * w
    return np.sqrt(err)


def integral_norm_on_mesh(func, cellquads, mesh, ndofs):
    """useful_routines.jl:195-221"""
    vals = np.zeros(ndofs)
    for cellid in range(mesh.number_of_cells()):
        s = mesh.cell_sign(cellid)
        for cellsign in (+1, -1):
            if not (s == cellsign or s == 0):
                continue
            cellmap = mesh.cell_map(cellsign, cellid)
            detjac = cellmap.determinant_jacobian()
            for p, w in cellquads[cellsign, cellid]:
                v = np.atleast_1d(func(cellmap(p)))
                vals += v ** 2 * detjac * w
    return np.sqrt(vals)


def convergence_rate(dx, err):
    """useful_routines.jl (`convergence_rate`): diff(log err) ./ diff(log dx)."""
    return np.diff(np.log(err)) / np.diff(np.log(dx))
