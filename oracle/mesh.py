"""Uniform (uncut) Cartesian DG mesh + COO containers (oracle, test-only).

Restates the subset of `CutCellDG` the hot path touches, for meshes WITHOUT cut
cells (cut-cell geometry is out of scope):

* cells numbered y-fastest: cellid = i * ny + j (0-based here, 1-based in Julia)
* DG nodes numbered consecutively per cell: node = cellid * NF + a
* faces (0-based here): 0 bottom (eta=-1), 1 right (xi=+1), 2 top, 3 left;
  `reference_face_normals` = (0,-1), (1,0), (0,1), (-1,0); `opposite_face` 0<->2, 1<->3
* `jacobian` = half the cell widths, `face_determinant_jacobian` = [jx, jy, jx, jy]
* face quadrature of opposite faces is parametrised identically
* `cell_sign` in {+1, -1}: phase of the cell (0 = cut cell, never produced here)
* `SystemMatrix` / `SystemRHS`: growable COO; `assemble_couple_cell_matrix!` takes
  rows from its first node list, cols from the second, `vals` column-major;
  dof id = node * dofspernode + dof.
"""
import numpy as np
import scipy.sparse as sp

from .basis import gauss_1d, tensor_product_quadrature

REFERENCE_FACE_NORMALS = [
    np.array([0.0, -1.0]),
    np.array([1.0, 0.0]),
    np.array([0.0, 1.0]),
    np.array([-1.0, 0.0]),
]
OPPOSITE_FACE = [2, 3, 0, 1]


def opposite_face(faceid):
    return OPPOSITE_FACE[faceid]


def reference_face_normals():
    return REFERENCE_FACE_NORMALS


def face_quadrature(faceid, n):
    """Reference-cell quadrature on one face, list of (point, weight)."""
    t, w = gauss_1d(n)
    if faceid == 0:
        return [(np.array([t[q], -1.0]), w[q]) for q in range(n)]
    if faceid == 1:
        return [(np.array([1.0, t[q]]), w[q]) for q in range(n)]
    if faceid == 2:
        return [(np.array([t[q], 1.0]), w[q]) for q in range(n)]
    return [(np.array([-1.0, t[q]]), w[q]) for q in range(n)]


class CellMap:
    """`CutCellDG.CellMap(yL, yR)`: xi -> yL + (1 + xi)/2 (yR - yL)."""

    def __init__(self, yL, yR):
        self.yL = np.atleast_1d(np.asarray(yL, dtype=float))
        self.yR = np.atleast_1d(np.asarray(yR, dtype=float))

    def __call__(self, xi):
        xi = np.atleast_1d(np.asarray(xi, dtype=float))
        return self.yL + 0.5 * (1.0 + xi) * (self.yR - self.yL)

    def jacobian(self):
        return 0.5 * (self.yR - self.yL)

    def determinant_jacobian(self):
        return float(np.prod(self.jacobian()))


class UniformMesh2D:
    """Uncut `CutCellDG` mesh on [x0, x0 + widths] with nx x ny cells."""

    def __init__(self, x0, widths, ncells, nf, cellsign=None):
        self.x0 = np.asarray(x0, dtype=float)
        self.widths = np.asarray(widths, dtype=float)
        self.nx, self.ny = int(ncells[0]), int(ncells[1])
        self.nf = nf
        self.ncells = self.nx * self.ny
        self.h = self.widths / np.array([self.nx, self.ny], dtype=float)
        if cellsign is None:
            cellsign = np.ones(self.ncells, dtype=np.int8)
        self.cellsign = np.asarray(cellsign, dtype=np.int8)
        assert self.cellsign.shape == (self.ncells,)

    def number_of_cells(self):
        return self.ncells

    def number_of_nodes(self):
        return self.ncells * self.nf

    def element_size(self):
        return self.h

    def jacobian(self):
        return 0.5 * self.h

    def determinant_jacobian(self):
        return float(np.prod(0.5 * self.h))

    def face_determinant_jacobian(self):
        jx, jy = 0.5 * self.h
        return [jx, jy, jx, jy]

    def cell_sign(self, cellid):
        return int(self.cellsign[cellid])

    def cell_ij(self, cellid):
        return divmod(cellid, self.ny)

    def cell_connectivity(self, faceid, cellid):
        """Neighbour cell id across `faceid`, or -1 on the boundary."""
        i, j = self.cell_ij(cellid)
        if faceid == 0:
            return cellid - 1 if j > 0 else -1
        if faceid == 1:
            return cellid + self.ny if i < self.nx - 1 else -1
        if faceid == 2:
            return cellid + 1 if j < self.ny - 1 else -1
        return cellid - self.ny if i > 0 else -1

    def nodal_connectivity(self, cellsign, cellid):
        return np.arange(cellid * self.nf, (cellid + 1) * self.nf)

    def solution_cell_id(self, cellsign, cellid):
        return cellid  # no merged cells on an uncut mesh

    def cell_map(self, cellsign, cellid):
        i, j = self.cell_ij(cellid)
        lo = self.x0 + self.h * np.array([i, j], dtype=float)
        return CellMap(lo, lo + self.h)

    def cell_centres(self):
        i, j = np.divmod(np.arange(self.ncells), self.ny)
        cx = self.x0[0] + (i + 0.5) * self.h[0]
        cy = self.x0[1] + (j + 0.5) * self.h[1]
        return cx, cy


class CellQuadratures:
    """`cellquads[s, cellid]` on an uncut mesh: the same tensor rule everywhere."""

    def __init__(self, numqp):
        self.quad = tensor_product_quadrature(2, numqp)

    def __getitem__(self, key):
        return self.quad


class FaceQuadratures:
    """`facequads[s, faceid, cellid]` on an uncut mesh."""

    def __init__(self, numqp):
        self.quads = [face_quadrature(f, numqp) for f in range(4)]

    def __getitem__(self, key):
        _, faceid, _ = key
        return self.quads[faceid]

    def number_of_faces_per_cell(self):
        return 4


def check_cellsign(s):
    assert s in (-1, 0, 1)


# ----------------------------------------------------------------------------
# COO containers
# ----------------------------------------------------------------------------
class SystemMatrix:
    def __init__(self):
        self.rows, self.cols, self.vals = [], [], []

    def append(self, rows, cols, vals):
        assert len(rows) == len(cols) == len(vals)
        self.rows.append(np.asarray(rows, dtype=np.int64))
        self.cols.append(np.asarray(cols, dtype=np.int64))
        self.vals.append(np.asarray(vals, dtype=np.float64))

    def triplets(self):
        if not self.rows:
            return (np.zeros(0, np.int64),) * 2 + (np.zeros(0),)
        return (
            np.concatenate(self.rows),
            np.concatenate(self.cols),
            np.concatenate(self.vals),
        )


class SystemRHS:
    def __init__(self):
        self.rows, self.vals = [], []

    def append(self, rows, vals):
        assert len(rows) == len(vals)
        self.rows.append(np.asarray(rows, dtype=np.int64))
        self.vals.append(np.asarray(vals, dtype=np.float64))


def _element_dofs(nodeids, dofs, dofspernode):
    nodeids = np.asarray(nodeids, dtype=np.int64)
    dofs = np.asarray(dofs, dtype=np.int64)
    # node-major, component-minor
    return (nodeids[:, None] * dofspernode + dofs[None, :]).ravel()


def assemble_couple_cell_matrix(sysmatrix, nodeids1, nodeids2, dofspernode, vals,
                                dofs1=None, dofs2=None):
    """`CutCellDG.assemble_couple_cell_matrix!` (5- and 7-argument forms).

    rows from `nodeids1` (x `dofs1`), cols from `nodeids2` (x `dofs2`); `vals` is
    the column-major `vec` of the (nrows x ncols) element matrix.
    """
    if dofs1 is None:
        dofs1 = np.arange(dofspernode)
    if dofs2 is None:
        dofs2 = np.arange(dofspernode)
    rowdofs = _element_dofs(nodeids1, dofs1, dofspernode)
    coldofs = _element_dofs(nodeids2, dofs2, dofspernode)
    nr, nc = len(rowdofs), len(coldofs)
    vals = np.asarray(vals, dtype=np.float64).ravel(order="F")
    assert vals.size == nr * nc
    rows = np.tile(rowdofs, nc)  # repeat(rowdofs, outer = nc)
    cols = np.repeat(coldofs, nr)  # repeat(coldofs, inner = nr)
    sysmatrix.append(rows, cols, vals)


def assemble_cell_matrix(sysmatrix, nodeids, dofspernode, vals):
    assemble_couple_cell_matrix(sysmatrix, nodeids, nodeids, dofspernode, vals)


def assemble_cell_rhs(sysrhs, nodeids, dofspernode, vals, dofs=None):
    if dofs is None:
        dofs = np.arange(dofspernode)
    rows = _element_dofs(nodeids, dofs, dofspernode)
    sysrhs.append(rows, np.asarray(vals, dtype=np.float64).ravel())


def vec(M):
    """Julia `vec` (column-major flatten)."""
    return np.asarray(M).ravel(order="F")


def sparse_operator(sysmatrix, ndofs):
    """`dropzeros!(sparse(rows, cols, vals, N, N))`: duplicates summed."""
    r, c, v = sysmatrix.triplets()
    A = sp.coo_matrix((v, (r, c)), shape=(ndofs, ndofs)).tocsr()
    A.sum_duplicates()
    A.eliminate_zeros()
    return A


def rhs_vector(sysrhs, ndofs):
    b = np.zeros(ndofs)
    for r, v in zip(sysrhs.rows, sysrhs.vals):
        np.add.at(b, r, v)
    return b
