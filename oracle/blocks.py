"""TEST INFRASTRUCTURE (oracle): the element-block path restated with the reference's own
face-level primitives (oracle/ip.py: gradient_operator IP_gradient_operator.jl:1-9,
flux_operator IP_flux_operator.jl:1-33, mass_operator IP_penalty_operator.jl:1-18) for
ARBITRARY per-cell / per-face quadrature data, in the cell-centric layout of
`stefan_blocks_data` (include/stefan_b200.h).  Parity unpinned against the reference itself
for cut-cell data (its geometry lives in un-vendored packages); pinned on uniform meshes
through oracle/fast.ip_matrix.  Not part of the product."""
import numpy as np

from . import ip
from .basis import LagrangeTensorProductBasis


def assemble_blocks(order, data):
    """data: dict of NumPy arrays shaped like stefan_blocks_data.  Returns blocks
    [ncells][1 + nslots][NF][NF] (diag, then one per slot; zeros for empty slots)."""
    basis = LagrangeTensorProductBasis(2, order)
    nf = basis.nf
    ncells, nslots = data["slot_nbr"].shape
    out = np.zeros((ncells, 1 + nslots, nf, nf))
    for c in range(ncells):
        jac1 = data["cell_jac"][c]
        k1 = data["cell_k"][c]
        quad = [(p, w) for p, w in zip(data["vol_pts"][c], data["vol_wts"][c]) if w != 0.0]
        out[c, 0] += ip.gradient_operator(basis, quad, k1, jac1, jac1[0] * jac1[1])
        for s in range(nslots):
            nbr = int(data["slot_nbr"][c, s])
            if nbr < -1:
                continue
            keep = data["face_wts"][c, s] != 0.0
            w = data["face_wts"][c, s][keep]
            q1 = list(zip(data["face_pts"][c, s][keep], w))
            normals = data["face_normals"][c, s][keep].T
            ones = np.ones(len(q1))
            pen, lam = data["slot_pen"][c, s], data["slot_lambda"][c, s]
            P11 = ip.mass_operator(basis, q1, q1, ones)
            if nbr == -1:
                M = ip.flux_operator(basis, q1, q1, normals, k1, jac1, ones)
                out[c, 0] += -(M + M.T) + pen * P11
                continue
            q2 = list(zip(data["face_nbr_pts"][c, s][keep], w))
            jac2, k2 = data["cell_jac"][nbr], data["cell_k"][nbr]
            M11 = -0.5 * ip.flux_operator(basis, q1, q1, normals, k1, jac1, ones)
            M12 = -0.5 * ip.flux_operator(basis, q1, q2, normals, k1, jac1, ones)
            M21 = -0.5 * ip.flux_operator(basis, q2, q1, normals, k2, jac2, ones)
            P12 = ip.mass_operator(basis, q1, q2, ones)
            out[c, 0] += M11 + M11.T + pen * P11 + 0.25 * lam * P11
            out[c, 1 + s] += -M12 + M21.T - pen * P12 + 0.25 * lam * P12
    return out


def dense_matrix(blocks, slot_nbr):
    ncells, nb, nf, _ = blocks.shape
    A = np.zeros((ncells * nf, ncells * nf))
    for c in range(ncells):
        A[c * nf:(c + 1) * nf, c * nf:(c + 1) * nf] += blocks[c, 0]
        for s in range(nb - 1):
            n = int(slot_nbr[c, s])
            if n >= 0:
                A[c * nf:(c + 1) * nf, n * nf:(n + 1) * nf] += blocks[c, 1 + s]
    return A
