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


def assemble_blocks(order, data, variant="current"):
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
                # (`legacy_boundary`: -M' only, oracle/ip.assemble_boundary_face_flux_operator)
                out[c, 0] += (-(M + M.T) if variant == "current" else -M.T) + pen * P11
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


def assemble_rhs(order, data, f_qp, g_qp, Tm=0.0, variant="current"):
    """Right-hand side for block data, from the reference's linear forms applied per solution cell:
    volume source `linear_form` (IP_source_term.jl:97-107), boundary slots `R1 - R2` =
    pen * linear_form - flux_linear_form (IP_source_term.jl:109-159), slots with a Robin coefficient
    1/2 lambda Tm * robin_interface_linear_form (IP_robin_interface_source.jl:1-25).
    f_qp [ncells][nqv]: source at the volume points; g_qp [ncells][nslots][nqf]: boundary data."""
    basis = LagrangeTensorProductBasis(2, order)
    nf = basis.nf
    ncells, nslots = data["slot_nbr"].shape
    ident = lambda p: p  # noqa: E731
    rhs = np.zeros((ncells, nf))
    for c in range(ncells):
        jac = data["cell_jac"][c]
        k = data["cell_k"][c]
        for p, w, fv in zip(data["vol_pts"][c], data["vol_wts"][c], f_qp[c]):
            if w != 0.0:   # (one point at a time: the sampled value belongs to the point, padding has weight 0)
                rhs[c] += ip.linear_form(lambda _p, fv=fv: fv, basis, [(p, w)], ident, jac[0] * jac[1])
        for s in range(nslots):
            nbr = int(data["slot_nbr"][c, s])
            keep = data["face_wts"][c, s] != 0.0
            pts, w = data["face_pts"][c, s][keep], data["face_wts"][c, s][keep]
            if nbr == -1:
                pen = data["slot_pen"][c, s]
                for p, wq, n, gq in zip(pts, w, data["face_normals"][c, s][keep], g_qp[c, s][keep]):
                    q1 = [(p, wq)]
                    g = lambda _p, gq=gq: gq  # noqa: E731
                    rhs[c] += pen * ip.linear_form(g, basis, q1, ident, 1.0)
                    if variant == "current":   # (`legacy_boundary`: R1 only, oracle/ip.assemble_boundary_face_source)
                        rhs[c] -= ip.flux_linear_form(g, basis, q1, n, k, ident, jac, 1.0)
            elif nbr >= 0 and data["slot_lambda"][c, s] != 0.0:
                quad = list(zip(pts, w))
                rhs[c] += 0.5 * data["slot_lambda"][c, s] * Tm * ip.robin_interface_linear_form(
                    basis, quad, np.ones(len(quad)))
    return rhs.ravel()


def l2_error(order, data, u, uex_qp):
    """(||u_h - u_ex||, ||u_ex||) over the volume points of all solution cells: mesh_L2_error and
    integral_norm_on_mesh (useful_routines.jl:95-133, :203-221) for per-cell quadratures."""
    basis = LagrangeTensorProductBasis(2, order)
    nf = basis.nf
    e2 = n2 = 0.0
    for c in range(data["slot_nbr"].shape[0]):
        jac = data["cell_jac"][c]
        uc = u[c * nf:(c + 1) * nf]
        for p, w, ue in zip(data["vol_pts"][c], data["vol_wts"][c], uex_qp[c]):
            if w == 0.0:
                continue
            e2 += (basis(p) @ uc - ue) ** 2 * jac[0] * jac[1] * w
            n2 += ue ** 2 * jac[0] * jac[1] * w
    return np.sqrt(e2), np.sqrt(n2)
