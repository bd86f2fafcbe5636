"""TEST INFRASTRUCTURE (oracle): point evaluation, front-velocity formulas and the interface error norm, restated.

* `gradient_at_reference_points`      StefanDG/InteriorPenalty/IP_postprocess.jl:1-28
* `interpolate_at_reference_points`   IP_postprocess.jl:30-52
* `front_velocities`                  robin_interface_velocity/IP_robin_interface_flux_velocity.jl:213-278
* `interface_L2_error`, `integral_norm_on_interface`   StefanDG/useful_routines.jl:239-274, :288-307
`nodal_connectivity(mesh, levelsetsign, cellid)` of the reference is the caller's choice of solution cell here:
`refcellids` index solution cells (NF consecutive dofs each).  Parity: literal restatement of the cited lines; the
un-vendored pieces they call (closest points, map_to_reference_on_merged_mesh) are inputs."""
import numpy as np

from .basis import LagrangeTensorProductBasis, gradient, transform_gradient


def gradient_at_reference_points(nodalvalues, basis, refpoints, refcellids, jacobian):
    """IP_postprocess.jl:1-28: refpoints [2, numpts]; returns [2, numpts]"""
    dim, numpts = refpoints.shape
    assert len(refcellids) == numpts
    nf = basis.nf
    out = np.zeros((dim, numpts))
    for idx in range(numpts):
        cellvals = nodalvalues[refcellids[idx] * nf:(refcellids[idx] + 1) * nf]
        grad = transform_gradient(gradient(basis, refpoints[:, idx]), jacobian)
        out[:, idx] = grad.T @ cellvals
    return out


def interpolate_at_reference_points(nodalvalues, basis, refpoints, refcellids):
    """IP_postprocess.jl:30-52"""
    dim, numpts = refpoints.shape
    assert len(refcellids) == numpts
    nf = basis.nf
    out = np.zeros(numpts)
    for idx in range(numpts):
        cellvals = nodalvalues[refcellids[idx] * nf:(refcellids[idx] + 1) * nf]
        out[idx] = cellvals @ basis(refpoints[:, idx])
    return out


def front_velocities(vals1, vals2, grads1, grads2, normals, lam, Tm, k1, k2):
    """IP_robin_interface_flux_velocity.jl:213-278: velocity1/2 = -lambda (Tm - vals), their mean, and
    gvelocity = normalflux2 - normalflux1 with normalflux_s = k_s sum(grads_s .* normals, dims=1)."""
    velocity1 = -lam * (Tm - vals1)
    velocity2 = -lam * (Tm - vals2)
    meanvelocity = 0.5 * (velocity1 + velocity2)
    normalflux1 = k1 * np.sum(grads1 * normals, axis=0)
    normalflux2 = k2 * np.sum(grads2 * normals, axis=0)
    return velocity1, velocity2, meanvelocity, normalflux2 - normalflux1


def interface_L2_error(order, data, slot_select, u, uex_fq):
    """useful_routines.jl:239-274 / :288-307 on block data: sqrt(sum (u_h - u_ex)^2 scalearea w) and the norm of
    u_ex over the selected slots' points (face_wts = scale area x weight)."""
    basis = LagrangeTensorProductBasis(2, order)
    nf = basis.nf
    e2 = n2 = 0.0
    ncells, nslots = data["slot_nbr"].shape
    for c in range(ncells):
        uc = u[c * nf:(c + 1) * nf]
        for s in range(nslots):
            if not slot_select[c, s]:
                continue
            for p, w, ue in zip(data["face_pts"][c, s], data["face_wts"][c, s], uex_fq[c, s]):
                if w == 0.0:
                    continue
                e2 += (basis(p) @ uc - ue) ** 2 * w
                n2 += ue ** 2 * w
    return np.sqrt(e2), np.sqrt(n2)


def interpolate_field_at_reference_points(nodalvalues, ndofs, basis, refpoints, refcellids):
    """CutCellDG.interpolate_at_reference_points(G, ndofs, basis, refpoints, refcellids, sign, mesh) as the reference's
    drivers call it for vector fields (lagrange_levelset_LDG_interface_flux.jl:256-291): nodalvalues [ndofs, nnodes]
    (the reference's `G`), returns [ndofs, numpts]."""
    dim, numpts = refpoints.shape
    nf = basis.nf
    out = np.zeros((ndofs, numpts))
    for idx in range(numpts):
        c = refcellids[idx]
        out[:, idx] = nodalvalues[:, c * nf:(c + 1) * nf] @ basis(refpoints[:, idx])
    return out


def ldg_interface_flux(grads1, grads2, normals, k1, k2, V0):
    """lagrange_levelset_LDG_interface_flux.jl:266-307: normalflux1, normalflux2 and the LDG numerical normal flux
    (beta = LocalDG.edge_switch(V0, normals), LDG_utilities.jl:1-4)."""
    from .ldg import edge_switch
    normalflux1 = k1 * np.sum(grads1 * normals, axis=0)
    normalflux2 = k2 * np.sum(grads2 * normals, axis=0)
    beta = edge_switch(np.asarray(V0, dtype=float), normals)
    betaqn = beta * (-normalflux1 + normalflux2)
    numericalflux = 0.5 * (k1 * grads1 + k2 * grads2) - betaqn
    return normalflux1, normalflux2, np.sum(numericalflux * normals, axis=0)
