"""`InteriorPenalty.interpolate_at_reference_points` / `gradient_at_reference_points` (IP_postprocess.jl:1-52) and the
front-velocity formulas of IP_robin_interface_flux_velocity.jl:213-278 over the CUDA kernels (csrc/postprocess.cu)."""
import ctypes

import numpy as np

from . import _lib
from .cutcelldg import _ptr, _stream_ptr


def _dev(a, dtype):
    import torch
    if a is None:
        return None
    if not isinstance(a, torch.Tensor):
        a = torch.from_numpy(np.ascontiguousarray(a, dtype=dtype))
    return a.cuda().contiguous()


def interpolate_at_reference_points(nodalvalues, basis, refpoints, refcellids, stream=None):
    """refpoints [2, numpts] reference coordinates in the solution cells refcellids [numpts] (0-based); returns the
    interpolated values [numpts] (device tensor)."""
    import torch
    pts = _dev(np.ascontiguousarray(np.asarray(refpoints).T) if not isinstance(refpoints, torch.Tensor) else refpoints.t(), np.float64)
    ids = _dev(refcellids, np.int32)
    out = torch.empty(ids.numel(), dtype=torch.float64, device="cuda")
    _lib.check(_lib.lib().stefan_interpolate_at_reference_points(basis.order, _ptr(nodalvalues), _ptr(pts), _ptr(ids),
                                                                 ids.numel(), _ptr(out), _stream_ptr(stream)))
    return out


def gradient_at_reference_points(nodalvalues, basis, refpoints, refcellids, jacobian, stream=None):
    """Physical gradients [2, numpts] (device tensor); jacobian = (hx / 2, hy / 2)."""
    import torch
    pts = _dev(np.ascontiguousarray(np.asarray(refpoints).T) if not isinstance(refpoints, torch.Tensor) else refpoints.t(), np.float64)
    ids = _dev(refcellids, np.int32)
    out = torch.empty((ids.numel(), 2), dtype=torch.float64, device="cuda")
    _lib.check(_lib.lib().stefan_gradient_at_reference_points(basis.order, _ptr(nodalvalues), _ptr(pts), _ptr(ids),
                                                              ids.numel(), float(jacobian[0]), float(jacobian[1]),
                                                              _ptr(out), _stream_ptr(stream)))
    return out.t()


def front_velocities(vals1, vals2, grads1, grads2, normals, lam, Tm, k1, k2, stream=None):
    """(velocity1, velocity2, meanvelocity, gvelocity) at the interface points; grads / normals [2, numpts]."""
    import torch
    n = vals1.numel()
    g1, g2, nr = (t.t().contiguous() for t in (grads1, grads2, _dev(normals, np.float64) if not isinstance(normals, torch.Tensor) else normals))
    outs = [torch.empty(n, dtype=torch.float64, device="cuda") for _ in range(4)]
    _lib.check(_lib.lib().stefan_front_velocity(n, float(lam), float(Tm), float(k1), float(k2), _ptr(vals1), _ptr(vals2),
                                                _ptr(g1), _ptr(g2), _ptr(nr), *[_ptr(o) for o in outs], _stream_ptr(stream)))
    return tuple(outs)


def interpolate_field_at_reference_points(nodalvalues, dofspernode, first_component, ncomponents, basis, refpoints,
                                          refcellids, stream=None):
    """`ncomponents` consecutive components of a field with `dofspernode` interleaved dofs per node (the LDG solution:
    T, q_x, q_y) at the reference points: [ncomponents, numpts] (CutCellDG.interpolate_at_reference_points(G, ndofs, ...)
    as lagrange_levelset_LDG_interface_flux.jl:256-291 calls it)."""
    import torch
    pts = _dev(np.ascontiguousarray(np.asarray(refpoints).T), np.float64)
    ids = _dev(refcellids, np.int32)
    out = torch.empty((ids.numel(), ncomponents), dtype=torch.float64, device="cuda")
    _lib.check(_lib.lib().stefan_interpolate_field_at_reference_points(
        basis.order, _ptr(nodalvalues), int(dofspernode), int(first_component), int(ncomponents), _ptr(pts), _ptr(ids),
        ids.numel(), _ptr(out), _stream_ptr(stream)))
    return out.t()


def ldg_interface_flux(grads1, grads2, normals, k1, k2, V0, stream=None):
    """(normalflux1, normalflux2, numericalnormalflux) of lagrange_levelset_LDG_interface_flux.jl:266-307;
    grads / normals [2, numpts] device tensors."""
    import torch
    n = grads1.shape[1]
    g1, g2, nr = (t.t().contiguous() for t in (grads1, grads2, normals))
    outs = [torch.empty(n, dtype=torch.float64, device="cuda") for _ in range(3)]
    _lib.check(_lib.lib().stefan_ldg_interface_flux(n, float(k1), float(k2), float(V0[0]), float(V0[1]), _ptr(g1), _ptr(g2),
                                                    _ptr(nr), *[_ptr(o) for o in outs], _stream_ptr(stream)))
    return tuple(outs)


def mesh_l2_error(dim, order, numqp, ncells, detjac, dofspernode, u, uex_qp, cell_detjac=None, stream=None):
    """mesh_L2_error / integral_norm_on_mesh (StefanDG/useful_routines.jl:95-133, :203-221; 1-D:
    StefanRobinIC/useful_routines.jl:52-78) on the device for a field with `dofspernode` dofs per node: returns
    (errors [dofspernode], norms [dofspernode]) as NumPy arrays.  uex_qp [ncells][points][dofspernode]."""
    import torch
    out = torch.empty(2 * dofspernode, dtype=torch.float64, device="cuda")
    ue, cd = _dev(uex_qp, np.float64), _dev(cell_detjac, np.float64)
    _lib.check(_lib.lib().stefan_mesh_l2_error(int(dim), int(order), int(numqp), int(ncells), float(detjac or 0.0), _ptr(cd),
                                               int(dofspernode), _ptr(u), _ptr(ue), _ptr(out), _stream_ptr(stream)))
    o = out.cpu().numpy().reshape(dofspernode, 2)
    return o[:, 0].copy(), o[:, 1].copy()


def map_to_reference(mesh, points):
    """Physical points [2, numpts] -> (reference points [2, numpts], cell ids) on a `cutcelldg.UniformMesh` (the uncut
    analogue of CutCellDG.map_to_reference_on_merged_mesh)."""
    p = np.asarray(points, dtype=float)
    rel = (p - mesh.x0[:, None]) / mesh.h[:, None]
    ij = np.clip(np.floor(rel).astype(np.int64), 0, np.array([[mesh.nx - 1], [mesh.ny - 1]]))
    ref = 2.0 * (rel - ij) - 1.0
    return ref, (ij[0] * mesh.ny + ij[1]).astype(np.int32)
