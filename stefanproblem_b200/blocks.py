"""Element-block path (`stefan_blocks_*`): the reference's element-local assembly for
arbitrary per-cell / per-face quadratures, and the application of the dense blocks.

`uniform_mesh_block_data` produces the `stefan_blocks_data` arrays for an uncut uniform mesh
(tensor Gauss rules, constant normals): the same system the matrix-free kernels apply, and
the template for feeding `CellQuadratures` / `FaceQuadratures` / `InterfaceQuadratures`
exported from a CutCellDG run (one solution cell per (cell, phase); the interface of a cut
cell is an extra face slot between its two solution cells)."""
import ctypes

import numpy as np

from . import _lib
from .cutcelldg import _ptr, _stream_ptr

_FIELDS = ("vol_pts", "vol_wts", "cell_jac", "cell_k", "slot_nbr", "face_pts", "face_nbr_pts", "face_wts",
           "face_normals", "slot_pen", "slot_lambda")


class BlockSystem:
    """Assembled dense element blocks on one GPU: `assemble(data)`, `A @ x`, `to_coo()`."""

    def __init__(self, ncells, order, nqv, nslots, nqf, variant="current"):
        """variant: "current" or "legacy_boundary" (the boundary form of the reference's stored cut-cell CSVs)."""
        import torch
        if not torch.cuda.is_available():
            raise _lib.StefanError(-1, "no CUDA device: stefanproblem_b200 has no CPU fallback")
        self.dims = _lib.BlocksDims(int(ncells), int(order), int(nqv), int(nslots), int(nqf))
        h = ctypes.c_void_p()
        _lib.check(_lib.lib().stefan_blocks_create(ctypes.byref(self.dims), ctypes.byref(h)))
        self._h = h
        self.ndofs = int(_lib.lib().stefan_blocks_ndofs(h))
        self.shape = (self.ndofs, self.ndofs)
        _lib.check(_lib.lib().stefan_blocks_set_variant(h, {"current": 0, "legacy_boundary": 1}[variant]))

    def __del__(self):
        h = getattr(self, "_h", None)
        if h and _lib is not None:
            _lib.lib().stefan_blocks_destroy(h)
            self._h = None

    def assemble(self, data, stream=None):
        """data: dict of NumPy / torch arrays named like the fields of stefan_blocks_data."""
        import torch
        dev = {}
        for name in _FIELDS:
            a = data[name]
            if not isinstance(a, torch.Tensor):
                a = torch.from_numpy(np.ascontiguousarray(a, dtype=np.int32 if name == "slot_nbr" else np.float64))
            dev[name] = a.cuda().contiguous()
        d = self.dims
        expect = {"vol_pts": d.ncells * d.nqv * 2, "vol_wts": d.ncells * d.nqv, "cell_jac": d.ncells * 2,
                  "cell_k": d.ncells, "slot_nbr": d.ncells * d.nslots, "face_pts": d.ncells * d.nslots * d.nqf * 2,
                  "face_nbr_pts": d.ncells * d.nslots * d.nqf * 2, "face_wts": d.ncells * d.nslots * d.nqf,
                  "face_normals": d.ncells * d.nslots * d.nqf * 2, "slot_pen": d.ncells * d.nslots,
                  "slot_lambda": d.ncells * d.nslots}
        for name, n in expect.items():
            if dev[name].numel() != n:
                raise ValueError(f"{name}: {dev[name].numel()} values, expected {n}")
        self._data = dev  # keep the device arrays alive until the kernel has run
        bd = _lib.BlocksData(*[dev[n].data_ptr() for n in _FIELDS])
        self._bd = bd
        _lib.check(_lib.lib().stefan_blocks_assemble(self._h, ctypes.byref(bd), _stream_ptr(stream)))
        return self

    def _dev(self, a):
        import torch
        if a is None:
            return None
        if not isinstance(a, torch.Tensor):
            a = torch.from_numpy(np.ascontiguousarray(a, dtype=np.float64))
        return a.cuda().contiguous()

    def rhs(self, f_qp=None, g_qp=None, Tm=0.0, stream=None):
        """Right-hand side of the assembled data (stefan_blocks_rhs): f_qp [ncells][nqv] source samples,
        g_qp [ncells][nslots][nqf] boundary data, Tm of the Robin source.  Device tensor [ndofs]."""
        import torch
        f, g = self._dev(f_qp), self._dev(g_qp)
        out = torch.empty(self.ndofs, dtype=torch.float64, device="cuda")
        _lib.check(_lib.lib().stefan_blocks_rhs(self._h, ctypes.byref(self._bd), _ptr(f), _ptr(g), float(Tm), _ptr(out),
                                                _stream_ptr(stream)))
        return out

    def l2_error(self, u, uex_qp):
        """(||u_h - u_ex||, ||u_ex||) over the volume points of all solution cells (stefan_blocks_l2_error)."""
        import torch
        out = torch.empty(2, dtype=torch.float64, device="cuda")
        ue = self._dev(uex_qp)
        _lib.check(_lib.lib().stefan_blocks_l2_error(self._h, ctypes.byref(self._bd), _ptr(u), _ptr(ue), _ptr(out),
                                                     _stream_ptr()))
        e = out.cpu().numpy()
        return float(e[0]), float(e[1])

    def interface_l2_error(self, slot_select, u, uex_fq):
        """interface_L2_error / integral_norm_on_interface over the slots with slot_select != 0
        (stefan_blocks_interface_l2_error); uex_fq [ncells][nslots][nqf]."""
        import torch
        sel = torch.from_numpy(np.ascontiguousarray(slot_select, dtype=np.int8)).cuda()
        out = torch.empty(2, dtype=torch.float64, device="cuda")
        ue = self._dev(uex_fq)
        _lib.check(_lib.lib().stefan_blocks_interface_l2_error(self._h, ctypes.byref(self._bd), _ptr(sel), _ptr(u), _ptr(ue),
                                                               _ptr(out), _stream_ptr()))
        e = out.cpu().numpy()
        return float(e[0]), float(e[1])

    def solve(self, b, x0=None, tol=1e-12, maxiter=20000, preconditioner="block_jacobi", check_every=10):
        """`matrix \\ rhs` by preconditioned CG on the device (stefan_blocks_cg_solve); returns (x, info)."""
        import torch
        x = torch.zeros_like(b) if x0 is None else x0.clone()
        prm = _lib.CgParams(int(maxiter), float(tol), {"none": 0, "block_jacobi": 1}[preconditioner], int(check_every))
        res = _lib.CgResult()
        _lib.check(_lib.lib().stefan_blocks_cg_solve(self._h, _ptr(b), _ptr(x), ctypes.byref(prm), ctypes.byref(res),
                                                     _stream_ptr()))
        return x, {"iterations": res.iterations, "residual_norm": res.residual_norm, "rhs_norm": res.rhs_norm,
                   "converged": bool(res.converged)}

    def apply(self, x, out=None, stream=None):
        import torch
        assert x.is_cuda and x.dtype == torch.float64 and x.is_contiguous() and x.numel() == self.ndofs
        if out is None:
            out = torch.empty_like(x)
        _lib.check(_lib.lib().stefan_blocks_apply(self._h, _ptr(x), _ptr(out), _stream_ptr(stream)))
        return out

    def __matmul__(self, x):
        return self.apply(x)

    def to_coo(self, index_base=0):
        import torch
        n = int(_lib.lib().stefan_blocks_coo_size(self._h))
        rows = torch.empty(n, dtype=torch.int64, device="cuda")
        cols = torch.empty(n, dtype=torch.int64, device="cuda")
        vals = torch.empty(n, dtype=torch.float64, device="cuda")
        nnz = ctypes.c_int64()
        _lib.check(_lib.lib().stefan_blocks_assemble_coo(self._h, _ptr(rows), _ptr(cols), _ptr(vals), n, index_base,
                                                         ctypes.byref(nnz), _stream_ptr()))
        return rows, cols, vals


class LDGBlockSystem:
    """Local-DG element blocks on one GPU (`stefan_ldgblocks_*`): the cut-cell form of
    `assemble_LDG_linear_system!` / `assemble_LDG_rhs!`.  3 dofs per node, interleaved (T, q_x, q_y).
    `data["slot_pen"]` carries the penalty of every slot (see `ldg_slot_penalties`)."""

    def __init__(self, ncells, order, nqv, nslots, nqf):
        import torch
        if not torch.cuda.is_available():
            raise _lib.StefanError(-1, "no CUDA device: stefanproblem_b200 has no CPU fallback")
        self.dims = _lib.BlocksDims(int(ncells), int(order), int(nqv), int(nslots), int(nqf))
        h = ctypes.c_void_p()
        _lib.check(_lib.lib().stefan_ldgblocks_create(ctypes.byref(self.dims), ctypes.byref(h)))
        self._h = h
        self.ndofs = int(_lib.lib().stefan_ldgblocks_ndofs(h))
        self.shape = (self.ndofs, self.ndofs)
        self._norms = None

    def __del__(self):
        h = getattr(self, "_h", None)
        if h and _lib is not None:
            _lib.lib().stefan_ldgblocks_destroy(h)
            self._h = None

    _dev = BlockSystem._dev

    def assemble(self, data, V0, stream=None):
        import torch
        dev = {}
        for name in _FIELDS:
            a = data[name]
            if not isinstance(a, torch.Tensor):
                a = torch.from_numpy(np.ascontiguousarray(a, dtype=np.int32 if name == "slot_nbr" else np.float64))
            dev[name] = a.cuda().contiguous()
        self._data = dev
        self._host_data = data
        self._bd = _lib.BlocksData(*[dev[n].data_ptr() for n in _FIELDS])
        _lib.check(_lib.lib().stefan_ldgblocks_assemble(self._h, ctypes.byref(self._bd), float(V0[0]), float(V0[1]),
                                                        _stream_ptr(stream)))
        return self

    def apply(self, x, out=None, stream=None):
        import torch
        assert x.is_cuda and x.dtype == torch.float64 and x.is_contiguous() and x.numel() == self.ndofs
        if out is None:
            out = torch.empty_like(x)
        _lib.check(_lib.lib().stefan_ldgblocks_apply(self._h, _ptr(x), _ptr(out), _stream_ptr(stream)))
        return out

    def __matmul__(self, x):
        return self.apply(x)

    def rhs(self, f_qp=None, g_qp=None, stream=None):
        import torch
        f, g = self._dev(f_qp), self._dev(g_qp)
        out = torch.empty(self.ndofs, dtype=torch.float64, device="cuda")
        _lib.check(_lib.lib().stefan_ldgblocks_rhs(self._h, ctypes.byref(self._bd), _ptr(f), _ptr(g), _ptr(out),
                                                   _stream_ptr(stream)))
        return out

    def solve(self, b, tol=1e-12, maxiter=20000, check_every=10):
        """`matrix \\ rhs` on the device: Schur complement on T + BiCGStab (stefan_ldgblocks_solve)."""
        import torch
        x = torch.zeros_like(b)
        prm = _lib.CgParams(int(maxiter), float(tol), 0, int(check_every))
        res = _lib.CgResult()
        _lib.check(_lib.lib().stefan_ldgblocks_solve(self._h, _ptr(b), _ptr(x), ctypes.byref(prm), ctypes.byref(res),
                                                     _stream_ptr()))
        return x, {"iterations": res.iterations, "residual_norm": res.residual_norm, "rhs_norm": res.rhs_norm,
                   "converged": bool(res.converged)}

    def to_coo(self, index_base=0):
        import torch
        n = int(_lib.lib().stefan_ldgblocks_coo_size(self._h))
        rows = torch.empty(n, dtype=torch.int64, device="cuda")
        cols = torch.empty(n, dtype=torch.int64, device="cuda")
        vals = torch.empty(n, dtype=torch.float64, device="cuda")
        nnz = ctypes.c_int64()
        _lib.check(_lib.lib().stefan_ldgblocks_assemble_coo(self._h, _ptr(rows), _ptr(cols), _ptr(vals), n, index_base,
                                                            ctypes.byref(nnz), _stream_ptr()))
        return rows, cols, vals

    def l2_error(self, sol, uex_qp, component):
        """mesh_L2_error of one component of the solution (0: T, 1 / 2: the flux variable) against samples at the
        volume points, through the scalar block norm kernel (stefan_blocks_l2_error) on the same quadrature data."""
        if self._norms is None:
            d = self.dims
            self._norms = BlockSystem(d.ncells, d.order, d.nqv, d.nslots, d.nqf).assemble(self._data)
        return self._norms.l2_error(sol[component::3].contiguous(), uex_qp)


def ldg_slot_penalties(data, V0, interiorpenalty, interfacepenalty, negboundarypenalty, posboundarypenalty,
                       slot_is_interface=None):
    """`slot_pen` of local-DG block data: interior / interface penalty on neighbour slots; on boundary slots the
    penalty the reference selects from sign(V0 . n) (LDG_vector_flux_operator.jl:391-439, LDG_source_term.jl:203-321)."""
    nbr = np.asarray(data["slot_nbr"])
    pen = np.zeros(nbr.shape)
    pen[nbr >= 0] = interiorpenalty
    if slot_is_interface is not None:
        pen[(nbr >= 0) & np.asarray(slot_is_interface, bool)] = interfacepenalty
    v0n = np.asarray(data["face_normals"])[:, :, 0, :] @ np.asarray(V0, float)
    pen[nbr == -1] = np.where(v0n < 0, negboundarypenalty, posboundarypenalty)[nbr == -1]
    return pen


def uniform_mesh_block_data(nx, ny, hx, hy, numqp, phase, k1, k2, penalty, lam=0.0, aligned_interface=True):
    """`stefan_blocks_data` arrays (NumPy) of an uncut uniform nx x ny mesh of two phases.

    Slots: 0 bottom, 1 right, 2 top, 3 left.  Faces between unlike phases: coupled with
    flux + penalty + Robin(lam) if aligned_interface, else empty slots
    (what the reference's drivers do on uncut cells, IP_flux_operator.jl:181)."""
    gx, gw = np.polynomial.legendre.leggauss(numqp)
    ncells = nx * ny
    phase = np.asarray(phase).reshape(nx, ny)
    jx, jy = 0.5 * hx, 0.5 * hy
    X, Y = np.meshgrid(gx, gx, indexing="ij")
    vol_pts = np.broadcast_to(np.stack([X.ravel(), Y.ravel()], axis=1), (ncells, numqp * numqp, 2)).copy()
    vol_wts = np.broadcast_to(np.outer(gw, gw).ravel(), (ncells, numqp * numqp)).copy()
    cell_jac = np.broadcast_to(np.array([jx, jy]), (ncells, 2)).copy()
    cell_k = np.where(phase.ravel() == 1, k1, k2).astype(float)
    t, one = gx, np.ones(numqp)
    own = [np.stack([t, -one], 1), np.stack([one, t], 1), np.stack([t, one], 1), np.stack([-one, t], 1)]
    nbr = [np.stack([t, one], 1), np.stack([-one, t], 1), np.stack([t, -one], 1), np.stack([one, t], 1)]
    normal = [(0.0, -1.0), (1.0, 0.0), (0.0, 1.0), (-1.0, 0.0)]
    sa = [jx, jy, jx, jy]
    step = [(0, -1), (1, 0), (0, 1), (-1, 0)]
    slot_nbr = np.full((ncells, 4), -2, dtype=np.int32)
    face_pts = np.zeros((ncells, 4, numqp, 2))
    face_nbr_pts = np.zeros((ncells, 4, numqp, 2))
    face_wts = np.zeros((ncells, 4, numqp))
    face_normals = np.zeros((ncells, 4, numqp, 2))
    slot_pen = np.zeros((ncells, 4))
    slot_lambda = np.zeros((ncells, 4))
    for s in range(4):
        face_pts[:, s], face_nbr_pts[:, s] = own[s], nbr[s]
        face_wts[:, s] = gw * sa[s]
        face_normals[:, s] = normal[s]
    cid = np.arange(ncells, dtype=np.int64).reshape(nx, ny)
    pp = np.zeros((nx + 2, ny + 2), dtype=np.int8)       # 0 = outside the mesh
    pp[1:-1, 1:-1] = np.where(phase == 1, 1, 2)
    idp = np.full((nx + 2, ny + 2), -1, dtype=np.int64)
    idp[1:-1, 1:-1] = cid
    for s, (di, dj) in enumerate(step):
        pn = pp[1 + di:nx + 1 + di, 1 + dj:ny + 1 + dj]   # neighbour's phase code
        idn = idp[1 + di:nx + 1 + di, 1 + dj:ny + 1 + dj]
        same = pn == pp[1:-1, 1:-1]
        outside = pn == 0
        coupled = same | ((~outside) & bool(aligned_interface))
        nb = np.where(outside, -1, np.where(coupled, idn, -2))
        slot_nbr[:, s] = nb.ravel()
        slot_pen[:, s] = np.where(outside | coupled, penalty, 0.0).ravel()
        slot_lambda[:, s] = np.where((~outside) & (~same) & bool(aligned_interface), lam, 0.0).ravel()
    return {"vol_pts": vol_pts, "vol_wts": vol_wts, "cell_jac": cell_jac, "cell_k": cell_k, "slot_nbr": slot_nbr,
            "face_pts": face_pts, "face_nbr_pts": face_nbr_pts, "face_wts": face_wts, "face_normals": face_normals,
            "slot_pen": slot_pen, "slot_lambda": slot_lambda}
