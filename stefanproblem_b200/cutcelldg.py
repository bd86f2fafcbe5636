"""Host-side stand-ins for the `CutCellDG` / `PolynomialBasis` objects the hot path
is called with (uncut uniform meshes only), and the two consumers of the COO seam:
`sparse_operator` and `rhs_vector`.

In the reference, `SystemMatrix` is a growable COO triplet list that the
`assemble_*!` functions append to, and `sparse_operator(sysmatrix, mesh, dpn)`
turns it into a `SparseMatrixCSC`.  Here `SystemMatrix` records WHICH assemble
passes were requested (and with what coefficients); `sparse_operator` hands that
description to libstefan_b200.so and returns a matrix-free device operator whose
`@` is the CUDA apply kernel.  Nothing is computed on the host.
"""
import ctypes

import numpy as np

from . import _lib


class LagrangeTensorProductBasis:
    """`PolynomialBasis.LagrangeTensorProductBasis(dim, order)` descriptor."""

    def __init__(self, dim, order):
        if dim not in (1, 2):
            raise ValueError("dim must be 1 or 2")
        if order not in (1, 2, 3):
            raise ValueError("order must be 1, 2 or 3")
        self.dim, self.order = dim, order
        self.nf = (order + 1) ** dim


def number_of_basis_functions(basis):
    return basis.nf


def interpolation_points(basis):
    nodes = np.linspace(-1.0, 1.0, basis.order + 1)
    if basis.dim == 1:
        return nodes.reshape(1, -1)
    return np.stack([np.repeat(nodes, basis.order + 1), np.tile(nodes, basis.order + 1)])


def reference_tables(order, numqp):
    """1-D reference matrices from the library (host call, no GPU needed)."""
    n = order + 1
    M, K, C = (np.zeros((n, n)) for _ in range(3))
    dL, dR = np.zeros(n), np.zeros(n)
    qx, qw = np.zeros(numqp), np.zeros(numqp)
    p = lambda a: a.ctypes.data_as(_lib.c_double_p)  # noqa: E731
    _lib.check(_lib.lib().stefan_reference_tables(order, numqp, p(M), p(K), p(C), p(dL), p(dR),
                                                  p(qx), p(qw)))
    return dict(mass=M, stiff=K, adv=C, dleft=dL, dright=dR, qpoints=qx, qweights=qw)


class UniformMesh:
    """Uncut Cartesian DG mesh: what `DGMesh` + `CutMesh` + `MergedMesh!` give when no
    cell is cut.  Cells y-fastest; `phase` (+1 / -1 per cell) plays `cell_sign`."""

    def __init__(self, x0, widths, nelements, basis, phase=None):
        self.x0 = np.asarray(x0, dtype=float)
        self.widths = np.asarray(widths, dtype=float)
        self.nx, self.ny = int(nelements[0]), int(nelements[1])
        self.basis = basis
        self.nf = basis.nf
        self.ncells = self.nx * self.ny
        self.h = self.widths / np.array([self.nx, self.ny], dtype=float)
        if phase is None:
            phase = np.ones(self.ncells, dtype=np.int8)
        phase = np.ascontiguousarray(phase, dtype=np.int8).reshape(-1)
        if phase.size != self.ncells:
            raise ValueError("phase must have one entry per cell")
        if not np.all(np.abs(phase) == 1):
            raise ValueError("cell_sign must be +1 or -1 (cut cells are not supported)")
        self.phase = phase

    def number_of_cells(self):
        return self.ncells

    def number_of_nodes(self):
        return self.ncells * self.nf

    def element_size(self):
        return self.h

    def jacobian(self):
        return 0.5 * self.h

    def cell_centres(self):
        i, j = np.divmod(np.arange(self.ncells), self.ny)
        return (self.x0[0] + (i + 0.5) * self.h[0], self.x0[1] + (j + 0.5) * self.h[1])


def element_size(mesh):
    return mesh.element_size()


def number_of_cells(mesh):
    return mesh.number_of_cells()


class CellQuadratures:
    def __init__(self, mesh, numqp):
        self.numqp = int(numqp)


class FaceQuadratures:
    def __init__(self, mesh, numqp):
        self.numqp = int(numqp)


class InterfaceQuadratures:
    """Empty on an uncut mesh (no cell has sign 0)."""

    def __init__(self, mesh=None, numqp=0):
        self.numqp = int(numqp)


class SystemMatrix:
    """Description of the operator being assembled (see module docstring)."""

    def __init__(self):
        self.kind = None
        self.spec = {}

    def _merge(self, kind, **kw):
        if self.kind is None:
            self.kind = kind
        elif self.kind != kind:
            raise ValueError(f"SystemMatrix already holds a {self.kind} operator, not {kind}")
        for key, val in kw.items():
            if key == "terms":
                self.spec["terms"] = self.spec.get("terms", 0) | int(val)
            elif key in self.spec and self.spec[key] != val and val is not None:
                raise ValueError(f"inconsistent {key}: {self.spec[key]} vs {val}")
            elif val is not None:
                self.spec[key] = val


class SystemRHS:
    """Records the rhs passes; `rhs_vector` samples the callbacks at the quadrature
    points on the host (they are Julia / Python closures in the reference's call surface)
    and runs the device kernel."""

    def __init__(self):
        self.kind = None
        self.spec = {}

    def _merge(self, kind, **kw):
        if self.kind is None:
            self.kind = kind
        elif self.kind != kind:
            raise ValueError(f"SystemRHS already holds a {self.kind} right-hand side, not {kind}")
        self.spec.update({k: v for k, v in kw.items() if v is not None})

    def evaluate(self, mesh, dofspernode):
        import torch
        s = self.spec
        if self.kind == "ip":
            numqp = s["numqp"]
            f_qp = g_qp = None
            if "f1" in s:
                x, y = cell_quadrature_points(mesh, numqp)
                P = np.stack([x, y])
                f = np.where((mesh.phase > 0)[:, None], s["f1"](P), s["f2"](P))
                f_qp = torch.from_numpy(np.ascontiguousarray(np.broadcast_to(f, x.shape), dtype=np.float64)).cuda()
            if "g" in s:
                x, y = boundary_quadrature_points(mesh, numqp)
                g = np.broadcast_to(s["g"](np.stack([x, y])), x.shape)
                g_qp = torch.from_numpy(np.ascontiguousarray(g, dtype=np.float64)).cuda()
            terms = (32 | 64 if "g" in s else 0) | (128 if "lam" in s else 0)
            op = IPOperator(mesh, s["order"], numqp, s.get("k1", 1.0), s.get("k2", 1.0), s.get("penalty", 0.0),
                            s.get("lam", 0.0), s.get("Tm", 0.0), s.get("variant", "current"), terms or 1)
            return op.rhs(f_qp, g_qp)
        if self.kind == "ldg":
            numqp = s["numqp"]
            x, y = cell_quadrature_points(mesh, numqp)
            f = np.broadcast_to(s["f"](np.stack([x, y])), x.shape)
            f_qp = torch.from_numpy(np.ascontiguousarray(f, dtype=np.float64)).cuda()
            x, y = boundary_quadrature_points(mesh, numqp)
            g = np.broadcast_to(s["g"](np.stack([x, y])), x.shape)
            g_qp = torch.from_numpy(np.ascontiguousarray(g, dtype=np.float64)).cuda()
            op = LDGOperator(mesh, s["order"], numqp, 1.0, 1.0, 0.0, s["negboundarypenalty"],
                             s["posboundarypenalty"], s["V0"])
            return op.rhs(f_qp, g_qp)
        raise ValueError(f"unknown right-hand side kind {self.kind}")


def cell_quadrature_points(mesh, numqp):
    """Physical coordinates (x, y), each [ncells, numqp*numqp], of the cell quadrature
    points in the order the kernels expect (x-point outer, y-point inner)."""
    q = reference_tables(mesh.basis.order, numqp)["qpoints"]
    i, j = np.divmod(np.arange(mesh.ncells), mesh.ny)
    qx, qy = np.repeat(q, numqp), np.tile(q, numqp)
    x = mesh.x0[0] + mesh.h[0] * (i[:, None] + 0.5 * (1.0 + qx[None, :]))
    y = mesh.x0[1] + mesh.h[1] * (j[:, None] + 0.5 * (1.0 + qy[None, :]))
    return x, y


def boundary_quadrature_points(mesh, numqp):
    """Physical coordinates of the boundary-face quadrature points, concatenated as
    [bottom nx*numqp][right ny*numqp][top nx*numqp][left ny*numqp]."""
    q = reference_tables(mesh.basis.order, numqp)["qpoints"]
    xs = (mesh.x0[0] + mesh.h[0] * (np.arange(mesh.nx)[:, None] + 0.5 * (1.0 + q[None, :]))).ravel()
    ys = (mesh.x0[1] + mesh.h[1] * (np.arange(mesh.ny)[:, None] + 0.5 * (1.0 + q[None, :]))).ravel()
    x1, y1 = mesh.x0[0] + mesh.widths[0], mesh.x0[1] + mesh.widths[1]
    x = np.concatenate([xs, np.full_like(ys, x1), xs, np.full_like(ys, mesh.x0[0])])
    y = np.concatenate([np.full_like(xs, mesh.x0[1]), ys, np.full_like(xs, y1), ys])
    return x, y


def rhs_vector(sysrhs, mesh, dofspernode):
    """`CutCellDG.rhs_vector(sysrhs, mesh, dofspernode)`: device vector."""
    if sysrhs.kind is None:
        raise ValueError("nothing assembled")
    return sysrhs.evaluate(mesh, dofspernode)


def _ptr(t):
    return ctypes.c_void_p(t.data_ptr()) if t is not None else None


def _stream_ptr(stream=None):
    import torch
    s = stream if stream is not None else torch.cuda.current_stream()
    return ctypes.c_void_p(s.cuda_stream)


def _aligned(t, align=16):
    """The TMA copies of the apply kernels need 16-byte aligned vectors; a view that starts
    8 bytes off (e.g. an odd row of a matrix) is copied into a fresh allocation."""
    if t is not None and t.data_ptr() % align:
        return t.clone()
    return t


class IPOperator:
    """Matrix-free `sparse_operator` of an interior-penalty system on one GPU."""

    MARCH, SIMPLE = 0, 1
    STREAM = MARCH  # former name

    def __init__(self, mesh, order, numqp, k1, k2, penalty, lam=0.0, Tm=0.0,
                 variant="current", terms=None, phase_lo=None, phase_hi=None):
        import torch
        if not torch.cuda.is_available():
            raise _lib.StefanError(-1, "no CUDA device: stefanproblem_b200 has no CPU fallback")
        if terms is None:
            terms = 255
        self.mesh, self.order, self.nf = mesh, order, (order + 1) ** 2
        self.params = _lib.IpdgParams(
            nx=mesh.nx, ny=mesh.ny, hx=float(mesh.h[0]), hy=float(mesh.h[1]), order=order,
            numqp=numqp, k1=float(k1), k2=float(k2), penalty=float(penalty), lam=float(lam),
            Tm=float(Tm), variant={"current": 0, "legacy_boundary": 1}[variant], terms=int(terms))
        i8 = lambda a: None if a is None else np.ascontiguousarray(a, dtype=np.int8)  # noqa: E731
        self._plo, self._phi = i8(phase_lo), i8(phase_hi)
        p8 = lambda a: None if a is None else a.ctypes.data_as(_lib.c_int8_p)  # noqa: E731
        h = ctypes.c_void_p()
        _lib.check(_lib.lib().stefan_ipdg_create(ctypes.byref(self.params), p8(mesh.phase),
                                                 p8(self._plo), p8(self._phi), ctypes.byref(h)))
        self._h = h
        self.ndofs = int(_lib.lib().stefan_ipdg_ndofs(h))
        self.shape = (self.ndofs, self.ndofs)

    def __del__(self):
        h = getattr(self, "_h", None)
        if h and _lib is not None:  # (None during interpreter shutdown)
            _lib.lib().stefan_ipdg_destroy(h)
            self._h = None

    def apply(self, x, out=None, x_lo=None, x_hi=None, algo=STREAM, stream=None):
        import torch
        assert x.is_cuda and x.dtype == torch.float64 and x.is_contiguous() and x.numel() == self.ndofs
        hal = 8 if self.order == 2 else 16
        x, x_lo, x_hi = _aligned(x), _aligned(x_lo, hal), _aligned(x_hi, hal)
        if out is None:
            out = torch.empty_like(x)
        _lib.check(_lib.lib().stefan_ipdg_apply(self._h, _ptr(x), _ptr(out), _ptr(x_lo), _ptr(x_hi),
                                                algo, _stream_ptr(stream)))
        return out

    def apply_columns(self, x, out, col_begin, col_end, x_lo=None, x_hi=None, algo=STREAM, stream=None):
        """Rows of the cell columns [col_begin, col_end) only (see stefan_ipdg_apply_columns)."""
        _lib.check(_lib.lib().stefan_ipdg_apply_columns(self._h, _ptr(x), _ptr(out), _ptr(x_lo), _ptr(x_hi),
                                                        col_begin, col_end, algo, _stream_ptr(stream)))
        return out

    def interface_sums(self, u, u_lo=None, u_hi=None, stream=None, out=None):
        """[length, int lambda (Tm - {T}), int (k1 dnT1 - k2 dnT2), int {T}] over the interface
        faces owned by this slab (device tensor of 4 doubles)."""
        import torch
        if out is None:
            out = torch.empty(4, dtype=torch.float64, device="cuda")
        _lib.check(_lib.lib().stefan_ipdg_interface_sums(self._h, _ptr(u), _ptr(u_lo), _ptr(u_hi), _ptr(out),
                                                         _stream_ptr(stream)))
        return out

    def euler_update(self, u, y, b, dt, stream=None):
        """u <- u + dt Minv (b - y), in place."""
        _lib.check(_lib.lib().stefan_ipdg_euler_update(self._h, _ptr(u), _ptr(y), _ptr(b), float(dt),
                                                       _stream_ptr(stream)))
        return u

    def step(self, u, b, dt, out=None, u_lo=None, u_hi=None, sync=None, stream=None):
        """out = u + dt Minv (b - A u): the fused explicit step, one pass of the apply kernel
        (stefan_ipdg_step); b may be None."""
        import torch
        assert u.is_cuda and u.dtype == torch.float64 and u.is_contiguous() and u.numel() == self.ndofs
        if out is None:
            out = torch.empty_like(u)
        _lib.check(_lib.lib().stefan_ipdg_step(self._h, _ptr(u), _ptr(out), _ptr(b), float(dt), _ptr(u_lo),
                                               _ptr(u_hi), ctypes.byref(sync) if sync is not None else None,
                                               _stream_ptr(stream)))
        return out

    def rhs(self, f_qp=None, g_qp=None, out=None, stream=None):
        """b of `assemble_interior_penalty_rhs!` (+ Robin source) from device samples."""
        import torch
        if out is None:
            out = torch.empty(self.ndofs, dtype=torch.float64, device="cuda")
        _lib.check(_lib.lib().stefan_ipdg_rhs(self._h, _ptr(f_qp), _ptr(g_qp), _ptr(out), _stream_ptr(stream)))
        return out

    def solve(self, b, x0=None, tol=1e-12, maxiter=10000, preconditioner="block_jacobi", check_every=10):
        """`matrix \\ rhs` by preconditioned CG on the device; returns (x, info dict)."""
        import torch
        x = torch.zeros_like(b) if x0 is None else x0.clone()
        prm = _lib.CgParams(int(maxiter), float(tol), {"none": 0, "block_jacobi": 1}[preconditioner], int(check_every))
        res = _lib.CgResult()
        _lib.check(_lib.lib().stefan_ipdg_cg_solve(self._h, _ptr(b), _ptr(x), ctypes.byref(prm), ctypes.byref(res),
                                                   _stream_ptr()))
        return x, {"iterations": res.iterations, "residual_norm": res.residual_norm, "rhs_norm": res.rhs_norm,
                   "converged": bool(res.converged)}

    def solve_slab(self, b, halo_lo, halo_hi, exchange, allreduce, x0=None, tol=1e-12, maxiter=10000,
                   preconditioner="block_jacobi", check_every=10):
        """The same CG across column slabs (stefan_ipdg_cg_solve_slab): this operator is one rank's slab,
        `exchange(v_ptr, stream_ptr)` fills halo_lo / halo_hi with the neighbours' edge columns of the slab
        vector at v_ptr, `allreduce(dev_ptr, count, stream_ptr)` sums device doubles over the ranks in place;
        both are enqueued on the raw stream they are given.  Returns (x, info)."""
        import torch
        x = torch.zeros_like(b) if x0 is None else x0.clone()
        prm = _lib.CgParams(int(maxiter), float(tol), {"none": 0, "block_jacobi": 1}[preconditioner], int(check_every))
        res = _lib.CgResult()
        errors = []

        def guard(fn, *a):
            try:
                fn(*a)
                return 0
            except Exception as e:  # an exception must not unwind through the C frames
                errors.append(e)
                return 1

        ex = _lib.HALO_FN(lambda ctx, v, st: guard(exchange, v, st)) if exchange is not None else _lib.HALO_FN()
        ar = _lib.ALLREDUCE_FN(lambda ctx, d, n, st: guard(allreduce, d, n, st)) if allreduce is not None \
            else _lib.ALLREDUCE_FN()
        rc = _lib.lib().stefan_ipdg_cg_solve_slab(self._h, _ptr(b), _ptr(x), ctypes.byref(prm), ctypes.byref(res),
                                                  _ptr(halo_lo), _ptr(halo_hi), ex, ar, None, _stream_ptr())
        if errors:
            raise errors[0]
        _lib.check(rc)
        return x, {"iterations": res.iterations, "residual_norm": res.residual_norm, "rhs_norm": res.rhs_norm,
                   "converged": bool(res.converged)}

    def l2_error(self, u, uex_qp):
        """(||u_h - u_ex||_L2, ||u_ex||_L2) with the handle's cell quadrature (mesh_L2_error)."""
        import torch
        out = torch.empty(2, dtype=torch.float64, device="cuda")
        _lib.check(_lib.lib().stefan_ipdg_l2_error(self._h, _ptr(u), _ptr(uex_qp), _ptr(out), _stream_ptr()))
        e = out.cpu().numpy()
        return float(e[0]), float(e[1])

    def to_coo(self, index_base=0):
        """(rows, cols, vals) device tensors of the assembled operator (the COO seam)."""
        import torch
        n = int(_lib.lib().stefan_ipdg_coo_size(self._h))
        rows = torch.empty(n, dtype=torch.int64, device="cuda")
        cols = torch.empty(n, dtype=torch.int64, device="cuda")
        vals = torch.empty(n, dtype=torch.float64, device="cuda")
        nnz = ctypes.c_int64()
        _lib.check(_lib.lib().stefan_ipdg_assemble_coo(self._h, _ptr(rows), _ptr(cols), _ptr(vals), n,
                                                       index_base, ctypes.byref(nnz), _stream_ptr()))
        return rows, cols, vals

    def apply_host(self, x_host, out_host=None, nslabs=0):
        """y = A x for HOST tensors / arrays (pinned torch tensors are fastest);
        copies and kernels are pipelined inside the C call."""
        import torch
        if isinstance(x_host, np.ndarray):
            x_host = torch.from_numpy(np.ascontiguousarray(x_host, dtype=np.float64))
        assert not x_host.is_cuda and x_host.dtype == torch.float64 and x_host.numel() == self.ndofs
        if out_host is None:
            out_host = torch.empty_like(x_host)
        _lib.check(_lib.lib().stefan_ipdg_apply_host(self._h, _ptr(x_host), _ptr(out_host), nslabs))
        return out_host

    def __matmul__(self, x):
        return self.apply(x)


class LDGOperator:
    """Matrix-free `sparse_operator(sysmatrix, mesh, 3)` of a local-DG system on one GPU."""

    def __init__(self, mesh, order, numqp, k1, k2, interiorpenalty, negboundarypenalty,
                 posboundarypenalty, V0, phase_lo=None, phase_hi=None, interfacepenalty=None):
        """interfacepenalty=None: unlike-phase faces are not coupled (the reference's drivers on uncut cells);
        a number: the face-level interface functions are applied there (mesh-aligned interface)."""
        import torch
        if not torch.cuda.is_available():
            raise _lib.StefanError(-1, "no CUDA device: stefanproblem_b200 has no CPU fallback")
        self.mesh, self.order, self.nf = mesh, order, (order + 1) ** 2
        self.params = _lib.LdgParams(
            nx=mesh.nx, ny=mesh.ny, hx=float(mesh.h[0]), hy=float(mesh.h[1]), order=order, numqp=numqp,
            k1=float(k1), k2=float(k2), interiorpenalty=float(interiorpenalty),
            negboundarypenalty=float(negboundarypenalty), posboundarypenalty=float(posboundarypenalty),
            V0x=float(V0[0]), V0y=float(V0[1]),
            interfacepenalty=0.0 if interfacepenalty is None else float(interfacepenalty),
            aligned_interface=0 if interfacepenalty is None else 1)
        i8 = lambda a: None if a is None else np.ascontiguousarray(a, dtype=np.int8)  # noqa: E731
        self._plo, self._phi = i8(phase_lo), i8(phase_hi)
        p8 = lambda a: None if a is None else a.ctypes.data_as(_lib.c_int8_p)  # noqa: E731
        h = ctypes.c_void_p()
        _lib.check(_lib.lib().stefan_ldg_create(ctypes.byref(self.params), p8(mesh.phase), p8(self._plo),
                                                p8(self._phi), ctypes.byref(h)))
        self._h = h
        self.ndofs = int(_lib.lib().stefan_ldg_ndofs(h))
        self.shape = (self.ndofs, self.ndofs)

    def __del__(self):
        h = getattr(self, "_h", None)
        if h and _lib is not None:  # (None during interpreter shutdown)
            _lib.lib().stefan_ldg_destroy(h)
            self._h = None

    def apply(self, x, out=None, x_lo=None, x_hi=None, stream=None, algo=0):
        """y = A x; algo 0 = production choice, 1 = plain one-thread-per-cell kernel, 2 = march kernel."""
        import torch
        assert x.is_cuda and x.dtype == torch.float64 and x.is_contiguous() and x.numel() == self.ndofs
        hal = 8 if self.order == 2 else 16
        x, x_lo, x_hi = _aligned(x), _aligned(x_lo, hal), _aligned(x_hi, hal)
        if out is None:
            out = torch.empty_like(x)
        _lib.check(_lib.lib().stefan_ldg_apply_algo(self._h, _ptr(x), _ptr(out), _ptr(x_lo), _ptr(x_hi),
                                                    int(algo), _stream_ptr(stream)))
        return out

    def solve(self, rhs, tol=1e-12, maxiter=20000, check_every=10, native=True):
        """`matrix \\ rhs` on the device: Schur complement on T + BiCGStab (stefan_ldg_solve: the whole
        iteration is native, scalars on the device; unpreconditioned, for the sizes of the reference's
        convergence studies).  native=False runs the host-side twin of the same algorithm
        (`ldg_solve.schur_solve`, kept for the CPU tests).  Returns (solution, info)."""
        import torch
        assert rhs.is_cuda and rhs.dtype == torch.float64 and rhs.numel() == self.ndofs
        if native:
            rhs = _aligned(rhs.contiguous())
            sol = torch.empty_like(rhs)
            prm = _lib.CgParams(int(maxiter), float(tol), 0, int(check_every))
            res = _lib.CgResult()
            _lib.check(_lib.lib().stefan_ldg_solve(self._h, _ptr(rhs), _ptr(sol), ctypes.byref(prm), ctypes.byref(res),
                                                   _stream_ptr()))
            rel = res.residual_norm / res.rhs_norm if res.rhs_norm > 0 else 0.0
            return sol, {"iterations": res.iterations, "converged": bool(res.converged), "relative_residual": rel,
                         "operator_applications": 1 + 2 * (2 * res.iterations + 1)}
        from .ldg_solve import cell_mass_inverse, schur_solve
        mass1d = reference_tables(self.order, self.params.numqp)["mass"]
        minv = torch.from_numpy(cell_mass_inverse(mass1d, 0.25 * self.params.hx * self.params.hy)).cuda()
        return schur_solve(lambda x: self.apply(x), minv, rhs.contiguous(), self.nf, tol, maxiter)

    def to_coo(self, index_base=0):
        """(rows, cols, vals) device tensors of the assembled operator (the COO seam)."""
        import torch
        n = int(_lib.lib().stefan_ldg_coo_size(self._h))
        rows = torch.empty(n, dtype=torch.int64, device="cuda")
        cols = torch.empty(n, dtype=torch.int64, device="cuda")
        vals = torch.empty(n, dtype=torch.float64, device="cuda")
        nnz = ctypes.c_int64()
        _lib.check(_lib.lib().stefan_ldg_assemble_coo(self._h, _ptr(rows), _ptr(cols), _ptr(vals), n, index_base,
                                                      ctypes.byref(nnz), _stream_ptr()))
        return rows, cols, vals

    def rhs(self, f_qp=None, g_qp=None, stream=None):
        """b of `assemble_LDG_rhs!` from device samples (layout as IPOperator.rhs)."""
        import torch
        out = torch.empty(self.ndofs, dtype=torch.float64, device="cuda")
        _lib.check(_lib.lib().stefan_ldg_rhs(self._h, _ptr(f_qp), _ptr(g_qp), _ptr(out), _stream_ptr(stream)))
        return out

    def __matmul__(self, x):
        return self.apply(x)


def sparse_operator(sysmatrix, mesh, dofspernode):
    """`CutCellDG.sparse_operator(sysmatrix, mesh, dofspernode)`: device operator."""
    if sysmatrix.kind == "ip":
        if dofspernode != 1:
            raise ValueError("interior penalty has 1 dof per node")
        s = sysmatrix.spec
        return IPOperator(mesh, s["order"], s["numqp"], s["k1"], s["k2"], s.get("penalty", 0.0),
                          s.get("lam", 0.0), 0.0, s.get("variant", "current"), s["terms"])
    if sysmatrix.kind == "ldg":
        if dofspernode != 3:
            raise ValueError("local DG has 3 dofs per node")
        s = sysmatrix.spec
        if s.get("terms", 0) != 63:
            raise ValueError("the LDG kernel applies the whole of assemble_LDG_linear_system!; "
                             f"passes recorded: {s.get('terms', 0):#x}")
        return LDGOperator(mesh, s["order"], s["numqp"], s["k1"], s["k2"], s["interiorpenalty"],
                           s["negboundarypenalty"], s["posboundarypenalty"], s["V0"],
                           interfacepenalty=s.get("interfacepenalty"))
    raise ValueError(f"nothing assembled (kind={sysmatrix.kind})")
