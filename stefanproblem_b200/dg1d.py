"""`DG1D` module of the reference (StefanRobinIC/DG1D/DG1D.jl), same function names and
argument order, over the CUDA path."""
import ctypes

import numpy as np

from . import _lib
from .cutcelldg import _ptr, _stream_ptr, reference_tables

GRADIENT, FLUX, PENALTY, ROBIN, BOUNDARY_FLUX, BOUNDARY_PENALTY = 1, 2, 4, 8, 16, 32


class DGMesh1D:
    """dgmesh_1d.jl:1-51"""

    def __init__(self, xL, xR, interfacepoint, nelmts1, nelmts2, refpoints):
        self.xL, self.xR, self.interfacepoint = float(xL), float(xR), float(interfacepoint)
        self.nelmts1, self.nelmts2 = int(nelmts1), int(nelmts2)
        self.nodesperelement = int(np.atleast_2d(refpoints).shape[1])
        self.numcells = self.nelmts1 + self.nelmts2
        self.numnodes = self.numcells * self.nodesperelement
        self.elementsize = [(self.interfacepoint - self.xL) / nelmts1, (self.xR - self.interfacepoint) / nelmts2]
        self.labels = np.array([1] * self.nelmts1 + [2] * self.nelmts2)

    def cell_bounds(self):
        a = np.linspace(self.xL, self.interfacepoint, self.nelmts1 + 1)
        b = np.linspace(self.interfacepoint, self.xR, self.nelmts2 + 1)
        lo = np.concatenate([a[:-1], b[:-1]])
        hi = np.concatenate([a[1:], b[1:]])
        return lo, hi


def element_size(mesh):
    return mesh.elementsize


def number_of_nodes(mesh):
    return mesh.numnodes


def nodal_coordinates(mesh):
    ref = np.linspace(-1.0, 1.0, mesh.nodesperelement)
    lo, hi = mesh.cell_bounds()
    return (lo[:, None] + 0.5 * (1.0 + ref[None, :]) * (hi - lo)[:, None]).reshape(1, -1)


class SystemMatrix:
    def __init__(self):
        self.spec = {}
        self.terms = 0

    def _merge(self, terms, **kw):
        self.terms |= terms
        for key, val in kw.items():
            if key in self.spec and self.spec[key] != val:
                raise ValueError(f"inconsistent {key}: {self.spec[key]} vs {val}")
            self.spec[key] = val


class SystemRHS:
    def __init__(self):
        self.spec = {}


def assemble_gradient_operator(sysmatrix, basis, quad, k1, k2, mesh):
    """assemble_gradient_operator.jl:25-48 (`quad` = number of Gauss points)"""
    sysmatrix._merge(GRADIENT, order=basis.order, numqp=int(quad), k1=float(k1), k2=float(k2))


def assemble_flux_operator(sysmatrix, basis, conductivity1, conductivity2, mesh):
    """assemble_flux_operator.jl:93-120"""
    sysmatrix._merge(FLUX, order=basis.order, k1=float(conductivity1), k2=float(conductivity2))


def assemble_penalty_operator(sysmatrix, basis, penalty, mesh):
    """assemble_penalty_operator.jl:52-70"""
    sysmatrix._merge(PENALTY, order=basis.order, penalty=float(penalty))


def assemble_interface_robin_operator(sysmatrix, basis, lam, mesh):
    """assemble_interface_robin_operator.jl:46-69"""
    sysmatrix._merge(ROBIN, order=basis.order, lam=float(lam))


def assemble_boundary_flux_operator(sysmatrix, basis, conductivity1, conductivity2, mesh):
    """assemble_flux_operator.jl:122-173"""
    sysmatrix._merge(BOUNDARY_FLUX, order=basis.order, k1=float(conductivity1), k2=float(conductivity2))


def assemble_boundary_penalty_operator(sysmatrix, basis, penalty, mesh):
    """assemble_penalty_operator.jl:72-102"""
    sysmatrix._merge(BOUNDARY_PENALTY, order=basis.order, penalty=float(penalty))


def assemble_interface_robin_source(sysrhs, basis, lam, Tm, mesh):
    """assemble_interface_robin_source.jl:19-43"""
    sysrhs.spec.update(order=basis.order, lam=float(lam), Tm=float(Tm), robin=True)


def assemble_two_phase_source(sysrhs, rhsfunc1, rhsfunc2, basis, quad, mesh):
    """assemble_source_term.jl:26-61"""
    sysrhs.spec.update(order=basis.order, numqp=int(quad), f1=rhsfunc1, f2=rhsfunc2)


def assemble_boundary_rhs(sysrhs, conductivity1, conductivity2, TL, TR, basis, penalty, mesh):
    """assemble_boundary_conditions.jl:1-44"""
    sysrhs.spec.update(order=basis.order, k1=float(conductivity1), k2=float(conductivity2), TL=float(TL),
                       TR=float(TR), penalty=float(penalty), boundary=True)


class DG1DOperator:
    def __init__(self, mesh, order, numqp, k1, k2, penalty, lam, Tm, terms):
        import torch
        if not torch.cuda.is_available():
            raise _lib.StefanError(-1, "no CUDA device: stefanproblem_b200 has no CPU fallback")
        self.mesh, self.order, self.numqp = mesh, order, numqp
        self.params = _lib.Dg1dParams(
            nelmts1=mesh.nelmts1, nelmts2=mesh.nelmts2, xL=mesh.xL, xR=mesh.xR,
            interfacepoint=mesh.interfacepoint, order=order, numqp=numqp, k1=float(k1), k2=float(k2),
            penalty=float(penalty), lam=float(lam), Tm=float(Tm), terms=int(terms))
        h = ctypes.c_void_p()
        _lib.check(_lib.lib().stefan_dg1d_create(ctypes.byref(self.params), ctypes.byref(h)))
        self._h = h
        self.ndofs = int(_lib.lib().stefan_dg1d_ndofs(h))
        self.shape = (self.ndofs, self.ndofs)

    def __del__(self):
        if getattr(self, "_h", None) and _lib is not None:  # (None during interpreter shutdown)
            _lib.lib().stefan_dg1d_destroy(self._h)
            self._h = None

    def apply(self, x, out=None, stream=None):
        import torch
        assert x.is_cuda and x.dtype == torch.float64 and x.numel() == self.ndofs
        if out is None:
            out = torch.empty_like(x)
        _lib.check(_lib.lib().stefan_dg1d_apply(self._h, _ptr(x), _ptr(out), _stream_ptr(stream)))
        return out

    def __matmul__(self, x):
        return self.apply(x)

    def to_coo(self, index_base=0):
        """(rows, cols, vals) device tensors of the assembled operator (the COO seam)."""
        import ctypes
        import torch
        n = int(_lib.lib().stefan_dg1d_coo_size(self._h))
        rows = torch.empty(n, dtype=torch.int64, device="cuda")
        cols = torch.empty(n, dtype=torch.int64, device="cuda")
        vals = torch.empty(n, dtype=torch.float64, device="cuda")
        nnz = ctypes.c_int64()
        _lib.check(_lib.lib().stefan_dg1d_assemble_coo(self._h, _ptr(rows), _ptr(cols), _ptr(vals), n, index_base,
                                                       ctypes.byref(nnz), _stream_ptr()))
        return rows, cols, vals

    def rhs(self, f_qp=None, TL=0.0, TR=0.0, with_boundary=False, stream=None):
        import torch
        out = torch.empty(self.ndofs, dtype=torch.float64, device="cuda")
        _lib.check(_lib.lib().stefan_dg1d_rhs(self._h, _ptr(f_qp), float(TL), float(TR), int(with_boundary),
                                              _ptr(out), _stream_ptr(stream)))
        return out


def sparse_operator(sysmatrix, mesh, dofspernode):
    """assembly.jl:1-5"""
    s = sysmatrix.spec
    return DG1DOperator(mesh, s["order"], s.get("numqp", s["order"] + 1), s.get("k1", 1.0), s.get("k2", 1.0),
                        s.get("penalty", 0.0), s.get("lam", 0.0), 0.0, sysmatrix.terms)


def rhs_vector(sysrhs, mesh, dofspernode):
    """assembly.jl:7-11: device vector (closures are sampled at the quadrature points here)."""
    import torch
    s = sysrhs.spec
    order = s["order"]
    numqp = s.get("numqp", order + 1)
    op = DG1DOperator(mesh, order, numqp, s.get("k1", 1.0), s.get("k2", 1.0), s.get("penalty", 0.0),
                      s.get("lam", 0.0), s.get("Tm", 0.0), ROBIN if s.get("robin") else 0)
    f_qp = None
    if "f1" in s:
        q = reference_tables(order, numqp)["qpoints"]
        lo, hi = mesh.cell_bounds()
        xq = lo[:, None] + 0.5 * (1.0 + q[None, :]) * (hi - lo)[:, None]
        f = np.where((mesh.labels == 1)[:, None], np.broadcast_to(s["f1"](xq), xq.shape),
                     np.broadcast_to(s["f2"](xq), xq.shape))
        f_qp = torch.from_numpy(np.ascontiguousarray(f, dtype=np.float64)).cuda()
    return op.rhs(f_qp, s.get("TL", 0.0), s.get("TR", 0.0), bool(s.get("boundary")))
