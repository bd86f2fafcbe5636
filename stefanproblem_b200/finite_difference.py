"""`FiniteDifference` module of the reference (StefanFD/finite-difference.jl), same
function names and argument order, over the CUDA path.

`SystemMatrix` records the passes (`assemble_laplace_operator!`,
`assemble_dirichlet_boundary_condition!`, `assemble_interface_continuity!`);
`sparse_operator(matrix, ndofs)` returns a device operator whose `@` is the CUDA
stencil kernel, with `.to_coo()` for the triplets the reference would have appended.
Node ids are 1-based here, as in the reference's call surface."""
import ctypes

import numpy as np

from . import _lib
from .cutcelldg import _ptr, _stream_ptr


class SystemMatrix:
    """finite-difference.jl:4-19 (description of the rows instead of triplets)."""

    def __init__(self):
        self.phaseid = None
        self.stepsize = None
        self.dirichlet_nodes = None
        self.continuity = []


def phase_id(phi):
    """finite-difference.jl:92-103"""
    phi = np.asarray(phi)
    return np.where(phi < 0, 1, 2).astype(np.int8)


def assemble_laplace_operator(matrix, phaseid, stepsize):
    """finite-difference.jl:71-85"""
    matrix.phaseid = np.ascontiguousarray(phaseid, dtype=np.int8)
    matrix.stepsize = float(stepsize)


def assemble_dirichlet_boundary_condition(matrix, numnodes):
    """finite-difference.jl:87-90"""
    matrix.dirichlet_nodes = int(numnodes)


def assemble_interface_continuity(matrix, r, s, nodeid):
    """finite-difference.jl:111-116 (nodeid 1-based)"""
    matrix.continuity.append((float(r), float(s), int(nodeid)))


class FDOperator:
    def __init__(self, phaseid, stepsize, dirichlet=True, continuity=(), phase_lo=None, phase_hi=None):
        import torch
        if not torch.cuda.is_available():
            raise _lib.StefanError(-1, "no CUDA device: stefanproblem_b200 has no CPU fallback")
        self.phaseid = np.ascontiguousarray(phaseid, dtype=np.int8)
        self.n = int(self.phaseid.size)
        i8 = lambda a: None if a is None else np.ascontiguousarray(a, dtype=np.int8)  # noqa: E731
        self._lo, self._hi = i8(phase_lo), i8(phase_hi)
        p8 = lambda a: None if a is None else a.ctypes.data_as(_lib.c_int8_p)  # noqa: E731
        h = ctypes.c_void_p()
        _lib.check(_lib.lib().stefan_fd_create(self.n, float(stepsize), p8(self.phaseid), int(bool(dirichlet)),
                                               p8(self._lo), p8(self._hi), ctypes.byref(h)))
        self._h = h
        for r, s, nodeid in continuity:
            _lib.check(_lib.lib().stefan_fd_add_interface_continuity(h, r, s, nodeid - 1))
        self.shape = (self.n, self.n)

    def __del__(self):
        if getattr(self, "_h", None) and _lib is not None:  # (None during interpreter shutdown)
            _lib.lib().stefan_fd_destroy(self._h)
            self._h = None

    def apply(self, u, out=None, u_lo=None, u_hi=None, stream=None):
        import torch
        assert u.is_cuda and u.dtype == torch.float64 and u.numel() == self.n
        if out is None:
            out = torch.empty_like(u)
        _lib.check(_lib.lib().stefan_fd_apply(self._h, _ptr(u), _ptr(out), _ptr(u_lo), _ptr(u_hi),
                                              _stream_ptr(stream)))
        return out

    def step(self, u, dt, kappa1, kappa2, out=None, u_lo=None, u_hi=None, stream=None):
        import torch
        if out is None:
            out = torch.empty_like(u)
        _lib.check(_lib.lib().stefan_fd_step(self._h, _ptr(u), _ptr(out), _ptr(u_lo), _ptr(u_hi), float(dt),
                                             float(kappa1), float(kappa2), _stream_ptr(stream)))
        return out

    def __matmul__(self, u):
        return self.apply(u)

    def to_coo(self):
        """(rows, cols, vals) device tensors, 0-based, ordered by row."""
        import torch
        nnz = int(_lib.lib().stefan_fd_nnz(self._h))
        rows = torch.empty(nnz, dtype=torch.int64, device="cuda")
        cols = torch.empty(nnz, dtype=torch.int64, device="cuda")
        vals = torch.empty(nnz, dtype=torch.float64, device="cuda")
        out = ctypes.c_int64()
        _lib.check(_lib.lib().stefan_fd_assemble_coo(self._h, _ptr(rows), _ptr(cols), _ptr(vals), nnz,
                                                     ctypes.byref(out), _stream_ptr()))
        return rows, cols, vals


def sparse_operator(sysmatrix, ndofs):
    """finite-difference.jl:105-109"""
    if sysmatrix.phaseid is None:
        raise ValueError("assemble_laplace_operator! was not called")
    if sysmatrix.phaseid.size != ndofs:
        raise ValueError("ndofs does not match the length of phaseid")
    if sysmatrix.dirichlet_nodes not in (None, ndofs):
        raise ValueError("assemble_dirichlet_boundary_condition! was called with a different numnodes")
    return FDOperator(sysmatrix.phaseid, sysmatrix.stepsize, sysmatrix.dirichlet_nodes is not None,
                      sysmatrix.continuity)
