"""ctypes wrapper of oracle/csrc/ip_oracle.c (oracle / CPU baseline, test-only).

`CIPOracle` assembles the interior-penalty matrix exactly as the reference does
(per-cell quadrature loops -> COO append -> compress) and applies it as a CSR
SpMV; `bench.py` times it as the CPU baseline ("port": the reference itself is
Julia and cannot run here)."""
import ctypes
import os
import subprocess

import numpy as np
import scipy.sparse as sp

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = os.path.join(_HERE, "_build", "libip_oracle.so")


def build():
    subprocess.run(["make", "-s", "-C", os.path.join(_HERE, "csrc")], check=True)
    return _LIB


_lib = None


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(_LIB):
            build()
        L = ctypes.CDLL(_LIB)
        L.ip_oracle_assemble.restype = ctypes.c_void_p
        L.ip_oracle_assemble.argtypes = [ctypes.c_int] * 4 + [ctypes.c_double] * 6 + [
            ctypes.c_int, ctypes.c_int, ctypes.c_void_p]
        L.ip_oracle_info.argtypes = [ctypes.c_void_p] + [ctypes.c_void_p] * 5
        L.ip_oracle_csr.argtypes = [ctypes.c_void_p] * 4
        L.ip_oracle_spmv.argtypes = [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int]
        L.ip_oracle_free.argtypes = [ctypes.c_void_p]
        L.ip_oracle_max_threads.restype = ctypes.c_int
        _lib = L
    return _lib


class CIPOracle:
    def __init__(self, order, numqp, nx, ny, hx, hy, k1, k2, penalty, lam=0.0,
                 variant="current", terms=255, phase=None):
        ph = None
        if phase is not None:
            ph = np.ascontiguousarray(phase, dtype=np.int8)
            assert ph.size == nx * ny
        self._ph = ph
        self._h = lib().ip_oracle_assemble(
            order, numqp, nx, ny, hx, hy, k1, k2, penalty, lam,
            {"current": 0, "legacy_boundary": 1}[variant], terms,
            None if ph is None else ph.ctypes.data)
        if not self._h:
            raise MemoryError("ip_oracle_assemble failed")
        nrows, nnz, coo = ctypes.c_int64(), ctypes.c_int64(), ctypes.c_int64()
        ta, tc = ctypes.c_double(), ctypes.c_double()
        lib().ip_oracle_info(self._h, ctypes.byref(nrows), ctypes.byref(nnz), ctypes.byref(coo),
                             ctypes.byref(ta), ctypes.byref(tc))
        self.nrows, self.nnz, self.coo_entries = nrows.value, nnz.value, coo.value
        self.assemble_seconds, self.compress_seconds = ta.value, tc.value

    def csr(self):
        indptr = np.zeros(self.nrows + 1, dtype=np.int64)
        indices = np.zeros(self.nnz, dtype=np.int32)
        data = np.zeros(self.nnz)
        lib().ip_oracle_csr(self._h, indptr.ctypes.data, indices.ctypes.data, data.ctypes.data)
        return sp.csr_matrix((data, indices, indptr), shape=(self.nrows, self.nrows))

    def spmv(self, x, y=None, nthreads=0):
        x = np.ascontiguousarray(x, dtype=np.float64)
        if y is None:
            y = np.empty_like(x)
        lib().ip_oracle_spmv(self._h, x.ctypes.data, y.ctypes.data, nthreads)
        return y

    @staticmethod
    def max_threads():
        return lib().ip_oracle_max_threads()

    def __del__(self):
        if getattr(self, "_h", None):
            lib().ip_oracle_free(self._h)
            self._h = None
