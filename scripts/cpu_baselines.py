"""CPU baselines of BASELINE.md section 4 (B1-B4), measured on the build container's host cores.
The reference itself (Julia) cannot run here; these are labelled proxies of its CPU path.
usage: python scripts/cpu_baselines.py  > profiles/r2_cpu_baselines.log"""
import ctypes
import os
import subprocess
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import fast  # noqa: E402
from oracle.cport import CIPOracle  # noqa: E402
from oracle.mesh import UniformMesh2D  # noqa: E402


def rate(fn, n, seconds=3.0):
    fn()
    reps, t0 = 0, time.perf_counter()
    while time.perf_counter() - t0 < seconds:
        fn()
        reps += 1
    return n / ((time.perf_counter() - t0) / reps)


def main():
    print(f"host: {os.cpu_count()} logical cores; OMP threads of the C port: {CIPOracle.max_threads()}")
    V0 = np.array([np.cos(np.pi / 4), np.sin(np.pi / 4)])
    # B1: SciPy CSR A @ x (1 core) of the oracle-assembled matrices
    for order, n in ((1, 700), (2, 470), (3, 350)):
        m = UniformMesh2D([0, 0], [1, 1], [n, n], (order + 1) ** 2)
        m.cellsign[:] = fast.circle_phase(m)
        t0 = time.perf_counter()
        A = fast.ip_matrix(m, order, order + 1, 1.0, 2.0, 1e3 * n, 1.0)
        ta = time.perf_counter() - t0
        x = np.random.default_rng(0).uniform(-1, 1, A.shape[0])
        r = rate(lambda: A @ x, A.shape[0])
        print(f"B1 IP  p={order} {n}^2 cells ({A.shape[0] / 1e6:.2f} M dofs, {A.nnz / A.shape[0]:.1f} nnz/row): SciPy CSR A@x, 1 core: "
              f"{r / 1e9:.3f} G dof-updates/s   (vectorised NumPy assembly {ta:.1f} s)", flush=True)
    for order, n in ((1, 300), (2, 200), (3, 130)):
        m = UniformMesh2D([0, 0], [1, 1], [n, n], (order + 1) ** 2)
        m.cellsign[:] = fast.circle_phase(m)
        t0 = time.perf_counter()
        A = fast.ldg_matrix(m, order, order + 1, 1.0, 2.0, 0.0, 0.0, 1.0, V0)
        ta = time.perf_counter() - t0
        x = np.random.default_rng(0).uniform(-1, 1, A.shape[0])
        r = rate(lambda: A @ x, A.shape[0])
        print(f"B1 LDG p={order} {n}^2 cells ({A.shape[0] / 1e6:.2f} M dofs, {A.nnz / A.shape[0]:.1f} nnz/row): SciPy CSR A@x, 1 core: "
              f"{r / 1e9:.3f} G dof-updates/s   (vectorised NumPy assembly {ta:.1f} s)", flush=True)
    # B2 + B4: the C port of the reference's path (COO append -> compress -> CSR SpMV), all threads and one thread
    for order, n in ((1, 1024), (2, 512), (3, 320)):
        phase = fast.circle_phase(UniformMesh2D([0, 0], [1, 1], [n, n], (order + 1) ** 2))
        C = CIPOracle(order, order + 1, n, n, 1.0 / n, 1.0 / n, 1.0, 2.0, 1e3 * n, 1.0, "current", 255, phase)
        x = np.random.default_rng(0).uniform(-1, 1, C.nrows)
        y = np.empty_like(x)
        rall = rate(lambda: C.spmv(x, y, 0), C.nrows)
        r1 = rate(lambda: C.spmv(x, y, 1), C.nrows)
        print(f"B2 IP  p={order} {n}^2 cells ({C.nrows / 1e6:.2f} M dofs, CSR {C.nnz * 12 / 1e9:.2f} GB): C/OpenMP CSR SpMV "
              f"{CIPOracle.max_threads()} threads: {rall / 1e9:.3f} G dof-updates/s; 1 thread: {r1 / 1e9:.3f} G", flush=True)
        print(f"B4 IP  p={order} {n}^2 cells: assembly as the reference does it (quadrature loops -> {C.coo_entries / 1e6:.0f} M COO triplets "
              f"-> compress): {C.assemble_seconds:.1f} s + {C.compress_seconds:.1f} s, 1 thread "
              f"= {C.nrows / (C.assemble_seconds + C.compress_seconds) / 1e6:.2f} M dofs/s", flush=True)
        del C
    # B3: matrix-free apply on the CPU (the kernels' own per-cell row functions compiled for the host), 1 core
    hc = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "hostcheck")
    so = os.path.join(hc, "_hostcheck.so")
    subprocess.run(["g++", "-O2", "-shared", "-fPIC", "-I/usr/local/cuda/include", "-o", so, os.path.join(hc, "hostcheck.cpp")], check=True)
    H = ctypes.CDLL(so)
    d, dp = ctypes.c_double, ctypes.POINTER(ctypes.c_double)
    for order, n in ((1, 700), (2, 470), (3, 350)):
        ph = np.ascontiguousarray(fast.circle_phase(UniformMesh2D([0, 0], [1, 1], [n, n], (order + 1) ** 2)), dtype=np.int8)
        nd = n * n * (order + 1) ** 2
        x = np.random.default_rng(0).uniform(-1, 1, nd)
        y = np.zeros_like(x)
        fn = lambda: H.hostcheck_ipdg_apply(order, n, n, order + 1, d(0.5 / n), d(0.5 / n), d(1.0), d(2.0), d(1e3 * n), d(1.0), 0, 255,  # noqa: E731
                                            ph.ctypes.data_as(ctypes.POINTER(ctypes.c_int8)), x.ctypes.data_as(dp), y.ctypes.data_as(dp))
        r = rate(fn, nd)
        print(f"B3 IP  p={order} {n}^2 cells ({nd / 1e6:.2f} M dofs): matrix-free apply with the kernels' row functions on the host, 1 core: "
              f"{r / 1e9:.3f} G dof-updates/s", flush=True)


if __name__ == "__main__":
    main()
