"""Time loops (oracle, test-only).  The reference NEVER time-steps: it
assembles a steady operator and solves once.  The explicit loops below are therefore
DEFINITIONS shared by the oracle and the kernels, not restatements; parity for them
is unpinned by construction.

* FD:  u <- u + dt * kappa(phase) * (K u) on the rows that carry a Laplace stencil
       (central / backward / forward, finite-difference.jl:34-53); all other rows
       (Dirichlet, empty phase-2 interface rows, continuity rows) are held.
* DG:  u <- u + dt * Minv (b - A u), M the block-diagonal DG mass matrix.
"""
import numpy as np

from . import fd


def fd_step(phaseid, stepsize, u, dt, kappa1, kappa2):
    K = fd.laplace_matrix(phaseid, stepsize, dirichlet=False)
    lap = np.diff(K.indptr) > 0
    kappa = np.where(np.asarray(phaseid) == 1, kappa1, kappa2)
    return np.where(lap, u + dt * kappa * (K @ u), u)


def fd_run(phaseid, stepsize, u0, dt, kappa1, kappa2, nsteps):
    u = np.array(u0, dtype=float)
    for _ in range(nsteps):
        u = fd_step(phaseid, stepsize, u, dt, kappa1, kappa2)
    return u
