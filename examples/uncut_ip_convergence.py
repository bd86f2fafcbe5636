"""The reference's convergence study, end to end on the GPU.

Mirrors StefanDG/InteriorPenalty/uncut_mesh_tests/test_uncut_IP_convergence.jl: same exact
solution / source term (:8-16), same `measure_error` steps (:18-101) -- assemble the
interior-penalty system and right-hand side, solve, `mesh_L2_error` -- for
nelmts = 2^k + 1.  What changes against the Julia script:
    CutCellDG / InteriorPenalty modules  ->  stefanproblem_b200.cutcelldg / interior_penalty
    matrix \\ rhs                          ->  matrix.solve(rhs)   (device CG, block Jacobi)
    mesh_L2_error(sol', ...)              ->  matrix.l2_error(sol, exact solution at the quadrature points)
The reference's stored CSV (IP_convergence.csv) corresponds to the legacy boundary form
(SURVEY.md N4), which is not symmetric; this driver runs the current, symmetric form, so
its numbers are compared with the oracle's direct solve in tests/test_ipdg_gpu.py and the
reference's own acceptance test applies: convergence rate > 1.95 (p = 1).

    python examples/uncut_ip_convergence.py [solverorder]
"""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def exact_solution(v):
    x, y = v
    return np.cos(2 * np.pi * x) * np.sin(2 * np.pi * y)


def source_term(v, k):
    x, y = v
    return 8 * np.pi ** 2 * k * np.cos(2 * np.pi * x) * np.sin(2 * np.pi * y)


def required_quadrature_order(polyorder):
    return polyorder + 1  # useful_routines.jl


def measure_error(nelmts, solverorder, sourceterm, exactsolution, k1, k2, penaltyfactor):
    import torch
    from stefanproblem_b200 import cutcelldg as CutCellDG
    from stefanproblem_b200 import interior_penalty as InteriorPenalty
    numqp = required_quadrature_order(solverorder)
    solverbasis = CutCellDG.LagrangeTensorProductBasis(2, solverorder)
    mesh = CutCellDG.UniformMesh([0.0, 0.0], [1.0, 1.0], [nelmts, nelmts], solverbasis)
    penalty = penaltyfactor / min(mesh.h)
    cellquads = CutCellDG.CellQuadratures(mesh, numqp)
    facequads = CutCellDG.FaceQuadratures(mesh, numqp)
    sysmatrix, sysrhs = CutCellDG.SystemMatrix(), CutCellDG.SystemRHS()
    InteriorPenalty.assemble_interior_penalty_linear_system(sysmatrix, solverbasis, cellquads, facequads, None,
                                                            k1, k2, penalty, mesh)
    InteriorPenalty.assemble_interior_penalty_rhs(sysrhs, sourceterm, exactsolution, solverbasis, cellquads, facequads,
                                                  k1, k2, penalty, mesh)
    matrix = CutCellDG.sparse_operator(sysmatrix, mesh, 1)
    rhs = CutCellDG.rhs_vector(sysrhs, mesh, 1)
    sol, info = matrix.solve(rhs, tol=1e-13, maxiter=50000)
    assert info["converged"], info
    x, y = CutCellDG.cell_quadrature_points(mesh, numqp)
    uex = torch.from_numpy(np.ascontiguousarray(exactsolution(np.stack([x, y])))).cuda()
    err, _ = matrix.l2_error(sol, uex)
    return err, info["iterations"]


def convergence_rate(dx, err):
    return np.diff(np.log(err)) / np.diff(np.log(dx))


def main(solverorder=1, powers=(1, 2, 3, 4, 5)):
    nelmts = [2 ** p + 1 for p in powers]
    k = 1.0
    out = [measure_error(n, solverorder, lambda v: source_term(v, k), exact_solution, k, k, 1e3) for n in nelmts]
    err = np.array([e for e, _ in out])
    dx = 1.0 / np.array(nelmts)
    rate = convergence_rate(dx, err)
    for n, (e, it), r in zip(nelmts, out, [np.nan] + list(rate)):
        print(f"nelmts {n:3d}  L2 error {e:.6e}  rate {r:5.2f}  CG iterations {it}")
    return nelmts, err, rate


if __name__ == "__main__":
    main(int(sys.argv[1]) if len(sys.argv) > 1 else 1)
