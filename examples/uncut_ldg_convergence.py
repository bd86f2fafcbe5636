"""The reference's local-DG convergence study through the COO seam.

Mirrors StefanDG/LocalDG/uncut_mesh_tests/test_uncut_LDG_convergence.jl (exact solution,
source and gradient :7-22; `measure_error` :25-120): assemble the LDG system and right-hand
side, `sparse_operator`, `matrix \\ rhs`, `mesh_L2_error` of T (relative) and of the flux
components.  Here the operator and the right-hand side are produced on the GPU, the
operator leaves through the COO seam (`to_coo`, the triplets CutCellDG.SystemMatrix holds)
and is solved on the host with a sparse direct solve exactly as the reference does -- the
LDG system is not symmetric, so the device CG does not apply.  With the reference's
parameters (penalties (0, 0, 0, 1), V0 at 45 degrees, numqp = p+1 for p = 1 and p+3 for
p = 2) the printed numbers are the rows of its stored MD-LDG_convergence.csv.

`--device-solve` replaces the host solve by `op.solve(rhs)`: Schur complement on T + BiCGStab
over the device operator (stefanproblem_b200/ldg_solve.py), nothing assembled at all; the
errors then agree with the CSV to the iteration's tolerance instead of to round-off.

    python examples/uncut_ldg_convergence.py [--device-solve]
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


def exact_gradient(v):
    x, y = v
    return (-2 * np.pi * np.sin(2 * np.pi * x) * np.sin(2 * np.pi * y),
            +2 * np.pi * np.cos(2 * np.pi * x) * np.cos(2 * np.pi * y))


def lagrange_values(order, pts):
    """N_a(pts) of the equispaced Lagrange basis on [-1, 1]: [len(pts), order+1]."""
    nodes = np.linspace(-1.0, 1.0, order + 1)
    V = np.ones((len(pts), order + 1))
    for a in range(order + 1):
        for b in range(order + 1):
            if b != a:
                V[:, a] *= (pts - nodes[b]) / (nodes[a] - nodes[b])
    return V


def mesh_L2_error(field, exact_qp, order, numqp, nelmts, h):
    """sqrt(sum_cells sum_q (u_h(q) - u_ex(q))^2 detjac w) (useful_routines.jl:95-133);
    field [ncells, NF] nodal values, exact_qp [ncells, numqp^2]."""
    q, w = np.polynomial.legendre.leggauss(numqp)
    N = lagrange_values(order, q)                      # [nq, n1]
    V = np.einsum("qa,rb->qrab", N, N).reshape(numqp * numqp, -1)   # x-point outer, eta node fastest
    W = np.outer(w, w).ravel() * (0.25 * h * h)
    uh = field @ V.T
    return np.sqrt(np.sum((uh - exact_qp) ** 2 * W[None, :]))


def measure_error(nelmts, solverorder, numqp, k1, k2, interiorpenalty, interfacepenalty, negboundarypenalty,
                  posboundarypenalty, V0, device_solve=False):
    import scipy.sparse as sp
    import scipy.sparse.linalg as spla
    from stefanproblem_b200 import cutcelldg as CutCellDG
    from stefanproblem_b200 import local_dg as LocalDG
    basis = CutCellDG.LagrangeTensorProductBasis(2, solverorder)
    mesh = CutCellDG.UniformMesh([0.0, 0.0], [1.0, 1.0], [nelmts, nelmts], basis)
    cellquads, facequads = CutCellDG.CellQuadratures(mesh, numqp), CutCellDG.FaceQuadratures(mesh, numqp)
    sysmatrix, sysrhs = CutCellDG.SystemMatrix(), CutCellDG.SystemRHS()
    LocalDG.assemble_LDG_linear_system(sysmatrix, basis, cellquads, facequads, None, k1, k2, interiorpenalty,
                                       interfacepenalty, negboundarypenalty, posboundarypenalty, V0, mesh)
    LocalDG.assemble_LDG_rhs(sysrhs, lambda v: source_term(v, k1), exact_solution, basis, cellquads, facequads,
                             negboundarypenalty, posboundarypenalty, V0, mesh)
    op = LocalDG.sparse_operator(sysmatrix, mesh, 3)            # device operator
    rhs = LocalDG.rhs_vector(sysrhs, mesh, 3)
    if device_solve:
        sol, info = op.solve(rhs, tol=1e-13)
        assert info["converged"], info
        sol = sol.cpu().numpy().reshape(-1, 3)
    else:
        rows, cols, vals = (t.cpu().numpy() for t in op.to_coo(index_base=1))      # the COO seam, Julia indices
        matrix = sp.coo_matrix((vals, (rows - 1, cols - 1)), shape=(op.ndofs, op.ndofs)).tocsc()
        matrix.eliminate_zeros()                                # dropzeros!(sparse(rows, cols, vals, N, N))
        sol = spla.spsolve(matrix, rhs.cpu().numpy()).reshape(-1, 3)               # [node, (T, qx, qy)]
    nf = (solverorder + 1) ** 2
    T, Gx, Gy = (sol[:, c].reshape(-1, nf) for c in range(3))
    x, y = CutCellDG.cell_quadrature_points(mesh, numqp)
    h = 1.0 / nelmts
    ex, (gx, gy) = exact_solution((x, y)), exact_gradient((x, y))
    zero = np.zeros_like(T)
    Tnorm = mesh_L2_error(zero, ex, solverorder, numqp, nelmts, h)
    return (mesh_L2_error(T, ex, solverorder, numqp, nelmts, h) / Tnorm,
            mesh_L2_error(Gx, gx, solverorder, numqp, nelmts, h), mesh_L2_error(Gy, gy, solverorder, numqp, nelmts, h))


def main(powers=(1, 2, 3, 4, 5), device_solve=False):
    V0 = (np.cos(np.radians(45.0)), np.sin(np.radians(45.0)))
    table = []
    for n in [2 ** p + 1 for p in powers]:
        row = [n]
        for order, numqp in ((1, 2), (2, 5)):
            row += list(measure_error(n, order, numqp, 1.0, 1.0, 0.0, 0.0, 0.0, 1.0, V0, device_solve))
        table.append(row)
        print("  ".join([f"{row[0]:3d}"] + [f"{v:.12e}" for v in row[1:]]))
    return table


if __name__ == "__main__":
    main(device_solve="--device-solve" in sys.argv)
