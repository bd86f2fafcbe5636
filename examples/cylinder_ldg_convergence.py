"""The reference's CUT-CELL local-DG cylinder study, end to end on the GPU (local-DG element-block path).

Mirrors StefanDG/LocalDG/cylindrical_bc_tests/LDG_test_cylindrical_bc_convergence.jl: unit square cut by the circle
r = 0.5 around the origin, phase +1 (k1, q1) inside, exact solution (cylinder-analytical-solution.jl) as Dirichlet data
on the whole boundary, penalties interior = interface = negboundary = 0, posboundary = 1, V0 at 45 degrees,
numqp = required_quadrature_order(p) + 2.  What changes against the Julia script:
    CutCellDG.CutMesh / CellQuadratures / FaceQuadratures / InterfaceQuadratures / MergedMesh!
                                    ->  stefanproblem_b200.cutcell.LevelSetCutMesh
    LocalDG.assemble_LDG_linear_system!   ->  stefan_ldgblocks_assemble (element-local assembly on the device)
    LocalDG.assemble_LDG_rhs! + assemble_two_phase_source!  ->  stefan_ldgblocks_rhs
    matrix \\ rhs                     ->  stefan_ldgblocks_solve (Schur complement on T, BiCGStab on the device), or with
                                        `coo`: the triplets of stefan_ldgblocks_assemble_coo + a host sparse direct solve
                                        (the reference's own step)
    mesh_L2_error / integral_norm_on_mesh -> stefan_blocks_l2_error per component (T against the exact solution, the
                                        flux variable against its gradient, as :163-171 does)
Geometry: stefanproblem_b200.cutcell.LevelSetCutMesh, the Q2 interpolant of circle_distance_function with
height-function Gauss rules (the reference's construction).  The stored CSV is reproduced to 6 digits in all six
columns at n = 3, 9, 17, 33 and to 1e-3 at n = 5.

    python examples/cylinder_ldg_convergence.py [coo]
"""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

PRM = dict(k1=1.0, k2=2.0, q1=1.0, q2=2.0, R=0.5, b=1.5, Tw=1.0, theta=45.0,
           interiorpenalty=0.0, interfacepenalty=0.0, negboundarypenalty=0.0, posboundarypenalty=1.0)


def exact_solution(prm):
    """cylinder-analytical-solution.jl:3-39 and its radial derivative (:61-69): T(x, y), (dT/dx, dT/dy)."""
    q1, k1, q2, k2, R, b, Tw = (prm[k] for k in ("q1", "k1", "q2", "k2", "R", "b", "Tw"))
    A2 = (q2 - q1) / (2 * k2) * R ** 2
    B2 = Tw + q2 / (4 * k2) * b ** 2 - A2 * np.log(b)
    B1 = R ** 2 * (q1 / k1 - q2 / k2) / 4 + B2 + A2 * np.log(R)

    def u(x, y):
        rr = np.hypot(x, y)
        return np.where(rr < R, -q1 / (4 * k1) * rr ** 2 + B1, -q2 / (4 * k2) * rr ** 2 + A2 * np.log(np.maximum(rr, 1e-300)) + B2)

    def grad(x, y):
        rr = np.maximum(np.hypot(x, y), 1e-300)
        tr = np.where(rr < R, -q1 / (2 * k1) * rr, -q2 / (2 * k2) * rr + A2 / rr)
        return tr * x / rr, tr * y / rr
    return u, grad


def measure_error(nelmts, solverorder, prm=PRM, merge_fraction=0.2, solver="schur"):
    import torch
    from stefanproblem_b200 import blocks as B
    from stefanproblem_b200.cutcell import LevelSetCutMesh
    numqp = solverorder + 1 + 2
    V0 = (np.cos(np.radians(prm["theta"])), np.sin(np.radians(prm["theta"])))
    # the reference's geometry: Q2 interpolant of circle_distance_function (levelsetorder = 2)
    mesh = LevelSetCutMesh([0.0, 0.0], [1.0, 1.0], [nelmts, nelmts], lambda x, y: prm["R"] - np.hypot(x, y), 2, numqp,
                           prm["k1"], prm["k2"], 0.0, 0.0, merge_fraction=merge_fraction)
    data = dict(mesh.data)
    data["slot_pen"] = B.ldg_slot_penalties(data, V0, prm["interiorpenalty"], prm["interfacepenalty"],
                                            prm["negboundarypenalty"], prm["posboundarypenalty"], mesh.slot_is_interface)
    u, grad = exact_solution(prm)
    system = B.LDGBlockSystem(mesh.nsol, solverorder, mesh.nqv, mesh.nslots, mesh.nqf).assemble(data, V0)
    f = mesh.sample_source(lambda x, y: prm["q1"] + 0 * x, lambda x, y: prm["q2"] + 0 * x)
    g = mesh.sample_boundary(u)
    rhs = system.rhs(f, g)
    sol = None
    if solver == "schur":
        sol, info = system.solve(rhs, tol=1e-12, maxiter=200000)
        # (the Schur complement of the finest p = 2 mesh has a condition number of 1e6: the TRUE residual the solver
        # reports may stall a little above the recurrence's; far below the discretisation error either way)
        if not info["residual_norm"] <= 1e-9 * info["rhs_norm"]:
            print(f"  [n = {nelmts}, p = {solverorder}] BiCGStab stopped at {info}; falling back to the triplets + host solve")
            sol = None
    if sol is None:
        import scipy.sparse as sp
        import scipy.sparse.linalg as spla
        r, c, v = (t.cpu().numpy() for t in system.to_coo())
        M = sp.coo_matrix((v, (r, c)), shape=system.shape).tocsc()
        sol = torch.from_numpy(spla.spsolve(M, rhs.cpu().numpy())).cuda()
        info = {"iterations": info["iterations"] if solver == "schur" else 0, "host_solve": True}
    errs = []
    for comp, exact in ((0, u), (1, lambda x, y: grad(x, y)[0]), (2, lambda x, y: grad(x, y)[1])):
        e, den = system.l2_error(sol, mesh.sample_exact(exact), comp)
        errs.append(e / den)
    return errs, info["iterations"], mesh


def main(powers=(1, 2, 3, 4, 5), solver="schur"):
    nelmts = [2 ** p + 1 for p in powers]
    table = {}
    for order in (1, 2):
        out = [measure_error(n, order, solver=solver) for n in nelmts]
        err = np.array([e for e, _, _ in out])                      # [n][T, G1, G2]
        rate = np.diff(np.log(err), axis=0) / np.diff(np.log(1.0 / np.array(nelmts)))[:, None]
        table[order] = (err, rate)
        for n, (e, it, m), r in zip(nelmts, out, [[np.nan] * 3] + list(rate)):
            print(f"p={order} nelmts {n:3d}  solution cells {m.nsol:5d}  rel. L2 errors T {e[0]:.4e} G1 {e[1]:.4e} G2 {e[2]:.4e}"
                  f"  rates {r[0]:5.2f} {r[1]:5.2f} {r[2]:5.2f}  BiCGStab iterations {it}")
    return nelmts, table


if __name__ == "__main__":
    main(solver="coo" if "coo" in sys.argv[1:] else "schur")
