"""Interface flux of the cut-cell local-DG solution, end to end on the GPU.

Mirrors StefanDG/LocalDG/cylindrical-bc-flux/lagrange_levelset_LDG_interface_flux.jl: unit square, circle of radius 0.3
around (0.5, 0.5) (a CLOSED interface inside the domain), phase +1 (k1, q1) inside, Dirichlet data from
cylinder-analytical-solution.jl, interior penalty 0, interface penalty 10, boundary penalties (0, 1), V0 at 45 degrees,
solverorder 2, numqp = required_quadrature_order + 2, levelsetorder 2, MergedMesh!(...; tinyratio = 0.3).  The script
solves the LDG system and evaluates, at points of the zero level set,
    normalflux_s = k_s (G_s . n)            (the flux variable of either phase, :256-291)
    numericalnormalflux = ({k G} - beta (-normalflux_1 + normalflux_2)) . n     (beta = edge_switch(V0, n), :300-307)
against the exact interface flux -q1 R / 2, and plots the relative errors (its stored PNGs have a 5 % axis).
Here:
    CutMesh / quadratures / MergedMesh!        ->  stefanproblem_b200.cutcell.LevelSetCutMesh (merge ratio 0.3)
    assemble_LDG_linear_system! / _rhs!        ->  stefan_ldgblocks_assemble / stefan_ldgblocks_rhs
    matrix \\ rhs                               ->  stefan_ldgblocks_solve (or the triplets + a host solve)
    closest points on the zero level set, map_to_reference_on_merged_mesh
                                               ->  the generator's interface quadrature points: they lie on the zero level
                                                   set of phi_h and come with their reference coordinates in both solution
                                                   cells and with the normal -grad phi_h / |grad phi_h|
    CutCellDG.interpolate_at_reference_points(G, 2, ...)  ->  stefan_interpolate_field_at_reference_points
    the flux formulas                          ->  stefan_ldg_interface_flux
`backend="oracle"` runs the same driver on the CPU restatement (tests/test_cutcell_cpu.py).

    python examples/cylinder_ldg_interface_flux.py [nelmts]
"""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

PRM = dict(k1=1.0, k2=2.0, q1=1.0, q2=2.0, R=0.3, b=1.0, Tw=1.0, center=(0.5, 0.5), theta=45.0,
           interiorpenalty=0.0, interfacepenalty=10.0, negboundarypenalty=0.0, posboundarypenalty=1.0, tinyratio=0.3)


def exact_solution(prm):
    """cylinder-analytical-solution.jl:3-39 around `center`"""
    q1, k1, q2, k2, R, b, Tw = (prm[k] for k in ("q1", "k1", "q2", "k2", "R", "b", "Tw"))
    cx, cy = prm["center"]
    A2 = (q2 - q1) / (2 * k2) * R ** 2
    B2 = Tw + q2 / (4 * k2) * b ** 2 - A2 * np.log(b)
    B1 = R ** 2 * (q1 / k1 - q2 / k2) / 4 + B2 + A2 * np.log(R)

    def u(x, y):
        rr = np.hypot(x - cx, y - cy)
        return np.where(rr < R, -q1 / (4 * k1) * rr ** 2 + B1, -q2 / (4 * k2) * rr ** 2 + A2 * np.log(np.maximum(rr, 1e-300)) + B2)
    return u


def interface_points(mesh):
    """The interface quadrature points seen from the + solution cells: (sol id of the + cell, sol id of the - cell,
    reference coordinates in either [2, npts], normals out of the + side [2, npts], angular position)."""
    d = mesh.data
    rows = []
    for sol in np.nonzero(mesh.sol_phase == 1)[0]:
        for s in np.nonzero(mesh.slot_is_interface[sol] & (d["slot_nbr"][sol] >= 0))[0]:
            keep = d["face_wts"][sol, s] != 0
            for q in np.nonzero(keep)[0]:
                rows.append((sol, int(d["slot_nbr"][sol, s]), d["face_pts"][sol, s, q], d["face_nbr_pts"][sol, s, q],
                             d["face_normals"][sol, s, q], mesh.face_xy[sol, s, q]))
    ids1 = np.array([r[0] for r in rows], dtype=np.int32)
    ids2 = np.array([r[1] for r in rows], dtype=np.int32)
    ref1 = np.array([r[2] for r in rows]).T
    ref2 = np.array([r[3] for r in rows]).T
    normals = np.array([r[4] for r in rows]).T
    xy = np.array([r[5] for r in rows])
    return ids1, ids2, ref1, ref2, normals, xy


def main(nelmts=33, solverorder=2, prm=PRM, backend="gpu", solver="schur"):
    from stefanproblem_b200.blocks import ldg_slot_penalties
    from stefanproblem_b200.cutcell import LevelSetCutMesh
    cx, cy = prm["center"]
    V0 = (np.cos(np.radians(prm["theta"])), np.sin(np.radians(prm["theta"])))
    numqp = solverorder + 1 + 2
    mesh = LevelSetCutMesh([0.0, 0.0], [1.0, 1.0], [nelmts, nelmts], lambda x, y: prm["R"] - np.hypot(x - cx, y - cy), 2,
                           numqp, prm["k1"], prm["k2"], 0.0, 0.0, merge_fraction=prm["tinyratio"])
    data = dict(mesh.data)
    data["slot_pen"] = ldg_slot_penalties(data, V0, prm["interiorpenalty"], prm["interfacepenalty"],
                                          prm["negboundarypenalty"], prm["posboundarypenalty"], mesh.slot_is_interface)
    u = exact_solution(prm)
    f = mesh.sample_source(lambda x, y: prm["q1"] + 0 * x, lambda x, y: prm["q2"] + 0 * x)
    g = mesh.sample_boundary(u)
    ids1, ids2, ref1, ref2, normals, xy = interface_points(mesh)
    if backend == "gpu":
        import torch
        from stefanproblem_b200 import blocks as B
        from stefanproblem_b200 import cutcelldg as cc
        from stefanproblem_b200 import postprocess as PP
        basis = cc.LagrangeTensorProductBasis(2, solverorder)
        system = B.LDGBlockSystem(mesh.nsol, solverorder, mesh.nqv, mesh.nslots, mesh.nqf).assemble(data, V0)
        rhs = system.rhs(f, g)
        sol, info = (system.solve(rhs, tol=1e-12, maxiter=200000) if solver == "schur" else (None, {"residual_norm": 1.0, "rhs_norm": 0.0}))
        if sol is None or not info["residual_norm"] <= 1e-9 * info["rhs_norm"]:
            import scipy.sparse as sp
            import scipy.sparse.linalg as spla
            r, c, v = (t.cpu().numpy() for t in system.to_coo())
            sol = torch.from_numpy(spla.spsolve(sp.coo_matrix((v, (r, c)), shape=system.shape).tocsc(), rhs.cpu().numpy())).cuda()
        grads1 = PP.interpolate_field_at_reference_points(sol, 3, 1, 2, basis, ref1, ids1)
        grads2 = PP.interpolate_field_at_reference_points(sol, 3, 1, 2, basis, ref2, ids2)
        nrm = torch.from_numpy(np.ascontiguousarray(normals)).cuda()
        nf1, nf2, numflux = (t.cpu().numpy() for t in PP.ldg_interface_flux(grads1, grads2, nrm, prm["k1"], prm["k2"], V0))
    else:
        import scipy.sparse as sp
        import scipy.sparse.linalg as spla
        from oracle import ldg_blocks as olb
        from oracle import postprocess as opp
        from oracle.basis import LagrangeTensorProductBasis
        basis = LagrangeTensorProductBasis(2, solverorder)
        A = sp.csc_matrix(olb.dense_matrix(olb.assemble_blocks(solverorder, data, V0), data["slot_nbr"]))
        sol = spla.spsolve(A, olb.assemble_rhs(solverorder, data, f, g))
        G = sol.reshape(-1, 3)[:, 1:3].T.copy()      # [2, nnodes], as `sol[2:3, :]` of the script
        grads1 = opp.interpolate_field_at_reference_points(G, 2, basis, ref1, ids1)
        grads2 = opp.interpolate_field_at_reference_points(G, 2, basis, ref2, ids2)
        nf1, nf2, numflux = opp.ldg_interface_flux(grads1, grads2, normals, prm["k1"], prm["k2"], V0)
    exactflux = -0.5 * prm["q1"] * prm["R"]
    err = {name: np.abs(v - exactflux) / abs(exactflux) for name, v in (("phase 1", nf1), ("phase 2", nf2), ("numerical", numflux))}
    theta = np.arctan2(xy[:, 1] - cy, xy[:, 0] - cx)
    print(f"nelmts {nelmts}  p = {solverorder}  solution cells {mesh.nsol}  interface points {len(theta)}  exact flux {exactflux:.6f}")
    for name, e in err.items():
        print(f"  relative error of the {name} normal flux: max {e.max():.3e}  mean {e.mean():.3e}")
    return theta, err


if __name__ == "__main__":
    main(int(sys.argv[1]) if len(sys.argv) > 1 else 33)
