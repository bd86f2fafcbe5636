"""CPU oracle for the StefanProblem hot path.  TEST INFRASTRUCTURE ONLY.

This package is a NumPy/SciPy restatement of the reference's Julia assembly
(`/root/reference/Stefan{DG,FD,RobinIC}`), written function by function with the
reference file:line each one follows.  Nothing in `stefanproblem_b200/` (the
product) imports it; only `tests/`, `__graft_entry__.smoke()` and the
`cpu_baseline` / `--impl reference` legs of `bench.py` may.

Why a restatement and not the reference itself: the reference is Julia (not
installed here or on the GPU box) and its mesh / basis / quadrature layer lives in
three un-vendored, un-pinned packages (`CutCellDG`, `PolynomialBasis`,
`ImplicitDomainQuadrature`, GitHub `ArjunNarayanan/*.jl`).  Their conventions are
restated in `oracle/basis.py` and `oracle/mesh.py`.

Pins (see `tests/test_oracle_golden.py`):
  * `StefanDG/LocalDG/uncut_mesh_tests/MD-LDG_convergence.csv`      -> <= 1e-12 rel
  * `StefanDG/InteriorPenalty/uncut_mesh_tests/IP_convergence.csv`  -> <= 1e-10 rel
    (legacy boundary variant)
  * the reference's own `@test`s for DG1D (symmetry, rate > 1.95, p=2 error
    < 1e6 eps, Robin flux residual < 1e3 eps)
  * the FD known-answer matrix of `StefanFD/test-linear-heat-equation.jl`
PARITY UNPINNED for: cut-cell / merged-cell configurations, the mesh-aligned 2-D
two-phase extension (face-level functions are the reference's, the driver loop is
ours), and time stepping (no reference implementation exists).
"""
