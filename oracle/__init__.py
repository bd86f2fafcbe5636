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
  * the cut-cell studies the reference stores (`tests/test_cutcell_cpu.py`; block
    oracles `oracle/blocks.py`, `oracle/ldg_blocks.py` on data from the product's
    host-side generators): `plane_interface_tests` IP and LDG p = 2 to 3-4 digits
    (both geometries exact), `curved_interface_tests` and the cylinder studies at
    discretisation level (exact circle against a Q2-interpolated level set);
    `oracle/ldg_blocks.py` itself is pinned on uniform meshes against `oracle/fast.py`
PARITY UNPINNED for: digit-level cut-cell results on curved interfaces (the
reference's geometry lives in un-vendored packages), the mesh-aligned 2-D
two-phase extension (face-level functions are the reference's, the driver loop is
ours), and time stepping (no reference implementation exists).
"""
