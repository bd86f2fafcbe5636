"""Cut-cell regime of the hot path on the GPU: the element-block path fed by the circle
cut-cell generator -- blocks, right-hand side and error norm against the oracle on the SAME data, the device CG
against the host's sparse direct solve, and the reference's two cylinder studies end to end (rates of
test_IP_convergence.jl:125,156 and the stored CSVs at discretisation level)."""
import importlib.util
import json
import os

import numpy as np
import pytest

from oracle import blocks as ob
from oracle.analytical import CylindricalSolver, RobinCylindricalSolver

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLD = json.load(open(os.path.join(ROOT, "tests", "golden", "reference_csv.json")))


def _example():
    spec = importlib.util.spec_from_file_location("cyl", os.path.join(ROOT, "examples", "cylinder_ip_convergence.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


@pytest.mark.parametrize("order,numqp", [(1, 2), (2, 5), (3, 4)])
@pytest.mark.parametrize("lam,merge", [(0.0, 0.0), (1.0, 0.25)])
def test_cut_cell_blocks_rhs_and_norm_match_oracle(cuda, order, numqp, lam, merge):
    import torch
    from stefanproblem_b200 import blocks as B
    from stefanproblem_b200.cutcell import CircleCutMesh
    n, R, Tm = 5, 0.62, 0.5
    m = CircleCutMesh([0, 0], [1, 1], [n, n], [0, 0], R, numqp, 1.0, 2.0, 1e3 * n, lam, merge_fraction=merge)
    sol = RobinCylindricalSolver(2.0, 1.0, 1.0, 2.0, 1.0, R, 1.5, 1.0, Tm) if lam else CylindricalSolver(1.0, 1.0, 2.0, 2.0, R, 1.5, 1.0)
    ref = ob.assemble_blocks(order, m.data)
    system = B.BlockSystem(m.nsol, order, m.nqv, m.nslots, m.nqf).assemble(m.data)
    nf = (order + 1) ** 2
    r, c, v = (t.cpu().numpy() for t in system.to_coo())
    got = v.reshape(m.nsol, 1 + m.nslots, nf, nf)
    live = np.concatenate([np.ones((m.nsol, 1), bool), m.data["slot_nbr"] >= 0], axis=1)
    assert np.abs(got[live] - ref[live]).max() <= 1e-12 * np.abs(ref).max()
    f = m.sample_source(lambda x, y: 1.0 + x * y, lambda x, y: 2.0 - x + 0 * y)
    g = m.sample_boundary(sol.at)
    b_ref = ob.assemble_rhs(order, m.data, f, g, Tm)
    b = system.rhs(f, g, Tm)
    assert np.linalg.norm(b.cpu().numpy() - b_ref) <= 1e-12 * np.linalg.norm(b_ref)
    u = np.random.default_rng(3).uniform(-1, 1, m.nsol * nf)
    uex = m.sample_exact(sol.at)
    e_ref, n_ref = ob.l2_error(order, m.data, u, uex)
    e, nn = system.l2_error(torch.from_numpy(u).cuda(), uex)
    assert abs(e - e_ref) <= 1e-12 * e_ref and abs(nn - n_ref) <= 1e-12 * n_ref
    # device CG (block Jacobi) against the sparse direct solve of the oracle's matrix
    import scipy.sparse as sp
    import scipy.sparse.linalg as spla
    A = sp.csc_matrix(ob.dense_matrix(ref, m.data["slot_nbr"]))
    x_ref = spla.spsolve(A, b_ref)
    x, info = system.solve(b, tol=1e-13, maxiter=100000)
    if merge:   # (without merging the tiny cells make the system too ill-conditioned for an iterative solve)
        assert info["converged"], info
        assert np.linalg.norm(x.cpu().numpy() - x_ref) <= 1e-7 * np.linalg.norm(x_ref)


@pytest.mark.parametrize("order", [1, 2, 3])
def test_blocks_legacy_boundary_variant_matches_oracle(cuda, order):
    """stefan_blocks_set_variant(1): boundary blocks -M' + pen P and boundary right-hand side pen sum V g w (the form the
    reference's stored cut-cell IP files were produced with), on interpolated-level-set data, against the oracle."""
    from stefanproblem_b200 import blocks as B
    from stefanproblem_b200.cutcell import LevelSetCutMesh
    n = 5
    m = LevelSetCutMesh([0, 0], [1, 1], [n, n], lambda x, y: 0.55 - np.hypot(x, y), 2, order + 3, 1.0, 2.0, 1e3 * n, 0.0,
                        merge_fraction=0.2)
    ref = ob.assemble_blocks(order, m.data, "legacy_boundary")
    system = B.BlockSystem(m.nsol, order, m.nqv, m.nslots, m.nqf, variant="legacy_boundary").assemble(m.data)
    nf = (order + 1) ** 2
    got = system.to_coo()[2].cpu().numpy().reshape(m.nsol, 1 + m.nslots, nf, nf)
    live = np.concatenate([np.ones((m.nsol, 1), bool), m.data["slot_nbr"] >= 0], axis=1)
    assert np.abs(got[live] - ref[live]).max() <= 1e-12 * np.abs(ref).max()
    assert np.abs(ref - ob.assemble_blocks(order, m.data)).max() > 1e-3 * np.abs(ref).max()   # (the variants differ)
    f = m.sample_source(lambda x, y: 1.0 + x * y, lambda x, y: 2.0 - x + 0 * y)
    g = m.sample_boundary(lambda x, y: np.cos(x) + y)
    b_ref = ob.assemble_rhs(order, m.data, f, g, 0.0, "legacy_boundary")
    assert np.linalg.norm(system.rhs(f, g, 0.0).cpu().numpy() - b_ref) <= 1e-12 * np.linalg.norm(b_ref)


def _gpu_tol(tight, n, loose_n5=6e-3):
    """GPU studies end with an iterative solve (or a host solve of device-assembled triplets): the digit-level pins live
    in tests/test_cutcell_cpu.py; here the same numbers to 2e-4 (n = 5: only reproduced to 5e-3 by the oracle as well, n = 33: solver
    accuracy against discretisation errors of 1e-8)."""
    if tight is None:
        return None
    return max(tight, loose_n5 if n == 5 else (2e-3 if n == 33 else 2e-4))


@pytest.mark.parametrize("robin", [False, True])
def test_cylinder_convergence_study_on_the_gpu(cuda, robin):
    """IP_test_cylindrical_bc_convergence.jl / IP_test_cylndrical_robin_bc_convergence.jl, n = 3 ... 33, p = 1, 2,
    end to end on the device (cut-cell data of the reference's Q2 level set -> blocks -> rhs -> CG or triplets + host
    solve -> L2 error).  Acceptance as in the reference's test_IP_convergence.jl:125,156 (mean rate > 1.8 / > 2.8) and
    the stored CSVs: Robin study and plain p = 2 to 2e-4 (CPU oracle: 1e-6 ... 1e-10), plain p = 1 (2-point rule) to 2 %."""
    mod = _example()
    nelmts, table = mod.main(robin=robin)
    gold = GOLD["ip_cylinder_robin" if robin else "ip_cylinder"]
    assert [g[0] for g in gold] == nelmts
    for order, col, bar in ((1, 1, 1.8), (2, 2, 2.8)):
        err, rate = table[order]
        assert np.mean(rate[1:]) > bar, (order, rate)
        for e, g in zip(err, gold):
            tol = 0.02 if (not robin and order == 1) else _gpu_tol(1e-6, g[0])
            assert abs(e / g[col] - 1.0) <= tol, (order, g[0], e, g[col])


# ----------------------------------------------------------------- local DG on element blocks (stefan_ldgblocks_*)
@pytest.mark.parametrize("order,numqp", [(1, 4), (2, 5), (3, 4)])
@pytest.mark.parametrize("merge", [0.0, 0.25])
def test_ldg_cut_cell_blocks_rhs_apply_and_solve_match_oracle(cuda, order, numqp, merge):
    """stefan_ldgblocks_assemble / _rhs / _apply / _assemble_coo / _solve on the circle cut-cell data against
    oracle/ldg_blocks.py (the reference's face-level LDG functions on the same data) and the host's sparse direct
    solve."""
    import scipy.sparse as sp
    import scipy.sparse.linalg as spla
    import torch
    from oracle import ldg_blocks as olb
    from stefanproblem_b200 import blocks as B
    from stefanproblem_b200.cutcell import CircleCutMesh
    n, R = 5, 0.62
    V0 = (np.cos(0.7), np.sin(0.7))
    m = CircleCutMesh([0, 0], [1, 1], [n, n], [0, 0], R, numqp, 1.0, 2.0, 0.0, 0.0, merge_fraction=merge)
    data = dict(m.data)
    data["slot_pen"] = B.ldg_slot_penalties(data, V0, 0.3, 0.8, 0.2, 1.1, m.slot_is_interface)
    sol = CylindricalSolver(1.0, 1.0, 2.0, 2.0, R, 1.5, 1.0)
    ref = olb.assemble_blocks(order, data, V0)
    system = B.LDGBlockSystem(m.nsol, order, m.nqv, m.nslots, m.nqf).assemble(data, V0)
    n3 = 3 * (order + 1) ** 2
    r, c, v = (t.cpu().numpy() for t in system.to_coo())
    got = v.reshape(m.nsol, 1 + m.nslots, n3, n3)
    live = np.concatenate([np.ones((m.nsol, 1), bool), data["slot_nbr"] >= 0], axis=1)
    assert np.abs(got[live] - ref[live]).max() <= 1e-12 * np.abs(ref).max()
    A = sp.csr_matrix(olb.dense_matrix(ref, data["slot_nbr"]))
    Acoo = sp.coo_matrix((v, (r, c)), shape=system.shape).tocsr()
    assert abs(Acoo - A).max() <= 1e-12 * abs(A).max()
    x = np.random.default_rng(5).uniform(-1, 1, system.ndofs)
    y = system.apply(torch.from_numpy(x).cuda()).cpu().numpy()
    assert np.linalg.norm(y - A @ x) <= 1e-13 * np.linalg.norm(A @ x)
    f = m.sample_source(lambda x, y: 1.0 + x * y, lambda x, y: 2.0 - x + 0 * y)
    g = m.sample_boundary(sol.at)
    b_ref = olb.assemble_rhs(order, data, f, g)
    b = system.rhs(f, g)
    assert np.linalg.norm(b.cpu().numpy() - b_ref) <= 1e-12 * np.linalg.norm(b_ref)
    # (slivers -- and, at order 3, the cubic mass blocks of the merged pieces -- make the Schur complement too
    # ill-conditioned for an unpreconditioned iteration: there the solver reports `converged: False` and a driver falls
    # back to the triplets + a direct solve, as examples/cylinder_ldg_convergence.py does)
    if merge and order <= 2:
        x_ref = spla.spsolve(A.tocsc(), b_ref)
        xs, info = system.solve(b, tol=1e-12, maxiter=200000)
        assert info["residual_norm"] <= 1e-10 * info["rhs_norm"], info
        assert np.linalg.norm(xs.cpu().numpy() - x_ref) <= 1e-6 * np.linalg.norm(x_ref)
        # component norms through the scalar block norm kernel
        e, den = system.l2_error(xs, m.sample_exact(sol.at), 0)
        e_ref, den_ref = ob.l2_error(order, data, x_ref[0::3], m.sample_exact(sol.at))
        assert abs(den - den_ref) <= 1e-12 * den_ref and abs(e - e_ref) <= 1e-6 * max(e_ref, 1e-9 * den_ref)


def test_ldg_blocks_on_a_uniform_mesh_equal_the_matrix_free_ldg_operator(cuda):
    """Same system two ways on the device: the local-DG block path fed uniform-mesh data against the matrix-free
    march / plain kernels of stefan_ldg_apply (two-phase mesh with the aligned-interface coupling)."""
    import torch
    from oracle import fast
    from oracle.mesh import UniformMesh2D
    from stefanproblem_b200 import blocks as B
    from stefanproblem_b200 import cutcelldg as cc
    for order in (1, 2, 3):
        nx, ny, nq = 9, 7, order + 1
        basis = cc.LagrangeTensorProductBasis(2, order)
        omesh = UniformMesh2D([0, 0], [1, 1], [nx, ny], basis.nf)
        ph = fast.circle_phase(omesh, radius=0.33)
        V0 = (np.cos(0.4), np.sin(0.4))
        op = cc.LDGOperator(cc.UniformMesh([0, 0], [1, 1], [nx, ny], basis, phase=ph), order, nq, 1.0, 2.0, 0.5, 0.2, 1.3,
                            V0, interfacepenalty=0.7)
        data = B.uniform_mesh_block_data(nx, ny, 1.0 / nx, 1.0 / ny, nq, ph, 1.0, 2.0, 0.0, 0.0, aligned_interface=True)
        nbr = data["slot_nbr"]
        phc = np.asarray(ph).reshape(-1)
        isif = (nbr >= 0) & (phc[np.maximum(nbr, 0)] != phc[:, None])
        data["slot_pen"] = B.ldg_slot_penalties(data, V0, 0.5, 0.7, 0.2, 1.3, isif)
        system = B.LDGBlockSystem(nx * ny, order, nq * nq, 4, nq).assemble(data, V0)
        x = torch.from_numpy(np.random.default_rng(order).uniform(-1, 1, system.ndofs)).cuda()
        y_blocks, y_free = system.apply(x), op.apply(x, algo=1)
        assert float((y_blocks - y_free).norm()) <= 1e-12 * float(y_free.norm())


def test_ldg_cylinder_convergence_study_on_the_gpu(cuda):
    """cylindrical_bc_tests/LDG_test_cylindrical_bc_convergence.jl, n = 3 ... 33, p = 1, 2, end to end on the device
    (cut-cell data of the reference's Q2 level set -> local-DG blocks -> rhs -> Schur BiCGStab -> component L2 errors)
    against the stored LDG_convergence.csv: all six columns to 2e-4 (n = 5: 6e-3, n = 33: 2e-3; the CPU oracle pins the
    same numbers to 2e-6), rates as the CSV's."""
    spec = importlib.util.spec_from_file_location("cyl_ldg", os.path.join(ROOT, "examples", "cylinder_ldg_convergence.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    nelmts, table = mod.main()
    gold = GOLD["ldg_cylinder"]
    assert [g[0] for g in gold] == nelmts
    err1, rate1 = table[1]
    err2, rate2 = table[2]
    for e1, e2, g in zip(err1, err2, gold):
        for col in range(3):
            assert abs(e1[col] / g[1 + col] - 1.0) <= _gpu_tol(2e-6, g[0]), (1, g[0], col, e1, g)
            assert abs(e2[col] / g[4 + col] - 1.0) <= _gpu_tol(2e-6, g[0]), (2, g[0], col, e2, g)
    assert np.mean(rate1[1:, 0]) > 1.7 and np.mean(rate1[1:, 1:]) > 1.2   # CSV: 1.95 (T), 1.45 (flux)
    assert np.mean(rate2[1:, 0]) > 2.6 and np.mean(rate2[1:, 1:]) > 2.0   # CSV: 2.98 (T), 2.66 (flux)


@pytest.mark.parametrize("kind", ["plane", "curved"])
@pytest.mark.parametrize("method", ["ip", "ldg"])
def test_cut_interface_convergence_studies_on_the_gpu(cuda, kind, method):
    """plane_interface_tests / curved_interface_tests of the reference (IP and local DG, p = 1, 2, n = 3 ... 33) end to
    end on the device against the stored CSVs: 2e-4 wherever tests/test_cutcell_cpu.py pins the same numbers to 1e-7
    (IP plane and curved, LDG plane and curved)."""
    spec = importlib.util.spec_from_file_location("cut_ex", os.path.join(ROOT, "examples", "cut_interface_convergence.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    nelmts, table = mod.main(kind, method)
    gold = GOLD[f"{method}_{kind}"]
    assert [g[0] for g in gold] == nelmts
    tight = {("plane", "ip", 1): 1e-7, ("plane", "ip", 2): 1e-7, ("plane", "ldg", 1): 2e-3, ("plane", "ldg", 2): 1e-7,
             ("curved", "ip", 1): 5e-5, ("curved", "ip", 2): 5e-5, ("curved", "ldg", 1): 3e-5, ("curved", "ldg", 2): 1e-5}
    nc = 1 if method == "ip" else 3
    for order in (1, 2):
        err, rate = table[order]
        for e, g in zip(err, gold):
            want = g[1 + nc * (order - 1):1 + nc * order]
            t = _gpu_tol(tight[kind, method, order], g[0])
            for ev, wv in zip(e, want):
                if t is None:
                    assert 0.4 * wv < ev < 3.0 * wv, (kind, method, order, g[0], e, want)
                else:
                    assert abs(ev / wv - 1.0) <= t, (kind, method, order, g[0], e, want)
        assert np.mean(rate[1:, 0]) > (1.8 if order == 1 else 2.8), (kind, method, order, rate)
