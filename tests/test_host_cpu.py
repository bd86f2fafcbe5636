"""CPU-side checks of the product's host logic (no GPU, no compute entry points):
the C-ABI library loads and exports every symbol include/stefan_b200.h declares; the
reference-element tables; the per-cell row functions run on the host (test harness
tests/hostcheck) against the oracle; the C oracle against the NumPy oracle."""
import ctypes
import os
import re
import subprocess
import sys

import numpy as np
import pytest

from oracle import fast
from oracle.basis import gauss_1d, lagrange_1d
from oracle.mesh import UniformMesh2D

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_library_exports_every_declared_symbol(built_lib):
    header = open(os.path.join(ROOT, "include", "stefan_b200.h")).read()
    declared = set(re.findall(r"\b(stefan_[a-z0-9_]+)\s*\(", header))
    L = ctypes.CDLL(built_lib)
    for name in sorted(declared):
        assert hasattr(L, name), f"{name} declared in stefan_b200.h but not exported"
    from stefanproblem_b200 import _lib
    assert declared == set(_lib.exported_symbols()), "ctypes binding and header disagree"


def test_missing_gpu_fails_loudly(built_lib):
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    from stefanproblem_b200 import StefanError, cutcelldg as cc
    basis = cc.LagrangeTensorProductBasis(2, 1)
    mesh = cc.UniformMesh([0, 0], [1, 1], [4, 4], basis)
    with pytest.raises(StefanError):
        cc.IPOperator(mesh, 1, 2, 1.0, 1.0, 1e3)


@pytest.mark.parametrize("order,numqp", [(1, 2), (2, 3), (3, 4), (2, 5)])
def test_reference_tables(built_lib, order, numqp):
    from stefanproblem_b200 import cutcelldg as cc
    t = cc.reference_tables(order, numqp)
    x, w = gauss_1d(numqp)
    assert np.allclose(t["qpoints"], x, atol=1e-15) and np.allclose(t["qweights"], w, atol=1e-15)
    n = order + 1
    M, K, C = np.zeros((n, n)), np.zeros((n, n)), np.zeros((n, n))
    for xq, wq in zip(x, w):
        N, dN = lagrange_1d(order, xq)
        M += wq * np.outer(N, N)
        K += wq * np.outer(dN, dN)
        C += wq * np.outer(dN, N)
    assert np.allclose(t["mass"], M, atol=1e-15)
    assert np.allclose(t["stiff"], K, atol=1e-14)
    assert np.allclose(t["adv"], C, atol=1e-15)
    assert np.allclose(t["dleft"], lagrange_1d(order, -1.0)[1], atol=1e-14)
    assert np.allclose(t["dright"], lagrange_1d(order, 1.0)[1], atol=1e-14)


@pytest.fixture(scope="module")
def hostcheck():
    hc = os.path.join(ROOT, "tests", "hostcheck")
    so = os.path.join(hc, "_hostcheck.so")
    subprocess.run(["g++", "-O2", "-shared", "-fPIC", "-I/usr/local/cuda/include", "-o", so,
                    os.path.join(hc, "hostcheck.cpp")], check=True)
    return ctypes.CDLL(so)


@pytest.mark.parametrize("order", [1, 2, 3])
@pytest.mark.parametrize("variant", ["current", "legacy_boundary"])
@pytest.mark.parametrize("terms", [255, 1 | 2 | 4 | 32 | 64, 1, 2 | 4, 32 | 64, 8 | 16 | 128])
def test_cell_row_functions_match_oracle(hostcheck, order, variant, terms):
    """The __host__ __device__ row functions the kernels call, looped over the cells on
    the CPU, reproduce `A_oracle @ x` for every combination of assemble passes."""
    nx, ny = 7, 6
    mesh = UniformMesh2D([0, 0], [1.0, 0.8], [nx, ny], (order + 1) ** 2)
    mesh.cellsign[:] = fast.circle_phase(mesh, (0.5, 0.4), 0.3)
    k1, k2, pen, lam = 1.0, 2.0, 1e3 * nx, 1.5
    A = ip_matrix_terms(mesh, order, k1, k2, pen, lam, variant, terms)
    x = np.random.default_rng(0).uniform(-1, 1, A.shape[0])
    y = np.zeros_like(x)
    ph = np.ascontiguousarray(mesh.cellsign, dtype=np.int8)
    jx, jy = mesh.jacobian()
    dp = ctypes.POINTER(ctypes.c_double)
    d = ctypes.c_double
    rc = hostcheck.hostcheck_ipdg_apply(order, nx, ny, order + 1, d(jx), d(jy), d(k1), d(k2), d(pen), d(lam),
                                        {"current": 0, "legacy_boundary": 1}[variant], terms,
                                        ph.ctypes.data_as(ctypes.POINTER(ctypes.c_int8)),
                                        x.ctypes.data_as(dp), y.ctypes.data_as(dp))
    assert rc == 0
    yr = A @ x
    scale = np.linalg.norm(yr) if np.linalg.norm(yr) > 0 else 1.0
    assert np.linalg.norm(y - yr) / scale < 1e-13


@pytest.mark.parametrize("order", [1, 2, 3])
@pytest.mark.parametrize("two_phase", [False, True])
@pytest.mark.parametrize("theta", [45.0, 143.0, 0.0, 270.0])
def test_ldg_cell_row_functions_match_oracle(hostcheck, order, two_phase, theta):
    nx, ny = 7, 6
    mesh = UniformMesh2D([0, 0], [1.0, 0.8], [nx, ny], (order + 1) ** 2)
    if two_phase:
        mesh.cellsign[:] = fast.circle_phase(mesh, (0.5, 0.4), 0.3)
    V0 = np.array([np.cos(np.radians(theta)), np.sin(np.radians(theta))])
    if theta in (0.0, 270.0):
        V0 = np.round(V0)  # exactly axis-aligned: sign(V0 . n) = 0 on the other axis
    k1, k2, ai, an, ap = 1.0, 2.0, 0.7, 0.3, 1.1
    A = fast.ldg_matrix(mesh, order, order + 1, k1, k2, ai, an, ap, V0)
    x = np.random.default_rng(0).uniform(-1, 1, A.shape[0])
    y = np.zeros_like(x)
    ph = np.ascontiguousarray(mesh.cellsign, dtype=np.int8)
    jx, jy = mesh.jacobian()
    dp, d = ctypes.POINTER(ctypes.c_double), ctypes.c_double
    rc = hostcheck.hostcheck_ldg_apply(order, nx, ny, order + 1, d(jx), d(jy), d(k1), d(k2), d(ai), d(an),
                                       d(ap), d(V0[0]), d(V0[1]),
                                       ph.ctypes.data_as(ctypes.POINTER(ctypes.c_int8)),
                                       x.ctypes.data_as(dp), y.ctypes.data_as(dp))
    assert rc == 0
    yr = A @ x
    assert np.linalg.norm(y - yr) / np.linalg.norm(yr) < 1e-13


@pytest.mark.parametrize("order", [1, 2, 3])
def test_ldg_aligned_interface_rows_match_oracle(hostcheck, order):
    """Mesh-aligned LDG interface (the face-level scalar / vector flux functions on unlike-phase faces,
    alpha = interfacepenalty, neighbour's conductivity on the neighbour's flux): the kernels' row function with the
    T_IFACE tables against the oracle (fast.ldg_matrix(interfacepenalty=...), itself equal to the literal drivers
    oracle/ldg.assemble_aligned_interface_*_operator)."""
    nx, ny = 7, 6
    mesh = UniformMesh2D([0, 0], [1.0, 0.8], [nx, ny], (order + 1) ** 2)
    mesh.cellsign[:] = fast.circle_phase(mesh, (0.5, 0.4), 0.3)
    V0 = np.array([np.cos(0.4), np.sin(0.4)])
    k1, k2, ai, an, ap, aif = 1.0, 2.0, 0.7, 0.3, 1.1, 0.9
    A = fast.ldg_matrix(mesh, order, order + 1, k1, k2, ai, an, ap, V0, interfacepenalty=aif)
    x = np.random.default_rng(0).uniform(-1, 1, A.shape[0])
    y = np.zeros_like(x)
    ph = np.ascontiguousarray(mesh.cellsign, dtype=np.int8)
    jx, jy = mesh.jacobian()
    dp, d = ctypes.POINTER(ctypes.c_double), ctypes.c_double
    rc = hostcheck.hostcheck_ldg_apply_aligned(order, nx, ny, order + 1, d(jx), d(jy), d(k1), d(k2), d(ai), d(an), d(ap),
                                               d(V0[0]), d(V0[1]), d(aif),
                                               ph.ctypes.data_as(ctypes.POINTER(ctypes.c_int8)),
                                               x.ctypes.data_as(dp), y.ctypes.data_as(dp))
    assert rc == 0
    yr = A @ x
    assert np.linalg.norm(y - yr) / np.linalg.norm(yr) < 1e-13


def ip_matrix_terms(mesh, order, k1, k2, pen, lam, variant, terms):
    """Oracle matrix for a subset of the reference's assemble passes (literal loops)."""
    from oracle import ip
    from oracle.basis import LagrangeTensorProductBasis
    from oracle.mesh import CellQuadratures, FaceQuadratures, SystemMatrix, sparse_operator
    basis = LagrangeTensorProductBasis(2, order)
    cq, fq = CellQuadratures(order + 1), FaceQuadratures(order + 1)
    A = SystemMatrix()
    if terms & 1:
        ip.assemble_gradient_operator(A, basis, cq, k1, k2, mesh)
    if terms & 2:
        ip.assemble_interelement_flux_operator(A, basis, fq, k1, k2, mesh)
    if terms & 4:
        ip.assemble_interelement_penalty_operator(A, basis, fq, pen, mesh)
    if terms & 8:
        ip.assemble_aligned_interface_flux_operator(A, basis, fq, k1, k2, mesh)
    if terms & 16:
        ip.assemble_aligned_interface_penalty_operator(A, basis, fq, pen, mesh)
    if terms & 32:
        ip.assemble_boundary_flux_operator(A, basis, fq, k1, k2, mesh, variant)
    if terms & 64:
        ip.assemble_boundary_penalty_operator(A, basis, fq, pen, mesh)
    if terms & 128:
        ip.assemble_aligned_robin_operator(A, basis, fq, lam, mesh)
    return sparse_operator(A, mesh.number_of_nodes())


@pytest.mark.parametrize("order", [1, 2, 3])
def test_c_oracle_matches_numpy_oracle(order):
    from oracle.cport import CIPOracle
    nx, ny = 9, 7
    mesh = UniformMesh2D([0, 0], [1.0, 0.7], [nx, ny], (order + 1) ** 2)
    mesh.cellsign[:] = fast.circle_phase(mesh, (0.5, 0.35), 0.3)
    for variant in ("current", "legacy_boundary"):
        A = fast.ip_matrix(mesh, order, order + 1, 1.0, 2.0, 1e3 * nx, 1.5, variant)
        C = CIPOracle(order, order + 1, nx, ny, mesh.h[0], mesh.h[1], 1.0, 2.0, 1e3 * nx, 1.5, variant,
                      255, mesh.cellsign)
        assert abs(A - C.csr()).max() < 1e-14 * abs(A).max()
        x = np.random.default_rng(0).uniform(-1, 1, A.shape[0])
        for threads in (1, 0):
            assert np.linalg.norm(C.spmv(x, nthreads=threads) - A @ x) < 1e-13 * np.linalg.norm(A @ x)


@pytest.mark.parametrize("nx,ny,warps,two_phase", [(40, 95, 8, True), (7, 6, 8, False), (257, 131, 4, True),
                                                    (33, 1, 8, False), (1, 64, 4, True), (600, 90, 8, True)])
@pytest.mark.parametrize("cbeg_frac,cend_frac,max_ctas", [(0.0, 1.0, 296), (0.0, 1.0, 7), (0.25, 0.8, 148)])
def test_march_plan_invariants(hostcheck, nx, ny, warps, two_phase, cbeg_frac, cend_frac, max_ctas):
    """Host plan of the march kernels (csrc/march_plan.h): the loop's store masks and the fix
    list partition the cells exactly (interior-uniform cells vs the rest), phase masks and
    modes are consistent with the phases, and the CTA tiles cover every (strip group, column)
    of the requested range exactly once within the CTA budget."""
    rng = np.random.default_rng(nx * 1000 + ny)
    phase = np.ones((nx, ny), dtype=np.int8)
    if two_phase:
        i, j = np.meshgrid(np.arange(nx), np.arange(ny), indexing="ij")
        phase[((i + 0.5) / nx - 0.5) ** 2 + ((j + 0.5) / ny - 0.45) ** 2 < 0.1] = -1
        phase[rng.integers(0, nx, 5), rng.integers(0, ny, 5)] *= -1  # a few isolated cells
    cbeg, cend = int(cbeg_frac * nx), max(int(cend_frac * nx), int(cbeg_frac * nx) + 1)
    nstrips = (ny + 29) // 30
    runs = np.zeros(4 * (nstrips * (nx + 3)), dtype=np.int32)
    run_ofs = np.zeros(nstrips + 1, dtype=np.int32)
    fix = np.zeros(nx * ny, dtype=np.int32)
    tiles = np.zeros(4 * 640, dtype=np.int32)
    n_out = np.zeros(4, dtype=np.int32)
    ip = lambda a: a.ctypes.data_as(ctypes.POINTER(ctypes.c_int))  # noqa: E731
    rc = hostcheck.hostcheck_march_plan(nx, ny, np.ascontiguousarray(phase).ctypes.data_as(ctypes.POINTER(ctypes.c_int8)),
                                        warps, cbeg, cend, max_ctas, ip(runs), ip(run_ofs), ip(fix), ip(tiles), ip(n_out))
    assert rc == 0
    nruns, ns, nfix, ntiles = (int(v) for v in n_out)
    assert ns == nstrips
    runs = runs[:4 * nruns].reshape(-1, 4)
    # expected classification
    P = np.zeros((nx + 2, ny + 2), dtype=np.int8)
    P[1:-1, 1:-1] = np.where(phase == 1, 1, 2)
    c = P[1:-1, 1:-1]
    uniform = (P[:-2, 1:-1] == c) & (P[2:, 1:-1] == c) & (P[1:-1, :-2] == c) & (P[1:-1, 2:] == c)
    # fix list = the non-uniform cells, sorted by cell id (hence by column)
    expect_fix = np.flatnonzero(~uniform.ravel())
    assert nfix == expect_fix.size and np.array_equal(fix[:nfix], expect_fix)
    # run lists reproduce the masks column by column
    for s in range(nstrips):
        k, col = int(run_ofs[s]), 0
        j0, j1 = 30 * s, min(30 * s + 30, ny)
        while col < nx:
            end, mode, pm, sm = (int(v) for v in runs[k])
            pm &= 0xffffffff
            sm &= 0xffffffff
            assert col < end <= nx
            for i in range(col, end):
                exp_s = sum(1 << (j - j0 + 1) for j in range(j0, j1) if uniform[i, j])
                exp_p = sum(1 << (j - j0 + 1) for j in range(j0, j1) if uniform[i, j] and c[i, j] == 2)
                assert sm == exp_s and pm == exp_p
            want = 0 if sm == 0 else (1 if pm == 0 else (2 if pm == sm else 3))
            assert (mode & 3) == want
            if sm:
                t = sm >> ((sm & -sm).bit_length() - 1)
                assert bool(mode & 4) == ((t & (t + 1)) == 0)
            col, k = end, k + 1
        assert all(int(v) == 2 ** 31 - 1 for v in runs[k:k + 3, 0])  # sentinels
    # tiles: exact cover of [cbeg, cend) per group, within budget (or one tile per group)
    ngroups = (nstrips + warps - 1) // warps
    tiles = tiles[:4 * ntiles].reshape(-1, 4)
    assert ntiles <= max(max_ctas, ngroups) and ntiles >= ngroups
    for g in range(ngroups):
        seg = sorted((int(t[1]), int(t[2])) for t in tiles if t[0] == g)
        assert seg[0][0] == cbeg and seg[-1][1] == cend
        assert all(a[1] == b[0] for a, b in zip(seg, seg[1:])) and all(a < b for a, b in seg)
        if not two_phase:  # equal-cost tiles: without mixed columns a group's tiles differ by one column at most
            lens = [b - a for a, b in seg]
            assert max(lens) - min(lens) <= 1
    if not two_phase and ngroups > 0:  # the whole budget is handed out (tiles of >= 4 columns)
        per = [sum(1 for t in tiles if t[0] == g) for g in range(ngroups)]
        assert max(per) - min(per) <= 1
        assert ntiles == max(ngroups, min(max_ctas, 640, ngroups * max(1, (cend - cbeg) // 4))) or max(per) == max(1, (cend - cbeg) // 4)


@pytest.mark.parametrize("order", [1, 2, 3])
@pytest.mark.parametrize("s", [0, 1])
def test_ldg_streamed_rows_equal_general_rows(hostcheck, order, s):
    """The march kernel's streamed LDG row function (traces in, no stored intermediates)
    equals the general per-cell row function, which is checked against the oracle above."""
    rng = np.random.default_rng(order * 10 + s)
    cd = 3 * (order + 1) ** 2
    x9 = rng.uniform(-1, 1, 9 * cd)
    # T_SAME everywhere, then all 81 face-type combinations, then the progressive form (order-3 march loop)
    g, st = np.zeros(83 * cd), np.zeros(83 * cd)
    d = ctypes.c_double
    dp = lambda a: a.ctypes.data_as(ctypes.POINTER(ctypes.c_double))  # noqa: E731
    rc = hostcheck.hostcheck_ldg_streamed(order, order + 1, d(0.013), d(0.021), d(1.0), d(2.0), d(0.7), d(0.3), d(1.1),
                                          d(np.cos(0.6)), d(np.sin(0.6)), s, dp(x9), dp(g), dp(st))
    assert rc == 0
    assert np.abs(g - st).max() <= 1e-13 * np.abs(g).max()


def test_bench_reference_arm_prints_the_contract_line():
    """`bench.py --impl reference` (the CPU path of the reference restated in C) runs without a
    GPU and prints one JSON line with the keys the driver reads."""
    import json
    import sys
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1",
                        "--warmup", "1"], capture_output=True, text=True, timeout=300, cwd=ROOT)
    assert r.returncode == 0, r.stderr[-2000:]
    line = json.loads(r.stdout.strip().splitlines()[-1])
    assert line["impl"] == "reference" and line["unit"] == "DOF-updates/s" and line["value"] > 0
    assert line["metric"].startswith("fp64 DOF-updates/sec") and line["higher_is_better"] is True
    assert line["cpu_baseline"]["kind"] == "port" and line["cpu_baseline"]["cores"] >= 1
    assert line["e2e"]["h2d_bytes_per_step"] == 0 and line["e2e"]["d2h_bytes_per_step"] == 0


def test_header_is_plain_c(tmp_path, built_lib):
    """include/stefan_b200.h is what a cgo / ccall / ctypes binding reads: it must compile as
    C99 on its own (no C++-isms, no CUDA headers), and link against the built library."""
    src = tmp_path / "abi.c"
    src.write_text('#include "stefan_b200.h"\n'
                   "int main(void) {\n"
                   "  stefan_ipdg_params p; stefan_blocks_dims d; stefan_peer_sync s; stefan_cg_params c;\n"
                   "  (void)p; (void)d; (void)s; (void)c;\n"
                   "  return stefan_version() > 0 && stefan_last_error() != 0 ? 0 : 1;\n"
                   "}\n")
    exe = tmp_path / "abi"
    libdir = os.path.join(ROOT, "stefanproblem_b200", "lib")
    subprocess.run(["gcc", "-std=c99", "-Wall", "-Werror", "-pedantic", "-I", os.path.join(ROOT, "include"), str(src),
                    "-o", str(exe), "-L", libdir, "-lstefan_b200", f"-Wl,-rpath,{libdir}"], check=True)
    assert subprocess.run([str(exe)]).returncode == 0


@pytest.mark.parametrize("order,nq", [(1, 2), (2, 5)])
def test_ldg_example_host_error_norm_matches_oracle(order, nq):
    """examples/uncut_ldg_convergence.py's host `mesh_L2_error` (node / point ordering of the
    device layout) against the oracle's restatement of useful_routines.jl:95-133."""
    import importlib.util
    from oracle.basis import LagrangeTensorProductBasis as OB
    from oracle.mesh import CellQuadratures, UniformMesh2D
    from oracle.norms import mesh_L2_error
    from stefanproblem_b200 import cutcelldg as cc
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    spec = importlib.util.spec_from_file_location("uncut_ldg_convergence", os.path.join(root, "examples", "uncut_ldg_convergence.py"))
    m = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(m)
    n, nf = 4, (order + 1) ** 2
    om = UniformMesh2D([0, 0], [1, 1], [n, n], nf)
    f = np.random.default_rng(3).standard_normal((1, n * n * nf))
    ref = mesh_L2_error(f, m.exact_solution, OB(2, order), CellQuadratures(nq), om)[0]
    mesh = cc.UniformMesh([0, 0], [1, 1], [n, n], cc.LagrangeTensorProductBasis(2, order))
    x, y = cc.cell_quadrature_points(mesh, nq)
    got = m.mesh_L2_error(f.reshape(-1, nf), m.exact_solution((x, y)), order, nq, n, 1.0 / n)
    assert abs(got - ref) <= 1e-13 * ref


@pytest.mark.parametrize("order,nq,n", [(1, 2, 9), (2, 5, 5), (3, 4, 4)])
def test_ldg_schur_bicgstab_against_direct_solve(order, nq, n):
    """`ldg_solve.schur_solve` (the device `matrix \\ rhs` of the LDG twin) driven on the CPU with
    the oracle's assembled matrix as the operator: same solution as the sparse direct solve."""
    import scipy.sparse.linalg as spla
    import torch
    from stefanproblem_b200 import cutcelldg as cc
    from stefanproblem_b200.ldg_solve import cell_mass_inverse, schur_solve
    V0 = np.array([np.cos(np.pi / 4), np.sin(np.pi / 4)])
    ex = lambda x, y: np.cos(2 * np.pi * x) * np.sin(2 * np.pi * y)  # noqa: E731
    src = lambda x, y: 8 * np.pi ** 2 * np.cos(2 * np.pi * x) * np.sin(2 * np.pi * y)  # noqa: E731
    nf = (order + 1) ** 2
    om = UniformMesh2D([0, 0], [1, 1], [n, n], nf)
    A = fast.ldg_matrix(om, order, nq, 1.0, 1.0, 0.0, 0.0, 1.0, V0).tocsr()
    b = fast.ldg_rhs(om, order, nq, 0.0, 1.0, V0, fast.sample_cell_quadrature(om, nq, src),
                     fast.sample_face_quadrature(om, nq, ex))
    minv = torch.from_numpy(cell_mass_inverse(cc.reference_tables(order, nq)["mass"], 0.25 / (n * n)))
    # the q-q block of the reference's system is exactly the cell mass matrix the elimination assumes
    qq = A[1:3 * nf:3, 1:3 * nf:3].toarray()
    assert np.allclose(qq @ minv.numpy(), np.eye(nf), atol=1e-12)
    sol, info = schur_solve(lambda x: torch.from_numpy(A @ x.numpy()), minv, torch.from_numpy(b), nf, tol=1e-13)
    ref = spla.spsolve(A.tocsc(), b)
    assert info["converged"] and info["relative_residual"] < 1e-12
    assert np.linalg.norm(sol.numpy() - ref) <= 1e-10 * np.linalg.norm(ref)


def test_production_kernels_carry_tma_and_mbarrier_instructions(built_lib):
    """Static check of the built sm_100a library: every march kernel holds TMA bulk copies (UBLKCP) and mbarrier
    operations (SYNCS), the p = 3 kernels TMA tensor loads (UTMALDG), and the vectorised FD kernel 256-bit accesses --
    i.e. the hand-written path is what got compiled (scripts/sass_evidence.py prints the full table)."""
    import shutil
    if shutil.which("cuobjdump") is None:
        pytest.skip("cuobjdump not on PATH")
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    out = subprocess.run([sys.executable, os.path.join(root, "scripts", "sass_evidence.py")], capture_output=True, text=True,
                         check=True).stdout.splitlines()
    keys = [k.strip() for k in out[0].split("|")]
    rows = {r.split("|")[0].strip(): dict(zip(keys[1:], (int(v) for v in r.split("|")[1:]))) for r in out[1:]}
    march = [k for k in rows if "apply_march" in k]
    assert len(march) == 12         # IP p = 1..3 x {apply, fused step}, LDG p = 1..3 in two CTA shapes
    for k in march:
        assert rows[k]["UBLKCP"] > 0 and rows[k]["SYNCS"] > 0 and rows[k]["REDUX"] > 0, k
    for k in ("ipdg_apply_march<3, 0>", "ipdg_apply_march<3, 1>", "ldg_apply_march<3, 0>", "ldg_apply_march<3, 1>"):
        assert rows[k]["UTMALDG"] > 0, k
    for k in ("fd_apply_vec<0>", "fd_apply_vec<1>"):
        assert rows[k]["STG.E.ENL2.256"] > 0 and rows[k]["LDG.E.ENL2.256"] > 0, k
