"""CPU checks of the cut-cell generator (stefanproblem_b200/cutcell.py: host-side geometry, no GPU) and of the
analytic cylinder solutions (oracle/analytical.py), plus the reference's cylinder study on small meshes with the
ORACLE's block assembly (oracle/blocks.py) -- the same data the GPU block path consumes in tests/test_cutcell_gpu.py."""
import json
import os

import numpy as np
import pytest
import scipy.sparse as sp
import scipy.sparse.linalg as spla

from oracle import blocks as ob
from oracle.analytical import CylindricalSolver, RobinCylindricalSolver
from stefanproblem_b200.cutcell import CircleCutMesh, arc_rule, disk_rect_rules, edge_rules

GOLD = json.load(open(os.path.join(os.path.dirname(__file__), "golden", "reference_csv.json")))


def test_analytical_solution_identities():
    """test_robin_interface_analytical_solution.jl:28-37 (continuity, Robin flux balance) and the outer / interface
    conditions of cylinder-analytical-solution.jl."""
    s = RobinCylindricalSolver(2.0, 1.0, 1.0, 2.0, 1.0, 0.3, 1.0, 1.0, 0.5)
    assert np.isclose(s.core_solution(0.3), s.rim_solution(0.3))               # TR1 ~ TR2 (:31)
    assert np.isclose(s(1.0), 1.0)                                             # Touter = Tw
    flux = 1.0 * s.core_radial_derivative(0.3) - 2.0 * s.rim_radial_derivative(0.3) - 1.0 * (0.5 - s.core_solution(0.3))
    assert abs(flux) < 1e-13                                                   # :36
    c = CylindricalSolver(1.0, 1.0, 2.0, 2.0, 0.5, 1.5, 1.0)
    assert np.isclose(c.core_solution(0.5), c.rim_solution(0.5)) and np.isclose(c(1.5), 1.0)
    assert abs(1.0 * c.radial_derivative(0.5 - 1e-9) - 2.0 * c.radial_derivative(0.5 + 1e-9)) < 1e-8   # flux continuity
    # the PDE itself: -k (T'' + T'/r) = q in both phases (finite differences)
    for r0, q, k in ((0.2, 1.0, 1.0), (1.0, 2.0, 2.0)):
        h = 1e-4
        lap = (c(r0 + h) - 2 * c(r0) + c(r0 - h)) / h ** 2 + (c(r0 + h) - c(r0 - h)) / (2 * h) / r0
        assert abs(-k * lap - q) < 1e-5


@pytest.mark.parametrize("box,circle", [((0.2, 0.5, 0.1, 0.4), (0.0, 0.0, 0.45)), ((0.0, 1.0, 0.0, 1.0), (0.5, 0.5, 0.3)),
                                        ((0.3, 0.4, 0.3, 0.4), (0.0, 0.0, 0.5)), ((-0.2, 0.3, 0.1, 0.9), (0.1, 0.5, 0.35))])
def test_disk_rect_rules_integrate_polynomials(box, circle):
    """inside + outside rules add up to the rectangle rule for polynomials; the inside area matches a brute-force
    estimate; face portions add up to the edge; the arc rule gives the arc length and unit radial normals."""
    x0, x1, y0, y1 = box
    cx, cy, R = circle
    (pi, wi), (po, wo) = disk_rect_rules(x0, x1, y0, y1, cx, cy, R, 5)
    f = lambda p: 1.0 + p[:, 0] ** 2 * p[:, 1] - 0.5 * p[:, 1] ** 3  # noqa: E731
    g, w = np.polynomial.legendre.leggauss(5)
    X, Y = np.meshgrid(0.5 * (x0 + x1) + 0.5 * (x1 - x0) * g, 0.5 * (y0 + y1) + 0.5 * (y1 - y0) * g, indexing="ij")
    exact = (np.outer(w, w) * 0.25 * (x1 - x0) * (y1 - y0) * f(np.stack([X.ravel(), Y.ravel()], 1)).reshape(5, 5)).sum()
    assert abs((wi * f(pi)).sum() + (wo * f(po)).sum() - exact) < 1e-12
    xs, ys = np.meshgrid(np.linspace(x0, x1, 1201), np.linspace(y0, y1, 1201), indexing="ij")
    brute = ((xs - cx) ** 2 + (ys - cy) ** 2 < R * R).mean() * (x1 - x0) * (y1 - y0)
    assert abs(wi.sum() - brute) < 2e-3 * (x1 - x0) * (y1 - y0)
    assert np.all((pi[:, 0] - cx) ** 2 + (pi[:, 1] - cy) ** 2 < R * R + 1e-12)
    assert np.all((po[:, 0] - cx) ** 2 + (po[:, 1] - cy) ** 2 > R * R - 1e-12)
    for p0, p1 in (((x0, y0), (x1, y0)), ((x1, y0), (x1, y1)), ((x0, y1), (x1, y1)), ((x0, y0), (x0, y1))):
        (_, a), (_, b) = edge_rules(p0, p1, cx, cy, R, 3)
        assert abs(a.sum() + b.sum() - max(abs(p1[0] - p0[0]), abs(p1[1] - p0[1]))) < 1e-13
    P, W, N = arc_rule(x0, x1, y0, y1, cx, cy, R, 4)
    if len(W):
        assert np.allclose(np.hypot(P[:, 0] - cx, P[:, 1] - cy), R) and np.allclose(np.hypot(N[:, 0], N[:, 1]), 1.0)
        assert np.all((P[:, 0] >= x0 - 1e-12) & (P[:, 0] <= x1 + 1e-12) & (P[:, 1] >= y0 - 1e-12) & (P[:, 1] <= y1 + 1e-12))


@pytest.mark.parametrize("merge", [0.0, 0.25])
def test_cut_mesh_geometry_is_consistent(merge):
    n, R = 9, 0.5
    m = CircleCutMesh([0, 0], [1, 1], [n, n], [0, 0], R, 3, 1.0, 2.0, 1e3 * n, 1.0, merge_fraction=merge)
    d = m.data
    detj = d["cell_jac"][:, 0] * d["cell_jac"][:, 1]
    area = (d["vol_wts"] * detj[:, None]).sum(axis=1)
    plus = m.sol_phase == 1
    assert abs(area.sum() - 1.0) < 1e-12                                # the pieces tile the square
    assert abs(area[plus].sum() - np.pi * R * R / 4) < 1e-6             # quarter disk (exact boundary; 3-point Gauss of the chord length)
    iface = m.slot_is_interface & (d["slot_nbr"] >= 0)
    assert abs(d["face_wts"][iface & plus[:, None]].sum() - np.pi * R / 2) < 1e-12     # quarter circumference, once per side
    assert abs(d["face_wts"][iface & ~plus[:, None]].sum() - np.pi * R / 2) < 1e-12
    bd = d["slot_nbr"] == -1
    assert abs(d["face_wts"][bd].sum() - 4.0) < 1e-12                   # the boundary of the square
    # every coupled slot has a partner slot on the neighbour with the same weights and opposite normals
    nsol, nslots = d["slot_nbr"].shape
    for c in range(nsol):
        for s in range(nslots):
            nb = int(d["slot_nbr"][c, s])
            if nb < 0:
                continue
            back = [t for t in range(nslots) if int(d["slot_nbr"][nb, t]) == c and
                    np.isclose(d["face_wts"][nb, t].sum(), d["face_wts"][c, s].sum(), rtol=1e-12, atol=1e-15)]
            assert back, (c, s, nb)
            t = back[0]
            k = d["face_wts"][c, s] != 0
            assert np.allclose(d["face_normals"][c, s][k].sum(axis=0), -d["face_normals"][nb, t][d["face_wts"][nb, t] != 0].sum(axis=0))
    if merge:
        assert m.nsol < len(m.pieces) and (area > 0.2 * (1.0 / n) ** 2).all()   # no tiny solution cell is left


def _study(n, order, prm, merge=0.2):
    sol = (RobinCylindricalSolver(prm["q1"], prm["k1"], prm["q2"], prm["k2"], prm["lam"], prm["R"], prm["b"], prm["Tw"], prm["Tm"])
           if prm["lam"] else CylindricalSolver(prm["q1"], prm["k1"], prm["q2"], prm["k2"], prm["R"], prm["b"], prm["Tw"]))
    m = CircleCutMesh([0, 0], [1, 1], [n, n], [0, 0], prm["R"], prm["numqp"][order], prm["k1"], prm["k2"], 1e3 * n, prm["lam"],
                      merge_fraction=merge)
    A = sp.csr_matrix(ob.dense_matrix(ob.assemble_blocks(order, m.data), m.data["slot_nbr"]))
    assert abs(A - A.T).max() <= 1e-12 * abs(A).max()          # `@assert issymmetric(matrix)` of the drivers
    rhs = ob.assemble_rhs(order, m.data, m.sample_source(lambda x, y: prm["q1"] + 0 * x, lambda x, y: prm["q2"] + 0 * x),
                          m.sample_boundary(sol.at), prm["Tm"])
    u = spla.spsolve(A.tocsc(), rhs)
    e, den = ob.l2_error(order, m.data, u, m.sample_exact(sol.at))
    return e / den


PLAIN = dict(k1=1.0, k2=2.0, q1=1.0, q2=2.0, lam=0.0, Tm=0.0, R=0.5, b=1.5, Tw=1.0, numqp={1: 2, 2: 5})
ROBIN = dict(k1=1.0, k2=2.0, q1=2.0, q2=1.0, lam=1.0, Tm=0.5, R=0.7, b=1.5, Tw=1.0, numqp={1: 4, 2: 5})


@pytest.mark.parametrize("name,prm", [("ip_cylinder", PLAIN), ("ip_cylinder_robin", ROBIN)])
def test_cylinder_study_small_meshes_against_the_reference_csv(name, prm):
    """n = 3, 5, 9 of the reference's cut-cell cylinder studies with the oracle's block assembly on the generator's
    data.  Discretisation-level parity: the stored CSVs come from a Q2-interpolated level set and other quadrature
    rules, this generator from the exact circle.  Observed ratio this / CSV over all 20 entries of the two studies
    (n = 3 ... 33, p = 1, 2): 0.73 ... 1.08, within 4 % from n = 9 on except Robin p = 2 (0.91); asserted 0.7 ... 1.12."""
    for row in GOLD[name][:3]:
        n = row[0]
        for order, want in ((1, row[1]), (2, row[2])):
            got = _study(n, order, prm)
            assert 0.7 * want < got < 1.12 * want, (name, n, order, got, want)


def test_distance_functions_restated():
    """StefanDG/useful_routines.jl:26-74 (oracle/geometry.py): signs, closest points, the corner function's branches;
    the cut-cell generator's phases agree with circle_distance_function (> 0 inside = phase +1)."""
    from oracle import geometry as geo
    pts = np.array([[0.1, 0.6, 0.3, 0.9], [0.2, 0.1, 0.4, 0.9]])
    d = geo.circle_distance_function(pts, [0.0, 0.0], 0.5)
    assert np.allclose(d, 0.5 - np.hypot(pts[0], pts[1])) and d[0] > 0 > d[1]
    cp = geo.closest_point_on_arc(pts, [0.0, 0.0], 0.5)
    assert np.allclose(np.hypot(cp[0], cp[1]), 0.5) and np.allclose(cp[0] * pts[1], cp[1] * pts[0])   # same ray
    n = np.array([0.6, 0.8])
    pd = geo.plane_distance_function(pts, n, [0.25, 0.0])
    on = geo.closest_point_on_plane(pts, n, [0.25, 0.0])
    assert np.allclose(geo.plane_distance_function(on, n, [0.25, 0.0]), 0.0) and np.allclose(np.linalg.norm(pts - on, axis=0), np.abs(pd))
    xc = [0.5, 0.5]
    assert np.isclose(geo.corner_distance_function(np.array([0.1, 0.2]), xc), 0.3)                 # inside the corner: min
    assert np.isclose(geo.corner_distance_function(np.array([0.9, 0.8]), xc), -np.hypot(0.4, 0.3))  # beyond: -distance
    assert np.isclose(geo.corner_distance_function(np.array([0.2, 0.9]), xc), -0.4)                 # above: v[2]
    assert np.isclose(geo.corner_distance_function(np.array([0.9, 0.2]), xc), -0.4)                 # right: v[1]
    assert np.allclose(geo.corner_distance_function(np.array([[0.1, 0.9], [0.2, 0.8]]), xc), [0.3, -0.5])
    m = CircleCutMesh([0, 0], [1, 1], [9, 9], [0, 0], 0.5, 2, 1.0, 2.0, 9e3)
    for sol in range(m.nsol):
        k = m.data["vol_wts"][sol] != 0
        dd = geo.circle_distance_function(m.vol_xy[sol][k].T, [0.0, 0.0], 0.5)
        assert np.all(dd * m.sol_phase[sol] > 0)


# ----------------------------------------------------------------- local DG on element blocks (oracle/ldg_blocks.py)
@pytest.mark.parametrize("order", [1, 2, 3])
@pytest.mark.parametrize("interfacepenalty", [None, 0.7])
def test_ldg_block_oracle_equals_the_uniform_ldg_oracle(order, interfacepenalty):
    """Pins oracle/ldg_blocks.py: on uniform-mesh block data its matrix and right-hand side are those of the
    golden-pinned uniform LDG oracle (oracle/fast.ldg_matrix / ldg_rhs), with and without the mesh-aligned interface
    coupling."""
    from oracle import fast, ldg_blocks as olb
    from oracle.mesh import UniformMesh2D
    from stefanproblem_b200.blocks import ldg_slot_penalties, uniform_mesh_block_data
    nx, ny, nq = 4, 3, order + 1
    nf = (order + 1) ** 2
    mesh = UniformMesh2D([0, 0], [1.0, 0.9], [nx, ny], nf)
    ph = fast.circle_phase(mesh, radius=0.33)
    assert set(np.unique(ph)) == {-1, 1}
    mesh.cellsign = ph
    V0 = (np.cos(0.6), np.sin(0.6))
    A = fast.ldg_matrix(mesh, order, nq, 1.0, 2.0, 0.5, 0.2, 1.3, V0, interfacepenalty=interfacepenalty).toarray()
    data = uniform_mesh_block_data(nx, ny, 1.0 / nx, 0.9 / ny, nq, ph, 1.0, 2.0, 0.0, 0.0,
                                   aligned_interface=interfacepenalty is not None)
    nbr = data["slot_nbr"]
    phc = ph.reshape(-1)
    isif = (nbr >= 0) & (phc[np.maximum(nbr, 0)] != phc[:, None])
    data["slot_pen"] = ldg_slot_penalties(data, V0, 0.5, interfacepenalty or 0.0, 0.2, 1.3, isif)
    assert np.array_equal(data["slot_pen"], olb.boundary_penalties(data, V0, 0.2, 1.3, 0.5, interfacepenalty, isif))
    B = olb.dense_matrix(olb.assemble_blocks(order, data, V0), nbr)
    assert np.abs(A - B).max() <= 1e-14 * np.abs(A).max()
    f = fast.sample_cell_quadrature(mesh, nq, lambda x, y: 1.0 + x * y)
    g = fast.sample_face_quadrature(mesh, nq, lambda x, y: np.sin(x) - y)
    b_ref = fast.ldg_rhs(mesh, order, nq, 0.2, 1.3, V0, f, g)
    b = olb.assemble_rhs(order, data, f, np.transpose(g, (1, 0, 2)))
    assert np.abs(b - b_ref).max() <= 1e-14 * np.abs(b_ref).max()


def _ldg_cylinder(n, order, merge=0.2, exact=None, k2=2.0, q=(1.0, 2.0), numqp=None):
    from oracle import ldg_blocks as olb
    from stefanproblem_b200.blocks import ldg_slot_penalties
    V0 = (np.cos(np.pi / 4), np.sin(np.pi / 4))
    sol = CylindricalSolver(q[0], 1.0, q[1], k2, 0.5, 1.5, 1.0)
    m = CircleCutMesh([0, 0], [1, 1], [n, n], [0, 0], 0.5, numqp or order + 3, 1.0, k2, 0.0, 0.0, merge_fraction=merge)
    data = dict(m.data)
    data["slot_pen"] = ldg_slot_penalties(data, V0, 0.0, 0.0, 0.0, 1.0, m.slot_is_interface)
    A = sp.csc_matrix(olb.dense_matrix(olb.assemble_blocks(order, data, V0), data["slot_nbr"]))
    if exact is None:
        T = sol.at

        def grad(x, y):
            r = np.maximum(np.hypot(x, y), 1e-300)
            tr = sol.radial_derivative(r)
            return tr * x / r, tr * y / r
    else:
        T, grad = exact
    f = m.sample_source(lambda x, y: q[0] + 0 * x, lambda x, y: q[1] + 0 * x)
    u = spla.spsolve(A, olb.assemble_rhs(order, data, f, m.sample_boundary(T)))
    out = []
    for comp, ex in ((0, T), (1, lambda x, y: grad(x, y)[0]), (2, lambda x, y: grad(x, y)[1])):
        e, den = ob.l2_error(order, data, u[comp::3], m.sample_exact(ex))
        out.append((e, den))
    return out


def test_ldg_cylinder_study_small_meshes_against_the_reference_csv():
    """n = 3, 5, 9 of cylindrical_bc_tests/LDG_test_cylindrical_bc_convergence.jl with the oracle's local-DG block
    assembly on the generator's data, against the stored LDG_convergence.csv (relative L2 errors of T and of the two
    flux components).  Observed ratio this / CSV: p = 1 0.99 ... 1.01 in all three columns; p = 2 temperature 0.97 / 0.99
    at n = 5 / 9 (0.41 at n = 3); p = 2 flux components 0.14 ... 0.84 -- the exact circle beats the CSV's Q2-interpolated
    interface there."""
    for row in GOLD["ldg_cylinder"][:3]:
        n = row[0]
        for order in (1, 2):
            got = [e / den for e, den in _ldg_cylinder(n, order)]
            want = row[1 + 3 * (order - 1):4 + 3 * (order - 1)]
            if order == 1:
                for gv, wv in zip(got, want):
                    assert 0.95 * wv < gv < 1.05 * wv, (n, order, got, want)
            else:
                assert (0.9 if n >= 5 else 0.3) * want[0] < got[0] < 1.1 * want[0], (n, got, want)
                assert got[1] < 1.05 * want[1] and got[2] < 1.05 * want[2], (n, got, want)


@pytest.mark.parametrize("order", [1, 2])
def test_ldg_cut_cell_patch_test(order):
    """The reference's linear-solution studies (plane_interface_tests/linear_solution_LDG_convergence.csv: errors
    ~1e-16): with k1 = k2 and no source a linear temperature is reproduced to rounding on the CUT mesh once the
    generator's Gauss rules on the circular pieces are accurate enough to be consistent with each other (volume rule
    against face + arc rules; 12 points: 1e-13 ... 1e-15, the study's p + 3 points: 1e-6 ... 1e-8, spectral in between).
    A wrong sign, normal or neighbour frame in any cut-cell term gives O(1)."""
    T = lambda x, y: 1.0 + 2.0 * x - 0.5 * y  # noqa: E731
    grad = lambda x, y: (2.0 + 0 * x, -0.5 + 0 * x)  # noqa: E731
    for e, den in _ldg_cylinder(5, order, exact=(T, grad), k2=1.0, q=(0.0, 0.0), numqp=12):
        assert e <= 1e-10 * den, (e, den)


# ------------------------------------------- manufactured-solution studies: straight and circular interface, IP and LDG
def _manufactured(kind, method, n, order):
    """examples/cut_interface_convergence.py with the ORACLE's block assembly and a sparse direct solve."""
    import importlib.util
    from oracle import ldg_blocks as olb
    from stefanproblem_b200.blocks import ldg_slot_penalties
    spec = importlib.util.spec_from_file_location(
        "cut_ex", os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "examples",
                               "cut_interface_convergence.py"))
    ex = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(ex)
    if method == "ip":
        m = ex.make_mesh(kind, n, order + 3, 1e3 * n)
        A = sp.csc_matrix(ob.dense_matrix(ob.assemble_blocks(order, m.data), m.data["slot_nbr"]))
        assert abs(A - A.T).max() <= 1e-12 * abs(A).max()
        u = spla.spsolve(A, ob.assemble_rhs(order, m.data, m.sample_source(ex.source, ex.source),
                                            m.sample_boundary(ex.exact), 0.0))
        return [ob.l2_error(order, m.data, u, m.sample_exact(ex.exact))[0]]
    V0 = (np.cos(np.pi / 4), np.sin(np.pi / 4))
    m = ex.make_mesh(kind, n, 2 if order == 1 else order + 3, 0.0)
    data = dict(m.data)
    data["slot_pen"] = ldg_slot_penalties(data, V0, 0.0, 0.0, 0.0, 1.0, m.slot_is_interface)
    A = sp.csc_matrix(olb.dense_matrix(olb.assemble_blocks(order, data, V0), data["slot_nbr"]))
    u = spla.spsolve(A, olb.assemble_rhs(order, data, m.sample_source(ex.source, ex.source), m.sample_boundary(ex.exact)))
    eT, nT = ob.l2_error(order, data, u[0::3], m.sample_exact(ex.exact))
    return [eT / nT, ob.l2_error(order, data, u[1::3], m.sample_exact(ex.grad_x))[0],
            ob.l2_error(order, data, u[2::3], m.sample_exact(ex.grad_y))[0]]


# |this / CSV - 1| allowed per (kind, method, order); observed over n = 3 ... 17 in the comment
CUT_STUDY_TOL = {
    ("plane", "ip", 1): 0.006, ("plane", "ip", 2): 0.002,        # 0.9958 ... 1.0018, 0.9996 ... 1.0008
    ("plane", "ldg", 1): 0.003, ("plane", "ldg", 2): 0.001,      # 0.9988 ... 1.0011 (2-point rule), 0.9994 ... 1.0002
    ("curved", "ip", 1): 0.04, ("curved", "ip", 2): 0.015,      # 1.004 ... 1.032, 0.989 ... 1.004
    ("curved", "ldg", 2): 0.005,                                  # 0.9962 ... 1.0008
}


@pytest.mark.parametrize("kind", ["plane", "curved"])
@pytest.mark.parametrize("method", ["ip", "ldg"])
def test_cut_interface_studies_small_meshes_against_the_reference_csvs(kind, method):
    """n = 3, 5, 9 of the reference's plane_interface_tests / curved_interface_tests studies (IP and local DG) with the
    oracle's block assembly on the generator's data against the stored CSVs.  Where the quadrature is exact for the
    polynomial integrands (everything but LDG p = 1, which the reference runs with a 2-point rule) the CSVs are
    reproduced to 3-4 digits -- the straight interface is represented exactly by both geometries, and at these
    resolutions the circle nearly so.  LDG p = 1: the 2-point rules on the cut pieces are construction dependent
    (observed 0.87 ... 2.2 of the CSV); only its order of magnitude is asserted."""
    nc = 1 if method == "ip" else 3
    for row in GOLD[f"{method}_{kind}"][:3]:
        n = row[0]
        for order in (1, 2):
            got = _manufactured(kind, method, n, order)
            want = row[1 + nc * (order - 1):1 + nc * order]
            tol = CUT_STUDY_TOL.get((kind, method, order))
            for gv, wv in zip(got, want):
                if tol is None:
                    assert 0.5 * wv < gv < 3.0 * wv, (kind, method, n, order, got, want)
                else:
                    assert abs(gv / wv - 1.0) <= tol, (kind, method, n, order, got, want)


@pytest.mark.parametrize("method", ["ip", "ldg"])
def test_plane_interface_linear_solution_is_reproduced_to_rounding(method):
    """plane_interface_tests/linear_solution_{IP,LDG}_convergence.csv (errors 1e-16): T = 3 x + 4 y, no source.  The
    polygon rules are exact, so the cut mesh reproduces it to rounding for both discretisations."""
    from oracle import ldg_blocks as olb
    from stefanproblem_b200.blocks import ldg_slot_penalties
    from stefanproblem_b200.cutcell import PlaneCutMesh
    T = lambda x, y: 3.0 * x + 4.0 * y  # noqa: E731
    zero = lambda x, y: 0.0 * x  # noqa: E731
    nrm = (np.cos(np.pi / 3), np.sin(np.pi / 3))
    for order in (1, 2):
        m = PlaneCutMesh([0, 0], [1, 1], [5, 5], nrm, [0.8, 0.0], order + 1 if method == "ldg" and order == 1 else order + 3,
                         1.0, 1.0, 1e3 * 5, 0.0, merge_fraction=0.2)
        if method == "ip":
            A = sp.csc_matrix(ob.dense_matrix(ob.assemble_blocks(order, m.data), m.data["slot_nbr"]))
            u = spla.spsolve(A, ob.assemble_rhs(order, m.data, m.sample_source(zero, zero), m.sample_boundary(T), 0.0))
            e, den = ob.l2_error(order, m.data, u, m.sample_exact(T))
            assert e <= 1e-9 * den, (order, e, den)      # (penalty 5e3: the system's conditioning, not the scheme)
        else:
            V0 = (np.cos(np.pi / 4), np.sin(np.pi / 4))
            data = dict(m.data)
            data["slot_pen"] = ldg_slot_penalties(data, V0, 0.0, 0.0, 0.0, 1.0, m.slot_is_interface)
            A = sp.csc_matrix(olb.dense_matrix(olb.assemble_blocks(order, data, V0), data["slot_nbr"]))
            u = spla.spsolve(A, olb.assemble_rhs(order, data, m.sample_source(zero, zero), m.sample_boundary(T)))
            for comp, ex in ((0, T), (1, lambda x, y: 3.0 + 0 * x), (2, lambda x, y: 4.0 + 0 * x)):
                e, den = ob.l2_error(order, data, u[comp::3], m.sample_exact(ex))
                assert e <= 1e-11 * den, (order, comp, e, den)


def test_polygon_rules_are_exact():
    from stefanproblem_b200.cutcell import (plane_edge_rules, plane_rect_rules, plane_rect_rules_polygon,
                                           plane_segment_rule, polygon_rule)
    poly = np.array([[0, 0], [1, 0], [1.2, 0.7], [0.4, 1.1], [-0.2, 0.5]], float)
    P, W = polygon_rule(poly, 4)
    P2, W2 = polygon_rule(poly, 12)
    x, y = poly[:, 0], poly[:, 1]
    assert abs(W.sum() - 0.5 * abs(np.dot(x, np.roll(y, -1)) - np.dot(y, np.roll(x, -1)))) < 1e-14
    f = lambda p: p[:, 0] ** 3 * p[:, 1] ** 3 + p[:, 0] ** 6  # noqa: E731   (total degree 6 = 2 nq - 2)
    assert abs((W * f(P)).sum() - (W2 * f(P2)).sum()) < 1e-13
    n, c = np.array([np.cos(1.0), np.sin(1.0)]), 0.45
    (pp, wp), (pm, wm) = plane_rect_rules(0.1, 0.6, 0.2, 0.9, n, c, 3)
    assert abs(wp.sum() + wm.sum() - 0.5 * 0.7) < 1e-14 and (pp @ n - c > 0).all() and (pm @ n - c < 0).all()
    # the height-function rule against the triangulated polygons: same integrals of a polynomial both integrate exactly
    (qp, vp), (qm, vm) = plane_rect_rules_polygon(0.1, 0.6, 0.2, 0.9, n, c, 4)
    mono = lambda p: p[:, 0] ** 2 * p[:, 1] ** 2 + p[:, 0] ** 3  # noqa: E731
    assert abs((wp * mono(pp)).sum() - (vp * mono(qp)).sum()) < 1e-14 and abs((wm * mono(pm)).sum() - (vm * mono(qm)).sum()) < 1e-14
    (ep, ewp), (em, ewm) = plane_edge_rules((0.1, 0.2), (0.6, 0.2), n, c, 3)
    assert abs(ewp.sum() + ewm.sum() - 0.5) < 1e-14
    sp_, sw, sn = plane_segment_rule(0.1, 0.6, 0.2, 0.9, n, c, 3)
    assert np.abs(sp_ @ n - c).max() < 1e-14 and np.allclose(sn, -n)
    # divergence theorem on the + piece: integral of div(x, y) = 2 |piece| = boundary integral of (x, y) . normal
    faces = [((0.1, 0.2), (0.6, 0.2), (0, -1)), ((0.6, 0.2), (0.6, 0.9), (1, 0)), ((0.1, 0.9), (0.6, 0.9), (0, 1)),
             ((0.1, 0.2), (0.1, 0.9), (-1, 0))]
    flux = (sw * (sp_ @ (-n))).sum()
    for p0, p1, nn in faces:
        (q, wq), _ = plane_edge_rules(p0, p1, n, c, 3)
        flux += (wq * (q @ np.array(nn, float))).sum()
    assert abs(flux - 2 * wp.sum()) < 1e-13
