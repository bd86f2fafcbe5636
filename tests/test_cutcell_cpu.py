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
from stefanproblem_b200.cutcell import CircleCutMesh, LevelSetCutMesh, arc_rule, disk_rect_rules, edge_rules

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
@pytest.mark.parametrize("generator", ["circle", "levelset2", "levelset1"])
def test_cut_mesh_geometry_is_consistent(merge, generator):
    n, R = 9, 0.5
    if generator == "circle":
        m = CircleCutMesh([0, 0], [1, 1], [n, n], [0, 0], R, 3, 1.0, 2.0, 1e3 * n, 1.0, merge_fraction=merge)
        tol_area, tol_len = 1e-6, 1e-12    # exact boundary; 3-point Gauss of the chord length
    else:
        order = 2 if generator == "levelset2" else 1
        m = LevelSetCutMesh([0, 0], [1, 1], [n, n], lambda x, y: R - np.hypot(x, y), order, 3, 1.0, 2.0, 1e3 * n, 1.0,
                            merge_fraction=merge)
        tol_area, tol_len = (1e-6, 1e-5) if order == 2 else (2e-3, 4e-3)   # geometry of the Q2 / Q1 interpolant
    d = m.data
    detj = d["cell_jac"][:, 0] * d["cell_jac"][:, 1]
    area = (d["vol_wts"] * detj[:, None]).sum(axis=1)
    plus = m.sol_phase == 1
    assert abs(area.sum() - 1.0) < 1e-12                                # the pieces tile the square
    assert abs(area[plus].sum() - np.pi * R * R / 4) < tol_area         # quarter disk
    iface = m.slot_is_interface & (d["slot_nbr"] >= 0)
    assert abs(d["face_wts"][iface & plus[:, None]].sum() - np.pi * R / 2) < tol_len     # quarter circumference, once per side
    assert abs(d["face_wts"][iface & ~plus[:, None]].sum() - np.pi * R / 2) < tol_len
    assert abs(d["face_wts"][iface & plus[:, None]].sum() - d["face_wts"][iface & ~plus[:, None]].sum()) < 1e-13
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


def _study(n, order, prm, merge=0.2, geometry="levelset"):
    sol = (RobinCylindricalSolver(prm["q1"], prm["k1"], prm["q2"], prm["k2"], prm["lam"], prm["R"], prm["b"], prm["Tw"], prm["Tm"])
           if prm["lam"] else CylindricalSolver(prm["q1"], prm["k1"], prm["q2"], prm["k2"], prm["R"], prm["b"], prm["Tw"]))
    if geometry == "exact":
        m = CircleCutMesh([0, 0], [1, 1], [n, n], [0, 0], prm["R"], prm["numqp"][order], prm["k1"], prm["k2"], 1e3 * n,
                          prm["lam"], merge_fraction=merge)
    else:
        m = LevelSetCutMesh([0, 0], [1, 1], [n, n], lambda x, y: prm["R"] - np.hypot(x, y), 2, prm["numqp"][order],
                            prm["k1"], prm["k2"], 1e3 * n, prm["lam"], merge_fraction=merge)
    variant = prm.get("variant", "current")
    A = sp.csr_matrix(ob.dense_matrix(ob.assemble_blocks(order, m.data, variant), m.data["slot_nbr"]))
    if variant == "current":
        assert abs(A - A.T).max() <= 1e-12 * abs(A).max()          # `@assert issymmetric(matrix)` of the drivers
    rhs = ob.assemble_rhs(order, m.data, m.sample_source(lambda x, y: prm["q1"] + 0 * x, lambda x, y: prm["q2"] + 0 * x),
                          m.sample_boundary(sol.at), prm["Tm"], variant)
    u = spla.spsolve(A.tocsc(), rhs)
    e, den = ob.l2_error(order, m.data, u, m.sample_exact(sol.at))
    return e / den


# (the stored plain-cylinder file was produced with the legacy boundary form -- SURVEY N4 --, the Robin one with the
# current form)
PLAIN = dict(k1=1.0, k2=2.0, q1=1.0, q2=2.0, lam=0.0, Tm=0.0, R=0.5, b=1.5, Tw=1.0, numqp={1: 2, 2: 5},
             variant="legacy_boundary")
ROBIN = dict(k1=1.0, k2=2.0, q1=2.0, q2=1.0, lam=1.0, Tm=0.5, R=0.7, b=1.5, Tw=1.0, numqp={1: 4, 2: 5})
# |this / CSV - 1| allowed per (study, order); observed over n = 3 ... 17 in the comment
CYL_TOL = {("ip_cylinder", 1): 0.02, ("ip_cylinder", 2): 1e-5,               # 2.4e-3 ... 1.3e-2 (2-point rule), 1e-8 ... 2e-6
           ("ip_cylinder_robin", 1): 1e-7, ("ip_cylinder_robin", 2): 1e-6}   # 2e-11 ... 2e-9, 8e-10 ... 1.4e-7


@pytest.mark.parametrize("name,prm", [("ip_cylinder", PLAIN), ("ip_cylinder_robin", ROBIN)])
def test_cylinder_study_small_meshes_against_the_reference_csv(name, prm):
    """n = 3, 5, 9 of the reference's cut-cell cylinder studies with the oracle's block assembly on the data of
    `LevelSetCutMesh` -- the Q2 interpolant of circle_distance_function with height-function Gauss rules, i.e. the
    reference's own geometry restated.  Wherever the quadrature is exact for the integrands the stored CSVs come back
    to 6-10 digits; plain p = 1 runs with a 2-point rule that is not exact on cut cells (within 1.3 %).  With the
    analytic circle instead (`geometry="exact"`) every entry is within 0.73 ... 1.08 of the CSV."""
    for row in GOLD[name][:3]:
        n = row[0]
        for order, want in ((1, row[1]), (2, row[2])):
            got = _study(n, order, prm)
            assert abs(got / want - 1.0) <= CYL_TOL[name, order], (name, n, order, got, want)
    got = _study(5, 2, dict(prm, variant="current"), geometry="exact")
    assert 0.7 * GOLD[name][1][2] < got < 1.12 * GOLD[name][1][2]


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
    m = LevelSetCutMesh([0, 0], [1, 1], [n, n], lambda x, y: 0.5 - np.hypot(x, y), 2, numqp or order + 3, 1.0, k2, 0.0, 0.0,
                        merge_fraction=merge)
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
    assembly on `LevelSetCutMesh` data, against the stored LDG_convergence.csv (relative L2 errors of T and of the two
    flux components): all six columns to 1e-6 at n = 3 and n = 9 (observed 1e-9 ... 9e-7; n = 17, 33 likewise), to 2e-3
    at n = 5 (observed 3e-5 ... 1.3e-3; cause not identified: the circle passes exactly through level-set nodes there, but
    neither the sign of those nodal values nor another merge target explains it)."""
    for row in GOLD["ldg_cylinder"][:3]:
        n = row[0]
        for order in (1, 2):
            got = [e / den for e, den in _ldg_cylinder(n, order)]
            want = row[1 + 3 * (order - 1):4 + 3 * (order - 1)]
            for gv, wv in zip(got, want):
                assert abs(gv / wv - 1.0) <= (2e-3 if n == 5 else 2e-6), (n, order, got, want)


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
        m = ex.make_mesh(kind, n, order + 3, 1e3 * n, ex.IP_LEVELSET_ORDER)
        var = ex.IP_VARIANT[kind]
        A = sp.csc_matrix(ob.dense_matrix(ob.assemble_blocks(order, m.data, var), m.data["slot_nbr"]))
        if var == "current":
            assert abs(A - A.T).max() <= 1e-12 * abs(A).max()
        u = spla.spsolve(A, ob.assemble_rhs(order, m.data, m.sample_source(ex.source, ex.source),
                                            m.sample_boundary(ex.exact), 0.0, var))
        return [ob.l2_error(order, m.data, u, m.sample_exact(ex.exact))[0]]
    V0 = (np.cos(np.pi / 4), np.sin(np.pi / 4))
    m = ex.make_mesh(kind, n, ex.LDG_NUMQP[kind][order], 0.0)
    data = dict(m.data)
    data["slot_pen"] = ldg_slot_penalties(data, V0, 0.0, 0.0, 0.0, 1.0, m.slot_is_interface)
    A = sp.csc_matrix(olb.dense_matrix(olb.assemble_blocks(order, data, V0), data["slot_nbr"]))
    u = spla.spsolve(A, olb.assemble_rhs(order, data, m.sample_source(ex.source, ex.source), m.sample_boundary(ex.exact)))
    eT, nT = ob.l2_error(order, data, u[0::3], m.sample_exact(ex.exact))
    return [eT / nT, ob.l2_error(order, data, u[1::3], m.sample_exact(ex.grad_x))[0],
            ob.l2_error(order, data, u[2::3], m.sample_exact(ex.grad_y))[0]]


# |this / CSV - 1| allowed per (kind, method, order) at n = 3, 9 (and 17, 33) / at n = 5 (the one size that is only
# reproduced to 1e-4 ... 5e-3; cause not identified); observed values in the comments.  None: only the order of magnitude is asserted.
CUT_STUDY_TOL = {
    ("plane", "ip", 1): (1e-7, 6e-3), ("plane", "ip", 2): (1e-7, 2e-4),      # 7e-15, 1.3e-8 / 4.7e-3; 3e-13, 6e-9 / 1.1e-4
    ("plane", "ldg", 1): (2e-3, 2e-3), ("plane", "ldg", 2): (1e-7, 5e-4),    # 3e-15, 1.3e-3 (2-point rule) / 1.1e-3; 2e-14, 2e-9 / 2.9e-4
    ("curved", "ip", 1): (5e-5, 5e-6), ("curved", "ip", 2): (5e-5, 5e-6),    # 1.5e-5, 5e-9 / 4e-7; 1.1e-5, 4e-9 / 4e-8 (level set of order 1)
    ("curved", "ldg", 1): (3e-5, 3e-5), ("curved", "ldg", 2): (1e-5, 1e-5),  # 1e-5, 3e-8 / 2.5e-6 (4-point rule); 6e-6, 9e-9 / 1e-6
}


@pytest.mark.parametrize("kind", ["plane", "curved"])
@pytest.mark.parametrize("method", ["ip", "ldg"])
def test_cut_interface_studies_small_meshes_against_the_reference_csvs(kind, method):
    """n = 3, 5, 9 of the reference's plane_interface_tests / curved_interface_tests studies (IP and local DG) with the
    oracle's block assembly on the generators' data (PlaneCutMesh; LevelSetCutMesh = the reference's Q2 level set)
    against the stored CSVs.  Wherever the quadrature is exact for the polynomial integrands the CSVs come back to 8-15
    digits (IP plane with the legacy boundary form it was produced with; LDG p = 2 for both interfaces); LDG p = 1 plane
    runs with a 2-point rule that is not exact on cut pieces and still agrees to 1e-3 because the rule construction is
    the reference's.  The IP curved file is reproduced (1e-5 at n = 3, 4e-9 at n = 9) by the legacy boundary form on a level
    set of ORDER 1 -- the settings it must have been written with; the script's present levelsetorder = 2 gives 1-3 %.
    Likewise the p = 1 columns of the LDG curved file come back (1e-5 ... 1e-9) with the 4-point rule, not with the 2-point
    rule the script now asks for (0.86 ... 1.8 of the CSV)."""
    nc = 1 if method == "ip" else 3
    for row in GOLD[f"{method}_{kind}"][:3]:
        n = row[0]
        for order in (1, 2):
            got = _manufactured(kind, method, n, order)
            want = row[1 + nc * (order - 1):1 + nc * order]
            tol = CUT_STUDY_TOL[kind, method, order]
            for gv, wv in zip(got, want):
                if tol is None:
                    assert 0.5 * wv < gv < 3.0 * wv, (kind, method, n, order, got, want)
                else:
                    assert abs(gv / wv - 1.0) <= tol[1 if n == 5 else 0], (kind, method, n, order, got, want)


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


def test_level_set_generator_on_a_plane_equals_the_plane_generator():
    """LevelSetCutMesh (Q2 interpolant, roots of the cell polynomials) fed plane_distance_function must produce the data
    of PlaneCutMesh (closed-form half-planes): same solution cells, adjacency, volume rules and face weights / normals
    (the points of a face slot may come in the opposite order)."""
    from stefanproblem_b200.cutcell import PlaneCutMesh
    nrm = np.array([np.cos(np.pi / 3), np.sin(np.pi / 3)])
    for n in (3, 5):
        a = PlaneCutMesh([0, 0], [1, 1], [n, n], nrm, [0.8, 0.0], 4, 1.0, 1.0, 1e3 * n, 0.0, merge_fraction=0.2)
        b = LevelSetCutMesh([0, 0], [1, 1], [n, n], lambda x, y: (x - 0.8) * nrm[0] + y * nrm[1], 2, 4, 1.0, 1.0, 1e3 * n,
                            0.0, merge_fraction=0.2)
        assert (a.nsol, a.nslots, a.nqv, a.nqf) == (b.nsol, b.nslots, b.nqv, b.nqf)
        assert np.array_equal(a.data["slot_nbr"], b.data["slot_nbr"])
        for key in ("vol_pts", "vol_wts", "face_normals"):
            assert np.abs(a.data[key] - b.data[key]).max() < 1e-13, key
        assert np.abs(np.sort(a.data["face_wts"], axis=-1) - np.sort(b.data["face_wts"], axis=-1)).max() < 1e-14


def test_level_set_generator_circle_geometry():
    """Area, interface length and the divergence theorem on the + region for the Q2-interpolated circle (the
    interpolation error of the geometry is 3e-8 relative at n = 9)."""
    m = LevelSetCutMesh([0, 0], [1, 1], [9, 9], lambda x, y: 0.5 - np.hypot(x, y), 2, 5, 1.0, 2.0, 9e3, 0.0, merge_fraction=0.2)
    detj = 0.25 * m.h[0] * m.h[1]
    inside = (m.data["vol_wts"][m.sol_phase == 1] * detj).sum()
    assert abs(inside - np.pi * 0.25 / 4) < 1e-7
    assert abs(m.data["face_wts"][m.slot_is_interface].sum() / 2 - np.pi * 0.5 / 2) < 1e-5
    # sum over the + solution cells of the boundary integrals of x . n: interior faces cancel, what is left is
    # 2 |region| = domain boundary part + interface part
    flux = 0.0
    for sol in np.nonzero(m.sol_phase == 1)[0]:
        for s in range(m.nslots):
            if m.data["slot_nbr"][sol, s] < -1:
                continue
            w = m.data["face_wts"][sol, s]
            flux += (w * np.einsum("qd,qd->q", m.face_xy[sol, s], m.data["face_normals"][sol, s])).sum()
    assert abs(flux - 2 * inside) < 1e-7


def test_ldg_interface_flux_driver_with_the_oracle_backend():
    """examples/cylinder_ldg_interface_flux.py (lagrange_levelset_LDG_interface_flux.jl: closed circular interface
    inside the domain, interface penalty 10, tinyratio 0.3) with the CPU restatement: the normal fluxes of both phases
    and the LDG numerical flux at the interface points against the exact -q1 R / 2.  The reference plots these errors on
    a 5 % axis at n = 33; here n = 9 (coarse: mean 1 %) and n = 17 (max below 2 %)."""
    import importlib.util
    spec = importlib.util.spec_from_file_location(
        "fx", os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "examples",
                           "cylinder_ldg_interface_flux.py"))
    fx = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(fx)
    theta, err = fx.main(9, backend="oracle")
    assert len(theta) == 120 and theta.min() < -3.0 and theta.max() > 3.0     # the whole circle
    assert all(e.mean() < 0.02 for e in err.values())
    theta, err = fx.main(17, backend="oracle")
    assert all(e.max() < 0.03 and e.mean() < 0.003 for e in err.values())
