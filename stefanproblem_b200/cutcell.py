"""Cut-cell data for the element-block path: a uniform Cartesian mesh cut by a CIRCLE.

In the reference the geometry of cut cells comes from three un-vendored packages
(`CutCellDG.CutMesh` / `CellQuadratures` / `FaceQuadratures` / `InterfaceQuadratures` over
`ImplicitDomainQuadrature`, level set = Q2 interpolant of `circle_distance_function`,
StefanDG/useful_routines.jl:40-44).  This module is an independent host-side generator of the
SAME KIND of data for the analytic circle, in the layout of `stefan_blocks_data`
(include/stefan_b200.h), so that the reference's cylinder studies
(InteriorPenalty/cylindrical_bc_tests/IP_test_cylindrical_bc_convergence.jl,
robin_cylindrical_bc_tests/IP_test_cylndrical_robin_bc_convergence.jl) run end to end on the
block path.  Parity with the reference's cut-cell CSVs is at DISCRETISATION level (same scheme,
different geometry approximation and quadrature rules), never digit for digit.

Conventions (as in the reference): level set phi = R - |x - c| (> 0 inside); phase +1 (k1) is
phi > 0, phase -1 (k2) is phi < 0; a cut cell carries two solution cells, (cell, +1) and
(cell, -1), coupled through an interface slot whose normal points out of the + side.

Quadrature ("height function" rules, exact geometry): a line parallel to an axis meets the
disk in ONE chord, so the part of a cell inside the disk is a union of strips
{t in [ta, tb], s in [a(t), b(t)]} with smooth a, b once [ta, tb] is split at the kinks (where a
chord end passes a cell corner or the chord appears); tensor Gauss rules on the strips.  The
outer variable is the one along which the interface is a graph of bounded slope.

Slots of a solution cell: 0 bottom, 1 right, 2 top, 3 left, 4 interface; further slots (5 ...)
hold the faces of tiny neighbouring pieces that were merged into the cell (see `merge`).
"""
import numpy as np

SLOT_STEP = [(0, -1), (1, 0), (0, 1), (-1, 0)]     # neighbour cell of slots 0..3
SLOT_NORMAL = [(0.0, -1.0), (1.0, 0.0), (0.0, 1.0), (-1.0, 0.0)]
OPPOSITE = [2, 3, 0, 1]


def _gauss(n):
    return np.polynomial.legendre.leggauss(n)


def _chord(t, c_t, c_s, R):
    """Chord of the disk on the line {coordinate_t = t}: (s_lo, s_hi) or None."""
    d2 = R * R - (t - c_t) ** 2
    if d2 <= 0.0:
        return None
    s = np.sqrt(d2)
    return c_s - s, c_s + s


def _split_interval(lo, hi, a, b, tol):
    """[lo, hi] cut by the chord [a, b] -> (inside intervals, outside intervals)."""
    if a is None:
        return [], [(lo, hi)]
    ia, ib = max(lo, a), min(hi, b)
    if ib - ia <= tol:
        return [], [(lo, hi)]
    ins = [(ia, ib)]
    out = []
    if ia - lo > tol:
        out.append((lo, ia))
    if hi - ib > tol:
        out.append((ib, hi))
    return ins, out


def disk_rect_rules(x0, x1, y0, y1, cx, cy, R, nq):
    """Quadrature of the parts of the rectangle inside / outside the disk.
    Returns ((pts_in [n,2], w_in [n]), (pts_out, w_out)) in physical coordinates."""
    xc, yc = 0.5 * (x0 + x1), 0.5 * (y0 + y1)
    swap = abs(xc - cx) < abs(yc - cy)       # interface is a graph over x: outer variable x
    if swap:
        x0, x1, y0, y1, cx, cy = y0, y1, x0, x1, cy, cx
    # outer variable t = y in [y0, y1], inner s = x in [x0, x1]
    h = max(x1 - x0, y1 - y0)
    tol = 1e-13 * h
    cuts = [y0, y1]
    for xe in (x0, x1):
        d2 = R * R - (xe - cx) ** 2
        if d2 > 0:
            cuts += [cy - np.sqrt(d2), cy + np.sqrt(d2)]
    cuts += [cy - R, cy + R]
    cuts = sorted(set(c for c in cuts if y0 <= c <= y1))
    g, w = _gauss(nq)
    pin, win, pout, wout = [], [], [], []
    for ta, tb in zip(cuts[:-1], cuts[1:]):
        if tb - ta <= tol:
            continue
        tq = 0.5 * (ta + tb) + 0.5 * (tb - ta) * g
        twq = 0.5 * (tb - ta) * w
        for t, tw in zip(tq, twq):
            ch = _chord(t, cy, cx, R)
            ins, outs = _split_interval(x0, x1, ch[0] if ch else None, ch[1] if ch else None, tol)
            for (lst_p, lst_w, ivs) in ((pin, win, ins), (pout, wout, outs)):
                for a, b in ivs:
                    sq = 0.5 * (a + b) + 0.5 * (b - a) * g
                    lst_p.append(np.stack([sq, np.full(nq, t)], axis=1))
                    lst_w.append(0.5 * (b - a) * w * tw)

    def pack(lp, lw):
        if not lp:
            return np.zeros((0, 2)), np.zeros(0)
        p, ww = np.concatenate(lp), np.concatenate(lw)
        if swap:
            p = p[:, ::-1].copy()
        return p, ww
    return pack(pin, win), pack(pout, wout)


def edge_rules(p0, p1, cx, cy, R, nq):
    """Gauss rules on the parts of the segment p0 -> p1 (axis parallel) inside / outside the disk.
    Returns ((pts_in, w_in), (pts_out, w_out)); weights are physical lengths."""
    p0, p1 = np.asarray(p0, float), np.asarray(p1, float)
    horizontal = abs(p1[1] - p0[1]) < abs(p1[0] - p0[0])
    L = np.abs(p1 - p0).max()
    tol = 1e-13 * L
    if horizontal:
        lo, hi = sorted((p0[0], p1[0]))
        ch = _chord(p0[1], cy, cx, R)
    else:
        lo, hi = sorted((p0[1], p1[1]))
        ch = _chord(p0[0], cx, cy, R)
    ins, outs = _split_interval(lo, hi, ch[0] if ch else None, ch[1] if ch else None, tol)
    g, w = _gauss(nq)

    def rule(ivs):
        P, W = [], []
        for a, b in ivs:
            s = 0.5 * (a + b) + 0.5 * (b - a) * g
            P.append(np.stack([s, np.full(nq, p0[1])], 1) if horizontal else np.stack([np.full(nq, p0[0]), s], 1))
            W.append(0.5 * (b - a) * w)
        if not P:
            return np.zeros((0, 2)), np.zeros(0)
        return np.concatenate(P), np.concatenate(W)
    return rule(ins), rule(outs)


def arc_rule(x0, x1, y0, y1, cx, cy, R, nq):
    """Gauss rule on the part of the circle inside the rectangle: points, weights (arc length),
    outward (radial) unit normals."""
    h = max(x1 - x0, y1 - y0)
    tol = 1e-12 * h
    ang = []
    for xe in (x0, x1):
        d2 = R * R - (xe - cx) ** 2
        if d2 > 0:
            for sgn in (-1.0, 1.0):
                y = cy + sgn * np.sqrt(d2)
                if y0 - tol <= y <= y1 + tol:
                    ang.append(np.arctan2(y - cy, xe - cx))
    for ye in (y0, y1):
        d2 = R * R - (ye - cy) ** 2
        if d2 > 0:
            for sgn in (-1.0, 1.0):
                x = cx + sgn * np.sqrt(d2)
                if x0 - tol <= x <= x1 + tol:
                    ang.append(np.arctan2(ye - cy, x - cx))
    ang = sorted(set(np.round(ang, 14)))
    arcs = []
    if not ang:
        if x0 < cx - R and cx + R < x1 and y0 < cy - R and cy + R < y1:
            arcs.append((-np.pi, np.pi))
    else:
        for a, b in zip(ang, ang[1:] + [ang[0] + 2 * np.pi]):
            if b - a <= 1e-13:
                continue
            m = 0.5 * (a + b)
            px, py = cx + R * np.cos(m), cy + R * np.sin(m)
            if x0 - tol <= px <= x1 + tol and y0 - tol <= py <= y1 + tol:
                arcs.append((a, b))
    g, w = _gauss(nq)
    P, W, N = [], [], []
    for a, b in arcs:
        th = 0.5 * (a + b) + 0.5 * (b - a) * g
        n = np.stack([np.cos(th), np.sin(th)], 1)
        P.append(np.array([cx, cy]) + R * n)
        W.append(0.5 * (b - a) * w * R)
        N.append(n)
    if not P:
        return np.zeros((0, 2)), np.zeros(0), np.zeros((0, 2))
    return np.concatenate(P), np.concatenate(W), np.concatenate(N)


# ----------------------------------------------------------------------------- straight interface
def _clip_halfplane(poly, n, c, keep_positive):
    """Sutherland-Hodgman: the part of the convex polygon `poly` [m, 2] with n . x - c >= 0 (or <= 0)."""
    sgn = 1.0 if keep_positive else -1.0
    d = sgn * (poly @ n - c)
    out = []
    m = len(poly)
    for i in range(m):
        p, q, dp, dq = poly[i], poly[(i + 1) % m], d[i], d[(i + 1) % m]
        if dp >= 0:
            out.append(p)
        if (dp > 0 and dq < 0) or (dp < 0 and dq > 0):
            t = dp / (dp - dq)
            out.append(p + t * (q - p))
    return np.array(out) if out else np.zeros((0, 2))


def polygon_rule(poly, nq):
    """Quadrature of a convex polygon: fan of triangles from vertex 0, collapsed tensor Gauss rule on each (exact for
    polynomials of total degree 2 nq - 2)."""
    if len(poly) < 3:
        return np.zeros((0, 2)), np.zeros(0)
    g, w = _gauss(nq)
    u, wu = 0.5 * (1 + g), 0.5 * w
    U, V = np.meshgrid(u, u, indexing="ij")
    WU = np.outer(wu, wu)
    P, W = [], []
    A = poly[0]
    for B, C in zip(poly[1:-1], poly[2:]):
        area2 = abs((B[0] - A[0]) * (C[1] - A[1]) - (B[1] - A[1]) * (C[0] - A[0]))
        if area2 <= 0.0:
            continue
        # x = A + u (B - A) + u v (C - B), |dx / d(u, v)| = u * 2 area
        X = A[None, None, :] + U[..., None] * (B - A)[None, None, :] + (U * V)[..., None] * (C - B)[None, None, :]
        P.append(X.reshape(-1, 2))
        W.append((WU * U * area2).ravel())
    if not P:
        return np.zeros((0, 2)), np.zeros(0)
    return np.concatenate(P), np.concatenate(W)


def plane_rect_rules_polygon(x0, x1, y0, y1, n, c, nq):
    """Quadrature of the parts of the rectangle with n . x - c > 0 / < 0 by triangulating the two convex polygons
    (exact for total degree 2 nq - 2; used as the cross-check of `plane_rect_rules`)."""
    rect = np.array([[x0, y0], [x1, y0], [x1, y1], [x0, y1]], float)
    return polygon_rule(_clip_halfplane(rect, n, c, True), nq), polygon_rule(_clip_halfplane(rect, n, c, False), nq)


def plane_rect_rules(x0, x1, y0, y1, n, c, nq):
    """The same two parts by the HEIGHT-FUNCTION construction the reference's quadrature package uses (and
    `disk_rect_rules` above): outer Gauss rule in the variable along which the line is a graph of bounded slope, split
    where the line leaves through the bottom / top of the rectangle; inner Gauss rules below and above the line.  Exact
    for the same polynomials as the polygon rule when nq is large enough -- and for the 2-point rule the reference
    runs its p = 1 local-DG studies with, where NO rule is exact, this construction is what reproduces its numbers
    (plane_interface_tests/LDG_convergence.csv linear columns: 1.0000 at n = 3, 17; the polygon rule: 1.15 ... 1.6)."""
    n = np.asarray(n, float)
    swap = abs(n[0]) > abs(n[1])           # the line is a graph over y: exchange the roles of x and y
    if swap:
        x0, x1, y0, y1 = y0, y1, x0, x1
        n = n[::-1]
    g, w = _gauss(nq)
    tol = 1e-13 * max(x1 - x0, y1 - y0)
    cuts = [x0, x1]
    if n[0] != 0.0:
        for ye in (y0, y1):
            xe = (c - n[1] * ye) / n[0]
            if x0 < xe < x1:
                cuts.append(xe)
    cuts = sorted(set(cuts))
    lists = {True: ([], []), False: ([], [])}
    for a, b in zip(cuts[:-1], cuts[1:]):
        if b - a <= tol:
            continue
        for x, wx in zip(0.5 * (a + b) + 0.5 * (b - a) * g, 0.5 * (b - a) * w):
            yr = (c - n[0] * x) / n[1]
            segs = [(y0, y1)] if (yr <= y0 + tol or yr >= y1 - tol) else [(y0, yr), (yr, y1)]
            for ya, yb in segs:
                plus = n[0] * x + n[1] * 0.5 * (ya + yb) - c > 0
                lists[plus][0].append(np.stack([np.full(nq, x), 0.5 * (ya + yb) + 0.5 * (yb - ya) * g], 1))
                lists[plus][1].append(0.5 * (yb - ya) * w * wx)

    def pack(lp, lw):
        if not lp:
            return np.zeros((0, 2)), np.zeros(0)
        p, ww = np.concatenate(lp), np.concatenate(lw)
        if swap:
            p = p[:, ::-1].copy()
        return p, ww
    return pack(*lists[True]), pack(*lists[False])


def plane_edge_rules(p0, p1, n, c, nq):
    """Gauss rules on the parts of the segment p0 -> p1 on the + / - side of the line n . x = c (weights: lengths)."""
    p0, p1 = np.asarray(p0, float), np.asarray(p1, float)
    d0, d1 = p0 @ n - c, p1 @ n - c
    L = np.linalg.norm(p1 - p0)
    g, w = _gauss(nq)

    def rule(a, b):   # parameters along the segment
        if b - a <= 1e-13:
            return np.zeros((0, 2)), np.zeros(0)
        t = 0.5 * (a + b) + 0.5 * (b - a) * g
        return p0[None, :] + t[:, None] * (p1 - p0)[None, :], 0.5 * (b - a) * w * L
    empty = (np.zeros((0, 2)), np.zeros(0))
    if d0 >= 0 and d1 >= 0:
        return rule(0.0, 1.0), empty
    if d0 <= 0 and d1 <= 0:
        return empty, rule(0.0, 1.0)
    tc = d0 / (d0 - d1)
    return (rule(0.0, tc), rule(tc, 1.0)) if d0 > 0 else (rule(tc, 1.0), rule(0.0, tc))


def plane_segment_rule(x0, x1, y0, y1, n, c, nq):
    """Gauss rule on the part of the line n . x = c inside the rectangle: points, weights (length), unit normals
    pointing OUT of the + side (= -n)."""
    rect = np.array([[x0, y0], [x1, y0], [x1, y1], [x0, y1]], float)
    d = rect @ n - c
    pts = []
    for i in range(4):
        dp, dq = d[i], d[(i + 1) % 4]
        if (dp > 0 and dq < 0) or (dp < 0 and dq > 0):
            pts.append(rect[i] + dp / (dp - dq) * (rect[(i + 1) % 4] - rect[i]))
        elif dp == 0.0:
            pts.append(rect[i])
    if len(pts) < 2:
        return np.zeros((0, 2)), np.zeros(0), np.zeros((0, 2))
    a, b = pts[0], pts[-1]
    for q in pts[1:]:
        if np.linalg.norm(q - a) > np.linalg.norm(b - a):
            b = q
    L = np.linalg.norm(b - a)
    if L <= 1e-13 * max(x1 - x0, y1 - y0):
        return np.zeros((0, 2)), np.zeros(0), np.zeros((0, 2))
    g, w = _gauss(nq)
    t = 0.5 * (1 + g)
    return a[None, :] + t[:, None] * (b - a)[None, :], 0.5 * w * L, np.tile(-np.asarray(n, float), (nq, 1))


class CircleCutMesh:
    """Uniform nx x ny mesh on [x0, x0 + widths] cut by the circle |x - center| = radius.

    Attributes after construction:
      sol_cell [nsol], sol_phase [nsol]   geometric cell / phase of every solution cell
      sol_id[(cell, phase)]               solution-cell id
      data                                dict of NumPy arrays in the stefan_blocks_data layout
      vol_xy [nsol][nqv][2]               physical coordinates of the volume points (source sampling)
      face_xy [nsol][nslots][nqf][2]      physical coordinates of the face points (boundary data)
      merged_into[sol]                    -1, or the solution cell a tiny piece was merged into
    `merge_fraction`: pieces whose area is below this fraction of a cell are merged into the
    same-phase neighbour across their longest face (the role of CutCellDG.MergedMesh!)."""

    def __init__(self, x0, widths, nelements, center, radius, numqp, k1, k2, penalty, lam=0.0,
                 merge_fraction=0.0, face_numqp=None):
        self.x0 = np.asarray(x0, float)
        self.widths = np.asarray(widths, float)
        self.nx, self.ny = int(nelements[0]), int(nelements[1])
        self.h = self.widths / np.array([self.nx, self.ny], float)
        self.center, self.R = np.asarray(center, float), float(radius)
        self.nq = int(numqp)
        self.nqf1 = int(face_numqp or numqp)
        self.k = {1: float(k1), -1: float(k2)}
        self.penalty, self.lam = float(penalty), float(lam)
        self.merge_fraction = float(merge_fraction)
        self._build()

    # ------------------------------------------------------------------ geometry
    # the three rules a subclass provides: the parts of a rectangle / of an axis-parallel segment on the + and on
    # the - side, and the piece of the interface inside a rectangle (normals pointing out of the + side)
    def _region_rules(self, xa, xb, ya, yb):
        return disk_rect_rules(xa, xb, ya, yb, self.center[0], self.center[1], self.R, self.nq)

    def _edge_rules(self, p0, p1):
        return edge_rules(p0, p1, self.center[0], self.center[1], self.R, self.nqf1)

    def _interface_rule(self, xa, xb, ya, yb):
        return arc_rule(xa, xb, ya, yb, self.center[0], self.center[1], self.R, self.nqf1)

    def cell_box(self, c):
        i, j = divmod(c, self.ny)
        xa = self.x0[0] + i * self.h[0]
        ya = self.x0[1] + j * self.h[1]
        return xa, xa + self.h[0], ya, ya + self.h[1]

    def _ref(self, c, pts):
        xa, xb, ya, yb = self.cell_box(c)
        ctr = np.array([0.5 * (xa + xb), 0.5 * (ya + yb)])
        return (np.asarray(pts) - ctr) / (0.5 * self.h)

    def _build(self):
        nx, ny, nq = self.nx, self.ny, self.nq
        ncell = nx * ny
        area = self.h[0] * self.h[1]
        pieces = {}          # (cell, phase) -> dict(vol=(pts, w), faces=[(pts, w)] * 4, area)
        iface = {}           # cell -> (pts, w, normals)
        eps = 1e-12
        for c in range(ncell):
            xa, xb, ya, yb = self.cell_box(c)
            (pi, wi), (po, wo) = self._region_rules(xa, xb, ya, yb)
            ain, aout = wi.sum(), wo.sum()
            corners = [((xa, ya), (xb, ya)), ((xb, ya), (xb, yb)), ((xa, yb), (xb, yb)), ((xa, ya), (xa, yb))]
            fr = [self._edge_rules(p0, p1) for p0, p1 in corners]
            if ain > eps * area:
                pieces[(c, 1)] = dict(vol=(pi, wi), faces=[f[0] for f in fr], area=ain)
            if aout > eps * area:
                pieces[(c, -1)] = dict(vol=(po, wo), faces=[f[1] for f in fr], area=aout)
            if (c, 1) in pieces and (c, -1) in pieces:
                iface[c] = self._interface_rule(xa, xb, ya, yb)
        self.pieces, self.iface = pieces, iface
        self.cell_sign = np.array([0 if ((c, 1) in pieces and (c, -1) in pieces) else (1 if (c, 1) in pieces else -1)
                                   for c in range(ncell)], dtype=np.int8)
        # ---- merging of tiny pieces
        host = {}
        if self.merge_fraction > 0:
            for key, pc in pieces.items():
                if pc["area"] >= self.merge_fraction * area:
                    continue
                c, s = key
                i, j = divmod(c, ny)
                best, best_len = None, 0.0
                for f, (di, dj) in enumerate(SLOT_STEP):
                    ii, jj = i + di, j + dj
                    if not (0 <= ii < nx and 0 <= jj < ny):
                        continue
                    nb = (ii * ny + jj, s)
                    if nb not in pieces or pieces[nb]["area"] < self.merge_fraction * area:
                        continue
                    # (across the longest shared face: of all single-piece alternatives this is the choice that
                    # reproduces the reference's stored results, e.g. plane_interface_tests n = 5)
                    ln = pc["faces"][f][1].sum()
                    if ln > best_len:
                        best, best_len = (nb, f), ln
                if best is not None:
                    host[key] = best
        self.host = host
        # ---- solution cells (pieces that own dofs)
        owners = [k for k in sorted(pieces, key=lambda t: (t[0], -t[1])) if k not in host]
        self.sol_id = {k: n for n, k in enumerate(owners)}
        self.sol_cell = np.array([k[0] for k in owners])
        self.sol_phase = np.array([k[1] for k in owners])
        nsol = len(owners)
        sol_of = lambda key: self.sol_id[host[key][0]] if key in host else self.sol_id[key]  # noqa: E731
        frame = lambda key: host[key][0][0] if key in host else key[0]   # geometric cell whose reference frame is used  # noqa: E731
        # ---- gather per solution cell: volume points, and face groups keyed by neighbour solution cell
        vol = [[] for _ in range(nsol)]
        groups = [dict() for _ in range(nsol)]   # sol -> {(kind, nbr_sol): [(pts_phys, w, normals, nbr_frame_cell)]}
        order_keys = [[] for _ in range(nsol)]
        def add(sol, slot_hint, nbr_sol, pts, w, normals, nbr_frame):
            if len(w) == 0:
                return
            key = (slot_hint, nbr_sol)
            if key not in groups[sol]:
                groups[sol][key] = []
                order_keys[sol].append(key)
            groups[sol][key].append((pts, w, normals, nbr_frame))
        for key, pc in pieces.items():
            c, s = key
            me = sol_of(key)
            vol[me].append(pc["vol"])
            i, j = divmod(c, ny)
            merged = key in host
            for f, (di, dj) in enumerate(SLOT_STEP):
                pts, w = pc["faces"][f]
                if len(w) == 0 or w.sum() <= 1e-13 * self.h.max():
                    continue
                ii, jj = i + di, j + dj
                nrm = np.tile(np.array(SLOT_NORMAL[f]), (len(w), 1))
                if not (0 <= ii < nx and 0 <= jj < ny):
                    add(me, f if not merged else 100 + f, -1, pts, w, nrm, -1)
                    continue
                nbk = (ii * ny + jj, s)
                if nbk not in pieces:
                    continue       # (a sliver below the tolerance on the other side)
                nb_sol = sol_of(nbk)
                if nb_sol == me:
                    continue       # the face between a merged piece and its host is interior to the solution cell
                add(me, f if not merged and nbk not in host else 100 + f, nb_sol, pts, w, nrm, frame(nbk))
            if c in iface:
                pts, w, nrm = iface[c]
                other = sol_of((c, -s))
                if other != me and len(w):
                    add(me, 4 if not merged and (c, -s) not in host else 104, other, pts, w,
                        nrm if s == 1 else -nrm, frame((c, -s)))
        # slots: own faces 0..3 and interface 4 keep their index; everything else is appended
        slot_lists = []
        for sol in range(nsol):
            fixed = {}
            extra = []
            for key in order_keys[sol]:
                hint = key[0]
                if hint < 100 and hint not in fixed:
                    fixed[hint] = key
                else:
                    extra.append(key)
            slots = [fixed.get(sidx) for sidx in range(5)] + extra
            slot_lists.append(slots)
        nslots = max(5, max(len(s) for s in slot_lists))
        nqv = max(sum(len(w) for _, w in v) for v in vol)
        nqf = 1
        for sol in range(nsol):
            for key in slot_lists[sol]:
                if key is not None:
                    nqf = max(nqf, sum(len(g[1]) for g in groups[sol][key]))
        self.nsol, self.nslots, self.nqv, self.nqf = nsol, nslots, nqv, nqf
        d = {
            "vol_pts": np.zeros((nsol, nqv, 2)), "vol_wts": np.zeros((nsol, nqv)),
            "cell_jac": np.tile(0.5 * self.h, (nsol, 1)), "cell_k": np.array([self.k[int(s)] for s in self.sol_phase]),
            "slot_nbr": np.full((nsol, nslots), -2, dtype=np.int32),
            "face_pts": np.zeros((nsol, nslots, nqf, 2)), "face_nbr_pts": np.zeros((nsol, nslots, nqf, 2)),
            "face_wts": np.zeros((nsol, nslots, nqf)), "face_normals": np.zeros((nsol, nslots, nqf, 2)),
            "slot_pen": np.zeros((nsol, nslots)), "slot_lambda": np.zeros((nsol, nslots)),
        }
        self.vol_xy = np.zeros((nsol, nqv, 2))
        self.face_xy = np.zeros((nsol, nslots, nqf, 2))
        self.slot_is_interface = np.zeros((nsol, nslots), dtype=bool)
        detj = 0.25 * area
        for sol in range(nsol):
            c = int(self.sol_cell[sol])
            p = np.concatenate([v[0] for v in vol[sol]])
            w = np.concatenate([v[1] for v in vol[sol]])
            n = len(w)
            d["vol_pts"][sol, :n] = self._ref(c, p)
            d["vol_wts"][sol, :n] = w / detj           # the kernels multiply by detjac = jx jy
            self.vol_xy[sol, :n] = p
            for sidx, key in enumerate(slot_lists[sol]):
                if key is None:
                    continue
                hint, nb_sol = key
                gl = groups[sol][key]
                p = np.concatenate([g[0] for g in gl])
                w = np.concatenate([g[1] for g in gl])
                nr = np.concatenate([g[2] for g in gl])
                n = len(w)
                d["slot_nbr"][sol, sidx] = nb_sol
                d["face_pts"][sol, sidx, :n] = self._ref(c, p)
                if nb_sol >= 0:
                    d["face_nbr_pts"][sol, sidx, :n] = self._ref(int(self.sol_cell[nb_sol]), p)
                d["face_wts"][sol, sidx, :n] = w
                d["face_normals"][sol, sidx, :n] = nr
                d["slot_pen"][sol, sidx] = self.penalty
                is_if = hint in (4, 104)
                self.slot_is_interface[sol, sidx] = is_if
                d["slot_lambda"][sol, sidx] = self.lam if is_if else 0.0
                self.face_xy[sol, sidx, :n] = p
        self.data = d

    # ------------------------------------------------------------------ sampling helpers
    def sample_source(self, f1, f2):
        """[nsol][nqv] values of the two-phase source (f1 on +1 cells, f2 on -1 cells; callables of (x, y))."""
        x, y = self.vol_xy[..., 0], self.vol_xy[..., 1]
        plus = (self.sol_phase == 1)[:, None]
        return np.where(plus, np.broadcast_to(f1(x, y), x.shape), np.broadcast_to(f2(x, y), x.shape)) * (self.data["vol_wts"] != 0)

    def sample_boundary(self, g):
        """[nsol][nslots][nqf] boundary data at the points of the boundary slots (0 elsewhere)."""
        x, y = self.face_xy[..., 0], self.face_xy[..., 1]
        bd = (self.data["slot_nbr"] == -1)[..., None]
        return np.where(bd & (self.data["face_wts"] != 0), np.broadcast_to(g(x, y), x.shape), 0.0)

    def sample_exact(self, u):
        x, y = self.vol_xy[..., 0], self.vol_xy[..., 1]
        return np.broadcast_to(u(x, y), x.shape) * (self.data["vol_wts"] != 0)


class PlaneCutMesh(CircleCutMesh):
    """The same data for a STRAIGHT interface: phase +1 where (x - point) . normal > 0 (`plane_distance_function`,
    StefanDG/useful_routines.jl:29-31), as in the reference's plane_interface_tests.  The cut pieces are convex polygons;
    height-function Gauss rules (`plane_rect_rules`) integrate them exactly once numqp is large enough, which is also
    what the reference's quadratures do for a level set that its Q2 interpolant represents exactly."""

    def __init__(self, x0, widths, nelements, normal, point, numqp, k1, k2, penalty, lam=0.0, merge_fraction=0.0,
                 face_numqp=None):
        n = np.asarray(normal, float)
        self.n = n / np.linalg.norm(n)
        self.c = float(self.n @ np.asarray(point, float))
        super().__init__(x0, widths, nelements, (0.0, 0.0), 1.0, numqp, k1, k2, penalty, lam, merge_fraction, face_numqp)

    def _region_rules(self, xa, xb, ya, yb):
        return plane_rect_rules(xa, xb, ya, yb, self.n, self.c, self.nq)

    def _edge_rules(self, p0, p1):
        return plane_edge_rules(p0, p1, self.n, self.c, self.nqf1)

    def _interface_rule(self, xa, xb, ya, yb):
        return plane_segment_rule(xa, xb, ya, yb, self.n, self.c, self.nqf1)


# ------------------------------------------------------------------- interpolated level set (the reference's geometry)
def _real_roots_in(coeffs_low_to_high, lo=-1.0, hi=1.0, tol=1e-12):
    """Real roots in (lo, hi) of the polynomial sum c_i t^i (degree <= 4 here)."""
    c = np.array(coeffs_low_to_high, float)
    scale = np.abs(c).max()
    if scale == 0.0:
        return []
    c = c / scale
    while len(c) > 1 and abs(c[-1]) < 1e-14:
        c = c[:-1]
    if len(c) <= 1:
        return []
    r = np.roots(c[::-1])
    out = sorted(float(z.real) for z in r if abs(z.imag) < 1e-10 and lo + tol < z.real < hi - tol)
    dedup = []
    for z in out:
        if not dedup or z - dedup[-1] > tol:
            dedup.append(z)
    return dedup


class LevelSetCutMesh(CircleCutMesh):
    """The same data for an INTERPOLATED level set: phi_h = the continuous piecewise-Q_m nodal interpolant (m =
    `levelset_order`, the reference uses 2) of `distance_function(x, y)` on the mesh of the cells, exactly what
    `CutCellDG.LevelSet(distancefunction, cgmesh, levelsetbasis)` is in the reference's drivers.  Phase +1 where
    phi_h > 0.  In a cell phi_h is a polynomial, so every geometric quantity is in closed form: along a line of the
    height direction (the coordinate with the larger |d phi_h| at the cell centre) phi_h is a polynomial of degree m
    whose roots split the line into + and - pieces; the outer Gauss rule is split where a root enters / leaves through
    the cell's faces or two roots merge.  Interface points: the roots at the outer Gauss points, weights = arc length
    elements, normals = -grad phi_h / |grad phi_h| (out of the + side)."""

    def __init__(self, x0, widths, nelements, distance_function, levelset_order, numqp, k1, k2, penalty, lam=0.0,
                 merge_fraction=0.0, face_numqp=None):
        m = int(levelset_order)
        nx, ny = int(nelements[0]), int(nelements[1])
        xs = np.asarray(x0, float)[0] + np.asarray(widths, float)[0] * np.arange(m * nx + 1) / (m * nx)
        ys = np.asarray(x0, float)[1] + np.asarray(widths, float)[1] * np.arange(m * ny + 1) / (m * ny)
        X, Y = np.meshgrid(xs, ys, indexing="ij")
        self.m = m
        self.phi_nodes = np.asarray(distance_function(X, Y), float)
        nodes = -1.0 + 2.0 * np.arange(m + 1) / m
        self._to_monomial = np.linalg.inv(np.vander(nodes, increasing=True))   # nodal values -> monomial coefficients
        super().__init__(x0, widths, nelements, (0.0, 0.0), 1.0, numqp, k1, k2, penalty, lam, merge_fraction, face_numqp)

    # monomial coefficients P[i][j] of xi^i eta^j of phi_h in cell (i, j)
    def _poly(self, ci, cj):
        m = self.m
        C = self.phi_nodes[m * ci:m * ci + m + 1, m * cj:m * cj + m + 1]
        return self._to_monomial @ C @ self._to_monomial.T

    def _cell_index(self, x, y):
        ci = int(np.clip(np.floor((x - self.x0[0]) / self.h[0]), 0, self.nx - 1))
        cj = int(np.clip(np.floor((y - self.x0[1]) / self.h[1]), 0, self.ny - 1))
        return ci, cj

    @staticmethod
    def _eval(P, t, s):
        m1 = P.shape[0]
        return float((t ** np.arange(m1)) @ P @ (s ** np.arange(m1)))

    def _oriented(self, xa, xb, ya, yb):
        """(P with the height variable second, swapped?, half widths (outer, height), centre (outer, height))"""
        ci, cj = self._cell_index(0.5 * (xa + xb), 0.5 * (ya + yb))
        P = self._poly(ci, cj)
        gx = P[1, 0] * 2.0 / self.h[0]      # d phi / dx and d phi / dy at the cell centre
        gy = P[0, 1] * 2.0 / self.h[1]
        swap = abs(gx) > abs(gy)            # height direction x: exchange the roles
        if swap:
            return P.T.copy(), True, (0.5 * self.h[1], 0.5 * self.h[0]), (0.5 * (ya + yb), 0.5 * (xa + xb))
        return P, False, (0.5 * self.h[0], 0.5 * self.h[1]), (0.5 * (xa + xb), 0.5 * (ya + yb))

    @staticmethod
    def _line_poly(P, t):
        """coefficients (low to high) in s of phi(t, s)"""
        return (t ** np.arange(P.shape[0])) @ P

    def _breakpoints(self, P):
        """outer values in (-1, 1) where the root structure along the height lines changes"""
        m1 = P.shape[0]
        cuts = [-1.0, 1.0]
        for s in (-1.0, 1.0):
            cuts += _real_roots_in(P @ (s ** np.arange(m1)))          # a root passes through the bottom / top face
        if m1 == 3:
            # two roots merge where the discriminant q1^2 - 4 q2 q0 of the quadratic in s vanishes (degree 4 in t)
            q0, q1, q2 = P[:, 0], P[:, 1], P[:, 2]
            disc = np.polynomial.polynomial.polysub(np.polynomial.polynomial.polymul(q1, q1),
                                                    4.0 * np.polynomial.polynomial.polymul(q2, q0))
            cuts += _real_roots_in(disc)
        return sorted(set(cuts))

    def _region_rules(self, xa, xb, ya, yb):
        P, swap, (ht, hs), (ct, cs) = self._oriented(xa, xb, ya, yb)
        g, w = _gauss(self.nq)
        lists = {True: ([], []), False: ([], [])}
        cuts = self._breakpoints(P)
        for a, b in zip(cuts[:-1], cuts[1:]):
            if b - a <= 1e-12:
                continue
            for t, wt in zip(0.5 * (a + b) + 0.5 * (b - a) * g, 0.5 * (b - a) * w):
                q = self._line_poly(P, t)
                ends = [-1.0] + _real_roots_in(q) + [1.0]
                for sa, sb in zip(ends[:-1], ends[1:]):
                    if sb - sa <= 1e-13:
                        continue
                    plus = np.polynomial.polynomial.polyval(0.5 * (sa + sb), q) > 0
                    sq = 0.5 * (sa + sb) + 0.5 * (sb - sa) * g
                    pts = np.stack([np.full(self.nq, ct + ht * t), cs + hs * sq], 1)
                    lists[plus][0].append(pts[:, ::-1] if swap else pts)
                    lists[plus][1].append(0.5 * (sb - sa) * w * wt * ht * hs)

        def pack(lp, lw):
            if not lp:
                return np.zeros((0, 2)), np.zeros(0)
            return np.concatenate(lp), np.concatenate(lw)
        return pack(*lists[True]), pack(*lists[False])

    def _edge_rules(self, p0, p1):
        p0, p1 = np.asarray(p0, float), np.asarray(p1, float)
        mid = 0.5 * (p0 + p1)
        horizontal = abs(p1[1] - p0[1]) < abs(p1[0] - p0[0])
        # a cell that has the edge on its boundary (phi_h is continuous: either neighbour gives the same restriction)
        eps = 1e-9 * self.h
        ci, cj = self._cell_index(mid[0] + (0.0 if horizontal else -eps[0]), mid[1] + (-eps[1] if horizontal else 0.0))
        if (not horizontal and mid[0] - eps[0] < self.x0[0]) or (horizontal and mid[1] - eps[1] < self.x0[1]):
            ci, cj = self._cell_index(mid[0], mid[1])
        P = self._poly(ci, cj)
        m1 = P.shape[0]
        xc = self.x0[0] + (ci + 0.5) * self.h[0]
        yc = self.x0[1] + (cj + 0.5) * self.h[1]
        if horizontal:
            eta = (mid[1] - yc) / (0.5 * self.h[1])
            q = P @ (eta ** np.arange(m1))              # polynomial in xi
            lo, hi = sorted(((p0[0] - xc) / (0.5 * self.h[0]), (p1[0] - xc) / (0.5 * self.h[0])))
            half = 0.5 * self.h[0]
        else:
            xi = (mid[0] - xc) / (0.5 * self.h[0])
            q = (xi ** np.arange(m1)) @ P               # polynomial in eta
            lo, hi = sorted(((p0[1] - yc) / (0.5 * self.h[1]), (p1[1] - yc) / (0.5 * self.h[1])))
            half = 0.5 * self.h[1]
        g, w = _gauss(self.nqf1)
        ends = [lo] + _real_roots_in(q, lo, hi) + [hi]
        lists = {True: ([], []), False: ([], [])}
        for a, b in zip(ends[:-1], ends[1:]):
            if b - a <= 1e-13:
                continue
            plus = np.polynomial.polynomial.polyval(0.5 * (a + b), q) > 0
            r = 0.5 * (a + b) + 0.5 * (b - a) * g
            if horizontal:
                pts = np.stack([xc + half * r, np.full(self.nqf1, mid[1])], 1)
            else:
                pts = np.stack([np.full(self.nqf1, mid[0]), yc + half * r], 1)
            lists[plus][0].append(pts)
            lists[plus][1].append(0.5 * (b - a) * w * half)

        def pack(lp, lw):
            if not lp:
                return np.zeros((0, 2)), np.zeros(0)
            return np.concatenate(lp), np.concatenate(lw)
        return pack(*lists[True]), pack(*lists[False])

    def _interface_rule(self, xa, xb, ya, yb):
        P, swap, (ht, hs), (ct, cs) = self._oriented(xa, xb, ya, yb)
        m1 = P.shape[0]
        g, w = _gauss(self.nqf1)
        pw = np.arange(m1)
        PT, WT, NT = [], [], []
        cuts = self._breakpoints(P)
        for a, b in zip(cuts[:-1], cuts[1:]):
            if b - a <= 1e-12:
                continue
            for t, wt in zip(0.5 * (a + b) + 0.5 * (b - a) * g, 0.5 * (b - a) * w):
                for s in _real_roots_in(self._line_poly(P, t)):
                    tp, sp_ = t ** pw, s ** pw
                    dt = float((pw[1:] * t ** pw[:-1]) @ P[1:, :] @ sp_) / ht      # d phi / d(physical outer)
                    ds = float(tp @ P[:, 1:] @ (pw[1:] * s ** pw[:-1])) / hs      # d phi / d(physical height)
                    if ds == 0.0:
                        continue
                    nrm = np.hypot(dt, ds)
                    pt = np.array([ct + ht * t, cs + hs * s])
                    nv = -np.array([dt, ds]) / nrm
                    PT.append(pt[::-1] if swap else pt)
                    NT.append(nv[::-1] if swap else nv)
                    WT.append(wt * ht * nrm / abs(ds))     # arc length element: |grad phi| / |d phi / d height| d(outer)
        if not PT:
            return np.zeros((0, 2)), np.zeros(0), np.zeros((0, 2))
        return np.array(PT), np.array(WT), np.array(NT)
