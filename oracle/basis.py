"""Basis / quadrature conventions of the un-vendored packages (oracle, test-only).

Restates the subset of `PolynomialBasis` and `ImplicitDomainQuadrature` that the
hot path calls:

* `LagrangeTensorProductBasis(dim, order)`: 1-D Lagrange on `order+1` equispaced
  nodes in [-1, 1]; 2-D basis is `kron(N(xi), N(eta))`, i.e. eta index fastest.
* `basis(p)` -> NF values, `gradient(basis, p)` -> NF x dim.
* `interpolation_points(basis)` -> dim x NF, same ordering.
* `tensor_product_quadrature(dim, n)` -> Gauss-Legendre tensor rule, list of
  `(point, weight)` pairs (order of the points only changes round-off).
"""
import numpy as np


def lagrange_nodes(order):
    return np.linspace(-1.0, 1.0, order + 1)


def lagrange_1d(order, x):
    """Values N_a(x) and derivatives N'_a(x) of the 1-D Lagrange basis."""
    nodes = lagrange_nodes(order)
    n = order + 1
    vals = np.ones(n)
    ders = np.zeros(n)
    for a in range(n):
        for b in range(n):
            if b != a:
                vals[a] *= (x - nodes[b]) / (nodes[a] - nodes[b])
        for c in range(n):
            if c == a:
                continue
            term = 1.0 / (nodes[a] - nodes[c])
            for b in range(n):
                if b != a and b != c:
                    term *= (x - nodes[b]) / (nodes[a] - nodes[b])
            ders[a] += term
    return vals, ders


class LagrangeTensorProductBasis:
    """`PolynomialBasis.LagrangeTensorProductBasis(dim, order)` (dim = 1 or 2)."""

    def __init__(self, dim, order):
        assert dim in (1, 2)
        self.dim = dim
        self.order = order
        self.nf = (order + 1) ** dim

    def __call__(self, p):
        p = np.atleast_1d(np.asarray(p, dtype=float))
        if self.dim == 1:
            return lagrange_1d(self.order, p[0])[0]
        nx, _ = lagrange_1d(self.order, p[0])
        ny, _ = lagrange_1d(self.order, p[1])
        return np.kron(nx, ny)

    def gradient(self, p):
        """NF x dim matrix of reference-coordinate derivatives."""
        p = np.atleast_1d(np.asarray(p, dtype=float))
        if self.dim == 1:
            return lagrange_1d(self.order, p[0])[1].reshape(-1, 1)
        nx, dx = lagrange_1d(self.order, p[0])
        ny, dy = lagrange_1d(self.order, p[1])
        return np.stack([np.kron(dx, ny), np.kron(nx, dy)], axis=1)

    def interpolation_points(self):
        nodes = lagrange_nodes(self.order)
        if self.dim == 1:
            return nodes.reshape(1, -1)
        xi = np.repeat(nodes, self.order + 1)
        eta = np.tile(nodes, self.order + 1)
        return np.stack([xi, eta], axis=0)


def number_of_basis_functions(basis):
    return basis.nf


def gradient(basis, p):
    return basis.gradient(p)


def interpolation_points(basis):
    return basis.interpolation_points()


def gauss_1d(n):
    return np.polynomial.legendre.leggauss(n)


def tensor_product_quadrature(dim, n):
    """List of (point, weight); point is a length-`dim` array."""
    x, w = gauss_1d(n)
    if dim == 1:
        return [(np.array([x[i]]), w[i]) for i in range(n)]
    return [
        (np.array([x[i], x[j]]), w[i] * w[j]) for i in range(n) for j in range(n)
    ]


def required_quadrature_order(polyorder):
    """`StefanDG/useful_routines.jl:76-78`: ceil(0.5 (2p + 1)) = p + 1."""
    return int(np.ceil(0.5 * (2 * polyorder + 1)))


def interpolation_matrix(v, nd):
    """`CutCellDG.interpolation_matrix(v, nd)` = kron(v', I_nd): nd x (nd*NF)."""
    return np.kron(np.asarray(v).reshape(1, -1), np.eye(nd))


def transform_gradient(G, jac):
    """`CutCellDG.transform_gradient(G, jac)` = G ./ jac' (per-axis scaling)."""
    return G / np.asarray(jac, dtype=float).reshape(1, -1)
