"""Two-phase 1-D finite-difference Laplace operator: literal restatement (oracle,
test-only).  Follows `/root/reference/StefanFD/finite-difference.jl` (module
`FiniteDifference`).  Node ids are 0-based here (1-based in Julia).

The reference indexes `phaseid[idx +- 3]` lazily inside `&&` chains; Julia would
throw a BoundsError where this restatement raises IndexError (negative indices are
rejected explicitly so NumPy wrap-around cannot hide it).
"""
import numpy as np
import scipy.sparse as sp


class SystemMatrix:
    """finite-difference.jl:4-19"""

    def __init__(self):
        self.rows, self.cols, self.vals = [], [], []


def assemble(matrix, rows, cols, vals):
    """finite-difference.jl:27-32"""
    assert len(rows) == len(cols) == len(vals)
    matrix.rows.extend(rows)
    matrix.cols.extend(cols)
    matrix.vals.extend(vals)


def assemble_central_second_derivative(matrix, nodeid, stepsize):
    """finite-difference.jl:34-39"""
    vals = np.array([1.0, -2.0, 1.0]) / stepsize ** 2
    assemble(matrix, [nodeid] * 3, [nodeid - 1, nodeid, nodeid + 1], vals)


def assemble_backward_second_derivative(matrix, nodeid, stepsize):
    """finite-difference.jl:41-46"""
    vals = np.array([-1.0, 4.0, -5.0, 2.0]) / stepsize ** 2
    assemble(matrix, [nodeid] * 4, [nodeid - 3, nodeid - 2, nodeid - 1, nodeid], vals)


def assemble_forward_second_derivative(matrix, nodeid, stepsize):
    """finite-difference.jl:48-53"""
    vals = np.array([2.0, -5.0, 4.0, -1.0]) / stepsize ** 2
    assemble(matrix, [nodeid] * 4, [nodeid, nodeid + 1, nodeid + 2, nodeid + 3], vals)


def _at(phaseid, idx):
    if idx < 0 or idx >= len(phaseid):
        raise IndexError(f"BoundsError: phaseid[{idx + 1}]")
    return phaseid[idx]


def away_from_interface(phaseid, idx):
    """finite-difference.jl:55-57"""
    return _at(phaseid, idx - 1) == _at(phaseid, idx) == _at(phaseid, idx + 1)


def interface_is_ahead(phaseid, idx):
    """finite-difference.jl:59-63 (short-circuit order preserved)."""
    c = _at(phaseid, idx)
    return _at(phaseid, idx + 1) != c and (
        _at(phaseid, idx - 3) == _at(phaseid, idx - 2)
        and _at(phaseid, idx - 2) == _at(phaseid, idx - 1)
        and _at(phaseid, idx - 1) == c)


def interface_is_behind(phaseid, idx):
    """finite-difference.jl:65-69"""
    c = _at(phaseid, idx)
    return _at(phaseid, idx - 1) != c and (
        _at(phaseid, idx + 3) == _at(phaseid, idx + 2)
        and _at(phaseid, idx + 2) == _at(phaseid, idx + 1)
        and _at(phaseid, idx + 1) == c)


def assemble_laplace_operator(matrix, phaseid, stepsize):
    """finite-difference.jl:71-85"""
    numnodes = len(phaseid)
    for idx in range(1, numnodes - 1):
        if away_from_interface(phaseid, idx):
            assemble_central_second_derivative(matrix, idx, stepsize)
        if phaseid[idx] == 1:
            if interface_is_ahead(phaseid, idx):
                assemble_backward_second_derivative(matrix, idx, stepsize)
            elif interface_is_behind(phaseid, idx):
                assemble_forward_second_derivative(matrix, idx, stepsize)


def assemble_dirichlet_boundary_condition(matrix, numnodes):
    """finite-difference.jl:87-90"""
    assemble(matrix, [0], [0], [1.0])
    assemble(matrix, [numnodes - 1], [numnodes - 1], [1.0])


def phase_id(phi):
    """finite-difference.jl:92-103: phi < 0 -> 1 else 2"""
    phi = np.asarray(phi)
    return np.where(phi < 0, 1, 2).astype(np.int64)


def sparse_operator(sysmatrix, ndofs):
    """finite-difference.jl:105-109"""
    A = sp.coo_matrix((np.asarray(sysmatrix.vals, dtype=float),
                       (np.asarray(sysmatrix.rows, dtype=np.int64),
                        np.asarray(sysmatrix.cols, dtype=np.int64))),
                      shape=(ndofs, ndofs)).tocsr()
    A.sum_duplicates()
    A.eliminate_zeros()
    return A


def assemble_interface_continuity(matrix, r, s, nodeid):
    """finite-difference.jl:111-116"""
    cols = [nodeid - 2, nodeid - 1, nodeid, nodeid + 1, nodeid + 2, nodeid + 3]
    vals = [r, -4.0 * r, (1.0 + 3.0 * r), -(1.0 + 3.0 * s), 4.0 * s, -s]
    assemble(matrix, [nodeid] * 6, cols, vals)


def laplace_matrix(phaseid, stepsize, dirichlet=True):
    """Driver of `StefanFD/test-linear-heat-equation.jl:11-15`."""
    m = SystemMatrix()
    assemble_laplace_operator(m, phaseid, stepsize)
    if dirichlet:
        assemble_dirichlet_boundary_condition(m, len(phaseid))
    return sparse_operator(m, len(phaseid))
