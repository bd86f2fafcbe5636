"""CPU side of the element-block path: the cell-centric face-level restatement
(oracle/blocks.py) fed with the uniform-mesh data generator of the product
(stefanproblem_b200.blocks.uniform_mesh_block_data, pure NumPy) reproduces the oracle's IP
matrix, which is pinned to the reference's CSV.  So the data layout handed to
stefan_blocks_assemble and the cell-centric block formulas are checked without a GPU."""
import numpy as np
import pytest

from oracle import blocks as oblocks
from oracle import fast
from oracle.mesh import UniformMesh2D


@pytest.mark.parametrize("order", [1, 2, 3])
@pytest.mark.parametrize("aligned", [True, False])
def test_cell_centric_blocks_reproduce_the_ip_matrix(order, aligned):
    from stefanproblem_b200.blocks import uniform_mesh_block_data
    nx, ny = 6, 5
    om = UniformMesh2D([0, 0], [1.0, 0.8], [nx, ny], (order + 1) ** 2)
    om.cellsign[:] = fast.circle_phase(om, (0.5, 0.4), 0.3)
    k1, k2, pen, lam = 1.0, 2.0, 1e3 * nx, 1.5
    A = fast.ip_matrix(om, order, order + 1, k1, k2, pen, lam if aligned else 0.0, aligned_interface=aligned)
    d = uniform_mesh_block_data(nx, ny, 1.0 / nx, 0.8 / ny, order + 1, om.cellsign, k1, k2, pen, lam,
                                aligned_interface=aligned)
    M = oblocks.dense_matrix(oblocks.assemble_blocks(order, d), d["slot_nbr"])
    assert abs(M - A.toarray()).max() <= 1e-13 * abs(A).max()
    assert abs(M - M.T).max() <= 1e-12 * abs(M).max()  # the reference asserts symmetry
