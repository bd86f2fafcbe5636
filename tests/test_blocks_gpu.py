"""Element-block path (stefan_blocks_*): GPU assembly + apply against the oracle.

1. uniform two-phase mesh, tensor quadrature: the assembled blocks give the oracle's IP
   matrix (fast.ip_matrix, itself pinned to the reference's CSV) and the block apply equals
   the matrix-free march kernel;
2. arbitrary data (random quadrature points / weights / per-point normals / per-cell
   conductivities and jacobians, empty and boundary slots, zero-weight padding): every
   block equals the restatement built from the reference's face-level primitives
   (oracle/blocks.py).  Tolerance 1e-12 relative (north star)."""
import numpy as np
import pytest

from oracle import blocks as oblocks
from oracle import fast
from oracle.mesh import UniformMesh2D

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("order", [1, 2, 3])
@pytest.mark.parametrize("nx,ny,two_phase", [(5, 4, False), (9, 7, True)])
def test_uniform_mesh_blocks_match_oracle_and_march(cuda, order, nx, ny, two_phase):
    import scipy.sparse as sp
    import torch
    from stefanproblem_b200 import blocks as B
    from stefanproblem_b200 import cutcelldg as cc
    nq = order + 1
    omesh = UniformMesh2D([0, 0], [1.0, 0.8], [nx, ny], (order + 1) ** 2)
    if two_phase:
        omesh.cellsign[:] = fast.circle_phase(omesh, (0.5, 0.4), 0.3)
    k1, k2, pen, lam = 1.0, 2.0, 1e3 * nx, 1.5
    A = fast.ip_matrix(omesh, order, nq, k1, k2, pen, lam)
    data = B.uniform_mesh_block_data(nx, ny, 1.0 / nx, 0.8 / ny, nq, omesh.cellsign, k1, k2, pen, lam)
    sys_ = B.BlockSystem(nx * ny, order, nq * nq, 4, nq).assemble(data)
    r, c, v = (t.cpu().numpy() for t in sys_.to_coo())
    M = sp.coo_matrix((v, (r, c)), shape=A.shape).tocsr()
    assert abs(M - A).max() <= 1e-12 * abs(A).max()
    x = torch.from_numpy(np.random.default_rng(0).uniform(-1, 1, A.shape[0])).cuda()
    basis = cc.LagrangeTensorProductBasis(2, order)
    mesh = cc.UniformMesh([0, 0], [1.0, 0.8], [nx, ny], basis, phase=omesh.cellsign)
    op = cc.IPOperator(mesh, order, nq, k1, k2, pen, lam, 0.5, "current", 255)
    y_blocks, y_march = sys_ @ x, op @ x
    yr = A @ x.cpu().numpy()
    assert np.linalg.norm(y_blocks.cpu().numpy() - yr) <= 1e-12 * np.linalg.norm(yr)
    assert float((y_blocks - y_march).norm() / y_march.norm()) <= 1e-12


@pytest.mark.parametrize("order", [1, 2, 3])
def test_arbitrary_quadrature_blocks_match_face_level_oracle(cuda, order):
    import torch
    from stefanproblem_b200 import blocks as B
    rng = np.random.default_rng(10 + order)
    ncells, nqv, nslots, nqf = 7, 6, 5, 4
    data = {
        "vol_pts": rng.uniform(-1, 1, (ncells, nqv, 2)), "vol_wts": rng.uniform(0.1, 1, (ncells, nqv)),
        "cell_jac": rng.uniform(0.05, 0.2, (ncells, 2)), "cell_k": rng.uniform(0.5, 3, ncells),
        "slot_nbr": rng.integers(-2, ncells, (ncells, nslots)).astype(np.int32),
        "face_pts": rng.uniform(-1, 1, (ncells, nslots, nqf, 2)),
        "face_nbr_pts": rng.uniform(-1, 1, (ncells, nslots, nqf, 2)),
        "face_wts": rng.uniform(0.05, 0.5, (ncells, nslots, nqf)),
        "slot_pen": rng.uniform(10, 100, (ncells, nslots)), "slot_lambda": rng.uniform(0, 2, (ncells, nslots)),
    }
    ang = rng.uniform(0, 2 * np.pi, (ncells, nslots, nqf))
    data["face_normals"] = np.stack([np.cos(ang), np.sin(ang)], axis=-1)  # curved-interface-like normals
    data["vol_wts"][:, -1] = 0.0          # padding
    data["face_wts"][:, 2, -1] = 0.0
    data["slot_nbr"][0, :3] = (-2, -1, 3)  # every kind of slot occurs
    ref = oblocks.assemble_blocks(order, data)
    sys_ = B.BlockSystem(ncells, order, nqv, nslots, nqf).assemble(data)
    r, c, v = (t.cpu().numpy() for t in sys_.to_coo())
    nf = (order + 1) ** 2
    got = v.reshape(ncells, 1 + nslots, nf, nf)
    live = np.concatenate([np.ones((ncells, 1), bool), data["slot_nbr"] >= 0], axis=1)
    assert np.all(got[~live] == 0.0)
    assert np.abs(got[live] - ref[live]).max() <= 1e-12 * np.abs(ref).max()
    # apply == dense matrix of the oracle's blocks
    A = oblocks.dense_matrix(ref, data["slot_nbr"])
    x = rng.uniform(-1, 1, ncells * nf)
    y = (sys_ @ torch.from_numpy(x).cuda()).cpu().numpy()
    assert np.linalg.norm(y - A @ x) <= 1e-12 * np.linalg.norm(A @ x)
