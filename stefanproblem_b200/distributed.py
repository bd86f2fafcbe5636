"""Column-slab partition of a uniform mesh across the GPUs of one node.

Cells are numbered y-fastest, so a range of cell columns is a contiguous range of the
dof vector: rank r owns columns [slab_begin(r), slab_begin(r+1)) and needs, per
operator application, ONE halo cell column from each neighbour (the IP-DG and LDG
face terms couple a cell to its four face neighbours only).  That halo trace exchange
is the only communication of the operator; the interface-flux reduction that drives
the front velocity is a 4-double all-reduce (see `allreduce_interface_sums`).

`torch.distributed` (NCCL on GPUs, gloo in the CPU tests) is the transport; the
exchange runs on a side stream while the interior columns are applied, and the two
edge columns follow once it has landed.
"""
import numpy as np


def slab_begin(ncols, world, rank):
    """First global cell column of `rank` (balanced, deterministic)."""
    return (ncols * rank) // world


def slab_range(ncols, world, rank):
    return slab_begin(ncols, world, rank), slab_begin(ncols, world, rank + 1)


class HaloExchange:
    """Exchange of the first / last cell column of a slab with the left / right neighbour.

    `column_len` = doubles per cell column (ny * NF * dofs-per-node).  Buffers live where
    `like` lives (CUDA tensor -> NCCL, CPU tensor -> gloo)."""

    def __init__(self, column_len, rank, world, like):
        import torch
        self.rank, self.world, self.col = rank, world, int(column_len)
        self.has_lo, self.has_hi = rank > 0, rank < world - 1
        mk = lambda: torch.zeros(self.col, dtype=like.dtype, device=like.device)  # noqa: E731
        self.x_lo = mk() if self.has_lo else None
        self.x_hi = mk() if self.has_hi else None

    def start(self, x):
        """Post the sends / receives for the slab vector `x`; returns the work handles."""
        import torch.distributed as dist
        if self.world == 1:
            return []
        ops = []
        if self.has_lo:
            ops.append(dist.P2POp(dist.isend, x[:self.col], self.rank - 1))
            ops.append(dist.P2POp(dist.irecv, self.x_lo, self.rank - 1))
        if self.has_hi:
            ops.append(dist.P2POp(dist.isend, x[-self.col:], self.rank + 1))
            ops.append(dist.P2POp(dist.irecv, self.x_hi, self.rank + 1))
        return dist.batch_isend_irecv(ops)

    @staticmethod
    def finish(works):
        for w in works:
            w.wait()

    def exchange(self, x):
        self.finish(self.start(x))
        return self.x_lo, self.x_hi


def allreduce_interface_sums(sums):
    """Sum over ranks of the per-rank interface integrals (a handful of doubles):
    the only global dependency of the front-velocity evaluation."""
    import torch.distributed as dist
    if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
        dist.all_reduce(sums, op=dist.ReduceOp.SUM)
    return sums


class DistributedIPOperator:
    """IP-DG operator on a global nxg x ny mesh, one column slab per rank.

    phase_of_column(i) -> int8[ny] gives the phases of global cell column i."""

    def __init__(self, nxg, ny, widths, order, numqp, k1, k2, penalty, lam, Tm, phase_of_column,
                 rank, world, variant="current", terms=255):
        import torch
        from . import cutcelldg as cc
        self.rank, self.world = rank, world
        self.i0, self.i1 = slab_range(nxg, world, rank)
        nxl = self.i1 - self.i0
        hx = widths[0] / nxg
        basis = cc.LagrangeTensorProductBasis(2, order)
        phase = np.concatenate([np.asarray(phase_of_column(i), dtype=np.int8) for i in range(self.i0, self.i1)])
        self.mesh = cc.UniformMesh([self.i0 * hx, 0.0], [nxl * hx, widths[1]], [nxl, ny], basis, phase=phase)
        self.mesh.h = np.array([hx, widths[1] / ny])
        plo = phase_of_column(self.i0 - 1) if rank > 0 else None
        phi = phase_of_column(self.i1) if rank < world - 1 else None
        self.op = cc.IPOperator(self.mesh, order, numqp, k1, k2, penalty, lam, Tm, variant, terms,
                                phase_lo=plo, phase_hi=phi)
        self.ndofs = self.op.ndofs
        self.nxl = nxl
        like = torch.empty(0, dtype=torch.float64, device="cuda")
        self.halo = HaloExchange(ny * basis.nf, rank, world, like)
        self.comm_stream = torch.cuda.Stream() if world > 1 else None
        self._ready = torch.cuda.Event() if world > 1 else None
        self._landed = torch.cuda.Event() if world > 1 else None

    def apply(self, x, out, overlap=True):
        """out = A x on this rank's slab (x is the slab's dof vector)."""
        import torch
        if self.world == 1:
            return self.op.apply(x, out=out)
        h = self.halo
        if not overlap or self.nxl < 3:
            h.exchange(x)
            return self.op.apply(x, out=out, x_lo=h.x_lo, x_hi=h.x_hi)
        main = torch.cuda.current_stream()
        self._ready.record(main)  # x is final on the main stream
        with torch.cuda.stream(self.comm_stream):
            self.comm_stream.wait_event(self._ready)
            h.finish(h.start(x))
            self._landed.record(self.comm_stream)
        self.op.apply_columns(x, out, 1, self.nxl - 1)  # interior: never reads the halos
        main.wait_event(self._landed)
        self.op.apply_columns(x, out, 0, 1, x_lo=h.x_lo, x_hi=h.x_hi)
        self.op.apply_columns(x, out, self.nxl - 1, self.nxl, x_lo=h.x_lo, x_hi=h.x_hi)
        return out
