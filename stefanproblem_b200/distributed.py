"""Column-slab partition of a uniform mesh across the GPUs of one node.

Cells are numbered y-fastest, so a range of cell columns is a contiguous range of the
dof vector: rank r owns columns [slab_begin(r), slab_begin(r+1)) and needs, per
operator application, ONE halo cell column from each neighbour (the IP-DG and LDG
face terms couple a cell to its four face neighbours only).  That halo trace exchange
is the only communication of the operator; the interface-flux reduction that drives
the front velocity is a 4-double all-reduce (see `allreduce_interface_sums`).

Two transports:
  * peer memory (`PeerSlab`, the default of `DistributedIPOperator.apply_peer`): every slab
    vector lives in CUDA-IPC memory that the neighbouring ranks map; the apply kernel of a
    rank reads the neighbours' edge columns straight out of their vectors -- TMA bulk copies
    over NVLink issued by the marching warps -- so an application is ONE kernel launch with
    no pack / send / recv / unpack step.  Epoch flags in the same peer memory order the reads.
  * `torch.distributed` send / recv (NCCL on GPUs, gloo in the CPU tests), `HaloExchange`:
    the exchange runs on a side stream while the interior columns are applied, and the two
    edge columns follow once it has landed.
"""
import ctypes

import numpy as np


def slab_begin(ncols, world, rank):
    """First global cell column of `rank` (balanced, deterministic)."""
    return (ncols * rank) // world


def slab_range(ncols, world, rank):
    return slab_begin(ncols, world, rank), slab_begin(ncols, world, rank + 1)


def balanced_bounds(ncols, seconds, widths, fixed_seconds=0.0, min_cols=4):
    """Column bounds [b_0 = 0, ..., b_world = ncols] that equalise the ranks' kernel times, from one measurement:
    rank r took seconds[r] for widths[r] columns; model t_r = fixed + c_r * width_r (the per-column cost c_r differs
    between slabs: exception cells, mixed columns).  Deterministic on every rank (same inputs -> same bounds)."""
    world = len(seconds)
    c = [max(seconds[r] - fixed_seconds, 1e-12) / max(widths[r], 1) for r in range(world)]
    inv = [1.0 / cr for cr in c]
    scale = ncols / sum(inv)
    w = [max(min_cols, int(round(scale * v))) for v in inv]
    # fix the rounding so that the widths add up, taking from / giving to the widest slabs
    while sum(w) != ncols:
        k = max(range(world), key=lambda r: w[r]) if sum(w) > ncols else min(range(world), key=lambda r: w[r] * c[r])
        w[k] += -1 if sum(w) > ncols else 1
    bounds = [0]
    for v in w:
        bounds.append(bounds[-1] + v)
    return bounds


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
                 rank, world, variant="current", terms=255, bounds=None):
        """bounds: optional column bounds [b_0 = 0, ..., b_world = nxg] (e.g. `balanced_bounds`); default: equal
        shares."""
        import torch
        from . import cutcelldg as cc
        self.rank, self.world = rank, world
        self.i0, self.i1 = (bounds[rank], bounds[rank + 1]) if bounds is not None else slab_range(nxg, world, rank)
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


def _wrap_device(ptr, n):
    """torch view of n device doubles at a raw pointer."""
    import torch
    return torch.as_tensor(_DevArray(ptr, n), device="cuda")


def cg_solve_distributed(dop, b, x0=None, tol=1e-12, maxiter=10000, preconditioner="block_jacobi", check_every=10):
    r"""`matrix \ rhs` for the global IP-DG system, every rank holding its slab of b and x
    (stefan_ipdg_cg_solve_slab).  Per iteration one halo exchange of the search direction (NCCL
    send / recv of one cell column each way) and two all-reduces of two doubles; the operator, the
    block-Jacobi preconditioner and the vector updates are rank-local kernels.  Collective."""
    import torch
    import torch.distributed as dist
    h = dop.halo
    col = h.col
    n = dop.ndofs

    def exchange(v_ptr, stream_ptr):
        with torch.cuda.stream(torch.cuda.ExternalStream(stream_ptr)):
            h.exchange(_wrap_device(v_ptr, n))

    def allreduce(dev_ptr, count, stream_ptr):
        with torch.cuda.stream(torch.cuda.ExternalStream(stream_ptr)):
            dist.all_reduce(_wrap_device(dev_ptr, count), op=dist.ReduceOp.SUM)

    if dop.world == 1:
        return dop.op.solve(b, x0=x0, tol=tol, maxiter=maxiter, preconditioner=preconditioner, check_every=check_every)
    assert h.col == col
    return dop.op.solve_slab(b, h.x_lo, h.x_hi, exchange, allreduce, x0=x0, tol=tol, maxiter=maxiter,
                             preconditioner=preconditioner, check_every=check_every)


class DistributedFDOperator:
    """The 1-D finite-difference operator (finite-difference.jl:34-116) on `numnodes` global
    nodes, one contiguous node range per rank; the one-sided interface rows reach three nodes,
    so the halo is the neighbour's three adjacent nodes (`HaloExchange(3, ...)`).  Dirichlet
    identity rows exist on the first node of rank 0 and the last node of the last rank."""

    def __init__(self, phaseid, stepsize, rank, world):
        import torch
        from . import finite_difference as FD
        phaseid = np.ascontiguousarray(phaseid, dtype=np.int8)
        self.rank, self.world = rank, world
        self.i0, self.i1 = slab_range(phaseid.size, world, rank)
        if world > 1 and self.i1 - self.i0 < 3:
            raise ValueError("a rank's node range must hold the 3-node halo of its neighbour")
        plo = phaseid[self.i0 - 3:self.i0] if rank > 0 else None
        phi = phaseid[self.i1:self.i1 + 3] if rank < world - 1 else None
        self.op = FD.FDOperator(phaseid[self.i0:self.i1], stepsize, phase_lo=plo, phase_hi=phi)
        self.n = self.i1 - self.i0
        self.halo = HaloExchange(3, rank, world, torch.empty(0, dtype=torch.float64, device="cuda"))

    def apply(self, u, out=None):
        """out = A u on this rank's nodes."""
        lo, hi = self.halo.exchange(u)
        return self.op.apply(u, out=out, u_lo=lo, u_hi=hi)

    def step(self, u, dt, kappa1, kappa2, out=None):
        """One fused explicit step u + dt * kappa(phase) * (A u) on this rank's nodes."""
        lo, hi = self.halo.exchange(u)
        return self.op.step(u, dt, kappa1, kappa2, out=out, u_lo=lo, u_hi=hi)


class _DevArray:
    """__cuda_array_interface__ view of raw device memory (lets torch wrap it without a copy)."""

    def __init__(self, ptr, n, typestr="<f8"):
        self.__cuda_array_interface__ = {"shape": (int(n),), "typestr": typestr, "data": (int(ptr), False),
                                         "version": 2}


class PeerSlab:
    """Slab vectors in peer-mappable memory plus the flags that order the neighbours' reads.

    Layout of the allocation: nbuf x [n doubles, padded to 128 B][ready flag u64][consumed flag u64].
    `x` is buffer 0 (`bufs[k]` the others: a time loop alternates between two).  Collective over
    the process group (handles are all-gathered).

    Protocol (see stefan_ipdg_apply_peer / stefan_ipdg_step in include/stefan_b200.h): the flags belong to
    the RANK, not to a buffer.  Epoch k of a rank's ready flag says "my kernel of epoch k has started", i.e.
    everything my stream wrote before it is final AND my kernel of epoch k-1 -- the last one that read my
    neighbours' vectors -- is complete.  The fused kernels wait for the neighbours' epoch-k flag before
    they read a halo column and before they store the slab's own edge columns, which is all a two-buffer
    time loop needs.  A caller that overwrites a slab vector by other means (a copy, its own kernel)
    calls `begin_write()` first."""

    FLAG_BYTES = 128

    def __init__(self, n, rank, world, nbuf=1):
        import torch
        import torch.distributed as dist
        from . import _lib
        L = _lib.lib()
        self.n, self.rank, self.world, self.nbuf = int(n), rank, world, int(nbuf)
        self._stride = (self.n * 8 + 127) // 128 * 128
        self._flag_off = self._stride * self.nbuf
        nbytes = self._flag_off + self.FLAG_BYTES
        p = ctypes.c_void_p()
        hbuf = ctypes.create_string_buffer(64)
        _lib.check(L.stefan_peer_alloc(nbytes, ctypes.byref(p), hbuf))
        self._ptr = p.value
        self.bufs = [torch.as_tensor(_DevArray(self._ptr + k * self._stride, self.n), device="cuda")
                     for k in range(self.nbuf)]
        self.x = self.bufs[0]
        self.status = torch.zeros(1, dtype=torch.int32, device="cuda")
        # device-resident epoch (graph-replayable launches, see stefan_peer_sync.epoch_dev)
        self.epoch_dev = torch.zeros(1, dtype=torch.int64, device="cuda")
        handles = [None] * world
        dist.all_gather_object(handles, (bytes(hbuf.raw), self.n, self.nbuf))
        self._peers = {}
        for r in (rank - 1, rank + 1):
            if 0 <= r < world:
                q = ctypes.c_void_p()
                _lib.check(L.stefan_peer_open(handles[r][0], ctypes.byref(q)))
                n_r, nbuf_r = handles[r][1], handles[r][2]
                stride_r = (n_r * 8 + 127) // 128 * 128
                self._peers[r] = (q.value, n_r, stride_r * nbuf_r, stride_r)
        self.epoch = 0

    # addresses -----------------------------------------------------------------
    def my_flag(self, which):
        return self._ptr + self._flag_off + 8 * which

    def peer_flag(self, r, which):
        base, _, off, _ = self._peers[r]
        return base + off + 8 * which

    def peer_column(self, r, col_len, last, buf=0):
        """Address of the neighbour r's first (last=False) / last (last=True) cell column of its buffer `buf`."""
        base, n_r, _, stride_r = self._peers[r]
        return base + buf * stride_r + (8 * (n_r - col_len) if last else 0)

    def sync_block(self, rank, world, timeout_ms=5000, device_epoch=False, chained=False, push=True):
        """The `stefan_peer_sync` of the NEXT epoch for the fused kernels.
        device_epoch: the epoch lives on the device (`epoch_dev`): the launch can be captured in a CUDA graph.
        push: the kernel also posts its ready / done values into the neighbours' mailbox slots and polls its
        own slots (flag slots of a rank: 0 my_ready, 1 my_done, 2 / 3 ready / done posted by the low neighbour,
        4 / 5 by the high neighbour) instead of reading the neighbours' flags over NVLink.
        chained: gate the halo reads on the neighbours' DONE value of the previous epoch (valid when the vector
        read in epoch k is final once the kernel of epoch k - 1 is complete: a static vector, or the two-buffer
        loop of `step_peer_fused`)."""
        from . import _lib
        has_lo, has_hi = rank > 0, rank < world - 1
        kind = 1 if chained else 0
        if push:
            lo_f = self.my_flag(2 + kind) if has_lo else None
            hi_f = self.my_flag(4 + kind) if has_hi else None
        else:
            lo_f = self.peer_flag(rank - 1, kind) if has_lo else None
            hi_f = self.peer_flag(rank + 1, kind) if has_hi else None
        sync = _lib.PeerSync()
        sync.my_ready, sync.my_done = self.my_flag(0), self.my_flag(1)
        sync.lo_ready, sync.hi_ready = lo_f, hi_f
        sync.timeout_ms = int(timeout_ms)
        sync.status = self.status.data_ptr()
        sync.chained = 1 if chained else 0
        if push:
            # my low neighbour sees me as ITS high neighbour (slots 4 / 5), my high neighbour as its low one (2 / 3)
            sync.post_ready[0] = self.peer_flag(rank - 1, 4) if has_lo else None
            sync.post_ready[1] = self.peer_flag(rank + 1, 2) if has_hi else None
            sync.post_done[0] = self.peer_flag(rank - 1, 5) if has_lo else None
            sync.post_done[1] = self.peer_flag(rank + 1, 3) if has_hi else None
        if device_epoch:
            sync.epoch, sync.epoch_dev = 1, self.epoch_dev.data_ptr()
        else:
            self.epoch += 1
            sync.epoch, sync.epoch_dev = self.epoch, None
        return sync

    def to_device_epochs(self):
        """Continue the epoch count on the device (before graph-captured / device_epoch=True launches)."""
        self.epoch_dev.fill_(self.epoch)

    def to_host_epochs(self):
        """Back to host-side epochs after device_epoch=True launches (synchronises the device)."""
        self.epoch = int(self.epoch_dev.item())

    # ordering ------------------------------------------------------------------
    def publish(self, stream=None):
        """My vector is final for the next epoch (call after the kernels that wrote it)."""
        from . import _lib
        from .cutcelldg import _stream_ptr
        self.epoch += 1
        _lib.check(_lib.lib().stefan_peer_signal(ctypes.c_void_p(self.my_flag(0)), self.epoch, _stream_ptr(stream)))

    def wait_neighbours(self, stream=None, timeout_ms=5000):
        """Hold the stream until both neighbours have published the current epoch."""
        from . import _lib
        from .cutcelldg import _stream_ptr
        for r in self._peers:
            _lib.check(_lib.lib().stefan_peer_wait(ctypes.c_void_p(self.peer_flag(r, 0)), self.epoch, timeout_ms,
                                                   ctypes.c_void_p(self.status.data_ptr()), _stream_ptr(stream)))

    def release(self, stream=None):
        """I am done reading my neighbours' vectors of the current epoch."""
        from . import _lib
        from .cutcelldg import _stream_ptr
        _lib.check(_lib.lib().stefan_peer_signal(ctypes.c_void_p(self.my_flag(1)), self.epoch, _stream_ptr(stream)))

    def wait_released(self, stream=None, timeout_ms=5000):
        """Hold the stream until the neighbours are done reading my vector (before overwriting it)."""
        from . import _lib
        from .cutcelldg import _stream_ptr
        for r in self._peers:
            _lib.check(_lib.lib().stefan_peer_wait(ctypes.c_void_p(self.peer_flag(r, 1)), self.epoch, timeout_ms,
                                                   ctypes.c_void_p(self.status.data_ptr()), _stream_ptr(stream)))

    def begin_write(self, stream=None, timeout_ms=5000):
        """Write-after-read guard: call before overwriting a slab vector that the neighbours read in the
        current epoch (a host upload, an update kernel).  Enqueues the wait for their `done` flags; no-op
        before the first epoch."""
        if self.epoch > 0:
            self.wait_released(stream, timeout_ms)

    def check(self, collective=True):
        """Raise on EVERY rank if a flag wait timed out anywhere (stale halos must not go unnoticed).
        Synchronises the device; collective=True all-reduces the status over the process group."""
        import torch
        import torch.distributed as dist
        st = self.status.clone()
        if collective and dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
            dist.all_reduce(st, op=dist.ReduceOp.MAX)
        if int(st.item()) != 0:
            raise RuntimeError("a peer-flag wait timed out (a neighbour did not publish its epoch): "
                               "the halo columns of that application are stale")

    def close(self):
        from . import _lib
        L = _lib.lib()
        for q, _, _, _ in self._peers.values():
            L.stefan_peer_close(ctypes.c_void_p(q))
        self._peers = {}
        if self._ptr:
            self.x = None
            self.bufs = []
            L.stefan_peer_free(ctypes.c_void_p(self._ptr))
            self._ptr = None


def step_peer_fused(dop, slab, src, dst, b, dt, timeout_ms=5000, device_epoch=False, chained=True):
    """slab.bufs[dst] = u + dt Minv (b - A u) with u = slab.bufs[src]: ONE launch (stefan_ipdg_step, peer form).
    The neighbours' edge columns of THEIR buffer `src` are read over NVLink inside the kernel.  chained (default):
    in a loop that alternates two buffers the neighbours' buffer `src` is final once THEIR previous step kernel is
    complete, so the halo gate is their done value of the previous epoch (see PeerSlab.sync_block); the very first
    step of a loop must follow a barrier (the initial vectors are final) -- epoch 0 counts as done."""
    from . import _lib
    from .cutcelldg import _stream_ptr
    col = dop.halo.col
    lo = ctypes.c_void_p(slab.peer_column(dop.rank - 1, col, last=True, buf=src)) if dop.rank > 0 else None
    hi = ctypes.c_void_p(slab.peer_column(dop.rank + 1, col, last=False, buf=src)) if dop.rank < dop.world - 1 else None
    sync = slab.sync_block(dop.rank, dop.world, timeout_ms, device_epoch, chained=chained)
    _lib.check(_lib.lib().stefan_ipdg_step(dop.op._h, ctypes.c_void_p(slab.bufs[src].data_ptr()),
                                           ctypes.c_void_p(slab.bufs[dst].data_ptr()),
                                           ctypes.c_void_p(b.data_ptr()) if b is not None else None, float(dt),
                                           lo, hi, ctypes.byref(sync), _stream_ptr()))
    return slab.bufs[dst]


def interface_sums_peer(dop, slab, buf, out, timeout_ms=5000):
    """Front-velocity sums (stefan_ipdg_interface_sums_peer) of slab.bufs[buf] right after the fused kernel of the
    current epoch wrote it: ONE launch; the threads that read a neighbour's edge column wait (inside the kernel) for
    that neighbour's done value of this epoch in the local mailbox.  `out`: 4 doubles on the device (this rank's
    part; all-reduce it)."""
    from . import _lib
    from .cutcelldg import _stream_ptr
    col = dop.halo.col
    has_lo, has_hi = dop.rank > 0, dop.rank < dop.world - 1
    lo = ctypes.c_void_p(slab.peer_column(dop.rank - 1, col, last=True, buf=buf)) if has_lo else None
    hi = ctypes.c_void_p(slab.peer_column(dop.rank + 1, col, last=False, buf=buf)) if has_hi else None
    sync = _lib.PeerSync()
    sync.lo_ready = slab.my_flag(3) if has_lo else None     # the neighbours' done values, posted into my slots
    sync.hi_ready = slab.my_flag(5) if has_hi else None
    sync.epoch, sync.timeout_ms, sync.status = slab.epoch, int(timeout_ms), slab.status.data_ptr()
    _lib.check(_lib.lib().stefan_ipdg_interface_sums_peer(dop.op._h, ctypes.c_void_p(slab.bufs[buf].data_ptr()), lo, hi,
                                                          ctypes.byref(sync), ctypes.c_void_p(out.data_ptr()),
                                                          _stream_ptr()))
    return out


def apply_peer_fused(dop, slab, out, timeout_ms=5000, device_epoch=False, buf=0, chained=False, push=True):
    """out = A x for the slab vector `slab.bufs[buf]`: ONE kernel launch and nothing else on the stream.
    The kernel publishes the slab's next epoch when it starts, its warps wait for the
    neighbours' epoch right before they fetch the neighbours' edge columns over NVLink, and
    its last CTA releases them (stefan_ipdg_apply_peer).  device_epoch=True keeps the epoch on the
    device so that the launch can be captured in a CUDA graph and replayed.  chained=True: the slab vector
    does not change between the applications of a loop (or only inside such kernels), so the halo gate is the
    neighbours' DONE value of the previous epoch (see PeerSlab.sync_block)."""
    from . import _lib
    from .cutcelldg import _stream_ptr
    col = dop.halo.col
    lo = ctypes.c_void_p(slab.peer_column(dop.rank - 1, col, last=True, buf=buf)) if dop.rank > 0 else None
    hi = ctypes.c_void_p(slab.peer_column(dop.rank + 1, col, last=False, buf=buf)) if dop.rank < dop.world - 1 else None
    sync = slab.sync_block(dop.rank, dop.world, timeout_ms, device_epoch, chained=chained, push=push)
    _lib.check(_lib.lib().stefan_ipdg_apply_peer(dop.op._h, ctypes.c_void_p(slab.bufs[buf].data_ptr()),
                                                 ctypes.c_void_p(out.data_ptr()), lo, hi, ctypes.byref(sync),
                                                 _stream_ptr()))
    return out


def apply_peer(dop, slab, out, publish=True):
    """out = A x for the slab vector `slab.x` of a DistributedIPOperator: ONE kernel launch;
    the neighbours' edge columns are read from their PeerSlab memory over NVLink.

    publish=True: publish my vector, wait for the neighbours', apply, release.  A caller whose
    x does not change between applications (a benchmark) can publish once and pass False."""
    import torch
    from . import _lib
    from .cutcelldg import _stream_ptr
    col = dop.halo.col
    if publish:
        slab.publish()
        slab.wait_neighbours()
    lo = ctypes.c_void_p(slab.peer_column(dop.rank - 1, col, last=True)) if dop.rank > 0 else None
    hi = ctypes.c_void_p(slab.peer_column(dop.rank + 1, col, last=False)) if dop.rank < dop.world - 1 else None
    _lib.check(_lib.lib().stefan_ipdg_apply(dop.op._h, ctypes.c_void_p(slab.x.data_ptr()),
                                            ctypes.c_void_p(out.data_ptr()), lo, hi, 0, _stream_ptr()))
    if publish:
        slab.release()
    return out


class DistributedLDGOperator:
    """Local-DG operator (3 dofs per node) on a global nxg x ny mesh, one column slab per rank.

    Halo = the neighbour's adjacent cell column (3 ny NF doubles).  `apply(x, out)` exchanges
    it with torch.distributed send / recv; `apply_peer(slab, out)` reads it straight out of the
    neighbour's PeerSlab over NVLink (ordering by stefan_peer_signal / stefan_peer_wait)."""

    def __init__(self, nxg, ny, widths, order, numqp, k1, k2, interiorpenalty, negboundarypenalty,
                 posboundarypenalty, V0, phase_of_column, rank, world):
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
        self.op = cc.LDGOperator(self.mesh, order, numqp, k1, k2, interiorpenalty, negboundarypenalty,
                                 posboundarypenalty, V0, phase_lo=plo, phase_hi=phi)
        self.ndofs = self.op.ndofs
        self.nxl = nxl
        like = torch.empty(0, dtype=torch.float64, device="cuda")
        self.halo = HaloExchange(3 * ny * basis.nf, rank, world, like)

    def apply(self, x, out):
        h = self.halo
        h.exchange(x)
        return self.op.apply(x, out=out, x_lo=h.x_lo, x_hi=h.x_hi)

    def apply_peer(self, slab, out):
        from . import _lib
        from .cutcelldg import _stream_ptr
        col = self.halo.col
        slab.publish()
        slab.wait_neighbours()
        lo = ctypes.c_void_p(slab.peer_column(self.rank - 1, col, last=True)) if self.rank > 0 else None
        hi = ctypes.c_void_p(slab.peer_column(self.rank + 1, col, last=False)) if self.rank < self.world - 1 else None
        _lib.check(_lib.lib().stefan_ldg_apply(self.op._h, ctypes.c_void_p(slab.x.data_ptr()),
                                               ctypes.c_void_p(out.data_ptr()), lo, hi, _stream_ptr()))
        slab.release()
        return out
