"""Multi-GPU check of the column-slab partition (run under torchrun, one rank per GPU):
peer-memory halos (one launch, neighbours' columns read over NVLink) and the NCCL
send/recv halos both reproduce the one-GPU operator.  Prints 'PEER CHECK OK' on rank 0.

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 \\
        --master-port 29511 tests/dist_peer_check.py [--time]
"""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from stefanproblem_b200 import cutcelldg as cc  # noqa: E402
from stefanproblem_b200.distributed import (DistributedIPOperator, DistributedLDGOperator, PeerSlab,  # noqa: E402
                                            apply_peer, apply_peer_fused, interface_sums_peer, slab_range,
                                            step_peer_fused)


def column_phase(nxg, ny, i):
    j = np.arange(ny)
    cx, cy = (i + 0.5) / nxg, (j + 0.5) / ny
    return np.where((cx - 0.5) ** 2 + (cy - 0.5) ** 2 < 0.16, -1, 1).astype(np.int8)


def check_fd(rank, world):
    # 1-D finite differences: node ranges with 3-node halos, apply and 20 fused explicit steps (bit-identical)
    from stefanproblem_b200 import finite_difference as FD
    from stefanproblem_b200.distributed import DistributedFDOperator
    n = 4001 * world + 7
    cut = slab_range(n, world, 1)[0]   # an interface two nodes beyond the first cut: one-sided rows cross the halo
    ph = np.where((np.arange(n) >= n // 5) & (np.arange(n) < cut + 2), 1, 2).astype(np.int8)
    hstep = 1.0 / (n - 1)
    ug = torch.from_numpy(np.random.default_rng(8).uniform(-1, 1, n)).cuda()
    fullfd = FD.FDOperator(ph, hstep)
    dfd = DistributedFDOperator(ph, hstep, rank, world)
    sl = slice(dfd.i0, dfd.i1)
    assert torch.equal(dfd.apply(ug[sl].clone()), fullfd.apply(ug)[sl])
    u1, ul = ug.clone(), ug[sl].clone()
    dt = 0.2 * hstep * hstep
    for _ in range(20):
        u1 = fullfd.step(u1, dt, 1.0, 2.0)
        ul = dfd.step(ul, dt, 1.0, 2.0)
    assert torch.equal(ul, u1[sl]), "distributed FD time loop differs from one GPU"


def main():
    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    torch.cuda.set_device(int(os.environ["LOCAL_RANK"]))
    dist.init_process_group("nccl", device_id=torch.device("cuda", int(os.environ["LOCAL_RANK"])))
    if "--fd-only" in sys.argv:
        check_fd(rank, world)
        dist.barrier()
        if rank == 0:
            print(f"FD CHECK OK world={world}", flush=True)
        return
    worst = 0.0
    for order in (1, 2, 3):
        for nxg, ny in ((24 * world + 5, 95), (16 * world, 64)):
            nf = (order + 1) ** 2
            pen = 1e3 * nxg
            dop = DistributedIPOperator(nxg, ny, (1.0, 1.0), order, order + 1, 1.0, 2.0, pen, 1.0, 0.5,
                                        lambda i: column_phase(nxg, ny, i), rank, world)
            # the whole mesh on every rank, plain kernel: the reference
            basis = cc.LagrangeTensorProductBasis(2, order)
            phase = np.concatenate([column_phase(nxg, ny, i) for i in range(nxg)])
            full = cc.IPOperator(cc.UniformMesh([0, 0], [1, 1], [nxg, ny], basis, phase=phase), order, order + 1,
                                 1.0, 2.0, pen, 1.0, 0.5, "current", 255)
            xg = torch.from_numpy(np.random.default_rng(5).uniform(-1, 1, nxg * ny * nf)).cuda()
            yref = full.apply(xg, algo=1)
            i0, i1 = slab_range(nxg, world, rank)
            sl = slice(i0 * ny * nf, i1 * ny * nf)
            # peer-memory halos, two epochs (the second after every rank rewrote its vector)
            slab = PeerSlab(dop.ndofs, rank, world)
            out = torch.empty(dop.ndofs, dtype=torch.float64, device="cuda")
            for epoch in range(4):
                slab.wait_released()
                slab.x.copy_(xg[sl] * (1.0 + epoch))
                if epoch % 2 == 0:
                    apply_peer(dop, slab, out)        # signal / wait / apply / release launches
                else:
                    apply_peer_fused(dop, slab, out)  # everything inside the one apply launch
                err = float((out - yref[sl] * (1.0 + epoch)).norm() / yref[sl].norm())
                worst = max(worst, err)
            torch.cuda.synchronize()
            assert int(slab.status.item()) == 0, "a peer wait timed out"
            # NCCL halos
            out2 = torch.empty_like(out)
            dop.apply(xg[sl].clone(), out2)
            worst = max(worst, float((out2 - yref[sl]).norm() / yref[sl].norm()))
            dist.barrier()
            slab.close()
    # explicit time loop across the slabs: u <- u + dt Minv (b - A u), 30 steps, vector rewritten
    # every step (the epoch protocol in earnest), front-velocity sums all-reduced; against the
    # same loop on one GPU.  North star: <= 1e-10 after N steps.
    from stefanproblem_b200.distributed import allreduce_interface_sums
    for order in (1, 2):
        nxg, ny, nq = 10 * world + 3, 37, order + 1
        nf = (order + 1) ** 2
        pen, lam, Tm, dt, nsteps = 10.0 * nxg, 1.0, 0.5, 2e-6, 30
        dop = DistributedIPOperator(nxg, ny, (1.0, 1.0), order, nq, 1.0, 2.0, pen, lam, Tm,
                                    lambda i: column_phase(nxg, ny, i), rank, world)
        basis = cc.LagrangeTensorProductBasis(2, order)
        phase = np.concatenate([column_phase(nxg, ny, i) for i in range(nxg)])
        full = cc.IPOperator(cc.UniformMesh([0, 0], [1, 1], [nxg, ny], basis, phase=phase), order, nq,
                             1.0, 2.0, pen, lam, Tm, "current", 255)
        rng = np.random.default_rng(8)
        f = rng.uniform(-1, 1, (nxg, ny, nq * nq))
        g = [rng.uniform(-1, 1, (n, nq)) for n in (nxg, ny, nxg, ny)]
        dev = lambda a: torch.from_numpy(np.ascontiguousarray(a)).cuda()  # noqa: E731
        bfull = full.rhs(dev(f), dev(np.concatenate([a.ravel() for a in g])))
        u0 = torch.from_numpy(rng.uniform(-1, 1, nxg * ny * nf)).cuda()
        uf, yf = u0.clone(), torch.empty_like(u0)
        for _ in range(nsteps):
            full.apply(uf, out=yf, algo=1)
            full.euler_update(uf, yf, bfull, dt)
        sums_ref = full.interface_sums(uf)
        i0, i1 = slab_range(nxg, world, rank)
        sl = slice(i0 * ny * nf, i1 * ny * nf)
        bloc = dop.op.rhs(dev(f[i0:i1]), dev(np.concatenate([g[0][i0:i1].ravel(), g[1].ravel(), g[2][i0:i1].ravel(),
                                                             g[3].ravel()])))
        slab = PeerSlab(dop.ndofs, rank, world)
        slab.x.copy_(u0[sl])
        y = torch.empty(dop.ndofs, dtype=torch.float64, device="cuda")
        for _ in range(nsteps):
            apply_peer_fused(dop, slab, y)   # publishes u's epoch, reads the neighbours', releases
            slab.wait_released()             # the neighbours are done reading u: it may be overwritten
            dop.op.euler_update(slab.x, y, bloc, dt)
        worst = max(worst, float((slab.x - uf[sl]).norm() / uf[sl].norm()) * 1e-3)  # 1e-10 bar on the 1e-13 scale
        assert float((slab.x - uf[sl]).norm()) <= 1e-10 * float(uf[sl].norm())
        col = ny * nf
        slab.publish()
        slab.wait_neighbours()
        import ctypes as _ct
        from stefanproblem_b200.cutcelldg import _stream_ptr
        from stefanproblem_b200 import _lib as _L
        lo = _ct.c_void_p(slab.peer_column(rank - 1, col, last=True)) if rank > 0 else None
        hi = _ct.c_void_p(slab.peer_column(rank + 1, col, last=False)) if rank < world - 1 else None
        sums = torch.empty(4, dtype=torch.float64, device="cuda")
        _L.check(_L.lib().stefan_ipdg_interface_sums(dop.op._h, _ct.c_void_p(slab.x.data_ptr()), lo, hi,
                                                     _ct.c_void_p(sums.data_ptr()), _stream_ptr()))
        slab.release()
        allreduce_interface_sums(sums)
        assert float((sums - sums_ref).abs().max()) <= 1e-9 * float(sums_ref.abs().max())
        torch.cuda.synchronize()
        assert int(slab.status.item()) == 0, "a peer wait timed out"
        dist.barrier()
        slab.close()
        # the same loop with the FUSED step kernel (stefan_ipdg_step, peer form): two buffers alternate and the only
        # ordering is the kernels' own flag protocol.  First half eager (host-side epochs), second half as CUDA-graph
        # replays of a captured pair of steps (device-resident epochs); interface sums of the last iterate.
        slab = PeerSlab(dop.ndofs, rank, world, nbuf=2)
        slab.bufs[0].copy_(u0[sl])
        torch.cuda.synchronize()
        dist.barrier()          # chained halo gate: every rank's initial vector is final before the first step
        half = nsteps // 2
        for k in range(half):
            step_peer_fused(dop, slab, k & 1, (k + 1) & 1, bloc, dt)
        torch.cuda.synchronize()
        dist.barrier()
        slab.to_device_epochs()
        st = torch.cuda.Stream()
        with torch.cuda.stream(st):
            step_peer_fused(dop, slab, half & 1, (half + 1) & 1, bloc, dt, device_epoch=True)       # (warm-up of the mode;
            step_peer_fused(dop, slab, (half + 1) & 1, half & 1, bloc, dt, device_epoch=True)       #  two real steps)
            st.synchronize()
            gr = torch.cuda.CUDAGraph()
            with torch.cuda.graph(gr, stream=st):
                step_peer_fused(dop, slab, half & 1, (half + 1) & 1, bloc, dt, device_epoch=True)
                step_peer_fused(dop, slab, (half + 1) & 1, half & 1, bloc, dt, device_epoch=True)
            for _ in range((nsteps - half - 2) // 2):
                gr.replay()
        torch.cuda.synchronize()
        slab.to_host_epochs()
        done_steps = half + 2 + 2 * ((nsteps - half - 2) // 2)
        for k in range(done_steps, nsteps):
            step_peer_fused(dop, slab, k & 1, (k + 1) & 1, bloc, dt)
        ufin = slab.bufs[nsteps & 1]
        assert float((ufin - uf[sl]).norm()) <= 1e-10 * float(uf[sl].norm()), "fused peer time loop differs from one GPU"
        worst = max(worst, float((ufin - uf[sl]).norm() / uf[sl].norm()) * 1e-3)
        sums2 = torch.zeros(4, dtype=torch.float64, device="cuda")
        interface_sums_peer(dop, slab, nsteps & 1, sums2)
        allreduce_interface_sums(sums2)
        assert float((sums2 - sums_ref).abs().max()) <= 1e-9 * float(sums_ref.abs().max())
        slab.check()
        dist.barrier()
        slab.close()
        # `matrix \\ rhs` across the slabs (stefan_ipdg_cg_solve_slab: NCCL halo of the search direction, all-reduced
        # dot products) against the one-GPU solve of the global system
        from stefanproblem_b200.distributed import cg_solve_distributed
        xfull, ifull = full.solve(bfull, tol=1e-11, maxiter=5000, check_every=5)
        xloc, iloc = cg_solve_distributed(dop, bloc, tol=1e-11, maxiter=5000, check_every=5)
        assert ifull["converged"] and iloc["converged"], (ifull, iloc)
        assert abs(ifull["iterations"] - iloc["iterations"]) <= 5, (ifull, iloc)
        assert float((xloc - xfull[sl]).norm()) <= 1e-6 * float(xfull[sl].norm()), "slab CG differs from the one-GPU solve"
        if rank == 0:
            print(f"slab CG order {order}: {iloc['iterations']} iterations (one GPU: {ifull['iterations']}), "
                  f"residual {iloc['residual_norm']:.2e}", flush=True)
        dist.barrier()
    # local DG (3 dofs per node): peer halos and NCCL halos
    V0 = (np.cos(0.4), np.sin(0.4))
    for order in (1, 2, 3):
        nxg, ny = 13 * world + 3, 65
        nf = (order + 1) ** 2
        dop = DistributedLDGOperator(nxg, ny, (1.0, 1.0), order, order + 1, 1.0, 2.0, 0.5, 0.2, 1.3, V0,
                                     lambda i: column_phase(nxg, ny, i), rank, world)
        basis = cc.LagrangeTensorProductBasis(2, order)
        phase = np.concatenate([column_phase(nxg, ny, i) for i in range(nxg)])
        full = cc.LDGOperator(cc.UniformMesh([0, 0], [1, 1], [nxg, ny], basis, phase=phase), order, order + 1,
                              1.0, 2.0, 0.5, 0.2, 1.3, V0)
        xg = torch.from_numpy(np.random.default_rng(6).uniform(-1, 1, nxg * ny * nf * 3)).cuda()
        yref = full.apply(xg, algo=1)
        i0, i1 = slab_range(nxg, world, rank)
        sl = slice(i0 * ny * nf * 3, i1 * ny * nf * 3)
        slab = PeerSlab(dop.ndofs, rank, world)
        slab.x.copy_(xg[sl])
        out = torch.empty(dop.ndofs, dtype=torch.float64, device="cuda")
        dop.apply_peer(slab, out)
        worst = max(worst, float((out - yref[sl]).norm() / yref[sl].norm()))
        out2 = torch.empty_like(out)
        dop.apply(xg[sl].clone(), out2)
        worst = max(worst, float((out2 - yref[sl]).norm() / yref[sl].norm()))
        torch.cuda.synchronize()
        assert int(slab.status.item()) == 0, "a peer wait timed out"
        dist.barrier()
        slab.close()
    check_fd(rank, world)
    t = torch.tensor([worst], dtype=torch.float64, device="cuda")
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    assert float(t.item()) < 1e-13, f"slab partition differs from the one-GPU operator: {float(t.item()):.3e}"
    if rank == 0:
        print(f"PEER CHECK OK world={world} worst relative difference {float(t.item()):.2e}", flush=True)

    if "--time" in sys.argv:
        order, nxl, ny = 1, 5000, 5000
        nxg = nxl * world
        dop = DistributedIPOperator(nxg, ny, (1.0, 1.0), order, order + 1, 1.0, 2.0, 1e3 * nxg, 1.0, 0.5,
                                    lambda i: column_phase(nxg, ny, i), rank, world)
        slab = PeerSlab(dop.ndofs, rank, world)
        slab.x.uniform_(-1, 1)
        out = torch.empty(dop.ndofs, dtype=torch.float64, device="cuda")
        x2 = slab.x.clone()
        for name, fn in (("peer, static x", lambda: apply_peer(dop, slab, out, publish=False)),
                         ("peer, publish/wait/release per apply", lambda: apply_peer(dop, slab, out)),
                         ("peer, ordering fused into the apply kernel", lambda: apply_peer_fused(dop, slab, out)),
                         ("nccl send/recv + 3 launches", lambda: dop.apply(x2, out))):
            slab.publish()
            slab.wait_neighbours()
            for _ in range(5):
                fn()
            dist.barrier()
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(50):
                fn()
            e1.record()
            dist.barrier()
            torch.cuda.synchronize()
            ms = torch.tensor([e0.elapsed_time(e1) / 50], dtype=torch.float64, device="cuda")
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
            if rank == 0:
                print(f"  {name}: {float(ms):.4f} ms/apply  {dop.ndofs * world / float(ms) / 1e6:.1f} GDOF/s on {world} GPUs",
                      flush=True)
        assert int(slab.status.item()) == 0
        slab.close()
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
