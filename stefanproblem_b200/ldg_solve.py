"""`matrix \\ rhs` for the local-DG system without the matrix: Schur complement on T.

The LDG system of LDG_assembly.jl:1-91 has, per node, the unknowns (T, q_x, q_y) and the
block structure

    rows T :  A_TT T + A_Tq q = b_T     (penalties; gradient + vector flux)
    rows q :  A_qT T + M    q = b_q     (divergence + scalar flux; mass, LDG_mass_operator.jl:1-25)

where the only q-q coupling is the cell mass matrix M = detjac * (M1 (x) M1) (x) I_2, block
diagonal with one NF x NF block per cell.  Eliminating q = M^-1 (b_q - A_qT T) leaves

    S T = b_T - A_Tq M^-1 b_q ,     S = A_TT - A_Tq M^-1 A_qT

and one S application is two applications of the device operator (first with q = 0, then
with T = 0) around the per-cell M^-1.  S is not symmetric for the reference's parameters
(the scalar flux switches with +-1/2, the vector flux with n . V0), so the
iteration is BiCGStab.  No preconditioner: meant for the sizes of the reference's
convergence studies; the operator applications are the library's kernels, the vector
updates and dot products are torch calls on the same stream.

The functions take the operator as a callable, so the CPU tests drive them with the
oracle's assembled matrix.
"""
import numpy as np


def cell_mass_inverse(mass1d, detjac):
    """Inverse of detjac * (M1 (x) M1): [NF, NF], node order x-node outer, y-node inner."""
    return np.linalg.inv(detjac * np.kron(mass1d, mass1d))


def schur_solve(apply, minv, rhs, nf, tol=1e-12, maxiter=20000):
    """Solve A [T; q] = rhs for a local-DG operator `apply` (vectors interleaved (T, qx, qy)
    per node, cells contiguous with `nf` nodes each); `minv` = `cell_mass_inverse` as a tensor
    on the vectors' device.  Returns (solution, info)."""
    import torch
    n = rhs.numel() // 3
    b = rhs.view(n, 3)
    work = torch.zeros_like(b)

    def mass_solve(q):                       # q: [n, 2] -> M^-1 q
        return torch.einsum("ab,cbk->cak", minv, q.reshape(-1, nf, 2)).reshape(n, 2)

    def a_times_T(T):                        # A [T; 0]: ([n] rows T, [n, 2] rows q)
        work.zero_()
        work[:, 0] = T
        y = apply(work.view(-1)).view(n, 3)
        return y[:, 0].clone(), y[:, 1:3].clone()

    def a_times_q_rows_T(q):                 # rows T of A [0; q]
        work.zero_()
        work[:, 1:3] = q
        return apply(work.view(-1)).view(n, 3)[:, 0].clone()

    def schur(T):
        yT, yq = a_times_T(T)
        return yT - a_times_q_rows_T(mass_solve(yq))

    g = b[:, 0] - a_times_q_rows_T(mass_solve(b[:, 1:3]))
    gnorm = float(g.norm())
    T = torch.zeros_like(g)
    info = {"iterations": 0, "converged": True, "relative_residual": 0.0, "operator_applications": 1}
    if gnorm > 0.0:
        # BiCGStab (van der Vorst), r0* = r0
        r = g.clone()
        r0 = r.clone()
        rho = alpha = omega = 1.0
        v = torch.zeros_like(g)
        p = torch.zeros_like(g)
        info["converged"] = False
        for it in range(1, maxiter + 1):
            rho_new = float(torch.dot(r0, r))
            if rho_new == 0.0 or not np.isfinite(rho_new):   # breakdown / divergence: report not converged
                break
            beta = (rho_new / rho) * (alpha / omega)
            p = r + beta * (p - omega * v)
            v = schur(p)
            alpha = rho_new / float(torch.dot(r0, v))
            s = r - alpha * v
            info["iterations"] = it
            if float(s.norm()) <= tol * gnorm:
                T += alpha * p
                r = s
                info["converged"] = True
                break
            t = schur(s)
            omega = float(torch.dot(t, s)) / float(torch.dot(t, t))
            T += alpha * p + omega * s
            r = s - omega * t
            rho = rho_new
            if float(r.norm()) <= tol * gnorm:
                info["converged"] = True
                break
        # report the TRUE residual of the Schur system, not the recurrence's
        info["relative_residual"] = float((g - schur(T)).norm()) / gnorm
        info["operator_applications"] = 1 + 2 * (2 * info["iterations"] + 1)
    _, yq = a_times_T(T)
    sol = torch.empty_like(b)
    sol[:, 0] = T
    sol[:, 1:3] = mass_solve(b[:, 1:3] - yq)
    return sol.view(-1), info
