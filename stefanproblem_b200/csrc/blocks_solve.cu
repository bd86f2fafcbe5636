// Element-block path, everything around the operator: right-hand side, error norm and the solve.
//
//   stefan_blocks_rhs        linear_form / flux_linear_form / robin_interface_linear_form per solution cell
//                            (IP_source_term.jl:97-159, IP_robin_interface_source.jl:1-25): what
//                            assemble_interior_penalty_rhs! + assemble_two_phase_source! + assemble_robin_source!
//                            add, for arbitrary per-cell / per-face quadratures
//   stefan_blocks_l2_error   mesh_L2_error / integral_norm_on_mesh (useful_routines.jl:95-133, :203-221)
//   stefan_blocks_cg_solve   `matrix \ rhs` (cylindrical_bc_tests/IP_test_cylindrical_bc_convergence.jl:87-90):
//                            conjugate gradients over stefan_blocks_apply, block-Jacobi preconditioner (the
//                            inverse of every solution cell's diagonal block, inverted on the device),
//                            iteration scalars on the device (krylov.cuh)
#include "blocks_handle.cuh"
#include "common.cuh"
#include "krylov.cuh"

#include <algorithm>
#include <cmath>

namespace stefan {

// one warp per cell; lane a (strided) owns basis function a
template <int P>
__global__ void __launch_bounds__(128)
blocks_rhs_kernel(const stefan_blocks_dims d, const stefan_blocks_data in, const double* __restrict__ f_qp,
                  const double* __restrict__ g_qp, double Tm, int legacy, double* __restrict__ rhs) {
  constexpr int NF = (P + 1) * (P + 1);
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, wpb = blockDim.x >> 5;
  for (int cell = blockIdx.x * wpb + warp; cell < d.ncells; cell += gridDim.x * wpb) {
    const double jx = in.cell_jac[2 * cell], jy = in.cell_jac[2 * cell + 1], k = in.cell_k[cell];
    for (int a = lane; a < NF; a += 32) {
      double acc = 0.0;
      if (f_qp != nullptr)
        for (int q = 0; q < d.nqv; ++q) {
          const size_t g = (size_t)cell * d.nqv + q;
          const double w = in.vol_wts[g];
          if (w == 0.0) continue;
          double v, gx, gy;
          basis_vg<P>(a, in.vol_pts[2 * g], in.vol_pts[2 * g + 1], jx, jy, v, gx, gy);
          acc += v * f_qp[g] * w * jx * jy;
        }
      for (int s = 0; s < d.nslots; ++s) {
        const size_t cs = (size_t)cell * d.nslots + s;
        const int nbr = in.slot_nbr[cs];
        const double lam = in.slot_lambda[cs];
        if (nbr == -1 && g_qp != nullptr) {
          const double pen = in.slot_pen[cs];
          for (int q = 0; q < d.nqf; ++q) {
            const size_t fq = cs * d.nqf + q;
            const double w = in.face_wts[fq];
            if (w == 0.0) continue;
            double v, gx, gy;
            basis_vg<P>(a, in.face_pts[2 * fq], in.face_pts[2 * fq + 1], jx, jy, v, gx, gy);
            const double gn = gx * in.face_normals[2 * fq] + gy * in.face_normals[2 * fq + 1];
            acc += (pen * v - (legacy ? 0.0 : k * gn)) * g_qp[fq] * w;   // R1 - R2 (IP_source_term.jl:131-159); legacy: R1
          }
        } else if (nbr >= 0 && lam != 0.0 && Tm != 0.0) {
          for (int q = 0; q < d.nqf; ++q) {
            const size_t fq = cs * d.nqf + q;
            const double w = in.face_wts[fq];
            if (w == 0.0) continue;
            double v, gx, gy;
            basis_vg<P>(a, in.face_pts[2 * fq], in.face_pts[2 * fq + 1], jx, jy, v, gx, gy);
            acc += 0.5 * lam * Tm * v * w;              // IP_robin_interface_source.jl:12-25
          }
        }
      }
      rhs[(size_t)cell * NF + a] = acc;
    }
  }
}

// per block: partial[2 b] = sum (u_h - u_ex)^2 detjac w, partial[2 b + 1] = sum u_ex^2 detjac w; one thread per point
template <int P>
__global__ void __launch_bounds__(kry::kThreads)
blocks_l2_kernel(const stefan_blocks_dims d, const double* __restrict__ vol_pts, const double* __restrict__ vol_wts,
                 const double* __restrict__ cell_jac, const double* __restrict__ u, const double* __restrict__ uex,
                 double* __restrict__ partial) {
  constexpr int NF = (P + 1) * (P + 1);
  const int64_t npts = (int64_t)d.ncells * d.nqv;
  double e2 = 0.0, n2 = 0.0;
  for (int64_t g = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; g < npts; g += (int64_t)gridDim.x * blockDim.x) {
    const double w = vol_wts[g];
    if (w == 0.0) continue;
    const int cell = (int)(g / d.nqv);
    const double jx = cell_jac[2 * cell], jy = cell_jac[2 * cell + 1];
    double uh = 0.0;
#pragma unroll
    for (int a = 0; a < NF; ++a) {
      double v, gx, gy;
      basis_vg<P>(a, vol_pts[2 * g], vol_pts[2 * g + 1], jx, jy, v, gx, gy);
      uh += v * u[(size_t)cell * NF + a];
    }
    const double ue = uex[g];
    e2 += (uh - ue) * (uh - ue) * jx * jy * w;
    n2 += ue * ue * jx * jy * w;
  }
  e2 = kry::block_sum(e2);
  n2 = kry::block_sum(n2);
  if (threadIdx.x == 0) {
    partial[2 * blockIdx.x] = e2;
    partial[2 * blockIdx.x + 1] = n2;
  }
}

__global__ void blocks_l2_final_kernel(const double* __restrict__ partial, int nblocks, double* __restrict__ out2) {
  if (threadIdx.x != 0 || blockIdx.x != 0) return;
  double a = 0.0, b = 0.0;
  for (int k = 0; k < nblocks; ++k) {
    a += partial[2 * k];
    b += partial[2 * k + 1];
  }
  out2[0] = sqrt(a);
  out2[1] = sqrt(b);
}

// inverse of every cell's diagonal block (Gauss-Jordan with partial pivoting, one thread per cell);
// a singular block (empty solution cell) gets the identity
__global__ void __launch_bounds__(64)
blocks_invert_diag_kernel(int ncells, int nslots, int nf, const double* __restrict__ blocks, double* __restrict__ dinv) {
  constexpr int MAXNF = 16;
  for (int cell = blockIdx.x * blockDim.x + threadIdx.x; cell < ncells; cell += gridDim.x * blockDim.x) {
    double m[MAXNF][2 * MAXNF];
    const double* B = blocks + (size_t)cell * (1 + nslots) * nf * nf;
    for (int i = 0; i < nf; ++i)
      for (int j = 0; j < nf; ++j) {
        m[i][j] = B[i * nf + j];
        m[i][nf + j] = i == j ? 1.0 : 0.0;
      }
    bool ok = true;
    for (int c = 0; c < nf && ok; ++c) {
      int piv = c;
      for (int r = c + 1; r < nf; ++r)
        if (fabs(m[r][c]) > fabs(m[piv][c])) piv = r;
      if (m[piv][c] == 0.0) {
        ok = false;
        break;
      }
      for (int j = 0; j < 2 * nf; ++j) {
        const double t = m[c][j];
        m[c][j] = m[piv][j];
        m[piv][j] = t;
      }
      const double dd = 1.0 / m[c][c];
      for (int j = 0; j < 2 * nf; ++j) m[c][j] *= dd;
      for (int r = 0; r < nf; ++r)
        if (r != c) {
          const double f = m[r][c];
          if (f != 0.0)
            for (int j = 0; j < 2 * nf; ++j) m[r][j] -= f * m[c][j];
        }
    }
    double* D = dinv + (size_t)cell * nf * nf;
    for (int i = 0; i < nf; ++i)
      for (int j = 0; j < nf; ++j) D[i * nf + j] = ok ? m[i][nf + j] : (i == j ? 1.0 : 0.0);
  }
}

// z = Dinv r, one thread per row
__global__ void __launch_bounds__(kry::kThreads)
blocks_precond_kernel(int64_t nrows, int nf, const double* __restrict__ dinv, const double* __restrict__ r,
                      double* __restrict__ z) {
  for (int64_t row = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; row < nrows; row += (int64_t)gridDim.x * blockDim.x) {
    const int64_t cell = row / nf;
    const int a = (int)(row - cell * nf);
    const double* D = dinv + ((size_t)cell * nf + a) * nf;
    const double* rc = r + cell * nf;
    double v = 0.0;
    for (int b = 0; b < nf; ++b) v += D[b] * rc[b];
    z[row] = v;
  }
}

}  // namespace stefan

using namespace stefan;

extern "C" {

int stefan_blocks_rhs(stefan_blocks* hv, const stefan_blocks_data* in, const double* f_qp, const double* g_qp,
                      double Tm, double* rhs, void* stream) {
  BlocksHandle* h = reinterpret_cast<BlocksHandle*>(hv);
  STEFAN_REQUIRE(h && in && rhs, "stefan_blocks_rhs: null argument");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const int blocks = (int)std::min<int64_t>(((int64_t)h->d.ncells + 3) / 4, (int64_t)sm_count() * 16);
  switch (h->d.order) {
    case 1: blocks_rhs_kernel<1><<<blocks, 128, 0, st>>>(h->d, *in, f_qp, g_qp, Tm, h->variant, rhs); break;
    case 2: blocks_rhs_kernel<2><<<blocks, 128, 0, st>>>(h->d, *in, f_qp, g_qp, Tm, h->variant, rhs); break;
    case 3: blocks_rhs_kernel<3><<<blocks, 128, 0, st>>>(h->d, *in, f_qp, g_qp, Tm, h->variant, rhs); break;
  }
  STEFAN_CUDA(cudaGetLastError());
  return ST_OK;
}

int stefan_blocks_l2_error(stefan_blocks* hv, const stefan_blocks_data* in, const double* u, const double* uex_qp,
                           double* out2_dev, void* stream) {
  BlocksHandle* h = reinterpret_cast<BlocksHandle*>(hv);
  STEFAN_REQUIRE(h && in && u && uex_qp && out2_dev, "stefan_blocks_l2_error: null argument");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const int64_t npts = (int64_t)h->d.ncells * h->d.nqv;
  const int nblk = (int)std::min<int64_t>((npts + kry::kThreads - 1) / kry::kThreads, (int64_t)sm_count() * 8);
  double* part = nullptr;
  STEFAN_CUDA(cudaMallocAsync(&part, 2 * (size_t)nblk * sizeof(double), st));
  switch (h->d.order) {
    case 1: blocks_l2_kernel<1><<<nblk, kry::kThreads, 0, st>>>(h->d, in->vol_pts, in->vol_wts, in->cell_jac, u, uex_qp, part); break;
    case 2: blocks_l2_kernel<2><<<nblk, kry::kThreads, 0, st>>>(h->d, in->vol_pts, in->vol_wts, in->cell_jac, u, uex_qp, part); break;
    case 3: blocks_l2_kernel<3><<<nblk, kry::kThreads, 0, st>>>(h->d, in->vol_pts, in->vol_wts, in->cell_jac, u, uex_qp, part); break;
  }
  blocks_l2_final_kernel<<<1, 1, 0, st>>>(part, nblk, out2_dev);
  STEFAN_CUDA(cudaFreeAsync(part, st));
  STEFAN_CUDA(cudaGetLastError());
  return ST_OK;
}

int stefan_blocks_cg_solve(stefan_blocks* hv, const double* b, double* x, const stefan_cg_params* prm,
                           stefan_cg_result* res, void* stream) {
  BlocksHandle* h = reinterpret_cast<BlocksHandle*>(hv);
  STEFAN_REQUIRE(h && b && x && prm && res, "stefan_blocks_cg_solve: null argument");
  STEFAN_REQUIRE(h->assembled, "stefan_blocks_cg_solve: call stefan_blocks_assemble first");
  STEFAN_REQUIRE(prm->max_iterations >= 0 && prm->rel_tolerance >= 0.0, "stefan_blocks_cg_solve: bad parameters");
  STEFAN_REQUIRE(prm->preconditioner == 0 || prm->preconditioner == 1, "stefan_blocks_cg_solve: unknown preconditioner");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const int nf = h->nf;
  const int64_t n = (int64_t)h->d.ncells * nf;
  const int nblk = (int)std::min<int64_t>((n + kry::kThreads - 1) / kry::kThreads, (int64_t)sm_count() * 8);
  const size_t dinv_n = prm->preconditioner == 1 ? (size_t)h->d.ncells * nf * nf : 0;
  double* work = nullptr;
  STEFAN_CUDA(cudaMalloc(&work, (size_t)(4 * n + dinv_n + 2 * nblk + kry::S_COUNT) * sizeof(double)));
  double *r = work, *z = r + n, *p = z + n, *Ap = p + n, *dinv = Ap + n, *partial = dinv + dinv_n,
         *scal = partial + 2 * nblk;
#define BS_CUDA(call)                                                                        \
  do {                                                                                       \
    cudaError_t e__ = (call);                                                                \
    if (e__ != cudaSuccess) {                                                                \
      cudaFree(work);                                                                        \
      return fail(ST_CUDA_ERROR, "%s:%d: %s -> %s", __FILE__, __LINE__, #call, cudaGetErrorString(e__)); \
    }                                                                                        \
  } while (0)
#define BS_RC(call)     \
  do {                  \
    int rc__ = (call);  \
    if (rc__) {         \
      cudaFree(work);   \
      return rc__;      \
    }                   \
  } while (0)
  kry::Reducer red{partial, scal, nblk, st, nullptr, nullptr};
  BS_CUDA(cudaMemsetAsync(scal, 0, kry::S_COUNT * sizeof(double), st));
  if (dinv_n)
    blocks_invert_diag_kernel<<<(h->d.ncells + 63) / 64, 64, 0, st>>>(h->d.ncells, h->d.nslots, nf, h->blocks, dinv);
  auto precond = [&](const double* rr, double* zz) {
    if (dinv_n) blocks_precond_kernel<<<nblk, kry::kThreads, 0, st>>>(n, nf, dinv, rr, zz);
    else cudaMemcpyAsync(zz, rr, n * sizeof(double), cudaMemcpyDeviceToDevice, st);
  };
  BS_RC(red.dot(b, b, nullptr, nullptr, n, kry::P_BB));
  BS_RC(stefan_blocks_apply(hv, x, Ap, st));
  kry::residual_kernel<<<nblk, kry::kThreads, 0, st>>>(r, b, Ap, n);
  precond(r, z);
  BS_RC(red.dot(r, z, r, r, n, kry::P_CG_START));
  kry::cg_p_kernel<<<nblk, kry::kThreads, 0, st>>>(p, z, scal, 1, n);
  double hs[kry::S_COUNT];
  BS_CUDA(cudaMemcpyAsync(hs, scal, sizeof(hs), cudaMemcpyDeviceToHost, st));
  BS_CUDA(cudaStreamSynchronize(st));
  const double bnorm = std::sqrt(hs[kry::S_BB]);
  const double target = prm->rel_tolerance * (bnorm > 0 ? bnorm : 1.0);
  const int check = prm->check_every > 0 ? prm->check_every : 10;
  double rnorm = std::sqrt(hs[kry::S_RR]);
  int it = 0;
  while (it < prm->max_iterations && rnorm > target) {
    const int batch = std::min(check, prm->max_iterations - it);
    for (int k = 0; k < batch; ++k) {
      BS_RC(stefan_blocks_apply(hv, p, Ap, st));
      BS_RC(red.dot(p, Ap, nullptr, nullptr, n, kry::P_CG_ALPHA));
      kry::cg_xr_kernel<<<nblk, kry::kThreads, 0, st>>>(x, r, p, Ap, scal, n);
      precond(r, z);
      BS_RC(red.dot(r, z, r, r, n, kry::P_CG_BETA));
      kry::cg_p_kernel<<<nblk, kry::kThreads, 0, st>>>(p, z, scal, 0, n);
    }
    it += batch;
    BS_CUDA(cudaMemcpyAsync(hs, scal, sizeof(hs), cudaMemcpyDeviceToHost, st));
    BS_CUDA(cudaStreamSynchronize(st));
    rnorm = std::sqrt(hs[kry::S_RR]);
    if (!(rnorm == rnorm)) break;  // NaN: breakdown
  }
  BS_CUDA(cudaGetLastError());
  cudaFree(work);
#undef BS_CUDA
#undef BS_RC
  res->iterations = it;
  res->residual_norm = rnorm;
  res->rhs_norm = bnorm;
  res->converged = rnorm <= target ? 1 : 0;
  return ST_OK;
}

}  // extern "C"
