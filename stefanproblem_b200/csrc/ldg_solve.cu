// `matrix \ rhs` for the local-DG system on the device, without the matrix: Schur complement on T,
// BiCGStab with device-resident scalars (krylov.cuh).  Native replacement of the round-1 host loop
// (stefanproblem_b200/ldg_solve.py, whose vector algebra were eager torch calls).
//
// The LDG system of LDG_assembly.jl:1-91 has, per node, the unknowns (T, q_x, q_y):
//     rows T :  A_TT T + A_Tq q = b_T     (penalties; gradient + vector flux)
//     rows q :  A_qT T + M    q = b_q     (divergence + scalar flux; mass, LDG_mass_operator.jl:1-25)
// The only q-q coupling is the cell mass matrix M = detjac (M1 (x) M1) (x) I_2 (block diagonal), so
//     S T = b_T - A_Tq M^-1 b_q,   S = A_TT - A_Tq M^-1 A_qT,   q = M^-1 (b_q - A_qT T).
// One application of S = two launches of the march kernel around a per-cell M^-1 (applied as two 1-D
// solves along the cell's lines).  S is not symmetric for the reference's parameters, hence
// BiCGStab.  Reference sizes are small (n <= 33 cells a side); no preconditioner.
#include "../../include/stefan_b200.h"
#include "common.cuh"
#include "krylov.cuh"
#include "ldg_handle.h"

#include <algorithm>
#include <cmath>
#include <cstring>
#include <utility>
#include <vector>

namespace stefan {

struct LdgMinvTab {
  int n1;
  double minv[kMaxN1][kMaxN1];  // inverse of the 1-D reference mass matrix
  double inv_detjac;
};

// w[3a] = T[a], w[3a+1] = w[3a+2] = 0
__global__ void __launch_bounds__(kry::kThreads)
ldg_pack_T_kernel(const double* __restrict__ T, double* __restrict__ w, int64_t nnodes) {
  for (int64_t a = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; a < nnodes; a += (int64_t)gridDim.x * blockDim.x) {
    w[3 * a] = T[a];
    w[3 * a + 1] = 0.0;
    w[3 * a + 2] = 0.0;
  }
}

// Per cell: out_q = M^-1 (u_q - v_q) (v may be null), out_T = Tsrc ? Tsrc : 0.  One thread per cell.
__global__ void __launch_bounds__(128)
ldg_mass_solve_kernel(const __grid_constant__ LdgMinvTab t, const double* __restrict__ u, const double* __restrict__ v,
                      const double* __restrict__ Tsrc, double* __restrict__ out, int64_t ncells) {
  const int n1 = t.n1, nf = n1 * n1;
  for (int64_t cell = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; cell < ncells;
       cell += (int64_t)gridDim.x * blockDim.x) {
    const int64_t n0 = cell * nf;
    for (int comp = 1; comp <= 2; ++comp) {
      double r[kMaxN1 * kMaxN1], tmp[kMaxN1 * kMaxN1];
      for (int a = 0; a < nf; ++a) r[a] = u[3 * (n0 + a) + comp] - (v ? v[3 * (n0 + a) + comp] : 0.0);
      for (int ix = 0; ix < n1; ++ix)
        for (int iy = 0; iy < n1; ++iy) {
          double s = 0.0;
          for (int c = 0; c < n1; ++c) s += t.minv[iy][c] * r[ix * n1 + c];
          tmp[ix * n1 + iy] = s;
        }
      for (int ix = 0; ix < n1; ++ix)
        for (int iy = 0; iy < n1; ++iy) {
          double s = 0.0;
          for (int c = 0; c < n1; ++c) s += t.minv[ix][c] * tmp[c * n1 + iy];
          out[3 * (n0 + ix * n1 + iy) + comp] = s * t.inv_detjac;
        }
    }
    for (int a = 0; a < nf; ++a) out[3 * (n0 + a)] = Tsrc ? Tsrc[n0 + a] : 0.0;
  }
}

// out[a] = y1[3a] - (y2 ? y2[3a] : 0)
__global__ void __launch_bounds__(kry::kThreads)
ldg_rows_T_kernel(const double* __restrict__ y1, const double* __restrict__ y2, double* __restrict__ out, int64_t nnodes) {
  for (int64_t a = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; a < nnodes; a += (int64_t)gridDim.x * blockDim.x)
    out[a] = y1[3 * a] - (y2 ? y2[3 * a] : 0.0);
}

// g[a] = b[3a] - y[3a]
__global__ void __launch_bounds__(kry::kThreads)
ldg_schur_rhs_kernel(const double* __restrict__ b, const double* __restrict__ y, double* __restrict__ g, int64_t nnodes) {
  for (int64_t a = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; a < nnodes; a += (int64_t)gridDim.x * blockDim.x)
    g[a] = b[3 * a] - y[3 * a];
}

}  // namespace stefan

using namespace stefan;

extern "C" {

int stefan_ldg_solve(stefan_ldg* hv, const double* rhs, double* sol, const stefan_cg_params* prm,
                     stefan_cg_result* res, void* stream) {
  LdgHandle* h = reinterpret_cast<LdgHandle*>(hv);
  STEFAN_REQUIRE(h && rhs && sol && prm && res, "stefan_ldg_solve: null argument");
  STEFAN_REQUIRE(!h->has_lo && !h->has_hi, "stefan_ldg_solve: slab handles are not supported");
  STEFAN_REQUIRE(prm->max_iterations >= 0 && prm->rel_tolerance >= 0.0, "stefan_ldg_solve: bad parameters");
  STEFAN_REQUIRE((reinterpret_cast<uintptr_t>(rhs) & 15) == 0 && (reinterpret_cast<uintptr_t>(sol) & 15) == 0,
                 "stefan_ldg_solve: vectors must be 16-byte aligned");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const int n1 = h->order + 1, nf = n1 * n1;
  const int64_t ncells = (int64_t)h->nx * h->ny, nn = ncells * nf, n3 = 3 * nn;
  LdgMinvTab mt;
  std::memset(&mt, 0, sizeof(mt));
  mt.n1 = n1;
  {
    RefElem1D ref;
    build_ref1d(h->order, h->numqp, &ref);
    // Gauss-Jordan inverse of the 1-D mass matrix
    double a[kMaxN1][2 * kMaxN1];
    for (int i = 0; i < n1; ++i)
      for (int j = 0; j < n1; ++j) {
        a[i][j] = ref.M[i][j];
        a[i][n1 + j] = i == j ? 1.0 : 0.0;
      }
    for (int c = 0; c < n1; ++c) {
      int piv = c;
      for (int r = c + 1; r < n1; ++r)
        if (std::fabs(a[r][c]) > std::fabs(a[piv][c])) piv = r;
      for (int j = 0; j < 2 * n1; ++j) std::swap(a[c][j], a[piv][j]);
      const double d = a[c][c];
      for (int j = 0; j < 2 * n1; ++j) a[c][j] /= d;
      for (int r = 0; r < n1; ++r)
        if (r != c) {
          const double f = a[r][c];
          for (int j = 0; j < 2 * n1; ++j) a[r][j] -= f * a[c][j];
        }
    }
    for (int i = 0; i < n1; ++i)
      for (int j = 0; j < n1; ++j) mt.minv[i][j] = a[i][n1 + j];
  }
  mt.inv_detjac = 1.0 / (0.25 * h->phys_hx * h->phys_hy);

  const int64_t npad = (nn + 3) / 4 * 4, n3pad = (n3 + 3) / 4 * 4;  // 32-byte aligned sub-vectors
  const int nblk = (int)std::min<int64_t>((nn + kry::kThreads - 1) / kry::kThreads, (int64_t)sm_count() * 8);
  const int cblk = (int)std::min<int64_t>((ncells + 127) / 128, (int64_t)sm_count() * 8);
  double* work = nullptr;
  const size_t wbytes = (size_t)(8 * npad + 3 * n3pad + 2 * nblk + kry::S_COUNT) * sizeof(double);
  STEFAN_CUDA(cudaMalloc(&work, wbytes));
  double *g = work, *T = g + npad, *r = T + npad, *r0 = r + npad, *p = r0 + npad, *v = p + npad, *s = v + npad,
         *t = s + npad, *w = t + npad, *y1 = w + n3pad, *y2 = y1 + n3pad, *partial = y2 + n3pad,
         *scal = partial + 2 * nblk;
  int rc = ST_OK;
#define LS_CUDA(call)                                                                        \
  do {                                                                                       \
    cudaError_t e__ = (call);                                                                \
    if (e__ != cudaSuccess) {                                                                \
      cudaFree(work);                                                                        \
      return fail(ST_CUDA_ERROR, "%s:%d: %s -> %s", __FILE__, __LINE__, #call, cudaGetErrorString(e__)); \
    }                                                                                        \
  } while (0)
#define LS_RC(call)            \
  do {                         \
    rc = (call);               \
    if (rc) {                  \
      cudaFree(work);          \
      return rc;               \
    }                          \
  } while (0)
  kry::Reducer red{partial, scal, nblk, st, nullptr, nullptr};
  // S x: out = (A [x; 0])_T - (A [0; M^-1 (A [x; 0])_q])_T
  auto schur = [&](const double* x, double* out) -> int {
    ldg_pack_T_kernel<<<nblk, kry::kThreads, 0, st>>>(x, w, nn);
    int e = stefan_ldg_apply(hv, w, y1, nullptr, nullptr, st);
    if (e) return e;
    ldg_mass_solve_kernel<<<cblk, 128, 0, st>>>(mt, y1, nullptr, nullptr, w, ncells);
    e = stefan_ldg_apply(hv, w, y2, nullptr, nullptr, st);
    if (e) return e;
    ldg_rows_T_kernel<<<nblk, kry::kThreads, 0, st>>>(y1, y2, out, nn);
    return ST_OK;
  };
  // g = b_T - A_Tq M^-1 b_q
  ldg_mass_solve_kernel<<<cblk, 128, 0, st>>>(mt, rhs, nullptr, nullptr, w, ncells);
  LS_RC(stefan_ldg_apply(hv, w, y1, nullptr, nullptr, st));
  ldg_schur_rhs_kernel<<<nblk, kry::kThreads, 0, st>>>(rhs, y1, g, nn);
  LS_CUDA(cudaMemsetAsync(T, 0, npad * sizeof(double), st));
  LS_CUDA(cudaMemsetAsync(p, 0, npad * sizeof(double), st));
  LS_CUDA(cudaMemsetAsync(v, 0, npad * sizeof(double), st));
  LS_CUDA(cudaMemcpyAsync(r, g, nn * sizeof(double), cudaMemcpyDeviceToDevice, st));
  LS_CUDA(cudaMemcpyAsync(r0, g, nn * sizeof(double), cudaMemcpyDeviceToDevice, st));
  const double init[kry::S_COUNT] = {1.0, 1.0, 1.0, 0.0, 0.0, 0.0, 0.0, 0.0};  // rho = alpha = omega = 1
  LS_CUDA(cudaMemcpyAsync(scal, init, sizeof(init), cudaMemcpyHostToDevice, st));
  LS_RC(red.dot(g, g, nullptr, nullptr, nn, kry::P_BB));
  double hs[kry::S_COUNT];
  LS_CUDA(cudaMemcpyAsync(hs, scal, sizeof(hs), cudaMemcpyDeviceToHost, st));
  LS_CUDA(cudaStreamSynchronize(st));
  const double gnorm = std::sqrt(hs[kry::S_BB]);
  const double target = prm->rel_tolerance * gnorm;
  const int check = prm->check_every > 0 ? prm->check_every : 10;
  double rnorm = gnorm;
  int it = 0;
  bool broke = false;
  while (gnorm > 0.0 && it < prm->max_iterations && rnorm > target && !broke) {
    const int batch = std::min(check, prm->max_iterations - it);
    for (int k = 0; k < batch; ++k) {
      LS_RC(red.dot(r0, r, nullptr, nullptr, nn, kry::P_BI_BETA));
      kry::bi_p_kernel<<<nblk, kry::kThreads, 0, st>>>(p, r, v, scal, nn);
      LS_RC(schur(p, v));
      LS_RC(red.dot(r0, v, nullptr, nullptr, nn, kry::P_BI_ALPHA));
      kry::bi_s_kernel<<<nblk, kry::kThreads, 0, st>>>(s, r, v, scal, nn);
      LS_RC(schur(s, t));
      LS_RC(red.dot(t, s, t, t, nn, kry::P_BI_OMEGA));
      kry::bi_xr_kernel<<<nblk, kry::kThreads, 0, st>>>(T, r, p, s, t, scal, nn, partial);
      LS_RC(red.finish(kry::P_RR));
    }
    it += batch;
    LS_CUDA(cudaMemcpyAsync(hs, scal, sizeof(hs), cudaMemcpyDeviceToHost, st));
    LS_CUDA(cudaStreamSynchronize(st));
    rnorm = std::sqrt(hs[kry::S_RR]);
    if (!(rnorm == rnorm) || hs[kry::S_RHO] == 0.0) broke = true;  // NaN or breakdown: report not converged
  }
  // the TRUE residual of the Schur system, not the recurrence's
  if (gnorm > 0.0) {
    LS_RC(schur(T, t));
    kry::residual_kernel<<<nblk, kry::kThreads, 0, st>>>(s, g, t, nn);
    LS_RC(red.dot(s, s, nullptr, nullptr, nn, kry::P_RR));
    LS_CUDA(cudaMemcpyAsync(hs, scal, sizeof(hs), cudaMemcpyDeviceToHost, st));
    LS_CUDA(cudaStreamSynchronize(st));
    rnorm = std::sqrt(hs[kry::S_RR]);
  }
  // q = M^-1 (b_q - A_qT T), T into the solution
  ldg_pack_T_kernel<<<nblk, kry::kThreads, 0, st>>>(T, w, nn);
  LS_RC(stefan_ldg_apply(hv, w, y1, nullptr, nullptr, st));
  ldg_mass_solve_kernel<<<cblk, 128, 0, st>>>(mt, rhs, y1, T, sol, ncells);
  LS_CUDA(cudaGetLastError());
  LS_CUDA(cudaStreamSynchronize(st));
  cudaFree(work);
#undef LS_CUDA
#undef LS_RC
  res->iterations = it;
  res->residual_norm = rnorm;
  res->rhs_norm = gnorm;
  // the recurrence's residual decides when to stop; `converged` reports the true one with a little slack
  res->converged = (gnorm == 0.0 || rnorm <= 10.0 * std::max(target, 1e-15 * gnorm)) ? 1 : 0;
  return ST_OK;
}

}  // extern "C"
