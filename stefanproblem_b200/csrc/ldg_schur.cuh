// Schur-complement BiCGStab for local-DG systems, shared by the uniform-mesh solver (ldg_solve.cu: the operator is
// the march kernel, M^-1 two 1-D solves per cell) and the element-block solver (ldg_blocks.cu: assembled blocks,
// per-cell dense M^-1).  See ldg_solve.cu for the algebra.
//   apply(w, y):  y = A w on interleaved (T, q_x, q_y) vectors, returns a status
//   mass_solve(u, v, Tsrc, out):  out_q = M^-1 (u_q - v_q) (v may be null), out_T = Tsrc ? Tsrc : 0
#pragma once
#include <cuda_runtime.h>

#include <algorithm>
#include <cmath>
#include <cstdint>

#include "../../include/stefan_b200.h"
#include "common.cuh"
#include "krylov.cuh"

namespace stefan {

// w[3a] = T[a], w[3a+1] = w[3a+2] = 0
static __global__ void __launch_bounds__(kry::kThreads)
ldg_pack_T_kernel(const double* __restrict__ T, double* __restrict__ w, int64_t nnodes) {
  for (int64_t a = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; a < nnodes; a += (int64_t)gridDim.x * blockDim.x) {
    w[3 * a] = T[a];
    w[3 * a + 1] = 0.0;
    w[3 * a + 2] = 0.0;
  }
}

// out[a] = y1[3a] - (y2 ? y2[3a] : 0)
static __global__ void __launch_bounds__(kry::kThreads)
ldg_rows_T_kernel(const double* __restrict__ y1, const double* __restrict__ y2, double* __restrict__ out, int64_t nnodes) {
  for (int64_t a = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; a < nnodes; a += (int64_t)gridDim.x * blockDim.x)
    out[a] = y1[3 * a] - (y2 ? y2[3 * a] : 0.0);
}

// g[a] = b[3a] - y[3a]
static __global__ void __launch_bounds__(kry::kThreads)
ldg_schur_rhs_kernel(const double* __restrict__ b, const double* __restrict__ y, double* __restrict__ g, int64_t nnodes) {
  for (int64_t a = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; a < nnodes; a += (int64_t)gridDim.x * blockDim.x)
    g[a] = b[3 * a] - y[3 * a];
}


template <class Apply, class MassSolve>
int ldg_schur_bicgstab(Apply&& apply, MassSolve&& mass_solve, int64_t nn, const double* rhs, double* sol,
                       const stefan_cg_params* prm, stefan_cg_result* res, cudaStream_t st) {
  const int64_t n3 = 3 * nn;
  const int64_t npad = (nn + 3) / 4 * 4, n3pad = (n3 + 3) / 4 * 4;  // 32-byte aligned sub-vectors
  const int nblk = (int)std::min<int64_t>((nn + kry::kThreads - 1) / kry::kThreads, (int64_t)sm_count() * 8);
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
    int e = apply(w, y1);
    if (e) return e;
    mass_solve(y1, nullptr, nullptr, w);
    e = apply(w, y2);
    if (e) return e;
    ldg_rows_T_kernel<<<nblk, kry::kThreads, 0, st>>>(y1, y2, out, nn);
    return ST_OK;
  };
  // g = b_T - A_Tq M^-1 b_q
  mass_solve(rhs, nullptr, nullptr, w);
  LS_RC(apply(w, y1));
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
    // NaN, breakdown or divergence (BiCGStab on a Schur complement that is too ill-conditioned): report not converged
    if (!(rnorm == rnorm) || hs[kry::S_RHO] == 0.0 || rnorm > 1e8 * gnorm) broke = true;
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
  LS_RC(apply(w, y1));
  mass_solve(rhs, y1, T, sol);
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

}  // namespace stefan
