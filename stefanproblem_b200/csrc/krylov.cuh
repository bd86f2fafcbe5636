// Device-resident Krylov building blocks shared by the solvers that replace the reference's
// `matrix \ rhs` (cg.cu: IP-DG CG; blocks.cu: CG on the element-block operator; ldg_solve.cu:
// Schur-complement BiCGStab for the local-DG systems).
//
// All scalars of an iteration (rho, alpha, beta, omega, norms) live in device memory: the vector
// kernels leave one partial sum per block, `kry_scalar_kernel` adds them in a fixed order (the
// reductions are deterministic) and runs a small "scalar program", the next vector kernel reads the
// result.  The host reads the residual only every `check_every` iterations.  An optional all-reduce
// hook between the summation and the scalar program makes the same loops run across column slabs
// (one process per GPU): the dots are summed over the ranks, everything else is rank-local.
#pragma once
#include <cuda_runtime.h>

#include <cstdint>

namespace stefan {
namespace kry {

constexpr int kThreads = 256;

// scalar slots (device doubles)
enum Slot : int {
  S_RHO = 0, S_ALPHA = 1, S_OMEGA = 2, S_BETA = 3,
  S_SUM0 = 4, S_SUM1 = 5,       // the sums of the last reduction (all-reduced across ranks when there is a hook)
  S_RR = 6,                     // ||r||^2 (or ||s||^2 in mid-iteration of BiCGStab)
  S_BB = 7,                     // ||b||^2
  S_COUNT = 8,
};

// scalar programs run by kry_scalar_kernel after the sums are final
enum Program : int {
  P_NONE = 0,
  P_BB,            // bb = sum0
  P_RR,            // rr = sum0
  P_CG_START,      // rho = sum0 (r.z), rr = sum1
  P_CG_ALPHA,      // alpha = rho / sum0 (p.Ap)
  P_CG_BETA,       // beta = sum0 / rho, rho = sum0 (r.z new), rr = sum1
  P_BI_BETA,       // beta = (sum0 / rho) (alpha / omega), rho = sum0   (sum0 = r0.r)
  P_BI_ALPHA,      // alpha = rho / sum0                                 (sum0 = r0.v)
  P_BI_OMEGA,      // omega = sum0 / sum1                                (t.s, t.t)
};

__device__ __forceinline__ double block_sum(double v) {
  __shared__ double sh[kThreads / 32];
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_down_sync(0xffffffffu, v, o);
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  __syncthreads();
  if (lane == 0) sh[w] = v;
  __syncthreads();
  double s = 0.0;
  if (threadIdx.x == 0)
    for (int k = 0; k < kThreads / 32; ++k) s += sh[k];
  return s;  // valid in thread 0
}

// partial[2 blk] = a.b, partial[2 blk + 1] = c.d (c may be null)
static __global__ void __launch_bounds__(kThreads)
dot2_kernel(const double* __restrict__ a, const double* __restrict__ b, const double* __restrict__ c,
            const double* __restrict__ d, int64_t n, double* __restrict__ partial) {
  double s0 = 0.0, s1 = 0.0;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    s0 += a[i] * b[i];
    if (c != nullptr) s1 += c[i] * d[i];
  }
  s0 = block_sum(s0);
  s1 = block_sum(s1);
  if (threadIdx.x == 0) {
    partial[2 * blockIdx.x] = s0;
    partial[2 * blockIdx.x + 1] = s1;
  }
}

// fixed-order sum of the partials into S_SUM0 / S_SUM1
static __global__ void sum_kernel(const double* __restrict__ partial, int nblocks, double* __restrict__ scal) {
  if (threadIdx.x != 0 || blockIdx.x != 0) return;
  double s0 = 0.0, s1 = 0.0;
  for (int k = 0; k < nblocks; ++k) {
    s0 += partial[2 * k];
    s1 += partial[2 * k + 1];
  }
  scal[S_SUM0] = s0;
  scal[S_SUM1] = s1;
}

static __global__ void scalar_kernel(double* __restrict__ scal, int program) {
  if (threadIdx.x != 0 || blockIdx.x != 0) return;
  const double s0 = scal[S_SUM0], s1 = scal[S_SUM1];
  switch (program) {
    case P_BB: scal[S_BB] = s0; break;
    case P_RR: scal[S_RR] = s0; break;
    case P_CG_START: scal[S_RHO] = s0; scal[S_RR] = s1; break;
    case P_CG_ALPHA: scal[S_ALPHA] = s0 != 0.0 ? scal[S_RHO] / s0 : 0.0; break;
    case P_CG_BETA:
      scal[S_BETA] = scal[S_RHO] != 0.0 ? s0 / scal[S_RHO] : 0.0;
      scal[S_RHO] = s0;
      scal[S_RR] = s1;
      break;
    case P_BI_BETA:
      scal[S_BETA] = (scal[S_RHO] != 0.0 && scal[S_OMEGA] != 0.0) ? (s0 / scal[S_RHO]) * (scal[S_ALPHA] / scal[S_OMEGA]) : 0.0;
      scal[S_RHO] = s0;
      break;
    case P_BI_ALPHA: scal[S_ALPHA] = s0 != 0.0 ? scal[S_RHO] / s0 : 0.0; break;
    case P_BI_OMEGA: scal[S_OMEGA] = s1 != 0.0 ? s0 / s1 : 0.0; break;
    default: break;
  }
}

// r = b - Ax
static __global__ void __launch_bounds__(kThreads)
residual_kernel(double* __restrict__ r, const double* __restrict__ b, const double* __restrict__ Ax, int64_t n) {
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
    r[i] = b[i] - Ax[i];
}

// CG: x += alpha p, r -= alpha Ap
static __global__ void __launch_bounds__(kThreads)
cg_xr_kernel(double* __restrict__ x, double* __restrict__ r, const double* __restrict__ p, const double* __restrict__ Ap,
             const double* __restrict__ scal, int64_t n) {
  const double alpha = scal[S_ALPHA];
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    x[i] += alpha * p[i];
    r[i] -= alpha * Ap[i];
  }
}

// CG: p = z + beta p (first: p = z)
static __global__ void __launch_bounds__(kThreads)
cg_p_kernel(double* __restrict__ p, const double* __restrict__ z, const double* __restrict__ scal, int first, int64_t n) {
  const double beta = first ? 0.0 : scal[S_BETA];
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
    p[i] = first ? z[i] : z[i] + beta * p[i];
}

// BiCGStab: p = r + beta (p - omega v)
static __global__ void __launch_bounds__(kThreads)
bi_p_kernel(double* __restrict__ p, const double* __restrict__ r, const double* __restrict__ v,
            const double* __restrict__ scal, int64_t n) {
  const double beta = scal[S_BETA], omega = scal[S_OMEGA];
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
    p[i] = r[i] + beta * (p[i] - omega * v[i]);
}

// BiCGStab: s = r - alpha v
static __global__ void __launch_bounds__(kThreads)
bi_s_kernel(double* __restrict__ s, const double* __restrict__ r, const double* __restrict__ v,
            const double* __restrict__ scal, int64_t n) {
  const double alpha = scal[S_ALPHA];
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
    s[i] = r[i] - alpha * v[i];
}

// BiCGStab: x += alpha p + omega s, r = s - omega t; partial[2 blk] = r.r
static __global__ void __launch_bounds__(kThreads)
bi_xr_kernel(double* __restrict__ x, double* __restrict__ r, const double* __restrict__ p, const double* __restrict__ s,
             const double* __restrict__ t, const double* __restrict__ scal, int64_t n, double* __restrict__ partial) {
  const double alpha = scal[S_ALPHA], omega = scal[S_OMEGA];
  double rr = 0.0;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    x[i] += alpha * p[i] + omega * s[i];
    const double rv = s[i] - omega * t[i];
    r[i] = rv;
    rr += rv * rv;
  }
  rr = block_sum(rr);
  if (threadIdx.x == 0) {
    partial[2 * blockIdx.x] = rr;
    partial[2 * blockIdx.x + 1] = 0.0;
  }
}

// all-reduce hook: sums `count` device doubles over the ranks on `stream`; 0 = ok
typedef int (*AllReduceFn)(void* ctx, double* dev, int count, void* stream);

struct Reducer {
  double* partial;   // [2 nblk]
  double* scal;      // [S_COUNT]
  int nblk;
  cudaStream_t st;
  AllReduceFn hook;
  void* hook_ctx;
  // sums -> (all-reduce) -> scalar program; returns 0 or the hook's error
  int finish(int program) const {
    sum_kernel<<<1, 1, 0, st>>>(partial, nblk, scal);
    if (hook != nullptr) {
      int rc = hook(hook_ctx, scal + S_SUM0, 2, st);
      if (rc) return rc;
    }
    scalar_kernel<<<1, 1, 0, st>>>(scal, program);
    return 0;
  }
  int dot(const double* a, const double* b, const double* c, const double* d, int64_t n, int program) const {
    dot2_kernel<<<nblk, kThreads, 0, st>>>(a, b, c, d, n, partial);
    return finish(program);
  }
};

}  // namespace kry
}  // namespace stefan
