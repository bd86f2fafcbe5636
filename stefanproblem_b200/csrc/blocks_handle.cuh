// Handle and basis helpers of the element-block path (shared by blocks.cu and blocks_solve.cu).
#pragma once
#include <cuda_runtime.h>

#include <cstdint>

#include "../../include/stefan_b200.h"

namespace stefan {

struct BlocksHandle {
  stefan_blocks_dims d;
  int nf;
  double* blocks;       // [ncells][1 + nslots][nf][nf], row-major (row = test function)
  int32_t* slot_nbr;    // [ncells][nslots] copy of the adjacency (apply needs it)
  bool assembled;
  int variant;          // 0 current, 1 legacy boundary form (stefan_blocks_set_variant)
};

// 1-D Lagrange basis on order+1 equispaced nodes of [-1, 1]: value and derivative of N_i at x
template <int P>
__device__ __forceinline__ void lagrange_vd(int i, double x, double& v, double& d) {
  constexpr int n = P + 1;
  const double xi = -1.0 + 2.0 * i / P;
  v = 1.0;
  d = 0.0;
#pragma unroll
  for (int j = 0; j < n; ++j) {
    if (j == i) continue;
    const double xj = -1.0 + 2.0 * j / P;
    v *= (x - xj) / (xi - xj);
    double t = 1.0 / (xi - xj);
#pragma unroll
    for (int k = 0; k < n; ++k) {
      if (k == i || k == j) continue;
      const double xk = -1.0 + 2.0 * k / P;
      t *= (x - xk) / (xi - xk);
    }
    d += t;
  }
}

// value V_a(p) and physical gradient (dV/dx, dV/dy) of the tensor basis function a = ax (P+1) + ay
template <int P>
__device__ __forceinline__ void basis_vg(int a, double px, double py, double jx, double jy, double& v, double& gx,
                                         double& gy) {
  const int ax = a / (P + 1), ay = a % (P + 1);
  double vx, dx, vy, dy;
  lagrange_vd<P>(ax, px, vx, dx);
  lagrange_vd<P>(ay, py, vy, dy);
  v = vx * vy;
  gx = dx * vy / jx;
  gy = vx * dy / jy;
}

// the COO seam of a block array with square blocks of size nf (IP: NF, local DG: 3 NF); explicit zeros for empty /
// boundary slots
static __global__ void blocks_coo_kernel(int ncells, int nslots, int nf, const double* __restrict__ blocks,
                                  const int32_t* __restrict__ slot_nbr, int64_t index_base, int64_t* __restrict__ rows,
                                  int64_t* __restrict__ cols, double* __restrict__ vals) {
  const int nb = 1 + nslots;
  const int64_t n = (int64_t)ncells * nb * nf * nf;
  for (int64_t e = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; e < n; e += (int64_t)gridDim.x * blockDim.x) {
    const int b = (int)(e % nf), a = (int)((e / nf) % nf);
    const int s = (int)((e / ((int64_t)nf * nf)) % nb);
    const int cell = (int)(e / ((int64_t)nf * nf * nb));
    const int nbr = s == 0 ? cell : slot_nbr[(size_t)cell * nslots + (s - 1)];
    rows[e] = (int64_t)cell * nf + a + index_base;
    cols[e] = (nbr >= 0 ? (int64_t)nbr * nf + b : (int64_t)cell * nf + b) + index_base;  // empty slots: explicit zeros
    vals[e] = nbr >= 0 ? blocks[e] : 0.0;
  }
}

}  // namespace stefan
