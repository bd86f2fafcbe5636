// Handle behind `stefan_ipdg*` (shared by ipdg.cu and ipdg_aux.cu).
#pragma once
#include <cuda_runtime.h>

#include <cmath>
#include <cstdint>
#include <utility>
#include <vector>

#include "ipdg_tables.h"
#include "march_plan.h"

namespace stefan {

constexpr int kTailRing = 8;

struct IpdgHandle {
  int nx, ny, order, numqp;
  double hx, hy;
  IpPhys phys;
  double Tm;
  bool has_lo, has_hi;
  int8_t* phase_dev;  // (nx+2)*(ny+2), padded
  MarchPlan plan;      // run lists, fix list and CTA tiles of the march kernel
  double *xh_dev, *yh_dev;       // device staging of the host path
  // tail-queue / finished-CTA counters of the march kernels: a ring of kTailRing {next, finished}
  // pairs, one per launch in flight (launches of one handle may overlap on different streams)
  unsigned* tail_ctr_dev;
  unsigned launch_seq;
  double minv1d[kMaxN1][kMaxN1];  // inverse of the 1-D reference mass matrix (fused explicit step)
  // interface faces owned by this slab (faces of +1 cells towards a -1 cell, halo columns included):
  // {cell, face 0 bottom / 1 right / 2 top / 3 left}; partial sums + ticket of the one-launch reduction
  int2* iface_dev;
  int niface;
  double* iface_partial_dev;
  unsigned* iface_ctr_dev;
  cudaStream_t hs[3];
  cudaEvent_t hev[64];
  bool host_ready;
  void* tab_host;     // IpTab<P>
  RefElem1D ref;
  int64_t ndofs;
};

// inverse of the (SPD, <= 4 x 4) 1-D reference mass matrix by Gauss-Jordan
inline void invert_mass_1d(const RefElem1D& ref, int n, double out[kMaxN1][kMaxN1]) {
  double a[kMaxN1][2 * kMaxN1];
  for (int i = 0; i < n; ++i)
    for (int j = 0; j < n; ++j) {
      a[i][j] = ref.M[i][j];
      a[i][n + j] = i == j ? 1.0 : 0.0;
    }
  for (int c = 0; c < n; ++c) {
    int piv = c;
    for (int r = c + 1; r < n; ++r)
      if (std::fabs(a[r][c]) > std::fabs(a[piv][c])) piv = r;
    for (int j = 0; j < 2 * n; ++j) std::swap(a[c][j], a[piv][j]);
    const double d = a[c][c];
    for (int j = 0; j < 2 * n; ++j) a[c][j] /= d;
    for (int r = 0; r < n; ++r)
      if (r != c) {
        const double f = a[r][c];
        for (int j = 0; j < 2 * n; ++j) a[r][j] -= f * a[c][j];
      }
  }
  for (int i = 0; i < kMaxN1; ++i)
    for (int j = 0; j < kMaxN1; ++j) out[i][j] = (i < n && j < n) ? a[i][n + j] : 0.0;
}

}  // namespace stefan
