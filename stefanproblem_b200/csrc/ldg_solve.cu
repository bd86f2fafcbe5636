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
#include "ldg_schur.cuh"

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
  const int64_t ncells = (int64_t)h->nx * h->ny, nn = ncells * nf;
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

  const int cblk = (int)std::min<int64_t>((ncells + 127) / 128, (int64_t)sm_count() * 8);
  auto apply = [&](const double* w, double* y) -> int { return stefan_ldg_apply(hv, w, y, nullptr, nullptr, st); };
  auto mass_solve = [&](const double* u, const double* v, const double* Tsrc, double* out) {
    ldg_mass_solve_kernel<<<cblk, 128, 0, st>>>(mt, u, v, Tsrc, out, ncells);
  };
  return ldg_schur_bicgstab(apply, mass_solve, nn, rhs, sol, prm, res, st);
}

}  // extern "C"
