// Local-DG element-block path: the reference's LDG assembly with ARBITRARY per-cell and per-face quadratures
// (cut cells, merged cells, curved interfaces), and the application of the assembled blocks.  The LDG twin of
// blocks.cu; same data layout (`stefan_blocks_data`), 3 dofs per node interleaved (T, q_x, q_y), so a block is
// 3 NF x 3 NF.
//
// Cell terms (per solution cell, its own volume points):
//   divergence_operator   LDG_div_operator.jl:1-11        A[q_d a][T b]   += sum (d_d V_a) V_b detjac w
//   mass_matrix           LDG_mass_operator.jl:1-25       A[q_d a][q_d b] += sum V_a V_b detjac w
//   gradient_operator     LDG_gradient_operator.jl:1-11   A[T a][q_d b]   += k sum (d_d V_a) V_b detjac w
// Face slot towards a neighbour (interelement or interface; the reference's drivers call the same face functions
// for both, LDG_scalar_flux_operator.jl:111-309, LDG_vector_flux_operator.jl:158-386), seen from the cell (side 1),
// per point with normal n, weight w (x scale area), bp = 1/2 sign(V0.n) (edge_switch, LDG_utilities.jl:1-4),
// bn = n.V0 (V0 passed as beta), alpha = slot_pen:
//   assemble_face_scalar_flux_operator!  LDG_scalar_flux_operator.jl:25-107
//       diag [q_d a][T b] += (-1/2 - bp) n_d V1_a V1_b w        off [q_d a][T b] += (-1/2 + bp) n_d V1_a V2_b w
//   assemble_face_vector_flux_operator!  LDG_vector_flux_operator.jl:47-154
//       diag [T a][q_d b] += k1 (-1/2 + bn) n_d V1_a V1_b w     off [T a][q_d b] += k2 (-1/2 - bn) n_d V1_a V2_b w
//       diag [T a][T b]   += alpha (n.n) V1_a V1_b w            off [T a][T b]   -= alpha (n.n) V1_a V2_b w
// Boundary slot (assemble_boundary_face_vector_flux_operator!, LDG_vector_flux_operator.jl:391-439):
//       diag [T a][q_d b] += -k1 n_d V1_a V1_b w,  diag [T a][T b] += alpha (n.n) V1_a V1_b w
// Right-hand side (assemble_LDG_rhs!, LDG_assembly.jl:93-131; LDG_source_term.jl): T rows sum V f detjac w; on
// boundary slots q rows sum n_d V g w and T rows + alpha (n.n) sum V g w.
#include "../../include/stefan_b200.h"
#include "blocks_handle.cuh"
#include "common.cuh"
#include "krylov.cuh"
#include "ldg_schur.cuh"

#include <algorithm>
#include <cmath>

namespace stefan {

struct LdgBlocksHandle {
  stefan_blocks_dims d;
  int nf, n3;
  double* blocks;     // [ncells][1 + nslots][n3][n3], row-major (row = test function)
  int32_t* slot_nbr;  // [ncells][nslots]
  double* minv;       // [ncells][nf][nf]: inverse of every cell's mass sub-block (built by assemble)
  bool assembled;
};

constexpr int kLbWarps = 4;
template <int P>
struct LbSmem {
  static constexpr int NF = (P + 1) * (P + 1);
  static constexpr int MAXQ = P == 3 ? 16 : 32;
  // per warp: three tables [MAXQ][NF] + four per-point scalars [MAXQ]
  static constexpr int WARP_DOUBLES = 3 * MAXQ * NF + 4 * MAXQ;
};

// One warp per solution cell.  Per batch of points the lanes tabulate basis values / gradients in shared memory,
// then every lane works through its entries (stride 32) of the 3 NF x 3 NF block as dot products over the batch
// and adds them to the block in global memory (an entry is owned by one lane: no atomics).
template <int P>
__global__ void __launch_bounds__(kLbWarps * 32)
ldgblocks_assemble_kernel(const stefan_blocks_dims d, const stefan_blocks_data in, double v0x, double v0y,
                          double* __restrict__ blocks) {
  constexpr int NF = (P + 1) * (P + 1), N3 = 3 * NF, MAXQ = LbSmem<P>::MAXQ;
  __shared__ double smem[kLbWarps * LbSmem<P>::WARP_DOUBLES];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  double* const T0 = smem + (size_t)warp * LbSmem<P>::WARP_DOUBLES;  // [MAXQ][NF]
  double* const T1 = T0 + MAXQ * NF;
  double* const T2 = T1 + MAXQ * NF;
  double* const S0 = T2 + MAXQ * NF;  // [MAXQ] per-point scalars
  double* const S1 = S0 + MAXQ;
  double* const S2 = S1 + MAXQ;
  double* const S3 = S2 + MAXQ;
  const int nb = 1 + d.nslots;
  for (int cell = blockIdx.x * kLbWarps + warp; cell < d.ncells; cell += gridDim.x * kLbWarps) {
    const double jx = in.cell_jac[2 * cell], jy = in.cell_jac[2 * cell + 1], k1 = in.cell_k[cell];
    double* const out = blocks + (size_t)cell * nb * N3 * N3;
    for (int e = lane; e < nb * N3 * N3; e += 32) out[e] = 0.0;
    __syncwarp();
    // ---- volume: T0 = V, T1 = dV/dx, T2 = dV/dy, S0 = w detjac ----
    for (int q0 = 0; q0 < d.nqv; q0 += MAXQ) {
      const int nq = min(MAXQ, d.nqv - q0);
      for (int t = lane; t < nq * NF; t += 32) {
        const int q = t / NF, a = t % NF;
        const size_t g = (size_t)cell * d.nqv + q0 + q;
        double v, gx, gy;
        basis_vg<P>(a, in.vol_pts[2 * g], in.vol_pts[2 * g + 1], jx, jy, v, gx, gy);
        T0[q * NF + a] = v;
        T1[q * NF + a] = gx;
        T2[q * NF + a] = gy;
        if (a == 0) S0[q] = in.vol_wts[g] * jx * jy;
      }
      __syncwarp();
      for (int e = lane; e < N3 * N3; e += 32) {
        const int r = e / N3, c = e % N3, a = r / 3, i = r % 3, b = c / 3, j = c % 3;
        const double* L = nullptr;  // left factor table (indexed by a)
        double f = 1.0;
        if (i >= 1 && j == 0) L = i == 1 ? T1 : T2;                    // divergence
        else if (i >= 1 && j == i) L = T0;                            // mass
        else if (i == 0 && j >= 1) { L = j == 1 ? T1 : T2; f = k1; }  // gradient
        if (L == nullptr) continue;
        double s = 0.0;
        for (int q = 0; q < nq; ++q) s += L[q * NF + a] * T0[q * NF + b] * S0[q];
        out[e] += f * s;
      }
      __syncwarp();
    }
    // ---- face slots: T0 = V1, T1 = V2; S0 = w, S1 = n_x, S2 = n_y ----
    for (int sidx = 0; sidx < d.nslots; ++sidx) {
      const size_t cs = (size_t)cell * d.nslots + sidx;
      const int nbr = in.slot_nbr[cs];
      if (nbr < -1) continue;
      const double alpha = in.slot_pen[cs];
      double jx2 = jx, jy2 = jy, k2 = k1;
      if (nbr >= 0) {
        jx2 = in.cell_jac[2 * nbr];
        jy2 = in.cell_jac[2 * nbr + 1];
        k2 = in.cell_k[nbr];
      }
      double* const off = out + (size_t)(1 + sidx) * N3 * N3;
      for (int q0 = 0; q0 < d.nqf; q0 += MAXQ) {
        const int nq = min(MAXQ, d.nqf - q0);
        for (int t = lane; t < nq * NF; t += 32) {
          const int q = t / NF, a = t % NF;
          const size_t fq = cs * d.nqf + q0 + q;
          double v, gx, gy;
          basis_vg<P>(a, in.face_pts[2 * fq], in.face_pts[2 * fq + 1], jx, jy, v, gx, gy);
          T0[q * NF + a] = v;
          if (nbr >= 0) {
            basis_vg<P>(a, in.face_nbr_pts[2 * fq], in.face_nbr_pts[2 * fq + 1], jx2, jy2, v, gx, gy);
            T1[q * NF + a] = v;
          }
          if (a == 0) {
            S0[q] = in.face_wts[fq];
            S1[q] = in.face_normals[2 * fq];
            S2[q] = in.face_normals[2 * fq + 1];
            S3[q] = 0.0;
          }
        }
        __syncwarp();
        for (int e = lane; e < N3 * N3; e += 32) {
          const int r = e / N3, c = e % N3, a = r / 3, i = r % 3, b = c / 3, j = c % 3;
          if (!((i >= 1 && j == 0) || (i == 0))) continue;  // q rows couple to T only; T rows to everything
          double dsum = 0.0, osum = 0.0;
          for (int q = 0; q < nq; ++q) {
            const double w = S0[q];
            if (w == 0.0) continue;
            const double nx = S1[q], ny = S2[q];
            const double v0n = v0x * nx + v0y * ny;
            const double nn = nx * nx + ny * ny;
            const double bp = (v0n > 0.0 ? 0.5 : (v0n < 0.0 ? -0.5 : 0.0)) * nn;  // edge_switch(V0, n) . n
            const double v1a = T0[q * NF + a], v1b = T0[q * NF + b];
            if (nbr >= 0) {
              const double v2b = T1[q * NF + b];
              if (i >= 1) {  // scalar flux, rows q_d
                const double nd = i == 1 ? nx : ny;
                dsum += (-0.5 - bp) * nd * v1a * v1b * w;
                osum += (-0.5 + bp) * nd * v1a * v2b * w;
              } else if (j >= 1) {  // vector flux, rows T, columns q_d
                const double nd = j == 1 ? nx : ny;
                dsum += k1 * (-0.5 + v0n) * nd * v1a * v1b * w;
                osum += k2 * (-0.5 - v0n) * nd * v1a * v2b * w;
              } else {  // penalty, rows T, columns T
                dsum += alpha * nn * v1a * v1b * w;
                osum -= alpha * nn * v1a * v2b * w;
              }
            } else if (i == 0) {  // boundary slot
              if (j >= 1) dsum += -k1 * (j == 1 ? nx : ny) * v1a * v1b * w;
              else dsum += alpha * nn * v1a * v1b * w;
            }
          }
          out[e] += dsum;
          if (nbr >= 0) off[e] += osum;
        }
        __syncwarp();
      }
    }
  }
}

// y = A x: G lanes per row (cell, r), coalesced block-row reads, shuffle reduction
template <int N3, int G>
__global__ void __launch_bounds__(256)
ldgblocks_apply_kernel(int ncells, int nslots, const double* __restrict__ blocks, const int32_t* __restrict__ slot_nbr,
                       const double* __restrict__ x, double* __restrict__ y) {
  const int64_t nrows = (int64_t)ncells * N3;
  const int nb = 1 + nslots;
  const int gl = threadIdx.x % G;
  const int64_t ngroups = (int64_t)gridDim.x * blockDim.x / G;
  const int64_t nrows_pad = (nrows + (32 / G) - 1) / (32 / G) * (32 / G);  // whole warps stay in the loop
  for (int64_t r = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) / G; r < nrows_pad; r += ngroups) {
    double v = 0.0;
    if (r < nrows) {
      const int cell = (int)(r / N3), a = (int)(r % N3);
      const double* B = blocks + ((size_t)cell * nb) * N3 * N3 + (size_t)a * N3;
      const double* xc = x + (size_t)cell * N3;
      for (int b = gl; b < N3; b += G) v += B[b] * xc[b];
      for (int s = 0; s < nslots; ++s) {
        const int nbr = slot_nbr[(size_t)cell * nslots + s];
        if (nbr < 0) continue;
        const double* Bo = B + (size_t)(1 + s) * N3 * N3;
        const double* xn = x + (size_t)nbr * N3;
        for (int b = gl; b < N3; b += G) v += Bo[b] * xn[b];
      }
    }
#pragma unroll
    for (int o = G / 2; o > 0; o >>= 1) v += __shfl_down_sync(0xffffffffu, v, o, G);
    if (gl == 0 && r < nrows) y[r] = v;
  }
}

// one warp per cell; lane a (strided) owns node a
template <int P>
__global__ void __launch_bounds__(128)
ldgblocks_rhs_kernel(const stefan_blocks_dims d, const stefan_blocks_data in, const double* __restrict__ f_qp,
                     const double* __restrict__ g_qp, double* __restrict__ rhs) {
  constexpr int NF = (P + 1) * (P + 1);
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, wpb = blockDim.x >> 5;
  for (int cell = blockIdx.x * wpb + warp; cell < d.ncells; cell += gridDim.x * wpb) {
    const double jx = in.cell_jac[2 * cell], jy = in.cell_jac[2 * cell + 1];
    for (int a = lane; a < NF; a += 32) {
      double rT = 0.0, rx = 0.0, ry = 0.0;
      if (f_qp != nullptr)
        for (int q = 0; q < d.nqv; ++q) {
          const size_t g = (size_t)cell * d.nqv + q;
          const double w = in.vol_wts[g];
          if (w == 0.0) continue;
          double v, gx, gy;
          basis_vg<P>(a, in.vol_pts[2 * g], in.vol_pts[2 * g + 1], jx, jy, v, gx, gy);
          rT += v * f_qp[g] * w * jx * jy;
        }
      if (g_qp != nullptr)
        for (int s = 0; s < d.nslots; ++s) {
          const size_t cs = (size_t)cell * d.nslots + s;
          if (in.slot_nbr[cs] != -1) continue;
          const double alpha = in.slot_pen[cs];
          for (int q = 0; q < d.nqf; ++q) {
            const size_t fq = cs * d.nqf + q;
            const double w = in.face_wts[fq];
            if (w == 0.0) continue;
            double v, gx, gy;
            basis_vg<P>(a, in.face_pts[2 * fq], in.face_pts[2 * fq + 1], jx, jy, v, gx, gy);
            const double nx = in.face_normals[2 * fq], ny = in.face_normals[2 * fq + 1], gd = g_qp[fq];
            rx += nx * v * gd * w;                            // boundary_face_scalar_flux, LDG_source_term.jl:64-89
            ry += ny * v * gd * w;
            rT += alpha * (nx * nx + ny * ny) * v * gd * w;   // -alpha (-n.n) linear_form, LDG_source_term.jl:203-321
          }
        }
      double* o = rhs + ((size_t)cell * NF + a) * 3;
      o[0] = rT;
      o[1] = rx;
      o[2] = ry;
    }
  }
}

// inverse of every cell's NF x NF mass sub-block (entries [3 a + 1][3 b + 1] of the diagonal block); Gauss-Jordan
// with partial pivoting, one thread per cell; a singular block (empty solution cell) gets the identity
__global__ void __launch_bounds__(64)
ldgblocks_invert_mass_kernel(int ncells, int nslots, int nf, const double* __restrict__ blocks, double* __restrict__ minv) {
  constexpr int MAXNF = 16;
  const int n3 = 3 * nf;
  for (int cell = blockIdx.x * blockDim.x + threadIdx.x; cell < ncells; cell += gridDim.x * blockDim.x) {
    const double* D = blocks + (size_t)cell * (1 + nslots) * n3 * n3;
    double a[MAXNF][2 * MAXNF];
    for (int i = 0; i < nf; ++i)
      for (int j = 0; j < nf; ++j) {
        a[i][j] = D[(size_t)(3 * i + 1) * n3 + (3 * j + 1)];
        a[i][nf + j] = i == j ? 1.0 : 0.0;
      }
    bool ok = true;
    for (int c = 0; c < nf && ok; ++c) {
      int piv = c;
      for (int r = c + 1; r < nf; ++r)
        if (fabs(a[r][c]) > fabs(a[piv][c])) piv = r;
      if (a[piv][c] == 0.0) { ok = false; break; }
      for (int j = 0; j < 2 * nf; ++j) { const double t = a[c][j]; a[c][j] = a[piv][j]; a[piv][j] = t; }
      const double dd = 1.0 / a[c][c];
      for (int j = 0; j < 2 * nf; ++j) a[c][j] *= dd;
      for (int r = 0; r < nf; ++r)
        if (r != c) {
          const double f = a[r][c];
          if (f != 0.0)
            for (int j = 0; j < 2 * nf; ++j) a[r][j] -= f * a[c][j];
        }
    }
    double* o = minv + (size_t)cell * nf * nf;
    for (int i = 0; i < nf; ++i)
      for (int j = 0; j < nf; ++j) o[i * nf + j] = ok ? a[i][nf + j] : (i == j ? 1.0 : 0.0);
  }
}

// Per cell: out_q = M^-1 (u_q - v_q) (v may be null), out_T = Tsrc ? Tsrc : 0.  One thread per (cell, node).
__global__ void __launch_bounds__(128)
ldgblocks_mass_solve_kernel(int64_t nnodes, int nf, const double* __restrict__ minv, const double* __restrict__ u,
                            const double* __restrict__ v, const double* __restrict__ Tsrc, double* __restrict__ out) {
  for (int64_t g = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; g < nnodes; g += (int64_t)gridDim.x * blockDim.x) {
    const int64_t cell = g / nf;
    const int a = (int)(g % nf);
    const double* M = minv + (size_t)cell * nf * nf + (size_t)a * nf;
    const int64_t n0 = cell * nf;
    double sx = 0.0, sy = 0.0;
    for (int b = 0; b < nf; ++b) {
      const double rx = u[3 * (n0 + b) + 1] - (v ? v[3 * (n0 + b) + 1] : 0.0);
      const double ry = u[3 * (n0 + b) + 2] - (v ? v[3 * (n0 + b) + 2] : 0.0);
      sx += M[b] * rx;
      sy += M[b] * ry;
    }
    out[3 * g] = Tsrc ? Tsrc[g] : 0.0;
    out[3 * g + 1] = sx;
    out[3 * g + 2] = sy;
  }
}

}  // namespace stefan

using namespace stefan;

extern "C" {

int stefan_ldgblocks_create(const stefan_blocks_dims* dims, stefan_ldgblocks** out) {
  STEFAN_REQUIRE(dims && out, "stefan_ldgblocks_create: null argument");
  STEFAN_REQUIRE(dims->ncells >= 1 && dims->order >= 1 && dims->order <= 3, "stefan_ldgblocks_create: bad ncells / order");
  STEFAN_REQUIRE(dims->nqv >= 1 && dims->nslots >= 1 && dims->nslots <= 16 && dims->nqf >= 1,
                 "stefan_ldgblocks_create: need nqv >= 1, 1 <= nslots <= 16, nqf >= 1");
  LdgBlocksHandle* h = new LdgBlocksHandle();
  h->d = *dims;
  h->nf = (dims->order + 1) * (dims->order + 1);
  h->n3 = 3 * h->nf;
  h->blocks = nullptr;
  h->slot_nbr = nullptr;
  h->minv = nullptr;
  h->assembled = false;
  const size_t nbytes = (size_t)dims->ncells * (1 + dims->nslots) * h->n3 * h->n3 * sizeof(double);
  cudaError_t e = cudaMalloc(&h->blocks, nbytes);
  if (e == cudaSuccess) e = cudaMalloc(&h->slot_nbr, (size_t)dims->ncells * dims->nslots * sizeof(int32_t));
  if (e == cudaSuccess) e = cudaMalloc(&h->minv, (size_t)dims->ncells * h->nf * h->nf * sizeof(double));
  if (e != cudaSuccess) {
    if (h->blocks) cudaFree(h->blocks);
    if (h->slot_nbr) cudaFree(h->slot_nbr);
    delete h;
    return fail(ST_CUDA_ERROR, "stefan_ldgblocks_create: %s", cudaGetErrorString(e));
  }
  *out = reinterpret_cast<stefan_ldgblocks*>(h);
  return ST_OK;
}

int stefan_ldgblocks_destroy(stefan_ldgblocks* hv) {
  LdgBlocksHandle* h = reinterpret_cast<LdgBlocksHandle*>(hv);
  if (!h) return ST_OK;
  if (h->blocks) cudaFree(h->blocks);
  if (h->slot_nbr) cudaFree(h->slot_nbr);
  if (h->minv) cudaFree(h->minv);
  delete h;
  return ST_OK;
}

int64_t stefan_ldgblocks_ndofs(const stefan_ldgblocks* hv) {
  const LdgBlocksHandle* h = reinterpret_cast<const LdgBlocksHandle*>(hv);
  return h ? (int64_t)h->d.ncells * h->n3 : -1;
}

int stefan_ldgblocks_assemble(stefan_ldgblocks* hv, const stefan_blocks_data* in, double V0x, double V0y, void* stream) {
  LdgBlocksHandle* h = reinterpret_cast<LdgBlocksHandle*>(hv);
  STEFAN_REQUIRE(h && in, "stefan_ldgblocks_assemble: null argument");
  STEFAN_REQUIRE(in->vol_pts && in->vol_wts && in->cell_jac && in->cell_k && in->slot_nbr && in->face_pts &&
                 in->face_nbr_pts && in->face_wts && in->face_normals && in->slot_pen,
                 "stefan_ldgblocks_assemble: every array of stefan_blocks_data except slot_lambda is required");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  STEFAN_CUDA(cudaMemcpyAsync(h->slot_nbr, in->slot_nbr, (size_t)h->d.ncells * h->d.nslots * sizeof(int32_t),
                              cudaMemcpyDeviceToDevice, st));
  const int grid = (int)std::min<int64_t>(((int64_t)h->d.ncells + kLbWarps - 1) / kLbWarps, (int64_t)sm_count() * 16);
  switch (h->d.order) {
    case 1: ldgblocks_assemble_kernel<1><<<grid, kLbWarps * 32, 0, st>>>(h->d, *in, V0x, V0y, h->blocks); break;
    case 2: ldgblocks_assemble_kernel<2><<<grid, kLbWarps * 32, 0, st>>>(h->d, *in, V0x, V0y, h->blocks); break;
    case 3: ldgblocks_assemble_kernel<3><<<grid, kLbWarps * 32, 0, st>>>(h->d, *in, V0x, V0y, h->blocks); break;
  }
  STEFAN_CUDA(cudaGetLastError());
  ldgblocks_invert_mass_kernel<<<(h->d.ncells + 63) / 64, 64, 0, st>>>(h->d.ncells, h->d.nslots, h->nf, h->blocks, h->minv);
  STEFAN_CUDA(cudaGetLastError());
  h->assembled = true;
  return ST_OK;
}

int stefan_ldgblocks_apply(stefan_ldgblocks* hv, const double* x, double* y, void* stream) {
  LdgBlocksHandle* h = reinterpret_cast<LdgBlocksHandle*>(hv);
  STEFAN_REQUIRE(h && x && y && x != y, "stefan_ldgblocks_apply: bad argument");
  STEFAN_REQUIRE(h->assembled, "stefan_ldgblocks_apply: call stefan_ldgblocks_assemble first");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const int64_t nrows = (int64_t)h->d.ncells * h->n3;
  switch (h->d.order) {
    case 1: {
      const int grid = (int)std::min<int64_t>((nrows * 4 + 255) / 256, (int64_t)sm_count() * 32);
      ldgblocks_apply_kernel<12, 4><<<grid, 256, 0, st>>>(h->d.ncells, h->d.nslots, h->blocks, h->slot_nbr, x, y);
      break;
    }
    case 2: {
      const int grid = (int)std::min<int64_t>((nrows * 8 + 255) / 256, (int64_t)sm_count() * 32);
      ldgblocks_apply_kernel<27, 8><<<grid, 256, 0, st>>>(h->d.ncells, h->d.nslots, h->blocks, h->slot_nbr, x, y);
      break;
    }
    case 3: {
      const int grid = (int)std::min<int64_t>((nrows * 16 + 255) / 256, (int64_t)sm_count() * 32);
      ldgblocks_apply_kernel<48, 16><<<grid, 256, 0, st>>>(h->d.ncells, h->d.nslots, h->blocks, h->slot_nbr, x, y);
      break;
    }
  }
  STEFAN_CUDA(cudaGetLastError());
  return ST_OK;
}

int stefan_ldgblocks_rhs(stefan_ldgblocks* hv, const stefan_blocks_data* in, const double* f_qp, const double* g_qp,
                         double* rhs, void* stream) {
  LdgBlocksHandle* h = reinterpret_cast<LdgBlocksHandle*>(hv);
  STEFAN_REQUIRE(h && in && rhs, "stefan_ldgblocks_rhs: null argument");
  STEFAN_REQUIRE(in->vol_pts && in->vol_wts && in->cell_jac && in->slot_nbr && in->face_pts && in->face_wts &&
                 in->face_normals && in->slot_pen, "stefan_ldgblocks_rhs: incomplete stefan_blocks_data");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const int grid = (int)std::min<int64_t>(((int64_t)h->d.ncells + 3) / 4, (int64_t)sm_count() * 16);
  switch (h->d.order) {
    case 1: ldgblocks_rhs_kernel<1><<<grid, 128, 0, st>>>(h->d, *in, f_qp, g_qp, rhs); break;
    case 2: ldgblocks_rhs_kernel<2><<<grid, 128, 0, st>>>(h->d, *in, f_qp, g_qp, rhs); break;
    case 3: ldgblocks_rhs_kernel<3><<<grid, 128, 0, st>>>(h->d, *in, f_qp, g_qp, rhs); break;
  }
  STEFAN_CUDA(cudaGetLastError());
  return ST_OK;
}

int64_t stefan_ldgblocks_coo_size(const stefan_ldgblocks* hv) {
  const LdgBlocksHandle* h = reinterpret_cast<const LdgBlocksHandle*>(hv);
  return h ? (int64_t)h->d.ncells * (1 + h->d.nslots) * h->n3 * h->n3 : -1;
}

int stefan_ldgblocks_assemble_coo(stefan_ldgblocks* hv, int64_t* rows, int64_t* cols, double* vals, int64_t capacity,
                                  int32_t index_base, int64_t* nnz_out, void* stream) {
  LdgBlocksHandle* h = reinterpret_cast<LdgBlocksHandle*>(hv);
  STEFAN_REQUIRE(h && rows && cols && vals && nnz_out, "stefan_ldgblocks_assemble_coo: null argument");
  STEFAN_REQUIRE(h->assembled, "stefan_ldgblocks_assemble_coo: call stefan_ldgblocks_assemble first");
  STEFAN_REQUIRE(index_base == 0 || index_base == 1, "stefan_ldgblocks_assemble_coo: index_base must be 0 or 1");
  *nnz_out = stefan_ldgblocks_coo_size(hv);
  if (*nnz_out > capacity)
    return fail(ST_CAPACITY, "stefan_ldgblocks_assemble_coo: %lld triplets, capacity %lld", (long long)*nnz_out,
                (long long)capacity);
  const int grid = (int)std::min<int64_t>((*nnz_out + 255) / 256, (int64_t)sm_count() * 32);
  blocks_coo_kernel<<<grid, 256, 0, static_cast<cudaStream_t>(stream)>>>(h->d.ncells, h->d.nslots, h->n3, h->blocks,
                                                                         h->slot_nbr, index_base, rows, cols, vals);
  STEFAN_CUDA(cudaGetLastError());
  return ST_OK;
}

int stefan_ldgblocks_solve(stefan_ldgblocks* hv, const double* rhs, double* sol, const stefan_cg_params* prm,
                           stefan_cg_result* res, void* stream) {
  LdgBlocksHandle* h = reinterpret_cast<LdgBlocksHandle*>(hv);
  STEFAN_REQUIRE(h && rhs && sol && prm && res, "stefan_ldgblocks_solve: null argument");
  STEFAN_REQUIRE(h->assembled, "stefan_ldgblocks_solve: call stefan_ldgblocks_assemble first");
  STEFAN_REQUIRE(prm->max_iterations >= 0 && prm->rel_tolerance >= 0.0, "stefan_ldgblocks_solve: bad parameters");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const int64_t nn = (int64_t)h->d.ncells * h->nf;
  const int mblk = (int)std::min<int64_t>((nn + 127) / 128, (int64_t)sm_count() * 16);
  auto apply = [&](const double* w, double* y) -> int { return stefan_ldgblocks_apply(hv, w, y, st); };
  auto mass_solve = [&](const double* u, const double* v, const double* Tsrc, double* out) {
    ldgblocks_mass_solve_kernel<<<mblk, 128, 0, st>>>(nn, h->nf, h->minv, u, v, Tsrc, out);
  };
  return ldg_schur_bicgstab(apply, mass_solve, nn, rhs, sol, prm, res, st);
}

}  // extern "C"
