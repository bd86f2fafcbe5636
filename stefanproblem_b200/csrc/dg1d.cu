// 1-D two-phase interior-penalty DG with a Robin interface (StefanRobinIC/DG1D):
// matrix-free application and right-hand side for sm_100a.
//
//   gradient_operator                 assemble_gradient_operator.jl:1-9    (k/j) sum w N' N''
//   assemble_edge_flux_operator!      assemble_flux_operator.jl:7-59       every interior edge,
//                                                                          incl. the interface
//   assemble_edge_penalty_operator!   assemble_penalty_operator.jl:7-50
//   assemble_robin_edge_operator!     assemble_interface_robin_operator.jl:1-44  (labels differ)
//   boundary flux / penalty           assemble_flux_operator.jl:122-173, assemble_penalty_operator.jl:72-102
//   rhs: two-phase source, Robin source, boundary rhs
//        assemble_source_term.jl:26-61, assemble_interface_robin_source.jl:1-43,
//        assemble_boundary_conditions.jl:1-44
// The mesh is DGMesh1D (dgmesh_1d.jl:1-51): ne1 uniform cells left of the interface
// point (label 1), ne2 right of it (label 2), consecutive node ids per cell.
// One thread per cell; the problem sizes of the reference are tiny (4..64 cells), the
// kernel is written for any size.
#include "../../include/stefan_b200.h"
#include "common.cuh"
#include "refelem.h"

#include <algorithm>

namespace stefan {

enum Dg1dTerms : int {
  DG1D_GRADIENT = 1, DG1D_FLUX = 2, DG1D_PENALTY = 4, DG1D_ROBIN = 8,
  DG1D_BOUNDARY_FLUX = 16, DG1D_BOUNDARY_PENALTY = 32,
};

struct Dg1dTab {
  int n1, nq, ne1, ncells, terms;
  double j1, j2, k1, k2, penalty, lambda, Tm;
  double K[kMaxN1][kMaxN1];
  double Nq[kMaxQ][kMaxN1], w[kMaxQ];
  double dNL[kMaxN1], dNR[kMaxN1];
};

struct Dg1dHandle {
  Dg1dTab tab;
};

// rows of cell c; w / wl / wr: the dofs of the cell and of its left / right neighbour
// (wl, wr are not read where there is no neighbour)
__device__ __forceinline__ void dg1d_cell_rows(const Dg1dTab& t, int c, const double* __restrict__ w,
                                               const double* __restrict__ wl, const double* __restrict__ wr,
                                               double* __restrict__ out) {
  const int n1 = t.n1, P = n1 - 1;
  const int lab = c < t.ne1 ? 1 : 2;
  const double j = lab == 1 ? t.j1 : t.j2, k = lab == 1 ? t.k1 : t.k2;
  for (int a = 0; a < n1; ++a) {
    double v = 0.0;
    if (t.terms & DG1D_GRADIENT)
      for (int b = 0; b < n1; ++b) v += (k / j) * t.K[a][b] * w[b];
    out[a] = v;
  }
  for (int side = 0; side < 2; ++side) {
    const int cn = side == 0 ? c - 1 : c + 1;
    const int fnode = side == 0 ? 0 : P;  // own face node; the neighbour's is the opposite end
    const double trK = w[fnode];
    double dK = 0.0;  // own outward normal derivative at the face
    for (int b = 0; b < n1; ++b) dK += (side == 0 ? -t.dNL[b] : t.dNR[b]) / j * w[b];
    double e_term = 0.0, g_coef = 0.0;  // contribution = e * e_term + g * g_coef
    if (cn < 0 || cn >= t.ncells) {
      if (t.terms & DG1D_BOUNDARY_FLUX) {
        e_term += -k * dK;
        g_coef += -k * trK;
      }
      if (t.terms & DG1D_BOUNDARY_PENALTY) e_term += t.penalty * trK;
    } else {
      const int labn = cn < t.ne1 ? 1 : 2;
      const double jn = labn == 1 ? t.j1 : t.j2, kn = labn == 1 ? t.k1 : t.k2;
      const double* wn = side == 0 ? wl : wr;
      const double trN = wn[side == 0 ? P : 0];
      double dN = 0.0;  // neighbour's derivative along MY outward normal, at the shared point
      for (int b = 0; b < n1; ++b) dN += (side == 0 ? -t.dNR[b] : t.dNL[b]) / jn * wn[b];
      if (t.terms & DG1D_FLUX) {
        e_term += -0.5 * k * dK - 0.5 * kn * dN;
        g_coef += -0.5 * k * (trK - trN);
      }
      if (t.terms & DG1D_PENALTY) e_term += t.penalty * (trK - trN);
      if ((t.terms & DG1D_ROBIN) && labn != lab) e_term += 0.25 * t.lambda * (trK + trN);
    }
    out[fnode] += e_term;
    for (int a = 0; a < n1; ++a) out[a] += (side == 0 ? -t.dNL[a] : t.dNR[a]) / j * g_coef;
  }
}

__global__ void dg1d_apply_kernel(const __grid_constant__ Dg1dTab t, const double* __restrict__ x,
                                  double* __restrict__ y) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= t.ncells) return;
  const int n1 = t.n1;
  const double* w = x + (size_t)c * n1;
  double out[kMaxN1];
  dg1d_cell_rows(t, c, w, w - n1, w + n1, out);
  for (int a = 0; a < n1; ++a) y[(size_t)c * n1 + a] = out[a];
}

// The COO seam: blocks (cell, which = 0 diag / 1 left / 2 right) of n1 x n1 triplets,
// existing neighbours only, cell by cell, column-major inside a block; every column is
// the apply row function probed with a unit vector.  One thread per (cell, which, column).
__global__ void dg1d_coo_kernel(const __grid_constant__ Dg1dTab t, int64_t index_base, int64_t* __restrict__ rows,
                                int64_t* __restrict__ cols, double* __restrict__ vals) {
  const int n1 = t.n1;
  const int64_t nwork = (int64_t)t.ncells * 3 * n1;
  for (int64_t wk = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; wk < nwork; wk += (int64_t)gridDim.x * blockDim.x) {
    const int bcol = (int)(wk % n1);
    const int64_t cw = wk / n1;
    const int c = (int)(cw / 3), which = (int)(cw % 3);
    if ((which == 1 && c == 0) || (which == 2 && c == t.ncells - 1)) continue;
    // blocks before cell c: 3 per cell minus the missing left of cell 0 and (only counted
    // for the last cell itself) nothing else; within the cell: diag, [left], [right]
    const int64_t nb = (int64_t)c * 3 - (c > 0 ? 1 : 0) + (which > 0 ? 1 : 0) + (which > 1 && c > 0 ? 1 : 0);
    const int64_t o = nb * n1 * n1 + (int64_t)bcol * n1;
    double U[3][kMaxN1], out[kMaxN1];
    for (int q = 0; q < 3; ++q)
      for (int k = 0; k < kMaxN1; ++k) U[q][k] = 0.0;
    U[which][bcol] = 1.0;
    dg1d_cell_rows(t, c, U[0], U[1], U[2], out);
    const int64_t cn = which == 0 ? c : which == 1 ? c - 1 : c + 1;
    for (int a = 0; a < n1; ++a) {
      rows[o + a] = (int64_t)c * n1 + a + index_base;
      cols[o + a] = cn * n1 + bcol + index_base;
      vals[o + a] = out[a];
    }
  }
}

__global__ void dg1d_rhs_kernel(const __grid_constant__ Dg1dTab t, const double* __restrict__ f_qp,
                                double TL, double TR, int with_boundary, double* __restrict__ b) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= t.ncells) return;
  const int n1 = t.n1, P = n1 - 1;
  const int lab = c < t.ne1 ? 1 : 2;
  const double j = lab == 1 ? t.j1 : t.j2, k = lab == 1 ? t.k1 : t.k2;
  double out[kMaxN1];
  for (int a = 0; a < n1; ++a) out[a] = 0.0;
  if (f_qp != nullptr)  // linear_form: sum V f j w
    for (int q = 0; q < t.nq; ++q)
      for (int a = 0; a < n1; ++a) out[a] += t.Nq[q][a] * f_qp[(size_t)c * t.nq + q] * j * t.w[q];
  if (t.terms & DG1D_ROBIN) {  // 0.5 lambda Tm basis(qp) on both sides of the label change
    if (c == t.ne1 - 1 && t.ncells > t.ne1) out[P] += 0.5 * t.lambda * t.Tm;
    if (c == t.ne1 && t.ne1 > 0) out[0] += 0.5 * t.lambda * t.Tm;
  }
  if (with_boundary) {  // pen T basis(qp) - T k n N'(qp)/j
    if (c == 0) {
      out[0] += t.penalty * TL;
      for (int a = 0; a < n1; ++a) out[a] += -TL * k * (-1.0) * t.dNL[a] / j;
    }
    if (c == t.ncells - 1) {
      out[P] += t.penalty * TR;
      for (int a = 0; a < n1; ++a) out[a] += -TR * k * (+1.0) * t.dNR[a] / j;
    }
  }
  for (int a = 0; a < n1; ++a) b[(size_t)c * n1 + a] = out[a];
}

}  // namespace stefan

using namespace stefan;

extern "C" {

int stefan_dg1d_create(const stefan_dg1d_params* p, stefan_dg1d** out) {
  STEFAN_REQUIRE(p && out, "stefan_dg1d_create: null argument");
  STEFAN_REQUIRE(p->nelmts1 >= 0 && p->nelmts2 >= 0 && p->nelmts1 + p->nelmts2 >= 1,
                 "stefan_dg1d_create: need at least one cell");
  STEFAN_REQUIRE(p->order >= 1 && p->order <= 3, "stefan_dg1d_create: order must be 1..3");
  STEFAN_REQUIRE(p->numqp >= 1 && p->numqp <= kMaxQ, "stefan_dg1d_create: numqp must be 1..%d", kMaxQ);
  STEFAN_REQUIRE(p->xL < p->xR && p->interfacepoint >= p->xL && p->interfacepoint <= p->xR,
                 "stefan_dg1d_create: need xL <= interfacepoint <= xR");
  RefElem1D r;
  build_ref1d(p->order, p->numqp, &r);
  Dg1dHandle* h = new Dg1dHandle();
  Dg1dTab& t = h->tab;
  std::memset(&t, 0, sizeof(t));
  t.n1 = p->order + 1;
  t.nq = p->numqp;
  t.ne1 = p->nelmts1;
  t.ncells = p->nelmts1 + p->nelmts2;
  t.terms = p->terms;
  t.j1 = p->nelmts1 > 0 ? 0.5 * (p->interfacepoint - p->xL) / p->nelmts1 : 1.0;
  t.j2 = p->nelmts2 > 0 ? 0.5 * (p->xR - p->interfacepoint) / p->nelmts2 : 1.0;
  t.k1 = p->k1;
  t.k2 = p->k2;
  t.penalty = p->penalty;
  t.lambda = p->lambda;
  t.Tm = p->Tm;
  for (int a = 0; a < t.n1; ++a) {
    for (int b = 0; b < t.n1; ++b) t.K[a][b] = r.K[a][b];
    t.dNL[a] = r.dNL[a];
    t.dNR[a] = r.dNR[a];
  }
  for (int q = 0; q < t.nq; ++q) {
    t.w[q] = r.qw[q];
    for (int a = 0; a < t.n1; ++a) t.Nq[q][a] = r.Nq[q][a];
  }
  *out = reinterpret_cast<stefan_dg1d*>(h);
  return ST_OK;
}

int stefan_dg1d_destroy(stefan_dg1d* hv) {
  delete reinterpret_cast<Dg1dHandle*>(hv);
  return ST_OK;
}

int64_t stefan_dg1d_ndofs(const stefan_dg1d* hv) {
  const Dg1dHandle* h = reinterpret_cast<const Dg1dHandle*>(hv);
  return h ? (int64_t)h->tab.ncells * h->tab.n1 : -1;
}

int stefan_dg1d_apply(stefan_dg1d* hv, const double* x, double* y, void* stream) {
  Dg1dHandle* h = reinterpret_cast<Dg1dHandle*>(hv);
  STEFAN_REQUIRE(h && x && y && x != y, "stefan_dg1d_apply: bad argument");
  dg1d_apply_kernel<<<(h->tab.ncells + 127) / 128, 128, 0, static_cast<cudaStream_t>(stream)>>>(h->tab, x, y);
  STEFAN_CUDA(cudaGetLastError());
  return ST_OK;
}

int64_t stefan_dg1d_coo_size(const stefan_dg1d* hv) {
  const Dg1dHandle* h = reinterpret_cast<const Dg1dHandle*>(hv);
  if (!h) return -1;
  const int64_t nc = h->tab.ncells, n1 = h->tab.n1;
  return n1 * n1 * (nc + 2 * (nc - 1));
}

int stefan_dg1d_assemble_coo(stefan_dg1d* hv, int64_t* rows, int64_t* cols, double* vals, int64_t capacity,
                             int32_t index_base, int64_t* nnz_out, void* stream) {
  Dg1dHandle* h = reinterpret_cast<Dg1dHandle*>(hv);
  STEFAN_REQUIRE(h && rows && cols && vals && nnz_out, "stefan_dg1d_assemble_coo: null argument");
  STEFAN_REQUIRE(index_base == 0 || index_base == 1, "stefan_dg1d_assemble_coo: index_base must be 0 or 1");
  *nnz_out = stefan_dg1d_coo_size(hv);
  if (*nnz_out > capacity)
    return fail(ST_CAPACITY, "stefan_dg1d_assemble_coo: %lld triplets, capacity %lld", (long long)*nnz_out,
                (long long)capacity);
  const int64_t nwork = (int64_t)h->tab.ncells * 3 * h->tab.n1;
  const int blocks = (int)std::min<int64_t>((nwork + 127) / 128, 4096);
  dg1d_coo_kernel<<<blocks, 128, 0, static_cast<cudaStream_t>(stream)>>>(h->tab, index_base, rows, cols, vals);
  STEFAN_CUDA(cudaGetLastError());
  return ST_OK;
}

int stefan_dg1d_rhs(stefan_dg1d* hv, const double* f_qp, double TL, double TR, int32_t with_boundary,
                    double* b, void* stream) {
  Dg1dHandle* h = reinterpret_cast<Dg1dHandle*>(hv);
  STEFAN_REQUIRE(h && b, "stefan_dg1d_rhs: null argument");
  dg1d_rhs_kernel<<<(h->tab.ncells + 127) / 128, 128, 0, static_cast<cudaStream_t>(stream)>>>(
      h->tab, f_qp, TL, TR, with_boundary, b);
  STEFAN_CUDA(cudaGetLastError());
  return ST_OK;
}

}  // extern "C"
