// Point evaluation of a DG solution and the front-velocity formulas that close the Stefan loop
// (SURVEY 8(f-1)), and the interface error norm.
//
//   stefan_interpolate_at_reference_points   InteriorPenalty.interpolate_at_reference_points, IP_postprocess.jl:30-52
//   stefan_gradient_at_reference_points      InteriorPenalty.gradient_at_reference_points,    IP_postprocess.jl:1-28
//   stefan_front_velocity                    robin_interface_velocity/IP_robin_interface_flux_velocity.jl:213-278:
//                                            velocity_s = -lambda (Tm - T_s), mean of the two, and the flux-jump
//                                            velocity k2 dnT2 - k1 dnT1
//   stefan_blocks_interface_l2_error         interface_L2_error / integral_norm_on_interface, useful_routines.jl:239-274
//                                            and :288-307, on the interface slots of the element-block data
// The evaluation kernels are handle-free gathers: a dof vector with NF = (order + 1)^2 dofs per (solution) cell,
// a list of reference points and the cell each lies in.  One thread per point; the NF dofs of a cell are one or
// two coalesced sectors, the points of one cell are adjacent in the reference's lists (closest points run along
// the interface), so neighbouring threads hit the same lines.
#include "blocks_handle.cuh"
#include "common.cuh"
#include "krylov.cuh"
#include "refelem.h"

#include <algorithm>

namespace stefan {

template <int P>
__global__ void __launch_bounds__(256)
point_eval_kernel(const double* __restrict__ u, const double* __restrict__ refpoints, const int32_t* __restrict__ cellids,
                  int64_t npts, double jx, double jy, double* __restrict__ vals, double* __restrict__ grads) {
  constexpr int NF = (P + 1) * (P + 1);
  for (int64_t q = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; q < npts; q += (int64_t)gridDim.x * blockDim.x) {
    const double px = refpoints[2 * q], py = refpoints[2 * q + 1];
    const double* uc = u + (size_t)cellids[q] * NF;
    double v = 0.0, gx = 0.0, gy = 0.0;
#pragma unroll
    for (int a = 0; a < NF; ++a) {
      double va, ga, gb;
      basis_vg<P>(a, px, py, jx, jy, va, ga, gb);
      const double ua = uc[a];
      v += va * ua;
      gx += ga * ua;
      gy += gb * ua;
    }
    if (vals != nullptr) vals[q] = v;
    if (grads != nullptr) {
      grads[2 * q] = gx;
      grads[2 * q + 1] = gy;
    }
  }
}

__global__ void __launch_bounds__(256)
front_velocity_kernel(int64_t npts, double lambda, double Tm, double k1, double k2, const double* __restrict__ vals1,
                      const double* __restrict__ vals2, const double* __restrict__ grads1,
                      const double* __restrict__ grads2, const double* __restrict__ normals, double* __restrict__ velocity1,
                      double* __restrict__ velocity2, double* __restrict__ meanvelocity, double* __restrict__ gvelocity) {
  for (int64_t q = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; q < npts; q += (int64_t)gridDim.x * blockDim.x) {
    const double v1 = -lambda * (Tm - vals1[q]), v2 = -lambda * (Tm - vals2[q]);
    if (velocity1) velocity1[q] = v1;
    if (velocity2) velocity2[q] = v2;
    if (meanvelocity) meanvelocity[q] = 0.5 * (v1 + v2);
    if (gvelocity && grads1 && grads2 && normals) {
      const double nx = normals[2 * q], ny = normals[2 * q + 1];
      const double f1 = k1 * (grads1[2 * q] * nx + grads1[2 * q + 1] * ny);
      const double f2 = k2 * (grads2[2 * q] * nx + grads2[2 * q + 1] * ny);
      gvelocity[q] = f2 - f1;
    }
  }
}

// one thread per face point of the selected slots
template <int P>
__global__ void __launch_bounds__(kry::kThreads)
blocks_iface_l2_kernel(const stefan_blocks_dims d, const double* __restrict__ face_pts, const double* __restrict__ face_wts,
                       const double* __restrict__ cell_jac, const int8_t* __restrict__ slot_select,
                       const double* __restrict__ u, const double* __restrict__ uex, double* __restrict__ partial) {
  constexpr int NF = (P + 1) * (P + 1);
  const int64_t npts = (int64_t)d.ncells * d.nslots * d.nqf;
  double e2 = 0.0, n2 = 0.0;
  for (int64_t g = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; g < npts; g += (int64_t)gridDim.x * blockDim.x) {
    const int64_t cs = g / d.nqf;
    const double w = face_wts[g];
    if (w == 0.0 || !slot_select[cs]) continue;
    const int cell = (int)(cs / d.nslots);
    const double jx = cell_jac[2 * cell], jy = cell_jac[2 * cell + 1];
    double uh = 0.0;
#pragma unroll
    for (int a = 0; a < NF; ++a) {
      double v, gx, gy;
      basis_vg<P>(a, face_pts[2 * g], face_pts[2 * g + 1], jx, jy, v, gx, gy);
      uh += v * u[(size_t)cell * NF + a];
    }
    const double ue = uex[g];
    e2 += (uh - ue) * (uh - ue) * w;
    n2 += ue * ue * w;
  }
  e2 = kry::block_sum(e2);
  n2 = kry::block_sum(n2);
  if (threadIdx.x == 0) {
    partial[2 * blockIdx.x] = e2;
    partial[2 * blockIdx.x + 1] = n2;
  }
}

__global__ void iface_l2_final_kernel(const double* __restrict__ partial, int nblocks, double* __restrict__ out2) {
  if (threadIdx.x != 0 || blockIdx.x != 0) return;
  double a = 0.0, b = 0.0;
  for (int k = 0; k < nblocks; ++k) {
    a += partial[2 * k];
    b += partial[2 * k + 1];
  }
  out2[0] = sqrt(a);
  out2[1] = sqrt(b);
}

// mesh_L2_error / integral_norm_on_mesh on a uniform mesh for `dpn` dofs per node (1: IP, 3: LDG), 1-D or 2-D:
// one thread per cell; per block partial[(2 c) nblk + b] / [(2 c + 1) nblk + b] for component c
struct MeshL2Tab {
  int dim, n1, nq, dpn;
  double Nq[kMaxQ][kMaxN1], w[kMaxQ];
  double detjac;
};
constexpr int kMaxDpn = 3;

__global__ void __launch_bounds__(kry::kThreads)
mesh_l2_kernel(const __grid_constant__ MeshL2Tab t, const double* __restrict__ u, const double* __restrict__ uex_qp,
               const double* __restrict__ cell_detjac, int64_t ncells, double* __restrict__ partial) {
  const int n1 = t.n1, nq = t.nq, nf = t.dim == 2 ? n1 * n1 : n1, npt = t.dim == 2 ? nq * nq : nq, dpn = t.dpn;
  double se[kMaxDpn] = {0.0, 0.0, 0.0}, sr[kMaxDpn] = {0.0, 0.0, 0.0};
  for (int64_t cell = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; cell < ncells;
       cell += (int64_t)gridDim.x * blockDim.x) {
    const double* uc = u + cell * nf * dpn;
    const double* ue = uex_qp + cell * (int64_t)npt * dpn;
    const double dj = cell_detjac ? cell_detjac[cell] : t.detjac;
    for (int q = 0; q < npt; ++q) {
      const int qx = t.dim == 2 ? q / nq : q, qy = t.dim == 2 ? q % nq : 0;
      const double wq = (t.dim == 2 ? t.w[qx] * t.w[qy] : t.w[qx]) * dj;
      for (int c = 0; c < dpn; ++c) {
        double v = 0.0;
        if (t.dim == 2) {
          for (int ix = 0; ix < n1; ++ix)
            for (int iy = 0; iy < n1; ++iy) v += t.Nq[qx][ix] * t.Nq[qy][iy] * uc[(ix * n1 + iy) * dpn + c];
        } else {
          for (int ix = 0; ix < n1; ++ix) v += t.Nq[qx][ix] * uc[ix * dpn + c];
        }
        const double r = ue[q * dpn + c], e = v - r;
        se[c] += e * e * wq;
        sr[c] += r * r * wq;
      }
    }
  }
  for (int c = 0; c < dpn; ++c) {
    const double a = kry::block_sum(se[c]);
    const double b = kry::block_sum(sr[c]);
    if (threadIdx.x == 0) {
      partial[(size_t)(2 * c) * gridDim.x + blockIdx.x] = a;
      partial[(size_t)(2 * c + 1) * gridDim.x + blockIdx.x] = b;
    }
  }
}

__global__ void mesh_l2_final_kernel(const double* __restrict__ partial, int nblocks, int dpn, double* __restrict__ out) {
  const int k = threadIdx.x;   // one thread per (component, err / norm)
  if (blockIdx.x != 0 || k >= 2 * dpn) return;
  double s = 0.0;
  for (int b = 0; b < nblocks; ++b) s += partial[(size_t)k * nblocks + b];
  out[k] = sqrt(s);
}

// interpolation of `ncomp` consecutive components (first: comp0) of a field with dpn dofs per node
template <int P>
__global__ void __launch_bounds__(256)
field_eval_kernel(const double* __restrict__ u, int dpn, int comp0, int ncomp, const double* __restrict__ refpoints,
                  const int32_t* __restrict__ cellids, int64_t npts, double* __restrict__ vals) {
  constexpr int NF = (P + 1) * (P + 1);
  for (int64_t q = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; q < npts; q += (int64_t)gridDim.x * blockDim.x) {
    const double px = refpoints[2 * q], py = refpoints[2 * q + 1];
    const double* uc = u + (size_t)cellids[q] * NF * dpn + comp0;
    double v[kMaxDpn] = {0.0, 0.0, 0.0};
#pragma unroll
    for (int a = 0; a < NF; ++a) {
      double va, ga, gb;
      basis_vg<P>(a, px, py, 1.0, 1.0, va, ga, gb);
      for (int c = 0; c < ncomp; ++c) v[c] += va * uc[a * dpn + c];
    }
    for (int c = 0; c < ncomp; ++c) vals[q * ncomp + c] = v[c];
  }
}

// lagrange_levelset_LDG_interface_flux.jl:266-307: normal fluxes of both sides and the LDG numerical flux
__global__ void __launch_bounds__(256)
ldg_interface_flux_kernel(int64_t npts, double k1, double k2, double V0x, double V0y, const double* __restrict__ g1,
                          const double* __restrict__ g2, const double* __restrict__ normals, double* __restrict__ nf1,
                          double* __restrict__ nf2, double* __restrict__ numflux) {
  for (int64_t q = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; q < npts; q += (int64_t)gridDim.x * blockDim.x) {
    const double nx = normals[2 * q], ny = normals[2 * q + 1];
    const double f1 = k1 * (g1[2 * q] * nx + g1[2 * q + 1] * ny), f2 = k2 * (g2[2 * q] * nx + g2[2 * q + 1] * ny);
    // beta = edge_switch(V0, normals) = 1/2 sign(V0 . n) n  (LDG_utilities.jl:1-4)
    const double v0n = V0x * nx + V0y * ny, sg = v0n > 0 ? 1.0 : (v0n < 0 ? -1.0 : 0.0);
    const double bx = 0.5 * sg * nx, by = 0.5 * sg * ny;
    // numericalflux = 1/2 (k1 grads1 + k2 grads2) - beta (-normalflux1 + normalflux2); its normal component
    const double fx = 0.5 * (k1 * g1[2 * q] + k2 * g2[2 * q]) - bx * (f2 - f1);
    const double fy = 0.5 * (k1 * g1[2 * q + 1] + k2 * g2[2 * q + 1]) - by * (f2 - f1);
    if (nf1) nf1[q] = f1;
    if (nf2) nf2[q] = f2;
    if (numflux) numflux[q] = fx * nx + fy * ny;
  }
}

}  // namespace stefan

using namespace stefan;

static int point_eval(int32_t order, const double* u, const double* refpoints, const int32_t* cellids, int64_t npts,
                      double jx, double jy, double* vals, double* grads, void* stream, const char* who) {
  STEFAN_REQUIRE(order >= 1 && order <= 3, "%s: order must be 1..3", who);
  STEFAN_REQUIRE(u && refpoints && cellids && npts >= 0 && (vals || grads), "%s: null argument", who);
  if (npts == 0) return ST_OK;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const int blocks = (int)std::min<int64_t>((npts + 255) / 256, (int64_t)sm_count() * 16);
  switch (order) {
    case 1: point_eval_kernel<1><<<blocks, 256, 0, st>>>(u, refpoints, cellids, npts, jx, jy, vals, grads); break;
    case 2: point_eval_kernel<2><<<blocks, 256, 0, st>>>(u, refpoints, cellids, npts, jx, jy, vals, grads); break;
    case 3: point_eval_kernel<3><<<blocks, 256, 0, st>>>(u, refpoints, cellids, npts, jx, jy, vals, grads); break;
  }
  STEFAN_CUDA(cudaGetLastError());
  return ST_OK;
}

extern "C" {

int stefan_interpolate_at_reference_points(int32_t order, const double* nodalvalues, const double* refpoints,
                                           const int32_t* refcellids, int64_t numpts, double* vals, void* stream) {
  return point_eval(order, nodalvalues, refpoints, refcellids, numpts, 1.0, 1.0, vals, nullptr, stream,
                    "stefan_interpolate_at_reference_points");
}

int stefan_gradient_at_reference_points(int32_t order, const double* nodalvalues, const double* refpoints,
                                        const int32_t* refcellids, int64_t numpts, double jx, double jy, double* grads,
                                        void* stream) {
  STEFAN_REQUIRE(jx > 0 && jy > 0, "stefan_gradient_at_reference_points: the jacobian must be positive");
  return point_eval(order, nodalvalues, refpoints, refcellids, numpts, jx, jy, nullptr, grads, stream,
                    "stefan_gradient_at_reference_points");
}

int stefan_front_velocity(int64_t numpts, double lambda, double Tm, double k1, double k2, const double* vals1,
                          const double* vals2, const double* grads1, const double* grads2, const double* normals,
                          double* velocity1, double* velocity2, double* meanvelocity, double* gvelocity, void* stream) {
  STEFAN_REQUIRE(numpts >= 0 && vals1 && vals2, "stefan_front_velocity: null argument");
  STEFAN_REQUIRE(!gvelocity || (grads1 && grads2 && normals), "stefan_front_velocity: gvelocity needs gradients and normals");
  if (numpts == 0) return ST_OK;
  const int blocks = (int)std::min<int64_t>((numpts + 255) / 256, (int64_t)sm_count() * 16);
  front_velocity_kernel<<<blocks, 256, 0, static_cast<cudaStream_t>(stream)>>>(numpts, lambda, Tm, k1, k2, vals1, vals2,
                                                                               grads1, grads2, normals, velocity1,
                                                                               velocity2, meanvelocity, gvelocity);
  STEFAN_CUDA(cudaGetLastError());
  return ST_OK;
}

int stefan_interpolate_field_at_reference_points(int32_t order, const double* nodalvalues, int32_t dofspernode,
                                                 int32_t first_component, int32_t ncomponents, const double* refpoints,
                                                 const int32_t* refcellids, int64_t numpts, double* vals, void* stream) {
  STEFAN_REQUIRE(order >= 1 && order <= 3, "stefan_interpolate_field_at_reference_points: order must be 1..3");
  STEFAN_REQUIRE(nodalvalues && refpoints && refcellids && vals && numpts >= 0,
                 "stefan_interpolate_field_at_reference_points: null argument");
  STEFAN_REQUIRE(dofspernode >= 1 && first_component >= 0 && ncomponents >= 1 && ncomponents <= kMaxDpn &&
                 first_component + ncomponents <= dofspernode,
                 "stefan_interpolate_field_at_reference_points: need 1 <= ncomponents <= 3 inside the node's dofs");
  if (numpts == 0) return ST_OK;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const int blocks = (int)std::min<int64_t>((numpts + 255) / 256, (int64_t)sm_count() * 16);
  switch (order) {
    case 1: field_eval_kernel<1><<<blocks, 256, 0, st>>>(nodalvalues, dofspernode, first_component, ncomponents, refpoints, refcellids, numpts, vals); break;
    case 2: field_eval_kernel<2><<<blocks, 256, 0, st>>>(nodalvalues, dofspernode, first_component, ncomponents, refpoints, refcellids, numpts, vals); break;
    case 3: field_eval_kernel<3><<<blocks, 256, 0, st>>>(nodalvalues, dofspernode, first_component, ncomponents, refpoints, refcellids, numpts, vals); break;
  }
  STEFAN_CUDA(cudaGetLastError());
  return ST_OK;
}

int stefan_ldg_interface_flux(int64_t numpts, double k1, double k2, double V0x, double V0y, const double* grads1,
                              const double* grads2, const double* normals, double* normalflux1, double* normalflux2,
                              double* numericalnormalflux, void* stream) {
  STEFAN_REQUIRE(numpts >= 0 && grads1 && grads2 && normals, "stefan_ldg_interface_flux: null argument");
  if (numpts == 0) return ST_OK;
  const int blocks = (int)std::min<int64_t>((numpts + 255) / 256, (int64_t)sm_count() * 16);
  ldg_interface_flux_kernel<<<blocks, 256, 0, static_cast<cudaStream_t>(stream)>>>(
      numpts, k1, k2, V0x, V0y, grads1, grads2, normals, normalflux1, normalflux2, numericalnormalflux);
  STEFAN_CUDA(cudaGetLastError());
  return ST_OK;
}

int stefan_mesh_l2_error(int32_t dim, int32_t order, int32_t numqp, int64_t ncells, double detjac,
                         const double* cell_detjac, int32_t dofspernode, const double* u, const double* uex_qp,
                         double* out_dev, void* stream) {
  STEFAN_REQUIRE((dim == 1 || dim == 2) && order >= 1 && order <= 3 && numqp >= 1 && numqp <= kMaxQ,
                 "stefan_mesh_l2_error: dim 1..2, order 1..3, numqp 1..%d", kMaxQ);
  STEFAN_REQUIRE(ncells >= 1 && dofspernode >= 1 && dofspernode <= kMaxDpn && u && uex_qp && out_dev &&
                 (cell_detjac || detjac > 0), "stefan_mesh_l2_error: bad argument");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  RefElem1D ref;
  build_ref1d(order, numqp, &ref);
  MeshL2Tab t;
  t.dim = dim;
  t.n1 = order + 1;
  t.nq = numqp;
  t.dpn = dofspernode;
  t.detjac = detjac;
  for (int q = 0; q < kMaxQ; ++q) {
    t.w[q] = q < numqp ? ref.qw[q] : 0.0;
    for (int a = 0; a < kMaxN1; ++a) t.Nq[q][a] = (q < numqp && a <= order) ? ref.Nq[q][a] : 0.0;
  }
  const int nblk = (int)std::min<int64_t>((ncells + kry::kThreads - 1) / kry::kThreads, (int64_t)sm_count() * 8);
  double* part = nullptr;
  STEFAN_CUDA(cudaMallocAsync(&part, (size_t)2 * dofspernode * nblk * sizeof(double), st));
  mesh_l2_kernel<<<nblk, kry::kThreads, 0, st>>>(t, u, uex_qp, cell_detjac, ncells, part);
  mesh_l2_final_kernel<<<1, 32, 0, st>>>(part, nblk, dofspernode, out_dev);
  STEFAN_CUDA(cudaFreeAsync(part, st));
  STEFAN_CUDA(cudaGetLastError());
  return ST_OK;
}

int stefan_blocks_interface_l2_error(stefan_blocks* hv, const stefan_blocks_data* in, const int8_t* slot_select,
                                     const double* u, const double* uex_fq, double* out2_dev, void* stream) {
  BlocksHandle* h = reinterpret_cast<BlocksHandle*>(hv);
  STEFAN_REQUIRE(h && in && slot_select && u && uex_fq && out2_dev, "stefan_blocks_interface_l2_error: null argument");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const int64_t npts = (int64_t)h->d.ncells * h->d.nslots * h->d.nqf;
  const int nblk = (int)std::min<int64_t>((npts + kry::kThreads - 1) / kry::kThreads, (int64_t)sm_count() * 8);
  double* part = nullptr;
  STEFAN_CUDA(cudaMallocAsync(&part, 2 * (size_t)nblk * sizeof(double), st));
  switch (h->d.order) {
    case 1: blocks_iface_l2_kernel<1><<<nblk, kry::kThreads, 0, st>>>(h->d, in->face_pts, in->face_wts, in->cell_jac, slot_select, u, uex_fq, part); break;
    case 2: blocks_iface_l2_kernel<2><<<nblk, kry::kThreads, 0, st>>>(h->d, in->face_pts, in->face_wts, in->cell_jac, slot_select, u, uex_fq, part); break;
    case 3: blocks_iface_l2_kernel<3><<<nblk, kry::kThreads, 0, st>>>(h->d, in->face_pts, in->face_wts, in->cell_jac, slot_select, u, uex_fq, part); break;
  }
  iface_l2_final_kernel<<<1, 1, 0, st>>>(part, nblk, out2_dev);
  STEFAN_CUDA(cudaFreeAsync(part, st));
  STEFAN_CUDA(cudaGetLastError());
  return ST_OK;
}

}  // extern "C"
