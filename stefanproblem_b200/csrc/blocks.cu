// Element-block path: the reference's element-local assembly with ARBITRARY per-cell and
// per-face quadratures (what a cut-cell mesh produces), and the application of the blocks.
//
// The uniform-mesh kernels (ipdg_march.cuh) never form an element matrix: on an uncut
// uniform mesh every block is a Kronecker product of 1-D tables.  A cut / merged cell has
// its own quadrature points, weights, interface normals and scale areas, so its blocks have
// to be assembled the way the reference does it -- quadrature loops over dense NF x NF
// outer products:
//   gradient_operator                 IP_gradient_operator.jl:1-9     sum k G G' detjac w
//   flux_operator                     IP_flux_operator.jl:1-33        sum k (G n) V' sa w
//   assemble_face_flux_operator!      IP_flux_operator.jl:35-118      M11 + M11', -M12 + M21'
//   mass_operator / face penalty      IP_penalty_operator.jl:1-65     pen sum V V' sa w (+, -)
//   assemble_face_robin_operator!     IP_robin_interface_operator.jl:1-45   lambda/4 sum V V' sa w
//   boundary flux / penalty           IP_flux_operator.jl:330-357, IP_penalty_operator.jl:239-254
// Cell-centric form: a solution cell (a cut cell contributes one per phase) has `nslots`
// face slots; for every slot the caller gives the quadrature points in the cell's own and
// in the neighbour's reference coordinates, weight x scale area, and the normal pointing out
// of the cell.  Seen from the cell (side 1), with M11 = -1/2 sum k1 (G1 n) V1' w etc.:
//     diag      += M11 + M11' + pen sum V1 V1' w + lambda/4 sum V1 V1' w
//     off[slot]  = -M12 + M21' - pen sum V1 V2' w + lambda/4 sum V1 V2' w
// which is exactly what the reference's face functions add to the (1,1) and (1,2) blocks;
// the neighbour's own slot produces the reference's (2,2) and (2,1) blocks (its normal is
// the opposite one).  Boundary slot: diag += -(M + M') + pen sum V1 V1' w, M = sum k1 (G1 n) V1' w.
//
// Assembly kernel: one warp per cell; basis tables at the cell's quadrature points are staged
// in shared memory, the block entries are dot products over the points (no atomics, coalesced
// block writes).  Apply kernel: a few lanes per row (cell, a), coalesced block reads, shuffle
// reduction; HBM-bound by the blocks: 16 + 8 (1 + nslots) NF bytes per dof-update
// ("general block path").
#include "../../include/stefan_b200.h"
#include "blocks_handle.cuh"
#include "common.cuh"

#include <algorithm>

namespace stefan {

// Assembly: one WARP per cell.  The lanes first tabulate, in shared memory, the basis values
// and physical gradients at the cell's quadrature points (volume points, then slot by slot
// the own side and the neighbour's side) -- NF evaluations per point instead of 2 per block
// entry -- and then each lane accumulates its share of the NF x NF entries as short dot
// products over the points.  Coalesced block writes, no atomics.
constexpr int kAsmWarps = 4;
template <int P>
struct AsmSmem {
  static constexpr int NF = (P + 1) * (P + 1);
  // per warp: tables T[4][maxq][NF] (own value, own directional / gradient data, neighbour value, neighbour gn)
  static constexpr int MAXQ = 32;
  static constexpr int WARP_DOUBLES = 4 * MAXQ * NF;
};

template <int P>
__global__ void __launch_bounds__(kAsmWarps * 32)
blocks_assemble_kernel(const stefan_blocks_dims d, const stefan_blocks_data in, int legacy, double* __restrict__ blocks) {
  constexpr int NF = (P + 1) * (P + 1);
  constexpr int MAXQ = AsmSmem<P>::MAXQ;
  extern __shared__ double asm_smem[];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  double* const T0 = asm_smem + (size_t)warp * AsmSmem<P>::WARP_DOUBLES;  // [MAXQ][NF]
  double* const T1 = T0 + MAXQ * NF;
  double* const T2 = T1 + MAXQ * NF;
  double* const T3 = T2 + MAXQ * NF;
  const int nb = 1 + d.nslots;
  for (int cell = blockIdx.x * kAsmWarps + warp; cell < d.ncells; cell += gridDim.x * kAsmWarps) {
    const double jx = in.cell_jac[2 * cell], jy = in.cell_jac[2 * cell + 1], k1 = in.cell_k[cell];
    double* const out = blocks + ((size_t)cell * nb) * NF * NF;
    // ---- volume: T0 = sqrt-free tables gx, T1 = gy, T2 = weight (per point, broadcast) ----
    double acc[(NF * NF + 31) / 32];
#pragma unroll
    for (int e = 0; e < (NF * NF + 31) / 32; ++e) acc[e] = 0.0;
    for (int q0 = 0; q0 < d.nqv; q0 += MAXQ) {
      const int nq = min(MAXQ, d.nqv - q0);
      for (int t = lane; t < nq * NF; t += 32) {
        const int q = t / NF, a = t % NF;
        const size_t g = (size_t)cell * d.nqv + q0 + q;
        double v, gx, gy;
        basis_vg<P>(a, in.vol_pts[2 * g], in.vol_pts[2 * g + 1], jx, jy, v, gx, gy);
        const double wq = in.vol_wts[g] * k1 * jx * jy;
        T0[q * NF + a] = gx;
        T1[q * NF + a] = gy;
        T2[q * NF + a] = gx * wq;
        T3[q * NF + a] = gy * wq;
      }
      __syncwarp();
#pragma unroll
      for (int e = 0; e < (NF * NF + 31) / 32; ++e) {
        const int ab = lane + 32 * e;
        if (ab < NF * NF) {
          const int a = ab / NF, b = ab % NF;
          double sacc = 0.0;
          for (int q = 0; q < nq; ++q) sacc += T2[q * NF + a] * T0[q * NF + b] + T3[q * NF + a] * T1[q * NF + b];
          acc[e] += sacc;
        }
      }
      __syncwarp();
    }
    // ---- face slots: T0 = V1, T1 = gn1 (G1 . n), T2 = V2, T3 = gn2 at the slot's points ----
    for (int s = 0; s < d.nslots; ++s) {
      const size_t cs = (size_t)cell * d.nslots + s;
      const int nbr = in.slot_nbr[cs];
      double off[(NF * NF + 31) / 32];
#pragma unroll
      for (int e = 0; e < (NF * NF + 31) / 32; ++e) off[e] = 0.0;
      if (nbr >= -1) {
        const double pen = in.slot_pen[cs], lam4 = 0.25 * in.slot_lambda[cs];
        double jx2 = jx, jy2 = jy, k2 = k1;
        if (nbr >= 0) {
          jx2 = in.cell_jac[2 * nbr];
          jy2 = in.cell_jac[2 * nbr + 1];
          k2 = in.cell_k[nbr];
        }
        for (int q0 = 0; q0 < d.nqf; q0 += MAXQ) {
          const int nq = min(MAXQ, d.nqf - q0);
          for (int t = lane; t < nq * NF; t += 32) {
            const int q = t / NF, a = t % NF;
            const size_t fq = cs * d.nqf + q0 + q;
            const double nx = in.face_normals[2 * fq], ny = in.face_normals[2 * fq + 1];
            double v, gx, gy;
            basis_vg<P>(a, in.face_pts[2 * fq], in.face_pts[2 * fq + 1], jx, jy, v, gx, gy);
            T0[q * NF + a] = v;
            T1[q * NF + a] = gx * nx + gy * ny;
            if (nbr >= 0) {
              basis_vg<P>(a, in.face_nbr_pts[2 * fq], in.face_nbr_pts[2 * fq + 1], jx2, jy2, v, gx, gy);
              T2[q * NF + a] = v;
              T3[q * NF + a] = gx * nx + gy * ny;
            }
          }
          __syncwarp();
#pragma unroll
          for (int e = 0; e < (NF * NF + 31) / 32; ++e) {
            const int ab = lane + 32 * e;
            if (ab < NF * NF) {
              const int a = ab / NF, b = ab % NF;
              double dsum = 0.0, osum = 0.0;
              for (int q = 0; q < nq; ++q) {
                const double wq = in.face_wts[cs * d.nqf + q0 + q];
                const double v1a = T0[q * NF + a], v1b = T0[q * NF + b], gn1a = T1[q * NF + a], gn1b = T1[q * NF + b];
                if (nbr >= 0) {
                  // (M11 + M11')[a][b] = -1/2 k1 (gn1a v1b + gn1b v1a); -M12[a][b] = 1/2 k1 gn1a v2b;
                  // M21'[a][b] = -1/2 k2 gn2b v1a
                  dsum += (-0.5 * k1 * (gn1a * v1b + gn1b * v1a) + (pen + lam4) * v1a * v1b) * wq;
                  const double v2b = T2[q * NF + b], gn2b = T3[q * NF + b];
                  osum += (0.5 * k1 * gn1a * v2b - 0.5 * k2 * gn2b * v1a + (lam4 - pen) * v1a * v2b) * wq;
                } else {
                  // boundary: -(M + M') + pen P, M[a][b] = k1 gn1a v1b; legacy boundary form: -M' + pen P
                  dsum += (-k1 * ((legacy ? 0.0 : gn1a * v1b) + gn1b * v1a) + pen * v1a * v1b) * wq;
                }
              }
              acc[e] += dsum;
              off[e] += osum;
            }
          }
          __syncwarp();
        }
      }
#pragma unroll
      for (int e = 0; e < (NF * NF + 31) / 32; ++e) {
        const int ab = lane + 32 * e;
        if (ab < NF * NF) out[(size_t)(1 + s) * NF * NF + ab] = off[e];
      }
    }
#pragma unroll
    for (int e = 0; e < (NF * NF + 31) / 32; ++e) {
      const int ab = lane + 32 * e;
      if (ab < NF * NF) out[ab] = acc[e];
    }
  }
}

// Apply, orders 1 and 2: one thread per row (cell, a); a warp's rows are contiguous in the
// block array, the short rows (4 / 9 doubles) stay in L1 between the loads of a thread.
// Measured at 2000^2 / 1000^2 cells: 94.6 % / 87.5 % of the measured HBM peak at 176 / 376 B
// per dof (the grouped kernel below: 69 % / 53 %).
template <int P>
__global__ void __launch_bounds__(256)
blocks_apply_rows_kernel(int ncells, int nslots, const double* __restrict__ blocks,
                         const int32_t* __restrict__ slot_nbr, const double* __restrict__ x, double* __restrict__ y) {
  constexpr int NF = (P + 1) * (P + 1);
  const int64_t nrows = (int64_t)ncells * NF;
  const int nb = 1 + nslots;
  for (int64_t r = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; r < nrows; r += (int64_t)gridDim.x * blockDim.x) {
    const int cell = (int)(r / NF), a = (int)(r % NF);
    const double* B = blocks + ((size_t)cell * nb) * NF * NF + (size_t)a * NF;
    const double* xc = x + (size_t)cell * NF;
    double v = 0.0;
#pragma unroll
    for (int b = 0; b < NF; ++b) v += B[b] * xc[b];
    for (int s = 0; s < nslots; ++s) {
      const int nbr = slot_nbr[(size_t)cell * nslots + s];
      if (nbr < 0) continue;
      const double* Bo = B + (size_t)(1 + s) * NF * NF;
      const double* xn = x + (size_t)nbr * NF;
#pragma unroll
      for (int b = 0; b < NF; ++b) v += Bo[b] * xn[b];
    }
    y[r] = v;
  }
}

// Apply, order 3: a group of 16 lanes per row (cell, a): the lanes read the 128-byte block
// row coalesced and reduce by shuffles (74 % of the measured HBM peak at 656 B per dof; one
// thread per row: 34 %).
template <int P>
__global__ void __launch_bounds__(256)
blocks_apply_kernel(int ncells, int nslots, const double* __restrict__ blocks, const int32_t* __restrict__ slot_nbr,
                    const double* __restrict__ x, double* __restrict__ y) {
  constexpr int NF = (P + 1) * (P + 1);
  constexpr int G = P == 1 ? 4 : (P == 2 ? 8 : 16);  // lanes per row
  const int64_t nrows = (int64_t)ncells * NF;
  const int nb = 1 + nslots;
  const int gl = threadIdx.x % G;
  const int64_t ngroups = (int64_t)gridDim.x * blockDim.x / G;
  const int64_t nrows_pad = (nrows + (32 / G) - 1) / (32 / G) * (32 / G);  // whole warps stay in the loop
  for (int64_t r = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) / G; r < nrows_pad; r += ngroups) {
    double v = 0.0;
    if (r < nrows) {
      const int cell = (int)(r / NF), a = (int)(r % NF);
      const double* B = blocks + ((size_t)cell * nb) * NF * NF + (size_t)a * NF;
      const double* xc = x + (size_t)cell * NF;
      for (int b = gl; b < NF; b += G) v += B[b] * xc[b];
      for (int s = 0; s < nslots; ++s) {
        const int nbr = slot_nbr[(size_t)cell * nslots + s];
        if (nbr < 0) continue;
        const double* Bo = B + (size_t)(1 + s) * NF * NF;
        const double* xn = x + (size_t)nbr * NF;
        for (int b = gl; b < NF; b += G) v += Bo[b] * xn[b];
      }
    }
#pragma unroll
    for (int o = G / 2; o > 0; o >>= 1) v += __shfl_down_sync(0xffffffffu, v, o, G);
    if (gl == 0 && r < nrows) y[r] = v;
  }
}

}  // namespace stefan

using namespace stefan;

extern "C" {

int stefan_blocks_create(const stefan_blocks_dims* dims, stefan_blocks** out) {
  STEFAN_REQUIRE(dims && out, "stefan_blocks_create: null argument");
  STEFAN_REQUIRE(dims->ncells >= 1 && dims->order >= 1 && dims->order <= 3, "stefan_blocks_create: bad ncells / order");
  STEFAN_REQUIRE(dims->nqv >= 1 && dims->nslots >= 1 && dims->nslots <= 16 && dims->nqf >= 1,
                 "stefan_blocks_create: need nqv >= 1, 1 <= nslots <= 16, nqf >= 1");
  BlocksHandle* h = new BlocksHandle();
  h->d = *dims;
  h->nf = (dims->order + 1) * (dims->order + 1);
  h->blocks = nullptr;
  h->slot_nbr = nullptr;
  h->assembled = false;
  h->variant = 0;
  const size_t nbytes = (size_t)dims->ncells * (1 + dims->nslots) * h->nf * h->nf * sizeof(double);
  cudaError_t e = cudaMalloc(&h->blocks, nbytes);
  if (e == cudaSuccess) e = cudaMalloc(&h->slot_nbr, (size_t)dims->ncells * dims->nslots * sizeof(int32_t));
  if (e != cudaSuccess) {
    if (h->blocks) cudaFree(h->blocks);
    delete h;
    return fail(ST_CUDA_ERROR, "stefan_blocks_create: %s", cudaGetErrorString(e));
  }
  *out = reinterpret_cast<stefan_blocks*>(h);
  return ST_OK;
}

int stefan_blocks_destroy(stefan_blocks* hv) {
  BlocksHandle* h = reinterpret_cast<BlocksHandle*>(hv);
  if (!h) return ST_OK;
  if (h->blocks) cudaFree(h->blocks);
  if (h->slot_nbr) cudaFree(h->slot_nbr);
  delete h;
  return ST_OK;
}

int stefan_blocks_set_variant(stefan_blocks* hv, int32_t variant) {
  BlocksHandle* h = reinterpret_cast<BlocksHandle*>(hv);
  STEFAN_REQUIRE(h, "stefan_blocks_set_variant: null handle");
  STEFAN_REQUIRE(variant == 0 || variant == 1, "stefan_blocks_set_variant: variant must be 0 (current) or 1 (legacy_boundary)");
  h->variant = variant;
  h->assembled = false;
  return ST_OK;
}

int64_t stefan_blocks_ndofs(const stefan_blocks* hv) {
  const BlocksHandle* h = reinterpret_cast<const BlocksHandle*>(hv);
  return h ? (int64_t)h->d.ncells * h->nf : -1;
}

int stefan_blocks_assemble(stefan_blocks* hv, const stefan_blocks_data* in, void* stream) {
  BlocksHandle* h = reinterpret_cast<BlocksHandle*>(hv);
  STEFAN_REQUIRE(h && in, "stefan_blocks_assemble: null argument");
  STEFAN_REQUIRE(in->vol_pts && in->vol_wts && in->cell_jac && in->cell_k && in->slot_nbr && in->face_pts &&
                 in->face_nbr_pts && in->face_wts && in->face_normals && in->slot_pen && in->slot_lambda,
                 "stefan_blocks_assemble: every array of stefan_blocks_data is required");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  STEFAN_CUDA(cudaMemcpyAsync(h->slot_nbr, in->slot_nbr, (size_t)h->d.ncells * h->d.nslots * sizeof(int32_t),
                              cudaMemcpyDeviceToDevice, st));
  const int blocks = (int)std::min<int64_t>(((int64_t)h->d.ncells + kAsmWarps - 1) / kAsmWarps, (int64_t)sm_count() * 16);
  switch (h->d.order) {
    case 1: {
      constexpr int smem = kAsmWarps * AsmSmem<1>::WARP_DOUBLES * 8;
      blocks_assemble_kernel<1><<<blocks, kAsmWarps * 32, smem, st>>>(h->d, *in, h->variant, h->blocks);
      break;
    }
    case 2: {
      constexpr int smem = kAsmWarps * AsmSmem<2>::WARP_DOUBLES * 8;
      static PerDeviceFlag set2;
      if (!set2.here()) { STEFAN_CUDA(cudaFuncSetAttribute(blocks_assemble_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem)); set2.here() = true; }
      blocks_assemble_kernel<2><<<blocks, kAsmWarps * 32, smem, st>>>(h->d, *in, h->variant, h->blocks);
      break;
    }
    case 3: {
      constexpr int smem = kAsmWarps * AsmSmem<3>::WARP_DOUBLES * 8;
      static PerDeviceFlag set3;
      if (!set3.here()) { STEFAN_CUDA(cudaFuncSetAttribute(blocks_assemble_kernel<3>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem)); set3.here() = true; }
      blocks_assemble_kernel<3><<<blocks, kAsmWarps * 32, smem, st>>>(h->d, *in, h->variant, h->blocks);
      break;
    }
  }
  STEFAN_CUDA(cudaGetLastError());
  h->assembled = true;
  return ST_OK;
}

int stefan_blocks_apply(stefan_blocks* hv, const double* x, double* y, void* stream) {
  BlocksHandle* h = reinterpret_cast<BlocksHandle*>(hv);
  STEFAN_REQUIRE(h && x && y && x != y, "stefan_blocks_apply: bad argument");
  STEFAN_REQUIRE(h->assembled, "stefan_blocks_apply: call stefan_blocks_assemble first");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const int64_t nrows = (int64_t)h->d.ncells * h->nf;
  const int blocks = (int)std::min<int64_t>((nrows + 255) / 256, (int64_t)sm_count() * 32);
  switch (h->d.order) {
    case 1: blocks_apply_rows_kernel<1><<<blocks, 256, 0, st>>>(h->d.ncells, h->d.nslots, h->blocks, h->slot_nbr, x, y); break;
    case 2: blocks_apply_rows_kernel<2><<<blocks, 256, 0, st>>>(h->d.ncells, h->d.nslots, h->blocks, h->slot_nbr, x, y); break;
    case 3: blocks_apply_kernel<3><<<blocks, 256, 0, st>>>(h->d.ncells, h->d.nslots, h->blocks, h->slot_nbr, x, y); break;
  }
  STEFAN_CUDA(cudaGetLastError());
  return ST_OK;
}

int64_t stefan_blocks_coo_size(const stefan_blocks* hv) {
  const BlocksHandle* h = reinterpret_cast<const BlocksHandle*>(hv);
  return h ? (int64_t)h->d.ncells * (1 + h->d.nslots) * h->nf * h->nf : -1;
}

int stefan_blocks_assemble_coo(stefan_blocks* hv, int64_t* rows, int64_t* cols, double* vals, int64_t capacity,
                               int32_t index_base, int64_t* nnz_out, void* stream) {
  BlocksHandle* h = reinterpret_cast<BlocksHandle*>(hv);
  STEFAN_REQUIRE(h && rows && cols && vals && nnz_out, "stefan_blocks_assemble_coo: null argument");
  STEFAN_REQUIRE(h->assembled, "stefan_blocks_assemble_coo: call stefan_blocks_assemble first");
  STEFAN_REQUIRE(index_base == 0 || index_base == 1, "stefan_blocks_assemble_coo: index_base must be 0 or 1");
  *nnz_out = stefan_blocks_coo_size(hv);
  if (*nnz_out > capacity)
    return fail(ST_CAPACITY, "stefan_blocks_assemble_coo: %lld triplets, capacity %lld", (long long)*nnz_out,
                (long long)capacity);
  const int blocks = (int)std::min<int64_t>((*nnz_out + 255) / 256, (int64_t)sm_count() * 32);
  blocks_coo_kernel<<<blocks, 256, 0, static_cast<cudaStream_t>(stream)>>>(h->d.ncells, h->d.nslots, h->nf, h->blocks,
                                                                           h->slot_nbr, index_base, rows, cols, vals);
  STEFAN_CUDA(cudaGetLastError());
  return ST_OK;
}

}  // extern "C"
