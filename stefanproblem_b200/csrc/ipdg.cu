// Interior-penalty DG heat operator on an uncut uniform mesh: matrix-free apply
// kernels for sm_100a and the C-ABI entry points around them.
//
// Replaces (see include/stefan_b200.h for the per-entry citations):
//   InteriorPenalty.assemble_interior_penalty_linear_system!  IP_assembly.jl:1-80
//   InteriorPenalty.assemble_robin_operator!                  IP_robin_interface_operator.jl:76-99
//   + CutCellDG.sparse_operator + `matrix * x`
//
// Data layout in HBM (the reference's dof order):
//   x[(i*ny + j)*NF + ix*(P+1) + iy], fp64; cells y-fastest, eta-node fastest.
//   phase: int8 per cell, padded to (nx+2) x (ny+2) with 0 = "no cell".
//
// Kernels
//   ipdg_apply_simple<P>: one thread per cell, neighbours straight from global
//       memory (L1/L2).  Plain, used as the second opinion for the stream kernel.
//   ipdg_apply_stream<P>: the production kernel.  One warp owns a strip of 30 cells
//       in y (+1 halo cell each side = 32 lanes) and marches along x.  Each column
//       chunk (32 cells, contiguous in memory) is staged by cp.async into a per-warp
//       shared-memory ring NST columns deep, so NST-2 columns per warp are in flight
//       while one is computed; the x-neighbours live in registers, the y-neighbours
//       are read from the ring (lane +-1).  Every dof is read from HBM once (plus
//       2/30 halo re-reads that hit L2) and every result written once: 16 B per
//       dof-update.  Warps whose 30 cells are all interior cells of one phase take
//       the path whose coefficients are compile-time offsets into the kernel
//       parameter bank; the others index the same tables per lane.
#include "../../include/stefan_b200.h"
#include "common.cuh"
#include "ipdg_tables.h"

#include <vector>

namespace stefan {

struct IpdgHandle {
  int nx, ny, order, numqp;
  double hx, hy;
  IpPhys phys;
  double Tm;
  bool has_lo, has_hi;
  int8_t* phase_dev;  // (nx+2)*(ny+2), padded
  void* tab_host;     // IpTab<P>
  RefElem1D ref;
  int64_t ndofs;
};

template <int P>
struct StreamCfg;
template <>
struct StreamCfg<1> {
  static constexpr int PIECE = 16, CELLB = 32, STRIDE = 48, NST = 6, WARPS = 8, MINB = 2;
};
template <>
struct StreamCfg<2> {
  static constexpr int PIECE = 8, CELLB = 72, STRIDE = 72, NST = 5, WARPS = 8, MINB = 2;
};
template <>
struct StreamCfg<3> {
  static constexpr int PIECE = 16, CELLB = 128, STRIDE = 144, NST = 4, WARPS = 6, MINB = 2;
};

constexpr int kStripCells = 30;

struct ApplyArgs {
  const double* x;
  const double* x_lo;  // neighbour slab's last column (ny*NF) or nullptr
  const double* x_hi;  // neighbour slab's first column or nullptr
  double* y;
  const int8_t* phase;  // padded
  int nx, ny;
  int nseg;  // x segments (stream kernel)
};

__device__ __forceinline__ int phase_at(const int8_t* __restrict__ ph, int ny, int i, int j) {
  return ph[(size_t)(i + 1) * (ny + 2) + (j + 1)];
}

// --------------------------------------------------------------------------
// simple kernel
// --------------------------------------------------------------------------
template <int P>
__global__ void __launch_bounds__(128)
ipdg_apply_simple(const __grid_constant__ IpTab<P> tab, const ApplyArgs a) {
  constexpr int NF = (P + 1) * (P + 1);
  const int64_t ncells = (int64_t)a.nx * a.ny;
  for (int64_t cell = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; cell < ncells;
       cell += (int64_t)gridDim.x * blockDim.x) {
    const int i = (int)(cell / a.ny), j = (int)(cell % a.ny);
    const int s = phase_at(a.phase, a.ny, i, j);
    const int sL = phase_at(a.phase, a.ny, i - 1, j), sR = phase_at(a.phase, a.ny, i + 1, j);
    const int sD = phase_at(a.phase, a.ny, i, j - 1), sU = phase_at(a.phase, a.ny, i, j + 1);
    double X[NF], XL[NF], XR[NF], XD[NF], XU[NF], out[NF];
    const double* xc = a.x + cell * NF;
    const double* pl = (i > 0) ? xc - (size_t)a.ny * NF : a.x_lo + (size_t)j * NF;
    const double* pr = (i < a.nx - 1) ? xc + (size_t)a.ny * NF : a.x_hi + (size_t)j * NF;
#pragma unroll
    for (int k = 0; k < NF; ++k) {
      X[k] = xc[k];
      XL[k] = sL ? pl[k] : 0.0;
      XR[k] = sR ? pr[k] : 0.0;
      XD[k] = sD ? xc[k - NF] : 0.0;
      XU[k] = sU ? xc[k + NF] : 0.0;
    }
    ip_cell_rows<P>(tab, s > 0 ? 0 : 1, ip_face_type(s, sL),
                    ip_face_type(s, sR), ip_face_type(s, sD),
                    ip_face_type(s, sU), X, XL, XR, XD, XU, out);
    double* yc = a.y + cell * NF;
#pragma unroll
    for (int k = 0; k < NF; ++k) yc[k] = out[k];
  }
}

// --------------------------------------------------------------------------
// stream kernel
// --------------------------------------------------------------------------
template <int P>
__device__ __forceinline__ void load_cell(const char* base, double* v) {
  constexpr int NF = (P + 1) * (P + 1);
  if constexpr (StreamCfg<P>::PIECE == 16) {
    const double2* p = reinterpret_cast<const double2*>(base);
#pragma unroll
    for (int k = 0; k < NF / 2; ++k) {
      double2 t = p[k];
      v[2 * k] = t.x;
      v[2 * k + 1] = t.y;
    }
  } else {
    const double* p = reinterpret_cast<const double*>(base);
#pragma unroll
    for (int k = 0; k < NF; ++k) v[k] = p[k];
  }
}

template <int P>
__device__ __forceinline__ void store_cell(double* dst, const double* v) {
  constexpr int NF = (P + 1) * (P + 1);
  if constexpr (StreamCfg<P>::PIECE == 16) {
    double2* p = reinterpret_cast<double2*>(dst);
#pragma unroll
    for (int k = 0; k < NF / 2; ++k) p[k] = make_double2(v[2 * k], v[2 * k + 1]);
  } else {
#pragma unroll
    for (int k = 0; k < NF; ++k) dst[k] = v[k];
  }
}

template <int P>
__global__ void __launch_bounds__(StreamCfg<P>::WARPS * 32, StreamCfg<P>::MINB)
ipdg_apply_stream(const __grid_constant__ IpTab<P> tab, const ApplyArgs a) {
  using C = StreamCfg<P>;
  constexpr int NF = (P + 1) * (P + 1);
  constexpr int NST = C::NST;
  constexpr int STAGE_BYTES = 32 * C::STRIDE;
  extern __shared__ __align__(16) char smem[];

  const int lane = threadIdx.x & 31;
  const int warp = threadIdx.x >> 5;
  const int nstrips = (a.ny + kStripCells - 1) / kStripCells;
  const int ngroups = (nstrips + C::WARPS - 1) / C::WARPS;
  const int group = blockIdx.x % ngroups;
  const int seg = blockIdx.x / ngroups;
  const int strip = group * C::WARPS + warp;
  if (strip >= nstrips) return;

  char* ring = smem + (size_t)warp * NST * STAGE_BYTES;
  const int j0 = strip * kStripCells;  // first output cell of the strip
  const int j = j0 - 1 + lane;         // this lane's cell
  const bool outlane = lane >= 1 && lane <= kStripCells && j < a.ny;
  const int jlo = max(j0 - 1, 0), jhi = min(j0 + kStripCells + 1, a.ny);
  const int nchunk = (jhi - jlo) * C::CELLB / C::PIECE;  // pieces in one column chunk
  const int rel0 = jlo - (j0 - 1);                        // first staged cell, lane units

  // column range of this segment
  const int c0 = (int)((int64_t)a.nx * seg / a.nseg);
  const int c1 = (int)((int64_t)a.nx * (seg + 1) / a.nseg);
  if (c0 >= c1) return;

  auto column_ptr = [&](int c) -> const double* {
    if (c < 0) return a.x_lo;
    if (c >= a.nx) return a.x_hi;
    return a.x + (size_t)c * a.ny * NF;
  };
  auto issue = [&](int c) {
    // stage column c (if it exists and lies in [c0-1, c1]) into ring slot c mod NST
    if (c <= c1 && c >= -1 && c <= a.nx) {
      const double* col = column_ptr(c);
      if (col != nullptr) {
        const char* src = reinterpret_cast<const char*>(col + (size_t)jlo * NF);
        char* dst = ring + ((c + NST) % NST) * STAGE_BYTES;
        for (int k = lane; k < nchunk; k += 32) {
          const int byte = k * C::PIECE;
          const int cellrel = rel0 + byte / C::CELLB;
          const int within = byte % C::CELLB;
          if constexpr (C::PIECE == 16)
            cp_async16(dst + cellrel * C::STRIDE + within, src + byte);
          else
            cp_async8(dst + cellrel * C::STRIDE + within, src + byte);
        }
      }
    }
    cp_async_commit();
  };
  auto col_exists = [&](int c) -> bool {
    return (c >= 0 && c < a.nx) || (c == -1 && a.x_lo != nullptr) ||
           (c == a.nx && a.x_hi != nullptr);
  };
  auto lane_phase = [&](int c) -> int {
    // phase of this lane's cell in column c (0 outside the mesh)
    const int jj = min(max(j, -1), a.ny);
    const int cc = min(max(c, -1), a.nx);
    return phase_at(a.phase, a.ny, cc, jj);
  };

  // prologue: fill the ring with columns c0-1 .. c0+NST-2
#pragma unroll
  for (int k = 0; k < NST; ++k) issue(c0 - 1 + k);

  int sL = lane_phase(c0 - 1), s = lane_phase(c0), sR = lane_phase(c0 + 1);
  int sRR = lane_phase(c0 + 2);
  if (a.x_lo == nullptr && c0 == 0) sL = 0;

  double X[NF], XL[NF];
  cp_async_wait<NST - 2>();
  __syncwarp();
  {
    const char* st = ring + ((c0 - 1 + NST) % NST) * STAGE_BYTES + lane * C::STRIDE;
    if (col_exists(c0 - 1) && sL != 0) {
      load_cell<P>(st, XL);
    } else {
#pragma unroll
      for (int k = 0; k < NF; ++k) XL[k] = 0.0;
    }
    st = ring + ((c0 + NST) % NST) * STAGE_BYTES + lane * C::STRIDE;
    if (s != 0) {
      load_cell<P>(st, X);
    } else {
#pragma unroll
      for (int k = 0; k < NF; ++k) X[k] = 0.0;
    }
  }
  __syncwarp();

  for (int c = c0; c < c1; ++c) {
    issue(c + NST - 1);  // reuses the slot of column c-1, which nobody reads any more
    cp_async_wait<NST - 2>();
    __syncwarp();
    const int sNext = lane_phase(c + 3);  // prefetched two iterations ahead

    double XR[NF], XD[NF], XU[NF], out[NF];
    const char* stc = ring + ((c + NST) % NST) * STAGE_BYTES;
    const char* str = ring + ((c + 1 + NST) % NST) * STAGE_BYTES;
    const int sD = __shfl_up_sync(0xffffffffu, s, 1);
    const int sU = __shfl_down_sync(0xffffffffu, s, 1);
    if (sR != 0 && col_exists(c + 1)) {
      load_cell<P>(str + lane * C::STRIDE, XR);
    } else {
#pragma unroll
      for (int k = 0; k < NF; ++k) XR[k] = 0.0;
    }
    if (outlane && sD != 0) {
      load_cell<P>(stc + (lane - 1) * C::STRIDE, XD);
    } else {
#pragma unroll
      for (int k = 0; k < NF; ++k) XD[k] = 0.0;
    }
    if (outlane && sU != 0) {
      load_cell<P>(stc + (lane + 1) * C::STRIDE, XU);
    } else {
#pragma unroll
      for (int k = 0; k < NF; ++k) XU[k] = 0.0;
    }
    const int sRe = col_exists(c + 1) ? sR : 0;
    const int sLe = col_exists(c - 1) ? sL : 0;
    const int tL = ip_face_type(s, sLe), tR = ip_face_type(s, sRe);
    const int tD = ip_face_type(s, sD), tU = ip_face_type(s, sU);
    const bool plain = !outlane || (tL == T_SAME && tR == T_SAME && tD == T_SAME && tU == T_SAME);
    const int sref = __shfl_sync(0xffffffffu, s, 1);
    const bool uniform = __all_sync(0xffffffffu, plain && (!outlane || s == sref));
    if (uniform) {
      if (sref > 0)
        ip_cell_rows<P>(tab, 0, T_SAME, T_SAME, T_SAME, T_SAME, X, XL, XR, XD, XU, out);
      else
        ip_cell_rows<P>(tab, 1, T_SAME, T_SAME, T_SAME, T_SAME, X, XL, XR, XD, XU, out);
    } else if (outlane) {
      ip_cell_rows<P>(tab, s > 0 ? 0 : 1, tL, tR, tD, tU, X, XL, XR, XD, XU, out);
    }
    if (outlane) store_cell<P>(a.y + ((size_t)c * a.ny + j) * NF, out);
#pragma unroll
    for (int k = 0; k < NF; ++k) {
      XL[k] = X[k];
      X[k] = XR[k];
    }
    sL = s;
    s = sR;
    sR = sRR;
    sRR = sNext;
    __syncwarp();
  }
  cp_async_wait<0>();
}

// --------------------------------------------------------------------------
// host side
// --------------------------------------------------------------------------
template <int P>
static int launch_apply(const IpdgHandle* h, const ApplyArgs& args0, int algo, cudaStream_t st) {
  using C = StreamCfg<P>;
  const IpTab<P>& tab = *static_cast<const IpTab<P>*>(h->tab_host);
  ApplyArgs args = args0;
  if (algo == 1) {
    const int64_t ncells = (int64_t)h->nx * h->ny;
    int blocks = (int)std::min<int64_t>((ncells + 127) / 128, (int64_t)sm_count() * 32);
    ipdg_apply_simple<P><<<blocks, 128, 0, st>>>(tab, args);
  } else {
    const int nstrips = (h->ny + kStripCells - 1) / kStripCells;
    const int ngroups = (nstrips + C::WARPS - 1) / C::WARPS;
    // aim at MINB resident CTAs per SM in one wave; segments no shorter than 16 columns
    int nseg = std::max(1, (sm_count() * C::MINB) / ngroups);
    nseg = std::min(nseg, std::max(1, h->nx / 16));
    args.nseg = nseg;
    const size_t smem = (size_t)C::WARPS * C::NST * 32 * C::STRIDE;
    static bool attr_set = false;
    if (!attr_set) {
      STEFAN_CUDA(cudaFuncSetAttribute(ipdg_apply_stream<P>,
                                       cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
      attr_set = true;
    }
    ipdg_apply_stream<P><<<ngroups * nseg, C::WARPS * 32, smem, st>>>(tab, args);
  }
  STEFAN_CUDA(cudaGetLastError());
  return ST_OK;
}

}  // namespace stefan

using namespace stefan;

extern "C" {

int stefan_ipdg_create(const stefan_ipdg_params* p, const int8_t* phase, const int8_t* phase_lo,
                       const int8_t* phase_hi, stefan_ipdg** out) {
  STEFAN_REQUIRE(p && out, "stefan_ipdg_create: null argument");
  STEFAN_REQUIRE(p->nx >= 1 && p->ny >= 1, "stefan_ipdg_create: nx, ny must be >= 1");
  STEFAN_REQUIRE(p->order >= 1 && p->order <= 3, "stefan_ipdg_create: order must be 1..3");
  STEFAN_REQUIRE(p->numqp >= 1 && p->numqp <= kMaxQ, "stefan_ipdg_create: numqp must be 1..%d", kMaxQ);
  STEFAN_REQUIRE(p->hx > 0 && p->hy > 0, "stefan_ipdg_create: cell widths must be positive");
  STEFAN_REQUIRE(p->variant == 0 || p->variant == 1, "stefan_ipdg_create: unknown variant");
  IpdgHandle* h = new IpdgHandle();
  h->nx = p->nx;
  h->ny = p->ny;
  h->order = p->order;
  h->numqp = p->numqp;
  h->hx = p->hx;
  h->hy = p->hy;
  h->phys = IpPhys{p->k1, p->k2, p->penalty, p->lambda, p->variant, p->terms};
  h->Tm = p->Tm;
  h->has_lo = phase_lo != nullptr;
  h->has_hi = phase_hi != nullptr;
  h->phase_dev = nullptr;
  h->tab_host = nullptr;
  const int NF = (p->order + 1) * (p->order + 1);
  h->ndofs = (int64_t)p->nx * p->ny * NF;
  build_ref1d(p->order, p->numqp, &h->ref);
  int rc = 0;
  const double jx = 0.5 * p->hx, jy = 0.5 * p->hy;
  switch (p->order) {
    case 1: { auto* t = new IpTab<1>(); rc = build_ip_tables<1>(h->ref, jx, jy, h->phys, t); h->tab_host = t; break; }
    case 2: { auto* t = new IpTab<2>(); rc = build_ip_tables<2>(h->ref, jx, jy, h->phys, t); h->tab_host = t; break; }
    case 3: { auto* t = new IpTab<3>(); rc = build_ip_tables<3>(h->ref, jx, jy, h->phys, t); h->tab_host = t; break; }
  }
  if (rc) { delete h; return fail(ST_BAD_ARGUMENT, "stefan_ipdg_create: table build failed (%d)", rc); }
  // padded phase array
  const size_t pn = (size_t)(p->nx + 2) * (p->ny + 2);
  std::vector<int8_t> ph(pn, 0);
  for (int i = 0; i < p->nx; ++i)
    for (int j = 0; j < p->ny; ++j) {
      int8_t s = phase ? phase[(size_t)i * p->ny + j] : 1;
      if (s != 1 && s != -1) { delete h; return fail(ST_BAD_ARGUMENT, "stefan_ipdg_create: phase[%d,%d] = %d (expected +1/-1; cut cells are not supported)", i, j, (int)s); }
      ph[(size_t)(i + 1) * (p->ny + 2) + (j + 1)] = s;
    }
  for (int j = 0; j < p->ny; ++j) {
    if (phase_lo) ph[(size_t)0 * (p->ny + 2) + (j + 1)] = phase_lo[j];
    if (phase_hi) ph[(size_t)(p->nx + 1) * (p->ny + 2) + (j + 1)] = phase_hi[j];
  }
  cudaError_t e = cudaMalloc(&h->phase_dev, pn);
  if (e == cudaSuccess) e = cudaMemcpy(h->phase_dev, ph.data(), pn, cudaMemcpyHostToDevice);
  if (e != cudaSuccess) { delete h; return fail(ST_CUDA_ERROR, "stefan_ipdg_create: %s", cudaGetErrorString(e)); }
  *out = reinterpret_cast<stefan_ipdg*>(h);
  return ST_OK;
}

int stefan_ipdg_destroy(stefan_ipdg* hv) {
  IpdgHandle* h = reinterpret_cast<IpdgHandle*>(hv);
  if (!h) return ST_OK;
  if (h->phase_dev) cudaFree(h->phase_dev);
  switch (h->order) {
    case 1: delete static_cast<IpTab<1>*>(h->tab_host); break;
    case 2: delete static_cast<IpTab<2>*>(h->tab_host); break;
    case 3: delete static_cast<IpTab<3>*>(h->tab_host); break;
  }
  delete h;
  return ST_OK;
}

int64_t stefan_ipdg_ndofs(const stefan_ipdg* hv) {
  return hv ? reinterpret_cast<const IpdgHandle*>(hv)->ndofs : -1;
}

// y = A x for the local slab.  x_lo / x_hi: the neighbouring slabs' adjacent cell
// columns (ny*NF doubles each, device or peer-device memory), required iff the
// handle was created with phase_lo / phase_hi.  algo: 0 = stream kernel, 1 = simple.
int stefan_ipdg_apply(stefan_ipdg* hv, const double* x, double* y, const double* x_lo,
                      const double* x_hi, int32_t algo, void* stream) {
  IpdgHandle* h = reinterpret_cast<IpdgHandle*>(hv);
  STEFAN_REQUIRE(h && x && y, "stefan_ipdg_apply: null argument");
  STEFAN_REQUIRE(x != y, "stefan_ipdg_apply: in-place application is not supported");
  STEFAN_REQUIRE(h->has_lo == (x_lo != nullptr) && h->has_hi == (x_hi != nullptr),
                 "stefan_ipdg_apply: halo columns must be given exactly where the handle has halo phases");
  STEFAN_REQUIRE((reinterpret_cast<uintptr_t>(x) & 15) == 0 && (reinterpret_cast<uintptr_t>(y) & 15) == 0 &&
                 (reinterpret_cast<uintptr_t>(x_lo) & 15) == 0 && (reinterpret_cast<uintptr_t>(x_hi) & 15) == 0,
                 "stefan_ipdg_apply: device pointers must be 16-byte aligned");
  STEFAN_REQUIRE(algo == 0 || algo == 1, "stefan_ipdg_apply: unknown algo %d", algo);
  ApplyArgs a{x, x_lo, x_hi, y, h->phase_dev, h->nx, h->ny, 1};
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  switch (h->order) {
    case 1: return launch_apply<1>(h, a, algo, st);
    case 2: return launch_apply<2>(h, a, algo, st);
    case 3: return launch_apply<3>(h, a, algo, st);
  }
  return fail(ST_UNSUPPORTED, "stefan_ipdg_apply: order %d", h->order);
}

}  // extern "C"
