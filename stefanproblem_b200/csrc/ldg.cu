// Local-DG heat operator (T, q_x, q_y per node) on an uncut uniform mesh: matrix-free
// application for sm_100a and its C-ABI entry points.
//
// Replaces LocalDG.assemble_LDG_linear_system! (StefanDG/LocalDG/LDG_assembly.jl:1-91)
// + CutCellDG.sparse_operator(sysmatrix, mesh, 3) + `matrix * x`.
//
// Data layout in HBM (the reference's dof order): x[3*((i*ny + j)*NF + a) + c],
// c = 0: T, 1: q_x, 2: q_y; cells y-fastest.  Phase: one byte per cell (1: sign +1,
// 2: sign -1), padded to (nx+2) x (ny+2) with 0 = "no cell".
//
// Kernel: one thread per cell; the cell's own 3 NF dofs and, of each neighbour, only
// the face nodes' T and normal flux component (2 (P+1) values per face) are read.
// At the BASELINE size (10 M dofs = 80 MB per vector) x and y fit the 126 MB L2
// together only partly, so the kernel is bound by L2 / HBM streaming, not by the
// 10 (P+1)/3 FMAs per dof of the sum-factorised rows (ldg_tables.h).
#include "../../include/stefan_b200.h"
#include "common.cuh"
#include "ldg_tables.h"

#include <vector>

namespace stefan {

struct LdgHandle {
  int nx, ny, order, numqp;
  LdgPhys phys;
  bool has_lo, has_hi;
  int8_t* phase_dev;
  void* tab_host;
  int64_t ndofs;
};

struct LdgArgs {
  const double* x;
  const double* x_lo;
  const double* x_hi;
  double* y;
  const int8_t* phase;
  int nx, ny;
};

__device__ __forceinline__ int ldg_phase_at(const int8_t* __restrict__ ph, int ny, int i, int j) {
  return ph[(size_t)(i + 1) * (ny + 2) + (j + 1)];
}

template <int P>
__global__ void __launch_bounds__(128)
ldg_apply_cells(const __grid_constant__ LdgTab<P> tab, const LdgArgs a) {
  constexpr int N1 = P + 1;
  constexpr int NF = N1 * N1;
  constexpr int CD = 3 * NF;  // doubles per cell
  const int64_t ncells = (int64_t)a.nx * a.ny;
  for (int64_t cell = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; cell < ncells;
       cell += (int64_t)gridDim.x * blockDim.x) {
    const int i = (int)(cell / a.ny), j = (int)(cell % a.ny);
    const int s = ldg_phase_at(a.phase, a.ny, i, j);
    const int sL = ldg_phase_at(a.phase, a.ny, i - 1, j), sR = ldg_phase_at(a.phase, a.ny, i + 1, j);
    const int sD = ldg_phase_at(a.phase, a.ny, i, j - 1), sU = ldg_phase_at(a.phase, a.ny, i, j + 1);
    double U[CD], UL[CD], UR[CD], UD[CD], UU[CD], out[CD];
    const double* xc = a.x + cell * CD;
    const double* pl = (i > 0) ? xc - (size_t)a.ny * CD : a.x_lo + (size_t)j * CD;
    const double* pr = (i < a.nx - 1) ? xc + (size_t)a.ny * CD : a.x_hi + (size_t)j * CD;
#pragma unroll
    for (int k = 0; k < CD; ++k) U[k] = xc[k];
    // of the neighbours only T and the normal flux component on the shared face
#pragma unroll
    for (int t = 0; t < N1; ++t) {
      const int nl = 3 * (P * N1 + t), nr = 3 * (0 * N1 + t);  // left nbr: ix = P; right: ix = 0
      UL[nl] = sL ? pl[nl] : 0.0;
      UL[nl + 1] = sL ? pl[nl + 1] : 0.0;
      UR[nr] = sR ? pr[nr] : 0.0;
      UR[nr + 1] = sR ? pr[nr + 1] : 0.0;
      const int nd = 3 * (t * N1 + P), nu = 3 * (t * N1 + 0);  // down nbr: iy = P; up: iy = 0
      UD[nd] = sD ? xc[nd - CD] : 0.0;
      UD[nd + 2] = sD ? xc[nd + 2 - CD] : 0.0;
      UU[nu] = sU ? xc[nu + CD] : 0.0;
      UU[nu + 2] = sU ? xc[nu + 2 + CD] : 0.0;
    }
    ldg_cell_rows<P>(tab, s == 1 ? 0 : 1, ip_face_type(s, sL), ip_face_type(s, sR),
                     ip_face_type(s, sD), ip_face_type(s, sU), U, UL, UR, UD, UU, out);
    double* yc = a.y + cell * CD;
#pragma unroll
    for (int k = 0; k < CD; ++k) yc[k] = out[k];
  }
}

template <int P>
static int ldg_launch(const LdgHandle* h, const LdgArgs& a, cudaStream_t st) {
  const LdgTab<P>& tab = *static_cast<const LdgTab<P>*>(h->tab_host);
  const int64_t ncells = (int64_t)h->nx * h->ny;
  const int blocks = (int)std::min<int64_t>((ncells + 127) / 128, (int64_t)sm_count() * 64);
  ldg_apply_cells<P><<<blocks, 128, 0, st>>>(tab, a);
  STEFAN_CUDA(cudaGetLastError());
  return ST_OK;
}

}  // namespace stefan

using namespace stefan;

extern "C" {

int stefan_ldg_create(const stefan_ldg_params* p, const int8_t* phase, const int8_t* phase_lo,
                      const int8_t* phase_hi, stefan_ldg** out) {
  STEFAN_REQUIRE(p && out, "stefan_ldg_create: null argument");
  STEFAN_REQUIRE(p->nx >= 1 && p->ny >= 1, "stefan_ldg_create: nx, ny must be >= 1");
  STEFAN_REQUIRE(p->order >= 1 && p->order <= 3, "stefan_ldg_create: order must be 1..3");
  STEFAN_REQUIRE(p->numqp >= 1 && p->numqp <= kMaxQ, "stefan_ldg_create: numqp must be 1..%d", kMaxQ);
  STEFAN_REQUIRE(p->hx > 0 && p->hy > 0, "stefan_ldg_create: cell widths must be positive");
  LdgHandle* h = new LdgHandle();
  h->nx = p->nx;
  h->ny = p->ny;
  h->order = p->order;
  h->numqp = p->numqp;
  h->phys = LdgPhys{p->k1, p->k2, p->interiorpenalty, p->negboundarypenalty, p->posboundarypenalty,
                    p->V0x, p->V0y};
  h->has_lo = phase_lo != nullptr;
  h->has_hi = phase_hi != nullptr;
  h->phase_dev = nullptr;
  h->tab_host = nullptr;
  const int NF = (p->order + 1) * (p->order + 1);
  h->ndofs = (int64_t)p->nx * p->ny * NF * 3;
  RefElem1D ref;
  build_ref1d(p->order, p->numqp, &ref);
  const double jx = 0.5 * p->hx, jy = 0.5 * p->hy;
  int rc = 0;
  switch (p->order) {
    case 1: { auto* t = new LdgTab<1>(); rc = build_ldg_tables<1>(ref, jx, jy, h->phys, t); h->tab_host = t; break; }
    case 2: { auto* t = new LdgTab<2>(); rc = build_ldg_tables<2>(ref, jx, jy, h->phys, t); h->tab_host = t; break; }
    case 3: { auto* t = new LdgTab<3>(); rc = build_ldg_tables<3>(ref, jx, jy, h->phys, t); h->tab_host = t; break; }
  }
  if (rc) { delete h; return fail(ST_BAD_ARGUMENT, "stefan_ldg_create: table build failed (%d)", rc); }
  const size_t pn = (size_t)(p->nx + 2) * (p->ny + 2);
  std::vector<int8_t> ph(pn, 0);
  for (int i = 0; i < p->nx; ++i)
    for (int j = 0; j < p->ny; ++j) {
      const int8_t s = phase ? phase[(size_t)i * p->ny + j] : 1;
      if (s != 1 && s != -1) {
        delete h;
        return fail(ST_BAD_ARGUMENT, "stefan_ldg_create: phase[%d,%d] = %d (expected +1/-1; cut cells are not supported)",
                    i, j, (int)s);
      }
      ph[(size_t)(i + 1) * (p->ny + 2) + (j + 1)] = (s == 1) ? 1 : 2;
    }
  for (int j = 0; j < p->ny; ++j) {
    if (phase_lo) ph[(size_t)(j + 1)] = (phase_lo[j] == 1) ? 1 : 2;
    if (phase_hi) ph[(size_t)(p->nx + 1) * (p->ny + 2) + (j + 1)] = (phase_hi[j] == 1) ? 1 : 2;
  }
  cudaError_t e = cudaMalloc(&h->phase_dev, pn);
  if (e == cudaSuccess) e = cudaMemcpy(h->phase_dev, ph.data(), pn, cudaMemcpyHostToDevice);
  if (e != cudaSuccess) { delete h; return fail(ST_CUDA_ERROR, "stefan_ldg_create: %s", cudaGetErrorString(e)); }
  *out = reinterpret_cast<stefan_ldg*>(h);
  return ST_OK;
}

int stefan_ldg_destroy(stefan_ldg* hv) {
  LdgHandle* h = reinterpret_cast<LdgHandle*>(hv);
  if (!h) return ST_OK;
  if (h->phase_dev) cudaFree(h->phase_dev);
  switch (h->order) {
    case 1: delete static_cast<LdgTab<1>*>(h->tab_host); break;
    case 2: delete static_cast<LdgTab<2>*>(h->tab_host); break;
    case 3: delete static_cast<LdgTab<3>*>(h->tab_host); break;
  }
  delete h;
  return ST_OK;
}

int64_t stefan_ldg_ndofs(const stefan_ldg* hv) {
  return hv ? reinterpret_cast<const LdgHandle*>(hv)->ndofs : -1;
}

int stefan_ldg_apply(stefan_ldg* hv, const double* x, double* y, const double* x_lo, const double* x_hi,
                     void* stream) {
  LdgHandle* h = reinterpret_cast<LdgHandle*>(hv);
  STEFAN_REQUIRE(h && x && y, "stefan_ldg_apply: null argument");
  STEFAN_REQUIRE(x != y, "stefan_ldg_apply: in-place application is not supported");
  STEFAN_REQUIRE(h->has_lo == (x_lo != nullptr) && h->has_hi == (x_hi != nullptr),
                 "stefan_ldg_apply: halo columns must be given exactly where the handle has halo phases");
  LdgArgs a{x, x_lo, x_hi, y, h->phase_dev, h->nx, h->ny};
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  switch (h->order) {
    case 1: return ldg_launch<1>(h, a, st);
    case 2: return ldg_launch<2>(h, a, st);
    case 3: return ldg_launch<3>(h, a, st);
  }
  return fail(ST_UNSUPPORTED, "stefan_ldg_apply: order %d", h->order);
}

}  // extern "C"
