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
// Kernels: ldg_apply_march<P> (ldg_march.cuh) is the production path -- warps march along
// x over TMA-staged column chunks, neighbours enter as face traces; ldg_apply_cells<P>
// (one thread per cell, everything straight from global memory) is the plain second
// opinion the tests compare it with (stefan_ldg_apply_algo, algo 1).
#include "../../include/stefan_b200.h"
#include "common.cuh"
#include "ldg_handle.h"
#include "ldg_tables.h"
#include "march_dev.cuh"

#include <cstdlib>
#include <cstring>
#include <vector>

namespace stefan {

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

// Right-hand side of assemble_LDG_rhs! (LDG_assembly.jl:93-131):
//   T rows: sum V f detjac w  (assemble_source!, LDG_source_term.jl:31-60)
//           + alpha_b sum V g sa w on boundary faces (assemble_boundary_vector_flux_source!
//             :265-321; the reference writes it as -alpha * linear_form(..., (-n.n) sa))
//   q rows: sum IM(V)' n g sa w on boundary faces (assemble_boundary_scalar_flux_source! :152-198)
struct LdgRhsTab {
  int n1, nq;
  double Nq[kMaxQ][kMaxN1], w[kMaxQ];
  double jx, jy;
  double alpha_b[4];  // per face 0 bottom, 1 right, 2 top, 3 left
};

__global__ void __launch_bounds__(128)
ldg_rhs_kernel(const __grid_constant__ LdgRhsTab t, const int8_t* __restrict__ phase, int nx, int ny,
               const double* __restrict__ f_qp, const double* __restrict__ g_qp, double* __restrict__ b) {
  const int n1 = t.n1, nq = t.nq, nf = n1 * n1;
  const double* g_bottom = g_qp;
  const double* g_right = g_bottom + (size_t)nx * nq;
  const double* g_top = g_right + (size_t)ny * nq;
  const double* g_left = g_top + (size_t)nx * nq;
  const int64_t ncells = (int64_t)nx * ny;
  for (int64_t cell = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; cell < ncells;
       cell += (int64_t)gridDim.x * blockDim.x) {
    const int i = (int)(cell / ny), j = (int)(cell % ny);
    double out[3 * kMaxN1 * kMaxN1];
    for (int a = 0; a < 3 * nf; ++a) out[a] = 0.0;
    if (f_qp != nullptr) {
      const double* f = f_qp + cell * (size_t)(nq * nq);
      const double detjac = t.jx * t.jy;
      for (int qx = 0; qx < nq; ++qx)
        for (int qy = 0; qy < nq; ++qy) {
          const double v = f[qx * nq + qy] * t.w[qx] * t.w[qy] * detjac;
          for (int ix = 0; ix < n1; ++ix)
            for (int iy = 0; iy < n1; ++iy) out[3 * (ix * n1 + iy)] += t.Nq[qx][ix] * t.Nq[qy][iy] * v;
        }
    }
    if (g_qp != nullptr)
      for (int face = 0; face < 4; ++face) {
        const int sn = face == 0 ? ldg_phase_at(phase, ny, i, j - 1)
                       : face == 1 ? ldg_phase_at(phase, ny, i + 1, j)
                       : face == 2 ? ldg_phase_at(phase, ny, i, j + 1)
                                   : ldg_phase_at(phase, ny, i - 1, j);
        if (sn != 0) continue;
        const bool xnormal = (face == 1 || face == 3), high = (face == 1 || face == 2);
        const double jt = xnormal ? t.jy : t.jx, n = high ? 1.0 : -1.0;
        const int fnode = high ? n1 - 1 : 0;
        const double* g = face == 0   ? g_bottom + (size_t)i * nq
                          : face == 1 ? g_right + (size_t)j * nq
                          : face == 2 ? g_top + (size_t)i * nq
                                      : g_left + (size_t)j * nq;
        for (int it = 0; it < n1; ++it) {
          double m = 0.0;
          for (int q = 0; q < nq; ++q) m += t.w[q] * t.Nq[q][it] * g[q];
          m *= jt;
          const int a = xnormal ? fnode * n1 + it : it * n1 + fnode;
          out[3 * a + 0] += t.alpha_b[face] * m;
          out[3 * a + (xnormal ? 1 : 2)] += n * m;
        }
      }
    for (int a = 0; a < 3 * nf; ++a) b[cell * 3 * nf + a] = out[a];
  }
}

// The COO seam of the LDG system: block (cell, which) -> CD x CD triplets, which = 0 diag,
// 1 left, 2 right, 3 down, 4 up; existing neighbours only; laid out cell by cell like the IP
// seam.  Every column of a block is the apply row function probed with a unit vector, so the
// triplets are by construction the matrix the apply kernels apply.  One thread per
// (cell, which, column).
template <int P>
__global__ void __launch_bounds__(128)
ldg_coo_kernel(const __grid_constant__ LdgTab<P> tab, const int8_t* __restrict__ phase, int nx, int ny,
               int64_t index_base, int64_t* __restrict__ rows, int64_t* __restrict__ cols,
               double* __restrict__ vals) {
  constexpr int N1 = P + 1, NF = N1 * N1, CD = 3 * NF;
  const int64_t nwork = (int64_t)nx * ny * 5 * CD;
  for (int64_t w = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; w < nwork;
       w += (int64_t)gridDim.x * blockDim.x) {
    const int bcol = (int)(w % CD);
    const int64_t cw = w / CD;
    const int64_t cell = cw / 5;
    const int which = (int)(cw % 5);
    const int i = (int)(cell / ny), j = (int)(cell % ny);
    const bool exists = which == 0 || (which == 1 && i > 0) || (which == 2 && i < nx - 1) ||
                        (which == 3 && j > 0) || (which == 4 && j < ny - 1);
    if (!exists) continue;
    const int64_t per = 1 + (i > 0) + (i < nx - 1);
    int64_t nb = (int64_t)i * (ny + 2 * (int64_t)(ny - 1)) +
                 (int64_t)ny * ((i > 1 ? i - 1 : 0) + (i < nx - 1 ? i : nx - 1));
    nb += (int64_t)j * per + (j > 1 ? j - 1 : 0) + (j < ny - 1 ? j : ny - 1);
    int before = 0;
    if (which > 0) before += 1;
    if (which > 1) before += (i > 0);
    if (which > 2) before += (i < nx - 1);
    if (which > 3) before += (j > 0);
    const int64_t o = (nb + before) * CD * CD + (int64_t)bcol * CD;  // column-major inside the block
    const int s = ldg_phase_at(phase, ny, i, j);
    const int sL = ldg_phase_at(phase, ny, i - 1, j), sR = ldg_phase_at(phase, ny, i + 1, j);
    const int sD = ldg_phase_at(phase, ny, i, j - 1), sU = ldg_phase_at(phase, ny, i, j + 1);
    double U[5][CD], out[CD];
    for (int q = 0; q < 5; ++q)
      for (int k = 0; k < CD; ++k) U[q][k] = 0.0;
    U[which][bcol] = 1.0;
    ldg_cell_rows<P>(tab, s == 1 ? 0 : 1, ip_face_type(s, sL), ip_face_type(s, sR), ip_face_type(s, sD),
                     ip_face_type(s, sU), U[0], U[1], U[2], U[3], U[4], out);
    const int64_t ncell = which == 0 ? cell : which == 1 ? cell - ny : which == 2 ? cell + ny
                          : which == 3 ? cell - 1 : cell + 1;
    for (int a = 0; a < CD; ++a) {
      rows[o + a] = cell * CD + a + index_base;
      cols[o + a] = ncell * CD + bcol + index_base;
      vals[o + a] = out[a];
    }
  }
}

}  // namespace stefan
#include "ldg_march.cuh"
namespace stefan {

// x (order 3: 384-byte cells) as a 2-D tensor {16 doubles, 3 ncells rows of 128 B}, tiles of
// 96 rows = 32 cells, 128-byte swizzle; cuTensorMapEncodeTiled through the runtime
static int ldg_make_tensor_map(const double* x, int64_t ncells, CUtensorMap* out) {
  typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                    const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                    CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
  static EncodeTiledFn enc = nullptr;
  if (enc == nullptr) {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
        q == cudaDriverEntryPointSuccess)
      enc = reinterpret_cast<EncodeTiledFn>(p);
  }
  if (enc == nullptr) return fail(ST_CUDA_ERROR, "cuTensorMapEncodeTiled is not available from this driver");
  const cuuint64_t gdim[2] = {16, (cuuint64_t)(3 * ncells)};
  const cuuint64_t gstr[1] = {128};
  const cuuint32_t box[2] = {16, 96};
  const cuuint32_t estr[2] = {1, 1};
  CUresult r = enc(out, CU_TENSOR_MAP_DATA_TYPE_FLOAT64, 2, const_cast<double*>(x), gdim, gstr, box, estr,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) return fail(ST_CUDA_ERROR, "cuTensorMapEncodeTiled failed (%d)", (int)r);
  return ST_OK;
}

// CTA shape of the march kernel for a mesh (see LdgMarchCfg)
static int ldg_march_variant(int order, int64_t ncells) {
  if (const char* e = getenv("STEFAN_LDG_VARIANT")) return atoi(e) ? 1 : 0;  // tuning override
  if (order == 1) return ncells >= 2000000 ? 1 : 0;
  if (order == 2) return ncells >= 1000000 ? 0 : 1;
  return 1;  // order 3: 4 warps x 2 CTAs (142 / 293 GDOF/s at 457^2 / 2000^2; 8 warps x 1: 121 / 288)
}
static int ldg_march_warps(int order, int variant) {
  (void)order;
  return variant ? 4 : 8;
}

template <int P, int V>
static int ldg_launch_march_v(LdgHandle* h, const LdgArgs& a, cudaStream_t st) {
  using C = LdgMarchCfg<P, V>;
  const LdgTab<P>& tab = *static_cast<const LdgTab<P>*>(h->tab_host);
  const SegCache& sc = h->plan.segments(0, h->nx, sm_count() * C::MINB);
  if (sc.n <= 0) return fail(ST_UNSUPPORTED, "stefan_ldg_apply: more than %d strip groups (ny too large)", kMaxMarchCtas);
  constexpr int smem = LdgMarchSmem<P, V>::TOTAL;
  static PerDeviceFlag attr_set;
  if (!attr_set.here()) {
    STEFAN_CUDA(cudaFuncSetAttribute(ldg_apply_march<P, V>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    attr_set.here() = true;
  }
  if (h->tail_ctr_dev == nullptr) {
    STEFAN_CUDA(cudaMalloc(&h->tail_ctr_dev, kLdgTailRing * 2 * sizeof(unsigned)));
    STEFAN_CUDA(cudaMemset(h->tail_ctr_dev, 0, kLdgTailRing * 2 * sizeof(unsigned)));
  }
  LdgMarchArgs m{a.x, a.x_lo, a.x_hi, a.y, a.phase, a.nx, a.ny, h->plan.fix_dev, h->plan.nfix, 0u, nullptr,
                 h->tail_ctr_dev + 2 * (h->launch_seq++ % kLdgTailRing)};
  static const bool skip_tail = getenv("STEFAN_DEV_SKIP_TAIL") != nullptr;  // timing experiments only (wrong results)
  if (skip_tail) m.nfix = 0;
  CUtensorMap tm;
  std::memset(&tm, 0, sizeof(tm));
  if (C::TENSOR) {
    int rc = ldg_make_tensor_map(a.x, (int64_t)h->nx * h->ny, &tm);
    if (rc) return rc;
  }
  // development only: STEFAN_DEV_CTA_TIMES=<file> dumps per-CTA {group, c0, c1, 4 x globaltimer ns}
  static const char* dbg_path = getenv("STEFAN_DEV_CTA_TIMES");
  if (dbg_path != nullptr) {
    unsigned long long* d = nullptr;
    STEFAN_CUDA(cudaMalloc(&d, (size_t)sc.n * 4 * sizeof(unsigned long long)));
    STEFAN_CUDA(cudaMemset(d, 0, (size_t)sc.n * 4 * sizeof(unsigned long long)));
    m.dbg_times = d;
    ldg_apply_march<P, V><<<sc.n, C::WARPS * 32, smem, st>>>(tab, m, h->plan.runs_dev, h->plan.run_ofs_dev, sc.tab, tm);
    STEFAN_CUDA(cudaStreamSynchronize(st));
    std::vector<unsigned long long> t((size_t)sc.n * 4);
    STEFAN_CUDA(cudaMemcpy(t.data(), d, t.size() * sizeof(unsigned long long), cudaMemcpyDeviceToHost));
    cudaFree(d);
    if (FILE* f = fopen(dbg_path, "w")) {
      for (int b = 0; b < sc.n; ++b)
        fprintf(f, "%d %d %d %llu %llu %llu %llu\n", sc.tab.e[b].x, sc.tab.e[b].y, sc.tab.e[b].z, t[4 * b],
                t[4 * b + 1], t[4 * b + 2], t[4 * b + 3]);
      fclose(f);
    }
    return ST_OK;
  }
  {
    // programmatic stream serialization: see griddep_wait in ipdg.cu (STEFAN_PDL=0 turns it off)
    static const bool pdl = [] { const char* e = getenv("STEFAN_PDL"); return e == nullptr || atoi(e) != 0; }();
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3((unsigned)sc.n);
    cfg.blockDim = dim3(C::WARPS * 32);
    cfg.dynamicSmemBytes = smem;
    cfg.stream = st;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr;
    cfg.numAttrs = pdl ? 1 : 0;
    STEFAN_CUDA(cudaLaunchKernelEx(&cfg, ldg_apply_march<P, V>, tab, m, (const int4*)h->plan.runs_dev,
                                   (const int*)h->plan.run_ofs_dev, sc.tab, tm));
  }
  STEFAN_CUDA(cudaGetLastError());
  return ST_OK;
}
template <int P>
static int ldg_launch_march(LdgHandle* h, const LdgArgs& a, cudaStream_t st) {
  return h->variant ? ldg_launch_march_v<P, 1>(h, a, st) : ldg_launch_march_v<P, 0>(h, a, st);
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
  h->phys_hx = p->hx;
  h->phys_hy = p->hy;
  h->phys = LdgPhys{p->k1, p->k2, p->interiorpenalty, p->negboundarypenalty, p->posboundarypenalty,
                    p->V0x, p->V0y, p->interfacepenalty, p->aligned_interface ? 1 : 0};
  h->has_lo = phase_lo != nullptr;
  h->has_hi = phase_hi != nullptr;
  h->phase_dev = nullptr;
  h->tab_host = nullptr;
  h->tail_ctr_dev = nullptr;
  h->launch_seq = 0;
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
  if ((int64_t)p->nx * p->ny >= (int64_t)INT32_MAX) {
    delete h;
    return fail(ST_UNSUPPORTED, "stefan_ldg_create: more than 2^31 cells per slab");
  }
  cudaError_t e = cudaMalloc(&h->phase_dev, pn);
  if (e == cudaSuccess) e = cudaMemcpy(h->phase_dev, ph.data(), pn, cudaMemcpyHostToDevice);
  h->variant = ldg_march_variant(p->order, (int64_t)p->nx * p->ny);
  // cost of a column in which a strip runs both phase bodies, in quarter columns (MarchPlan::segments).
  // The loops are memory-bound at every order: a least-squares fit of the per-CTA loop times of the
  // 457^2 order-3 mesh (profiles/r2_ldg_cta_times.txt) gives 4.15 us per column, +0.29 us per mixed
  // column and 4.3 us per tile -- the second body hides behind the loads.
  h->plan.mixed_cost_q = p->order == 1 ? 4 : 8;   // sixteenths of a column (sweep: profiles/r2_sweep_mixed_cost.log)
  if (e == cudaSuccess) e = h->plan.build(ph.data(), p->nx, p->ny, ldg_march_warps(p->order, h->variant));
  if (e != cudaSuccess) {
    h->plan.destroy();
    if (h->phase_dev) cudaFree(h->phase_dev);
    delete h;
    return fail(ST_CUDA_ERROR, "stefan_ldg_create: %s", cudaGetErrorString(e));
  }
  *out = reinterpret_cast<stefan_ldg*>(h);
  return ST_OK;
}

int stefan_ldg_destroy(stefan_ldg* hv) {
  LdgHandle* h = reinterpret_cast<LdgHandle*>(hv);
  if (!h) return ST_OK;
  if (h->phase_dev) cudaFree(h->phase_dev);
  if (h->tail_ctr_dev) cudaFree(h->tail_ctr_dev);
  h->plan.destroy();
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

int64_t stefan_ldg_coo_size(const stefan_ldg* hv) {
  const LdgHandle* h = reinterpret_cast<const LdgHandle*>(hv);
  if (!h) return -1;
  const int64_t nx = h->nx, ny = h->ny, CD = 3 * (h->order + 1) * (h->order + 1);
  return CD * CD * (nx * ny + 2 * (nx - 1) * ny + 2 * nx * (ny - 1));
}

int stefan_ldg_assemble_coo(stefan_ldg* hv, int64_t* rows, int64_t* cols, double* vals, int64_t capacity,
                            int32_t index_base, int64_t* nnz_out, void* stream) {
  LdgHandle* h = reinterpret_cast<LdgHandle*>(hv);
  STEFAN_REQUIRE(h && rows && cols && vals && nnz_out, "stefan_ldg_assemble_coo: null argument");
  STEFAN_REQUIRE(!h->has_lo && !h->has_hi, "stefan_ldg_assemble_coo: slab handles are not supported");
  STEFAN_REQUIRE(index_base == 0 || index_base == 1, "stefan_ldg_assemble_coo: index_base must be 0 or 1");
  *nnz_out = stefan_ldg_coo_size(hv);
  if (*nnz_out > capacity)
    return fail(ST_CAPACITY, "stefan_ldg_assemble_coo: %lld triplets, capacity %lld", (long long)*nnz_out,
                (long long)capacity);
  const int CD = 3 * (h->order + 1) * (h->order + 1);
  const int64_t nwork = (int64_t)h->nx * h->ny * 5 * CD;
  const int blocks = (int)std::min<int64_t>((nwork + 127) / 128, (int64_t)sm_count() * 32);
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  switch (h->order) {
    case 1: ldg_coo_kernel<1><<<blocks, 128, 0, st>>>(*static_cast<LdgTab<1>*>(h->tab_host), h->phase_dev, h->nx, h->ny, index_base, rows, cols, vals); break;
    case 2: ldg_coo_kernel<2><<<blocks, 128, 0, st>>>(*static_cast<LdgTab<2>*>(h->tab_host), h->phase_dev, h->nx, h->ny, index_base, rows, cols, vals); break;
    case 3: ldg_coo_kernel<3><<<blocks, 128, 0, st>>>(*static_cast<LdgTab<3>*>(h->tab_host), h->phase_dev, h->nx, h->ny, index_base, rows, cols, vals); break;
  }
  STEFAN_CUDA(cudaGetLastError());
  return ST_OK;
}

int stefan_ldg_rhs(stefan_ldg* hv, const double* f_qp, const double* g_qp, double* b, void* stream) {
  LdgHandle* h = reinterpret_cast<LdgHandle*>(hv);
  STEFAN_REQUIRE(h && b, "stefan_ldg_rhs: null argument");
  RefElem1D ref;
  build_ref1d(h->order, h->numqp, &ref);
  LdgRhsTab t;
  std::memset(&t, 0, sizeof(t));
  t.n1 = h->order + 1;
  t.nq = h->numqp;
  for (int q = 0; q < t.nq; ++q) {
    t.w[q] = ref.qw[q];
    for (int a = 0; a < t.n1; ++a) t.Nq[q][a] = ref.Nq[q][a];
  }
  t.jx = 0.5 * h->phys_hx;
  t.jy = 0.5 * h->phys_hy;
  const double nx_[4] = {0, 1, 0, -1}, ny_[4] = {-1, 0, 1, 0};
  for (int f = 0; f < 4; ++f) {
    const double v0n = h->phys.V0x * nx_[f] + h->phys.V0y * ny_[f];
    t.alpha_b[f] = v0n < 0 ? h->phys.negboundarypenalty : h->phys.posboundarypenalty;
  }
  const int64_t ncells = (int64_t)h->nx * h->ny;
  const int blocks = (int)std::min<int64_t>((ncells + 127) / 128, (int64_t)sm_count() * 32);
  ldg_rhs_kernel<<<blocks, 128, 0, static_cast<cudaStream_t>(stream)>>>(t, h->phase_dev, h->nx, h->ny, f_qp,
                                                                         g_qp, b);
  STEFAN_CUDA(cudaGetLastError());
  return ST_OK;
}

static int ldg_apply_impl(stefan_ldg* hv, const double* x, double* y, const double* x_lo, const double* x_hi,
                          int algo, void* stream) {
  LdgHandle* h = reinterpret_cast<LdgHandle*>(hv);
  STEFAN_REQUIRE(h && x && y, "stefan_ldg_apply: null argument");
  STEFAN_REQUIRE(x != y, "stefan_ldg_apply: in-place application is not supported");
  STEFAN_REQUIRE(h->has_lo == (x_lo != nullptr) && h->has_hi == (x_hi != nullptr),
                 "stefan_ldg_apply: halo columns must be given exactly where the handle has halo phases");
  const uintptr_t hmask = h->order == 2 ? 7 : 15;  // see stefan_ipdg_apply
  STEFAN_REQUIRE((reinterpret_cast<uintptr_t>(x) & 15) == 0 && (reinterpret_cast<uintptr_t>(y) & 15) == 0 &&
                 (reinterpret_cast<uintptr_t>(x_lo) & hmask) == 0 && (reinterpret_cast<uintptr_t>(x_hi) & hmask) == 0,
                 "stefan_ldg_apply: x, y must be 16-byte aligned (halo columns: 16, order 2: 8)");
  LdgArgs a{x, x_lo, x_hi, y, h->phase_dev, h->nx, h->ny};
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  if (algo == 1) {
    switch (h->order) {
      case 1: return ldg_launch<1>(h, a, st);
      case 2: return ldg_launch<2>(h, a, st);
      case 3: return ldg_launch<3>(h, a, st);
    }
  } else {
    switch (h->order) {
      case 1: return ldg_launch_march<1>(h, a, st);
      case 2: return ldg_launch_march<2>(h, a, st);
      case 3: return ldg_launch_march<3>(h, a, st);
    }
  }
  return fail(ST_UNSUPPORTED, "stefan_ldg_apply: order %d", h->order);
}

int stefan_ldg_apply(stefan_ldg* hv, const double* x, double* y, const double* x_lo, const double* x_hi,
                     void* stream) {
  return ldg_apply_impl(hv, x, y, x_lo, x_hi, 0, stream);
}

int stefan_ldg_apply_algo(stefan_ldg* hv, const double* x, double* y, const double* x_lo, const double* x_hi,
                          int32_t algo, void* stream) {
  STEFAN_REQUIRE(algo >= 0 && algo <= 2, "stefan_ldg_apply_algo: unknown algo %d", algo);
  return ldg_apply_impl(hv, x, y, x_lo, x_hi, algo, stream);
}

}  // extern "C"
