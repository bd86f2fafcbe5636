// Two-phase 1-D finite-difference Laplace operator (StefanFD/finite-difference.jl):
// matrix-free application, fused explicit heat step and COO emission for sm_100a.
//
// Row of node idx (0-based; the reference is 1-based), as assembled by
//   assemble_laplace_operator!                finite-difference.jl:71-85
//   assemble_dirichlet_boundary_condition!    finite-difference.jl:87-90
//   assemble_interface_continuity!            finite-difference.jl:111-116
//   * same phase on both sides               -> [1,-2,1]/h^2 on idx-1..idx+1
//   * phase 1, interface ahead                -> [-1,4,-5,2]/h^2 on idx-3..idx
//   * phase 1, interface behind               -> [2,-5,4,-1]/h^2 on idx..idx+3
//   * phase 2 next to the interface           -> no Laplace row
//   * first / last node (if requested)        -> identity
//   * interface-continuity rows (if added)    -> [r,-4r,1+3r,-(1+3s),4s,-s] on idx-2..idx+3
// The row kind is decided once at handle creation (one byte per node, validated on
// the host exactly where Julia would throw a BoundsError) so the kernels are pure
// streaming: 8 B read + 8 B written + 1 B of row kind per node.
#include "../../include/stefan_b200.h"
#include "common.cuh"

#include <algorithm>
#include <cstdlib>

#include <vector>

namespace stefan {

enum FdRow : uint8_t {
  FD_EMPTY = 0,
  FD_CENTRAL = 1,
  FD_BACKWARD = 2,
  FD_FORWARD = 3,
  FD_IDENTITY = 4,
  FD_PHASE2 = 8,      // bit: node belongs to phase 2 (selects kappa2 in the step)
  // bit: interface-continuity rows were added to this node (coefficients in the side table).  They
  // ADD to whatever row the node already has, as the reference's triplets do in sparse()
  // (assemble_interface_continuity!, finite-difference.jl:111-116 only appends).
  FD_CONTINUITY = 16,
};

struct FdContinuity {
  int64_t node;
  double r, s;
};

struct FdHandle {
  int64_t n;
  double h;
  bool has_lo, has_hi;
  std::vector<uint8_t> kind_host;
  uint8_t* kind_dev;
  std::vector<FdContinuity> cont;
  FdContinuity* cont_dev;
  bool cont_dirty;
  bool force_scalar;  // STEFAN_FD_SCALAR=1: always the shared-memory kernel (tests)
};

constexpr int kFdThreads = 256;
constexpr int kFdPerThread = 4;
constexpr int kFdTile = kFdThreads * kFdPerThread;
constexpr int kFdHalo = 4;  // 3 needed; 4 keeps the tile loads 32-byte aligned

// MODE 0: y = K u.  MODE 1: y = u + dt * kappa(phase) * (K u) on Laplace rows, u elsewhere.
// the central row and the explicit update, with the roundings spelled out: the fast path, the general path of the vector
// kernel and the scalar kernel must give the same bits (a slab split moves nodes from one path to the other; tests compare with torch.equal)
__device__ __forceinline__ double fd_central(const double* c, double inv_h2) {
  return __dmul_rn(__dadd_rn(__fma_rn(-2.0, c[0], c[-1]), c[1]), inv_h2);
}
__device__ __forceinline__ double fd_update(double c0, double coef, double t) { return __fma_rn(coef, t, c0); }

template <int MODE>
__global__ void __launch_bounds__(kFdThreads)
fd_apply_kernel(const double* __restrict__ u, const double* __restrict__ u_lo,
                const double* __restrict__ u_hi, double* __restrict__ y,
                const uint8_t* __restrict__ kind, int64_t n, double inv_h2, double dtk1,
                double dtk2, const FdContinuity* __restrict__ cont, int ncont) {
  __shared__ double tile[kFdTile + 2 * kFdHalo];
  const int64_t base = (int64_t)blockIdx.x * kFdTile;
  // stage u[base-4 .. base+tile+4) ; outside [0,n): the neighbouring slab's nodes or 0
  for (int t = threadIdx.x; t < kFdTile + 2 * kFdHalo; t += kFdThreads) {
    const int64_t g = base - kFdHalo + t;
    double v = 0.0;
    if (g >= 0 && g < n)
      v = u[g];
    else if (g < 0 && g >= -3 && u_lo != nullptr)
      v = u_lo[3 + g];  // u_lo holds the neighbour's last three nodes
    else if (g >= n && g < n + 3 && u_hi != nullptr)
      v = u_hi[g - n];
    tile[t] = v;
  }
  __syncthreads();
#pragma unroll
  for (int k = 0; k < kFdPerThread; ++k) {
    const int t = threadIdx.x + k * kFdThreads;
    const int64_t g = base + t;
    if (g >= n) continue;
    const double* c = tile + kFdHalo + t;
    const uint8_t kd = kind[g];
    double r = 0.0;
    switch (kd & 7) {
      case FD_CENTRAL: r = fd_central(c, inv_h2); break;
      case FD_BACKWARD: r = (-c[-3] + 4.0 * c[-2] - 5.0 * c[-1] + 2.0 * c[0]) * inv_h2; break;
      case FD_FORWARD: r = (2.0 * c[0] - 5.0 * c[1] + 4.0 * c[2] - c[3]) * inv_h2; break;
      case FD_IDENTITY: r = c[0]; break;
      default: break;
    }
    if (kd & FD_CONTINUITY) {
      for (int q = 0; q < ncont; ++q)
        if (cont[q].node == g) {
          const double rr = cont[q].r, ss = cont[q].s;
          r += rr * c[-2] - 4.0 * rr * c[-1] + (1.0 + 3.0 * rr) * c[0] - (1.0 + 3.0 * ss) * c[1] +
               4.0 * ss * c[2] - ss * c[3];
        }
    }
    if (MODE == 0) {
      y[g] = r;
    } else {
      const int row = kd & 7;
      const bool laplace = row == FD_CENTRAL || row == FD_BACKWARD || row == FD_FORWARD;
      y[g] = laplace ? fd_update(c[0], (kd & FD_PHASE2) ? dtk2 : dtk1, r) : c[0];
    }
  }
}

// Vector form of the same kernel (production path when u, y are 32-byte aligned): a thread
// owns 4 consecutive nodes -- one 32-byte load, one 32-byte store (whole sectors), the 4
// row kinds in one 32-bit load -- and gets the 3 + 3 stencil neighbours from lanes -+1 by
// warp shuffles; only the first / last lane of a warp reads its outer neighbours from
// global memory.  No shared memory, no barrier; 17 B of HBM traffic per node.
__device__ __forceinline__ void ld_global_v4(const double* p, double* v) {
  asm volatile("ld.global.nc.v4.f64 {%0,%1,%2,%3}, [%4];" : "=d"(v[0]), "=d"(v[1]), "=d"(v[2]), "=d"(v[3]) : "l"(p));
}
__device__ __forceinline__ void st_global_v4f(double* p, const double* v) {
  asm volatile("st.global.v4.f64 [%0], {%1,%2,%3,%4};" ::"l"(p), "d"(v[0]), "d"(v[1]), "d"(v[2]), "d"(v[3]) : "memory");
}

#ifndef FD_PREFETCH
#define FD_PREFETCH 1
#endif
template <int MODE>
__global__ void __launch_bounds__(256)
fd_apply_vec(const double* __restrict__ u, const double* __restrict__ u_lo, const double* __restrict__ u_hi,
             double* __restrict__ y, const uint8_t* __restrict__ kind, int64_t n, double inv_h2, double dtk1,
             double dtk2, const FdContinuity* __restrict__ cont, int ncont) {
  const int lane = threadIdx.x & 31;
  const int64_t nvec = (n + 3) / 4;
  const int64_t nvec_w = (nvec + 31) / 32 * 32;  // whole warps stay in the loop (shuffles)
  auto node = [&](int64_t g) -> double {       // u[g] with the slab halos / zeros outside
    if (g >= 0 && g < n) return u[g];
    if (g < 0 && g >= -3 && u_lo != nullptr) return u_lo[3 + g];
    if (g >= n && g < n + 3 && u_hi != nullptr) return u_hi[g - n];
    return 0.0;
  };
  // the thread's own 4 nodes and row kinds of vector vv
  auto load_own = [&](int64_t vv, double* o, unsigned& kk) {
    const int64_t gg = 4 * vv;
    if (gg + 3 < n) {
      ld_global_v4(u + gg, o);
    } else {
#pragma unroll
      for (int k = 0; k < 4; ++k) o[k] = node(gg + k);
    }
    kk = vv < nvec ? *reinterpret_cast<const unsigned*>(kind + gg) : 0u;
  };
  // One iteration ahead: the loads of the next vector are issued before this one is worked on, so a thread always
  // has a request in flight (round 2; without it the kernel sat at 73-80 % of the HBM peak with 64 KB of loads per
  // SM in flight only at the instant every thread had just issued).
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  int64_t v = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  double cur[4];
  unsigned kcur = 0;
  if (v < nvec_w) load_own(v, cur, kcur);
  for (; v < nvec_w; v += stride) {
    const int64_t g0 = 4 * v;
    double nxt[4] = {0.0, 0.0, 0.0, 0.0};
    unsigned knext = 0;
    if (FD_PREFETCH && v + stride < nvec_w) load_own(v + stride, nxt, knext);
    double w[10];  // u[g0-3 .. g0+6]
#pragma unroll
    for (int k = 0; k < 4; ++k) w[3 + k] = cur[k];
    w[0] = __shfl_up_sync(0xffffffffu, w[4], 1);
    w[1] = __shfl_up_sync(0xffffffffu, w[5], 1);
    w[2] = __shfl_up_sync(0xffffffffu, w[6], 1);
    w[7] = __shfl_down_sync(0xffffffffu, w[3], 1);
    w[8] = __shfl_down_sync(0xffffffffu, w[4], 1);
    w[9] = __shfl_down_sync(0xffffffffu, w[5], 1);
    if (lane == 0) {
#pragma unroll
      for (int k = 0; k < 3; ++k) w[k] = node(g0 - 3 + k);
    }
    if (lane == 31) {
#pragma unroll
      for (int k = 0; k < 3; ++k) w[7 + k] = node(g0 + 4 + k);
    }
    if (v < nvec) {
      const unsigned kinds = kcur;
      const bool any_cont = (kinds & (0x01010101u * FD_CONTINUITY)) != 0u;
      double r[4];
      // fast path: four central rows without continuity terms (all but a handful of vectors of a mesh) -- straight-line
      // code instead of the per-node row dispatch (ncu: the dispatch kept the issue slots 62 % busy)
      if ((kinds & (0x01010101u * (7u | FD_CONTINUITY))) == 0x01010101u * FD_CENTRAL) {
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          const double* c = w + 3 + k;
          const double t = fd_central(c, inv_h2);
          if (MODE == 0) r[k] = t;
          else r[k] = fd_update(c[0], ((kinds >> (8 * k)) & FD_PHASE2) ? dtk2 : dtk1, t);
        }
      } else
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        const double* c = w + 3 + k;
        const unsigned kd = (kinds >> (8 * k)) & 0xffu;
        const unsigned row = kd & 7u;
        double t = 0.0;
        if (row == FD_CENTRAL) t = fd_central(c, inv_h2);
        else if (row == FD_BACKWARD) t = (-c[-3] + 4.0 * c[-2] - 5.0 * c[-1] + 2.0 * c[0]) * inv_h2;
        else if (row == FD_FORWARD) t = (2.0 * c[0] - 5.0 * c[1] + 4.0 * c[2] - c[3]) * inv_h2;
        else if (row == FD_IDENTITY) t = c[0];
        // interface rows are a handful per mesh: one test of the packed word keeps the loop off the common path
        if (any_cont && (kd & FD_CONTINUITY)) {
          for (int q = 0; q < ncont; ++q)
            if (cont[q].node == g0 + k) {
              const double rr = cont[q].r, ss = cont[q].s;
              t += rr * c[-2] - 4.0 * rr * c[-1] + (1.0 + 3.0 * rr) * c[0] - (1.0 + 3.0 * ss) * c[1] +
                   4.0 * ss * c[2] - ss * c[3];
            }
        }
        if (MODE == 0) {
          r[k] = t;
        } else {
          const bool laplace = row == FD_CENTRAL || row == FD_BACKWARD || row == FD_FORWARD;
          r[k] = laplace ? fd_update(c[0], (kd & FD_PHASE2) ? dtk2 : dtk1, t) : c[0];
        }
      }
      if (g0 + 3 < n) {
        st_global_v4f(y + g0, r);
      } else {
#pragma unroll
        for (int k = 0; k < 4; ++k)
          if (g0 + k < n) y[g0 + k] = r[k];
      }
    }
    if (!FD_PREFETCH && v + stride < nvec_w) load_own(v + stride, nxt, knext);
#pragma unroll
    for (int k = 0; k < 4; ++k) cur[k] = nxt[k];
    kcur = knext;
  }
}

// COO triplets in row order; one thread per node writes its (at most 6) entries at a
// precomputed offset.  Values and structure are exactly those the reference appends
// (the order of the triplets is by row, which `sparse()` does not care about).
__global__ void fd_coo_kernel(const uint8_t* __restrict__ kind, const int64_t* __restrict__ ofs,
                              int64_t n, double inv_h2, const FdContinuity* __restrict__ cont,
                              int ncont, int64_t* __restrict__ rows, int64_t* __restrict__ cols,
                              double* __restrict__ vals) {
  const int64_t g = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (g >= n) return;
  int64_t o = ofs[g];
  auto put = [&](int64_t c, double v) {
    rows[o] = g;
    cols[o] = c;
    vals[o] = v;
    ++o;
  };
  switch (kind[g] & 7) {
    case FD_CENTRAL:
      put(g - 1, 1.0 * inv_h2); put(g, -2.0 * inv_h2); put(g + 1, 1.0 * inv_h2);
      break;
    case FD_BACKWARD:
      put(g - 3, -1.0 * inv_h2); put(g - 2, 4.0 * inv_h2); put(g - 1, -5.0 * inv_h2); put(g, 2.0 * inv_h2);
      break;
    case FD_FORWARD:
      put(g, 2.0 * inv_h2); put(g + 1, -5.0 * inv_h2); put(g + 2, 4.0 * inv_h2); put(g + 3, -1.0 * inv_h2);
      break;
    case FD_IDENTITY: put(g, 1.0); break;
    default: break;
  }
  if (kind[g] & FD_CONTINUITY)
    for (int q = 0; q < ncont; ++q)
      if (cont[q].node == g) {
        const double r = cont[q].r, s = cont[q].s;
        put(g - 2, r); put(g - 1, -4.0 * r); put(g, 1.0 + 3.0 * r);
        put(g + 1, -(1.0 + 3.0 * s)); put(g + 2, 4.0 * s); put(g + 3, -s);
      }
}

static int row_nnz(uint8_t kind) {
  switch (kind & 7) {
    case FD_CENTRAL: return 3;
    case FD_BACKWARD:
    case FD_FORWARD: return 4;
    case FD_IDENTITY: return 1;
    default: return 0;
  }
}

static int sync_continuity(FdHandle* h) {
  if (!h->cont_dirty) return ST_OK;
  if (h->cont_dev) cudaFree(h->cont_dev);
  h->cont_dev = nullptr;
  if (!h->cont.empty()) {
    STEFAN_CUDA(cudaMalloc(&h->cont_dev, h->cont.size() * sizeof(FdContinuity)));
    STEFAN_CUDA(cudaMemcpy(h->cont_dev, h->cont.data(), h->cont.size() * sizeof(FdContinuity),
                           cudaMemcpyHostToDevice));
  }
  STEFAN_CUDA(cudaMemcpy(h->kind_dev, h->kind_host.data(), (size_t)h->n, cudaMemcpyHostToDevice));
  h->cont_dirty = false;
  return ST_OK;
}

}  // namespace stefan

using namespace stefan;

extern "C" {

int stefan_fd_create(int64_t numnodes, double stepsize, const int8_t* phaseid, int32_t dirichlet,
                     const int8_t* phase_lo, const int8_t* phase_hi, stefan_fd** out) {
  STEFAN_REQUIRE(out && phaseid, "stefan_fd_create: null argument");
  STEFAN_REQUIRE(numnodes >= 3, "stefan_fd_create: need at least 3 nodes");
  STEFAN_REQUIRE(stepsize > 0, "stefan_fd_create: stepsize must be positive");
  const int64_t n = numnodes;
  // phase with three ghost nodes each side (0 = outside the domain)
  auto ph = [&](int64_t i) -> int {
    if (i >= 0 && i < n) return phaseid[i];
    if (i < 0 && i >= -3 && phase_lo) return phase_lo[3 + i];
    if (i >= n && i < n + 3 && phase_hi) return phase_hi[i - n];
    return 0;
  };
  for (int64_t i = 0; i < n; ++i)
    STEFAN_REQUIRE(phaseid[i] == 1 || phaseid[i] == 2, "stefan_fd_create: phaseid[%lld] = %d (expected 1 or 2)",
                   (long long)i, (int)phaseid[i]);
  FdHandle* h = new FdHandle();
  h->n = n;
  h->h = stepsize;
  h->has_lo = phase_lo != nullptr;
  h->has_hi = phase_hi != nullptr;
  h->kind_dev = nullptr;
  h->force_scalar = getenv("STEFAN_FD_SCALAR") != nullptr && getenv("STEFAN_FD_SCALAR")[0] == '1';
  h->cont_dev = nullptr;
  h->cont_dirty = false;
  h->kind_host.assign((size_t)n, FD_EMPTY);
  // global interior nodes: all but the first / last node of the whole domain
  const int64_t lo = h->has_lo ? 0 : 1, hi = h->has_hi ? n : n - 1;
  for (int64_t idx = lo; idx < hi; ++idx) {
    const int c = ph(idx);
    uint8_t kind = FD_EMPTY;
    if (ph(idx - 1) == c && ph(idx + 1) == c) kind = FD_CENTRAL;  // away_from_interface :55-57
    if (c == 1) {
      bool ahead = false;
      if (ph(idx + 1) != c) {  // interface_is_ahead :59-63 then reads idx-3 .. idx-1
        if (ph(idx - 3) == 0) {
          delete h;
          return fail(ST_OUT_OF_RANGE,
                      "stefan_fd_create: BoundsError: phaseid[%lld] (interface within 3 nodes of the left end)",
                      (long long)(idx - 3 + 1));
        }
        ahead = ph(idx - 3) == c && ph(idx - 2) == c && ph(idx - 1) == c;
      }
      if (ahead) {
        kind = FD_BACKWARD;
      } else if (ph(idx - 1) != c) {  // interface_is_behind :65-69 then reads idx+3 .. idx+1
        if (ph(idx + 3) == 0) {
          delete h;
          return fail(ST_OUT_OF_RANGE,
                      "stefan_fd_create: BoundsError: phaseid[%lld] (interface within 3 nodes of the right end)",
                      (long long)(idx + 3 + 1));
        }
        if (ph(idx + 3) == c && ph(idx + 2) == c && ph(idx + 1) == c) kind = FD_FORWARD;
      }
    }
    h->kind_host[(size_t)idx] = kind;
  }
  if (dirichlet) {
    if (!h->has_lo) h->kind_host[0] = FD_IDENTITY;
    if (!h->has_hi) h->kind_host[(size_t)n - 1] = FD_IDENTITY;
  }
  for (int64_t i = 0; i < n; ++i)
    if (phaseid[i] == 2) h->kind_host[(size_t)i] |= FD_PHASE2;
  cudaError_t e = cudaMalloc(&h->kind_dev, (size_t)n + 4);  // + 4: the vector kernel reads 4 row kinds at once
  if (e == cudaSuccess) e = cudaMemcpy(h->kind_dev, h->kind_host.data(), (size_t)n, cudaMemcpyHostToDevice);
  if (e != cudaSuccess) {
    delete h;
    return fail(ST_CUDA_ERROR, "stefan_fd_create: %s", cudaGetErrorString(e));
  }
  *out = reinterpret_cast<stefan_fd*>(h);
  return ST_OK;
}

int stefan_fd_destroy(stefan_fd* hv) {
  FdHandle* h = reinterpret_cast<FdHandle*>(hv);
  if (!h) return ST_OK;
  if (h->kind_dev) cudaFree(h->kind_dev);
  if (h->cont_dev) cudaFree(h->cont_dev);
  delete h;
  return ST_OK;
}

int stefan_fd_add_interface_continuity(stefan_fd* hv, double r, double s, int64_t nodeid) {
  FdHandle* h = reinterpret_cast<FdHandle*>(hv);
  STEFAN_REQUIRE(h, "stefan_fd_add_interface_continuity: null handle");
  STEFAN_REQUIRE(nodeid - 2 >= 0 && nodeid + 3 < h->n,
                 "stefan_fd_add_interface_continuity: row needs nodes nodeid-2 .. nodeid+3");
  h->cont.push_back(FdContinuity{nodeid, r, s});
  h->kind_host[(size_t)nodeid] |= FD_CONTINUITY;
  h->cont_dirty = true;
  return ST_OK;
}

static int fd_launch(FdHandle* h, int mode, const double* u, double* y, const double* u_lo,
                     const double* u_hi, double dtk1, double dtk2, void* stream) {
  STEFAN_REQUIRE(h && u && y, "stefan_fd: null argument");
  STEFAN_REQUIRE(u != y, "stefan_fd: in-place application is not supported");
  STEFAN_REQUIRE(h->has_lo == (u_lo != nullptr) && h->has_hi == (u_hi != nullptr),
                 "stefan_fd: halo nodes must be given exactly where the handle has halo phases");
  int rc = sync_continuity(h);
  if (rc) return rc;
  const int blocks = (int)((h->n + kFdTile - 1) / kFdTile);
  const double inv_h2 = 1.0 / (h->h * h->h);
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const bool vec = ((reinterpret_cast<uintptr_t>(u) | reinterpret_cast<uintptr_t>(y)) & 31) == 0 && !h->force_scalar;
  if (vec) {
    const int64_t nthreads = ((h->n + 3) / 4 + 31) / 32 * 32;
    static const int per_sm = [] { const char* e = getenv("STEFAN_FD_BLOCKS_PER_SM"); return e ? atoi(e) : 32; }();
    const int vb = (int)std::min<int64_t>((nthreads + 255) / 256, (int64_t)sm_count() * per_sm);
    if (mode == 0)
      fd_apply_vec<0><<<vb, 256, 0, st>>>(u, u_lo, u_hi, y, h->kind_dev, h->n, inv_h2, 0, 0, h->cont_dev,
                                          (int)h->cont.size());
    else
      fd_apply_vec<1><<<vb, 256, 0, st>>>(u, u_lo, u_hi, y, h->kind_dev, h->n, inv_h2, dtk1, dtk2, h->cont_dev,
                                          (int)h->cont.size());
    STEFAN_CUDA(cudaGetLastError());
    return ST_OK;
  }
  if (mode == 0)
    fd_apply_kernel<0><<<blocks, kFdThreads, 0, st>>>(u, u_lo, u_hi, y, h->kind_dev, h->n, inv_h2, 0, 0,
                                                      h->cont_dev, (int)h->cont.size());
  else
    fd_apply_kernel<1><<<blocks, kFdThreads, 0, st>>>(u, u_lo, u_hi, y, h->kind_dev, h->n, inv_h2, dtk1,
                                                      dtk2, h->cont_dev, (int)h->cont.size());
  STEFAN_CUDA(cudaGetLastError());
  return ST_OK;
}

int stefan_fd_apply(stefan_fd* hv, const double* u, double* y, const double* u_lo, const double* u_hi,
                    void* stream) {
  return fd_launch(reinterpret_cast<FdHandle*>(hv), 0, u, y, u_lo, u_hi, 0, 0, stream);
}

int stefan_fd_step(stefan_fd* hv, const double* u, double* unew, const double* u_lo, const double* u_hi,
                   double dt, double kappa1, double kappa2, void* stream) {
  return fd_launch(reinterpret_cast<FdHandle*>(hv), 1, u, unew, u_lo, u_hi, dt * kappa1, dt * kappa2, stream);
}

int64_t stefan_fd_nnz(stefan_fd* hv) {
  FdHandle* h = reinterpret_cast<FdHandle*>(hv);
  if (!h) return -1;
  int64_t nnz = 0;
  for (int64_t i = 0; i < h->n; ++i) nnz += row_nnz(h->kind_host[(size_t)i]);
  return nnz + 6 * (int64_t)h->cont.size();
}

int stefan_fd_assemble_coo(stefan_fd* hv, int64_t* rows, int64_t* cols, double* vals, int64_t capacity,
                           int64_t* nnz_out, void* stream) {
  FdHandle* h = reinterpret_cast<FdHandle*>(hv);
  STEFAN_REQUIRE(h && rows && cols && vals && nnz_out, "stefan_fd_assemble_coo: null argument");
  int rc = sync_continuity(h);
  if (rc) return rc;
  std::vector<int64_t> ofs((size_t)h->n + 1, 0);
  std::vector<int> ncont((size_t)h->n, 0);
  for (const auto& c : h->cont) ncont[(size_t)c.node] += 6;
  for (int64_t i = 0; i < h->n; ++i) ofs[(size_t)i + 1] = ofs[(size_t)i] + row_nnz(h->kind_host[(size_t)i]) + ncont[(size_t)i];
  *nnz_out = ofs[(size_t)h->n];
  if (*nnz_out > capacity)
    return fail(ST_CAPACITY, "stefan_fd_assemble_coo: %lld triplets, capacity %lld", (long long)*nnz_out,
                (long long)capacity);
  int64_t* ofs_dev = nullptr;
  STEFAN_CUDA(cudaMalloc(&ofs_dev, ofs.size() * sizeof(int64_t)));
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  STEFAN_CUDA(cudaMemcpyAsync(ofs_dev, ofs.data(), ofs.size() * sizeof(int64_t), cudaMemcpyHostToDevice, st));
  fd_coo_kernel<<<(int)((h->n + 255) / 256), 256, 0, st>>>(h->kind_dev, ofs_dev, h->n, 1.0 / (h->h * h->h),
                                                          h->cont_dev, (int)h->cont.size(), rows, cols, vals);
  STEFAN_CUDA(cudaGetLastError());
  STEFAN_CUDA(cudaStreamSynchronize(st));
  cudaFree(ofs_dev);
  return ST_OK;
}

}  // extern "C"
