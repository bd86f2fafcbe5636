// Interior-penalty DG: right-hand-side assembly and COO emission (the reference's
// `SystemRHS` / `SystemMatrix` seam) on the device.
//
//   stefan_ipdg_rhs           assemble_interior_penalty_rhs!   IP_assembly.jl:82-115
//                             = assemble_boundary_source!      IP_source_term.jl:201-258
//                               (R1 - R2 at :145-156; legacy variant: R1 only)
//                             + assemble_source! / assemble_two_phase_source!
//                                                              IP_source_term.jl:19-91
//                             + assemble_robin_source!         IP_robin_interface_source.jl
//                               (face-level form, :12-25, on unlike-phase mesh faces)
//   stefan_ipdg_assemble_coo  the triplets of assemble_interior_penalty_linear_system!
//                             after `sparse()` has summed duplicates, block by block
//
// The reference evaluates its `sourceterm(x)` / `boundaryfunc(x)` closures at the
// physical quadrature points inside the loops; callbacks cannot cross the C ABI, so
// the host samples them once (same points, same order) and passes the samples.
#include "../../include/stefan_b200.h"
#include "common.cuh"
#include "ipdg_handle.h"
#include "peer_dev.cuh"

namespace stefan {

struct RhsTab {
  int n1, nq;
  double Nq[kMaxQ][kMaxN1];  // N_a(x_q)
  double w[kMaxQ];
  double dNL[kMaxN1], dNR[kMaxN1];
  double jx, jy, k1, k2, penalty, lamTm;  // lamTm = 0.5 lambda Tm if the Robin term is on
  int flux_term;                           // 1: R1 - R2 (current), 0: R1 only (legacy)
};

__device__ __forceinline__ int aux_phase_at(const int8_t* __restrict__ ph, int ny, int i, int j) {
  return ph[(size_t)(i + 1) * (ny + 2) + (j + 1)];
}

// one thread per cell; P is small, so plain loops over the tables
__global__ void __launch_bounds__(128)
ipdg_rhs_kernel(const __grid_constant__ RhsTab t, const int8_t* __restrict__ phase, int nx, int ny,
                const double* __restrict__ f_qp, const double* __restrict__ g_qp,
                double* __restrict__ b) {
  const int64_t ncells = (int64_t)nx * ny;
  const int n1 = t.n1, nq = t.nq, nf = n1 * n1;
  // boundary samples: [bottom nx*nq][right ny*nq][top nx*nq][left ny*nq]
  const double* g_bottom = g_qp;
  const double* g_right = g_bottom + (size_t)nx * nq;
  const double* g_top = g_right + (size_t)ny * nq;
  const double* g_left = g_top + (size_t)nx * nq;
  for (int64_t cell = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; cell < ncells;
       cell += (int64_t)gridDim.x * blockDim.x) {
    const int i = (int)(cell / ny), j = (int)(cell % ny);
    const int s = aux_phase_at(phase, ny, i, j);
    const double k = s == 1 ? t.k1 : t.k2;
    double out[kMaxN1 * kMaxN1];
    for (int a = 0; a < nf; ++a) out[a] = 0.0;
    // volume: sum_q V f detjac w   (linear_form, IP_source_term.jl:97-107)
    if (f_qp != nullptr) {
      const double* f = f_qp + cell * (size_t)(nq * nq);
      const double detjac = t.jx * t.jy;
      for (int qx = 0; qx < nq; ++qx)
        for (int qy = 0; qy < nq; ++qy) {
          const double v = f[qx * nq + qy] * t.w[qx] * t.w[qy] * detjac;
          for (int ix = 0; ix < n1; ++ix)
            for (int iy = 0; iy < n1; ++iy) out[ix * n1 + iy] += t.Nq[qx][ix] * t.Nq[qy][iy] * v;
        }
    }
    // faces: 0 bottom, 1 right, 2 top, 3 left
    for (int face = 0; face < 4; ++face) {
      const int sn = face == 0 ? aux_phase_at(phase, ny, i, j - 1)
                     : face == 1 ? aux_phase_at(phase, ny, i + 1, j)
                     : face == 2 ? aux_phase_at(phase, ny, i, j + 1)
                                 : aux_phase_at(phase, ny, i - 1, j);
      const bool xnormal = (face == 1 || face == 3);
      const double jn = xnormal ? t.jx : t.jy, jt = xnormal ? t.jy : t.jx;
      const bool high = (face == 1 || face == 2);
      const int fnode = high ? n1 - 1 : 0;
      if (sn == 0 && g_qp != nullptr) {
        // boundary: pen sum V g sa w - k sum (G n) g sa w   (IP_source_term.jl:131-159)
        const double* g = face == 0   ? g_bottom + (size_t)i * nq
                          : face == 1 ? g_right + (size_t)j * nq
                          : face == 2 ? g_top + (size_t)i * nq
                                      : g_left + (size_t)j * nq;
        for (int it = 0; it < n1; ++it) {
          double m = 0.0;  // tangential moment of the data
          for (int q = 0; q < nq; ++q) m += t.w[q] * t.Nq[q][it] * g[q];
          m *= jt;
          for (int in = 0; in < n1; ++in) {
            const double gn = (high ? t.dNR[in] : -t.dNL[in]) / jn;  // outward normal derivative
            double v = 0.0;
            if (in == fnode) v += t.penalty * m;
            if (t.flux_term) v -= k * gn * m;
            const int a = xnormal ? in * n1 + it : it * n1 + in;
            out[a] += v;
          }
        }
      } else if (sn != 0 && sn != s && t.lamTm != 0.0) {
        // Robin source on an unlike-phase face: 0.5 lambda Tm sum V sa w
        for (int it = 0; it < n1; ++it) {
          double m = 0.0;
          for (int q = 0; q < nq; ++q) m += t.w[q] * t.Nq[q][it];
          const int a = xnormal ? fnode * n1 + it : it * n1 + fnode;
          out[a] += t.lamTm * jt * m;
        }
      }
    }
    double* bc = b + cell * nf;
    for (int a = 0; a < nf; ++a) bc[a] = out[a];
  }
}

// ---------------------------------------------------------------------------
// COO emission: block (cell, which) -> NF x NF triplets; which = 0 diag, 1 left,
// 2 right, 3 down, 4 up.  Blocks are laid out cell by cell, existing neighbours only.
template <int P>
__global__ void __launch_bounds__(128)
ipdg_coo_kernel(const __grid_constant__ IpTab<P> tab, const int8_t* __restrict__ phase, int nx, int ny,
                int64_t index_base, int64_t* __restrict__ rows, int64_t* __restrict__ cols,
                double* __restrict__ vals) {
  constexpr int N1 = P + 1;
  constexpr int NF = N1 * N1;
  const int64_t nwork = (int64_t)nx * ny * 5;
  for (int64_t w = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; w < nwork;
       w += (int64_t)gridDim.x * blockDim.x) {
    const int64_t cell = w / 5;
    const int which = (int)(w % 5);
    const int i = (int)(cell / ny), j = (int)(cell % ny);
    const bool exists = which == 0 || (which == 1 && i > 0) || (which == 2 && i < nx - 1) ||
                        (which == 3 && j > 0) || (which == 4 && j < ny - 1);
    if (!exists) continue;
    // blocks before this cell: closed form of the sum over earlier cells of (1 + #neighbours)
    const int64_t per = 1 + (i > 0) + (i < nx - 1);
    int64_t nb = (int64_t)i * (ny + 2 * (int64_t)(ny - 1)) +
                 (int64_t)ny * ((i > 1 ? i - 1 : 0) + (i < nx - 1 ? i : nx - 1));
    nb += (int64_t)j * per + (j > 1 ? j - 1 : 0) + (j < ny - 1 ? j : ny - 1);
    int before = 0;  // blocks of this cell ahead of `which`
    if (which > 0) before += 1;
    if (which > 1) before += (i > 0);
    if (which > 2) before += (i < nx - 1);
    if (which > 3) before += (j > 0);
    const int64_t o = (nb + before) * NF * NF;
    const int s = aux_phase_at(phase, ny, i, j);
    const int si = s == 1 ? 0 : 1;
    const int tL = ip_face_type(s, aux_phase_at(phase, ny, i - 1, j));
    const int tR = ip_face_type(s, aux_phase_at(phase, ny, i + 1, j));
    const int tD = ip_face_type(s, aux_phase_at(phase, ny, i, j - 1));
    const int tU = ip_face_type(s, aux_phase_at(phase, ny, i, j + 1));
    const int64_t ncol = which == 0   ? cell
                         : which == 1 ? cell - ny
                         : which == 2 ? cell + ny
                         : which == 3 ? cell - 1
                                      : cell + 1;
    for (int a = 0; a < NF; ++a) {
      const int ix = a / N1, iy = a % N1;
      for (int bb = 0; bb < NF; ++bb) {
        const int bx = bb / N1, by = bb % N1;
        double v;
        if (which == 0) {
          v = tab.My[iy * N1 + by] * tab.D[0][si][tL][tR][ix * N1 + bx] +
              tab.Mx[ix * N1 + bx] * tab.D[1][si][tD][tU][iy * N1 + by];
        } else if (which == 1) {
          const double l = (ix == 0 ? tab.L0[0][si][tL][bx] : (bx == P ? tab.Lc[0][si][tL][ix] : 0.0));
          v = tab.My[iy * N1 + by] * l;
        } else if (which == 2) {
          const double r = (ix == P ? tab.Rp[0][si][tR][bx] : (bx == 0 ? tab.Rc[0][si][tR][ix] : 0.0));
          v = tab.My[iy * N1 + by] * r;
        } else if (which == 3) {
          const double l = (iy == 0 ? tab.L0[1][si][tD][by] : (by == P ? tab.Lc[1][si][tD][iy] : 0.0));
          v = tab.Mx[ix * N1 + bx] * l;
        } else {
          const double r = (iy == P ? tab.Rp[1][si][tU][by] : (by == 0 ? tab.Rc[1][si][tU][iy] : 0.0));
          v = tab.Mx[ix * N1 + bx] * r;
        }
        const int64_t e = o + (int64_t)bb * NF + a;  // column-major inside the block, like vec(M)
        rows[e] = cell * NF + a + index_base;
        cols[e] = ncol * NF + bb + index_base;
        vals[e] = v;
      }
    }
  }
}

}  // namespace stefan

using namespace stefan;

extern "C" {

int stefan_ipdg_rhs(stefan_ipdg* hv, const double* f_qp, const double* g_qp, double* b, void* stream) {
  IpdgHandle* h = reinterpret_cast<IpdgHandle*>(hv);
  STEFAN_REQUIRE(h && b, "stefan_ipdg_rhs: null argument");
  RhsTab t;
  std::memset(&t, 0, sizeof(t));
  t.n1 = h->order + 1;
  t.nq = h->numqp;
  for (int q = 0; q < t.nq; ++q) {
    t.w[q] = h->ref.qw[q];
    for (int a = 0; a < t.n1; ++a) t.Nq[q][a] = h->ref.Nq[q][a];
  }
  for (int a = 0; a < t.n1; ++a) {
    t.dNL[a] = h->ref.dNL[a];
    t.dNR[a] = h->ref.dNR[a];
  }
  t.jx = 0.5 * h->hx;
  t.jy = 0.5 * h->hy;
  t.k1 = h->phys.k1;
  t.k2 = h->phys.k2;
  t.penalty = (h->phys.terms & IP_BOUNDARY_PENALTY) ? h->phys.penalty : 0.0;
  t.lamTm = (h->phys.terms & IP_ROBIN) ? 0.5 * h->phys.lambda * h->Tm : 0.0;
  t.flux_term = ((h->phys.terms & IP_BOUNDARY_FLUX) && h->phys.variant == IP_CURRENT) ? 1 : 0;
  const int64_t ncells = (int64_t)h->nx * h->ny;
  const int blocks = (int)std::min<int64_t>((ncells + 127) / 128, (int64_t)sm_count() * 32);
  ipdg_rhs_kernel<<<blocks, 128, 0, static_cast<cudaStream_t>(stream)>>>(t, h->phase_dev, h->nx, h->ny, f_qp,
                                                                          g_qp, b);
  STEFAN_CUDA(cudaGetLastError());
  return ST_OK;
}

int64_t stefan_ipdg_coo_size(const stefan_ipdg* hv) {
  const IpdgHandle* h = reinterpret_cast<const IpdgHandle*>(hv);
  if (!h) return -1;
  const int64_t nx = h->nx, ny = h->ny, NF = (h->order + 1) * (h->order + 1);
  return NF * NF * (nx * ny + 2 * (nx - 1) * ny + 2 * nx * (ny - 1));
}

int stefan_ipdg_assemble_coo(stefan_ipdg* hv, int64_t* rows, int64_t* cols, double* vals, int64_t capacity,
                             int32_t index_base, int64_t* nnz_out, void* stream) {
  IpdgHandle* h = reinterpret_cast<IpdgHandle*>(hv);
  STEFAN_REQUIRE(h && rows && cols && vals && nnz_out, "stefan_ipdg_assemble_coo: null argument");
  STEFAN_REQUIRE(!h->has_lo && !h->has_hi, "stefan_ipdg_assemble_coo: slab handles are not supported");
  STEFAN_REQUIRE(index_base == 0 || index_base == 1, "stefan_ipdg_assemble_coo: index_base must be 0 or 1");
  *nnz_out = stefan_ipdg_coo_size(hv);
  if (*nnz_out > capacity)
    return fail(ST_CAPACITY, "stefan_ipdg_assemble_coo: %lld triplets, capacity %lld", (long long)*nnz_out,
                (long long)capacity);
  const int64_t nwork = (int64_t)h->nx * h->ny * 5;
  const int blocks = (int)std::min<int64_t>((nwork + 127) / 128, (int64_t)sm_count() * 32);
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  switch (h->order) {
    case 1: ipdg_coo_kernel<1><<<blocks, 128, 0, st>>>(*static_cast<IpTab<1>*>(h->tab_host), h->phase_dev, h->nx, h->ny, index_base, rows, cols, vals); break;
    case 2: ipdg_coo_kernel<2><<<blocks, 128, 0, st>>>(*static_cast<IpTab<2>*>(h->tab_host), h->phase_dev, h->nx, h->ny, index_base, rows, cols, vals); break;
    case 3: ipdg_coo_kernel<3><<<blocks, 128, 0, st>>>(*static_cast<IpTab<3>*>(h->tab_host), h->phase_dev, h->nx, h->ny, index_base, rows, cols, vals); break;
  }
  STEFAN_CUDA(cudaGetLastError());
  return ST_OK;
}

}  // extern "C"

// ---------------------------------------------------------------------------
// Interface sums that drive the front velocity, and the explicit time-step update.
//
// On every mesh face between a +1 cell (side 1, conductivity k1) and a -1 cell (side 2),
// with n the outward normal of side 1, sa = face det-jacobian and the face Gauss rule:
//   S[0] = sum sa w                              length of the interface
//   S[1] = sum lambda (Tm - {T}) sa w            kinetic term  lambda (Tm - T_I)
//   S[2] = sum (k1 dn T1 - k2 dn T2) sa w        jump of the normal heat flux
//   S[3] = sum {T} sa w                          mean interface temperature x length
// {T} = (T1 + T2)/2.  These are the face-integrated forms of what the reference samples
// point-wise after the solve (IP_robin_interface_flux_velocity.jl:213-278: velocity
// -lambda (Tm - T) per side, gvelocity = k2 dn T2 - k1 dn T1; values / gradients at
// reference points as in IP_postprocess.jl:1-52).  Faces are owned by their +1 cell, so a
// column-slab partition sums every face exactly once; the per-rank sums are added with a
// 4-double all-reduce.  Deterministic: one partial per block, then a fixed-order sum.
namespace stefan {

struct IfaceTab {
  int n1, nq;
  double Nq[kMaxQ][kMaxN1], w[kMaxQ], dNL[kMaxN1], dNR[kMaxN1];
  double jx, jy, k1, k2, lambda, Tm;
};

// One launch: one thread per interface face of the handle's list, one partial per block, and the last block to
// finish (atomic ticket) adds the partials in a fixed order -- deterministic -- and resets the ticket.
// Peer form (wait_lo / wait_hi non-null): a thread whose -1 neighbour lies in a halo column first waits until the
// neighbour's flag has reached `epoch` (its kernel that wrote the vector is complete).
struct IfaceSync {
  const unsigned long long* wait_lo;
  const unsigned long long* wait_hi;
  unsigned long long epoch, timeout_ns;
  int* status;
};

__global__ void __launch_bounds__(128)
ipdg_interface_list_kernel(const __grid_constant__ IfaceTab t, const int2* __restrict__ faces, int nfaces, int nx, int ny,
                           const double* __restrict__ u, const double* __restrict__ u_lo,
                           const double* __restrict__ u_hi, const IfaceSync sync, double* __restrict__ partial,
                           unsigned* __restrict__ ticket, double* __restrict__ out) {
  const int n1 = t.n1, nq = t.nq, nf = n1 * n1, P = n1 - 1;
  double s0 = 0, s1 = 0, s2 = 0, s3 = 0;
  for (int k = blockIdx.x * blockDim.x + threadIdx.x; k < nfaces; k += gridDim.x * blockDim.x) {
    const int2 cf = faces[k];
    const int cell = cf.x, face = cf.y;
    const int i = cell / ny, j = cell - i * ny;
    const int di = face == 1 ? 1 : (face == 3 ? -1 : 0), dj = face == 2 ? 1 : (face == 0 ? -1 : 0);
    const double* u1 = u + (size_t)cell * nf;
    const double* u2;
    if (i + di < 0) {
      if (sync.wait_lo) peer_flag_wait(sync.wait_lo, sync.epoch, sync.timeout_ns, sync.status);
      u2 = u_lo + (size_t)j * nf;
    } else if (i + di >= nx) {
      if (sync.wait_hi) peer_flag_wait(sync.wait_hi, sync.epoch, sync.timeout_ns, sync.status);
      u2 = u_hi + (size_t)j * nf;
    } else {
      u2 = u + ((size_t)(i + di) * ny + (j + dj)) * nf;
    }
    const bool xnormal = di != 0, high = (face == 1 || face == 2);
    const double jn = xnormal ? t.jx : t.jy, sa = xnormal ? t.jy : t.jx;
    // nodal traces and outward (side-1) normal derivatives along the face, per tangential node
    for (int q = 0; q < nq; ++q) {
      double T1 = 0, T2 = 0, d1 = 0, d2 = 0;
      for (int it = 0; it < n1; ++it) {
        const double Nt = t.Nq[q][it];
        double a1 = 0, a2 = 0;  // normal derivative of side 1 / side 2 at tangential node it
        for (int in = 0; in < n1; ++in) {
          const int idx1 = xnormal ? in * n1 + it : it * n1 + in;
          a1 += (high ? t.dNR[in] : -t.dNL[in]) / jn * u1[idx1];
          a2 += (high ? t.dNL[in] : -t.dNR[in]) / jn * u2[idx1];  // neighbour: opposite end, same direction
        }
        const int f1 = high ? P : 0, f2 = high ? 0 : P;
        T1 += Nt * u1[xnormal ? f1 * n1 + it : it * n1 + f1];
        T2 += Nt * u2[xnormal ? f2 * n1 + it : it * n1 + f2];
        d1 += Nt * a1;
        d2 += Nt * a2;
      }
      const double wq = t.w[q] * sa, Tavg = 0.5 * (T1 + T2);
      s0 += wq;
      s1 += t.lambda * (t.Tm - Tavg) * wq;
      s2 += (t.k1 * d1 - t.k2 * d2) * wq;
      s3 += Tavg * wq;
    }
  }
  // warp shuffle reduction, then one partial per block (fixed order -> deterministic)
  __shared__ double red[4][4];
  __shared__ bool last;
  double v[4] = {s0, s1, s2, s3};
  for (int k = 0; k < 4; ++k)
    for (int off = 16; off > 0; off >>= 1) v[k] += __shfl_down_sync(0xffffffffu, v[k], off);
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  if (lane == 0)
    for (int k = 0; k < 4; ++k) red[warp][k] = v[k];
  __syncthreads();
  if (threadIdx.x < 4) {
    double s = 0;
    for (int w = 0; w < (int)(blockDim.x >> 5); ++w) s += red[w][threadIdx.x];
    partial[(size_t)blockIdx.x * 4 + threadIdx.x] = s;
  }
  __threadfence();
  __syncthreads();
  if (threadIdx.x == 0) last = atomicAdd(ticket, 1u) == gridDim.x - 1;
  __syncthreads();
  if (last && threadIdx.x < 4) {
    __threadfence();
    double s = 0;
    for (unsigned b = 0; b < gridDim.x; ++b) s += partial[(size_t)b * 4 + threadIdx.x];
    out[threadIdx.x] = s;
    if (threadIdx.x == 0) *ticket = 0;
  }
}

// u <- u + dt * Minv (b - y), M = jx jy (Mh (x) Mh) the block-diagonal DG mass matrix:
// the explicit Euler step of M du/dt = b - A u (no reference implementation; definition in
// oracle/timeloop.py).  minv = inverse of the 1-D reference mass matrix.
struct EulerTab {
  int n1;
  double minv[kMaxN1][kMaxN1];
  double scale;  // dt / (jx jy)
};

__global__ void __launch_bounds__(128)
ipdg_euler_kernel(const __grid_constant__ EulerTab t, int64_t ncells, const double* __restrict__ y,
                  const double* __restrict__ b, double* __restrict__ u) {
  const int n1 = t.n1, nf = n1 * n1;
  for (int64_t cell = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; cell < ncells;
       cell += (int64_t)gridDim.x * blockDim.x) {
    double r[kMaxN1 * kMaxN1], tmp[kMaxN1 * kMaxN1];
    for (int a = 0; a < nf; ++a) r[a] = (b ? b[cell * nf + a] : 0.0) - y[cell * nf + a];
    for (int ix = 0; ix < n1; ++ix)
      for (int iy = 0; iy < n1; ++iy) {
        double v = 0;
        for (int c = 0; c < n1; ++c) v += t.minv[iy][c] * r[ix * n1 + c];
        tmp[ix * n1 + iy] = v;
      }
    for (int ix = 0; ix < n1; ++ix)
      for (int iy = 0; iy < n1; ++iy) {
        double v = 0;
        for (int c = 0; c < n1; ++c) v += t.minv[ix][c] * tmp[c * n1 + iy];
        u[cell * nf + ix * n1 + iy] += t.scale * v;
      }
  }
}

}  // namespace stefan

extern "C" {

static int interface_sums_impl(IpdgHandle* h, const double* u, const double* u_lo, const double* u_hi,
                               const stefan_peer_sync* sync, double* sums_dev, cudaStream_t st) {
  IfaceTab t;
  std::memset(&t, 0, sizeof(t));
  t.n1 = h->order + 1;
  t.nq = h->numqp;
  for (int q = 0; q < t.nq; ++q) {
    t.w[q] = h->ref.qw[q];
    for (int a = 0; a < t.n1; ++a) t.Nq[q][a] = h->ref.Nq[q][a];
  }
  for (int a = 0; a < t.n1; ++a) {
    t.dNL[a] = h->ref.dNL[a];
    t.dNR[a] = h->ref.dNR[a];
  }
  t.jx = 0.5 * h->hx;
  t.jy = 0.5 * h->hy;
  t.k1 = h->phys.k1;
  t.k2 = h->phys.k2;
  t.lambda = h->phys.lambda;
  t.Tm = h->Tm;
  const int blocks = std::max(1, std::min((h->niface + 127) / 128, sm_count() * 4));
  if (h->iface_partial_dev == nullptr) {
    STEFAN_CUDA(cudaMalloc(&h->iface_partial_dev, (size_t)sm_count() * 4 * 4 * sizeof(double)));
    STEFAN_CUDA(cudaMalloc(&h->iface_ctr_dev, sizeof(unsigned)));
    STEFAN_CUDA(cudaMemset(h->iface_ctr_dev, 0, sizeof(unsigned)));
  }
  IfaceSync sy{nullptr, nullptr, 0, 0, nullptr};
  if (sync != nullptr) {
    sy.wait_lo = u_lo ? reinterpret_cast<const unsigned long long*>(sync->lo_ready) : nullptr;
    sy.wait_hi = u_hi ? reinterpret_cast<const unsigned long long*>(sync->hi_ready) : nullptr;
    sy.epoch = sync->epoch;
    sy.timeout_ns = (unsigned long long)sync->timeout_ms * 1000000ull;
    sy.status = sync->status;
  }
  ipdg_interface_list_kernel<<<blocks, 128, 0, st>>>(t, h->iface_dev, h->niface, h->nx, h->ny, u, u_lo, u_hi, sy,
                                                     h->iface_partial_dev, h->iface_ctr_dev, sums_dev);
  STEFAN_CUDA(cudaGetLastError());
  return ST_OK;
}

int stefan_ipdg_interface_sums(stefan_ipdg* hv, const double* u, const double* u_lo, const double* u_hi,
                               double* sums_dev, void* stream) {
  IpdgHandle* h = reinterpret_cast<IpdgHandle*>(hv);
  STEFAN_REQUIRE(h && u && sums_dev, "stefan_ipdg_interface_sums: null argument");
  STEFAN_REQUIRE(h->has_lo == (u_lo != nullptr) && h->has_hi == (u_hi != nullptr),
                 "stefan_ipdg_interface_sums: halo columns must be given exactly where the handle has halo phases");
  return interface_sums_impl(h, u, u_lo, u_hi, nullptr, sums_dev, static_cast<cudaStream_t>(stream));
}

// The same with the neighbours' columns read from peer memory: the threads that touch a halo column first wait
// until the flag sync->lo_ready / sync->hi_ready points at has reached sync->epoch (point them at DONE-type flags
// or mailboxes: the neighbour's kernel that wrote the vector must be complete).  One launch.
int stefan_ipdg_interface_sums_peer(stefan_ipdg* hv, const double* u, const double* u_lo, const double* u_hi,
                                    const stefan_peer_sync* sync, double* sums_dev, void* stream) {
  IpdgHandle* h = reinterpret_cast<IpdgHandle*>(hv);
  STEFAN_REQUIRE(h && u && sums_dev && sync, "stefan_ipdg_interface_sums_peer: null argument");
  STEFAN_REQUIRE(h->has_lo == (u_lo != nullptr) && h->has_hi == (u_hi != nullptr),
                 "stefan_ipdg_interface_sums_peer: halo columns must be given exactly where the handle has halo phases");
  STEFAN_REQUIRE(sync->timeout_ms > 0 && (!u_lo || sync->lo_ready) && (!u_hi || sync->hi_ready),
                 "stefan_ipdg_interface_sums_peer: incomplete sync block");
  return interface_sums_impl(h, u, u_lo, u_hi, sync, sums_dev, static_cast<cudaStream_t>(stream));
}

int stefan_ipdg_euler_update(stefan_ipdg* hv, double* u, const double* y, const double* b, double dt,
                             void* stream) {
  IpdgHandle* h = reinterpret_cast<IpdgHandle*>(hv);
  STEFAN_REQUIRE(h && u && y, "stefan_ipdg_euler_update: null argument");
  EulerTab t;
  std::memset(&t, 0, sizeof(t));
  const int n = h->order + 1;
  t.n1 = n;
  for (int i = 0; i < n; ++i)
    for (int j = 0; j < n; ++j) t.minv[i][j] = h->minv1d[i][j];
  t.scale = dt / (0.25 * h->hx * h->hy);
  const int64_t ncells = (int64_t)h->nx * h->ny;
  const int blocks = (int)std::min<int64_t>((ncells + 127) / 128, (int64_t)sm_count() * 32);
  ipdg_euler_kernel<<<blocks, 128, 0, static_cast<cudaStream_t>(stream)>>>(t, ncells, y, b, u);
  STEFAN_CUDA(cudaGetLastError());
  return ST_OK;
}

}  // extern "C"
