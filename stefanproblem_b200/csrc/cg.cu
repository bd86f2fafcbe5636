// Solver and error norm on top of the matrix-free IP-DG apply.
//
// The reference solves its (symmetric positive definite: `@test issymmetric`,
// uncut_mesh_tests/test_uncut_IP_convergence.jl:88) interior-penalty systems with a sparse
// direct solve `matrix \ rhs` and then measures `mesh_L2_error` (useful_routines.jl:95-133).
// Here: preconditioned conjugate gradients whose operator is stefan_ipdg_apply, with a
// block-Jacobi preconditioner (the inverse of every cell's NF x NF diagonal block, one
// inverse per (phase, four face types) combination, inverted once on the host), and the L2
// error of a dof vector against samples of the exact solution at the cell quadrature points.
//
// Everything in the iteration stays on the device: the scalars alpha / beta live in device
// memory and are produced by the reduction kernels, so an iteration is 1 apply + 3 vector
// kernels + 2 small reductions and the host only reads the residual every `check_every`
// iterations.  Reductions are deterministic (one partial per block, fixed-order final sum).
#include "../../include/stefan_b200.h"
#include "common.cuh"
#include "ipdg_handle.h"
#include "krylov.cuh"

#include <algorithm>
#include <cmath>
#include <cstdlib>
#include <vector>

namespace stefan {

constexpr int kRedThreads = 256;

__device__ __forceinline__ double block_sum(double v) {
  __shared__ double sh[kRedThreads / 32];
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_down_sync(0xffffffffu, v, o);
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  __syncthreads();
  if (lane == 0) sh[w] = v;
  __syncthreads();
  double s = 0.0;
  if (threadIdx.x == 0)
    for (int k = 0; k < kRedThreads / 32; ++k) s += sh[k];
  return s;  // valid in thread 0
}

// x += alpha p, r -= alpha Ap (skipped when first), z = Binv r per cell, partials of r.z and r.r.
// One thread per cell; Binv: [162][NF*NF] inverse diagonal blocks, or nullptr (z = r).
template <int P>
__global__ void __launch_bounds__(kRedThreads)
cg_update_kernel(double* __restrict__ x, double* __restrict__ r, double* __restrict__ z, const double* __restrict__ p,
                 const double* __restrict__ Ap, const double* __restrict__ scal, int first,
                 const double* __restrict__ Binv, const int8_t* __restrict__ phase, int nx, int ny,
                 double* __restrict__ partial) {
  constexpr int NF = (P + 1) * (P + 1);
  const double alpha = first ? 0.0 : scal[kry::S_ALPHA];
  const int64_t ncells = (int64_t)nx * ny;
  double srz = 0.0, srr = 0.0;
  for (int64_t cell = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; cell < ncells;
       cell += (int64_t)gridDim.x * blockDim.x) {
    double rc[NF];
#pragma unroll
    for (int a = 0; a < NF; ++a) {
      const int64_t i = cell * NF + a;
      double rv = r[i];
      if (!first) {
        x[i] += alpha * p[i];
        rv -= alpha * Ap[i];
        r[i] = rv;
      }
      rc[a] = rv;
      srr += rv * rv;
    }
    if (Binv != nullptr) {
      const int ci = (int)(cell / ny), cj = (int)(cell % ny);
      const size_t ld = (size_t)ny + 2;
      const int8_t* ph = phase + (size_t)(ci + 1) * ld + (cj + 1);
      const int s = ph[0];
      const int type = ((((s == 1 ? 0 : 1) * 3 + ip_face_type(s, ph[-(ptrdiff_t)ld])) * 3 + ip_face_type(s, ph[ld])) * 3 +
                        ip_face_type(s, ph[-1])) * 3 + ip_face_type(s, ph[1]);
      const double* B = Binv + (size_t)type * NF * NF;
#pragma unroll
      for (int a = 0; a < NF; ++a) {
        double v = 0.0;
#pragma unroll
        for (int b = 0; b < NF; ++b) v += B[a * NF + b] * rc[b];
        z[cell * NF + a] = v;
        srz += v * rc[a];
      }
    } else {
#pragma unroll
      for (int a = 0; a < NF; ++a) z[cell * NF + a] = rc[a];
      srz = srr;
    }
  }
  srz = block_sum(srz);
  srr = block_sum(srr);
  if (threadIdx.x == 0) {  // (kry::Reducer layout: sum0 = r.z, sum1 = r.r)
    partial[2 * blockIdx.x] = srz;
    partial[2 * blockIdx.x + 1] = srr;
  }
}

// dense inverse by Gauss-Jordan with partial pivoting; false if singular
static bool invert_dense(int n, const double* A, double* inv) {
  std::vector<double> m((size_t)n * 2 * n, 0.0);
  for (int i = 0; i < n; ++i) {
    for (int j = 0; j < n; ++j) m[(size_t)i * 2 * n + j] = A[i * n + j];
    m[(size_t)i * 2 * n + n + i] = 1.0;
  }
  for (int c = 0; c < n; ++c) {
    int piv = c;
    for (int r = c + 1; r < n; ++r)
      if (std::fabs(m[(size_t)r * 2 * n + c]) > std::fabs(m[(size_t)piv * 2 * n + c])) piv = r;
    if (m[(size_t)piv * 2 * n + c] == 0.0) return false;
    if (piv != c)
      for (int j = 0; j < 2 * n; ++j) std::swap(m[(size_t)piv * 2 * n + j], m[(size_t)c * 2 * n + j]);
    const double d = 1.0 / m[(size_t)c * 2 * n + c];
    for (int j = 0; j < 2 * n; ++j) m[(size_t)c * 2 * n + j] *= d;
    for (int r = 0; r < n; ++r) {
      if (r == c) continue;
      const double f = m[(size_t)r * 2 * n + c];
      if (f != 0.0)
        for (int j = 0; j < 2 * n; ++j) m[(size_t)r * 2 * n + j] -= f * m[(size_t)c * 2 * n + j];
    }
  }
  for (int i = 0; i < n; ++i)
    for (int j = 0; j < n; ++j) inv[i * n + j] = m[(size_t)i * 2 * n + n + j];
  return true;
}

// inverse diagonal blocks for every (phase index, tL, tR, tD, tU): the diagonal block of a
// cell is  B[(ix,iy),(bx,by)] = Dx[ix][bx] My[iy][by] + Mx[ix][bx] Dy[iy][by]  (ipdg_tables.h)
template <int P>
static void build_block_inverses(const IpTab<P>& tab, std::vector<double>* out) {
  constexpr int N1 = P + 1, NF = N1 * N1;
  out->assign((size_t)162 * NF * NF, 0.0);
  std::vector<double> B(NF * NF);
  for (int s = 0; s < 2; ++s)
    for (int tL = 0; tL < 3; ++tL)
      for (int tR = 0; tR < 3; ++tR)
        for (int tD = 0; tD < 3; ++tD)
          for (int tU = 0; tU < 3; ++tU) {
            for (int a = 0; a < NF; ++a)
              for (int b = 0; b < NF; ++b) {
                const int ix = a / N1, iy = a % N1, bx = b / N1, by = b % N1;
                B[a * NF + b] = tab.D[0][s][tL][tR][ix * N1 + bx] * tab.My[iy * N1 + by] +
                                tab.Mx[ix * N1 + bx] * tab.D[1][s][tD][tU][iy * N1 + by];
              }
            const int type = (((s * 3 + tL) * 3 + tR) * 3 + tD) * 3 + tU;
            double* inv = out->data() + (size_t)type * NF * NF;
            if (!invert_dense(NF, B.data(), inv)) {
              for (int a = 0; a < NF * NF; ++a) inv[a] = 0.0;
              for (int a = 0; a < NF; ++a) inv[a * NF + a] = 1.0;
            }
          }
}

struct L2Tab {
  int n1, nq;
  double Nq[kMaxQ][kMaxN1], w[kMaxQ];
  double detjac;
};

// sum over cells and quadrature points of (u_h(q) - uex(q))^2 detjac w, and of uex(q)^2 detjac w
__global__ void __launch_bounds__(kRedThreads)
ipdg_l2_error_kernel(const __grid_constant__ L2Tab t, const double* __restrict__ u, const double* __restrict__ uex_qp,
                     int64_t ncells, double* __restrict__ partial_err, double* __restrict__ partial_ref) {
  const int n1 = t.n1, nq = t.nq, nf = n1 * n1;
  double se = 0.0, sr = 0.0;
  for (int64_t cell = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; cell < ncells;
       cell += (int64_t)gridDim.x * blockDim.x) {
    const double* uc = u + cell * nf;
    const double* ue = uex_qp + cell * (int64_t)(nq * nq);
    for (int qx = 0; qx < nq; ++qx)
      for (int qy = 0; qy < nq; ++qy) {
        double v = 0.0;
        for (int ix = 0; ix < n1; ++ix)
          for (int iy = 0; iy < n1; ++iy) v += t.Nq[qx][ix] * t.Nq[qy][iy] * uc[ix * n1 + iy];
        const double e = v - ue[qx * nq + qy], wq = t.w[qx] * t.w[qy] * t.detjac;
        se += e * e * wq;
        sr += ue[qx * nq + qy] * ue[qx * nq + qy] * wq;
      }
  }
  se = block_sum(se);
  sr = block_sum(sr);
  if (threadIdx.x == 0) {
    partial_err[blockIdx.x] = se;
    partial_ref[blockIdx.x] = sr;
  }
}

__global__ void l2_final_kernel(const double* __restrict__ pe, const double* __restrict__ pr, int nblocks,
                                double* __restrict__ out) {
  if (threadIdx.x != 0 || blockIdx.x != 0) return;
  double a = 0.0, b = 0.0;
  for (int k = 0; k < nblocks; ++k) {
    a += pe[k];
    b += pr[k];
  }
  out[0] = sqrt(a);
  out[1] = sqrt(b);
}

}  // namespace stefan

using namespace stefan;

// One implementation for the single-GPU solve and the column-slab solve (one process per GPU):
// `exchange` fills the caller's halo buffers with the neighbours' edge columns of a slab vector
// before every operator application, `allreduce` sums the two dot products of a reduction over
// the ranks (kry::Reducer's hook).  Both null: the single-GPU solver.
static int cg_solve_impl(stefan_ipdg* hv, const double* b, double* x, const stefan_cg_params* prm,
                         stefan_cg_result* res, const double* halo_lo, const double* halo_hi,
                         stefan_halo_fn exchange, stefan_allreduce_fn allreduce, void* ctx, void* stream) {
  IpdgHandle* h = reinterpret_cast<IpdgHandle*>(hv);
  const bool slab = exchange != nullptr || allreduce != nullptr;
  // The iteration runs on a stream of its own (ordered after the caller's stream): a batch of
  // iterations is captured into a CUDA graph and replayed, which the legacy default stream
  // cannot do.  The call is synchronous, so the caller's stream sees the result on return.
  cudaStream_t caller = static_cast<cudaStream_t>(stream), st = nullptr;
  cudaEvent_t ordered = nullptr;
  cudaGraph_t graph = nullptr;
  cudaGraphExec_t gexec = nullptr;
  double *work = nullptr, *Binv = nullptr;
  int rc = ST_OK;
  // defined before the first fallible call: every early return releases what exists so far
  auto cleanup = [&]() {
    if (st) cudaStreamSynchronize(st);
    if (gexec) cudaGraphExecDestroy(gexec);
    if (graph) cudaGraphDestroy(graph);
    if (work) cudaFree(work);
    if (Binv) cudaFree(Binv);
    if (ordered) cudaEventDestroy(ordered);
    if (st) cudaStreamDestroy(st);
  };
#define CG_CUDA(call)                                                                        \
  do {                                                                                       \
    cudaError_t e__ = (call);                                                                \
    if (e__ != cudaSuccess) {                                                                \
      cleanup();                                                                             \
      return fail(ST_CUDA_ERROR, "%s:%d: %s -> %s", __FILE__, __LINE__, #call, cudaGetErrorString(e__)); \
    }                                                                                        \
  } while (0)
#define CG_RC(expr)          \
  do {                       \
    rc = (expr);             \
    if (rc) {                \
      cleanup();             \
      return rc;             \
    }                        \
  } while (0)
  CG_CUDA(cudaStreamCreateWithFlags(&st, cudaStreamNonBlocking));
  CG_CUDA(cudaEventCreateWithFlags(&ordered, cudaEventDisableTiming));
  CG_CUDA(cudaEventRecord(ordered, caller));
  CG_CUDA(cudaStreamWaitEvent(st, ordered, 0));
  stream = st;
  const int64_t n = h->ndofs;
  const int64_t ncells = (int64_t)h->nx * h->ny;
  const int nblk = (int)std::min<int64_t>((ncells + kRedThreads - 1) / kRedThreads, (int64_t)sm_count() * 8);
  const int64_t ns = (n + 3) / 4 * 4;  // 32-byte aligned vectors (the apply kernel needs 16)
  const size_t wbytes = (size_t)(4 * ns + 2 * (size_t)nblk + kry::S_COUNT) * sizeof(double);
  CG_CUDA(cudaMalloc(&work, wbytes));
  double *r = work, *z = r + ns, *p = z + ns, *Ap = p + ns, *part = Ap + ns, *scal = part + 2 * (size_t)nblk;
  CG_CUDA(cudaMemsetAsync(scal, 0, kry::S_COUNT * sizeof(double), st));
  const kry::Reducer red{part, scal, nblk, st, allreduce, ctx};
  if (prm->preconditioner == 1) {
    std::vector<double> inv;
    switch (h->order) {
      case 1: build_block_inverses<1>(*static_cast<IpTab<1>*>(h->tab_host), &inv); break;
      case 2: build_block_inverses<2>(*static_cast<IpTab<2>*>(h->tab_host), &inv); break;
      case 3: build_block_inverses<3>(*static_cast<IpTab<3>*>(h->tab_host), &inv); break;
    }
    CG_CUDA(cudaMalloc(&Binv, inv.size() * sizeof(double)));
    CG_CUDA(cudaMemcpyAsync(Binv, inv.data(), inv.size() * sizeof(double), cudaMemcpyHostToDevice, st));
    CG_CUDA(cudaStreamSynchronize(st));  // `inv` goes out of scope
  }
  auto update = [&](int first) {
    switch (h->order) {
      case 1: cg_update_kernel<1><<<nblk, kRedThreads, 0, st>>>(x, r, z, p, Ap, scal, first, Binv, h->phase_dev, h->nx, h->ny, part); break;
      case 2: cg_update_kernel<2><<<nblk, kRedThreads, 0, st>>>(x, r, z, p, Ap, scal, first, Binv, h->phase_dev, h->nx, h->ny, part); break;
      case 3: cg_update_kernel<3><<<nblk, kRedThreads, 0, st>>>(x, r, z, p, Ap, scal, first, Binv, h->phase_dev, h->nx, h->ny, part); break;
    }
  };
  // Ap = A v on this slab (halo columns of v fetched first)
  auto apply = [&](const double* v) -> int {
    if (exchange != nullptr) {
      int e = exchange(ctx, v, stream);
      if (e) return fail(ST_BAD_ARGUMENT, "stefan_ipdg_cg_solve_slab: the halo exchange callback returned %d", e);
    }
    return stefan_ipdg_apply(hv, v, Ap, halo_lo, halo_hi, 0, stream);
  };
  auto hook_failed = [&](int e) {
    return fail(ST_BAD_ARGUMENT, "stefan_ipdg_cg_solve_slab: the all-reduce callback returned %d", e);
  };
  // r = b - A x0
  CG_RC(apply(x));
  kry::residual_kernel<<<nblk, kRedThreads, 0, st>>>(r, b, Ap, n);
  // ||b||
  if (int e = red.dot(b, b, nullptr, nullptr, n, kry::P_BB)) { cleanup(); return hook_failed(e); }
  update(1);  // z = Binv r, r.z, r.r
  if (int e = red.finish(kry::P_CG_START)) { cleanup(); return hook_failed(e); }
  kry::cg_p_kernel<<<nblk, kRedThreads, 0, st>>>(p, z, scal, 1, n);
  double hs[kry::S_COUNT];
  CG_CUDA(cudaMemcpyAsync(hs, scal, sizeof(hs), cudaMemcpyDeviceToHost, st));
  CG_CUDA(cudaStreamSynchronize(st));
  const double bnorm = std::sqrt(hs[kry::S_BB]);
  double rnorm = std::sqrt(hs[kry::S_RR]);
  const double target = prm->rel_tolerance * (bnorm > 0 ? bnorm : 1.0);
  const int check = prm->check_every > 0 ? prm->check_every : 10;
  int it = 0;
  auto iterations = [&](int count) -> int {
    for (int k = 0; k < count; ++k) {
      int rc2 = apply(p);
      if (rc2) return rc2;
      if (int e = red.dot(p, Ap, nullptr, nullptr, n, kry::P_CG_ALPHA)) return hook_failed(e);
      update(0);
      if (int e = red.finish(kry::P_CG_BETA)) return hook_failed(e);
      kry::cg_p_kernel<<<nblk, kRedThreads, 0, st>>>(p, z, scal, 0, n);
    }
    return ST_OK;
  };
  // one batch of `check` iterations (8 launches each) as a CUDA graph: one launch per batch;
  // what makes the small systems of the reference's convergence studies launch-bound.  The slab
  // form calls back into the host every iteration and is launched plainly.
  const bool use_graph = !slab && prm->max_iterations >= check && getenv("STEFAN_CG_NO_GRAPH") == nullptr;
  if (use_graph && rnorm > target) {
    CG_RC(stefan_ipdg_apply(hv, p, Ap, nullptr, nullptr, 0, stream));  // warm-up outside the capture (attributes)
    CG_CUDA(cudaStreamBeginCapture(st, cudaStreamCaptureModeThreadLocal));
    rc = iterations(check);
    cudaError_t ce = cudaStreamEndCapture(st, &graph);
    if (rc) { cleanup(); return rc; }
    CG_CUDA(ce);
    CG_CUDA(cudaGraphInstantiate(&gexec, graph, 0));
  }
  while (it < prm->max_iterations && rnorm > target) {
    const int batch = std::min(check, prm->max_iterations - it);
    if (gexec != nullptr && batch == check) {
      CG_CUDA(cudaGraphLaunch(gexec, st));
    } else {
      CG_RC(iterations(batch));
    }
    it += batch;
    CG_CUDA(cudaMemcpyAsync(hs, scal, sizeof(hs), cudaMemcpyDeviceToHost, st));
    CG_CUDA(cudaStreamSynchronize(st));
    rnorm = std::sqrt(hs[kry::S_RR]);  // (all-reduced: every rank takes the same decision)
    if (!(rnorm == rnorm)) break;  // NaN: breakdown
  }
  CG_CUDA(cudaGetLastError());
  res->iterations = it;
  res->residual_norm = rnorm;
  res->rhs_norm = bnorm;
  res->converged = rnorm <= target ? 1 : 0;
  cleanup();
#undef CG_CUDA
#undef CG_RC
  return ST_OK;
}

extern "C" {

int stefan_ipdg_cg_solve(stefan_ipdg* hv, const double* b, double* x, const stefan_cg_params* prm,
                         stefan_cg_result* res, void* stream) {
  IpdgHandle* h = reinterpret_cast<IpdgHandle*>(hv);
  STEFAN_REQUIRE(h && b && x && prm && res, "stefan_ipdg_cg_solve: null argument");
  STEFAN_REQUIRE(!h->has_lo && !h->has_hi, "stefan_ipdg_cg_solve: slab handles need stefan_ipdg_cg_solve_slab");
  STEFAN_REQUIRE(prm->max_iterations >= 0 && prm->rel_tolerance >= 0.0, "stefan_ipdg_cg_solve: bad parameters");
  STEFAN_REQUIRE(prm->preconditioner == 0 || prm->preconditioner == 1, "stefan_ipdg_cg_solve: unknown preconditioner");
  return cg_solve_impl(hv, b, x, prm, res, nullptr, nullptr, nullptr, nullptr, nullptr, stream);
}

int stefan_ipdg_cg_solve_slab(stefan_ipdg* hv, const double* b, double* x, const stefan_cg_params* prm,
                              stefan_cg_result* res, const double* halo_lo, const double* halo_hi,
                              stefan_halo_fn exchange, stefan_allreduce_fn allreduce, void* ctx, void* stream) {
  IpdgHandle* h = reinterpret_cast<IpdgHandle*>(hv);
  STEFAN_REQUIRE(h && b && x && prm && res, "stefan_ipdg_cg_solve_slab: null argument");
  STEFAN_REQUIRE((!h->has_lo || halo_lo) && (!h->has_hi || halo_hi),
                 "stefan_ipdg_cg_solve_slab: the handle has a neighbour slab whose halo buffer is null");
  STEFAN_REQUIRE((!h->has_lo && !h->has_hi) || exchange, "stefan_ipdg_cg_solve_slab: a slab handle needs the halo exchange callback");
  STEFAN_REQUIRE(prm->max_iterations >= 0 && prm->rel_tolerance >= 0.0, "stefan_ipdg_cg_solve_slab: bad parameters");
  STEFAN_REQUIRE(prm->preconditioner == 0 || prm->preconditioner == 1, "stefan_ipdg_cg_solve_slab: unknown preconditioner");
  return cg_solve_impl(hv, b, x, prm, res, h->has_lo ? halo_lo : nullptr, h->has_hi ? halo_hi : nullptr, exchange,
                       allreduce, ctx, stream);
}

int stefan_ipdg_l2_error(stefan_ipdg* hv, const double* u, const double* uex_qp, double* out2_dev, void* stream) {
  IpdgHandle* h = reinterpret_cast<IpdgHandle*>(hv);
  STEFAN_REQUIRE(h && u && uex_qp && out2_dev, "stefan_ipdg_l2_error: null argument");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  L2Tab t;
  std::memset(&t, 0, sizeof(t));
  t.n1 = h->order + 1;
  t.nq = h->numqp;
  for (int q = 0; q < t.nq; ++q) {
    t.w[q] = h->ref.qw[q];
    for (int a = 0; a < t.n1; ++a) t.Nq[q][a] = h->ref.Nq[q][a];
  }
  t.detjac = 0.25 * h->hx * h->hy;
  const int64_t ncells = (int64_t)h->nx * h->ny;
  const int nblk = (int)std::min<int64_t>((ncells + kRedThreads - 1) / kRedThreads, (int64_t)sm_count() * 8);
  double* part = nullptr;
  STEFAN_CUDA(cudaMallocAsync(&part, 2 * (size_t)nblk * sizeof(double), st));
  ipdg_l2_error_kernel<<<nblk, kRedThreads, 0, st>>>(t, u, uex_qp, ncells, part, part + nblk);
  l2_final_kernel<<<1, 1, 0, st>>>(part, part + nblk, nblk, out2_dev);
  STEFAN_CUDA(cudaFreeAsync(part, st));
  STEFAN_CUDA(cudaGetLastError());
  return ST_OK;
}

}  // extern "C"
