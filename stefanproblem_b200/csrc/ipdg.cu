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
//   phase: one byte per cell (1: sign +1, 2: sign -1), padded to (nx+2) x (ny+2)
//          with 0 = "no cell".
//
// Kernels
//   ipdg_apply_simple<P>: one thread per cell, neighbours straight from global
//       memory (L1/L2).  Plain; the second opinion for the march kernels.
//   ipdg_apply_march<P> (ipdg_march.cuh): the production kernels.  One warp owns a
//       strip of 30 cells in y (+1 halo cell each side = 32 lanes) and marches along
//       x; column chunks arrive by TMA bulk copies into a per-warp ring; neighbours
//       enter as face traces / moments (registers, warp shuffles).  Every dof is read
//       from HBM once (plus 2/30 halo re-reads that hit L2) and every result written
//       once: 16 B per dof-update.  The loop computes the cells whose four neighbours
//       exist and have the cell's own phase ("interior-uniform" cells); all others
//       (boundary cells, cells next to the other phase) are listed at handle creation
//       and recomputed with the general tables by each CTA's tail (fixup_tail).
#include "../../include/stefan_b200.h"
#include "common.cuh"
#include "ipdg_handle.h"
#include "peer_dev.cuh"

#include <algorithm>
#include <cstdlib>
#include <utility>

namespace stefan {



struct ApplyArgs {
  const double* x;
  const double* x_lo;  // neighbour slab's last column (ny*NF) or nullptr
  const double* x_hi;  // neighbour slab's first column or nullptr
  double* y;
  const int8_t* phase;  // padded
  int nx, ny;
  int cbeg, cend;  // columns to produce (a sub-range when the host path pipelines slabs)
  const int2* fix;  // {cell, face-type code} of the cells the march kernels leave to their tail (see fixup_tail)
  int nfix;
  unsigned zero;  // 0 at run time, unknown at compile time (see ipdg_march.cuh, slot release)
  // peer-memory halos with the ordering fused into the kernel (stefan_ipdg_apply_peer);
  // all null / 0 otherwise
  unsigned long long* sync_my_ready;         // published at kernel start: "my x of this epoch is final"
  unsigned long long* sync_my_done;          // published by the last CTA: "I am done reading my neighbours"
  const unsigned long long* sync_lo_ready;   // neighbours' ready flags (peer memory)
  const unsigned long long* sync_hi_ready;
  unsigned long long sync_epoch, sync_timeout_ns;  // epoch != 0: peer mode
  // graph-replayable form: the epoch lives on the device -- the kernel uses *sync_epoch_dev + 1 and
  // its last CTA stores that back -- so the same captured launch advances the protocol on every replay
  unsigned long long* sync_epoch_dev;
  // push-style mailboxes in the NEIGHBOURS' memory (null = none): the kernel also stores its ready value (at
  // start) / done value (by its last CTA) there, so that the neighbours poll their own memory instead of
  // reading a flag over NVLink (a remote poll is a 1-2 us round trip)
  unsigned long long* sync_post_ready[2];
  unsigned long long* sync_post_done[2];
  int sync_chained;  // 1: the halo gate is the neighbours' DONE value of the previous epoch (see stefan_peer_sync)
  int sync_dev;      // development only (STEFAN_DEV_PEER): 1 = never wait for a flag, 2 = never post one (timing experiments)
  int* sync_status;                          // set to 1 if a wait timed out
  unsigned long long* dbg_times;             // development only: per-CTA {start, prologue, loop, end} globaltimer, or null
  // dynamic tail queue: [0] next fix-list index, [1] CTAs finished; both reset by the last CTA
  unsigned* tail_ctr;
  // fused explicit step (MODE 1 of the kernels): y = x + step_scale * Minv (b - A x), Minv the
  // Kronecker square of the inverse 1-D reference mass matrix
  const double* b;
  double step_scale;
  double minv[kMaxN1][kMaxN1];
};

// Programmatic dependent launch: the kernels may start while the previous kernel of the stream
// drains (their CTAs take SM slots as they free up and run their prologue); griddep_wait()
// returns once the previous kernel has completed and its memory is visible, and must precede
// every access to memory a previous kernel may have written.  Both are no-ops in a launch
// without the attribute.
__device__ __forceinline__ void griddep_launch() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
__device__ __forceinline__ void griddep_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }

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
  const int64_t cell_end = (int64_t)a.cend * a.ny;
  for (int64_t cell = (int64_t)a.cbeg * a.ny + blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
       cell < cell_end; cell += (int64_t)gridDim.x * blockDim.x) {
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
    ip_cell_rows<P>(tab, s == 1 ? 0 : 1, ip_face_type(s, sL),
                    ip_face_type(s, sR), ip_face_type(s, sD),
                    ip_face_type(s, sU), X, XL, XR, XD, XU, out);
    double* yc = a.y + cell * NF;
#pragma unroll
    for (int k = 0; k < NF; ++k) yc[k] = out[k];
  }
}

// The step epilogue of one cell: out (= A x rows) becomes x + scale * Minv (b - out).
template <int P>
__device__ __forceinline__ void step_epilogue(const ApplyArgs& a, const double* __restrict__ X,
                                              const double* __restrict__ B, double* __restrict__ out) {
  constexpr int N1 = P + 1, NF = N1 * N1;
  double t[NF];
#pragma unroll
  for (int ix = 0; ix < N1; ++ix)
#pragma unroll
    for (int iy = 0; iy < N1; ++iy) {
      double v = 0.0;
#pragma unroll
      for (int c = 0; c < N1; ++c) v += a.minv[iy][c] * (B[ix * N1 + c] - out[ix * N1 + c]);
      t[ix * N1 + iy] = v;
    }
#pragma unroll
  for (int ix = 0; ix < N1; ++ix)
#pragma unroll
    for (int iy = 0; iy < N1; ++iy) {
      double v = 0.0;
#pragma unroll
      for (int c = 0; c < N1; ++c) v += a.minv[ix][c] * t[c * N1 + iy];
      out[ix * N1 + iy] = X[ix * N1 + iy] + a.step_scale * v;
    }
}

// Cells the marching loops do not produce (they skip the store): boundary cells and cells
// next to the other phase.  They are disjoint from what the loops write.  Round 2: the list is
// a QUEUE -- every warp that is done with its strip takes a batch of cells through an atomic
// counter until the list is empty, so the warps whose tiles finish first absorb the tail
// (round 1 gave CTA b the cells b*blockDim .. : the first CTAs of the grid, which also own
// the boundary column, paid for all of it) -- and the cells are processed COOPERATIVELY: one
// lane per NODE, NF lanes per cell (8 / 3 / 2 cells per warp at orders 1 / 2 / 3).  The line
// operators of ipdg_tables.h become warp shuffles along the x- / y-lines of the cell, every
// load and store is coalesced, and a cell costs ~150 instructions at ~40 registers instead of a
// whole per-thread ip_cell_rows with run-time table indices (10-25 us per 32-cell chunk).
template <int P, int MODE>
__device__ __forceinline__ void fixup_tail(const IpTab<P>& tab, const ApplyArgs& a) {
  constexpr int N1 = P + 1, NF = N1 * N1;
  constexpr int CW = 32 / NF;      // cells per warp and round
  constexpr int BATCH = 4;         // rounds per queue access; their loads are issued together
  if (a.nfix <= 0) return;
  const int lane = threadIdx.x & 31;
  const bool lane_live = lane < CW * NF;
  const int grp = lane_live ? lane / NF : 0;          // (idle lanes shadow group 0 and never store)
  const int node = lane_live ? lane - grp * NF : 0;
  const int ix = node / N1, iy = node % N1;
  const int gbase = grp * NF;
  unsigned long long epoch = 0;   // the value the neighbours' flags must have reached
  if (a.sync_epoch != 0) epoch = peer_epoch(a.sync_epoch_dev, a.sync_epoch) - (unsigned long long)a.sync_chained;
  for (;;) {
    unsigned base = 0;
    if (lane == 0) base = atomicAdd(a.tail_ctr, (unsigned)(CW * BATCH));
    base = __shfl_sync(0xffffffffu, base, 0);
    if (base >= (unsigned)a.nfix) break;
    // ---- all loads of the batch first (two dependent levels: record, then dofs) ----
    int2 rec[BATCH];
    double X[BATCH], XL[BATCH], XR[BATCH], XD[BATCH], XU[BATCH];
    [[maybe_unused]] double Bv[BATCH];
#pragma unroll
    for (int rd = 0; rd < BATCH; ++rd) rec[rd] = a.fix[min((int)base + rd * CW + grp, a.nfix - 1)];
#pragma unroll
    for (int rd = 0; rd < BATCH; ++rd) {
      const int64_t cell = rec[rd].x;
      const int code = rec[rd].y;
      const int i = (int)(cell / a.ny), j = (int)(cell - (int64_t)i * a.ny);
      const int tL = (code >> 1) & 3, tR = (code >> 3) & 3, tD = (code >> 5) & 3, tU = (code >> 7) & 3;
      const double* xc = a.x + cell * NF;
      const double* pl = (i > 0) ? xc - (size_t)a.ny * NF : a.x_lo + (size_t)j * NF;
      const double* pr = (i < a.nx - 1) ? xc + (size_t)a.ny * NF : a.x_hi + (size_t)j * NF;
      if (a.sync_epoch != 0) {  // peer halos: the neighbour's column must be published
        if (i == 0 && tL != T_BDRY && a.sync_lo_ready && !(a.sync_dev & 1)) peer_flag_wait(a.sync_lo_ready, epoch, a.sync_timeout_ns, a.sync_status);
        if (i == a.nx - 1 && tR != T_BDRY && a.sync_hi_ready && !(a.sync_dev & 1)) peer_flag_wait(a.sync_hi_ready, epoch, a.sync_timeout_ns, a.sync_status);
      }
      // this lane's node of the cell and of its four neighbours (zeros where there is none)
      X[rd] = xc[node];
      XL[rd] = tL != T_BDRY ? pl[node] : 0.0;
      XR[rd] = tR != T_BDRY ? pr[node] : 0.0;
      XD[rd] = tD != T_BDRY ? xc[node - NF] : 0.0;
      XU[rd] = tU != T_BDRY ? xc[node + NF] : 0.0;
      if constexpr (MODE == 1) Bv[rd] = a.b ? a.b[cell * NF + node] : 0.0;
    }
    // ---- the line operators of each cell, by shuffles along its x- / y-lines ----
#pragma unroll
    for (int rd = 0; rd < BATCH; ++rd) {
      const int t0 = (int)base + rd * CW;
      if (t0 < a.nfix) {  // (warp-uniform)
        const bool live = lane_live && (t0 + grp) < a.nfix;
        const int code = rec[rd].y;
        const int S = code & 1, tL = (code >> 1) & 3, tR = (code >> 3) & 3, tD = (code >> 5) & 3, tU = (code >> 7) & 3;
        const double* Dx = tab.D[0][S][tL][tR];
        const double* Dy = tab.D[1][S][tD][tU];
        const double* L0x = tab.L0[0][S][tL];
        const double* Rpx = tab.Rp[0][S][tR];
        const double* L0y = tab.L0[1][S][tD];
        const double* Rpy = tab.Rp[1][S][tU];
        double ax = 0.0, sl = 0.0, sr = 0.0, ay = 0.0, sd = 0.0, su = 0.0;
#pragma unroll
        for (int b = 0; b < N1; ++b) {
          const int lx = gbase + b * N1 + iy, ly = gbase + ix * N1 + b;
          ax += Dx[ix * N1 + b] * __shfl_sync(0xffffffffu, X[rd], lx);
          sl += L0x[b] * __shfl_sync(0xffffffffu, XL[rd], lx);
          sr += Rpx[b] * __shfl_sync(0xffffffffu, XR[rd], lx);
          ay += Dy[iy * N1 + b] * __shfl_sync(0xffffffffu, X[rd], ly);
          sd += L0y[b] * __shfl_sync(0xffffffffu, XD[rd], ly);
          su += Rpy[b] * __shfl_sync(0xffffffffu, XU[rd], ly);
        }
        const double xlP = __shfl_sync(0xffffffffu, XL[rd], gbase + P * N1 + iy);
        const double xr0 = __shfl_sync(0xffffffffu, XR[rd], gbase + iy);
        const double xdP = __shfl_sync(0xffffffffu, XD[rd], gbase + ix * N1 + P);
        const double xu0 = __shfl_sync(0xffffffffu, XU[rd], gbase + ix * N1);
        ax += (ix == 0) ? sl : tab.Lc[0][S][tL][ix] * xlP;
        ax += (ix == P) ? sr : tab.Rc[0][S][tR][ix] * xr0;
        ay += (iy == 0) ? sd : tab.Lc[1][S][tD][iy] * xdP;
        ay += (iy == P) ? su : tab.Rc[1][S][tU][iy] * xu0;
        double out = 0.0;
#pragma unroll
        for (int c = 0; c < N1; ++c) {
          out += tab.My[iy * N1 + c] * __shfl_sync(0xffffffffu, ax, gbase + ix * N1 + c);
          out += tab.Mx[ix * N1 + c] * __shfl_sync(0xffffffffu, ay, gbase + c * N1 + iy);
        }
        if constexpr (MODE == 1) {
          // step epilogue: out <- X + scale (Minv (x) Minv) (b - out), again along the lines
          const double r = Bv[rd] - out;
          double tt = 0.0, z = 0.0;
#pragma unroll
          for (int c = 0; c < N1; ++c) tt += a.minv[iy][c] * __shfl_sync(0xffffffffu, r, gbase + ix * N1 + c);
#pragma unroll
          for (int c = 0; c < N1; ++c) z += a.minv[ix][c] * __shfl_sync(0xffffffffu, tt, gbase + c * N1 + iy);
          out = X[rd] + a.step_scale * z;
        }
        if (live) a.y[(int64_t)rec[rd].x * NF + node] = out;
      }
    }
  }
}

// End of a march kernel: the last CTA to get here resets the tail queue for the next launch
// and, with peer halos, tells the neighbours that their columns are free.
__device__ __forceinline__ void march_finish(const ApplyArgs& a) {
  __syncthreads();
  if (threadIdx.x == 0) {
    __threadfence();
    if (atomicAdd(a.tail_ctr + 1, 1u) == gridDim.x - 1) {
      a.tail_ctr[0] = 0;
      a.tail_ctr[1] = 0;
      __threadfence();
      if (a.sync_epoch != 0) {
        const unsigned long long epoch = peer_epoch(a.sync_epoch_dev, a.sync_epoch);
        if (!(a.sync_dev & 2)) peer_flag_publish(a.sync_my_done, epoch, a.sync_post_done[0], a.sync_post_done[1]);
        if (a.sync_epoch_dev != nullptr) *a.sync_epoch_dev = epoch;  // every CTA has read it: all are done
      }
    }
  }
}

}  // namespace stefan
#include "ipdg_march.cuh"
namespace stefan {

// --------------------------------------------------------------------------
// host side
// --------------------------------------------------------------------------
// cuTensorMapEncodeTiled through the runtime (no link against libcuda)
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*,
                                  CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion,
                                  CUtensorMapFloatOOBfill);
static EncodeTiledFn encode_tiled_fn() {
  static EncodeTiledFn fn = nullptr;
  if (fn == nullptr) {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
        q == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<EncodeTiledFn>(p);
  }
  return fn;
}

// x as a 2-D tensor {16 doubles, ncells}, tiles of 32 cells, 128-byte swizzle
static int make_cell_tensor_map(const double* x, int64_t ncells, CUtensorMap* out) {
  EncodeTiledFn enc = encode_tiled_fn();
  if (enc == nullptr) return fail(ST_CUDA_ERROR, "cuTensorMapEncodeTiled is not available from this driver");
  const cuuint64_t gdim[2] = {16, (cuuint64_t)ncells};
  const cuuint64_t gstr[1] = {128};
  const cuuint32_t box[2] = {16, 32};
  const cuuint32_t estr[2] = {1, 1};
  CUresult r = enc(out, CU_TENSOR_MAP_DATA_TYPE_FLOAT64, 2, const_cast<double*>(x), gdim, gstr, box, estr,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) return fail(ST_CUDA_ERROR, "cuTensorMapEncodeTiled failed (%d)", (int)r);
  return ST_OK;
}

static int march_warps(int order) {
  return order == 1 ? MarchCfg<1>::WARPS : order == 2 ? MarchCfg<2>::WARPS : MarchCfg<3>::WARPS;
}

// Launch with the programmatic-stream-serialization attribute (see griddep_wait above);
// STEFAN_PDL=0 turns it off (A/B timing).
template <typename... KArgs, typename... Args>
static cudaError_t launch_pdl(void (*kernel)(KArgs...), int grid, int block, size_t smem, cudaStream_t st,
                              Args&&... args) {
  static const bool pdl = [] { const char* e = getenv("STEFAN_PDL"); return e == nullptr || atoi(e) != 0; }();
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3((unsigned)grid);
  cfg.blockDim = dim3((unsigned)block);
  cfg.dynamicSmemBytes = smem;
  cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = pdl ? 1 : 0;
  return cudaLaunchKernelEx(&cfg, kernel, KArgs(std::forward<Args>(args))...);
}

template <int P, int MODE>
static int launch_march(IpdgHandle* h, const IpTab<P>& tab, const ApplyArgs& args0, cudaStream_t st) {
  using C = MarchCfg<P>;
  const SegCache& sc = h->plan.segments(args0.cbeg, args0.cend, sm_count() * march_minb<P, MODE>());
  const SegTab& segs = sc.tab;
  const int grid = sc.n;
  if (grid <= 0) return fail(ST_UNSUPPORTED, "stefan_ipdg_apply: more than %d strip groups (ny too large)", kMaxMarchCtas);
  constexpr int smem = MarchSmem<P, MODE>::TOTAL;
  static PerDeviceFlag attr_set;
  if (!attr_set.here()) {
    STEFAN_CUDA(cudaFuncSetAttribute(ipdg_apply_march<P, MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    attr_set.here() = true;
  }
  ApplyArgs args = args0;
  if (h->tail_ctr_dev == nullptr) {
    STEFAN_CUDA(cudaMalloc(&h->tail_ctr_dev, kTailRing * 2 * sizeof(unsigned)));
    STEFAN_CUDA(cudaMemset(h->tail_ctr_dev, 0, kTailRing * 2 * sizeof(unsigned)));
  }
  args.tail_ctr = h->tail_ctr_dev + 2 * (h->launch_seq++ % kTailRing);
  CUtensorMap tm;
  std::memset(&tm, 0, sizeof(tm));
  if (C::TENSOR) {
    int rc = make_cell_tensor_map(args.x, (int64_t)h->nx * h->ny, &tm);
    if (rc) return rc;
  }
  // development only: STEFAN_DEV_CTA_TIMES=<file> dumps per-CTA {group, c0, c1, 4 x globaltimer ns}
  static const char* dbg_path = getenv("STEFAN_DEV_CTA_TIMES");
  if (dbg_path != nullptr) {
    unsigned long long* d = nullptr;
    STEFAN_CUDA(cudaMalloc(&d, (size_t)grid * 4 * sizeof(unsigned long long)));
    STEFAN_CUDA(cudaMemset(d, 0, (size_t)grid * 4 * sizeof(unsigned long long)));
    args.dbg_times = d;
    ipdg_apply_march<P, MODE><<<grid, C::WARPS * 32, smem, st>>>(tab, args, h->plan.runs_dev, h->plan.run_ofs_dev, tm, segs);
    STEFAN_CUDA(cudaStreamSynchronize(st));
    std::vector<unsigned long long> t((size_t)grid * 4);
    STEFAN_CUDA(cudaMemcpy(t.data(), d, t.size() * sizeof(unsigned long long), cudaMemcpyDeviceToHost));
    cudaFree(d);
    if (FILE* f = fopen(dbg_path, "w")) {
      for (int b = 0; b < grid; ++b)
        fprintf(f, "%d %d %d %llu %llu %llu %llu\n", segs.e[b].x, segs.e[b].y, segs.e[b].z, t[4 * b], t[4 * b + 1],
                t[4 * b + 2], t[4 * b + 3]);
      fclose(f);
    }
    return ST_OK;
  }
  STEFAN_CUDA(launch_pdl(ipdg_apply_march<P, MODE>, grid, C::WARPS * 32, smem, st, tab, args,
                         (const int4*)h->plan.runs_dev, (const int*)h->plan.run_ofs_dev, tm, segs));
  return ST_OK;
}

template <int P>
static int launch_apply(IpdgHandle* h, const ApplyArgs& args0, int algo, cudaStream_t st) {
  const IpTab<P>& tab = *static_cast<const IpTab<P>*>(h->tab_host);
  ApplyArgs args = args0;
  const int ncol = args.cend - args.cbeg;
  if (ncol <= 0) return ST_OK;
  const bool step = args.step_scale != 0.0;
  if (algo == 1) {
    if (step) return fail(ST_UNSUPPORTED, "stefan_ipdg_step: the plain kernel has no fused step");
    const int64_t ncells = (int64_t)ncol * h->ny;
    int blocks = (int)std::min<int64_t>((ncells + 127) / 128, (int64_t)sm_count() * 32);
    ipdg_apply_simple<P><<<blocks, 128, 0, st>>>(tab, args);
    STEFAN_CUDA(cudaGetLastError());
    return ST_OK;
  }
  // exception cells of the produced columns (the list is sorted by column)
  args.fix = h->plan.fix_dev + h->plan.fix_col_ofs[args.cbeg];
  args.nfix = h->plan.fix_col_ofs[args.cend] - h->plan.fix_col_ofs[args.cbeg];
  static const bool skip_tail = getenv("STEFAN_DEV_SKIP_TAIL") != nullptr;  // timing experiments only (wrong results)
  if (skip_tail) args.nfix = 0;
  // CTA = (group of adjacent strips, x segment); one wave of MINB CTAs per SM
  int rc = step ? launch_march<P, 1>(h, tab, args, st) : launch_march<P, 0>(h, tab, args, st);
  if (rc) return rc;
  STEFAN_CUDA(cudaGetLastError());
  return ST_OK;
}

static int dispatch_apply(IpdgHandle* h, const ApplyArgs& a, int algo, cudaStream_t st) {
  switch (h->order) {
    case 1: return launch_apply<1>(h, a, algo, st);
    case 2: return launch_apply<2>(h, a, algo, st);
    case 3: return launch_apply<3>(h, a, algo, st);
  }
  return fail(ST_UNSUPPORTED, "order %d", h->order);
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
  h->xh_dev = h->yh_dev = nullptr;
  h->tail_ctr_dev = nullptr;
  h->launch_seq = 0;
  h->iface_dev = nullptr;
  h->niface = 0;
  h->iface_partial_dev = nullptr;
  h->iface_ctr_dev = nullptr;
  h->host_ready = false;
  h->tab_host = nullptr;
  const int NF = (p->order + 1) * (p->order + 1);
  h->ndofs = (int64_t)p->nx * p->ny * NF;
  build_ref1d(p->order, p->numqp, &h->ref);
  invert_mass_1d(h->ref, p->order + 1, h->minv1d);
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
      ph[(size_t)(i + 1) * (p->ny + 2) + (j + 1)] = (s == 1) ? 1 : 2;
    }
  for (int j = 0; j < p->ny; ++j) {
    if (phase_lo) ph[(size_t)0 * (p->ny + 2) + (j + 1)] = (phase_lo[j] == 1) ? 1 : 2;
    if (phase_hi) ph[(size_t)(p->nx + 1) * (p->ny + 2) + (j + 1)] = (phase_hi[j] == 1) ? 1 : 2;
  }
  if ((int64_t)p->nx * p->ny >= (int64_t)INT32_MAX) {
    delete h;
    return fail(ST_UNSUPPORTED, "stefan_ipdg_create: more than 2^31 cells per slab");
  }
  cudaError_t e = cudaMalloc(&h->phase_dev, pn);
  if (e == cudaSuccess) e = cudaMemcpy(h->phase_dev, ph.data(), pn, cudaMemcpyHostToDevice);
  if (e == cudaSuccess) e = h->plan.build(ph.data(), p->nx, p->ny, march_warps(p->order));
  {
    // the interface faces, owned by their +1 cell (every face is summed exactly once across slabs)
    std::vector<int2> faces;
    const int ld = p->ny + 2;
    const int step[4] = {-1, ld, 1, -ld};  // bottom, right, top, left in the padded array
    for (int i = 0; i < p->nx; ++i)
      for (int j = 0; j < p->ny; ++j) {
        const int8_t* c = ph.data() + (size_t)(i + 1) * ld + (j + 1);
        if (*c != 1) continue;
        for (int f = 0; f < 4; ++f)
          if (c[step[f]] == 2) faces.push_back(make_int2(i * p->ny + j, f));
      }
    h->niface = (int)faces.size();
    if (e == cudaSuccess && h->niface > 0) e = cudaMalloc(&h->iface_dev, faces.size() * sizeof(int2));
    if (e == cudaSuccess && h->niface > 0)
      e = cudaMemcpy(h->iface_dev, faces.data(), faces.size() * sizeof(int2), cudaMemcpyHostToDevice);
  }
  // measured (sixteenths of a column, see MarchPlan::segments; profiles/r2_sweep_mixed_cost.log): a quarter column
  // at order 1, half a column at orders 2 and 3 (order 2: a whole column is 2 % better on the 416-column slab
  // and 8 % worse on the 3334-column mesh)
  h->plan.mixed_cost_q = p->order == 1 ? 4 : 8;
  // tiles that touch a neighbouring slab get two columns less: they may have to wait for its flag
  h->plan.edge_lo_q = h->has_lo ? 32 : 0;
  h->plan.edge_hi_q = h->has_hi ? 32 : 0;
  if (e != cudaSuccess) { h->plan.destroy(); if (h->phase_dev) cudaFree(h->phase_dev); delete h; return fail(ST_CUDA_ERROR, "stefan_ipdg_create: %s", cudaGetErrorString(e)); }
  *out = reinterpret_cast<stefan_ipdg*>(h);
  return ST_OK;
}

int stefan_ipdg_destroy(stefan_ipdg* hv) {
  IpdgHandle* h = reinterpret_cast<IpdgHandle*>(hv);
  if (!h) return ST_OK;
  if (h->phase_dev) cudaFree(h->phase_dev);
  h->plan.destroy();
  if (h->tail_ctr_dev) cudaFree(h->tail_ctr_dev);
  if (h->iface_dev) cudaFree(h->iface_dev);
  if (h->iface_partial_dev) cudaFree(h->iface_partial_dev);
  if (h->iface_ctr_dev) cudaFree(h->iface_ctr_dev);
  if (h->xh_dev) cudaFree(h->xh_dev);
  if (h->yh_dev) cudaFree(h->yh_dev);
  if (h->host_ready) {
    for (auto& s : h->hs) cudaStreamDestroy(s);
    for (auto& e : h->hev) cudaEventDestroy(e);
  }
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
  // halo columns may be columns of a neighbour's slab vector (peer memory): with 72-byte
  // cells (order 2) such a column can start 8 bytes off a 16-byte boundary, which the
  // kernel handles like any odd column chunk
  const uintptr_t hmask = h->order == 2 ? 7 : 15;
  STEFAN_REQUIRE((reinterpret_cast<uintptr_t>(x) & 15) == 0 && (reinterpret_cast<uintptr_t>(y) & 15) == 0 &&
                 (reinterpret_cast<uintptr_t>(x_lo) & hmask) == 0 && (reinterpret_cast<uintptr_t>(x_hi) & hmask) == 0,
                 "stefan_ipdg_apply: x, y must be 16-byte aligned (halo columns: 16, order 2: 8)");
  STEFAN_REQUIRE(algo >= 0 && algo <= 1, "stefan_ipdg_apply: unknown algo %d", algo);
  ApplyArgs a{x, x_lo, x_hi, y, h->phase_dev, h->nx, h->ny, 0, h->nx, nullptr, 0};
  return dispatch_apply(h, a, algo, static_cast<cudaStream_t>(stream));
}

// y = A x on a column slab whose halo columns are the neighbours' own vectors in peer memory,
// with the ordering fused into the kernel: it publishes sync->my_ready = epoch when it starts
// (everything the stream wrote to x before is final), the warps that fetch a neighbour's column
// first wait for that neighbour's ready flag, and the last CTA publishes sync->my_done = epoch.
int stefan_ipdg_apply_peer(stefan_ipdg* hv, const double* x, double* y, const double* x_lo, const double* x_hi,
                           const stefan_peer_sync* sync, void* stream) {
  IpdgHandle* h = reinterpret_cast<IpdgHandle*>(hv);
  STEFAN_REQUIRE(h && x && y && sync && x != y, "stefan_ipdg_apply_peer: bad argument");
  STEFAN_REQUIRE(h->has_lo == (x_lo != nullptr) && h->has_hi == (x_hi != nullptr),
                 "stefan_ipdg_apply_peer: halo columns must be given exactly where the handle has halo phases");
  STEFAN_REQUIRE(sync->epoch > 0 && sync->my_ready && sync->my_done && sync->timeout_ms > 0 &&
                 (!x_lo || sync->lo_ready) && (!x_hi || sync->hi_ready),
                 "stefan_ipdg_apply_peer: incomplete sync block");
  const uintptr_t hmask = h->order == 2 ? 7 : 15;
  STEFAN_REQUIRE((reinterpret_cast<uintptr_t>(x) & 15) == 0 && (reinterpret_cast<uintptr_t>(y) & 15) == 0 &&
                 (reinterpret_cast<uintptr_t>(x_lo) & hmask) == 0 && (reinterpret_cast<uintptr_t>(x_hi) & hmask) == 0,
                 "stefan_ipdg_apply_peer: x, y must be 16-byte aligned (halo columns: 16, order 2: 8)");
  ApplyArgs a{x, x_lo, x_hi, y, h->phase_dev, h->nx, h->ny, 0, h->nx, nullptr, 0};
  a.sync_my_ready = reinterpret_cast<unsigned long long*>(sync->my_ready);
  a.sync_my_done = reinterpret_cast<unsigned long long*>(sync->my_done);
  a.sync_lo_ready = reinterpret_cast<const unsigned long long*>(sync->lo_ready);
  a.sync_hi_ready = reinterpret_cast<const unsigned long long*>(sync->hi_ready);
  a.sync_epoch = sync->epoch;
  a.sync_epoch_dev = reinterpret_cast<unsigned long long*>(sync->epoch_dev);
  a.sync_timeout_ns = (unsigned long long)sync->timeout_ms * 1000000ull;
  a.sync_status = sync->status;
  for (int k = 0; k < 2; ++k) {
    a.sync_post_ready[k] = reinterpret_cast<unsigned long long*>(sync->post_ready[k]);
    a.sync_post_done[k] = reinterpret_cast<unsigned long long*>(sync->post_done[k]);
  }
  a.sync_chained = sync->chained ? 1 : 0;
  static const int dev_peer = [] { const char* e = getenv("STEFAN_DEV_PEER"); return e ? atoi(e) : 0; }();
  a.sync_dev = dev_peer;
  return dispatch_apply(h, a, 0, static_cast<cudaStream_t>(stream));
}

// The fused explicit step: unew = u + dt Minv (b - A u) in ONE pass of the march
// kernel -- u and b read once, unew written once (24 B per dof; apply + stefan_ipdg_euler_update
// move 48).  M = the block-diagonal DG mass matrix.  b may be NULL (no right-hand side).
// sync != NULL: the peer-halo form (see stefan_ipdg_apply_peer); u_lo / u_hi are then the
// neighbours' edge columns of THEIR u in peer memory.
int stefan_ipdg_step(stefan_ipdg* hv, const double* u, double* unew, const double* b, double dt, const double* u_lo,
                     const double* u_hi, const stefan_peer_sync* sync, void* stream) {
  IpdgHandle* h = reinterpret_cast<IpdgHandle*>(hv);
  STEFAN_REQUIRE(h && u && unew && u != unew, "stefan_ipdg_step: bad argument (in-place steps are not supported)");
  STEFAN_REQUIRE(dt != 0.0, "stefan_ipdg_step: dt must not be 0");
  STEFAN_REQUIRE(h->has_lo == (u_lo != nullptr) && h->has_hi == (u_hi != nullptr),
                 "stefan_ipdg_step: halo columns must be given exactly where the handle has halo phases");
  const uintptr_t hmask = h->order == 2 ? 7 : 15;
  const uintptr_t bmask = h->order == 2 ? 7 : 31;  // b cells are fetched with 256-bit loads (orders 1, 3)
  STEFAN_REQUIRE((reinterpret_cast<uintptr_t>(u) & 15) == 0 && (reinterpret_cast<uintptr_t>(unew) & 15) == 0 &&
                 (reinterpret_cast<uintptr_t>(b) & bmask) == 0 &&
                 (reinterpret_cast<uintptr_t>(u_lo) & hmask) == 0 && (reinterpret_cast<uintptr_t>(u_hi) & hmask) == 0,
                 "stefan_ipdg_step: u, unew must be 16-byte aligned, b 32-byte (order 2: 8), halo columns 16 (order 2: 8)");
  ApplyArgs a{u, u_lo, u_hi, unew, h->phase_dev, h->nx, h->ny, 0, h->nx, nullptr, 0};
  a.b = b;
  a.step_scale = dt / (0.25 * h->hx * h->hy);
  for (int i = 0; i < kMaxN1; ++i)
    for (int j = 0; j < kMaxN1; ++j) a.minv[i][j] = h->minv1d[i][j];
  if (sync != nullptr) {
    STEFAN_REQUIRE(sync->epoch > 0 && sync->my_ready && sync->my_done && sync->timeout_ms > 0 &&
                   (!u_lo || sync->lo_ready) && (!u_hi || sync->hi_ready),
                   "stefan_ipdg_step: incomplete sync block");
    a.sync_my_ready = reinterpret_cast<unsigned long long*>(sync->my_ready);
    a.sync_my_done = reinterpret_cast<unsigned long long*>(sync->my_done);
    a.sync_lo_ready = reinterpret_cast<const unsigned long long*>(sync->lo_ready);
    a.sync_hi_ready = reinterpret_cast<const unsigned long long*>(sync->hi_ready);
    for (int k = 0; k < 2; ++k) {
      a.sync_post_ready[k] = reinterpret_cast<unsigned long long*>(sync->post_ready[k]);
      a.sync_post_done[k] = reinterpret_cast<unsigned long long*>(sync->post_done[k]);
    }
    a.sync_chained = sync->chained ? 1 : 0;
    a.sync_epoch = sync->epoch;
    a.sync_epoch_dev = reinterpret_cast<unsigned long long*>(sync->epoch_dev);
    a.sync_timeout_ns = (unsigned long long)sync->timeout_ms * 1000000ull;
    a.sync_status = sync->status;
  }
  return dispatch_apply(h, a, 0, static_cast<cudaStream_t>(stream));
}

// Rows of the cell columns [col_begin, col_end) only (same arguments otherwise).  Lets a
// caller overlap the halo exchange of a column-slab partition with the interior:
// columns [1, nx-1) do not read x_lo / x_hi.
int stefan_ipdg_apply_columns(stefan_ipdg* hv, const double* x, double* y, const double* x_lo,
                              const double* x_hi, int32_t col_begin, int32_t col_end, int32_t algo,
                              void* stream) {
  IpdgHandle* h = reinterpret_cast<IpdgHandle*>(hv);
  STEFAN_REQUIRE(h && x && y && x != y, "stefan_ipdg_apply_columns: bad argument");
  STEFAN_REQUIRE(0 <= col_begin && col_begin <= col_end && col_end <= h->nx,
                 "stefan_ipdg_apply_columns: need 0 <= col_begin <= col_end <= nx");
  STEFAN_REQUIRE(algo >= 0 && algo <= 1, "stefan_ipdg_apply_columns: unknown algo %d", algo);
  const bool needs_lo = col_begin == 0, needs_hi = col_end == h->nx;
  STEFAN_REQUIRE(!(needs_lo && h->has_lo && !x_lo) && !(needs_hi && h->has_hi && !x_hi),
                 "stefan_ipdg_apply_columns: the halo column next to the requested range is missing");
  ApplyArgs a{x, h->has_lo ? x_lo : nullptr, h->has_hi ? x_hi : nullptr, y, h->phase_dev, h->nx, h->ny, col_begin, col_end, nullptr, 0};
  return dispatch_apply(h, a, algo, static_cast<cudaStream_t>(stream));
}

// y = A x with HOST buffers (pinned memory recommended): the mesh is cut into
// `nslabs` column slabs that are pipelined on three streams -- host->device copy
// of slab s+1, kernel on slab s, device->host copy of slab s-1 run concurrently --
// so the call is bounded by max(H2D, D2H) over PCIe, not their sum.
int stefan_ipdg_apply_host(stefan_ipdg* hv, const double* x_host, double* y_host, int32_t nslabs) {
  IpdgHandle* h = reinterpret_cast<IpdgHandle*>(hv);
  STEFAN_REQUIRE(h && x_host && y_host, "stefan_ipdg_apply_host: null argument");
  STEFAN_REQUIRE(!h->has_lo && !h->has_hi,
                 "stefan_ipdg_apply_host: handles with halo columns take device buffers");
  if (nslabs <= 0) nslabs = 32;  // measured: 16 slabs 5.38, 32 slabs 5.56 G dof/s (PCIe bound 5.84)
  nslabs = std::min<int>(std::min<int>(nslabs, 32), h->nx);
  if (!h->host_ready) {
    STEFAN_CUDA(cudaMalloc(&h->xh_dev, h->ndofs * sizeof(double)));
    STEFAN_CUDA(cudaMalloc(&h->yh_dev, h->ndofs * sizeof(double)));
    for (auto& s : h->hs) STEFAN_CUDA(cudaStreamCreateWithFlags(&s, cudaStreamNonBlocking));
    for (auto& e : h->hev) STEFAN_CUDA(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
    h->host_ready = true;
  }
  const int NF = (h->order + 1) * (h->order + 1);
  const size_t col = (size_t)h->ny * NF;  // doubles per column
  auto slab_begin = [&](int s) { return (int)((int64_t)h->nx * s / nslabs); };
  for (int s = 0; s <= nslabs; ++s) {
    if (s < nslabs) {  // upload slab s
      const int b = slab_begin(s), e = slab_begin(s + 1);
      STEFAN_CUDA(cudaMemcpyAsync(h->xh_dev + b * col, x_host + b * col, (size_t)(e - b) * col * sizeof(double),
                                  cudaMemcpyHostToDevice, h->hs[0]));
      STEFAN_CUDA(cudaEventRecord(h->hev[s], h->hs[0]));
    }
    if (s >= 1) {  // slab s-1 needs the first column of slab s: wait for that upload
      const int k = s - 1;
      STEFAN_CUDA(cudaStreamWaitEvent(h->hs[1], h->hev[std::min(s, nslabs - 1)], 0));
      ApplyArgs a{h->xh_dev, nullptr, nullptr, h->yh_dev, h->phase_dev, h->nx, h->ny, slab_begin(k), slab_begin(k + 1), nullptr, 0};
      int rc = dispatch_apply(h, a, 0, h->hs[1]);
      if (rc) return rc;
      STEFAN_CUDA(cudaEventRecord(h->hev[32 + k], h->hs[1]));
      STEFAN_CUDA(cudaStreamWaitEvent(h->hs[2], h->hev[32 + k], 0));
      const int b = slab_begin(k), e = slab_begin(k + 1);
      STEFAN_CUDA(cudaMemcpyAsync(y_host + b * col, h->yh_dev + b * col, (size_t)(e - b) * col * sizeof(double),
                                  cudaMemcpyDeviceToHost, h->hs[2]));
    }
  }
  STEFAN_CUDA(cudaStreamSynchronize(h->hs[2]));
  return ST_OK;
}

}  // extern "C"
