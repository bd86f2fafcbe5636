// Interior-penalty DG apply, "march" kernels (production path for every order).
// Included by ipdg.cu after its device helpers (ApplyArgs, mbarrier / bulk-copy
// wrappers, fixup_tail).
//
// Decomposition: one warp owns a strip of 30 cells in y (+1 halo cell each side =
// 32 lanes, one cell per lane) and marches along x over a segment of columns.
//
// What changed against ipdg_apply_stream / ipdg_apply_p1_tma: a cell never reads its
// neighbours' dofs.  In the 1-D line operator of ipdg_tables.h the off-diagonal block
// towards a neighbour is one dense row (the face node's row) plus one column (the
// neighbour's face node), so all a cell needs from a neighbour, per face line, is
//     t = the neighbour's trace value            (N1 values per face)
//     d = sum_b L0[b] x_nbr[b]  (or Rp)          (N1 values per face)
// i.e. 2 N1 numbers instead of NF.  The x-neighbours' (t, d) are carried in registers
// from the previous column / computed from the next column as it is loaded (it is
// needed next iteration anyway); the y-neighbours' (t, d) come from lanes -+1 by warp
// shuffles.  Consequences:
//   * every column chunk is read from shared memory exactly once (the lane's own
//     cell), so its ring slot is free for the next bulk copy immediately;
//   * shared-memory traffic per cell drops from 3 NF to NF doubles;
//   * registers hold X, XR and 6 N1 trace numbers instead of five full cells.
//
// Data movement: one bulk asynchronous copy (TMA) per column chunk into a per-warp
// ring guarded by mbarriers.
//   P = 1: 32-byte cells, 1-D bulk copy (cp.async.bulk), direct STG.256 stores.
//   P = 2: 72-byte cells: a chunk may start 8 bytes off a 16-byte boundary, the copy
//          then starts 8 bytes early (`sh`).  Results are staged in shared memory and
//          leave with ONE bulk store per column (cp.async.bulk shared -> global): the
//          per-lane 8-byte stores of the stream kernel touched every 32-byte sector
//          four times (ncu: 100 M store sectors for 100 M doubles).
//   P = 3: 128-byte cells: a dense stage would put all 32 lanes on the same banks, so
//          the chunk is fetched as a 2-D tensor tile {16 doubles, 32 cells} with
//          CU_TENSOR_MAP_SWIZZLE_128B (cp.async.bulk.tensor): the TMA unit writes the
//          16-byte pieces of row r at piece ^ (r & 7) and the lanes read them back
//          conflict-free with LDS.128.
// Coefficients are compile-time offsets into the __grid_constant__ table (constant
// bank operands of the DFMAs); the loop body exists once per phase.
#pragma once

#include <type_traits>

#include <cuda.h>  // CUtensorMap (types only; the encoder is fetched at run time)

#include "march_dev.cuh"

namespace stefan {

#ifndef MARCH_NST1
#define MARCH_NST1 2
#endif
#ifndef MARCH_NST2
#define MARCH_NST2 2
#endif
#ifndef MARCH_NST3
#define MARCH_NST3 2
#endif
#ifndef MARCH_WARPS1
#define MARCH_WARPS1 8
#endif
#ifndef MARCH_WARPS2
#define MARCH_WARPS2 8
#endif
#ifndef MARCH_WARPS3
#define MARCH_WARPS3 4
#endif
#ifndef MARCH_MINB1
#define MARCH_MINB1 2
#endif
#ifndef MARCH_MINB2
#define MARCH_MINB2 2
#endif
#ifndef MARCH_MINB3
#define MARCH_MINB3 2
#endif

template <int P>
struct MarchCfg;
template <>
struct MarchCfg<1> {
  static constexpr int CB = 32, STAGE = 32 * 32, NST = MARCH_NST1, WARPS = MARCH_WARPS1, MINB = MARCH_MINB1;
  static constexpr int OUTB = 0;  // direct stores
  static constexpr bool TENSOR = false;
};
template <>
struct MarchCfg<2> {
  // 32 cells x 72 B + 8 (lane shift of strip 0) + 8 (sh) rounded up to 16, + 16 over-copy
  static constexpr int CB = 72, STAGE = 32 * 72 + 32, NST = MARCH_NST2, WARPS = MARCH_WARPS2, MINB = MARCH_MINB2;
  static constexpr int OUTB = 30 * 72 + 16;  // one output buffer (two per warp)
  static constexpr bool TENSOR = false;
};
template <>
struct MarchCfg<3> {
  static constexpr int CB = 128, STAGE = 32 * 128, NST = MARCH_NST3, WARPS = MARCH_WARPS3, MINB = MARCH_MINB3;
  static constexpr int OUTB = 0;
  static constexpr bool TENSOR = true;
};

// CTAs per SM of the kernel <P, MODE>.  The fused step (MODE 1) of order 2 needs more than the 128 registers
// that two 8-warp CTAs leave per thread (138 B of spills in every column): one CTA of 192 registers per SM.
// Measured (100 M dofs, us per step; profiles/r2_p2_step_variants.log): b per cell / staged, 2 CTAs: 597 / 660;
// 1 CTA: 588 / 569.
#ifndef MARCH_MINB2_STEP
#define MARCH_MINB2_STEP 1
#endif
#ifndef MARCH_BSTAGE2
#define MARCH_BSTAGE2 1
#endif
template <int P, int MODE>
constexpr int march_minb() {
  return (P == 2 && MODE == 1) ? MARCH_MINB2_STEP : MarchCfg<P>::MINB;
}

// SegTab (ipdg_handle.h): one entry per CTA, {strip group, first column, one past the last
// column, -}, built by the host so that every CTA gets the same COST rather than the same
// column count (columns where a strip mixes phases run both phase bodies).
template <int P, int MODE = 0>
struct MarchSmem {
  using C = MarchCfg<P>;
  // P = 2 fused step: one strip of b (30 cells x 72 B), loaded coalesced and read back per cell
  static constexpr int BBUF = (P == 2 && MODE == 1 && MARCH_BSTAGE2) ? 30 * 72 + 16 : 0;
  static constexpr int WARP = C::NST * C::STAGE + 2 * C::OUTB + BBUF;
  // rings (1024-byte aligned for the swizzled tiles) + mbarriers + alignment slack
  static constexpr int TOTAL = C::WARPS * WARP + C::WARPS * C::NST * 8 + 1024;
};

// ---- the arithmetic of one interior cell of phase index S with same-phase neighbours ----
template <int P, int S>
struct MarchMath {
  static constexpr int N1 = P + 1;
  // d handed to the RIGHT neighbour (it is that cell's low-side moment): sum_b L0x[b] X[b][iy]
  static __device__ __forceinline__ void x_moment_lo(const IpTab<P>& tab, const double* X, double* d) {
#pragma unroll
    for (int iy = 0; iy < N1; ++iy) {
      double v = 0.0;
#pragma unroll
      for (int b = 0; b < N1; ++b) v += tab.L0[0][S][T_SAME][b] * X[b * N1 + iy];
      d[iy] = v;
    }
  }
  // d handed to the LEFT neighbour (its high-side moment): sum_b Rpx[b] X[b][iy]
  static __device__ __forceinline__ void x_moment_hi(const IpTab<P>& tab, const double* X, double* d) {
#pragma unroll
    for (int iy = 0; iy < N1; ++iy) {
      double v = 0.0;
#pragma unroll
      for (int b = 0; b < N1; ++b) v += tab.Rp[0][S][T_SAME][b] * X[b * N1 + iy];
      d[iy] = v;
    }
  }
  // handed to the cell ABOVE (its low-side moment) / BELOW (its high-side moment)
  static __device__ __forceinline__ void y_moment_lo(const IpTab<P>& tab, const double* X, double* d) {
#pragma unroll
    for (int ix = 0; ix < N1; ++ix) {
      double v = 0.0;
#pragma unroll
      for (int c = 0; c < N1; ++c) v += tab.L0[1][S][T_SAME][c] * X[ix * N1 + c];
      d[ix] = v;
    }
  }
  static __device__ __forceinline__ void y_moment_hi(const IpTab<P>& tab, const double* X, double* d) {
#pragma unroll
    for (int ix = 0; ix < N1; ++ix) {
      double v = 0.0;
#pragma unroll
      for (int c = 0; c < N1; ++c) v += tab.Rp[1][S][T_SAME][c] * X[ix * N1 + c];
      d[ix] = v;
    }
  }
  // rows of the cell: tL/dL .. tU/dU are the neighbours' traces / moments per face line
  static __device__ __forceinline__ void rows(const IpTab<P>& tab, const double* X, const double* tL,
                                              const double* dL, const double* tR, const double* dR,
                                              const double* tD, const double* dD, const double* tU,
                                              const double* dU, double* out) {
    double ax[N1 * N1], ay[N1 * N1];
#pragma unroll
    for (int ix = 0; ix < N1; ++ix) {
#pragma unroll
      for (int iy = 0; iy < N1; ++iy) {
        double v = 0.0;
#pragma unroll
        for (int b = 0; b < N1; ++b) v += tab.D[0][S][T_SAME][T_SAME][ix * N1 + b] * X[b * N1 + iy];
        if (ix == 0) v += dL[iy];
        else v += tab.Lc[0][S][T_SAME][ix] * tL[iy];
        if (ix == P) v += dR[iy];
        else v += tab.Rc[0][S][T_SAME][ix] * tR[iy];
        ax[ix * N1 + iy] = v;
        double u = 0.0;
#pragma unroll
        for (int c = 0; c < N1; ++c) u += tab.D[1][S][T_SAME][T_SAME][iy * N1 + c] * X[ix * N1 + c];
        if (iy == 0) u += dD[ix];
        else u += tab.Lc[1][S][T_SAME][iy] * tD[ix];
        if (iy == P) u += dU[ix];
        else u += tab.Rc[1][S][T_SAME][iy] * tU[ix];
        ay[ix * N1 + iy] = u;
      }
    }
#pragma unroll
    for (int ix = 0; ix < N1; ++ix) {
#pragma unroll
      for (int iy = 0; iy < N1; ++iy) {
        double v = 0.0;
#pragma unroll
        for (int c = 0; c < N1; ++c) v += tab.My[iy * N1 + c] * ax[ix * N1 + c];
#pragma unroll
        for (int b = 0; b < N1; ++b) v += tab.Mx[ix * N1 + b] * ay[b * N1 + iy];
        out[ix * N1 + iy] = v;
      }
    }
  }
};

// MODE 0: y = A x.  MODE 1: the fused explicit step y = x + dt Minv (b - A x) (SURVEY 2.4 K9:
// 24 B per dof -- x and b read once, y written once -- instead of the 48 B of apply + update).
template <int P, int MODE>
__global__ void __launch_bounds__(MarchCfg<P>::WARPS * 32, march_minb<P, MODE>())
ipdg_apply_march(const __grid_constant__ IpTab<P> tab, const ApplyArgs a, const int4* __restrict__ runs,
                 const int* __restrict__ run_ofs, const __grid_constant__ CUtensorMap tmx,
                 const __grid_constant__ SegTab segs) {
  using C = MarchCfg<P>;
  constexpr int N1 = P + 1;
  constexpr int NF = N1 * N1;
  constexpr int NST = C::NST;
  constexpr int CB = C::CB;
  extern __shared__ char smem_raw[];
  const int lane = threadIdx.x & 31;
  // broadcast from lane 0: the compiler then knows that everything derived from the warp
  // index is warp-uniform and keeps it in uniform registers (TMA operands included)
  const int warp = __shfl_sync(0xffffffffu, threadIdx.x >> 5, 0);
  // 1024-byte aligned base (swizzle atoms of the tensor tiles)
  const unsigned smem0 = (smem_addr(smem_raw) + 1023u) & ~1023u;
  const unsigned ring = smem0 + (unsigned)warp * MarchSmem<P, MODE>::WARP;
  [[maybe_unused]] const unsigned obuf = ring + NST * C::STAGE;  // P = 2: two output buffers
  [[maybe_unused]] const unsigned bbuf = obuf + 2 * C::OUTB;     // P = 2 fused step: the strip of b
  const unsigned bars = smem0 + (unsigned)C::WARPS * MarchSmem<P, MODE>::WARP + (unsigned)warp * NST * 8;
  griddep_launch();  // the next kernel of the stream may start its prologue while this one drains
  if (lane == 0) {
#pragma unroll
    for (int s = 0; s < NST; ++s)
      asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bars + 8 * s), "r"(1));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncwarp();

  const int nstrips = (a.ny + kStripCells - 1) / kStripCells;
  const int4 myseg = segs.e[blockIdx.x];
  const int group = myseg.x;
  const int strip = group * C::WARPS + warp;
  const int c0 = myseg.y, c1 = myseg.z;
  if (strip < nstrips && c0 < c1) {
    const int j0 = strip * kStripCells;
    const int j = j0 - 1 + lane;                      // this lane's cell
    const int jlo = max(j0 - 1, 0), jhi = min(j0 + kStripCells + 1, a.ny);
    const int rel0 = jlo - (j0 - 1);  // lane of the first staged cell (0, or 1 in strip 0)
    const unsigned nbytes = (unsigned)(jhi - jlo) * CB;  // bytes of one column chunk
    const size_t colbytes = (size_t)a.ny * CB;
    const char* const xb = reinterpret_cast<const char*>(a.x);
    // Lanes outside the mesh (j = -1 in strip 0, j >= ny in the top strip) and columns
    // that do not exist (k = -1 / nx without a halo) read whatever their ring slot
    // holds: every cell next to them is a boundary cell, i.e. an exception cell that
    // the loop never stores (it is recomputed by fixup_tail), so the values go nowhere.

    // ---- general forms (prologue, halo / missing columns, P = 2 array end) ----
    auto col_base = [&](int k) -> const char* {
      if (k >= 0 && k < a.nx) return xb + (size_t)k * colbytes + (size_t)jlo * CB;
      if (k < 0 && a.x_lo != nullptr) return reinterpret_cast<const char*>(a.x_lo) + (size_t)jlo * CB;
      if (k >= a.nx && a.x_hi != nullptr) return reinterpret_cast<const char*>(a.x_hi) + (size_t)jlo * CB;
      return nullptr;
    };
    // P = 2: the last chunk of an array whose byte length is 8 mod 16 cannot be copied to
    // its end in 16-byte units; the lane of the last cell patches its last double itself
    auto chunk_is_array_end = [&](int k) -> bool {
      if constexpr (P != 2) return false;
      if (jhi != a.ny) return false;
      if (k >= 0 && k < a.nx) return k == a.nx - 1;
      return true;  // halo columns are arrays of ny cells
    };
    auto issue_general = [&](int k, unsigned stage, unsigned bar) {
      if (k > c1 || lane != 0) return;
      const char* src = col_base(k);
      if (src == nullptr) {  // no such column: nothing to copy, complete the phase
        mbar_arrive_s(bar);
        return;
      }
      if (a.sync_epoch != 0 && (k < 0 || k >= a.nx) && !(a.sync_dev & 1))  // a neighbour's column: wait until it is published
        peer_flag_wait(k < 0 ? a.sync_lo_ready : a.sync_hi_ready,
                       peer_epoch(a.sync_epoch_dev, a.sync_epoch) - (unsigned long long)a.sync_chained,
                       a.sync_timeout_ns, a.sync_status);
      if constexpr (C::TENSOR) {
        if (k >= 0 && k < a.nx) {
          mbar_expect_tx_s(bar, (unsigned)C::STAGE);
          tensor2d_load_s(stage, &tmx, 0, (int)((int64_t)k * a.ny + (j0 - 1)), bar);
        } else {  // halo column: plain bulk copy, read back unswizzled
          mbar_expect_tx_s(bar, nbytes);
          bulk_load_s(stage + rel0 * CB, src, nbytes, bar);
        }
      } else if constexpr (P == 2) {
        const unsigned sh = (unsigned)(reinterpret_cast<uintptr_t>(src) & 15u);  // 0 or 8
        unsigned bytes = sh + nbytes;
        if (chunk_is_array_end(k)) bytes &= ~15u;
        else bytes = (bytes + 15u) & ~15u;
        mbar_expect_tx_s(bar, bytes);
        bulk_load_s(stage + rel0 * 80, src - sh, bytes, bar);
      } else {
        mbar_expect_tx_s(bar, nbytes);
        bulk_load_s(stage + rel0 * CB, src, nbytes, bar);
      }
    };
    // Slot release.  A ring slot may be refilled (an asynchronous-proxy write) only after
    // every lane's shared-memory loads of it have COMPLETED; having issued them is not
    // enough (with 12+ warps per SM a queued LDS was overtaken by the refill of an
    // L2-resident column: wrong results, non-deterministically).  The readers therefore
    // return a word h that depends on one destination register of every load instruction;
    // lane 0 adds  redux.xor(h) & a.zero  (= 0) to the address operands of the refill, so
    // the copy cannot be issued before all 32 lanes hold their data.
    auto hi32 = [](double v) -> unsigned { return (unsigned)__double2hiint(v); };
    auto read_general = [&](int k, unsigned stage, double* v) -> unsigned {
      unsigned h = 0;
      if constexpr (C::TENSOR) {
        const unsigned m = (k >= 0 && k < a.nx) ? (unsigned)(lane & 7) : 0u;
        const unsigned p = stage + lane * CB;
#pragma unroll
        for (int q = 0; q < NF / 2; ++q) {
          const double2 t = lds_v2(p + ((q ^ m) << 4));
          v[2 * q] = t.x;
          v[2 * q + 1] = t.y;
          h ^= hi32(t.x);
        }
      } else if constexpr (P == 2) {
        const char* src = col_base(k);
        const unsigned sh = (unsigned)(reinterpret_cast<uintptr_t>(src) & 15u);
        const unsigned p = stage + lane * CB + 8 * rel0 + sh;
#pragma unroll
        for (int q = 0; q < NF; ++q) {
          v[q] = lds_f64(p + 8 * q);
          h ^= hi32(v[q]);
        }
        if (chunk_is_array_end(k) && j == a.ny - 1 && src != nullptr) {
          // (sh + nbytes) was 8 mod 16: the copy stopped 8 bytes short of the array end.  A halo column may be peer
          // memory: this plain load is ordered after lane 0's flag acquire through the mbarrier this lane waited on
          // (acquire -> bulk copy issued -> complete_tx -> this lane's try_wait), it never runs ahead of the copy.
          if ((sh + nbytes) & 8u) v[NF - 1] = *reinterpret_cast<const double*>(src + nbytes - 8);
        }
      } else {
        const unsigned p = stage + lane * CB;
#pragma unroll
        for (int q = 0; q < NF / 2; ++q) {
          const double2 t = lds_v2(p + 16 * q);
          v[2 * q] = t.x;
          v[2 * q + 1] = t.y;
          h ^= hi32(t.x);
        }
      }
      return h;
    };
    auto released = [&](unsigned h) -> unsigned { return __reduce_xor_sync(0xffffffffu, h) & a.zero; };

    // run list of this strip (built by stefan_ipdg_create): {end column, mode, phase mask,
    // store mask}.  Bit `lane` of the store mask: this lane's cell is interior-uniform and
    // is stored by the loop; of the phase mask: it has phase code 2.  mode & 3: 0 nothing
    // to store, 1 / 2 every stored cell has phase code 1 / 2, 3 both occur (both phase
    // bodies run, each storing its own lanes); mode & 4: the stored lanes are contiguous.
    // the run that holds column c0: bisection over the strip's list (end columns increase);
    // walking it from the start cost a chain of dependent L2 loads per tile -- hundreds in the
    // strips that follow the interface
    const int4* rp;
    {
      int lo = run_ofs[strip], hi = run_ofs[strip + 1] - 1;  // the sentinels end at INT_MAX
      while (lo < hi) {
        const int mid = (lo + hi) >> 1;
        if (__ldg(&runs[mid].x) > c0) hi = mid;
        else lo = mid + 1;
      }
      rp = runs + lo;
    }
    int4 cur = ld_run(rp);
    int4 nxt = ld_run(rp + 1);
    bool mine1 = (cur.z >> lane) & 1;  // this lane's cell has phase code 2

    // everything above read only what stefan_ipdg_create wrote; from here on the kernel touches
    // the vectors, the flags and the queue counters: the previous kernel must be complete
    griddep_wait();
    if (a.dbg_times != nullptr && threadIdx.x == 0) {
      unsigned long long t;
      asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
      a.dbg_times[4 * blockIdx.x] = t;
    }
    // peer halos: everything the stream wrote before this kernel is final -> publish my epoch
    // (warp 0 of CTA 0 always owns a strip: tiles are never empty)
    // (chained mode: nobody waits for the ready value -- the gate is the previous epoch's done value -- and the
    // fence would sit on this warp's critical path: the value is stored without one, for bookkeeping only)
    if (a.sync_epoch != 0 && blockIdx.x == 0 && threadIdx.x == 0) {
      const unsigned long long epoch = peer_epoch(a.sync_epoch_dev, a.sync_epoch);
      if (a.sync_chained || (a.sync_dev & 2)) *reinterpret_cast<volatile unsigned long long*>(a.sync_my_ready) = epoch;
      else peer_flag_publish(a.sync_my_ready, epoch, a.sync_post_ready[0], a.sync_post_ready[1]);
    }

    // prologue: columns c0-1 .. c0+NST-2 into slots 0 .. NST-1
#pragma unroll
    for (int k = 0; k < NST; ++k) issue_general(c0 - 1 + k, ring + k * C::STAGE, bars + 8 * k);
    double X[NF], tL[N1], dL[N1];
    {
      double XLfull[NF];
      mbar_wait_s(bars + 0, 0);
      unsigned z = released(read_general(c0 - 1, ring, XLfull));
      issue_general(c0 - 1 + NST, ring + z, bars + z);
      mbar_wait_s(bars + 8, 0);
      z = released(read_general(c0, ring + C::STAGE, X));
      issue_general(c0 + NST, ring + C::STAGE + z, bars + 8 + z);
#pragma unroll
      for (int t = 0; t < N1; ++t) tL[t] = XLfull[P * N1 + t];
      double d1[N1];
      MarchMath<P, 0>::x_moment_lo(tab, XLfull, dL);
      MarchMath<P, 1>::x_moment_lo(tab, XLfull, d1);
#pragma unroll
      for (int t = 0; t < N1; ++t) dL[t] = mine1 ? d1[t] : dL[t];
    }

    if (a.dbg_times != nullptr && threadIdx.x == 0) {  // development: end of the prologue
      unsigned long long t;
      asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
      a.dbg_times[4 * blockIdx.x + 1] = t;
    }

    // ---- steady-state bookkeeping, all incremental ----
    int slot = 2 % NST;                               // slot of column c+1
    unsigned stage_r = ring + slot * C::STAGE;        // ... its stage
    unsigned bar_r = bars + 8 * slot;                 // ... its mbarrier
    unsigned par_r = (2 / NST) & 1;                   // ... its phase parity
    // columns k <= k_plain are issued / read by the short forms below
    const int k_plain = min(c1, a.nx - 1) - (P == 2 ? 1 : 0);
    const char* src_i = xb + (size_t)(c0 + 1 + NST) * colbytes + (size_t)jlo * CB;  // chunk of column c+1+NST
    [[maybe_unused]] const unsigned tog = (P == 2 && (a.ny & 1)) ? 8u : 0u;  // sh alternates when ny is odd
    [[maybe_unused]] unsigned sh_r = 0, sh_i = 0, shy = 0;
    if constexpr (P == 2) {
      sh_r = (unsigned)((reinterpret_cast<uintptr_t>(xb) + (size_t)(c0 + 1) * colbytes + (size_t)jlo * CB) & 15u);
      sh_i = (unsigned)(reinterpret_cast<uintptr_t>(src_i) & 15u);
    }
    [[maybe_unused]] int trow = 0;
    if constexpr (C::TENSOR) trow = (int)((int64_t)(c0 + 1 + NST) * a.ny + (j0 - 1));
    const unsigned own = (unsigned)lane * CB + (P == 2 ? 8u * rel0 : 0u);
    const unsigned dst_off = P == 2 ? rel0 * 80u : (C::TENSOR ? 0u : rel0 * (unsigned)CB);
    const size_t ycol = (size_t)a.ny * NF;  // doubles per column
    double* yp = a.y + (size_t)c0 * ycol + (size_t)j * NF;
    [[maybe_unused]] const double* bp = a.b + (size_t)c0 * ycol + (size_t)j * NF;
    [[maybe_unused]] const bool bp_valid = a.b != nullptr && lane >= 1 && lane <= kStripCells && j < a.ny;
    // P = 2 fused step: 72-byte cells make a per-cell read of b touch every 32-byte sector from several
    // instructions (measured: the fused step at 62 % of the roofline against 79 % for the apply).  The warp
    // reads the strip's 270 doubles of b COALESCED instead (lane + 32 r), one column ahead, parks them in shared
    // memory at the top of the column and every lane picks up its own cell for the epilogue.
    constexpr bool BSTAGE = (P == 2 && MODE == 1 && MARCH_BSTAGE2);
    constexpr int BR = BSTAGE ? (kStripCells * NF + 31) / 32 : 1;  // rounds of 32 doubles
    [[maybe_unused]] double Bnext[BR];
    [[maybe_unused]] const double* bchunk = a.b + (size_t)c0 * ycol + (size_t)j0 * NF;
    [[maybe_unused]] const int bvalid = a.b != nullptr ? min(kStripCells, a.ny - j0) * NF : 0;  // doubles of the strip
    [[maybe_unused]] auto load_b_chunk = [&]() {
#pragma unroll
      for (int r = 0; r < BR; ++r) {
        const int d = lane + 32 * r;
        Bnext[r] = d < bvalid ? __ldg(bchunk + d) : 0.0;
      }
      bchunk += ycol;
    };
    if constexpr (BSTAGE) load_b_chunk();
    [[maybe_unused]] char* gstrip = reinterpret_cast<char*>(a.y) + ((size_t)c0 * a.ny + j0) * CB;
    if constexpr (P == 2) shy = (unsigned)(reinterpret_cast<uintptr_t>(gstrip) & 15u);
    [[maybe_unused]] unsigned ob = obuf;

    // One column: Xc = column c (in registers), Xn receives column c+1.  The callers
    // alternate the two arrays, so no register copies are needed between columns.
    auto column = [&](int c, double* __restrict__ Xc, double* __restrict__ Xn) {
      // fused step: this cell's b, requested before the wait for the next column (in flight
      // behind it); only the lanes whose cells the loop may store
      [[maybe_unused]] double Bc[NF];
      if constexpr (BSTAGE) {
        // (the previous column's reads of bbuf precede these stores in program order)
#pragma unroll
        for (int r = 0; r < BR; ++r) {
          const int d = lane + 32 * r;
          if (d < kStripCells * NF) sts_f64(bbuf + 8 * d, Bnext[r]);
        }
        __syncwarp();
        if (c + 1 < c1) load_b_chunk();  // in flight behind the whole column
      } else if constexpr (MODE == 1) {
        if (bp_valid) {
          if constexpr (P == 2) {
#pragma unroll
            for (int q = 0; q < NF; ++q) Bc[q] = __ldg(bp + q);
          } else {
#pragma unroll
            for (int q = 0; q < NF; q += 4) ld_global_nc_v4(bp + q, Bc + q);
          }
        } else {
#pragma unroll
          for (int q = 0; q < NF; ++q) Bc[q] = 0.0;
        }
      }
      mbar_wait_s(bar_r, par_r);
      // where this lane's cell of column c+1 sits: plain columns differ from halo columns
      // only in the swizzle mask (P = 3) / the 8-byte shift (P = 2)
      unsigned p = stage_r + own, m = 0, h = 0;
      const bool plain = c + 1 <= k_plain;
      if constexpr (C::TENSOR) m = (plain || c + 1 < a.nx) ? (unsigned)(lane & 7) << 4 : 0u;
      if constexpr (P == 2) {
        p += sh_r;
        if (!plain) {
          const char* src = col_base(c + 1);
          p = stage_r + own + (unsigned)(reinterpret_cast<uintptr_t>(src) & 15u);
        }
      }
      if constexpr (P == 2) {
#pragma unroll
        for (int q = 0; q < NF; ++q) {
          Xn[q] = lds_f64(p + 8 * q);
          h ^= hi32(Xn[q]);
        }
        if (!plain && chunk_is_array_end(c + 1) && j == a.ny - 1) {
          const char* src = col_base(c + 1);
          // (sh + nbytes) was 8 mod 16: the copy stopped 8 bytes short of the array end
          if (src != nullptr && (((unsigned)(reinterpret_cast<uintptr_t>(src) & 15u) + nbytes) & 8u))
            Xn[NF - 1] = *reinterpret_cast<const double*>(src + nbytes - 8);
        }
      } else {
#pragma unroll
        for (int q = 0; q < NF / 2; ++q) {
          const double2 t = lds_v2(p + ((16u * q) ^ m));
          Xn[2 * q] = t.x;
          Xn[2 * q + 1] = t.y;
          h ^= hi32(t.x);
        }
      }
      {
        // every lane holds its cell (see "slot release" above): refill the slot
        const unsigned z = released(h);
        const unsigned stage_i = stage_r + z, bar_i = bar_r + z;
        if (c + 1 + NST <= k_plain) {
          if (lane == 0) {
            if constexpr (C::TENSOR) {
              mbar_expect_tx_s(bar_i, (unsigned)C::STAGE);
              tensor2d_load_s(stage_i, &tmx, 0, trow, bar_i);
            } else if constexpr (P == 2) {
              const unsigned bytes = (sh_i + nbytes + 15u) & ~15u;
              mbar_expect_tx_s(bar_i, bytes);
              bulk_load_s(stage_i + dst_off, src_i - sh_i, bytes, bar_i);
            } else {
              mbar_expect_tx_s(bar_i, nbytes);
              bulk_load_s(stage_i + dst_off, src_i, nbytes, bar_i);
            }
          }
        } else {
          issue_general(c + 1 + NST, stage_i, bar_i);
        }
      }
      src_i += colbytes;
      if constexpr (P == 2) {
        sh_r ^= tog;
        sh_i ^= tog;
      }
      if constexpr (C::TENSOR) trow += a.ny;
      if (++slot == NST) {
        slot = 0;
        stage_r = ring;
        bar_r = bars;
        par_r ^= 1;
      } else {
        stage_r += C::STAGE;
        bar_r += 8;
      }

      // the arithmetic, once per phase (compile-time table offsets)
      auto body = [&](auto phase_c) {
        constexpr int S = decltype(phase_c)::value;
        // lanes this body stores: the interior-uniform cells of phase index S
        const unsigned ms = S ? ((unsigned)cur.w & (unsigned)cur.z) : ((unsigned)cur.w & ~(unsigned)cur.z);
        const bool st = (ms >> lane) & 1;
        using M = MarchMath<P, S>;
        double out[NF];
        double tR[N1], dR[N1], sU[N1], sD[N1], tD[N1], dD[N1], tU[N1], dU[N1];
#pragma unroll
        for (int t = 0; t < N1; ++t) tR[t] = Xn[0 * N1 + t];
        M::x_moment_hi(tab, Xn, dR);
        M::y_moment_lo(tab, Xc, sU);
        M::y_moment_hi(tab, Xc, sD);
#pragma unroll
        for (int t = 0; t < N1; ++t) {
          // from the cell below (lane - 1): its top trace and its moment for the cell above it
          tD[t] = __shfl_up_sync(0xffffffffu, Xc[t * N1 + P], 1);
          dD[t] = __shfl_up_sync(0xffffffffu, sU[t], 1);
          // from the cell above (lane + 1)
          tU[t] = __shfl_down_sync(0xffffffffu, Xc[t * N1 + 0], 1);
          dU[t] = __shfl_down_sync(0xffffffffu, sD[t], 1);
        }
        M::rows(tab, Xc, tL, dL, tR, dR, tD, dD, tU, dU, out);
        if constexpr (BSTAGE) {
          const unsigned pb = bbuf + (unsigned)(lane >= 1 && lane <= kStripCells ? lane - 1 : 0) * CB;
#pragma unroll
          for (int q = 0; q < NF; ++q) Bc[q] = lds_f64(pb + 8 * q);
        }
        if constexpr (MODE == 1) step_epilogue<P>(a, Xc, Bc, out);

        if constexpr (C::OUTB == 0) {
          if (st) {
#pragma unroll
            for (int q = 0; q < NF; q += 4) st_global_v4(yp + q, out[q], out[q + 1], out[q + 2], out[q + 3]);
          }
        } else if (ms != 0 && (((ms >> (__ffs(ms) - 1)) + 1) & (ms >> (__ffs(ms) - 1))) == 0) {
          // P = 2, contiguous stored lanes [l0, l1): stage the results, one bulk store per
          // column.  The stored byte range of the strip is [b0, b1); a head / tail double
          // that does not fill a 16-byte unit is written by the lane that owns it.
          const unsigned l0 = __ffs(ms) - 1, l1 = 32 - __clz(ms);
          const unsigned b0 = (l0 - 1) * CB, b1 = (l1 - 1) * CB;
          const unsigned head = (shy + b0) & 8u, tail = (shy + b1) & 8u;
          if (lane == 0) tma_store_wait_read<1>();  // the buffer used two columns ago is drained
          __syncwarp();
          if (st) {
            const unsigned o = ob + shy + (lane - 1) * CB;
#pragma unroll
            for (int q = 0; q < NF; ++q) sts_f64(o + 8 * q, out[q]);
            if (head && (unsigned)lane == l0) yp[0] = out[0];
            if (tail && (unsigned)lane + 1 == l1) yp[NF - 1] = out[NF - 1];
          }
          fence_proxy_async();
          __syncwarp();
          if (lane == 0) {
            const unsigned s0 = b0 + head, s1 = b1 - tail;
            if (s1 > s0) bulk_store_s(gstrip + s0, ob + shy + s0, s1 - s0);
            tma_store_commit();
          }
          ob = (ob == obuf) ? obuf + C::OUTB : obuf;
        } else {
          // P = 2, stored lanes with holes (rare): plain stores
          if (st) {
#pragma unroll
            for (int q = 0; q < NF; ++q) yp[q] = out[q];
          }
        }
      };
      // mode 1 / 2 / 3 = bit 0 / bit 1 / both: every phase body is instantiated exactly once
      // (the code footprint matters: see the instruction-cache note at the end of the kernel)
      const int ph = cur.y & 3;
      if (ph & 1) body(std::integral_constant<int, 0>{});
      if (ph & 2) body(std::integral_constant<int, 1>{});
      yp += ycol;
      if constexpr (MODE == 1) bp += ycol;
      if constexpr (P == 2) {
        gstrip += colbytes;
        shy ^= tog;
      }

      // next column: its run, and the left-neighbour traces it needs from this column
      if (c + 1 >= cur.x) {
        cur = nxt;
        nxt = ld_run(++rp + 1);
        mine1 = (cur.z >> lane) & 1;
      }
#pragma unroll
      for (int t = 0; t < N1; ++t) tL[t] = Xc[P * N1 + t];
      const int phn = cur.y & 3;
      if (phn & 1) MarchMath<P, 0>::x_moment_lo(tab, Xc, dL);
      if (phn & 2) {
        double d1[N1];
        MarchMath<P, 1>::x_moment_lo(tab, Xc, d1);
#pragma unroll
        for (int t = 0; t < N1; ++t) dL[t] = (phn == 2 || mine1) ? d1[t] : dL[t];
      }
    };

    double Y[NF];  // the second cell buffer
#ifndef MARCH_PINGPONG3
#define MARCH_PINGPONG3 0
#endif
    if constexpr (P != 3 || MARCH_PINGPONG3) {
      // the two buffers alternate roles: no register copies between columns
      int c = c0;
      for (; c + 1 < c1; c += 2) {
        column(c, X, Y);
        column(c + 1, Y, X);
      }
      if (c < c1) column(c, X, Y);
    } else {
      // order 3: one copy of the (19 KB per phase) loop body instead of two, at the price of 32
      // register moves per column -- where warps of one SM run both phase bodies the
      // instruction cache is what limits (vertical-split mesh: 318 vs 352 GDOF/s single-phase)
#pragma unroll 1
      for (int c = c0; c < c1; ++c) {
        column(c, X, Y);
#pragma unroll
        for (int q = 0; q < NF; ++q) X[q] = Y[q];
      }
    }
    if constexpr (C::OUTB != 0) {
      if (lane == 0) tma_store_wait_read<0>();
      __syncwarp();
    }
  }
  if (a.dbg_times != nullptr) {  // time of the marching part (before the tail)
    __syncthreads();
    if (threadIdx.x == 0) {
      unsigned long long t;
      asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
      a.dbg_times[4 * blockIdx.x + 2] = t;
    }
  }
  griddep_wait();  // (warps without a strip have not waited yet)
  fixup_tail<P, MODE>(tab, a);
  if (a.dbg_times != nullptr) {  // development: end of the tail
    __syncthreads();
    if (threadIdx.x == 0) {
      unsigned long long t;
      asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
      a.dbg_times[4 * blockIdx.x + 3] = t;
    }
  }
  march_finish(a);
}

}  // namespace stefan
