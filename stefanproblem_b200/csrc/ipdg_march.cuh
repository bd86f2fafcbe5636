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

#include <cuda.h>  // CUtensorMap (types only; the encoder is fetched at run time)

namespace stefan {

#ifndef MARCH_NST1
#define MARCH_NST1 6
#endif
#ifndef MARCH_NST2
#define MARCH_NST2 5
#endif
#ifndef MARCH_NST3
#define MARCH_NST3 5
#endif
#ifndef MARCH_WARPS1
#define MARCH_WARPS1 8
#endif
#ifndef MARCH_WARPS2
#define MARCH_WARPS2 8
#endif
#ifndef MARCH_WARPS3
#define MARCH_WARPS3 8
#endif
#ifndef MARCH_MINB1
#define MARCH_MINB1 2
#endif
#ifndef MARCH_MINB2
#define MARCH_MINB2 1
#endif
#ifndef MARCH_MINB3
#define MARCH_MINB3 1
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

template <int P>
struct MarchSmem {
  using C = MarchCfg<P>;
  static constexpr int WARP = C::NST * C::STAGE + 2 * C::OUTB;
  // rings (1024-byte aligned for the swizzled tiles) + mbarriers + alignment slack
  static constexpr int TOTAL = C::WARPS * WARP + C::WARPS * C::NST * 8 + 1024;
};

__device__ __forceinline__ void tma_tensor2d_load(void* dst, const CUtensorMap* map, int c0, int c1, void* bar) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];" ::
          "r"(smem_addr(dst)),
      "l"(map), "r"(c0), "r"(c1), "r"(smem_addr(bar))
      : "memory");
}
__device__ __forceinline__ void tma_bulk_store(void* gdst, const void* ssrc, unsigned bytes) {
  asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(gdst),
               "r"(smem_addr(ssrc)), "r"(bytes)
               : "memory");
}
__device__ __forceinline__ void tma_store_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void tma_store_wait_read() {
  asm volatile("cp.async.bulk.wait_group.read %0;" ::"n"(N) : "memory");
}

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

template <int P>
__global__ void __launch_bounds__(MarchCfg<P>::WARPS * 32, MarchCfg<P>::MINB)
ipdg_apply_march(const __grid_constant__ IpTab<P> tab, const ApplyArgs a, const int2* __restrict__ runs,
                 const int* __restrict__ run_ofs, const __grid_constant__ CUtensorMap tmx) {
  using C = MarchCfg<P>;
  constexpr int N1 = P + 1;
  constexpr int NF = N1 * N1;
  constexpr int NST = C::NST;
  constexpr int CB = C::CB;
  extern __shared__ char smem_raw[];
  const int lane = threadIdx.x & 31;
  const int warp = threadIdx.x >> 5;
  // 1024-byte aligned base (swizzle atoms of the tensor tiles)
  char* const smem = smem_raw + ((1024u - (smem_addr(smem_raw) & 1023u)) & 1023u);
  char* const ring = smem + (size_t)warp * MarchSmem<P>::WARP;
  char* const obuf = ring + NST * C::STAGE;  // P = 2: two output buffers
  unsigned long long* const bars =
      reinterpret_cast<unsigned long long*>(smem + (size_t)C::WARPS * MarchSmem<P>::WARP) + warp * NST;
  if (lane == 0) {
#pragma unroll
    for (int s = 0; s < NST; ++s) mbar_init(&bars[s], 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncwarp();

  const int nstrips = (a.ny + kStripCells - 1) / kStripCells;
  const int ngroups = (nstrips + C::WARPS - 1) / C::WARPS;
  const int group = blockIdx.x % ngroups;
  const int seg = blockIdx.x / ngroups;
  const int strip = group * C::WARPS + warp;
  const int c0 = a.cbeg + (int)((int64_t)(a.cend - a.cbeg) * seg / a.nseg);
  const int c1 = a.cbeg + (int)((int64_t)(a.cend - a.cbeg) * (seg + 1) / a.nseg);
  if (strip < nstrips && c0 < c1) {
    const int j0 = strip * kStripCells;
    const int jhi_out = min(j0 + kStripCells, a.ny);  // one past the strip's top output cell
    const int j = j0 - 1 + lane;                      // this lane's cell
    const bool inmesh = j >= 0 && j < a.ny;
    const bool outlane = lane >= 1 && lane <= kStripCells && j < a.ny;
    const int jlo = max(j0 - 1, 0), jhi = min(j0 + kStripCells + 1, a.ny);
    const int rel0 = jlo - (j0 - 1);  // lane of the first staged cell (0, or 1 in strip 0)
    const size_t colcells = (size_t)a.ny;
    const char* const xb = reinterpret_cast<const char*>(a.x);

    // ---- issue of column k into its ring slot (lane 0 only does the work) ----
    // returns nothing; the reader recomputes where the data landed from (k, slot)
    auto col_base = [&](int k) -> const char* {
      if (k >= 0 && k < a.nx) return xb + ((size_t)k * colcells + jlo) * CB;
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
    auto issue = [&](int k, int slot) {
      if (k > c1) return;
      if (lane != 0) return;
      char* dst = ring + slot * C::STAGE;
      const char* src = col_base(k);
      if (src == nullptr) {  // no such column: nothing to copy, complete the phase
        mbar_arrive(&bars[slot]);
        return;
      }
      if constexpr (C::TENSOR) {
        if (k >= 0 && k < a.nx) {
          mbar_arrive_expect_tx(&bars[slot], (unsigned)C::STAGE);
          tma_tensor2d_load(dst, &tmx, 0, (int)((int64_t)k * a.ny + (j0 - 1)), &bars[slot]);
        } else {  // halo column: plain bulk copy, read back unswizzled
          const unsigned bytes = (unsigned)(jhi - jlo) * CB;
          mbar_arrive_expect_tx(&bars[slot], bytes);
          tma_bulk_load(dst + rel0 * CB, src, bytes, &bars[slot]);
        }
      } else if constexpr (P == 2) {
        const unsigned sh = (unsigned)(reinterpret_cast<uintptr_t>(src) & 15u);  // 0 or 8
        unsigned bytes = sh + (unsigned)(jhi - jlo) * CB;
        if (chunk_is_array_end(k)) bytes &= ~15u;
        else bytes = (bytes + 15u) & ~15u;
        mbar_arrive_expect_tx(&bars[slot], bytes);
        tma_bulk_load(dst + rel0 * 80, src - sh, bytes, &bars[slot]);
      } else {
        const unsigned bytes = (unsigned)(jhi - jlo) * CB;
        mbar_arrive_expect_tx(&bars[slot], bytes);
        tma_bulk_load(dst + rel0 * CB, src, bytes, &bars[slot]);
      }
    };
    auto wait_slot = [&](int slot, unsigned parity) {
      while (!mbar_try_wait(&bars[slot], parity)) {
      }
    };
    // this lane's cell of column k out of ring slot `slot`
    auto read_cell = [&](int k, int slot, double* v) {
      const char* st = ring + slot * C::STAGE;
      if constexpr (C::TENSOR) {
        const unsigned m = (k >= 0 && k < a.nx) ? (unsigned)(lane & 7) : 0u;
        const char* p = st + lane * CB;
#pragma unroll
        for (int q = 0; q < NF / 2; ++q) {
          const double2 t = *reinterpret_cast<const double2*>(p + ((q ^ m) << 4));
          v[2 * q] = t.x;
          v[2 * q + 1] = t.y;
        }
      } else if constexpr (P == 2) {
        const char* src = col_base(k);
        const unsigned sh = (unsigned)(reinterpret_cast<uintptr_t>(src) & 15u);
        const double* p = reinterpret_cast<const double*>(st + lane * CB + 8 * rel0 + sh);
#pragma unroll
        for (int q = 0; q < NF; ++q) v[q] = p[q];
        if (chunk_is_array_end(k) && j == a.ny - 1 && src != nullptr) {
          // (sh + n CB) was 8 mod 16: the copy stopped 8 bytes short of the array end
          const unsigned total = sh + (unsigned)(jhi - jlo) * CB;
          if (total & 8u) v[NF - 1] = *reinterpret_cast<const double*>(src + (size_t)(jhi - jlo) * CB - 8);
        }
      } else {
        const double2* p = reinterpret_cast<const double2*>(st + lane * CB);
#pragma unroll
        for (int q = 0; q < NF / 2; ++q) {
          const double2 t = p[q];
          v[2 * q] = t.x;
          v[2 * q + 1] = t.y;
        }
      }
      const bool have = inmesh && col_base(k) != nullptr;
#pragma unroll
      for (int q = 0; q < NF; ++q) v[q] = have ? v[q] : 0.0;
    };

    // run list of this strip (see ipdg_apply_p1_tma)
    const int2* rp = runs + run_ofs[strip];
    int2 cur = rp[0];
    while (cur.x <= c0) cur = *++rp;
    int2 nxt = rp[1];
    auto store_flag = [&](int f) {
      return outlane && (f & 3) != 0 && !((f & 4) && lane == 1) && !((f & 8) && j == jhi_out - 1);
    };
    bool store = store_flag(cur.y);

    // prologue: columns c0-1 .. c0+NST-2 into slots 0 .. NST-1
#pragma unroll
    for (int k = 0; k < NST; ++k) issue(c0 - 1 + k, k);
    double X[NF], tL[N1], dL[N1];
    {
      double XLfull[NF];
      wait_slot(0, 0);
      read_cell(c0 - 1, 0, XLfull);
      __syncwarp();
      issue(c0 - 1 + NST, 0);
      wait_slot(1, 0);
      read_cell(c0, 1, X);
      __syncwarp();
      issue(c0 + NST, 1);
#pragma unroll
      for (int t = 0; t < N1; ++t) tL[t] = XLfull[P * N1 + t];
      if ((cur.y & 3) == 2) MarchMath<P, 1>::x_moment_lo(tab, XLfull, dL);
      else MarchMath<P, 0>::x_moment_lo(tab, XLfull, dL);
    }

    int slot_r = 2 % NST;                       // slot of column c+1
    unsigned par_r = (2 / NST) & 1;             // its phase parity
    const size_t ycol = (size_t)a.ny * NF;      // doubles per column
    double* yp = a.y + (size_t)c0 * ycol + (size_t)j * NF;
    [[maybe_unused]] int obsel = 0;

    for (int c = c0; c < c1; ++c) {
      double XR[NF];
      wait_slot(slot_r, par_r);
      read_cell(c + 1, slot_r, XR);
      __syncwarp();  // every lane has its cell: the slot can be refilled
      issue(c + 1 + NST, slot_r);
      if (++slot_r == NST) {
        slot_r = 0;
        par_r ^= 1;
      }

      const int ph = cur.y & 3;
      if (ph != 0) {
        double out[NF];
        double tR[N1], dR[N1], sU[N1], sD[N1], tD[N1], dD[N1], tU[N1], dU[N1];
#pragma unroll
        for (int t = 0; t < N1; ++t) tR[t] = XR[0 * N1 + t];
        if (ph == 1) {
          MarchMath<P, 0>::x_moment_hi(tab, XR, dR);
          MarchMath<P, 0>::y_moment_lo(tab, X, sU);
          MarchMath<P, 0>::y_moment_hi(tab, X, sD);
        } else {
          MarchMath<P, 1>::x_moment_hi(tab, XR, dR);
          MarchMath<P, 1>::y_moment_lo(tab, X, sU);
          MarchMath<P, 1>::y_moment_hi(tab, X, sD);
        }
#pragma unroll
        for (int t = 0; t < N1; ++t) {
          // from the cell below (lane - 1): its top trace and its moment for the cell above it
          tD[t] = __shfl_up_sync(0xffffffffu, X[t * N1 + P], 1);
          dD[t] = __shfl_up_sync(0xffffffffu, sU[t], 1);
          // from the cell above (lane + 1)
          tU[t] = __shfl_down_sync(0xffffffffu, X[t * N1 + 0], 1);
          dU[t] = __shfl_down_sync(0xffffffffu, sD[t], 1);
        }
        if (ph == 1) MarchMath<P, 0>::rows(tab, X, tL, dL, tR, dR, tD, dD, tU, dU, out);
        else MarchMath<P, 1>::rows(tab, X, tL, dL, tR, dR, tD, dD, tU, dU, out);

        if constexpr (C::OUTB == 0) {
          if (store) {
#pragma unroll
            for (int q = 0; q < NF; q += 4) st_global_v4(yp + q, out[q], out[q + 1], out[q + 2], out[q + 3]);
          }
        } else {
          // P = 2: stage the strip's results, one bulk store per column.  The gmem range is
          // [first stored cell, one past the last stored cell) of the strip; head / tail
          // doubles that do not fill a 16-byte unit go out as plain stores.
          char* ob = obuf + obsel * C::OUTB;
          obsel ^= 1;
          if (lane == 0) tma_store_wait_read<1>();  // the buffer used two columns ago is drained
          __syncwarp();
          char* const gstrip = reinterpret_cast<char*>(a.y) + ((size_t)c * colcells + j0) * CB;
          const unsigned shy = (unsigned)(reinterpret_cast<uintptr_t>(gstrip) & 15u);
          if (outlane) {
            double* o = reinterpret_cast<double*>(ob + shy + (lane - 1) * CB);
#pragma unroll
            for (int q = 0; q < NF; ++q) o[q] = out[q];
          }
          fence_proxy_async();
          __syncwarp();
          if (lane == 0) {
            const int f = cur.y;
            unsigned b0 = (f & 4) ? CB : 0u;
            unsigned b1 = (unsigned)(jhi_out - j0) * CB - ((f & 8) ? CB : 0u);
            if (b1 > b0) {
              if ((shy + b0) & 8u) {
                *reinterpret_cast<double*>(gstrip + b0) = *reinterpret_cast<const double*>(ob + shy + b0);
                b0 += 8;
              }
              if ((shy + b1) & 8u) {
                b1 -= 8;
                *reinterpret_cast<double*>(gstrip + b1) = *reinterpret_cast<const double*>(ob + shy + b1);
              }
              if (b1 > b0) tma_bulk_store(gstrip + b0, ob + shy + b0, b1 - b0);
            }
            tma_store_commit();
          }
        }
      }
      yp += ycol;

      // next column: its run, and the left-neighbour traces it needs from this column
      if (c + 1 >= cur.x) {
        cur = nxt;
        nxt = *(++rp + 1);
        store = store_flag(cur.y);
      }
#pragma unroll
      for (int t = 0; t < N1; ++t) tL[t] = X[P * N1 + t];
      if ((cur.y & 3) == 2) MarchMath<P, 1>::x_moment_lo(tab, X, dL);
      else MarchMath<P, 0>::x_moment_lo(tab, X, dL);
#pragma unroll
      for (int q = 0; q < NF; ++q) X[q] = XR[q];
    }
    if constexpr (C::OUTB != 0) {
      if (lane == 0) tma_store_wait_read<0>();
      __syncwarp();
    }
  }
  fixup_tail<P>(tab, a);
}

}  // namespace stefan
