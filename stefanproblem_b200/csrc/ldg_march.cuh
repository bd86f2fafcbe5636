// Local-DG apply, "march" kernel (production path).  Included by ldg.cu.
//
// Same decomposition as ipdg_march.cuh: one warp owns a strip of 30 cells in y (+1 halo
// cell each side = 32 lanes, one cell per lane) and marches along x; every column chunk
// (32 cells, contiguous in memory: 3, 6.75 or 12 KiB) arrives by ONE bulk asynchronous
// copy (TMA) into a per-warp ring guarded by mbarriers.
//
// In the LDG rows of a cell (ldg_tables.h) a face neighbour enters only through the face
// nodes' T and normal flux component -- 2 (P+1) numbers per face.  Per column the lane
// therefore holds its own cell (3 NF doubles) in registers and gets
//   left  traces: kept in registers from the previous column,
//   right traces: 2 (P+1) doubles read from the next column's ring slot,
//   down / up traces: warp shuffles from lanes -1 / +1,
// computes its rows with compile-time table offsets (one body per phase), stores them,
// and then replaces its registers by the next column's cell out of the same ring slot,
// which is released for the next copy.  Each dof is read from HBM once (+2/30 halo
// re-reads) and written once: 16 B per dof-update.
//   P = 1: 96-byte cells, STG.256 stores.
//   P = 2: 216-byte cells: chunks may start 8 bytes off a 16-byte boundary (the copy then
//          starts 8 bytes early); results leave through a staged bulk store.
//   P = 3: 384-byte cells: a dense stage would put all 32 lanes on the same banks (measured:
//          157 GDOF/s), so the chunk is a 2-D tensor tile {16 doubles, 96 rows} fetched with
//          CU_TENSOR_MAP_SWIZZLE_128B (three 128-byte rows per cell; 290 GDOF/s); STG.256 stores.
// Cells that are not interior-uniform (march_plan.h) are recomputed by the tail.
#pragma once

#include <type_traits>

namespace stefan {

#ifndef LDGM_NST
#define LDGM_NST 2
#endif

// V = 0: few large CTAs, V = 1: the other shape.  Measured (GDOF/s, 10 M dofs / 2000^2 cells):
// P = 1: 8 warps x 2 CTAs 278 / 294, 4 warps x 2 CTAs - / 334 (x 4 CTAs: 243 / 306);
// P = 2: 8 warps x 1 CTA 195 / 333, 4 warps x 2 CTAs 241 / 325.  ldg.cu picks per mesh size.
template <int P, int V>
struct LdgMarchCfg;
template <int V>
struct LdgMarchCfg<1, V> {
  static constexpr int CB = 96, STAGE = 32 * 96, NST = LDGM_NST, WARPS = V ? 4 : 8, MINB = 2;
  static constexpr int OUTB = 0;
  static constexpr bool TENSOR = false;
};
template <int V>
struct LdgMarchCfg<2, V> {
  static constexpr int CB = 216, STAGE = 32 * 216 + 32, NST = LDGM_NST, WARPS = V ? 4 : 8, MINB = V ? 2 : 1;
  static constexpr int OUTB = 30 * 216 + 16;
  static constexpr bool TENSOR = false;
};
template <int V>
struct LdgMarchCfg<3, V> {
  static constexpr int CB = 384, STAGE = 32 * 384, NST = LDGM_NST, WARPS = V ? 4 : 8, MINB = V ? 2 : 1;
  static constexpr int OUTB = 0;
  // 384-byte cells put all 32 lanes on the same banks: the chunk is fetched as a 2-D tensor
  // tile {16 doubles, 96 rows of 128 B} with CU_TENSOR_MAP_SWIZZLE_128B (three rows per cell)
  static constexpr bool TENSOR = true;
};
template <int P, int V>
struct LdgMarchSmem {
  using C = LdgMarchCfg<P, V>;
  static constexpr int WARP = C::NST * C::STAGE + 2 * C::OUTB;
  static constexpr int TOTAL = C::WARPS * WARP + C::WARPS * C::NST * 8 + 1024;  // 1024: alignment of the swizzled tiles
};

struct LdgMarchArgs {
  const double* x;
  const double* x_lo;
  const double* x_hi;
  double* y;
  const int8_t* phase;  // padded
  int nx, ny;
  const int2* fix;  // {cell, face-type code}, see MarchPlan::fix_dev
  int nfix;
  unsigned zero;  // 0 at run time, unknown at compile time (slot release, see ipdg_march.cuh)
  unsigned long long* dbg_times;  // development only: per-CTA {start, prologue end, loop end, end} globaltimer, or null
  unsigned* tail_ctr;  // dynamic tail queue: [0] next fix-list index, [1] CTAs finished (reset by the last CTA)
};
__device__ __forceinline__ void ldg_griddep_launch() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
__device__ __forceinline__ void ldg_griddep_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }

__device__ __forceinline__ void ldg_dbg_stamp(const LdgMarchArgs& a, int slot) {
  if (a.dbg_times != nullptr && threadIdx.x == 0) {
    unsigned long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    a.dbg_times[4 * blockIdx.x + slot] = t;
  }
}

// Cells the loop does not store (boundary cells, cells next to the other phase).  The list is a
// QUEUE: every warp that is done with its strip takes a batch of cells through an atomic counter
// until the list is empty, so the warps whose tiles finish first absorb the tail.
// Round 2: COOPERATIVE form -- one lane per NODE, NF lanes per cell (8 / 3 / 2 cells per warp at
// orders 1 / 2 / 3).  A lane holds its node's (T, q_x, q_y); the 1-D line operators of
// ldg_tables.h become warp shuffles along the x- / y-lines of the cell, loads and stores are
// coalesced (24 contiguous bytes per lane), about 40 registers and ~200 instructions per cell
// instead of one thread working through a whole cell with run-time table indices (measured in
// round 1 / 2: 10-25 us per 32-cell chunk at orders 2, 3 -- the kernel's critical path).
template <int P>
__device__ __forceinline__ void ldg_fixup_tail(const LdgTab<P>& tab, const LdgMarchArgs& a) {
  constexpr int N1 = P + 1, NF = N1 * N1, CD = 3 * NF;
  constexpr int CW = 32 / NF;      // cells per warp and round
  constexpr int BATCH = 4;         // rounds per queue access; their loads are issued together
  if (a.nfix <= 0) return;
  const int lane_t = threadIdx.x & 31;
  const bool lane_live = lane_t < CW * NF;
  const int grp = lane_live ? lane_t / NF : 0;          // (idle lanes shadow group 0 and never store)
  const int node = lane_live ? lane_t - grp * NF : 0;
  const int ix = node / N1, iy = node % N1;
  const int gbase = grp * NF;
  for (;;) {
    unsigned base = 0;
    if (lane_t == 0) base = atomicAdd(a.tail_ctr, (unsigned)(CW * BATCH));
    base = __shfl_sync(0xffffffffu, base, 0);
    if (base >= (unsigned)a.nfix) break;
    // ---- all loads of the batch first (two dependent levels: record, then dofs) ----
    int2 rec[BATCH];
    double T[BATCH], Qx[BATCH], Qy[BATCH], nxT[BATCH], nxQ[BATCH], nyT[BATCH], nyQ[BATCH];
#pragma unroll
    for (int rd = 0; rd < BATCH; ++rd) rec[rd] = a.fix[min((int)base + rd * CW + grp, a.nfix - 1)];
#pragma unroll
    for (int rd = 0; rd < BATCH; ++rd) {
      const int64_t cell = rec[rd].x;
      const int code = rec[rd].y;
      const int i = (int)(cell / a.ny), j = (int)(cell - (int64_t)i * a.ny);
      const int tL = (code >> 1) & 3, tR = (code >> 3) & 3, tD = (code >> 5) & 3, tU = (code >> 7) & 3;
      const double* xc = a.x + cell * CD;
      const double* pl = (i > 0) ? xc - (size_t)a.ny * CD : a.x_lo + (size_t)j * CD;
      const double* pr = (i < a.nx - 1) ? xc + (size_t)a.ny * CD : a.x_hi + (size_t)j * CD;
      T[rd] = xc[3 * node];
      Qx[rd] = xc[3 * node + 1];
      Qy[rd] = xc[3 * node + 2];
      // the neighbours' face-node T and normal flux component, on the lanes of the face nodes (a lane is on
      // at most one x-face and one y-face unless N1 = 2, where ix = 0 and ix = P are different lanes anyway)
      nxT[rd] = nxQ[rd] = nyT[rd] = nyQ[rd] = 0.0;
      if (ix == 0 && tL != T_BDRY) { nxT[rd] = pl[3 * (P * N1 + iy)]; nxQ[rd] = pl[3 * (P * N1 + iy) + 1]; }
      if (ix == P && tR != T_BDRY) { nxT[rd] = pr[3 * iy]; nxQ[rd] = pr[3 * iy + 1]; }
      if (iy == 0 && tD != T_BDRY) { nyT[rd] = xc[3 * (ix * N1 + P) - CD]; nyQ[rd] = xc[3 * (ix * N1 + P) + 2 - CD]; }
      if (iy == P && tU != T_BDRY) { nyT[rd] = xc[3 * (ix * N1) + CD]; nyQ[rd] = xc[3 * (ix * N1) + 2 + CD]; }
    }
#pragma unroll
    for (int rd = 0; rd < BATCH; ++rd) {
      const int t0 = (int)base + rd * CW;
      if (t0 < a.nfix) {  // (warp-uniform)
        const bool live = lane_live && (t0 + grp) < a.nfix;
        const int code = rec[rd].y;
        const int S = code & 1, tL = (code >> 1) & 3, tR = (code >> 3) & 3, tD = (code >> 5) & 3, tU = (code >> 7) & 3;
        const double k = tab.kk[S];
        // ---- 1-D operators along the x-line (fixed iy) and the y-line (fixed ix) of the node
        double cqx = 0.0, g_x = 0.0, mqx = 0.0, cqy = 0.0, g_y = 0.0, mqy = 0.0;
#pragma unroll
        for (int b = 0; b < N1; ++b) {
          const int lx = gbase + b * N1 + iy, ly = gbase + ix * N1 + b;
          const double tx = __shfl_sync(0xffffffffu, T[rd], lx), qx = __shfl_sync(0xffffffffu, Qx[rd], lx);
          const double ty = __shfl_sync(0xffffffffu, T[rd], ly), qy = __shfl_sync(0xffffffffu, Qy[rd], ly);
          const double cx = tab.C[ix * N1 + b], cy = tab.C[iy * N1 + b];
          cqx += cx * tx;
          g_x += cx * qx;
          mqx += tab.Mx[ix * N1 + b] * qx;
          cqy += cy * ty;
          g_y += cy * qy;
          mqy += tab.My[iy * N1 + b] * qy;
        }
        g_x *= k;
        g_y *= k;
        // face terms on the face nodes' lanes (side 0 = low, 1 = high); N1 >= 2, so a node is on one x-face at most
        if (ix == 0 || ix == P) {
          const int side = ix == 0 ? 0 : 1, ty_ = ix == 0 ? tL : tR;
          cqx += tab.sqK[0][side][ty_] * T[rd] + tab.sqN[0][side][ty_] * nxT[rd];
          g_x += tab.tqK[S][0][side][ty_] * Qx[rd] + tab.tqN[S][0][side][ty_] * nxQ[rd] + tab.tpK[0][side][ty_] * T[rd] +
                 tab.tpN[0][side][ty_] * nxT[rd];
        }
        if (iy == 0 || iy == P) {
          const int side = iy == 0 ? 0 : 1, ty_ = iy == 0 ? tD : tU;
          cqy += tab.sqK[1][side][ty_] * T[rd] + tab.sqN[1][side][ty_] * nyT[rd];
          g_y += tab.tqK[S][1][side][ty_] * Qy[rd] + tab.tqN[S][1][side][ty_] * nyQ[rd] + tab.tpK[1][side][ty_] * T[rd] +
                 tab.tpN[1][side][ty_] * nyT[rd];
        }
        const double a_x = cqx + mqx, a_y = cqy + mqy;
        // ---- mass matrix across the other direction
        double oT = 0.0, oqx = 0.0, oqy = 0.0;
#pragma unroll
        for (int c = 0; c < N1; ++c) {
          const int ly = gbase + ix * N1 + c, lx = gbase + c * N1 + iy;
          const double my = tab.My[iy * N1 + c], mx = tab.Mx[ix * N1 + c];
          oqx += my * __shfl_sync(0xffffffffu, a_x, ly);
          oT += my * __shfl_sync(0xffffffffu, g_x, ly);
          oqy += mx * __shfl_sync(0xffffffffu, a_y, lx);
          oT += mx * __shfl_sync(0xffffffffu, g_y, lx);
        }
        if (live) {
          double* yc = a.y + (int64_t)rec[rd].x * CD + 3 * node;
          yc[0] = oT;
          yc[1] = oqx;
          yc[2] = oqy;
        }
      }
    }
  }
}

template <int P, int V>
__global__ void __launch_bounds__(LdgMarchCfg<P, V>::WARPS * 32, LdgMarchCfg<P, V>::MINB)
ldg_apply_march(const __grid_constant__ LdgTab<P> tab, const LdgMarchArgs a, const int4* __restrict__ runs,
                const int* __restrict__ run_ofs, const __grid_constant__ SegTab segs,
                const __grid_constant__ CUtensorMap tmx) {
  using C = LdgMarchCfg<P, V>;
  constexpr int N1 = P + 1, NF = N1 * N1, CD = 3 * NF;
  constexpr int NST = C::NST, CB = C::CB;
  constexpr bool ODD = (CB % 16) != 0;  // cells are only 8-byte aligned (P = 2)
  extern __shared__ char smem_raw[];
  const int lane = threadIdx.x & 31;
  const int warp = __shfl_sync(0xffffffffu, threadIdx.x >> 5, 0);  // warp-uniform for the compiler
  const unsigned smem0 = (smem_addr(smem_raw) + 1023u) & ~1023u;
  const unsigned ring = smem0 + (unsigned)warp * LdgMarchSmem<P, V>::WARP;
  [[maybe_unused]] const unsigned obuf = ring + NST * C::STAGE;
  const unsigned bars = smem0 + (unsigned)C::WARPS * LdgMarchSmem<P, V>::WARP + (unsigned)warp * NST * 8;
  if (lane == 0) {
#pragma unroll
    for (int s = 0; s < NST; ++s) asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bars + 8 * s), "r"(1));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncwarp();
  ldg_griddep_launch();  // the next kernel of the stream may start its prologue while this one drains

  const int nstrips = (a.ny + kStripCells - 1) / kStripCells;
  const int4 myseg = segs.e[blockIdx.x];
  const int strip = myseg.x * C::WARPS + warp;
  const int c0 = myseg.y, c1 = myseg.z;
  if (strip < nstrips && c0 < c1) {
    const int j0 = strip * kStripCells;
    const int j = j0 - 1 + lane;
    const int jlo = max(j0 - 1, 0), jhi = min(j0 + kStripCells + 1, a.ny);
    const int rel0 = jlo - (j0 - 1);
    const unsigned nbytes = (unsigned)(jhi - jlo) * CB;
    const size_t colbytes = (size_t)a.ny * CB;
    const char* const xb = reinterpret_cast<const char*>(a.x);
    // Lanes outside the mesh and columns that do not exist read whatever their ring slot
    // holds: every cell next to them is a boundary cell, which the loop never stores.

    auto col_base = [&](int k) -> const char* {
      if (k >= 0 && k < a.nx) return xb + (size_t)k * colbytes + (size_t)jlo * CB;
      if (k < 0 && a.x_lo != nullptr) return reinterpret_cast<const char*>(a.x_lo) + (size_t)jlo * CB;
      if (k >= a.nx && a.x_hi != nullptr) return reinterpret_cast<const char*>(a.x_hi) + (size_t)jlo * CB;
      return nullptr;
    };
    // ODD cells: the last chunk of an array whose byte length is 8 mod 16 cannot be copied to
    // its end in 16-byte units; the lane of the last cell patches its last double itself
    auto chunk_is_array_end = [&](int k) -> bool {
      if constexpr (!ODD) return false;
      if (jhi != a.ny) return false;
      if (k >= 0 && k < a.nx) return k == a.nx - 1;
      return true;
    };
    const unsigned dst_off = ODD ? rel0 * (unsigned)(CB + 8) : rel0 * (unsigned)CB;
    const unsigned own = (unsigned)lane * CB + (ODD ? 8u * rel0 : 0u);
    auto issue = [&](int k, unsigned stage, unsigned bar) {
      if (k > c1 || lane != 0) return;
      const char* src = col_base(k);
      if (src == nullptr) {
        mbar_arrive_s(bar);
        return;
      }
      if constexpr (C::TENSOR) {
        if (k >= 0 && k < a.nx) {
          mbar_expect_tx_s(bar, (unsigned)C::STAGE);
          tensor2d_load_s(stage, &tmx, 0, 3 * (int)((int64_t)k * a.ny + (j0 - 1)), bar);
        } else {  // halo column: plain bulk copy, read back unswizzled
          mbar_expect_tx_s(bar, nbytes);
          bulk_load_s(stage + dst_off, src, nbytes, bar);
        }
      } else if constexpr (ODD) {
        const unsigned sh = (unsigned)(reinterpret_cast<uintptr_t>(src) & 15u);
        unsigned bytes = sh + nbytes;
        if (chunk_is_array_end(k)) bytes &= ~15u;
        else bytes = (bytes + 15u) & ~15u;
        mbar_expect_tx_s(bar, bytes);
        bulk_load_s(stage + dst_off, src - sh, bytes, bar);
      } else {
        mbar_expect_tx_s(bar, nbytes);
        bulk_load_s(stage + dst_off, src, nbytes, bar);
      }
    };
    // shared address of this lane's cell of column k in `stage`
    auto cell_addr = [&](int k, unsigned stage) -> unsigned {
      unsigned p = stage + own;
      if constexpr (ODD) p += (unsigned)(reinterpret_cast<uintptr_t>(col_base(k)) & 15u);
      return p;
    };
    // shared address of double `d` (compile-time) of this lane's cell whose base is p; swizzled
    // tiles (TENSOR, in-mesh columns): 16-byte piece c of row r sits at piece c ^ (r & 7)
    [[maybe_unused]] const unsigned m0 = ((3u * lane + 0u) & 7u) << 4, m1 = ((3u * lane + 1u) & 7u) << 4,
                                    m2 = ((3u * lane + 2u) & 7u) << 4;
    auto elem_addr = [&](unsigned p, bool swz, int d) -> unsigned {
      if constexpr (C::TENSOR) {
        const int part = d / 16, piece = (d % 16) / 2, sub = (d % 2) * 8;
        const unsigned m = swz ? (part == 0 ? m0 : (part == 1 ? m1 : m2)) : 0u;
        return p + part * 128 + (((unsigned)piece << 4) ^ m) + sub;
      } else {
        return p + 8 * d;
      }
    };
    auto swizzled = [&](int k) -> bool { return C::TENSOR && k >= 0 && k < a.nx; };
    auto hi32 = [](double v) -> unsigned { return (unsigned)__double2hiint(v); };
    // the lane's whole cell; returns the release word (see ipdg_march.cuh, "slot release")
    auto read_cell = [&](int k, unsigned stage, double* v) -> unsigned {
      unsigned h = 0;
      const unsigned p = cell_addr(k, stage);
      if constexpr (ODD) {
#pragma unroll
        for (int q = 0; q < CD; ++q) {
          v[q] = lds_f64(p + 8 * q);
          h ^= hi32(v[q]);
        }
        if (chunk_is_array_end(k) && j == a.ny - 1) {
          const char* src = col_base(k);
          if (src != nullptr && (((unsigned)(reinterpret_cast<uintptr_t>(src) & 15u) + nbytes) & 8u))
            v[CD - 1] = *reinterpret_cast<const double*>(src + nbytes - 8);
        }
      } else {
        const bool swz = swizzled(k);
#pragma unroll
        for (int q = 0; q < CD / 2; ++q) {
          const double2 t = lds_v2(elem_addr(p, swz, 2 * q));
          v[2 * q] = t.x;
          v[2 * q + 1] = t.y;
          h ^= hi32(t.x);
        }
      }
      return h;
    };
    auto released = [&](unsigned h) -> unsigned { return __reduce_xor_sync(0xffffffffu, h) & a.zero; };

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
    // up to here only the handle's own tables were read; the vectors and the queue counters may
    // have been written by the previous kernel of the stream
    ldg_griddep_wait();
    ldg_dbg_stamp(a, 0);

    // prologue: columns c0-1 .. c0+NST-2 into slots 0 .. NST-1
#pragma unroll
    for (int k = 0; k < NST; ++k) issue(c0 - 1 + k, ring + k * C::STAGE, bars + 8 * k);
    double U[CD];
    double tLT[N1], tLQ[N1];  // left neighbour: T and q_x at its ix = P nodes
    {
      mbar_wait_s(bars + 0, 0);
      const unsigned p = cell_addr(c0 - 1, ring);
      unsigned h = 0;
#pragma unroll
      for (int q = 0; q < N1; ++q) {
        tLT[q] = lds_f64(elem_addr(p, swizzled(c0 - 1), 3 * (P * N1 + q) + 0));
        tLQ[q] = lds_f64(elem_addr(p, swizzled(c0 - 1), 3 * (P * N1 + q) + 1));
        h ^= hi32(tLT[q]) ^ hi32(tLQ[q]);
      }
      unsigned z = released(h);
      issue(c0 - 1 + NST, ring + z, bars + z);
      const int s1 = 1 % NST;
      mbar_wait_s(bars + 8 * s1, (1 / NST) & 1);
      z = released(read_cell(c0, ring + s1 * C::STAGE, U));
      issue(c0 + NST, ring + s1 * C::STAGE + z, bars + 8 * s1 + z);
    }

    ldg_dbg_stamp(a, 1);
    int slot = 2 % NST;
    unsigned stage_r = ring + slot * C::STAGE, bar_r = bars + 8 * slot, par_r = (2 / NST) & 1;
    double* yp = a.y + ((size_t)c0 * a.ny + j) * CD;
    [[maybe_unused]] char* gstrip = reinterpret_cast<char*>(a.y) + ((size_t)c0 * a.ny + j0) * CB;
    [[maybe_unused]] unsigned ob = obuf;

    for (int c = c0; c < c1; ++c) {
      mbar_wait_s(bar_r, par_r);
      const unsigned pn = cell_addr(c + 1, stage_r);
      [[maybe_unused]] double tRT[N1], tRQ[N1];  // right neighbour: T and q_x at its ix = 0 nodes
      if constexpr (P != 3) {  // (order 3 fetches them when its last x-line needs them)
#pragma unroll
        for (int q = 0; q < N1; ++q) {
          tRT[q] = lds_f64(elem_addr(pn, swizzled(c + 1), 3 * q + 0));
          tRQ[q] = lds_f64(elem_addr(pn, swizzled(c + 1), 3 * q + 1));
        }
      }
      auto body = [&](auto phase_c) {
        constexpr int S = decltype(phase_c)::value;
        const unsigned ms = S ? ((unsigned)cur.w & (unsigned)cur.z) : ((unsigned)cur.w & ~(unsigned)cur.z);
        const bool st = (ms >> lane) & 1;
        if constexpr (P == 3) {
          // order 3: rows leave as they complete (ldg_rows_progressive: no spills)
          ldg_rows_progressive<P, S>(
              tab, U, tLT, tLQ,
              [&](int ix, double& dT, double& dQ, double& uT, double& uQ) {
                dT = __shfl_up_sync(0xffffffffu, U[3 * (ix * N1 + P) + 0], 1);
                dQ = __shfl_up_sync(0xffffffffu, U[3 * (ix * N1 + P) + 2], 1);
                uT = __shfl_down_sync(0xffffffffu, U[3 * (ix * N1 + 0) + 0], 1);
                uQ = __shfl_down_sync(0xffffffffu, U[3 * (ix * N1 + 0) + 2], 1);
              },
              [&](double* rT, double* rQ) {
#pragma unroll
                for (int q = 0; q < N1; ++q) {
                  rT[q] = lds_f64(elem_addr(pn, swizzled(c + 1), 3 * q + 0));
                  rQ[q] = lds_f64(elem_addr(pn, swizzled(c + 1), 3 * q + 1));
                }
              },
              [&](int jx, const double* rows) {
                if (st) {
#pragma unroll
                  for (int q = 0; q < 3 * N1; q += 4)
                    st_global_v4(yp + 3 * N1 * jx + q, rows[q], rows[q + 1], rows[q + 2], rows[q + 3]);
                }
              });
        } else {
        double tDT[N1], tDQ[N1], tUT[N1], tUQ[N1];
#pragma unroll
        for (int q = 0; q < N1; ++q) {
          // from the cell below (lane - 1): T and q_y at its iy = P nodes; from above: iy = 0
          tDT[q] = __shfl_up_sync(0xffffffffu, U[3 * (q * N1 + P) + 0], 1);
          tDQ[q] = __shfl_up_sync(0xffffffffu, U[3 * (q * N1 + P) + 2], 1);
          tUT[q] = __shfl_down_sync(0xffffffffu, U[3 * (q * N1 + 0) + 0], 1);
          tUQ[q] = __shfl_down_sync(0xffffffffu, U[3 * (q * N1 + 0) + 2], 1);
        }
        double out[CD];
        ldg_rows_streamed<P, S>(tab, U, tLT, tLQ, tRT, tRQ, tDT, tDQ, tUT, tUQ, out);
        if constexpr (C::OUTB == 0) {
          if (st) {
#pragma unroll
            for (int q = 0; q < CD; q += 4) st_global_v4(yp + q, out[q], out[q + 1], out[q + 2], out[q + 3]);
          }
        } else if (ms != 0 && (((ms >> (__ffs(ms) - 1)) + 1) & (ms >> (__ffs(ms) - 1))) == 0) {
          // contiguous stored lanes [l0, l1): staged, one bulk store per column; a head / tail
          // double that does not fill a 16-byte unit is written by the lane that owns it
          const unsigned shy = (unsigned)(reinterpret_cast<uintptr_t>(gstrip) & 15u);
          const unsigned l0 = __ffs(ms) - 1, l1 = 32 - __clz(ms);
          const unsigned b0 = (l0 - 1) * CB, b1 = (l1 - 1) * CB;
          const unsigned head = (shy + b0) & 8u, tail = (shy + b1) & 8u;
          if (lane == 0) tma_store_wait_read<1>();
          __syncwarp();
          if (st) {
            const unsigned o = ob + shy + (lane - 1) * CB;
#pragma unroll
            for (int q = 0; q < CD; ++q) sts_f64(o + 8 * q, out[q]);
            if (head && (unsigned)lane == l0) yp[0] = out[0];
            if (tail && (unsigned)lane + 1 == l1) yp[CD - 1] = out[CD - 1];
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
          if (st) {
#pragma unroll
            for (int q = 0; q < CD; ++q) yp[q] = out[q];
          }
        }
        }  // P != 3
      };
      // mode 1 / 2 / 3 = bit 0 / bit 1 / both: every phase body is instantiated exactly once
      // (instruction-cache footprint)
      const int ph = cur.y & 3;
      if (ph & 1) body(std::integral_constant<int, 0>{});
      if (ph & 2) body(std::integral_constant<int, 1>{});
      yp += (size_t)a.ny * CD;
      if constexpr (C::OUTB != 0) gstrip += colbytes;
      if (c + 1 >= cur.x) {
        cur = nxt;
        nxt = ld_run(++rp + 1);
      }
      // this column becomes the left neighbour; the next column's cell replaces it
#pragma unroll
      for (int q = 0; q < N1; ++q) {
        tLT[q] = U[3 * (P * N1 + q) + 0];
        tLQ[q] = U[3 * (P * N1 + q) + 1];
      }
      const unsigned z = released(read_cell(c + 1, stage_r, U));
      issue(c + 1 + NST, stage_r + z, bar_r + z);
      if (++slot == NST) {
        slot = 0;
        stage_r = ring;
        bar_r = bars;
        par_r ^= 1;
      } else {
        stage_r += C::STAGE;
        bar_r += 8;
      }
    }
    if constexpr (C::OUTB != 0) {
      if (lane == 0) tma_store_wait_read<0>();
      __syncwarp();
    }
  }
  if (a.dbg_times != nullptr) {
    __syncthreads();
    ldg_dbg_stamp(a, 2);
  }
  ldg_griddep_wait();  // (warps without a strip have not waited yet)
  ldg_fixup_tail<P>(tab, a);
  if (a.dbg_times != nullptr) {
    __syncthreads();
    ldg_dbg_stamp(a, 3);
  }
  // the last CTA to get here resets the tail queue for the next launch
  __syncthreads();
  if (threadIdx.x == 0) {
    __threadfence();
    if (atomicAdd(a.tail_ctr + 1, 1u) == gridDim.x - 1) {
      a.tail_ctr[0] = 0;
      a.tail_ctr[1] = 0;
    }
  }
}

}  // namespace stefan
