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
#include "ipdg_handle.h"

namespace stefan {


#ifndef P1_MINB
#define P1_MINB 2
#endif
#ifndef P1_NST
#define P1_NST 6
#endif
template <int P>
struct StreamCfg;
template <>
struct StreamCfg<1> {
  static constexpr int PIECE = 16, CELLB = 32, STRIDE = 48, NST = P1_NST, WARPS = 8, MINB = P1_MINB;
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
  const uint8_t* flags;  // [nstrips][nxpad] strip classification (stream kernel)
  int nxpad;             // nx rounded up to 16 (flag rows are 16-byte aligned)
  int nx, ny;
  int nseg;  // x segments (stream kernels)
  int cbeg, cend;  // columns to produce (a sub-range when the host path pipelines slabs)
  const int32_t* fix;  // cells the stream kernels leave to their tail (see fixup_tail)
  int nfix;
  unsigned zero;  // 0 at run time, unknown at compile time (see ipdg_march.cuh, slot release)
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

// Cells the streaming loops do not produce (they skip the store): boundary cells,
// cells next to the other phase, and every cell of a (strip, column) that mixes
// phases or has such a column next to it.  They are disjoint from what the loops
// write, so each CTA simply works through its share of the list after its strips:
// one thread per cell, general tables, neighbours straight from global memory.
template <int P>
__device__ __forceinline__ void fixup_tail(const IpTab<P>& tab, const ApplyArgs& a) {
  constexpr int NF = (P + 1) * (P + 1);
  for (int t = blockIdx.x * blockDim.x + threadIdx.x; t < a.nfix; t += gridDim.x * blockDim.x) {
    const int64_t cell = a.fix[t];
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
    ip_cell_rows<P>(tab, s == 1 ? 0 : 1, ip_face_type(s, sL), ip_face_type(s, sR),
                    ip_face_type(s, sD), ip_face_type(s, sU), X, XL, XR, XD, XU, out);
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

// 32 bytes per lane in one instruction (STG.256, sm_100): a warp writes whole
// 32-byte sectors; two STG.128 per lane would send every sector to L2 twice, half
// filled (measured: L1->XBAR request path 66% busy, 43% of it stores).
__device__ __forceinline__ void st_global_v4(double* p, double a, double b, double c, double d) {
  asm volatile("st.global.v4.f64 [%0], {%1,%2,%3,%4};" ::"l"(p), "d"(a), "d"(b), "d"(c), "d"(d)
               : "memory");
}

template <int P>
__device__ __forceinline__ void store_cell(double* dst, const double* v) {
  constexpr int NF = (P + 1) * (P + 1);
  if constexpr (NF % 4 == 0) {
#pragma unroll
    for (int k = 0; k < NF; k += 4) st_global_v4(dst + k, v[k], v[k + 1], v[k + 2], v[k + 3]);
  } else {
#pragma unroll
    for (int k = 0; k < NF; ++k) dst[k] = v[k];
  }
}

// phase byte, zero-extended by the load itself (no conversion instruction that
// would wait on the load right after it is issued)
__device__ __forceinline__ unsigned ld_phase(const uint8_t* p) {
  unsigned v;
  asm volatile("ld.global.nc.u8 %0, [%1];" : "=r"(v) : "l"(p));
  return v;
}

template <int P>
__global__ void __launch_bounds__(StreamCfg<P>::WARPS * 32, StreamCfg<P>::MINB)
ipdg_apply_stream(const __grid_constant__ IpTab<P> tab, const ApplyArgs a) {
  using C = StreamCfg<P>;
  constexpr int NF = (P + 1) * (P + 1);
  constexpr int NST = C::NST;
  constexpr int FLAG_OFF = 32 * C::STRIDE;        // the column's strip flag rides with its data
  constexpr int STAGE_BYTES = 32 * C::STRIDE + 16;
  constexpr int NPIECE = C::CELLB / C::PIECE;     // pieces per lane per column
  extern __shared__ __align__(128) char smem[];

  const int lane = threadIdx.x & 31;
  const int warp = threadIdx.x >> 5;
  const int nstrips = (a.ny + kStripCells - 1) / kStripCells;
  const int ngroups = (nstrips + C::WARPS - 1) / C::WARPS;
  const int group = blockIdx.x % ngroups;
  const int seg = blockIdx.x / ngroups;
  const int strip = group * C::WARPS + warp;
  const int c0 = a.cbeg + (int)((int64_t)(a.cend - a.cbeg) * seg / a.nseg);
  const int c1 = a.cbeg + (int)((int64_t)(a.cend - a.cbeg) * (seg + 1) / a.nseg);
  if (strip < nstrips && c0 < c1) {
  char* ring = smem + (size_t)warp * NST * STAGE_BYTES;
  const int j0 = strip * kStripCells;  // first output cell of the strip
  const int j = j0 - 1 + lane;         // this lane's cell
  const bool outlane = lane >= 1 && lane <= kStripCells && j < a.ny;
  const int jhi_out = min(j0 + kStripCells, a.ny);  // one past the strip's top output cell
  const int jlo = max(j0 - 1, 0), jhi = min(j0 + kStripCells + 1, a.ny);
  const int nchunk = (jhi - jlo) * NPIECE;  // 16- or 8-byte pieces in one column chunk
  const int rel0 = jlo - (j0 - 1);          // first staged cell, in lane units
  const size_t colbytes = (size_t)a.ny * NF * sizeof(double);

  // per-lane constants: ring offsets of the pieces this lane copies, of its own
  // cell and of its y-neighbours
  int dsto[NPIECE];
#pragma unroll
  for (int q = 0; q < NPIECE; ++q) {
    const int byte = (lane + 32 * q) * C::PIECE;
    dsto[q] = (rel0 + byte / C::CELLB) * C::STRIDE + byte % C::CELLB;
  }
  const int own = lane * C::STRIDE;
  const int dn = max(lane - 1, 0) * C::STRIDE;
  const int up = min(lane + 1, 31) * C::STRIDE;

  // cells of this strip that lie outside the mesh are never staged: zero them once
  if (j < 0 || j >= a.ny) {
#pragma unroll
    for (int st = 0; st < NST; ++st)
      for (int k = 0; k < C::STRIDE; k += 8)
        *reinterpret_cast<double*>(ring + st * STAGE_BYTES + own + k) = 0.0;
  }
  __syncwarp();

  // source of this lane's first piece in column c (valid for 0 <= c < nx)
  const char* const src0 =
      reinterpret_cast<const char*>(a.x) + (size_t)jlo * (NF * sizeof(double)) + lane * C::PIECE;
  // per (strip, column) classification computed at handle creation:
  // 1 / 2 = every output cell of the strip is an interior cell of that phase code
  // whose four neighbours have the same phase; 0 = anything else.  The flag rides
  // with the column data: the 16-byte block holding the flags of columns
  // [16 floor(c/16), +16) is staged by a .cg cp.async like the data (measured: a
  // 4-byte .ca cp.async per column costs 35% of the kernel, an LDG in the loop waits
  // behind every cp.async in flight because L1TEX returns in order).
  const uint8_t* const flag0 = a.flags + (size_t)strip * a.nxpad;
  auto stage = [&](const char* src, char* dst, const uint8_t* flag) {
#pragma unroll
    for (int q = 0; q < NPIECE; ++q) {
#ifdef EXP_NOLOAD
      if (lane + 32 * q < nchunk && a.nseg < 0) {
#else
      if (lane + 32 * q < nchunk) {
#endif
        if constexpr (C::PIECE == 16)
          cp_async16(dst + dsto[q], src + q * 32 * C::PIECE);
        else
          cp_async8(dst + dsto[q], src + q * 32 * C::PIECE);
      }
    }
    if (lane == 0 && flag != nullptr) cp_async16(dst + FLAG_OFF, flag);
  };
  auto issue_general = [&](int c, int slot) {
    // stage column c into ring slot `slot`; columns beyond c1 are not needed
    if (c <= c1) {
      char* dst = ring + slot * STAGE_BYTES;
      const char* halo = reinterpret_cast<const char*>(c < 0 ? a.x_lo : a.x_hi);
      if (c >= 0 && c < a.nx) {
        stage(src0 + (size_t)c * colbytes, dst, flag0 + (c & ~15));
      } else if (halo != nullptr) {
        stage(halo + (size_t)jlo * (NF * sizeof(double)) + lane * C::PIECE, dst, nullptr);
      } else {
        // no such column (domain boundary): the slot must read as zeros
        for (int k = 0; k < C::STRIDE; k += 8) *reinterpret_cast<double*>(dst + own + k) = 0.0;
      }
    }
    cp_async_commit();
  };

  // prologue: fill the ring with columns c0-1 .. c0+NST-2 (column c0-1+k -> slot k)
#pragma unroll
  for (int k = 0; k < NST; ++k) issue_general(c0 - 1 + k, k);

  double X[NF], XL[NF];
  cp_async_wait<NST - 2>();
  __syncwarp();
  load_cell<P>(ring + 0 * STAGE_BYTES + own, XL);
  load_cell<P>(ring + 1 * STAGE_BYTES + own, X);
  __syncwarp();

  const int csteady = min(c1, a.nx - 1) - (NST - 1);  // last c whose prefetch column is a plain one
  const char* srcn = src0 + (size_t)(c0 + NST - 1) * colbytes;  // column c + NST - 1
  double* yp = a.y + ((size_t)c0 * a.ny + j) * NF;
  int slot_c = 1;  // ring slot of column c
  P1Coef coef;
  unsigned coef_phase = 0;
  if constexpr (P == 1) {
    p1_load_coef(tab, 0, coef);
    coef_phase = 1;
  }
  for (int c = c0; c < c1; ++c) {
    const int slot_l = slot_c == 0 ? NST - 1 : slot_c - 1;
    const int slot_r = slot_c == NST - 1 ? 0 : slot_c + 1;
    // prefetch column c+NST-1 into the slot of column c-1, which nobody reads any more
    if (c <= csteady) {
      stage(srcn, ring + slot_l * STAGE_BYTES, flag0 + ((c + NST - 1) & ~15));
      cp_async_commit();
    } else {
      issue_general(c + NST - 1, slot_l);
    }
    srcn += colbytes;
    cp_async_wait<NST - 2>();  // column c+1 has landed
    __syncwarp();

    double XR[NF], XD[NF], XU[NF], out[NF];
    const char* stc = ring + slot_c * STAGE_BYTES;
    const unsigned fb = *reinterpret_cast<const uint8_t*>(stc + FLAG_OFF + (c & 15));
    const unsigned f0 = fb & 3u;  // phase code; bits 2 / 3: do not store the bottom / top cell
    load_cell<P>(ring + slot_r * STAGE_BYTES + own, XR);
    load_cell<P>(stc + dn, XD);
    load_cell<P>(stc + up, XU);
    // f0 = phase code all 30 cells are treated as (their exceptions -- boundary,
    // interface or other-phase cells -- are recomputed by ipdg_apply_fixup), or 0 =
    // the strip mixes phases in this column and is left to the fix-up kernel entirely
    if (f0 != 0) {
      if constexpr (P == 1) {
        if (f0 != coef_phase) {  // rare: the strip crossed into the other phase
          p1_load_coef(tab, f0 == 1 ? 0 : 1, coef);
          coef_phase = f0;
        }
#ifdef EXP_NOCOMPUTE
#pragma unroll
        for (int k = 0; k < NF; ++k) out[k] = X[k] + XL[k] + XR[k] + XD[k] + XU[k] * coef.dx0;
#else
        p1_uniform_rows(coef, X, XL, XR, XD, XU, out);
#endif
      } else {
        if (f0 == 1)
          ip_cell_rows<P>(tab, 0, T_SAME, T_SAME, T_SAME, T_SAME, X, XL, XR, XD, XU, out);
        else
          ip_cell_rows<P>(tab, 1, T_SAME, T_SAME, T_SAME, T_SAME, X, XL, XR, XD, XU, out);
      }
    }
#ifdef EXP_NOSTORE
    if (outlane && f0 != 0 && out[0] == 1.2345e300) store_cell<P>(yp, out);
#else
    if (outlane && f0 != 0 && !((fb & 4u) && lane == 1) && !((fb & 8u) && j == jhi_out - 1))
      store_cell<P>(yp, out);
#endif
    yp += (size_t)a.ny * NF;
#pragma unroll
    for (int k = 0; k < NF; ++k) {
      XL[k] = X[k];
      X[k] = XR[k];
    }
    slot_c = slot_r;
    __syncwarp();  // everyone is done with column c's slot before it is refilled
  }
  cp_async_wait<0>();
  }
  fixup_tail<P>(tab, a);
}

// --------------------------------------------------------------------------
// order-1 TMA kernel (production path for P = 1)
// --------------------------------------------------------------------------
// Same decomposition again (warp = strip of 30 cells + 1 halo cell each side,
// marching along x), but every column chunk -- 32 cells = 1 KiB, contiguous in
// memory -- is fetched by ONE bulk asynchronous copy (cp.async.bulk, the TMA unit;
// SASS UBLKCP) issued by lane 0 into a per-warp ring of kTmaStages stages, each
// guarded by an mbarrier.  Why TMA: measured with the arithmetic removed
// (scripts/micro/stream_pattern.cu), 16 warps/SM moving 1 KiB per warp per column
// reach 5.0 TB/s with LDG.256 into registers (the six scoreboards per warp cap the
// loads a warp can keep in flight), 4.5 TB/s with 16-byte cp.async (LDGSTS does not
// merge the two halves of a sector, so L2 serves every sector twice), and 5.9 TB/s
// with bulk copies, against 6.5 TB/s for a plain copy.
constexpr int kTmaStages = 6;
constexpr int kTmaWarps = 8;
constexpr int kTmaStageBytes = 32 * 32;  // 32 cells x 32 B, dense (bulk copies cannot pad)
constexpr int kTmaWarpSmem = kTmaStages * kTmaStageBytes + 128;  // ring + mbarriers

__device__ __forceinline__ unsigned smem_addr(const void* p) {
  return (unsigned)__cvta_generic_to_shared(p);
}
__device__ __forceinline__ void mbar_init(void* bar, int count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_addr(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_arrive_expect_tx(void* bar, unsigned bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_addr(bar)),
               "r"(bytes)
               : "memory");
}
__device__ __forceinline__ void mbar_arrive(void* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_addr(bar)) : "memory");
}
__device__ __forceinline__ void tma_bulk_load(void* dst, const void* src, unsigned bytes, void* bar) {
  asm volatile(
      "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
          smem_addr(dst)),
      "l"(src), "r"(bytes), "r"(smem_addr(bar))
      : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(void* bar, unsigned parity) {
  unsigned ok;
  asm volatile(
      "{\n.reg .pred p;\nmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\nselp.u32 %0, 1, 0, p;\n}"
      : "=r"(ok)
      : "r"(smem_addr(bar)), "r"(parity)
      : "memory");
  return ok != 0;
}
__device__ __forceinline__ void fence_proxy_async() {
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}

__global__ void __launch_bounds__(kTmaWarps * 32, 2)
ipdg_apply_p1_tma(const __grid_constant__ IpTab<1> tab, const ApplyArgs a,
                  const int2* __restrict__ runs, const int* __restrict__ run_ofs) {
  constexpr int NST = kTmaStages;
  extern __shared__ __align__(128) char smem[];
  const int lane = threadIdx.x & 31;
  const int warp = threadIdx.x >> 5;
  char* const ring = smem + (size_t)warp * kTmaWarpSmem;
  unsigned long long* const bars = reinterpret_cast<unsigned long long*>(ring + NST * kTmaStageBytes);
  if (lane == 0) {
#pragma unroll
    for (int s = 0; s < NST; ++s) mbar_init(&bars[s], 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncwarp();

  const int nstrips = (a.ny + kStripCells - 1) / kStripCells;
  const int ngroups = (nstrips + kTmaWarps - 1) / kTmaWarps;
  const int group = blockIdx.x % ngroups;
  const int seg = blockIdx.x / ngroups;
  const int strip = group * kTmaWarps + warp;
  const int c0 = a.cbeg + (int)((int64_t)(a.cend - a.cbeg) * seg / a.nseg);
  const int c1 = a.cbeg + (int)((int64_t)(a.cend - a.cbeg) * (seg + 1) / a.nseg);
  if (strip < nstrips && c0 < c1) {
  const int j0 = strip * kStripCells;
  const int jhi_out = min(j0 + kStripCells, a.ny);  // one past the strip's top output cell
  const int j = j0 - 1 + lane;
  const bool outlane = lane >= 1 && lane <= kStripCells && j < a.ny;
  const int jlo = max(j0 - 1, 0), jhi = min(j0 + kStripCells + 1, a.ny);
  const unsigned bytes = (unsigned)(jhi - jlo) * 32u;  // one column chunk
  const int dst0 = (jlo - (j0 - 1)) * 32;              // where the chunk starts inside a stage
  const size_t colstride = (size_t)a.ny * 4;           // doubles per column
  const int own = lane * 32;
  const int dn = max(lane - 1, 0) * 32;
  const int up = min(lane + 1, 31) * 32;

  // cells of the strip outside the mesh are never written by a copy: zero them once
  if (j < 0 || j >= a.ny) {
#pragma unroll
    for (int s = 0; s < NST; ++s) {
      *reinterpret_cast<double2*>(ring + s * kTmaStageBytes + own) = make_double2(0.0, 0.0);
      *reinterpret_cast<double2*>(ring + s * kTmaStageBytes + own + 16) = make_double2(0.0, 0.0);
    }
  }
  fence_proxy_async();
  __syncwarp();

  const double* const src0 = a.x + (size_t)jlo * 4;
  auto issue = [&](int c, int slot) {
    // column c into ring slot `slot` (columns beyond c1 are not needed)
    if (c > c1) return;
    char* dst = ring + slot * kTmaStageBytes;
    const double* col = nullptr;
    if (c >= 0 && c < a.nx)
      col = src0 + (size_t)c * colstride;
    else if (c < 0 && a.x_lo != nullptr)
      col = a.x_lo + (size_t)jlo * 4;
    else if (c >= a.nx && a.x_hi != nullptr)
      col = a.x_hi + (size_t)jlo * 4;
    if (col != nullptr) {
      if (lane == 0) {
        mbar_arrive_expect_tx(&bars[slot], bytes);
        tma_bulk_load(dst + dst0, col, bytes, &bars[slot]);
      }
    } else {
      // no such column (domain boundary): the slot reads as zeros
      *reinterpret_cast<double2*>(dst + own) = make_double2(0.0, 0.0);
      *reinterpret_cast<double2*>(dst + own + 16) = make_double2(0.0, 0.0);
      fence_proxy_async();
      __syncwarp();
      if (lane == 0) mbar_arrive(&bars[slot]);
    }
  };
  auto wait_slot = [&](int slot, unsigned parity) {
    while (!mbar_try_wait(&bars[slot], parity)) {
    }
  };
  auto lds4 = [&](const char* p, double* v) {
    const double2 t0 = *reinterpret_cast<const double2*>(p);
    const double2 t1 = *reinterpret_cast<const double2*>(p + 16);
    v[0] = t0.x; v[1] = t0.y; v[2] = t1.x; v[3] = t1.y;
  };

  // run list of this strip: (end column, flag) with increasing end; flag & 3 = phase
  // code all cells are treated as (0: none, the column is left to the tail), bit 2 /
  // bit 3 = the bottom / top cell of the strip is an exception and is not stored
  const int2* rp = runs + run_ofs[strip];
  int2 cur = rp[0];
  while (cur.x <= c0) cur = *++rp;
  int2 nxt = rp[1];  // the list ends with sentinel runs, so rp[1] always exists
  P1Coef coef;
  p1_load_coef(tab, (cur.y & 3) == 2 ? 1 : 0, coef);
  bool store = outlane && (cur.y & 3) != 0 && !((cur.y & 4) && lane == 1) &&
               !((cur.y & 8) && j == jhi_out - 1);

  // prologue: columns c0-1 .. c0+NST-2 into slots 0 .. NST-1
#pragma unroll
  for (int k = 0; k < NST; ++k) issue(c0 - 1 + k, k);
  double X[4], XL[4];
  wait_slot(0, 0);
  wait_slot(1, 0);
  lds4(ring + 0 * kTmaStageBytes + own, XL);
  lds4(ring + 1 * kTmaStageBytes + own, X);
  __syncwarp();

  double* yp = a.y + (size_t)c0 * colstride + (size_t)j * 4;
  int slot_c = 1;              // ring slot of column c
  unsigned par_r = 0;          // phase parity of the slot of column c+1
  for (int c = c0; c < c1; ++c) {
    const int slot_l = slot_c == 0 ? NST - 1 : slot_c - 1;
    const int slot_r = slot_c == NST - 1 ? 0 : slot_c + 1;
    // refill the slot of column c-1 (nobody reads it any more) with column c+NST-1
    issue(c + NST - 1, slot_l);
    // wait for column c+1
    if (slot_r == 0) par_r ^= 1;  // the ring wrapped: slot 0 is in its next phase
    wait_slot(slot_r, par_r);

    double XR[4], XD[4], XU[4], out[4];
    const char* stc = ring + slot_c * kTmaStageBytes;
    lds4(ring + slot_r * kTmaStageBytes + own, XR);
    lds4(stc + dn, XD);
    lds4(stc + up, XU);
    if (c >= cur.x) {  // rare: next run of this strip
      cur = nxt;
      nxt = *(++rp + 1);
      if ((cur.y & 3) != 0) p1_load_coef(tab, (cur.y & 3) == 2 ? 1 : 0, coef);
      store = outlane && (cur.y & 3) != 0 && !((cur.y & 4) && lane == 1) &&
              !((cur.y & 8) && j == jhi_out - 1);
    }
    if ((cur.y & 3) != 0) {
      p1_uniform_rows(coef, X, XL, XR, XD, XU, out);
      if (store) st_global_v4(yp, out[0], out[1], out[2], out[3]);
    }
    yp += colstride;
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      XL[k] = X[k];
      X[k] = XR[k];
    }
    slot_c = slot_r;
    __syncwarp();  // everyone is done with column c's slot before it is refilled
  }
  }
  fixup_tail<1>(tab, a);
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

template <int P>
static int launch_march(const IpdgHandle* h, const IpTab<P>& tab, const ApplyArgs& args, int grid, cudaStream_t st) {
  using C = MarchCfg<P>;
  constexpr int smem = MarchSmem<P>::TOTAL;
  static bool attr_set = false;
  if (!attr_set) {
    STEFAN_CUDA(cudaFuncSetAttribute(ipdg_apply_march<P>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    attr_set = true;
  }
  CUtensorMap tm;
  std::memset(&tm, 0, sizeof(tm));
  if (C::TENSOR) {
    int rc = make_cell_tensor_map(args.x, (int64_t)h->nx * h->ny, &tm);
    if (rc) return rc;
  }
  ipdg_apply_march<P><<<grid, C::WARPS * 32, smem, st>>>(tab, args, h->runs_dev, h->run_ofs_dev, tm);
  return ST_OK;
}

template <int P>
static int launch_apply(const IpdgHandle* h, const ApplyArgs& args0, int algo, cudaStream_t st) {
  using C = StreamCfg<P>;
  const IpTab<P>& tab = *static_cast<const IpTab<P>*>(h->tab_host);
  ApplyArgs args = args0;
  const int ncol = args.cend - args.cbeg;
  if (ncol <= 0) return ST_OK;
  if (algo == 1) {
    const int64_t ncells = (int64_t)ncol * h->ny;
    int blocks = (int)std::min<int64_t>((ncells + 127) / 128, (int64_t)sm_count() * 32);
    ipdg_apply_simple<P><<<blocks, 128, 0, st>>>(tab, args);
    STEFAN_CUDA(cudaGetLastError());
    return ST_OK;
  }
  // exception cells of the produced columns (the list is sorted by column)
  args.fix = h->fix_dev + h->fix_col_ofs[args.cbeg];
  args.nfix = h->fix_col_ofs[args.cend] - h->fix_col_ofs[args.cbeg];
  // stream kernels: CTA = (group of adjacent strips, x segment); one wave of
  // 2 CTAs per SM, segments no shorter than 16 columns
  const int nstrips = (h->ny + kStripCells - 1) / kStripCells;
  auto grid_for = [&](int warps, int ctas_per_sm) {
    const int ngroups = (nstrips + warps - 1) / warps;
    int nseg = std::max(1, (sm_count() * ctas_per_sm) / ngroups);
    nseg = std::min(nseg, std::max(1, ncol / 16));
    args.nseg = nseg;
    return ngroups * nseg;
  };
  if (algo == 4 || (algo == 0 && P != 1)) {
    const int grid = grid_for(MarchCfg<P>::WARPS, MarchCfg<P>::MINB);
    int rc = launch_march<P>(h, tab, args, grid, st);
    if (rc) return rc;
  } else if (P == 1 && (algo == 0 || algo == 3)) {
    const int grid = grid_for(kTmaWarps, 2);
    if constexpr (P == 1) {
      const size_t smem = (size_t)kTmaWarps * kTmaWarpSmem;
      static bool attr_set = false;
      if (!attr_set) {
        STEFAN_CUDA(cudaFuncSetAttribute(ipdg_apply_p1_tma, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                         (int)smem));
        attr_set = true;
      }
      ipdg_apply_p1_tma<<<grid, kTmaWarps * 32, smem, st>>>(tab, args, h->runs_dev, h->run_ofs_dev);
    }
  } else {
    const int grid = grid_for(C::WARPS, C::MINB);
    const size_t smem = (size_t)C::WARPS * C::NST * (32 * C::STRIDE + 16);
    static bool attr_set = false;
    if (!attr_set) {
      STEFAN_CUDA(cudaFuncSetAttribute(ipdg_apply_stream<P>,
                                       cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
      attr_set = true;
    }
    ipdg_apply_stream<P><<<grid, C::WARPS * 32, smem, st>>>(tab, args);
  }
  STEFAN_CUDA(cudaGetLastError());
  return ST_OK;
}

static int dispatch_apply(const IpdgHandle* h, const ApplyArgs& a, int algo, cudaStream_t st) {
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
  h->flags_dev = nullptr;
  h->fix_dev = nullptr;
  h->nfix = 0;
  h->runs_dev = nullptr;
  h->xh_dev = h->yh_dev = nullptr;
  h->host_ready = false;
  h->run_ofs_dev = nullptr;
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
      ph[(size_t)(i + 1) * (p->ny + 2) + (j + 1)] = (s == 1) ? 1 : 2;
    }
  for (int j = 0; j < p->ny; ++j) {
    if (phase_lo) ph[(size_t)0 * (p->ny + 2) + (j + 1)] = (phase_lo[j] == 1) ? 1 : 2;
    if (phase_hi) ph[(size_t)(p->nx + 1) * (p->ny + 2) + (j + 1)] = (phase_hi[j] == 1) ? 1 : 2;
  }
  // strip classification for the stream kernel: flag = the phase code shared by
  // the (up to) 30 cells of the strip in that column, or 0 if they differ; and the
  // list of cells the fix-up kernel recomputes
  const int nstrips = (p->ny + kStripCells - 1) / kStripCells;
  const int nxpad = (p->nx + 15) / 16 * 16;
  std::vector<uint8_t> flags((size_t)nstrips * nxpad, 0);
  std::vector<int32_t> fix;
  if ((int64_t)p->nx * p->ny >= (int64_t)INT32_MAX) {
    delete h;
    return fail(ST_UNSUPPORTED, "stefan_ipdg_create: more than 2^31 cells per slab");
  }
  {
    const int ld = p->ny + 2;
    for (int i = 0; i < p->nx; ++i) {
      const int8_t* col = ph.data() + (size_t)(i + 1) * ld + 1;
      for (int sidx = 0; sidx < nstrips; ++sidx) {
        const int ja = sidx * kStripCells, jb = std::min(ja + kStripCells, (int)p->ny);
        const int8_t ref = col[ja];
        // the loop produces the strip in this column iff all its cells, and all their
        // x-neighbours, have one phase; the bottom / top cell additionally need the
        // cell below / above them to have it (else: skip bit, and the cell is listed)
        bool same = true;
        for (int j = ja; j < jb && same; ++j)
          same = col[j] == ref && col[j - ld] == ref && col[j + ld] == ref;
        uint8_t flag = 0;
        if (same) {
          flag = (uint8_t)ref;
          if (col[ja - 1] != ref) flag |= 4;
          if (col[jb] != ref) flag |= 8;
        }
        flags[(size_t)sidx * nxpad + i] = flag;
        for (int j = ja; j < jb; ++j)
          if (!same || ((flag & 4) && j == ja) || ((flag & 8) && j == jb - 1))
            fix.push_back((int32_t)((int64_t)i * p->ny + j));
      }
    }
  }
  h->nfix = (int)fix.size();
  h->fix_col_ofs.assign(p->nx + 1, 0);
  for (int32_t cell : fix) h->fix_col_ofs[cell / p->ny + 1]++;
  for (int i = 0; i < p->nx; ++i) h->fix_col_ofs[i + 1] += h->fix_col_ofs[i];
  // run lists: maximal column ranges over which a strip keeps the same flag, as
  // (end column, flag); each list ends with a sentinel that is never reached
  std::vector<int2> runs;
  std::vector<int> run_ofs(nstrips);
  for (int sidx = 0; sidx < nstrips; ++sidx) {
    run_ofs[sidx] = (int)runs.size();
    const uint8_t* f = flags.data() + (size_t)sidx * nxpad;
    for (int i = 0; i < p->nx;) {
      int e = i + 1;
      while (e < p->nx && f[e] == f[i]) ++e;
      runs.push_back(make_int2(e, (int)f[i]));
      i = e;
    }
    runs.push_back(make_int2(INT32_MAX, 0));
    runs.push_back(make_int2(INT32_MAX, 0));
  }
  cudaError_t e = cudaMalloc(&h->phase_dev, pn);
  if (e == cudaSuccess) e = cudaMemcpy(h->phase_dev, ph.data(), pn, cudaMemcpyHostToDevice);
  if (e == cudaSuccess) e = cudaMalloc(&h->flags_dev, flags.size());
  if (e == cudaSuccess) e = cudaMemcpy(h->flags_dev, flags.data(), flags.size(), cudaMemcpyHostToDevice);
  if (e == cudaSuccess) e = cudaMalloc(&h->runs_dev, runs.size() * sizeof(int2));
  if (e == cudaSuccess)
    e = cudaMemcpy(h->runs_dev, runs.data(), runs.size() * sizeof(int2), cudaMemcpyHostToDevice);
  if (e == cudaSuccess) e = cudaMalloc(&h->run_ofs_dev, run_ofs.size() * sizeof(int));
  if (e == cudaSuccess)
    e = cudaMemcpy(h->run_ofs_dev, run_ofs.data(), run_ofs.size() * sizeof(int), cudaMemcpyHostToDevice);
  if (e == cudaSuccess && h->nfix > 0) e = cudaMalloc(&h->fix_dev, fix.size() * sizeof(int32_t));
  if (e == cudaSuccess && h->nfix > 0)
    e = cudaMemcpy(h->fix_dev, fix.data(), fix.size() * sizeof(int32_t), cudaMemcpyHostToDevice);
  if (e != cudaSuccess) { delete h; return fail(ST_CUDA_ERROR, "stefan_ipdg_create: %s", cudaGetErrorString(e)); }
  *out = reinterpret_cast<stefan_ipdg*>(h);
  return ST_OK;
}

int stefan_ipdg_destroy(stefan_ipdg* hv) {
  IpdgHandle* h = reinterpret_cast<IpdgHandle*>(hv);
  if (!h) return ST_OK;
  if (h->phase_dev) cudaFree(h->phase_dev);
  if (h->flags_dev) cudaFree(h->flags_dev);
  if (h->fix_dev) cudaFree(h->fix_dev);
  if (h->runs_dev) cudaFree(h->runs_dev);
  if (h->xh_dev) cudaFree(h->xh_dev);
  if (h->yh_dev) cudaFree(h->yh_dev);
  if (h->host_ready) {
    for (auto& s : h->hs) cudaStreamDestroy(s);
    for (auto& e : h->hev) cudaEventDestroy(e);
  }
  if (h->run_ofs_dev) cudaFree(h->run_ofs_dev);
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
  STEFAN_REQUIRE(algo >= 0 && algo <= 4 && !(algo == 3 && h->order != 1), "stefan_ipdg_apply: unknown algo %d", algo);
  ApplyArgs a{x, x_lo, x_hi, y, h->phase_dev, h->flags_dev, (h->nx + 15) / 16 * 16, h->nx, h->ny, 1, 0, h->nx, nullptr, 0};
  return dispatch_apply(h, a, algo, static_cast<cudaStream_t>(stream));
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
  STEFAN_REQUIRE(algo >= 0 && algo <= 4 && !(algo == 3 && h->order != 1), "stefan_ipdg_apply_columns: unknown algo %d", algo);
  const bool needs_lo = col_begin == 0, needs_hi = col_end == h->nx;
  STEFAN_REQUIRE(!(needs_lo && h->has_lo && !x_lo) && !(needs_hi && h->has_hi && !x_hi),
                 "stefan_ipdg_apply_columns: the halo column next to the requested range is missing");
  ApplyArgs a{x, h->has_lo ? x_lo : nullptr, h->has_hi ? x_hi : nullptr, y, h->phase_dev, h->flags_dev,
              (h->nx + 15) / 16 * 16, h->nx, h->ny, 1, col_begin, col_end, nullptr, 0};
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
  if (nslabs <= 0) nslabs = 16;
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
      ApplyArgs a{h->xh_dev, nullptr, nullptr, h->yh_dev, h->phase_dev, h->flags_dev, (h->nx + 15) / 16 * 16,
                  h->nx, h->ny, 1, slab_begin(k), slab_begin(k + 1), nullptr, 0};
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
