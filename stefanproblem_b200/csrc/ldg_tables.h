// Local DG (3 dofs per node: T, q_x, q_y) on an uncut uniform Cartesian mesh: tables
// and the per-cell row function the matrix-free kernels apply.
//
// Reference passes (StefanDG/LocalDG/), all Kronecker products on a uniform mesh:
//   divergence_operator   LDG_div_operator.jl:1-11      rows q_d, cols T:  int d_d(phi_a) phi_b
//   mass_matrix           LDG_mass_operator.jl:1-25     rows q_d, cols q_d: int phi_a phi_b
//   gradient_operator     LDG_gradient_operator.jl:1-11 rows T, cols q_d:  k int d_d(phi_a) phi_b
//   scalar flux           LDG_scalar_flux_operator.jl:25-107   rows q_d of the cell itself:
//        n_d e [(-1/2 - bp) tr T_K + (-1/2 + bp) tr T_N],  bp = 1/2 sign(V0 . n)
//   vector flux           LDG_vector_flux_operator.jl:47-154   rows T of the cell itself:
//        n_d e k [(-1/2 + bn) tr q_d,K + (-1/2 - bn) tr q_d,N] + alpha e (tr T_K - tr T_N),
//        bn = n . V0   (the raw unit vector: LDG_assembly.jl:44-53 passes V0 as `beta`)
//   boundary vector flux  LDG_vector_flux_operator.jl:391-439
//        e [-k n_d tr q_d,K + alpha_b tr T_K],  alpha_b = (V0 . n < 0) ? negpen : pospen
// Every cell assembles its own rows for every face with a same-phase neighbour
// (LDG_scalar_flux_operator.jl:159); faces between unlike phases get nothing -- unless
// LdgPhys::aligned_interface asks for the face-level interface functions there (T_IFACE entries).
// With M = sum_q w N N', C = sum_q w N' N' (1-D, reference cell):
//   y_qx = jy (I (x) M) [ jx (M (x) I) q_x + (Cx (x) I) T ]      Cx = C + face terms along x
//   y_qy = jx (M (x) I) [ jy (I (x) M) q_y + (I (x) Cy) T ]
//   y_T  = jy (I (x) M) Gx(q_x, T) + jx (M (x) I) Gy(q_y, T),   G = k C q + face terms
#pragma once
#include "ipdg_tables.h"

namespace stefan {

struct LdgPhys {
  double k1, k2;
  double interiorpenalty, negboundarypenalty, posboundarypenalty;
  double V0x, V0y;
  // mesh-aligned interface (extension): the face-level scalar / vector flux functions also on
  // the faces between uncut cells of unlike phase, with alpha = interfacepenalty and the neighbour's conductivity
  // on the neighbour's flux (LDG_scalar_flux_operator.jl:241-283, LDG_vector_flux_operator.jl:318-355)
  double interfacepenalty = 0.0;
  int aligned_interface = 0;
};

template <int P>
struct LdgTab {
  static constexpr int N1 = P + 1;
  double Mx[N1 * N1], My[N1 * N1];  // jx Mh, jy Mh
  double C[N1 * N1];                // sum_q w N'_a N_b
  double kk[2];                     // conductivity per phase index
  // face coefficients [axis][side 0 = low / 1 = high][face type]
  double sqK[2][2][T_NTYPES], sqN[2][2][T_NTYPES];      // q-rows <- tr T (own / neighbour)
  double tqK[2][2][2][T_NTYPES], tqN[2][2][2][T_NTYPES];  // [phase] T-rows <- tr q_d
  double tpK[2][2][T_NTYPES], tpN[2][2][T_NTYPES];      // T-rows <- tr T (penalty)
};

inline double sign0(double v) { return v > 0 ? 1.0 : (v < 0 ? -1.0 : 0.0); }

template <int P>
int build_ldg_tables(const RefElem1D& r, double jx, double jy, const LdgPhys& ph, LdgTab<P>* t) {
  constexpr int N1 = P + 1;
  if (r.order != P) return 1;
  std::memset(t, 0, sizeof(*t));
  for (int a = 0; a < N1; ++a)
    if (r.NL[a] != (a == 0 ? 1.0 : 0.0) || r.NR[a] != (a == P ? 1.0 : 0.0)) return 2;
  for (int a = 0; a < N1; ++a)
    for (int b = 0; b < N1; ++b) {
      t->Mx[a * N1 + b] = jx * r.M[a][b];
      t->My[a * N1 + b] = jy * r.M[a][b];
      t->C[a * N1 + b] = r.C[a][b];
    }
  t->kk[0] = ph.k1;
  t->kk[1] = ph.k2;
  const double V0[2] = {ph.V0x, ph.V0y};
  for (int ax = 0; ax < 2; ++ax)
    for (int side = 0; side < 2; ++side) {
      const double n = side == 0 ? -1.0 : 1.0;
      const double v0n = V0[ax] * n;
      const double bp = 0.5 * sign0(v0n);  // edge_switch, LDG_utilities.jl:1-4
      const double bn = v0n;               // posnormals' * beta with beta = V0
      const double alpha_b = v0n < 0 ? ph.negboundarypenalty : ph.posboundarypenalty;
      t->sqK[ax][side][T_SAME] = n * (-0.5 - bp);
      t->sqN[ax][side][T_SAME] = n * (-0.5 + bp);
      t->tpK[ax][side][T_SAME] = ph.interiorpenalty;
      t->tpN[ax][side][T_SAME] = -ph.interiorpenalty;
      t->tpK[ax][side][T_BDRY] = alpha_b;
      if (ph.aligned_interface) {
        t->sqK[ax][side][T_IFACE] = n * (-0.5 - bp);
        t->sqN[ax][side][T_IFACE] = n * (-0.5 + bp);
        t->tpK[ax][side][T_IFACE] = ph.interfacepenalty;
        t->tpN[ax][side][T_IFACE] = -ph.interfacepenalty;
      }
      for (int s = 0; s < 2; ++s) {
        const double k = t->kk[s];
        t->tqK[s][ax][side][T_SAME] = n * k * (-0.5 + bn);
        t->tqN[s][ax][side][T_SAME] = n * k * (-0.5 - bn);
        t->tqK[s][ax][side][T_BDRY] = -k * n;
        if (ph.aligned_interface) {
          t->tqK[s][ax][side][T_IFACE] = n * k * (-0.5 + bn);
          t->tqN[s][ax][side][T_IFACE] = n * t->kk[1 - s] * (-0.5 - bn);   // the NEIGHBOUR's conductivity
        }
      }
    }
  return 0;
}

// Rows of one cell.  U* hold the 3 NF dofs of a cell interleaved per node (T, q_x,
// q_y); of the neighbours only the face nodes' T and normal q component are read.
template <int P>
#if defined(__CUDACC__)
__host__ __device__ __forceinline__
#else
inline
#endif
void ldg_cell_rows(const LdgTab<P>& tab, int s, int tL, int tR, int tD, int tU,
                   const double* __restrict__ U, const double* __restrict__ UL,
                   const double* __restrict__ UR, const double* __restrict__ UD,
                   const double* __restrict__ UU, double* __restrict__ out) {
  constexpr int N1 = P + 1;
  constexpr int NF = N1 * N1;
  const double k = tab.kk[s];
  double aqx[NF], aqy[NF], gx[NF], gy[NF];
#define T_(W, ix, iy) W[3 * ((ix)*N1 + (iy)) + 0]
#define QX_(W, ix, iy) W[3 * ((ix)*N1 + (iy)) + 1]
#define QY_(W, ix, iy) W[3 * ((ix)*N1 + (iy)) + 2]
#pragma unroll
  for (int ix = 0; ix < N1; ++ix) {
#pragma unroll
    for (int iy = 0; iy < N1; ++iy) {
      double cqx = 0.0, g_x = 0.0, mqx = 0.0, cqy = 0.0, g_y = 0.0, mqy = 0.0;
#pragma unroll
      for (int b = 0; b < N1; ++b) {
        cqx += tab.C[ix * N1 + b] * T_(U, b, iy);
        g_x += tab.C[ix * N1 + b] * QX_(U, b, iy);
        mqx += tab.Mx[ix * N1 + b] * QX_(U, b, iy);
        cqy += tab.C[iy * N1 + b] * T_(U, ix, b);
        g_y += tab.C[iy * N1 + b] * QY_(U, ix, b);
        mqy += tab.My[iy * N1 + b] * QY_(U, ix, b);
      }
      g_x *= k;
      g_y *= k;
      if (ix == 0) {
        cqx += tab.sqK[0][0][tL] * T_(U, 0, iy) + tab.sqN[0][0][tL] * T_(UL, P, iy);
        g_x += tab.tqK[s][0][0][tL] * QX_(U, 0, iy) + tab.tqN[s][0][0][tL] * QX_(UL, P, iy) +
               tab.tpK[0][0][tL] * T_(U, 0, iy) + tab.tpN[0][0][tL] * T_(UL, P, iy);
      }
      if (ix == P) {
        cqx += tab.sqK[0][1][tR] * T_(U, P, iy) + tab.sqN[0][1][tR] * T_(UR, 0, iy);
        g_x += tab.tqK[s][0][1][tR] * QX_(U, P, iy) + tab.tqN[s][0][1][tR] * QX_(UR, 0, iy) +
               tab.tpK[0][1][tR] * T_(U, P, iy) + tab.tpN[0][1][tR] * T_(UR, 0, iy);
      }
      if (iy == 0) {
        cqy += tab.sqK[1][0][tD] * T_(U, ix, 0) + tab.sqN[1][0][tD] * T_(UD, ix, P);
        g_y += tab.tqK[s][1][0][tD] * QY_(U, ix, 0) + tab.tqN[s][1][0][tD] * QY_(UD, ix, P) +
               tab.tpK[1][0][tD] * T_(U, ix, 0) + tab.tpN[1][0][tD] * T_(UD, ix, P);
      }
      if (iy == P) {
        cqy += tab.sqK[1][1][tU] * T_(U, ix, P) + tab.sqN[1][1][tU] * T_(UU, ix, 0);
        g_y += tab.tqK[s][1][1][tU] * QY_(U, ix, P) + tab.tqN[s][1][1][tU] * QY_(UU, ix, 0) +
               tab.tpK[1][1][tU] * T_(U, ix, P) + tab.tpN[1][1][tU] * T_(UU, ix, 0);
      }
      aqx[ix * N1 + iy] = cqx + mqx;
      aqy[ix * N1 + iy] = cqy + mqy;
      gx[ix * N1 + iy] = g_x;
      gy[ix * N1 + iy] = g_y;
    }
  }
#pragma unroll
  for (int ix = 0; ix < N1; ++ix) {
#pragma unroll
    for (int iy = 0; iy < N1; ++iy) {
      double oT = 0.0, oqx = 0.0, oqy = 0.0;
#pragma unroll
      for (int c = 0; c < N1; ++c) {
        oqx += tab.My[iy * N1 + c] * aqx[ix * N1 + c];
        oT += tab.My[iy * N1 + c] * gx[ix * N1 + c];
        oqy += tab.Mx[ix * N1 + c] * aqy[c * N1 + iy];
        oT += tab.Mx[ix * N1 + c] * gy[c * N1 + iy];
      }
      T_(out, ix, iy) = oT;
      QX_(out, ix, iy) = oqx;
      QY_(out, ix, iy) = oqy;
    }
  }
#undef T_
#undef QX_
#undef QY_
}

// The same rows (phase index S, face types tL .. tU) in the form the march kernel uses: the neighbours enter as face traces (T and the normal flux
// component at the face nodes), and the sum-factorisation intermediates are never stored --
// every x-line (fixed iy) / y-line (fixed ix) result is scattered straight into the output
// accumulators through the 1-D mass matrix.  ROLLED keeps the two outer loop levels as loops:
// the kernels' tail runs this code ONCE per warp, and straight-line code that is executed once
// is one instruction-cache miss after the other (measured: 8-10 us per tail chunk at orders 2, 3).  Live data: the cell (3 NF), the output (3 NF) and
// 2 (P+1) line temporaries, instead of 4 NF intermediates on top (order 3: 110 instead of
// 160 doubles, which is what fits the register file).
template <int P, bool ROLLED = false>
#if defined(__CUDACC__)
__host__ __device__ __forceinline__
#else
inline
#endif
void ldg_rows_streamed_general(const LdgTab<P>& tab, const int S, const int tL, const int tR, const int tD,
                               const int tU, const double* __restrict__ U, const double* __restrict__ tLT,
                       const double* __restrict__ tLQ, const double* __restrict__ tRT,
                       const double* __restrict__ tRQ, const double* __restrict__ tDT,
                       const double* __restrict__ tDQ, const double* __restrict__ tUT,
                       const double* __restrict__ tUQ, double* __restrict__ out) {
  constexpr int N1 = P + 1;
  constexpr int NF = N1 * N1;
  const double k = tab.kk[S];
#pragma unroll(ROLLED ? 1 : 3 * NF)
  for (int a = 0; a < 3 * NF; ++a) out[a] = 0.0;
  // x-lines
#pragma unroll(ROLLED ? 1 : N1)
  for (int iy = 0; iy < N1; ++iy) {
#pragma unroll(ROLLED ? 1 : N1)
    for (int ix = 0; ix < N1; ++ix) {
      double cqx = 0.0, g_x = 0.0, mqx = 0.0;
#pragma unroll
      for (int b = 0; b < N1; ++b) {
        const double t = U[3 * (b * N1 + iy) + 0], q = U[3 * (b * N1 + iy) + 1];
        cqx += tab.C[ix * N1 + b] * t;
        g_x += tab.C[ix * N1 + b] * q;
        mqx += tab.Mx[ix * N1 + b] * q;
      }
      g_x *= k;
      if (ix == 0) {
        const double t = U[3 * (0 * N1 + iy) + 0], q = U[3 * (0 * N1 + iy) + 1];
        cqx += tab.sqK[0][0][tL] * t + tab.sqN[0][0][tL] * tLT[iy];
        g_x += tab.tqK[S][0][0][tL] * q + tab.tqN[S][0][0][tL] * tLQ[iy] + tab.tpK[0][0][tL] * t +
               tab.tpN[0][0][tL] * tLT[iy];
      }
      if (ix == P) {
        const double t = U[3 * (P * N1 + iy) + 0], q = U[3 * (P * N1 + iy) + 1];
        cqx += tab.sqK[0][1][tR] * t + tab.sqN[0][1][tR] * tRT[iy];
        g_x += tab.tqK[S][0][1][tR] * q + tab.tqN[S][0][1][tR] * tRQ[iy] + tab.tpK[0][1][tR] * t +
               tab.tpN[0][1][tR] * tRT[iy];
      }
      const double a = cqx + mqx;
#pragma unroll
      for (int jy = 0; jy < N1; ++jy) {  // y-mass: rows (ix, jy) get My[jy][iy] times the line value
        out[3 * (ix * N1 + jy) + 1] += tab.My[jy * N1 + iy] * a;
        out[3 * (ix * N1 + jy) + 0] += tab.My[jy * N1 + iy] * g_x;
      }
    }
  }
  // y-lines
#pragma unroll(ROLLED ? 1 : N1)
  for (int ix = 0; ix < N1; ++ix) {
#pragma unroll(ROLLED ? 1 : N1)
    for (int iy = 0; iy < N1; ++iy) {
      double cqy = 0.0, g_y = 0.0, mqy = 0.0;
#pragma unroll
      for (int c = 0; c < N1; ++c) {
        const double t = U[3 * (ix * N1 + c) + 0], q = U[3 * (ix * N1 + c) + 2];
        cqy += tab.C[iy * N1 + c] * t;
        g_y += tab.C[iy * N1 + c] * q;
        mqy += tab.My[iy * N1 + c] * q;
      }
      g_y *= k;
      if (iy == 0) {
        const double t = U[3 * (ix * N1 + 0) + 0], q = U[3 * (ix * N1 + 0) + 2];
        cqy += tab.sqK[1][0][tD] * t + tab.sqN[1][0][tD] * tDT[ix];
        g_y += tab.tqK[S][1][0][tD] * q + tab.tqN[S][1][0][tD] * tDQ[ix] + tab.tpK[1][0][tD] * t +
               tab.tpN[1][0][tD] * tDT[ix];
      }
      if (iy == P) {
        const double t = U[3 * (ix * N1 + P) + 0], q = U[3 * (ix * N1 + P) + 2];
        cqy += tab.sqK[1][1][tU] * t + tab.sqN[1][1][tU] * tUT[ix];
        g_y += tab.tqK[S][1][1][tU] * q + tab.tqN[S][1][1][tU] * tUQ[ix] + tab.tpK[1][1][tU] * t +
               tab.tpN[1][1][tU] * tUT[ix];
      }
      const double a = cqy + mqy;
#pragma unroll
      for (int jx = 0; jx < N1; ++jx) {  // x-mass: rows (jx, iy) get Mx[jx][ix] times the line value
        out[3 * (jx * N1 + iy) + 2] += tab.Mx[jx * N1 + ix] * a;
        out[3 * (jx * N1 + iy) + 0] += tab.Mx[jx * N1 + ix] * g_y;
      }
    }
  }
}

// compile-time phase, four same-phase neighbours: what the march loop calls
template <int P, int S>
#if defined(__CUDACC__)
__host__ __device__ __forceinline__
#else
inline
#endif
void ldg_rows_streamed(const LdgTab<P>& tab, const double* __restrict__ U, const double* __restrict__ tLT,
                       const double* __restrict__ tLQ, const double* __restrict__ tRT,
                       const double* __restrict__ tRQ, const double* __restrict__ tDT,
                       const double* __restrict__ tDQ, const double* __restrict__ tUT,
                       const double* __restrict__ tUQ, double* __restrict__ out) {
  ldg_rows_streamed_general<P>(tab, S, T_SAME, T_SAME, T_SAME, T_SAME, U, tLT, tLQ, tRT, tRQ, tDT, tDQ, tUT, tUQ, out);
}

// The all-interior rows once more, ordered so that the OUTPUT leaves progressively (order 3 of the
// march kernel: the streamed form above keeps the cell, the whole output and the traces live --
// 48 + 48 + 32 doubles, which does not fit 255 registers and spilled 292 bytes per thread).
// First every y-line (fixed ix) is contracted: that consumes q_y and the down / up traces and
// leaves 2 NF line values.  Then, for one jx after the other, the N1 x-lines' values at ix = jx
// are formed and the 3 N1 rows of the nodes (jx, 0 .. P) are complete: `emit(jx, rows)` stores
// them and their registers are free again.  traceR(T, Q) fetches the right neighbour's face
// values only when jx reaches P.  Live: T and q_x of the cell (2 NF), the y-line values (2 NF),
// the left traces, 2 N1 line values and 3 N1 rows -- about 90 doubles at order 3.
template <int P, int S, class TraceDU, class TraceR, class Emit>
#if defined(__CUDACC__)
__host__ __device__ __forceinline__
#else
inline
#endif
void ldg_rows_progressive(const LdgTab<P>& tab, const double* __restrict__ U, const double* __restrict__ tLT,
                          const double* __restrict__ tLQ, TraceDU&& traceDU, TraceR&& traceR, Emit&& emit) {
  constexpr int N1 = P + 1;
  constexpr int NF = N1 * N1;
  constexpr int tS = T_SAME;
  const double k = tab.kk[S];
  double ay[NF], gy[NF];
#pragma unroll
  for (int ix = 0; ix < N1; ++ix) {
    double dT, dQ, uT, uQ;  // the y-neighbours' T and q_y at the face nodes of this line
    traceDU(ix, dT, dQ, uT, uQ);
#pragma unroll
    for (int iy = 0; iy < N1; ++iy) {
      double cqy = 0.0, g_y = 0.0, mqy = 0.0;
#pragma unroll
      for (int c = 0; c < N1; ++c) {
        const double t = U[3 * (ix * N1 + c) + 0], q = U[3 * (ix * N1 + c) + 2];
        cqy += tab.C[iy * N1 + c] * t;
        g_y += tab.C[iy * N1 + c] * q;
        mqy += tab.My[iy * N1 + c] * q;
      }
      g_y *= k;
      if (iy == 0) {
        const double t = U[3 * (ix * N1 + 0) + 0], q = U[3 * (ix * N1 + 0) + 2];
        cqy += tab.sqK[1][0][tS] * t + tab.sqN[1][0][tS] * dT;
        g_y += tab.tqK[S][1][0][tS] * q + tab.tqN[S][1][0][tS] * dQ + tab.tpK[1][0][tS] * t +
               tab.tpN[1][0][tS] * dT;
      }
      if (iy == P) {
        const double t = U[3 * (ix * N1 + P) + 0], q = U[3 * (ix * N1 + P) + 2];
        cqy += tab.sqK[1][1][tS] * t + tab.sqN[1][1][tS] * uT;
        g_y += tab.tqK[S][1][1][tS] * q + tab.tqN[S][1][1][tS] * uQ + tab.tpK[1][1][tS] * t +
               tab.tpN[1][1][tS] * uT;
      }
      ay[ix * N1 + iy] = cqy + mqy;
      gy[ix * N1 + iy] = g_y;
    }
  }
#ifndef LDG_PROG_UNROLL
#define LDG_PROG_UNROLL 1
#endif
  constexpr int kUnrollX = LDG_PROG_UNROLL;
#pragma unroll(kUnrollX)
  for (int jx = 0; jx < N1; ++jx) {
    double ax[N1], gx[N1];
    double tRT[N1], tRQ[N1];
    if (jx == P) traceR(tRT, tRQ);
#pragma unroll
    for (int iy = 0; iy < N1; ++iy) {
      double cqx = 0.0, g_x = 0.0, mqx = 0.0;
#pragma unroll
      for (int b = 0; b < N1; ++b) {
        const double t = U[3 * (b * N1 + iy) + 0], q = U[3 * (b * N1 + iy) + 1];
        cqx += tab.C[jx * N1 + b] * t;
        g_x += tab.C[jx * N1 + b] * q;
        mqx += tab.Mx[jx * N1 + b] * q;
      }
      g_x *= k;
      if (jx == 0) {
        const double t = U[3 * (0 * N1 + iy) + 0], q = U[3 * (0 * N1 + iy) + 1];
        cqx += tab.sqK[0][0][tS] * t + tab.sqN[0][0][tS] * tLT[iy];
        g_x += tab.tqK[S][0][0][tS] * q + tab.tqN[S][0][0][tS] * tLQ[iy] + tab.tpK[0][0][tS] * t +
               tab.tpN[0][0][tS] * tLT[iy];
      }
      if (jx == P) {
        const double t = U[3 * (P * N1 + iy) + 0], q = U[3 * (P * N1 + iy) + 1];
        cqx += tab.sqK[0][1][tS] * t + tab.sqN[0][1][tS] * tRT[iy];
        g_x += tab.tqK[S][0][1][tS] * q + tab.tqN[S][0][1][tS] * tRQ[iy] + tab.tpK[0][1][tS] * t +
               tab.tpN[0][1][tS] * tRT[iy];
      }
      ax[iy] = cqx + mqx;
      gx[iy] = g_x;
    }
    double rows[3 * N1];
#pragma unroll
    for (int iy = 0; iy < N1; ++iy) {
      double oT = 0.0, oqx = 0.0, oqy = 0.0;
#pragma unroll
      for (int c = 0; c < N1; ++c) {
        oqx += tab.My[iy * N1 + c] * ax[c];
        oT += tab.My[iy * N1 + c] * gx[c];
        oqy += tab.Mx[jx * N1 + c] * ay[c * N1 + iy];
        oT += tab.Mx[jx * N1 + c] * gy[c * N1 + iy];
      }
      rows[3 * iy + 0] = oT;
      rows[3 * iy + 1] = oqx;
      rows[3 * iy + 2] = oqy;
    }
    emit(jx, rows);
  }
}

}  // namespace stefan
