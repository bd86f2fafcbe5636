// Interior-penalty DG on an uncut uniform Cartesian mesh: the 1-D line-operator
// tables the matrix-free kernels apply, built on the host from the reference's
// element-level formulas.
//
// For a cell K of phase s (conductivity k), axis d (jacobian j, tangential
// jacobian jt) and the low / high faces along that axis, the rows of the global
// matrix that belong to K are
//
//   y_K = jy (Ax (x) Mh) [x_L, x_K, x_R]  +  jx (Mh (x) Ay) [x_D, x_K, x_U]
//
// where Mh = sum_q w N N' (1-D mass on the reference cell) and A_d is the 1-D
// interior-penalty operator with
//   volume   (k/j) Kh                         IP_gradient_operator.jl:1-9
//   flux     -{k dn u}[v] - {k dn v}[u]       IP_flux_operator.jl:35-118
//   penalty  pen [u][v]                       IP_penalty_operator.jl:21-65
//   Robin    lambda {u}{v}  (interface only)  IP_robin_interface_operator.jl:1-45
//   boundary -(M + M') + pen P                IP_flux_operator.jl:330-357,
//                                             IP_penalty_operator.jl:239-254
// (legacy variant: -M' + pen P).  Writing e for the trace vector
// N(+-1) of K at the face, g for its outward normal derivative N'(+-1) n / j, and
// eN, gN for the neighbour's trace / derivative along K's outward normal:
//   diag  += e (cK g' + cpK e') + g cgK e'
//   off   += e (cN gN' + cpN eN') + g cgN eN'
// with the six coefficients per face type below.  Because the Lagrange nodes
// include the end points, e is a unit vector, so the off-diagonal 1-D blocks have
// one dense row (the face node) plus one column (the neighbour's face node).
#pragma once
#include "refelem.h"

namespace stefan {

enum FaceType : int { T_SAME = 0, T_IFACE = 1, T_BDRY = 2, T_NTYPES = 3 };
// which of the reference's assemble_* passes are present in the operator
enum IpTerms : int {
  IP_VOLUME = 1,               // assemble_gradient_operator!
  IP_INTERELEMENT_FLUX = 2,    // assemble_interelement_flux_operator!
  IP_INTERELEMENT_PENALTY = 4, // assemble_interelement_penalty_operator!
  IP_INTERFACE_FLUX = 8,       // face-level flux on unlike-phase faces (mesh-aligned interface)
  IP_INTERFACE_PENALTY = 16,   // face-level penalty on unlike-phase faces
  IP_BOUNDARY_FLUX = 32,       // assemble_boundary_flux_operator!
  IP_BOUNDARY_PENALTY = 64,    // assemble_boundary_penalty_operator!
  IP_ROBIN = 128,              // assemble_robin_operator! (face-level, unlike-phase faces)
  IP_ALL_TERMS = 255,
};
enum IpVariant : int { IP_CURRENT = 0, IP_LEGACY_BOUNDARY = 1 };

template <int P>
struct IpTab {
  static constexpr int N1 = P + 1;
  // [axis][phase index (0: sign +1 / k1, 1: sign -1 / k2)][type lo][type hi]
  double D[2][2][T_NTYPES][T_NTYPES][N1 * N1];
  // off-diagonal blocks towards the low (L*) and high (R*) neighbour:
  // L0 = row of the low-face node, Lc = column of the neighbour's face node (rows >= 1)
  double L0[2][2][T_NTYPES][N1];
  double Lc[2][2][T_NTYPES][N1];
  double Rp[2][2][T_NTYPES][N1];
  double Rc[2][2][T_NTYPES][N1];
  double Mx[N1 * N1];  // jx * Mh
  double My[N1 * N1];  // jy * Mh
};

struct IpPhys {
  double k1, k2, penalty, lambda;
  int variant;
  int terms;
};

struct FaceCoef {
  double cK, cN, cpK, cpN, cgK, cgN;
};

inline FaceCoef ip_face_coef(int type, double k, double kother, const IpPhys& ph) {
  FaceCoef c = {0, 0, 0, 0, 0, 0};
  const bool fl = type == T_SAME ? (ph.terms & IP_INTERELEMENT_FLUX)
                  : type == T_IFACE ? (ph.terms & IP_INTERFACE_FLUX) : (ph.terms & IP_BOUNDARY_FLUX);
  const bool pe = type == T_SAME ? (ph.terms & IP_INTERELEMENT_PENALTY)
                  : type == T_IFACE ? (ph.terms & IP_INTERFACE_PENALTY) : (ph.terms & IP_BOUNDARY_PENALTY);
  if (type == T_BDRY) {
    if (fl) {
      c.cK = -k;
      c.cgK = (ph.variant == IP_LEGACY_BOUNDARY) ? 0.0 : -k;
    }
    if (pe) c.cpK = ph.penalty;
  } else {
    if (fl) {
      c.cK = -0.5 * k;
      c.cN = -0.5 * (type == T_IFACE ? kother : k);
      c.cgK = -0.5 * k;
      c.cgN = 0.5 * k;
    }
    if (pe) {
      c.cpK = ph.penalty;
      c.cpN = -ph.penalty;
    }
    if (type == T_IFACE && (ph.terms & IP_ROBIN)) {
      c.cpK += 0.25 * ph.lambda;
      c.cpN += 0.25 * ph.lambda;
    }
  }
  return c;
}

template <int P>
int build_ip_tables(const RefElem1D& r, double jx, double jy, const IpPhys& ph, IpTab<P>* t) {
  constexpr int N1 = P + 1;
  if (r.order != P) return 1;
  std::memset(t, 0, sizeof(*t));
  // the structured off-diagonal form needs N(-1) = e_0 and N(+1) = e_P
  for (int a = 0; a < N1; ++a) {
    if (r.NL[a] != (a == 0 ? 1.0 : 0.0) || r.NR[a] != (a == P ? 1.0 : 0.0)) return 2;
  }
  const double jac[2] = {jx, jy};
  for (int ax = 0; ax < 2; ++ax) {
    const double j = jac[ax];
    double gm[N1], gp[N1], gNlo[N1], gNhi[N1];
    for (int b = 0; b < N1; ++b) {
      gm[b] = -r.dNL[b] / j;    // own outward derivative at the low face
      gp[b] = +r.dNR[b] / j;    // own outward derivative at the high face
      gNlo[b] = -r.dNR[b] / j;  // low neighbour, at its high end, along my outward (-) normal
      gNhi[b] = +r.dNL[b] / j;  // high neighbour, at its low end, along my outward (+) normal
    }
    for (int s = 0; s < 2; ++s) {
      const double k = s == 0 ? ph.k1 : ph.k2;
      const double ko = s == 0 ? ph.k2 : ph.k1;
      double Dlo[T_NTYPES][N1][N1], Dhi[T_NTYPES][N1][N1];
      for (int ty = 0; ty < T_NTYPES; ++ty) {
        const FaceCoef c = ip_face_coef(ty, k, ko, ph);
        for (int a = 0; a < N1; ++a)
          for (int b = 0; b < N1; ++b) {
            Dlo[ty][a][b] = r.NL[a] * (c.cK * gm[b] + c.cpK * r.NL[b]) + gm[a] * c.cgK * r.NL[b];
            Dhi[ty][a][b] = r.NR[a] * (c.cK * gp[b] + c.cpK * r.NR[b]) + gp[a] * c.cgK * r.NR[b];
          }
        for (int b = 0; b < N1; ++b) {
          // low neighbour: its trace vector is N(+1) = e_P
          t->L0[ax][s][ty][b] = c.cN * gNlo[b] + c.cpN * r.NR[b] + gm[0] * c.cgN * r.NR[b];
          t->Rp[ax][s][ty][b] = c.cN * gNhi[b] + c.cpN * r.NL[b] + gp[P] * c.cgN * r.NL[b];
        }
        for (int a = 0; a < N1; ++a) {
          t->Lc[ax][s][ty][a] = (a == 0) ? 0.0 : gm[a] * c.cgN;
          t->Rc[ax][s][ty][a] = (a == P) ? 0.0 : gp[a] * c.cgN;
        }
      }
      for (int tl = 0; tl < T_NTYPES; ++tl)
        for (int th = 0; th < T_NTYPES; ++th)
          for (int a = 0; a < N1; ++a)
            for (int b = 0; b < N1; ++b)
              t->D[ax][s][tl][th][a * N1 + b] =
                  ((ph.terms & IP_VOLUME) ? (k / j) * r.K[a][b] : 0.0) + Dlo[tl][a][b] + Dhi[th][a][b];
    }
  }
  for (int a = 0; a < N1; ++a)
    for (int b = 0; b < N1; ++b) {
      t->Mx[a * N1 + b] = jx * r.M[a][b];
      t->My[a * N1 + b] = jy * r.M[a][b];
    }
  // The interior (same-phase neighbours on both sides) blocks are persymmetric and
  // the high-side blocks mirror the low-side ones; make that exact in floating point
  // so that kernels that exploit it agree bitwise with those that do not.
  for (int ax = 0; ax < 2; ++ax)
    for (int s = 0; s < 2; ++s) {
      double* D = t->D[ax][s][T_SAME][T_SAME];
      for (int a = 0; a < N1; ++a)
        for (int b = 0; b < N1; ++b) {
          const int m = (P - a) * N1 + (P - b);
          if (m > a * N1 + b) D[m] = D[a * N1 + b];
        }
      for (int b = 0; b < N1; ++b) {
        t->Rp[ax][s][T_SAME][b] = t->L0[ax][s][T_SAME][P - b];
        t->Rc[ax][s][T_SAME][b] = t->Lc[ax][s][T_SAME][P - b];
      }
    }
  for (int a = 0; a < N1; ++a)
    for (int b = 0; b < N1; ++b) {
      const int m = (P - a) * N1 + (P - b);
      if (m > a * N1 + b) {
        t->Mx[m] = t->Mx[a * N1 + b];
        t->My[m] = t->My[a * N1 + b];
      }
    }
  return 0;
}

// Order-1 interior cell of one phase with same-phase neighbours: the 14 distinct
// coefficients of the two 1-D stencils and mass matrices, held in registers by the
// stream kernel (the uniform datapath cannot hold two phases' worth of them).
struct P1Coef {
  double dx0, dx1, lx0, lx1, lcx, mx0, mx1;
  double dy0, dy1, ly0, ly1, lcy, my0, my1;
};

#if defined(__CUDACC__)
__host__ __device__ __forceinline__
#else
inline
#endif
void p1_load_coef(const IpTab<1>& tab, int s, P1Coef& c) {
  c.dx0 = tab.D[0][s][T_SAME][T_SAME][0];
  c.dx1 = tab.D[0][s][T_SAME][T_SAME][1];
  c.lx0 = tab.L0[0][s][T_SAME][0];
  c.lx1 = tab.L0[0][s][T_SAME][1];
  c.lcx = tab.Lc[0][s][T_SAME][1];
  c.mx0 = tab.Mx[0];
  c.mx1 = tab.Mx[1];
  c.dy0 = tab.D[1][s][T_SAME][T_SAME][0];
  c.dy1 = tab.D[1][s][T_SAME][T_SAME][1];
  c.ly0 = tab.L0[1][s][T_SAME][0];
  c.ly1 = tab.L0[1][s][T_SAME][1];
  c.lcy = tab.Lc[1][s][T_SAME][1];
  c.my0 = tab.My[0];
  c.my1 = tab.My[1];
}

#if defined(__CUDACC__)
__host__ __device__ __forceinline__
#else
inline
#endif
void p1_uniform_rows(const P1Coef& c, const double* __restrict__ X, const double* __restrict__ XL,
                     const double* __restrict__ XR, const double* __restrict__ XD,
                     const double* __restrict__ XU, double* __restrict__ out) {
  double ax[4], ay[4];
#pragma unroll
  for (int iy = 0; iy < 2; ++iy) {
    ax[0 + iy] = c.dx0 * X[0 + iy] + c.dx1 * X[2 + iy] + c.lx0 * XL[0 + iy] + c.lx1 * XL[2 + iy] +
                 c.lcx * XR[0 + iy];
    ax[2 + iy] = c.dx1 * X[0 + iy] + c.dx0 * X[2 + iy] + c.lcx * XL[2 + iy] + c.lx1 * XR[0 + iy] +
                 c.lx0 * XR[2 + iy];
  }
#pragma unroll
  for (int ix = 0; ix < 2; ++ix) {
    ay[2 * ix + 0] = c.dy0 * X[2 * ix] + c.dy1 * X[2 * ix + 1] + c.ly0 * XD[2 * ix] +
                     c.ly1 * XD[2 * ix + 1] + c.lcy * XU[2 * ix];
    ay[2 * ix + 1] = c.dy1 * X[2 * ix] + c.dy0 * X[2 * ix + 1] + c.lcy * XD[2 * ix + 1] +
                     c.ly1 * XU[2 * ix] + c.ly0 * XU[2 * ix + 1];
  }
  out[0] = c.my0 * ax[0] + c.my1 * ax[1] + c.mx0 * ay[0] + c.mx1 * ay[2];
  out[1] = c.my1 * ax[0] + c.my0 * ax[1] + c.mx0 * ay[1] + c.mx1 * ay[3];
  out[2] = c.my0 * ax[2] + c.my1 * ax[3] + c.mx1 * ay[0] + c.mx0 * ay[2];
  out[3] = c.my1 * ax[2] + c.my0 * ax[3] + c.mx1 * ay[1] + c.mx0 * ay[3];
}

// Face type seen from a cell of sign s towards a neighbour of sign sn (0 = no cell).
#if defined(__CUDACC__)
__host__ __device__
#endif
inline int ip_face_type(int s, int sn) {
  return sn == 0 ? T_BDRY : (sn == s ? T_SAME : T_IFACE);
}

// Rows of one cell.  X* are the NF = (P+1)^2 dofs (a = ix (P+1) + iy) of the cell
// and of its left / right / down / up neighbours (zeros where there is none).
// All table indices are either compile-time constants (interior fast path) or
// per-thread values (general path); the body is the same.  ROLLED keeps the two outer loop
// levels as loops (the march kernels' tail: code that runs once per warp must be small, every
// instruction line of an unrolled body is a cold instruction-cache miss).
template <int P, bool ROLLED = false>
#if defined(__CUDACC__)
__host__ __device__ __forceinline__
#else
inline
#endif
void ip_cell_rows(const IpTab<P>& tab, int s, int tL, int tR, int tD, int tU,
                  const double* __restrict__ X, const double* __restrict__ XL,
                  const double* __restrict__ XR, const double* __restrict__ XD,
                  const double* __restrict__ XU, double* __restrict__ out) {
  constexpr int N1 = P + 1;
  const double* Dx = tab.D[0][s][tL][tR];
  const double* Dy = tab.D[1][s][tD][tU];
  const double* L0x = tab.L0[0][s][tL];
  const double* Lcx = tab.Lc[0][s][tL];
  const double* Rpx = tab.Rp[0][s][tR];
  const double* Rcx = tab.Rc[0][s][tR];
  const double* L0y = tab.L0[1][s][tD];
  const double* Lcy = tab.Lc[1][s][tD];
  const double* Rpy = tab.Rp[1][s][tU];
  const double* Rcy = tab.Rc[1][s][tU];
  double ax[N1 * N1], ay[N1 * N1];
#pragma unroll(ROLLED ? 1 : N1)
  for (int ix = 0; ix < N1; ++ix) {
#pragma unroll(ROLLED ? 1 : N1)
    for (int iy = 0; iy < N1; ++iy) {
      // x-line through node (ix, iy): couples ix' in K, L, R at fixed iy
      double v = 0.0;
#pragma unroll
      for (int b = 0; b < N1; ++b) v += Dx[ix * N1 + b] * X[b * N1 + iy];
      if (ix == 0) {
#pragma unroll
        for (int b = 0; b < N1; ++b) v += L0x[b] * XL[b * N1 + iy];
      } else {
        v += Lcx[ix] * XL[P * N1 + iy];
      }
      if (ix == P) {
#pragma unroll
        for (int b = 0; b < N1; ++b) v += Rpx[b] * XR[b * N1 + iy];
      } else {
        v += Rcx[ix] * XR[0 * N1 + iy];
      }
      ax[ix * N1 + iy] = v;
      // y-line through node (ix, iy): couples iy' in K, D, U at fixed ix
      double u = 0.0;
#pragma unroll
      for (int c = 0; c < N1; ++c) u += Dy[iy * N1 + c] * X[ix * N1 + c];
      if (iy == 0) {
#pragma unroll
        for (int c = 0; c < N1; ++c) u += L0y[c] * XD[ix * N1 + c];
      } else {
        u += Lcy[iy] * XD[ix * N1 + P];
      }
      if (iy == P) {
#pragma unroll
        for (int c = 0; c < N1; ++c) u += Rpy[c] * XU[ix * N1 + c];
      } else {
        u += Rcy[iy] * XU[ix * N1 + 0];
      }
      ay[ix * N1 + iy] = u;
    }
  }
#pragma unroll(ROLLED ? 1 : N1)
  for (int ix = 0; ix < N1; ++ix) {
#pragma unroll(ROLLED ? 1 : N1)
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

}  // namespace stefan
