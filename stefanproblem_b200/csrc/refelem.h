// Reference-element math (host side): Lagrange bases, Gauss rules and the 1-D
// element matrices that the uniform-mesh operators factor into.
//
// Conventions are those of the packages the reference builds on:
// 1-D Lagrange basis on order+1 equispaced nodes in [-1,1]; the 2-D basis is
// kron(N(xi), N(eta)) (eta index fastest); Gauss-Legendre tensor rules.
//
// On an uncut uniform mesh every 2-D element / face matrix of the reference is a
// Kronecker product of these 1-D matrices, e.g. for
//   gradient_operator  (StefanDG/InteriorPenalty/IP_gradient_operator.jl:1-9)
//     sum_q k G G' detjac w = k [ (jy/jx) K (x) M  +  (jx/jy) M (x) K ]
//   flux_operator on an x-normal face (IP_flux_operator.jl:1-33), sa = jy
//     sum_q k (G n) V' sa w = k (N'(+-1)/jx) N(+-1)'  (x)  jy M
// with M = sum_q w N N', K = sum_q w N' N'', C = sum_q w N' N'.
#pragma once
#include <cmath>
#include <cstring>

namespace stefan {

constexpr int kMaxOrder = 3;
constexpr int kMaxN1 = kMaxOrder + 1;
constexpr int kMaxQ = 8;

// Gauss-Legendre nodes/weights on [-1,1] by Newton iteration on P_n.
inline void gauss_legendre(int n, double* x, double* w) {
  for (int i = 0; i < n; ++i) {
    double z = std::cos(M_PI * (i + 0.75) / (n + 0.5));
    double pp = 0.0;
    for (int it = 0; it < 100; ++it) {
      double p1 = 1.0, p2 = 0.0;
      for (int j = 0; j < n; ++j) {
        double p3 = p2;
        p2 = p1;
        p1 = ((2.0 * j + 1.0) * z * p2 - j * p3) / (j + 1.0);
      }
      pp = n * (z * p1 - p2) / (z * z - 1.0);
      double dz = p1 / pp;
      z -= dz;
      if (std::fabs(dz) < 1e-16) break;
    }
    // ascending order like numpy.polynomial.legendre.leggauss
    x[n - 1 - i] = z;
    w[n - 1 - i] = 2.0 / ((1.0 - z * z) * pp * pp);
  }
  // enforce exact symmetry
  for (int i = 0; i < n / 2; ++i) {
    double a = 0.5 * (x[n - 1 - i] - x[i]);
    x[i] = -a;
    x[n - 1 - i] = a;
    double ww = 0.5 * (w[i] + w[n - 1 - i]);
    w[i] = w[n - 1 - i] = ww;
  }
  if (n % 2) x[n / 2] = 0.0;
}

// Values N_a(x) and derivatives N'_a(x) of the equispaced Lagrange basis.
inline void lagrange_1d(int order, double x, double* val, double* der) {
  const int n = order + 1;
  double nodes[kMaxN1];
  for (int a = 0; a < n; ++a) nodes[a] = (order == 0) ? 0.0 : -1.0 + 2.0 * a / order;
  for (int a = 0; a < n; ++a) {
    double v = 1.0;
    for (int b = 0; b < n; ++b)
      if (b != a) v *= (x - nodes[b]) / (nodes[a] - nodes[b]);
    val[a] = v;
    double d = 0.0;
    for (int c = 0; c < n; ++c) {
      if (c == a) continue;
      double t = 1.0 / (nodes[a] - nodes[c]);
      for (int b = 0; b < n; ++b)
        if (b != a && b != c) t *= (x - nodes[b]) / (nodes[a] - nodes[b]);
      d += t;
    }
    der[a] = d;
  }
}

struct RefElem1D {
  int order, nq;
  double qx[kMaxQ], qw[kMaxQ];
  double Nq[kMaxQ][kMaxN1];   // N_a(q)
  double dNq[kMaxQ][kMaxN1];  // N'_a(q)
  double M[kMaxN1][kMaxN1];   // sum_q w N_a N_b
  double K[kMaxN1][kMaxN1];   // sum_q w N'_a N'_b
  double C[kMaxN1][kMaxN1];   // sum_q w N'_a N_b
  double NL[kMaxN1], NR[kMaxN1];    // N(-1), N(+1)
  double dNL[kMaxN1], dNR[kMaxN1];  // N'(-1), N'(+1)
};

inline int build_ref1d(int order, int nq, RefElem1D* r) {
  if (order < 1 || order > kMaxOrder || nq < 1 || nq > kMaxQ) return 1;
  std::memset(r, 0, sizeof(*r));
  r->order = order;
  r->nq = nq;
  gauss_legendre(nq, r->qx, r->qw);
  const int n = order + 1;
  for (int q = 0; q < nq; ++q) lagrange_1d(order, r->qx[q], r->Nq[q], r->dNq[q]);
  for (int q = 0; q < nq; ++q)
    for (int a = 0; a < n; ++a)
      for (int b = 0; b < n; ++b) {
        r->M[a][b] += r->qw[q] * r->Nq[q][a] * r->Nq[q][b];
        r->K[a][b] += r->qw[q] * r->dNq[q][a] * r->dNq[q][b];
        r->C[a][b] += r->qw[q] * r->dNq[q][a] * r->Nq[q][b];
      }
  lagrange_1d(order, -1.0, r->NL, r->dNL);
  lagrange_1d(order, +1.0, r->NR, r->dNR);
  return 0;
}

}  // namespace stefan
