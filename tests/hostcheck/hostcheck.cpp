// TEST INFRASTRUCTURE ONLY (not part of the product, not a fallback).
// Runs the host-device per-cell row functions of stefanproblem_b200/csrc on the
// CPU, cell by cell, so that the operator algebra (tables + row formulas) can be
// checked against the oracle on the build box, where there is no GPU.  The GPU
// kernels call exactly the same functions; what this cannot cover is the data
// movement of the kernels, which the `-m gpu` tests cover.
#include <cstdint>
#include <vector>
#include "../../stefanproblem_b200/csrc/ipdg_tables.h"
#include "../../stefanproblem_b200/csrc/ldg_tables.h"

using namespace stefan;

template <int P>
static int run_ip(int nx, int ny, int nq, double jx, double jy, IpPhys ph,
                  const int8_t* phase, const double* x, double* y) {
  constexpr int NF = (P + 1) * (P + 1);
  RefElem1D r;
  if (build_ref1d(P, nq, &r)) return 1;
  static IpTab<P> tab;
  if (build_ip_tables<P>(r, jx, jy, ph, &tab)) return 2;
  auto sign = [&](int i, int j) -> int {
    return (i < 0 || i >= nx || j < 0 || j >= ny) ? 0 : phase[i * ny + j];
  };
  std::vector<double> zero(NF, 0.0);
  for (int i = 0; i < nx; ++i)
    for (int j = 0; j < ny; ++j) {
      const int s = sign(i, j);
      auto cell = [&](int ii, int jj) -> const double* {
        return sign(ii, jj) == 0 ? zero.data() : x + (size_t)(ii * ny + jj) * NF;
      };
      ip_cell_rows<P>(tab, s > 0 ? 0 : 1, ip_face_type(s, sign(i - 1, j)),
                      ip_face_type(s, sign(i + 1, j)),
                      ip_face_type(s, sign(i, j - 1)),
                      ip_face_type(s, sign(i, j + 1)), cell(i, j), cell(i - 1, j),
                      cell(i + 1, j), cell(i, j - 1), cell(i, j + 1),
                      y + (size_t)(i * ny + j) * NF);
    }
  return 0;
}

extern "C" int hostcheck_ipdg_apply(int order, int nx, int ny, int nq, double jx, double jy,
                                    double k1, double k2, double penalty, double lambda,
                                    int variant, int terms, const int8_t* phase,
                                    const double* x, double* y) {
  IpPhys ph{k1, k2, penalty, lambda, variant, terms};
  switch (order) {
    case 1: return run_ip<1>(nx, ny, nq, jx, jy, ph, phase, x, y);
    case 2: return run_ip<2>(nx, ny, nq, jx, jy, ph, phase, x, y);
    case 3: return run_ip<3>(nx, ny, nq, jx, jy, ph, phase, x, y);
  }
  return 3;
}

template <int P>
static int run_ldg(int nx, int ny, int nq, double jx, double jy, LdgPhys ph, const int8_t* phase,
                   const double* x, double* y) {
  constexpr int NF = (P + 1) * (P + 1);
  RefElem1D r;
  if (build_ref1d(P, nq, &r)) return 1;
  static LdgTab<P> tab;
  if (build_ldg_tables<P>(r, jx, jy, ph, &tab)) return 2;
  auto sign = [&](int i, int j) -> int {
    return (i < 0 || i >= nx || j < 0 || j >= ny) ? 0 : phase[i * ny + j];
  };
  std::vector<double> zero(3 * NF, 0.0);
  for (int i = 0; i < nx; ++i)
    for (int j = 0; j < ny; ++j) {
      const int s = sign(i, j);
      auto cell = [&](int ii, int jj) -> const double* {
        return sign(ii, jj) == 0 ? zero.data() : x + (size_t)(ii * ny + jj) * 3 * NF;
      };
      ldg_cell_rows<P>(tab, s > 0 ? 0 : 1, ip_face_type(s, sign(i - 1, j)), ip_face_type(s, sign(i + 1, j)),
                       ip_face_type(s, sign(i, j - 1)), ip_face_type(s, sign(i, j + 1)), cell(i, j),
                       cell(i - 1, j), cell(i + 1, j), cell(i, j - 1), cell(i, j + 1),
                       y + (size_t)(i * ny + j) * 3 * NF);
    }
  return 0;
}

extern "C" int hostcheck_ldg_apply(int order, int nx, int ny, int nq, double jx, double jy, double k1,
                                   double k2, double interiorpenalty, double negboundarypenalty,
                                   double posboundarypenalty, double V0x, double V0y, const int8_t* phase,
                                   const double* x, double* y) {
  LdgPhys ph{k1, k2, interiorpenalty, negboundarypenalty, posboundarypenalty, V0x, V0y};
  switch (order) {
    case 1: return run_ldg<1>(nx, ny, nq, jx, jy, ph, phase, x, y);
    case 2: return run_ldg<2>(nx, ny, nq, jx, jy, ph, phase, x, y);
    case 3: return run_ldg<3>(nx, ny, nq, jx, jy, ph, phase, x, y);
  }
  return 3;
}
