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
#include "../../stefanproblem_b200/csrc/march_plan.h"

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

// the same with the mesh-aligned interface terms (LdgPhys::aligned_interface)
extern "C" int hostcheck_ldg_apply_aligned(int order, int nx, int ny, int nq, double jx, double jy, double k1,
                                           double k2, double interiorpenalty, double negboundarypenalty,
                                           double posboundarypenalty, double V0x, double V0y, double interfacepenalty,
                                           const int8_t* phase, const double* x, double* y) {
  LdgPhys ph{k1, k2, interiorpenalty, negboundarypenalty, posboundarypenalty, V0x, V0y, interfacepenalty, 1};
  switch (order) {
    case 1: return run_ldg<1>(nx, ny, nq, jx, jy, ph, phase, x, y);
    case 2: return run_ldg<2>(nx, ny, nq, jx, jy, ph, phase, x, y);
    case 3: return run_ldg<3>(nx, ny, nq, jx, jy, ph, phase, x, y);
  }
  return 3;
}


// The host plan of the march kernels (march_plan.h) for a mesh: run lists, fix list and the
// CTA tiles of the column range [cbeg, cend).  phase: int8 +1/-1 per cell (no halos).
// Outputs (caller-sized): runs [4 * nruns] (with the 3 sentinels per strip), run_ofs
// [nstrips + 1], fix [nx*ny], tiles [4 * 640]; counts come back through n_out[4] =
// {nruns, nstrips, nfix, ntiles}.
extern "C" int hostcheck_march_plan(int nx, int ny, const int8_t* phase, int group_warps, int cbeg, int cend,
                                    int max_ctas, int* runs, int* run_ofs, int* fix, int* tiles, int* n_out) {
  std::vector<int8_t> ph((size_t)(nx + 2) * (ny + 2), 0);
  for (int i = 0; i < nx; ++i)
    for (int j = 0; j < ny; ++j) ph[(size_t)(i + 1) * (ny + 2) + (j + 1)] = phase[(size_t)i * ny + j] == 1 ? 1 : 2;
  MarchPlan plan;
  plan.compute(ph.data(), nx, ny, group_warps);
  const SegCache& sc = plan.segments(cbeg, cend, max_ctas);
  for (size_t k = 0; k < plan.runs_host.size(); ++k) {
    runs[4 * k + 0] = plan.runs_host[k].x;
    runs[4 * k + 1] = plan.runs_host[k].y;
    runs[4 * k + 2] = plan.runs_host[k].z;
    runs[4 * k + 3] = plan.runs_host[k].w;
  }
  for (int s = 0; s <= plan.nstrips; ++s) run_ofs[s] = plan.run_ofs_host[s];
  for (int k = 0; k < plan.nfix; ++k) fix[k] = plan.fix_host[k];
  for (int k = 0; k < sc.n; ++k) {
    tiles[4 * k + 0] = sc.tab.e[k].x;
    tiles[4 * k + 1] = sc.tab.e[k].y;
    tiles[4 * k + 2] = sc.tab.e[k].z;
    tiles[4 * k + 3] = sc.tab.e[k].w;
  }
  n_out[0] = (int)plan.runs_host.size();
  n_out[1] = plan.nstrips;
  n_out[2] = plan.nfix;
  n_out[3] = sc.n;
  return 0;
}


// ldg_rows_streamed (the march kernel's form) on the host for an all-interior evaluation:
// a 3 x 3 patch of cells of one phase; returns the rows of the centre cell computed by the
// general row function and by the streamed one (traces taken from the neighbours).
template <int P>
static int run_ldg_streamed(int nq, double jx, double jy, LdgPhys ph, int s, const double* x9, double* out_general,
                            double* out_streamed) {
  constexpr int N1 = P + 1, NF = N1 * N1, CD = 3 * NF;
  RefElem1D r;
  if (build_ref1d(P, nq, &r)) return 1;
  static LdgTab<P> tab;
  if (build_ldg_tables<P>(r, jx, jy, ph, &tab)) return 2;
  auto cell = [&](int i, int j) { return x9 + (size_t)(i * 3 + j) * CD; };  // y fastest
  const double *U = cell(1, 1), *UL = cell(0, 1), *UR = cell(2, 1), *UD = cell(1, 0), *UU = cell(1, 2);
  ldg_cell_rows<P>(tab, s, T_SAME, T_SAME, T_SAME, T_SAME, U, UL, UR, UD, UU, out_general);
  double tLT[N1], tLQ[N1], tRT[N1], tRQ[N1], tDT[N1], tDQ[N1], tUT[N1], tUQ[N1];
  for (int q = 0; q < N1; ++q) {
    tLT[q] = UL[3 * (P * N1 + q) + 0];
    tLQ[q] = UL[3 * (P * N1 + q) + 1];
    tRT[q] = UR[3 * (0 * N1 + q) + 0];
    tRQ[q] = UR[3 * (0 * N1 + q) + 1];
    tDT[q] = UD[3 * (q * N1 + P) + 0];
    tDQ[q] = UD[3 * (q * N1 + P) + 2];
    tUT[q] = UU[3 * (q * N1 + 0) + 0];
    tUQ[q] = UU[3 * (q * N1 + 0) + 2];
  }
  if (s == 0) ldg_rows_streamed<P, 0>(tab, U, tLT, tLQ, tRT, tRQ, tDT, tDQ, tUT, tUQ, out_streamed);
  else ldg_rows_streamed<P, 1>(tab, U, tLT, tLQ, tRT, tRQ, tDT, tDQ, tUT, tUQ, out_streamed);
  // every combination of face types: the general streamed form (the kernels' tail) against the
  // general row function; both outputs are appended after the first CD entries
  int n = 1;
  for (int tL = 0; tL < 3; ++tL)
    for (int tR = 0; tR < 3; ++tR)
      for (int tD = 0; tD < 3; ++tD)
        for (int tU = 0; tU < 3; ++tU, ++n) {
          ldg_cell_rows<P>(tab, s, tL, tR, tD, tU, U, UL, UR, UD, UU, out_general + (size_t)n * CD);
          ldg_rows_streamed_general<P>(tab, s, tL, tR, tD, tU, U, tLT, tLQ, tRT, tRQ, tDT, tDQ, tUT, tUQ,
                                       out_streamed + (size_t)n * CD);
        }
  // entry 82: the progressive form (the order-3 march loop), all-interior
  ldg_cell_rows<P>(tab, s, T_SAME, T_SAME, T_SAME, T_SAME, U, UL, UR, UD, UU, out_general + (size_t)n * CD);
  auto du = [&](int ix, double& dT, double& dQ, double& uT, double& uQ) {
    dT = tDT[ix]; dQ = tDQ[ix]; uT = tUT[ix]; uQ = tUQ[ix];
  };
  auto tr = [&](double* rT, double* rQ) {
    for (int q = 0; q < N1; ++q) { rT[q] = tRT[q]; rQ[q] = tRQ[q]; }
  };
  double* op = out_streamed + (size_t)n * CD;
  auto emit = [&](int jx, const double* rows) {
    for (int q = 0; q < 3 * N1; ++q) op[3 * N1 * jx + q] = rows[q];
  };
  if (s == 0) ldg_rows_progressive<P, 0>(tab, U, tLT, tLQ, du, tr, emit);
  else ldg_rows_progressive<P, 1>(tab, U, tLT, tLQ, du, tr, emit);
  return 0;
}

extern "C" int hostcheck_ldg_streamed(int order, int nq, double jx, double jy, double k1, double k2, double ai,
                                      double an, double ap, double v0x, double v0y, int s, const double* x9,
                                      double* out_general, double* out_streamed) {
  LdgPhys ph{k1, k2, ai, an, ap, v0x, v0y};
  switch (order) {
    case 1: return run_ldg_streamed<1>(nq, jx, jy, ph, s, x9, out_general, out_streamed);
    case 2: return run_ldg_streamed<2>(nq, jx, jy, ph, s, x9, out_general, out_streamed);
    case 3: return run_ldg_streamed<3>(nq, jx, jy, ph, s, x9, out_general, out_streamed);
  }
  return 3;
}
