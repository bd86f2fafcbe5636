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


// The host plan of the march kernels (march_plan.h) for a mesh: run lists, fix list and the
// CTA tiles of the column range [cbeg, cend).  phase: int8 +1/-1 per cell (no halos).
// Outputs (caller-sized): runs [4 * nruns] (with the 3 sentinels per strip), run_ofs
// [nstrips], fix [nx*ny], tiles [4 * 640]; counts come back through n_out[4] =
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
  for (int s = 0; s < plan.nstrips; ++s) run_ofs[s] = plan.run_ofs_host[s];
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
