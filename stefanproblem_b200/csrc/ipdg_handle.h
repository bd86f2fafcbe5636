// Handle behind `stefan_ipdg*` (shared by ipdg.cu and ipdg_aux.cu).
#pragma once
#include <cuda_runtime.h>

#include <cstdint>
#include <vector>

#include "ipdg_tables.h"

namespace stefan {

// One entry per CTA of the march kernel: {strip group, first column, one past the last, -}
constexpr int kMaxMarchCtas = 320;
struct SegTab {
  int4 e[kMaxMarchCtas];
};
struct SegCache {
  int cbeg, cend, max_ctas, n;
  SegTab tab;
};

struct IpdgHandle {
  int nx, ny, order, numqp;
  double hx, hy;
  IpPhys phys;
  double Tm;
  bool has_lo, has_hi;
  int8_t* phase_dev;  // (nx+2)*(ny+2), padded
  int32_t* fix_dev;    // cells recomputed by the kernels' tail (not interior-uniform)
  int nfix;
  int ngroups, group_warps;      // strip groups (the strips of one CTA) of the march kernel
  std::vector<int> mixed_cum;    // [nstrips][nx+1] running count of phase-mixing columns
  std::vector<SegCache> seg_cache;  // CTA tiles per requested column range
  std::vector<int> fix_col_ofs;  // fix cells are sorted by column: [nx+1] offsets
  double *xh_dev, *yh_dev;       // device staging of the host path
  cudaStream_t hs[3];
  cudaEvent_t hev[64];
  bool host_ready;
  int4* runs_dev;      // per-strip run lists of the march kernels (see stefan_ipdg_create)
  int* run_ofs_dev;
  void* tab_host;     // IpTab<P>
  RefElem1D ref;
  int64_t ndofs;
};

}  // namespace stefan
