// Handle behind `stefan_ipdg*` (shared by ipdg.cu and ipdg_aux.cu).
#pragma once
#include <cuda_runtime.h>

#include <cstdint>
#include <vector>

#include "ipdg_tables.h"

namespace stefan {

struct IpdgHandle {
  int nx, ny, order, numqp;
  double hx, hy;
  IpPhys phys;
  double Tm;
  bool has_lo, has_hi;
  int8_t* phase_dev;  // (nx+2)*(ny+2), padded
  uint8_t* flags_dev;  // [nstrips][nxpad]
  int32_t* fix_dev;    // cells recomputed by the fix-up kernel
  int nfix;
  std::vector<int> fix_col_ofs;  // fix cells are sorted by column: [nx+1] offsets
  double *xh_dev, *yh_dev;       // device staging of the host path
  cudaStream_t hs[3];
  cudaEvent_t hev[64];
  bool host_ready;
  int2* runs_dev;      // per-strip run lists (order-1 kernel)
  int* run_ofs_dev;
  void* tab_host;     // IpTab<P>
  RefElem1D ref;
  int64_t ndofs;
};

}  // namespace stefan
