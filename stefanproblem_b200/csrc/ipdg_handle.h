// Handle behind `stefan_ipdg*` (shared by ipdg.cu and ipdg_aux.cu).
#pragma once
#include <cuda_runtime.h>

#include <cstdint>
#include <vector>

#include "ipdg_tables.h"
#include "march_plan.h"

namespace stefan {

struct IpdgHandle {
  int nx, ny, order, numqp;
  double hx, hy;
  IpPhys phys;
  double Tm;
  bool has_lo, has_hi;
  int8_t* phase_dev;  // (nx+2)*(ny+2), padded
  MarchPlan plan;      // run lists, fix list and CTA tiles of the march kernel
  double *xh_dev, *yh_dev;       // device staging of the host path
  unsigned* sync_counter_dev;    // finished-CTA counter of stefan_ipdg_apply_peer
  cudaStream_t hs[3];
  cudaEvent_t hev[64];
  bool host_ready;
  void* tab_host;     // IpTab<P>
  RefElem1D ref;
  int64_t ndofs;
};

}  // namespace stefan
