// Handle behind `stefan_ldg*` (shared by ldg.cu and ldg_solve.cu).
#pragma once
#include <cuda_runtime.h>

#include <cstdint>

#include "ldg_tables.h"
#include "march_plan.h"

namespace stefan {

constexpr int kLdgTailRing = 8;

struct LdgHandle {
  int nx, ny, order, numqp;
  double phys_hx, phys_hy;
  LdgPhys phys;
  bool has_lo, has_hi;
  int8_t* phase_dev;
  MarchPlan plan;  // run lists, fix list and CTA tiles of the march kernel
  int variant;     // CTA shape of the march kernel (ldg_march_variant)
  unsigned* tail_ctr_dev;  // ring of {next, finished} counter pairs of the march kernel's tail queue
  unsigned launch_seq;
  void* tab_host;
  int64_t ndofs;
};

}  // namespace stefan
