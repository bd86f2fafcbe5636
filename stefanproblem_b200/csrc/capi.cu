// Library-wide C-ABI entry points (error string, version).
#include "../../include/stefan_b200.h"
#include "common.cuh"
#include "refelem.h"

namespace stefan {
char* last_error_buffer() {
  static thread_local char buf[512] = {0};
  return buf;
}
}  // namespace stefan

extern "C" {
const char* stefan_last_error(void) { return stefan::last_error_buffer(); }
int stefan_version(void) { return 100; }

int stefan_reference_tables(int32_t order, int32_t numqp, double* mass, double* stiff, double* adv,
                            double* dleft, double* dright, double* qpoints, double* qweights) {
  stefan::RefElem1D r;
  if (stefan::build_ref1d(order, numqp, &r))
    return stefan::fail(stefan::ST_BAD_ARGUMENT, "stefan_reference_tables: order must be 1..3, numqp 1..%d",
                        stefan::kMaxQ);
  const int n = order + 1;
  for (int a = 0; a < n; ++a) {
    for (int b = 0; b < n; ++b) {
      if (mass) mass[a * n + b] = r.M[a][b];
      if (stiff) stiff[a * n + b] = r.K[a][b];
      if (adv) adv[a * n + b] = r.C[a][b];
    }
    if (dleft) dleft[a] = r.dNL[a];
    if (dright) dright[a] = r.dNR[a];
  }
  for (int q = 0; q < numqp; ++q) {
    if (qpoints) qpoints[q] = r.qx[q];
    if (qweights) qweights[q] = r.qw[q];
  }
  return stefan::ST_OK;
}
}
