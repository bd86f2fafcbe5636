// Device helpers shared by the march kernels (ipdg_march.cuh, ldg_march.cuh): TMA bulk /
// tensor copies, mbarriers, shared-memory accesses through 32-bit shared-window addresses.
#pragma once
#include <cuda.h>  // CUtensorMap

#include "march_plan.h"

namespace stefan {

// 32 bytes per lane in one instruction (STG.256, sm_100): a warp writes whole
// 32-byte sectors; two STG.128 per lane would send every sector to L2 twice, half
// filled (measured: L1->XBAR request path 66% busy, 43% of it stores).
__device__ __forceinline__ void st_global_v4(double* p, double a, double b, double c, double d) {
  asm volatile("st.global.v4.f64 [%0], {%1,%2,%3,%4};" ::"l"(p), "d"(a), "d"(b), "d"(c), "d"(d)
               : "memory");
}
__device__ __forceinline__ void ld_global_nc_v4(const double* p, double* v) {
  asm volatile("ld.global.nc.v4.f64 {%0,%1,%2,%3}, [%4];" : "=d"(v[0]), "=d"(v[1]), "=d"(v[2]), "=d"(v[3]) : "l"(p));
}
__device__ __forceinline__ unsigned smem_addr(const void* p) {
  return (unsigned)__cvta_generic_to_shared(p);
}
__device__ __forceinline__ void fence_proxy_async() {
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}
__device__ __forceinline__ void tma_store_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void tma_store_wait_read() {
  asm volatile("cp.async.bulk.wait_group.read %0;" ::"n"(N) : "memory");
}

// 32-bit shared-window addresses keep the loop's address arithmetic in single registers
__device__ __forceinline__ double2 lds_v2(unsigned addr) {
  double2 v;
  asm volatile("ld.shared.v2.f64 {%0,%1}, [%2];" : "=d"(v.x), "=d"(v.y) : "r"(addr) : "memory");
  return v;
}
__device__ __forceinline__ double lds_f64(unsigned addr) {
  double v;
  asm volatile("ld.shared.f64 %0, [%1];" : "=d"(v) : "r"(addr) : "memory");
  return v;
}
__device__ __forceinline__ void sts_f64(unsigned addr, double v) {
  asm volatile("st.shared.f64 [%0], %1;" ::"r"(addr), "d"(v) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx_s(unsigned bar, unsigned bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive_s(unsigned bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void mbar_wait_s(unsigned bar, unsigned parity) {
  asm volatile(
      "{\n.reg .pred p;\nWAIT_%=:\nmbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n@!p bra WAIT_%=;\n}" ::"r"(bar),
      "r"(parity)
      : "memory");
}
#ifndef MARCH_L2_HINT
#define MARCH_L2_HINT 0  // 1: evict_first, 2: evict_last on the streaming TMA loads (tuning experiment)
#endif
__device__ __forceinline__ void bulk_load_s(unsigned dst, const void* src, unsigned bytes, unsigned bar) {
#if MARCH_L2_HINT
  unsigned long long pol;
#if MARCH_L2_HINT == 1
  asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(pol));
#else
  asm volatile("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(pol));
#endif
  asm volatile(
      "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1], %2, [%3], %4;" ::"r"(dst),
      "l"(src), "r"(bytes), "r"(bar), "l"(pol)
      : "memory");
#else
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst),
               "l"(src), "r"(bytes), "r"(bar)
               : "memory");
#endif
}
__device__ __forceinline__ void tensor2d_load_s(unsigned dst, const CUtensorMap* map, int c0, int c1, unsigned bar) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];" ::"r"(dst),
      "l"(map), "r"(c0), "r"(c1), "r"(bar)
      : "memory");
}
__device__ __forceinline__ void bulk_store_s(void* gdst, unsigned ssrc, unsigned bytes) {
  asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(gdst), "r"(ssrc), "r"(bytes)
               : "memory");
}
// A run {end column, mode, phase mask, store mask}, loaded only when a run ends (an
// ordinary load would be hoisted out of the rare branch and every iteration would wait
// for it); the kernel keeps one run of look-ahead so the load is never waited for.
__device__ __forceinline__ int4 ld_run(const int4* p) {
  int4 v;
  asm volatile("ld.global.nc.v4.s32 {%0,%1,%2,%3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(p));
  return v;
}


}  // namespace stefan
