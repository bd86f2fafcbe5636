// Device side of the peer-halo protocol (epoch flags in peer-mappable memory), shared by the march kernels
// (ipdg.cu) and the interface-sums kernel (ipdg_aux.cu).
#pragma once
#include <cuda_runtime.h>

namespace stefan {

// Wait (acquire, system scope) until the neighbour has published `epoch`; bounded.
__device__ __forceinline__ void peer_flag_wait(const unsigned long long* flag, unsigned long long epoch,
                                               unsigned long long timeout_ns, int* status) {
  unsigned long long t0, t, v;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t0));
  for (;;) {
    asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(flag) : "memory");
    if (v >= epoch) {
      // the neighbour's column is fetched next by a bulk copy (asynchronous proxy): order the
      // acquire (generic proxy) before it
      asm volatile("fence.proxy.async.global;" ::: "memory");
      return;
    }
    __nanosleep(100);
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    if (t - t0 > timeout_ns) {
      if (status) atomicExch(status, 1);
      return;
    }
  }
}
// the epoch of this launch (call after griddep_wait: the previous launch's last CTA wrote the counter)
__device__ __forceinline__ unsigned long long peer_epoch(const unsigned long long* epoch_dev, unsigned long long epoch) {
  if (epoch_dev == nullptr) return epoch;
  unsigned long long v;
  asm volatile("ld.relaxed.gpu.global.u64 %0, [%1];" : "=l"(v) : "l"(epoch_dev) : "memory");
  return v + 1;
}
// Publish `epoch` to the local flag and to the neighbours' mailboxes (null = none): ONE system-scope fence, then
// relaxed stores (fence + relaxed store is a release pattern).  A release store per target costs a system fence
// each -- measured: three of them at the kernel's start and end added 9 us to a 48 us step at 8 GPUs.
__device__ __forceinline__ void peer_flag_publish(unsigned long long* flag, unsigned long long epoch,
                                                  unsigned long long* post0 = nullptr, unsigned long long* post1 = nullptr) {
  __threadfence_system();
  asm volatile("st.relaxed.sys.global.u64 [%0], %1;" ::"l"(flag), "l"(epoch) : "memory");
  if (post0) asm volatile("st.relaxed.sys.global.u64 [%0], %1;" ::"l"(post0), "l"(epoch) : "memory");
  if (post1) asm volatile("st.relaxed.sys.global.u64 [%0], %1;" ::"l"(post1), "l"(epoch) : "memory");
}

}  // namespace stefan
