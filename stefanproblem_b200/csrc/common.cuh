// Shared helpers for the sm_100a kernels and the C-ABI layer.
#pragma once
#include <cuda_runtime.h>
#include <cstdarg>
#include <cstdint>
#include <cstdio>

namespace stefan {

enum Status : int {
  ST_OK = 0,
  ST_BAD_ARGUMENT = 1,
  ST_CUDA_ERROR = 2,
  ST_UNSUPPORTED = 3,
  ST_OUT_OF_RANGE = 4,  // e.g. the FD interface closer than 3 nodes to an end (BoundsError in the reference)
  ST_CAPACITY = 5,
};

// thread-local message returned by stefan_last_error()
char* last_error_buffer();

inline int fail(int status, const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(last_error_buffer(), 512, fmt, ap);
  va_end(ap);
  return status;
}

#define STEFAN_CUDA(call)                                                              \
  do {                                                                                 \
    cudaError_t e__ = (call);                                                          \
    if (e__ != cudaSuccess)                                                            \
      return ::stefan::fail(::stefan::ST_CUDA_ERROR, "%s:%d: %s -> %s", __FILE__,      \
                            __LINE__, #call, cudaGetErrorString(e__));                 \
  } while (0)

#define STEFAN_REQUIRE(cond, ...)                                                      \
  do {                                                                                 \
    if (!(cond)) return ::stefan::fail(::stefan::ST_BAD_ARGUMENT, __VA_ARGS__);        \
  } while (0)

// Function attributes (the > 48 KB shared-memory opt-in) and SM counts are per DEVICE, not per
// process: everything cached about them is keyed by the current device.
constexpr int kMaxDevices = 64;
inline int current_device() {
  int dev = 0;
  cudaGetDevice(&dev);
  return (dev >= 0 && dev < kMaxDevices) ? dev : 0;
}
struct PerDeviceFlag {
  bool set[kMaxDevices] = {};
  bool& here() { return set[current_device()]; }
};

inline int sm_count() {
  static int n[kMaxDevices] = {};
  const int dev = current_device();
  if (n[dev] == 0) {
    cudaDeviceGetAttribute(&n[dev], cudaDevAttrMultiProcessorCount, dev);
    if (n[dev] <= 0) n[dev] = 148;
  }
  return n[dev];
}

#if defined(__CUDACC__)
// Ampere-style asynchronous global->shared copies (LDGSTS); 16-byte pieces bypass
// L1 (.cg), 8-byte pieces go through it (.ca).
__device__ __forceinline__ void cp_async16(void* smem, const void* gmem) {
  unsigned s = (unsigned)__cvta_generic_to_shared(smem);
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;\n" ::"r"(s), "l"(gmem) : "memory");
}
__device__ __forceinline__ void cp_async8(void* smem, const void* gmem) {
  unsigned s = (unsigned)__cvta_generic_to_shared(smem);
  asm volatile("cp.async.ca.shared.global [%0], [%1], 8;\n" ::"r"(s), "l"(gmem) : "memory");
}
__device__ __forceinline__ void cp_async4(void* smem, const void* gmem) {
  unsigned s = (unsigned)__cvta_generic_to_shared(smem);
  asm volatile("cp.async.ca.shared.global [%0], [%1], 4;\n" ::"r"(s), "l"(gmem) : "memory");
}
__device__ __forceinline__ void cp_async_commit() {
  asm volatile("cp.async.commit_group;\n" ::: "memory");
}
template <int N>
__device__ __forceinline__ void cp_async_wait() {
  asm volatile("cp.async.wait_group %0;\n" ::"n"(N) : "memory");
}
#endif

}  // namespace stefan
