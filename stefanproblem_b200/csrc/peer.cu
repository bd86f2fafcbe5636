// Peer memory for the column-slab partition: slab vectors that the neighbouring ranks'
// kernels read directly over NVLink, and the epoch flags that order those reads.
//
// One process per GPU.  A rank allocates its slab vector with stefan_peer_alloc (plain
// cudaMalloc memory + a CUDA IPC handle), ships the 64-byte handle to its neighbours
// through whatever the host program uses (torch.distributed in the Python twin), and the
// neighbours map it with stefan_peer_open.  The apply kernels take the neighbour's edge
// column as x_lo / x_hi pointers, so a mapped peer pointer makes the halo "exchange" a
// TMA bulk copy issued by the marching warps themselves: no pack / send / recv / unpack
// kernels and no extra launch for the edge columns.
//
// Ordering: stefan_peer_signal(flag, epoch) publishes "my vector of this epoch is final"
// (release, system scope) on the stream; stefan_peer_wait(peer_flag, epoch) holds the
// stream until the neighbour has published that epoch (acquire, system scope).  The same
// pair, on a second flag, tells a rank that its neighbours are done reading before it
// overwrites its vector.  The waiting kernel gives up after `timeout_ms` and records it in
// *status instead of hanging the GPU.
#include "../../include/stefan_b200.h"
#include "common.cuh"

#include <cstring>

namespace stefan {

__global__ void peer_signal_kernel(unsigned long long* flag, unsigned long long epoch) {
  __threadfence_system();
  asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(flag), "l"(epoch) : "memory");
}

__global__ void peer_wait_kernel(const unsigned long long* flag, unsigned long long epoch,
                                 unsigned long long timeout_ns, int* status) {
  unsigned long long t0, t, v;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t0));
  for (;;) {
    asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(flag) : "memory");
    if (v >= epoch) return;
    __nanosleep(200);
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    if (t - t0 > timeout_ns) {
      if (status) atomicExch(status, 1);
      return;
    }
  }
}

}  // namespace stefan

using namespace stefan;

extern "C" {

int stefan_peer_alloc(int64_t nbytes, void** dev_ptr, stefan_ipc_handle* handle) {
  STEFAN_REQUIRE(nbytes > 0 && dev_ptr && handle, "stefan_peer_alloc: bad argument");
  static_assert(sizeof(cudaIpcMemHandle_t) <= sizeof(stefan_ipc_handle), "handle size");
  void* p = nullptr;
  STEFAN_CUDA(cudaMalloc(&p, (size_t)nbytes));
  cudaError_t e = cudaMemset(p, 0, (size_t)nbytes);
  cudaIpcMemHandle_t h;
  if (e == cudaSuccess) e = cudaIpcGetMemHandle(&h, p);
  if (e != cudaSuccess) {
    cudaFree(p);
    return fail(ST_CUDA_ERROR, "stefan_peer_alloc: %s", cudaGetErrorString(e));
  }
  std::memset(handle, 0, sizeof(*handle));
  std::memcpy(handle, &h, sizeof(h));
  *dev_ptr = p;
  return ST_OK;
}

int stefan_peer_free(void* dev_ptr) {
  if (dev_ptr) STEFAN_CUDA(cudaFree(dev_ptr));
  return ST_OK;
}

int stefan_peer_open(const stefan_ipc_handle* handle, void** peer_ptr) {
  STEFAN_REQUIRE(handle && peer_ptr, "stefan_peer_open: null argument");
  cudaIpcMemHandle_t h;
  std::memcpy(&h, handle, sizeof(h));
  STEFAN_CUDA(cudaIpcOpenMemHandle(peer_ptr, h, cudaIpcMemLazyEnablePeerAccess));
  return ST_OK;
}

int stefan_peer_close(void* peer_ptr) {
  if (peer_ptr) STEFAN_CUDA(cudaIpcCloseMemHandle(peer_ptr));
  return ST_OK;
}

int stefan_peer_signal(uint64_t* flag, uint64_t epoch, void* stream) {
  STEFAN_REQUIRE(flag, "stefan_peer_signal: null flag");
  peer_signal_kernel<<<1, 1, 0, static_cast<cudaStream_t>(stream)>>>(reinterpret_cast<unsigned long long*>(flag),
                                                                     (unsigned long long)epoch);
  STEFAN_CUDA(cudaGetLastError());
  return ST_OK;
}

int stefan_peer_wait(const uint64_t* peer_flag, uint64_t epoch, int32_t timeout_ms, int32_t* status_dev,
                     void* stream) {
  STEFAN_REQUIRE(peer_flag && timeout_ms > 0, "stefan_peer_wait: bad argument");
  peer_wait_kernel<<<1, 1, 0, static_cast<cudaStream_t>(stream)>>>(
      reinterpret_cast<const unsigned long long*>(peer_flag), (unsigned long long)epoch,
      (unsigned long long)timeout_ms * 1000000ull, status_dev);
  STEFAN_CUDA(cudaGetLastError());
  return ST_OK;
}

}  // extern "C"
