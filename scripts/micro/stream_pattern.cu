// Microbenchmark: what does the B200 memory system give for the traversal pattern of
// the IP-DG stream kernels (warp = strip of cells in y, marching along x), with the
// arithmetic removed?  Variants isolate strip width / alignment, halo re-reads,
// prefetch depth and occupancy.  Build: nvcc -O3 -arch=sm_100a stream_pattern.cu
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("%s: %s\n", #x, cudaGetErrorString(e)); exit(1);} } while (0)

__device__ __forceinline__ void ld4(const double* p, double* v) {
  asm volatile("ld.global.v4.f64 {%0,%1,%2,%3}, [%4];" : "=d"(v[0]), "=d"(v[1]), "=d"(v[2]), "=d"(v[3]) : "l"(p));
}
__device__ __forceinline__ void st4(double* p, const double* v) {
  asm volatile("st.global.v4.f64 [%0], {%1,%2,%3,%4};" ::"l"(p), "d"(v[0]), "d"(v[1]), "d"(v[2]), "d"(v[3]) : "memory");
}

// plain grid-stride copy, 32 B per thread per access, UNROLL accesses in flight
template <int UNROLL>
__global__ void copy_plain(const double* __restrict__ x, double* __restrict__ y, size_t n4) {
  size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  const size_t stride = (size_t)gridDim.x * blockDim.x;
  for (; i + (UNROLL - 1) * stride < n4; i += UNROLL * stride) {
    double v[UNROLL][4];
#pragma unroll
    for (int u = 0; u < UNROLL; ++u) ld4(x + 4 * (i + u * stride), v[u]);
#pragma unroll
    for (int u = 0; u < UNROLL; ++u) st4(y + 4 * (i + u * stride), v[u]);
  }
}

// strip traversal: warp owns OUT cells (+HALO each side when HALO) of 32 B in y and
// marches along x over [c0,c1); R register columns deep.
template <int OUT, int HALO, int R, int WARPS, int SCATTER = 0, int MODE = 0>
__global__ void __launch_bounds__(WARPS * 32)
copy_strips(const double* __restrict__ x, double* __restrict__ y, int nx, int ny, int nseg) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int nstrips = (ny + OUT - 1) / OUT;
  const int ngroups = (nstrips + WARPS - 1) / WARPS;
  const int group = blockIdx.x % ngroups, seg = blockIdx.x / ngroups;
  const int strip = SCATTER ? warp * ngroups + group : group * WARPS + warp;
  if (strip >= nstrips) return;
  const int c0 = (int)((long long)nx * seg / nseg), c1 = (int)((long long)nx * (seg + 1) / nseg);
  const int j = strip * OUT - HALO + lane;
  const bool valid = j >= 0 && j < ny && lane < OUT + 2 * HALO;
  const bool outlane = lane >= HALO && lane < OUT + HALO && j < ny;
  const size_t cs = (size_t)ny * 4;
  double A[R][4];
#pragma unroll
  for (int k = 0; k < R; ++k) {
    if (valid && c0 + k < c1) ld4(x + (size_t)(c0 + k) * cs + (size_t)j * 4, A[k]);
    else A[k][0] = A[k][1] = A[k][2] = A[k][3] = 0;
  }
  const double* pn = x + (size_t)(c0 + R) * cs + (size_t)j * 4;
  double* yp = y + (size_t)c0 * cs + (size_t)j * 4;
  double acc = 0;
  for (int cb = c0; cb < c1; cb += R) {
#pragma unroll
    for (int k = 0; k < R; ++k) {
      const int c = cb + k;
      if (c < c1) {
        if (MODE == 1) { acc += A[k][0] + A[k][1] + A[k][2] + A[k][3]; }
        else if (outlane) st4(yp, A[k]);
        yp += cs;
        if (MODE != 2) { if (valid && c + R < c1) ld4(pn, A[k]); }
        pn += cs;
      }
    }
  }
  if (MODE == 1 && acc == 1.2345e300) y[0] = acc;
}

// same traversal, but the CTAs of one x-segment keep within SLACK blocks of K columns
// of each other (software barrier on a per-segment counter; all CTAs are resident)
template <int OUT, int HALO, int R, int WARPS, int K>
__global__ void __launch_bounds__(WARPS * 32)
copy_strips_sync(const double* __restrict__ x, double* __restrict__ y, int nx, int ny, int nseg,
                 unsigned* progress, int slack, unsigned epoch_base) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int nstrips = (ny + OUT - 1) / OUT;
  const int ngroups = (nstrips + WARPS - 1) / WARPS;
  const int group = blockIdx.x % ngroups, seg = blockIdx.x / ngroups;
  const int strip = group * WARPS + warp;
  const bool active = strip < nstrips;
  const int c0 = (int)((long long)nx * seg / nseg), c1 = (int)((long long)nx * (seg + 1) / nseg);
  const int j = strip * OUT - HALO + lane;
  const bool valid = active && j >= 0 && j < ny && lane < OUT + 2 * HALO;
  const bool outlane = active && lane >= HALO && lane < OUT + HALO && j < ny;
  const size_t cs = (size_t)ny * 4;
  double A[R][4];
#pragma unroll
  for (int k = 0; k < R; ++k) {
    if (valid && c0 + k < c1) ld4(x + (size_t)(c0 + k) * cs + (size_t)j * 4, A[k]);
    else A[k][0] = A[k][1] = A[k][2] = A[k][3] = 0;
  }
  const double* pn = x + (size_t)(c0 + R) * cs + (size_t)j * 4;
  double* yp = y + (size_t)c0 * cs + (size_t)j * 4;
  unsigned blk = 0;
  for (int cb = c0; cb < c1; cb += R) {
#pragma unroll
    for (int k = 0; k < R; ++k) {
      const int c = cb + k;
      if (c < c1) {
        if (outlane) st4(yp, A[k]);
        yp += cs;
        if (valid && c + R < c1) ld4(pn, A[k]);
        pn += cs;
      }
    }
    if (((cb - c0) / R + 1) % (K / R) == 0) {
      ++blk;
      __syncthreads();
      if (threadIdx.x == 0) {
        atomicAdd(&progress[seg], 1u);
        const unsigned need = epoch_base + (unsigned)ngroups * (blk > (unsigned)slack ? blk - slack : 0u);
        while (*((volatile unsigned*)&progress[seg]) < need) {}
      }
      __syncthreads();
    }
  }
}

template <typename F>
float timeit(F f, int reps) {
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0); cudaEventCreate(&e1);
  for (int i = 0; i < 3; ++i) f();
  CK(cudaDeviceSynchronize());
  cudaEventRecord(e0);
  for (int i = 0; i < reps; ++i) f();
  cudaEventRecord(e1);
  CK(cudaDeviceSynchronize());
  float ms; cudaEventElapsedTime(&ms, e0, e1);
  return ms / reps;
}


// ---- TMA (cp.async.bulk) version: per-warp ring of NST x 1 KiB stages, one mbarrier each
__device__ __forceinline__ unsigned smem_u32(const void* p) { return (unsigned)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(void* bar, int count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(void* bar, unsigned bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bulk_load(void* dst, const void* src, unsigned bytes, void* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
               ::"r"(smem_u32(dst)), "l"(src), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_wait(void* bar, unsigned parity) {
  asm volatile("{\n.reg .pred p;\nWAIT_LOOP:\nmbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n@!p bra WAIT_LOOP;\n}"
               ::"r"(smem_u32(bar)), "r"(parity) : "memory");
}

template <int NST, int WARPS>
__global__ void __launch_bounds__(WARPS * 32)
copy_strips_tma(const double* __restrict__ x, double* __restrict__ y, int nx, int ny, int nseg) {
  extern __shared__ __align__(128) char smem[];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  constexpr int OUT = 32;
  const int nstrips = (ny + OUT - 1) / OUT;
  const int ngroups = (nstrips + WARPS - 1) / WARPS;
  const int group = blockIdx.x % ngroups, seg = blockIdx.x / ngroups;
  const int strip = group * WARPS + warp;
  char* ring = smem + (size_t)warp * (NST * 1024 + 128);
  unsigned long long* bars = reinterpret_cast<unsigned long long*>(ring + NST * 1024);
  if (lane == 0) {
    for (int s = 0; s < NST; ++s) mbar_init(&bars[s], 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncwarp();
  if (strip >= nstrips) return;
  const int c0 = (int)((long long)nx * seg / nseg), c1 = (int)((long long)nx * (seg + 1) / nseg);
  const int jlo = strip * OUT, jhi = min(jlo + OUT, ny);
  const unsigned bytes = (unsigned)(jhi - jlo) * 32u;
  const int j = jlo + lane;
  const bool outlane = j < ny;
  const size_t cs = (size_t)ny * 4;
  const double* src = x + (size_t)c0 * cs + (size_t)jlo * 4;
  // prologue
  if (lane == 0) {
    for (int k = 0; k < NST; ++k)
      if (c0 + k < c1) { mbar_expect_tx(&bars[k], bytes); bulk_load(ring + k * 1024, src + (size_t)k * cs, bytes, &bars[k]); }
  }
  double* yp = y + (size_t)c0 * cs + (size_t)j * 4;
  int slot = 0; unsigned parity = 0;
  for (int c = c0; c < c1; ++c) {
    mbar_wait(&bars[slot], parity);
    double v[4];
    const double2 a = *reinterpret_cast<const double2*>(ring + slot * 1024 + lane * 32);
    const double2 b = *reinterpret_cast<const double2*>(ring + slot * 1024 + lane * 32 + 16);
    v[0] = a.x; v[1] = a.y; v[2] = b.x; v[3] = b.y;
    if (outlane) st4(yp, v);
    yp += cs;
    __syncwarp();
    if (lane == 0 && c + NST < c1) {
      mbar_expect_tx(&bars[slot], bytes);
      bulk_load(ring + slot * 1024, src + (size_t)(c - c0 + NST) * cs, bytes, &bars[slot]);
    }
    if (++slot == NST) { slot = 0; parity ^= 1; }
  }
}

template <int NST, int WARPS>
void run_tma(const double* x, double* y, int nx, int ny, int ctas_per_sm) {
  const int nstrips = (ny + 31) / 32, ngroups = (nstrips + WARPS - 1) / WARPS;
  int nseg = (148 * ctas_per_sm) / ngroups; if (nseg < 1) nseg = 1;
  const size_t smem = (size_t)WARPS * (NST * 1024 + 128);
  CK(cudaFuncSetAttribute(copy_strips_tma<NST, WARPS>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  float ms = timeit([&] { copy_strips_tma<NST, WARPS><<<ngroups * nseg, WARPS * 32, smem>>>(x, y, nx, ny, nseg); }, 20);
  CK(cudaGetLastError());
  printf("TMA bulk strips NST=%d WARPS=%d ctas/sm=%d grid=%d smem=%zu: %.3f ms  %.0f GB/s\n", NST, WARPS, ctas_per_sm,
         ngroups * nseg, smem, ms, 2.0 * nx * ny * 32 / ms / 1e6);
}

template <int OUT, int HALO, int R, int WARPS, int SCATTER = 0, int MODE = 0>
void run_strips(const double* x, double* y, int nx, int ny, int ctas_per_sm, const char* name) {
  const int nstrips = (ny + OUT - 1) / OUT, ngroups = (nstrips + WARPS - 1) / WARPS;
  int nseg = (148 * ctas_per_sm) / ngroups; if (nseg < 1) nseg = 1;
  float ms = timeit([&] { copy_strips<OUT, HALO, R, WARPS, SCATTER, MODE><<<ngroups * nseg, WARPS * 32>>>(x, y, nx, ny, nseg); }, 20);
  printf("%-44s OUT=%d HALO=%d R=%d WARPS=%d ctas/sm=%d grid=%d: %.3f ms  %.0f GB/s\n", name, OUT, HALO, R, WARPS,
         ctas_per_sm, ngroups * nseg, ms, 2.0 * nx * ny * 32 / ms / 1e6);
}

int main() {
  const int nx = 5000, ny = 5000;
  const size_t n4 = (size_t)nx * ny;
  double *x, *y;
  CK(cudaMalloc(&x, n4 * 32)); CK(cudaMalloc(&y, n4 * 32));
  CK(cudaMemset(x, 0, n4 * 32)); CK(cudaMemset(y, 0, n4 * 32));
  float ms = timeit([&] { cudaMemcpyAsync(y, x, n4 * 32, cudaMemcpyDeviceToDevice); }, 20);
  printf("cudaMemcpy D2D: %.3f ms %.0f GB/s\n", ms, 2.0 * n4 * 32 / ms / 1e6);
  ms = timeit([&] { copy_plain<4><<<148 * 8, 256>>>(x, y, n4); }, 20);
  printf("copy_plain<4> 148x8x256: %.3f ms %.0f GB/s\n", ms, 2.0 * n4 * 32 / ms / 1e6);
  ms = timeit([&] { copy_plain<4><<<148 * 2, 256>>>(x, y, n4); }, 20);
  printf("copy_plain<4> 148x2x256 (16 warps/SM): %.3f ms %.0f GB/s\n", ms, 2.0 * n4 * 32 / ms / 1e6);
  ms = timeit([&] { copy_plain<8><<<148 * 2, 256>>>(x, y, n4); }, 20);
  printf("copy_plain<8> 148x2x256 (16 warps/SM): %.3f ms %.0f GB/s\n", ms, 2.0 * n4 * 32 / ms / 1e6);
  ms = timeit([&] { copy_plain<8><<<148 * 4, 256>>>(x, y, n4); }, 20);
  printf("copy_plain<8> 148x4x256 (32 warps/SM): %.3f ms %.0f GB/s\n", ms, 2.0 * n4 * 32 / ms / 1e6);
  run_strips<32, 0, 8, 32>(x, y, nx, ny, 1, "strips aligned, 32 warps/CTA");
  run_strips<32, 0, 8, 16>(x, y, nx, ny, 2, "strips aligned, 16 warps/CTA");
  run_tma<4, 8>(x, y, nx, ny, 2);
  run_tma<8, 8>(x, y, nx, ny, 2);
  run_tma<12, 8>(x, y, nx, ny, 2);
  run_tma<8, 8>(x, y, nx, ny, 1);
  run_tma<6, 8>(x, y, nx, ny, 3);
  run_tma<16, 4>(x, y, nx, ny, 2);
  return 0;
}
