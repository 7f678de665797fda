// Micro-benchmark: cycles per tcgen05.mma (kind::f16, M = 128, K = 16) issued back to back by one thread, as a function
// of N, operand swizzle and accumulator reuse.  nvcc -gencode arch=compute_100a,code=sm_100a -O3 -I../../hfnet_slam_b200/csrc
#include <cstdio>
#include <cuda_runtime.h>
#include "tc.cuh"

__global__ void __launch_bounds__(128) rate_kernel(int N, int reps, int mode, int sw64, long long* out) {
  extern __shared__ uint8_t raw[];
  uint8_t* smem = raw + ((1024u - (tc::smem_u32(raw) & 1023u)) & 1023u);
  __shared__ uint64_t bar;
  __shared__ uint32_t slot;
  const int tid = threadIdx.x;
  for (int i = tid; i < 96 * 1024 / 4; i += 128) reinterpret_cast<uint32_t*>(smem)[i] = 0x3c003c00u;   // ones
  if (tid == 0) {
    tc::mbar_init(&bar, 1);
    tc::fence_barrier_init();
  }
  if (tid < 32) tc::tmem_alloc(&slot, 512);
  tc::fence_proxy_async();
  tc::fence_before_sync();
  __syncthreads();
  tc::fence_after_sync();
  const uint32_t tm = slot;
  if (tid == 0) {
    const uint32_t sa = tc::smem_u32(smem), sb = sa + 32 * 1024;
    const uint64_t da = sw64 ? tc::make_sdesc_sw64(sa) : tc::make_sdesc_sw128(sa);
    const uint64_t db = sw64 ? tc::make_sdesc_sw64(sb) : tc::make_sdesc_sw128(sb);
    const uint32_t idesc = tc::make_idesc_f16(N);
    for (int warm = 0; warm < 2; ++warm) {
      const long long t0 = clock64();
      for (int r = 0; r < reps; ++r) {
        // mode 0: one accumulator, dependent chain; mode 1: two accumulators alternating; mode 2: fresh D every time (no accumulate)
        const uint32_t d = tm + (mode == 1 ? (uint32_t)((r & 1) * 256) : 0u);
        tc::umma_f16(d, da + (uint64_t)(2 * (r & 3)), db + (uint64_t)(2 * (r & 3)), idesc, mode == 2 ? 0u : (r > 1 ? 1u : 0u));
      }
      const long long t1 = clock64();
      tc::umma_commit(&bar);
      tc::mbar_wait(&bar, (uint32_t)warm);
      const long long t2 = clock64();
      out[0] = t1 - t0;
      out[1] = t2 - t0;
    }
  }
  tc::fence_before_sync();
  __syncthreads();
  if (tid < 32) tc::tmem_dealloc(tm, 512);
}

int main() {
  long long* d;
  cudaMalloc(&d, 16);
  cudaFuncSetAttribute(rate_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024);
  const int reps = 64;
  for (int sw64 = 0; sw64 < 2; ++sw64)
    for (int mode = 0; mode < 3; ++mode)
      for (int N : {16, 32, 64, 96, 128, 160, 256}) {
        rate_kernel<<<1, 128, 100 * 1024>>>(N, reps, mode, sw64, d);
        long long h[2];
        cudaMemcpy(h, d, 16, cudaMemcpyDeviceToHost);
        cudaError_t e = cudaGetLastError();
        printf("sw%s mode %d N %3d: issue %6.1f cyc/mma, complete %6.1f cyc/mma%s\n", sw64 ? "64 " : "128", mode, N,
               (double)h[0] / reps, (double)h[1] / reps, e == cudaSuccess ? "" : cudaGetErrorString(e));
      }
  return 0;
}
