// TMEM read (tcgen05.ld 32x32b.x32 / .x16) and write (tcgen05.st) throughput per SM as a function of the number of warps
// reading (1, 2 or 4 warps per lane quadrant = 4 / 8 / 16 warps), and the latency of one dependent load.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -I../../weatherconverter_b200/csrc -o tmem_rate tmem_rate.cu
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
#include "wc_ptx.cuh"
using namespace wc;

template <int MODE>   // 0: ld x32 throughput (8 in flight), 1: ld x32 dependent (latency), 2: st x32 throughput, 3: ld x16 throughput
__global__ void __launch_bounds__(512, 1) k(long long* clk, uint32_t* sink, int iters) {
  __shared__ uint32_t slot;
  const int warp = threadIdx.x >> 5;
  if (warp == 0) { tmem_alloc(smem_u32(&slot), 512); tmem_relinquish(); }
  tc_fence_before(); __syncthreads(); tc_fence_after();
  const uint32_t base = slot + ((static_cast<uint32_t>((warp & 3) * 32)) << 16);
  uint32_t acc = 0;
  uint32_t r[32];
#pragma unroll
  for (int i = 0; i < 32; ++i) r[i] = threadIdx.x + i;
  if (MODE == 2) { tmem_st16(base, reinterpret_cast<uint32_t(&)[16]>(r)); tmem_wait_st(); }
  __syncthreads();
  const long long t0 = clock64();
  for (int it = 0; it < iters; ++it) {
    if (MODE == 0) {
#pragma unroll
      for (int c = 0; c < 8; ++c) { tmem_ld32(base + ((it * 8 + c) * 32) % 480, r); }
      tmem_wait_ld();
      acc ^= r[0] ^ r[31];
    } else if (MODE == 1) {
      tmem_ld32(base + (acc & 31), r);
      tmem_wait_ld();
      acc ^= r[3] & 1;
    } else if (MODE == 2) {
#pragma unroll
      for (int c = 0; c < 8; ++c) tmem_st16(base + ((it * 8 + c) * 16) % 496, reinterpret_cast<uint32_t(&)[16]>(r));
      tmem_wait_st();
    } else {
#pragma unroll
      for (int c = 0; c < 8; ++c) { tmem_ld16(base + ((it * 8 + c) * 16) % 496, reinterpret_cast<uint32_t(&)[16]>(r)); }
      tmem_wait_ld();
      acc ^= r[0] ^ r[15];
    }
  }
  const long long t1 = clock64();
  sink[blockIdx.x * blockDim.x + threadIdx.x] = acc;
  if (threadIdx.x == 0) clk[blockIdx.x] = t1 - t0;
  tc_fence_before(); __syncthreads();
  if (warp == 0) { tc_fence_after(); tmem_dealloc(slot, 512); }
}

int main() {
  long long* clk; uint32_t* sink;
  cudaMalloc(&clk, 8 * 148); cudaMalloc(&sink, 4 * 148 * 512);
  const int iters = 2000;
  for (int warps : {4, 8, 16}) {
    long long h[148];
    auto run = [&](auto kern, const char* name, double bytes_per_warp_iter) {
      kern<<<148, warps * 32>>>(clk, sink, iters);
      kern<<<148, warps * 32>>>(clk, sink, iters);
      cudaError_t e = cudaDeviceSynchronize();
      cudaMemcpy(h, clk, sizeof(h), cudaMemcpyDeviceToHost);
      long long mx = 0; for (long long v : h) mx = v > mx ? v : mx;
      printf("%2d warps  %-34s %8.1f B/clk/SM   %7.1f clk per iteration  (%s)\n", warps, name, warps * bytes_per_warp_iter * iters / double(mx),
             double(mx) / iters, cudaGetErrorString(e));
    };
    run(k<0>, "tcgen05.ld 32x32b.x32, 8 in flight", 8 * 4096.0);
    run(k<3>, "tcgen05.ld 32x32b.x16, 8 in flight", 8 * 2048.0);
    run(k<1>, "tcgen05.ld x32 dependent (latency)", 4096.0);
    run(k<2>, "tcgen05.st 32x32b.x16, 8 in flight", 8 * 2048.0);
  }
  return 0;
}
