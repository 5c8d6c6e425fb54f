// Micro-benchmark: clocks per tcgen05.mma (cta_group::1, kind::f16, M = 128, K = 16) as a function of N, of the A operand
// source (shared memory descriptor vs tensor memory) and of the number of back-to-back instructions.  One CTA per SM, one
// issuing thread, operands resident (no TMA in the timed loop).  Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3
// -I weatherconverter_b200/csrc tools/micro/mma_rate.cu -o tools/micro/mma_rate
#include <cstdio>
#include <cstdlib>
#include "wc_ptx.cuh"
using namespace wc;

template <int N, bool TS>
__global__ void __launch_bounds__(128, 1) mma_rate_kernel(long long* out, int iters) {
  extern __shared__ uint8_t smem_raw[];
  __shared__ uint32_t slot;
  __shared__ __align__(8) unsigned long long bar;
  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  // A: 128 rows x 64 bf16 (16 KB, SW128 K-major); B: N rows x 64 bf16
  for (int i = threadIdx.x; i < (16384 + N * 128) / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(smem_raw + (base - smem_u32(smem_raw)))[i] = 0x3c003c00u;
  if (threadIdx.x < 32) {
    tmem_alloc(smem_u32(&slot), 512);
    tmem_relinquish();
  }
  if (threadIdx.x == 0) { mbar_init(smem_u32(&bar), 1); fence_barrier_init(); }
  fence_proxy_async_smem();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = slot;
  // Warp 0 runs the loop in uniform control flow and one elected lane issues (as the library kernels do since the end of
  // round 1).  The table in profiles/r1_mma_issue_rate.txt was measured with the issue inside `if (threadIdx.x == 0)`: there the
  // compiler wraps every UTCHMMA in an ELECT / BRA.U.ANY loop, and the "45-clk floor" for N <= 64 is that wrapper, not the tensor
  // pipe (the igemm trace shows 37 clk per 128x64x16 MMA with the elected issue).  Define MMA_RATE_DIVERGENT to reproduce it.
#ifdef MMA_RATE_DIVERGENT
  if (threadIdx.x == 0) {
#else
  if (__shfl_sync(0xffffffffu, threadIdx.x >> 5, 0) == 0 && elect_one_sync()) {
#endif
    const uint32_t idesc = umma_idesc_bf16(128, N);
    const uint64_t da = umma_smem_desc(base, 128, 1024), db = umma_smem_desc(base + 16384, 128, 1024);
    const uint32_t hi = umma_desc_hi(da), alo = umma_desc_lo(da), blo = umma_desc_lo(db);
    uint32_t phase = 0;
    for (int rep = 0; rep < 2; ++rep) {   // rep 0 warms up
      const long long t0 = clock64();
      for (int i = 0; i < iters; ++i) {
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          if (TS) umma_bf16_ts(tmem, tmem + 256 + 8 * k, umma_desc_join(blo + 2u * k, hi), idesc, 1u);
          else umma_bf16(tmem, umma_desc_join(alo + 2u * k, hi), umma_desc_join(blo + 2u * k, hi), idesc, 1u);
        }
      }
      umma_commit(smem_u32(&bar));
      mbar_wait(smem_u32(&bar), phase);
      phase ^= 1u;
      const long long t1 = clock64();
      if (rep == 1 && blockIdx.x == 0) out[0] = t1 - t0;
    }
  }
  tc_fence_before();
  __syncthreads();
  if (threadIdx.x < 32) { tc_fence_after(); tmem_dealloc(tmem, 512); }
}

template <int N, bool TS>
void run(long long* d_out, int iters, int grid) {
  const size_t smem = 16384 + N * 128 + 1024;
  cudaFuncSetAttribute(mma_rate_kernel<N, TS>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  mma_rate_kernel<N, TS><<<grid, 128, smem>>>(d_out, iters);
  cudaError_t e = cudaDeviceSynchronize();
  long long clk = 0;
  cudaMemcpy(&clk, d_out, 8, cudaMemcpyDeviceToHost);
  printf("N=%3d A=%s grid=%3d: %7.1f clk per tcgen05.mma (128xNx16)  [%s]  floor %d\n", N, TS ? "tmem" : "smem", grid,
         (double)clk / (4.0 * iters), cudaGetErrorString(e), N / 2);
}

int main() {
  long long* d_out;
  cudaMalloc(&d_out, 8);
  const int iters = 2000;
  for (int grid : {1, 148}) {
    run<16, false>(d_out, iters, grid);  run<32, false>(d_out, iters, grid);  run<64, false>(d_out, iters, grid);
    run<128, false>(d_out, iters, grid); run<256, false>(d_out, iters, grid);
    run<16, true>(d_out, iters, grid);   run<64, true>(d_out, iters, grid);   run<128, true>(d_out, iters, grid);  run<256, true>(d_out, iters, grid);
  }
  return 0;
}
