// Throughput of bulk-tensor (TMA) stores and loads issued from an epilogue-like pattern: W warps per CTA (one CTA per SM), each lane 0
// issues boxes of ROWS rows x INNER bytes (bf16 2-D map, the box is dense in shared memory), double-buffered per warp
// (cp.async.bulk.wait_group.read 1).  Question answered: is the cost of the igemm epilogue's 2 KiB boxes (32 rows x 64 bytes) per
// operation, per row or per byte?
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -I../../weatherconverter_b200/csrc -o tma_store_rate tma_store_rate.cu -lcuda
#include <cstdio>
#include <cstdint>
#include <cstdlib>
#include <cuda.h>
#include <cuda_runtime.h>
#include "wc_ptx.cuh"
using namespace wc;

__device__ __forceinline__ void tma_store_2d(const void* desc, uint32_t src, int c0, int c1) {
  asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group [%0, {%2, %3}], [%1];" ::"l"(reinterpret_cast<uint64_t>(desc)), "r"(src),
               "r"(c0), "r"(c1) : "memory");
}

// mode 0: stores only; mode 1: one load + one store per iteration (residual pattern, one load in flight per warp);
// mode 2..4: the same with `mode` load buffers per warp (loads requested mode - 1 boxes ahead) and two separate store buffers
__global__ void __launch_bounds__(256, 1) k(const __grid_constant__ CUtensorMap map, long long* clk, int iters, int rows, int inner_elems,
                                            int warps, int mode, int rows_total) {
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ __align__(8) unsigned long long bars[8 * 4];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t box_bytes = static_cast<uint32_t>(rows) * inner_elems * 2;
  const uint32_t base = (smem_u32(smem) + 1023u) & ~1023u;
  const uint32_t buf = base + warp * 2 * 8192;
  if (mode >= 2) {
    // [2 store buffers | nl load buffers] of box_bytes each inside this warp's 16 KiB (box_bytes <= 2 KiB here)
    const int nl = mode;
    const uint32_t lbuf = buf + 2 * box_bytes;
    const int my_rows2 = rows_total / (gridDim.x * warps);
    const int row02 = (blockIdx.x * warps + warp) * my_rows2;
    __syncwarp();
    const long long t0 = clock64();
    auto row_of = [&](int it) { return row02 + (it * rows) % (my_rows2 - rows); };
    auto request = [&](int it) {
      const int slot = it % nl;
      if (lane == 0) {
        mbar_arrive_expect_tx(smem_u32(&bars[warp * 4 + slot]), box_bytes);
        asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];" ::"r"(lbuf + slot * box_bytes),
                     "l"(reinterpret_cast<uint64_t>(&map)), "r"(smem_u32(&bars[warp * 4 + slot])), "r"(0), "r"(row_of(it)) : "memory");
      }
    };
    for (int i = 0; i < nl && i < iters; ++i) request(i);
    for (int it = 0; it < iters; ++it) {
      const int slot = it % nl;
      mbar_wait(smem_u32(&bars[warp * 4 + slot]), (it / nl) & 1);
      // "use" the residual: copy the row to the store buffer through registers
      const uint32_t sb = buf + (it & 1) * box_bytes;
      if (lane == 0) tma_store_wait_read<1>();
      __syncwarp();
      for (uint32_t o = lane * 16; o < box_bytes; o += 32 * 16) {
        uint32_t v[4];
        asm volatile("ld.shared.v4.b32 {%0,%1,%2,%3}, [%4];" : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]) : "r"(lbuf + slot * box_bytes + o) : "memory");
        asm volatile("st.shared.v4.b32 [%0], {%1,%2,%3,%4};" ::"r"(sb + o), "r"(v[0]), "r"(v[1]), "r"(v[2]), "r"(v[3]) : "memory");
      }
      fence_proxy_async_smem();
      __syncwarp();
      if (lane == 0) {
        tma_store_2d(&map, sb, 0, row_of(it));
        tma_store_commit();
      }
      if (it + nl < iters) request(it + nl);
    }
    if (lane == 0) tma_store_wait_read<0>();
    __syncwarp();
    if (threadIdx.x == 0) clk[blockIdx.x] = clock64() - t0;
    return;
  }
  if (lane == 0 && warp < warps) { for (int i = 0; i < 4; ++i) mbar_init(smem_u32(&bars[warp * 4 + i]), 1); fence_barrier_init(); }
  __syncthreads();
  if (warp >= warps) return;
  // this warp's row range of the global tensor (disjoint per warp and CTA, cycled)
  const int my_rows = rows_total / (gridDim.x * warps);
  const int row0 = (blockIdx.x * warps + warp) * my_rows;
  uint32_t phase = 0;
  const long long t0 = clock64();
  for (int it = 0; it < iters; ++it) {
    const int r = row0 + (it * rows) % (my_rows - rows);
    const uint32_t b = buf + (it & 1) * 8192;
    if (lane == 0) {
      tma_store_wait_read<1>();
      if (mode == 1) {
        mbar_arrive_expect_tx(smem_u32(&bars[warp]), box_bytes);
        asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];" ::"r"(b),
                     "l"(reinterpret_cast<uint64_t>(&map)), "r"(smem_u32(&bars[warp])), "r"(0), "r"(r) : "memory");
      }
    }
    __syncwarp();
    if (mode == 1) { mbar_wait(smem_u32(&bars[warp]), phase); phase ^= 1u; }
    fence_proxy_async_smem();
    __syncwarp();
    if (lane == 0) {
      tma_store_2d(&map, b, 0, r);
      tma_store_commit();
    }
  }
  if (lane == 0) tma_store_wait_read<0>();
  __syncwarp();
  if (threadIdx.x == 0) clk[blockIdx.x] = clock64() - t0;
}

int main() {
  cudaSetDevice(0);
  int sms = 0;
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
  const size_t rows_total = 1 << 22;            // 4M rows
  void* g = nullptr;
  cudaMalloc(&g, rows_total * 256);             // up to 256 bytes per row
  cudaMemset(g, 0, rows_total * 256);
  long long* clk;
  cudaMallocManaged(&clk, sms * sizeof(long long));
  cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, 8 * 2 * 8192 + 1024);
  struct Cfg { int rows, inner_bytes, warps, mode; };
  const Cfg cfgs[] = {{32, 64, 8, 0}, {32, 128, 8, 0}, {16, 128, 8, 0}, {64, 64, 8, 0}, {128, 64, 8, 0}, {32, 64, 4, 0}, {32, 64, 1, 0},
                      {32, 64, 8, 1}, {32, 128, 8, 1}, {128, 64, 2, 1}, {32, 64, 8, 2}, {32, 64, 8, 3}, {32, 64, 8, 4}};
  printf("rows x inner B, warps, mode (0 store, 1 load+store): clk per box per SM, bytes/clk/SM (each direction), TB/s chip\n");
  for (const Cfg& c : cfgs) {
    CUtensorMap map;
    cuuint64_t gdim[2] = {static_cast<cuuint64_t>(c.inner_bytes / 2), rows_total};
    cuuint64_t gstride[1] = {static_cast<cuuint64_t>(c.inner_bytes)};
    cuuint32_t box[2] = {static_cast<cuuint32_t>(c.inner_bytes / 2), static_cast<cuuint32_t>(c.rows)};
    cuuint32_t es[2] = {1, 1};
    CUresult r = cuTensorMapEncodeTiled(&map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, g, gdim, gstride, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                                        CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) { printf("encode failed %d\n", (int)r); return 1; }
    const int iters = 2000;
    for (int rep = 0; rep < 2; ++rep) {
      k<<<sms, 256, 8 * 2 * 8192 + 1024>>>(map, clk, iters, c.rows, c.inner_bytes / 2, c.warps, c.mode, static_cast<int>(rows_total));
      cudaError_t e = cudaDeviceSynchronize();
      if (e != cudaSuccess) { printf("kernel failed: %s\n", cudaGetErrorString(e)); return 1; }
    }
    long long mx = 0;
    for (int i = 0; i < sms; ++i) mx = clk[i] > mx ? clk[i] : mx;
    const double boxes = static_cast<double>(iters) * c.warps;
    const double bytes = boxes * c.rows * c.inner_bytes;
    printf("%4d x %3d B, %d warps, mode %d: %7.1f clk/box/SM  %6.1f B/clk/SM  %5.2f TB/s\n", c.rows, c.inner_bytes, c.warps, c.mode, mx / boxes,
           bytes / mx, bytes / mx * sms * 1.965e9 / 1e12);
  }
  return 0;
}
