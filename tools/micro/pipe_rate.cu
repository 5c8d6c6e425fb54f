// Issue / pipe throughput of the instructions the small-head attention softmax is made of, on one B200:
//   MUFU.EX2 (f32), MUFU.EX2.F16 / .BF16 (ex2.approx.f16x2 / bf16x2 = two MUFU ops + PRMT), FFMA, FFMA2 (fma.rn.f32x2),
//   HFMA2 (f16x2 / bf16x2), F2FP pack, FMNMX, LOP3, IADD3-shift, and two MIXES (the current hot loop's instruction mix).
// Every kernel runs 8 independent dependency chains per thread, 1024 threads per SM-resident block set, and reports
// warp-instructions per clock per SM (clock64 around the loop, max over blocks).
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o pipe_rate pipe_rate.cu && ./pipe_rate
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

#define ITER 2048
#define CHAINS 8

template <int OP>
__device__ __forceinline__ void body(uint32_t (&r)[CHAINS], uint32_t a, uint32_t b) {
#pragma unroll
  for (int c = 0; c < CHAINS; ++c) {
    if (OP == 0) asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+r"(r[c]));
    if (OP == 1) asm volatile("ex2.approx.f16x2 %0, %0;" : "+r"(r[c]));
    if (OP == 2) asm volatile("ex2.approx.ftz.bf16x2 %0, %0;" : "+r"(r[c]));
    if (OP == 3) asm volatile("fma.rn.f32 %0, %0, %1, %2;" : "+r"(r[c]) : "r"(a), "r"(b));
    if (OP == 5) asm volatile("fma.rn.f16x2 %0, %0, %1, %2;" : "+r"(r[c]) : "r"(a), "r"(b));
    if (OP == 6) asm volatile("fma.rn.bf16x2 %0, %0, %1, %2;" : "+r"(r[c]) : "r"(a), "r"(b));
    if (OP == 7) asm volatile("{.reg .f32 t; mov.b32 t, %0; cvt.rn.bf16x2.f32 %0, t, t;}" : "+r"(r[c]));
    if (OP == 8) asm volatile("max.f32 %0, %0, %1;" : "+r"(r[c]) : "r"(a));
    if (OP == 9) asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(r[c]) : "r"(a), "r"(b));
    if (OP == 10) asm volatile("{.reg .b32 t; shl.b32 t, %0, 23; add.u32 %0, t, %1;}" : "+r"(r[c]) : "r"(a));
    if (OP == 11) asm volatile("{.reg .f32 t; mov.b32 t, %0; cvt.rn.f16x2.f32 %0, t, t;}" : "+r"(r[c]));
  }
}

template <int OP>
__global__ void __launch_bounds__(256) k_single(uint32_t* out, long long* clk, uint32_t a, uint32_t b) {
  uint32_t r[CHAINS];
#pragma unroll
  for (int c = 0; c < CHAINS; ++c) r[c] = threadIdx.x * 7 + c + 0x3c003c00u;
  __syncthreads();
  const long long t0 = clock64();
#pragma unroll 1
  for (int i = 0; i < ITER; ++i) body<OP>(r, a, b);
  const long long t1 = clock64();
  uint32_t s = 0;
#pragma unroll
  for (int c = 0; c < CHAINS; ++c) s ^= r[c];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
  if (threadIdx.x == 0) clk[blockIdx.x] = t1 - t0;
}

// FFMA2: 64-bit packed operands
__global__ void __launch_bounds__(256) k_ffma2(unsigned long long* out, long long* clk, unsigned long long a, unsigned long long b) {
  unsigned long long r[CHAINS];
#pragma unroll
  for (int c = 0; c < CHAINS; ++c) r[c] = threadIdx.x * 7 + c + 0x3f8000003f800000ull;
  __syncthreads();
  const long long t0 = clock64();
#pragma unroll 1
  for (int i = 0; i < ITER; ++i) {
#pragma unroll
    for (int c = 0; c < CHAINS; ++c) asm volatile("fma.rn.f32x2 %0, %0, %1, %2;" : "+l"(r[c]) : "l"(a), "l"(b));
  }
  const long long t1 = clock64();
  unsigned long long s = 0;
#pragma unroll
  for (int c = 0; c < CHAINS; ++c) s ^= r[c];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
  if (threadIdx.x == 0) clk[blockIdx.x] = t1 - t0;
}

// MIX A (today's MUFU lanes, per PAIR of elements): 1 mul.f32x2 + 2 ex2.f32 + 1 cvt.bf16x2 + 0.5 lop3
// MIX B (candidate, per pair): 1 cvt.f16x2.f32 (scaled beforehand) + 2 MUFU.EX2.F16 (+PRMT)      [ex2.approx.f16x2]
// MIX C (candidate, per pair, FMA pipe only): cvt.f16x2 + 6 fma.f16x2 + 2 integer ops
template <int MIX>
__global__ void __launch_bounds__(256) k_mix(uint32_t* out, long long* clk, unsigned long long cc, uint32_t a, uint32_t b) {
  unsigned long long x[CHAINS];
  uint32_t acc = 0;
#pragma unroll
  for (int c = 0; c < CHAINS; ++c) x[c] = 0xc0000000c0400000ull + threadIdx.x + c;
  __syncthreads();
  const long long t0 = clock64();
#pragma unroll 1
  for (int i = 0; i < ITER; ++i) {
#pragma unroll
    for (int c = 0; c < CHAINS; ++c) {
      uint32_t w;
      if (MIX == 0) {
        unsigned long long y;
        asm volatile("mul.rn.f32x2 %0, %1, %2;" : "=l"(y) : "l"(x[c]), "l"(cc));
        uint32_t lo = static_cast<uint32_t>(y), hi = static_cast<uint32_t>(y >> 32);
        asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+r"(lo));
        asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+r"(hi));
        asm volatile("{.reg .f32 l, h; mov.b32 l, %1; mov.b32 h, %2; cvt.rn.bf16x2.f32 %0, h, l;}" : "=r"(w) : "r"(lo), "r"(hi));
      } else if (MIX == 1) {
        uint32_t lo = static_cast<uint32_t>(x[c]), hi = static_cast<uint32_t>(x[c] >> 32);
        asm volatile("{.reg .f32 l, h; mov.b32 l, %1; mov.b32 h, %2; cvt.rn.f16x2.f32 %0, h, l;}" : "=r"(w) : "r"(lo), "r"(hi));
        asm volatile("ex2.approx.f16x2 %0, %0;" : "+r"(w));
      } else {
        uint32_t lo = static_cast<uint32_t>(x[c]), hi = static_cast<uint32_t>(x[c] >> 32);
        asm volatile("{.reg .f32 l, h; mov.b32 l, %1; mov.b32 h, %2; cvt.rn.f16x2.f32 %0, h, l;}" : "=r"(w) : "r"(lo), "r"(hi));
        uint32_t r = w, f, p;
        asm volatile("fma.rn.f16x2 %0, %1, %2, %3;" : "=r"(r) : "r"(w), "r"(a), "r"(b));       // round
        asm volatile("sub.rn.f16x2 %0, %1, %2;" : "=r"(f) : "r"(r), "r"(b));                   // n
        asm volatile("fma.rn.f16x2 %0, %1, %2, %3;" : "=r"(f) : "r"(w), "r"(a), "r"(f));       // frac
        asm volatile("fma.rn.f16x2 %0, %1, %2, %3;" : "=r"(p) : "r"(f), "r"(a), "r"(b));
        asm volatile("fma.rn.f16x2 %0, %1, %2, %3;" : "=r"(p) : "r"(p), "r"(f), "r"(b));
        asm volatile("fma.rn.f16x2 %0, %1, %2, %3;" : "=r"(p) : "r"(p), "r"(f), "r"(a));
        asm volatile("{.reg .b32 t; shl.b32 t, %1, 10; add.u32 %0, t, %2;}" : "=r"(w) : "r"(r), "r"(p));
      }
      acc |= w;
      x[c] += w & 1u;   // keeps the chain data-dependent without adding work on the measured pipes (1 IADD)
    }
  }
  const long long t1 = clock64();
  out[blockIdx.x * blockDim.x + threadIdx.x] = acc;
  if (threadIdx.x == 0) clk[blockIdx.x] = t1 - t0;
}

static double report(const char* name, long long* clk_d, int blocks, double warp_instr_per_thread_iter, int blocks_per_sm) {
  static long long h[4096];
  cudaMemcpy(h, clk_d, sizeof(long long) * blocks, cudaMemcpyDeviceToHost);
  long long mx = 0;
  for (int i = 0; i < blocks; ++i) mx = h[i] > mx ? h[i] : mx;
  // per SM: blocks_per_sm blocks x 8 warps, each issuing ITER*CHAINS*k warp-instructions
  const double wi = static_cast<double>(blocks_per_sm) * 8 * ITER * CHAINS * warp_instr_per_thread_iter;
  const double rate = wi / static_cast<double>(mx);
  printf("%-44s %8.3f warp-instr/clk/SM  (%6.1f lanes/clk/SM, %lld clk)\n", name, rate, rate * 32, mx);
  return rate;
}

int main() {
  int sms = 0;
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
  const int bps = 4, blocks = sms * bps;
  uint32_t* out; unsigned long long* out64; long long* clk;
  cudaMalloc(&out, sizeof(uint32_t) * blocks * 256);
  cudaMalloc(&out64, sizeof(unsigned long long) * blocks * 256);
  cudaMalloc(&clk, sizeof(long long) * blocks);
  const uint32_t fa = 0x3f7fff00u, fb = 0x3a000000u, ha = 0x3bff3bffu, hb = 0x10001000u;
#define RUN1(OP, NAME, K)                                              \
  k_single<OP><<<blocks, 256>>>(out, clk, (OP == 5 || OP == 6) ? ha : fa, (OP == 5 || OP == 6) ? hb : fb); \
  k_single<OP><<<blocks, 256>>>(out, clk, (OP == 5 || OP == 6) ? ha : fa, (OP == 5 || OP == 6) ? hb : fb); \
  cudaDeviceSynchronize();                                             \
  report(NAME, clk, blocks, K, bps);
  printf("SMs %d, %d blocks of 256 threads per SM, %d independent chains per thread\n", sms, bps, CHAINS);
  RUN1(0, "MUFU.EX2 f32 (ex2.approx.ftz.f32)", 1.0)
  RUN1(1, "ex2.approx.f16x2 (2 MUFU.EX2.F16 + PRMT)", 1.0)
  RUN1(2, "ex2.approx.ftz.bf16x2 (2 MUFU.EX2.BF16 + PRMT)", 1.0)
  RUN1(3, "FFMA (fma.rn.f32, 3 reg)", 1.0)
  k_ffma2<<<blocks, 256>>>(out64, clk, 0x3f7fff003f7fff00ull, 0x3a0000003a000000ull);
  k_ffma2<<<blocks, 256>>>(out64, clk, 0x3f7fff003f7fff00ull, 0x3a0000003a000000ull);
  cudaDeviceSynchronize();
  report("FFMA2 (fma.rn.f32x2)", clk, blocks, 1.0, bps);
  RUN1(5, "HFMA2 (fma.rn.f16x2)", 1.0)
  RUN1(6, "HFMA2.BF16 (fma.rn.bf16x2)", 1.0)
  RUN1(7, "F2FP.BF16.F32.PACK_AB (cvt.rn.bf16x2.f32)", 1.0)
  RUN1(11, "F2FP.F16.F32.PACK_AB (cvt.rn.f16x2.f32)", 1.0)
  RUN1(8, "FMNMX (max.f32)", 1.0)
  RUN1(9, "LOP3", 1.0)
  RUN1(10, "SHL + IADD (exponent patch)", 1.0)
#define RUNM(M, NAME)                                                                   \
  k_mix<M><<<blocks, 256>>>(out, clk, 0x3fb8aa3b3fb8aa3bull, M == 2 ? ha : fa, M == 2 ? hb : fb); \
  k_mix<M><<<blocks, 256>>>(out, clk, 0x3fb8aa3b3fb8aa3bull, M == 2 ? ha : fa, M == 2 ? hb : fb); \
  cudaDeviceSynchronize();                                                              \
  { double r = report(NAME, clk, blocks, 1.0, bps); printf("    -> %.1f exponentials/clk/SM\n", r * 64); }
  RUNM(0, "MIX A pairs/..: fmul2 + 2 ex2.f32 + cvt.bf16x2")
  RUNM(1, "MIX B pairs/..: cvt.f16x2 + ex2.f16x2")
  RUNM(2, "MIX C pairs/..: cvt.f16x2 + 6 HFMA2 + shl/add (FMA pipe)")
  cudaError_t e = cudaDeviceSynchronize();
  printf("status: %s\n", cudaGetErrorString(e));
  return e != cudaSuccess;
}
