// GroupNorm(8 groups, eps 1e-5, affine) with optional fused SiLU on NHWC bf16 activations.
// Reference ops: nn.GroupNorm(8, C) [+ nn.SiLU] at unet_base.py:89-90,101-102,107,155-156,483-484.
// Two bandwidth-bound passes: (1) per-(sample, slab) partial sums, (2) deterministic combine + normalise.
// Statistics are identical for the 4-D [B,C,H,W] and 3-D [B,C,HW] uses (biased variance over C/8 x HW).
#include "wc_host.h"
#include "wc_ptx.cuh"

namespace wc {

namespace {

constexpr int kGroups = 8;

// Each thread owns one 8-channel vector column (fixed group) and strides over pixels.
__global__ void gn_stats_kernel(const __nv_bfloat16* __restrict__ x, int HW, int C, int ld, int nsplit,
                                float2* __restrict__ partial) {
  extern __shared__ float2 stash[];
  const int vpp = C / 8, gv = (C / kGroups) / 8;  // vectors per pixel, vectors per group
  const int rows = blockDim.x / vpp;
  const int v = threadIdx.x % vpp, r = threadIdx.x / vpp;
  const int b = blockIdx.y, split = blockIdx.x;
  const int p0 = static_cast<int>(static_cast<long long>(HW) * split / nsplit);
  const int p1 = static_cast<int>(static_cast<long long>(HW) * (split + 1) / nsplit);
  const __nv_bfloat16* xb = x + static_cast<size_t>(b) * HW * ld + v * 8;
  float s = 0.f, ss = 0.f;
  for (int p = p0 + r; p < p1; p += rows) {
    const uint4 u = __ldg(reinterpret_cast<const uint4*>(xb + static_cast<size_t>(p) * ld));
    const float2 a = unpack_bf16(u.x), bq = unpack_bf16(u.y), c = unpack_bf16(u.z), d = unpack_bf16(u.w);
    s += (a.x + a.y) + (bq.x + bq.y) + (c.x + c.y) + (d.x + d.y);
    ss += (a.x * a.x + a.y * a.y) + (bq.x * bq.x + bq.y * bq.y) + (c.x * c.x + c.y * c.y) + (d.x * d.x + d.y * d.y);
  }
  stash[threadIdx.x] = make_float2(s, ss);
  __syncthreads();
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int nwarps = blockDim.x >> 5;
  for (int g = warp; g < kGroups; g += nwarps) {  // one warp reduces one group in a fixed order (deterministic)
    const int cnt = rows * gv;
    float as = 0.f, ass = 0.f;
    for (int i = lane; i < cnt; i += 32) {
      const int tid = (i / gv) * vpp + g * gv + (i % gv);
      const float2 t = stash[tid];
      as += t.x; ass += t.y;
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      as += __shfl_xor_sync(0xffffffffu, as, o);
      ass += __shfl_xor_sync(0xffffffffu, ass, o);
    }
    if (lane == 0) partial[(static_cast<size_t>(b) * nsplit + split) * kGroups + g] = make_float2(as, ass);
  }
}

__global__ void gn_apply_kernel(const __nv_bfloat16* __restrict__ x, __nv_bfloat16* __restrict__ y, int HW, int C,
                                int ld, int ldy, int nsplit_stats, int nsplit, const float2* __restrict__ partial,
                                const float* __restrict__ gamma, const float* __restrict__ beta, float eps, int silu) {
  __shared__ float s_mean[kGroups], s_rstd[kGroups];
  const int b = blockIdx.y, split = blockIdx.x;
  if (threadIdx.x < kGroups) {
    double s = 0.0, ss = 0.0;
    for (int i = 0; i < nsplit_stats; ++i) {
      const float2 t = partial[(static_cast<size_t>(b) * nsplit_stats + i) * kGroups + threadIdx.x];
      s += t.x; ss += t.y;
    }
    const double n = static_cast<double>(HW) * (C / kGroups);
    const double mean = s / n;
    double var = ss / n - mean * mean;
    if (var < 0.0) var = 0.0;
    s_mean[threadIdx.x] = static_cast<float>(mean);
    s_rstd[threadIdx.x] = static_cast<float>(1.0 / sqrt(var + static_cast<double>(eps)));
  }
  __syncthreads();
  const int vpp = C / 8, cpg = C / kGroups;
  const int rows = blockDim.x / vpp;
  const int v = threadIdx.x % vpp, r = threadIdx.x / vpp;
  const int g = (v * 8) / cpg;
  const float mean = s_mean[g], rstd = s_rstd[g];
  float ga[8], be[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    ga[j] = gamma[v * 8 + j] * rstd;
    be[j] = beta[v * 8 + j] - mean * ga[j];
  }
  const int p0 = static_cast<int>(static_cast<long long>(HW) * split / nsplit);
  const int p1 = static_cast<int>(static_cast<long long>(HW) * (split + 1) / nsplit);
  const __nv_bfloat16* xb = x + static_cast<size_t>(b) * HW * ld + v * 8;
  __nv_bfloat16* yb = y + static_cast<size_t>(b) * HW * ldy + v * 8;
  for (int p = p0 + r; p < p1; p += rows) {
    const uint4 u = __ldg(reinterpret_cast<const uint4*>(xb + static_cast<size_t>(p) * ld));
    float f[8];
    float2 t;
    t = unpack_bf16(u.x); f[0] = t.x; f[1] = t.y;
    t = unpack_bf16(u.y); f[2] = t.x; f[3] = t.y;
    t = unpack_bf16(u.z); f[4] = t.x; f[5] = t.y;
    t = unpack_bf16(u.w); f[6] = t.x; f[7] = t.y;
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      float o = fmaf(f[j], ga[j], be[j]);
      if (silu) o = __fdividef(o, 1.f + __expf(-o));
      f[j] = o;
    }
    uint4 w;
    w.x = pack_bf16(f[0], f[1]); w.y = pack_bf16(f[2], f[3]); w.z = pack_bf16(f[4], f[5]); w.w = pack_bf16(f[6], f[7]);
    *reinterpret_cast<uint4*>(yb + static_cast<size_t>(p) * ldy) = w;
  }
}

}  // namespace

size_t groupnorm_workspace_bytes(int B) { return static_cast<size_t>(B) * 64 * kGroups * sizeof(float2); }

// x: [B,HW,C] bf16 with pixel stride ld; y likewise with ldy; workspace >= groupnorm_workspace_bytes(B).
int groupnorm_silu(const __nv_bfloat16* x, __nv_bfloat16* y, int B, int HW, int C, int ld, int ldy, const float* gamma,
                   const float* beta, float eps, int silu, void* workspace, cudaStream_t st) {
  WC_REQUIRE(C % 64 == 0 && C / 8 <= 256, "GroupNorm(8) kernel needs C % 64 == 0 and C <= 2048");
  WC_REQUIRE(ld % 8 == 0 && ldy % 8 == 0, "pixel strides must be multiples of 8 elements");
  const int vpp = C / 8;
  const int threads = (256 / vpp) * vpp;
  int nsplit = (4 * num_sms() + B - 1) / B;
  const int max_split = (HW + 31) / 32;
  if (nsplit > max_split) nsplit = max_split;
  if (nsplit > 64) nsplit = 64;
  if (nsplit < 1) nsplit = 1;
  float2* partial = reinterpret_cast<float2*>(workspace);
  ProfScope prof(kProfGroupNorm, st, 6.0 * B * static_cast<double>(HW) * C);  // algorithmic bytes: 2 reads + 1 write, bf16
  gn_stats_kernel<<<dim3(nsplit, B), threads, threads * sizeof(float2), st>>>(x, HW, C, ld, nsplit, partial);
  WC_LAUNCH_CHECK();
  int asplit = (8 * num_sms() + B - 1) / B;
  if (asplit > max_split) asplit = max_split;
  if (asplit < 1) asplit = 1;
  gn_apply_kernel<<<dim3(asplit, B), threads, 0, st>>>(x, y, HW, C, ld, ldy, nsplit, asplit, partial, gamma, beta, eps,
                                                       silu);
  WC_LAUNCH_CHECK();
  return 0;
}

}  // namespace wc
