// Small kernels of the denoising-loss training step (diffusion_model/train_ddpm.py:95-114): MSE loss + gradient,
// weight gradients of the two 3-channel boundary convolutions, time-embedding MLP forward (saving activations) and
// backward, fused Adam.  The tensor-core parts (data / weight gradients of the wide convolutions, attention backward)
// live in igemm.cu, wgrad.cu and attention_bwd.cu; GroupNorm backward in norm.cu.
#include "wc_host.h"
#include "wc_ptx.cuh"

namespace wc {

namespace {

// ------------------------------------------------------------------------------------------------ MSE
// loss = mean((pred - target)^2) (nn.MSELoss, train_ddpm.py:107); dpred = 2 (pred - target) / numel * grad_scale.
__global__ void mse_partial_kernel(const float* __restrict__ pred, const float* __restrict__ target, float* __restrict__ dpred,
                                   size_t n, float gscale, double* __restrict__ part) {
  pdl_prologue();
  __shared__ double sh[32];
  double acc = 0.0;
  for (size_t i = blockIdx.x * static_cast<size_t>(blockDim.x) + threadIdx.x; i < n; i += static_cast<size_t>(gridDim.x) * blockDim.x) {
    const float d = pred[i] - target[i];
    dpred[i] = d * gscale;
    acc += static_cast<double>(d) * d;
  }
  for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
  if ((threadIdx.x & 31) == 0) sh[threadIdx.x >> 5] = acc;
  __syncthreads();
  if (threadIdx.x == 0) {
    double s = 0.0;
    for (int i = 0; i < (blockDim.x >> 5); ++i) s += sh[i];
    part[blockIdx.x] = s;
  }
}
__global__ void mse_finish_kernel(const double* __restrict__ part, int nparts, size_t n, float* __restrict__ loss) {
  pdl_prologue();
  if (threadIdx.x == 0 && blockIdx.x == 0) {
    double s = 0.0;
    for (int i = 0; i < nparts; ++i) s += part[i];
    *loss = static_cast<float>(s / static_cast<double>(n));
  }
}

// ------------------------------------------------------------------------------------------------ boundary wgrad
// res[wc][nc][tap] = sum_p wide[p][wc] * narrow[nc][p + SIGN*off(tap)],  wide: NHWC bf16 with 64 channels,
// narrow: NCHW fp32 with 3 channels, 3x3 taps, zero padding.  SIGN = +1: conv_in (wide = d conv_in output, narrow =
// the input image; result index [co=wc][ci=nc]); SIGN = -1: conv_out (wide = its input activation, narrow = dpred;
// result index [co=nc][ci=wc]).  Also wsum[wc] = sum_p wide[p][wc] and nsum[nc] = sum_p narrow[nc][p].
constexpr int kBwThreads = 256;
constexpr int kBwTile = 64;  // pixels of one image row per tile
// Tile-based: a block stages 64 pixels x 64 channels of `wide` (as fp32) and the 3 x 3 rows x 66 pixels halo of `narrow`
// in shared memory, then thread (wc, q) accumulates its 16 pixels from shared memory (wide: conflict-free, narrow:
// broadcast).  Partials per block are reduced in a fixed order by the finish kernel (deterministic).
__global__ void __launch_bounds__(kBwThreads)
boundary_wgrad_kernel(const __nv_bfloat16* __restrict__ wide, int ldw, const float* __restrict__ narrow, int B, int H, int W,
                      int sign, float* __restrict__ part /*[blocks][64*28 + 4]*/) {
  pdl_prologue();
  __shared__ float s_wide[kBwTile][65];
  __shared__ float s_nar[3][3][kBwTile + 2];
  __shared__ float red[kBwThreads / 64][64 * 28 + 4];
  const int wc = threadIdx.x & 63, q = threadIdx.x >> 6;  // q in 0..3
  const size_t plane = static_cast<size_t>(H) * W;
  const int tiles_x = (W + kBwTile - 1) / kBwTile;
  const long ntiles = static_cast<long>(B) * H * tiles_x;
  float acc[27];
#pragma unroll
  for (int i = 0; i < 27; ++i) acc[i] = 0.f;
  float wsum = 0.f, ns0 = 0.f, ns1 = 0.f, ns2 = 0.f;
  for (long t = blockIdx.x; t < ntiles; t += gridDim.x) {
    const int x0 = static_cast<int>(t % tiles_x) * kBwTile;
    const int y = static_cast<int>((t / tiles_x) % H), b = static_cast<int>(t / (static_cast<long>(tiles_x) * H));
    __syncthreads();
    // wide tile: 64 pixels x 64 channels, 8 channels (16 bytes) per thread and step
    for (int i = threadIdx.x; i < kBwTile * 8; i += kBwThreads) {
      const int px = i >> 3, v = i & 7;
      float f[8];
#pragma unroll
      for (int j = 0; j < 8; ++j) f[j] = 0.f;
      if (x0 + px < W) {
        const uint4 u = __ldg(reinterpret_cast<const uint4*>(wide + ((static_cast<size_t>(b) * H + y) * W + x0 + px) * ldw + v * 8));
        float2 tt;
        tt = unpack_bf16(u.x); f[0] = tt.x; f[1] = tt.y;
        tt = unpack_bf16(u.y); f[2] = tt.x; f[3] = tt.y;
        tt = unpack_bf16(u.z); f[4] = tt.x; f[5] = tt.y;
        tt = unpack_bf16(u.w); f[6] = tt.x; f[7] = tt.y;
      }
#pragma unroll
      for (int j = 0; j < 8; ++j) s_wide[px][v * 8 + j] = f[j];
    }
    // narrow halo: rows y-1..y+1, columns x0-1..x0+64, 3 channels (zero outside the image)
    for (int i = threadIdx.x; i < 9 * (kBwTile + 2); i += kBwThreads) {
      const int cx = i % (kBwTile + 2), r = (i / (kBwTile + 2)) % 3, nc = i / (3 * (kBwTile + 2));
      const int yy = y + r - 1, xx = x0 + cx - 1;
      float v = 0.f;
      if (yy >= 0 && yy < H && xx >= 0 && xx < W) v = __ldg(narrow + (static_cast<size_t>(b) * 3 + nc) * plane + static_cast<size_t>(yy) * W + xx);
      s_nar[nc][r][cx] = v;
    }
    __syncthreads();
#pragma unroll 4
    for (int px = q; px < kBwTile; px += 4) {
      if (x0 + px >= W) break;
      const float wv = s_wide[px][wc];
      wsum += wv;
#pragma unroll
      for (int ky = 0; ky < 3; ++ky)
#pragma unroll
        for (int kx = 0; kx < 3; ++kx) {
          // narrow pixel of tap (ky,kx): p + sign*(k-1)  ->  halo index (1 + sign*(k-1))
          const int r = 1 + sign * (ky - 1), c = px + 1 + sign * (kx - 1);
          acc[0 * 9 + ky * 3 + kx] = fmaf(wv, s_nar[0][r][c], acc[0 * 9 + ky * 3 + kx]);
          acc[1 * 9 + ky * 3 + kx] = fmaf(wv, s_nar[1][r][c], acc[1 * 9 + ky * 3 + kx]);
          acc[2 * 9 + ky * 3 + kx] = fmaf(wv, s_nar[2][r][c], acc[2 * 9 + ky * 3 + kx]);
        }
      if (wc == 0) { ns0 += s_nar[0][1][px + 1]; ns1 += s_nar[1][1][px + 1]; ns2 += s_nar[2][1][px + 1]; }
    }
  }
#pragma unroll
  for (int i = 0; i < 27; ++i) red[q][wc * 28 + i] = acc[i];
  red[q][wc * 28 + 27] = wsum;
  if (wc == 0) { red[q][64 * 28] = ns0; red[q][64 * 28 + 1] = ns1; red[q][64 * 28 + 2] = ns2; }
  __syncthreads();
  float* out = part + static_cast<size_t>(blockIdx.x) * (64 * 28 + 4);
  for (int i = threadIdx.x; i < 64 * 28 + 3; i += blockDim.x) out[i] = (red[0][i] + red[1][i]) + (red[2][i] + red[3][i]);
}

__global__ void boundary_wgrad_finish_kernel(const float* __restrict__ part, int nblocks, int sign, float* __restrict__ dw,
                                             float* __restrict__ dbias) {
  pdl_prologue();
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= 64 * 28 + 3) return;
  float s = 0.f;
  for (int b = 0; b < nblocks; ++b) s += part[static_cast<size_t>(b) * (64 * 28 + 4) + i];
  if (i < 64 * 28) {
    const int wc = i / 28, r = i % 28;
    if (r < 27) {
      const int nc = r / 9, tap = r % 9;
      if (sign > 0) dw[(wc * 3 + nc) * 9 + tap] = s;       // conv_in.weight [64][3][3][3]
      else dw[(nc * 64 + wc) * 9 + tap] = s;               // conv_out.weight [3][64][3][3]
    } else if (sign > 0) {
      dbias[wc] = s;                                       // conv_in.bias [64]
    }
  } else if (sign < 0) {
    dbias[i - 64 * 28] = s;                                // conv_out.bias [3]
  }
}

// w'[ci][co][ky][kx] = w[co][ci][2-ky][2-kx]: the conv_out data gradient as a 3 -> 64 convolution of dpred
__global__ void flip_transpose_3x3_kernel(const float* __restrict__ w, float* __restrict__ wt, int Co, int Ci) {
  pdl_prologue();
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= Co * Ci * 9) return;
  const int tap = i % 9, ci = (i / 9) % Ci, co = i / (9 * Ci);
  wt[(ci * Co + co) * 9 + (8 - tap)] = w[i];
}

// ------------------------------------------------------------------------------------------------ time embedding
// forward, saving what the backward needs: emb [B][dim], h1 (pre-SiLU) [B][dim], temb (pre-SiLU) [B][dim], silu(temb)
__global__ void temb_mlp_train_kernel(const long long* __restrict__ t, int dim, const float* __restrict__ factor,
                                      const float* __restrict__ w1,
                                      const float* __restrict__ b1, const float* __restrict__ w2, const float* __restrict__ b2,
                                      float* __restrict__ emb_out, float* __restrict__ h1_out, float* __restrict__ temb,
                                      float* __restrict__ temb_silu) {
  pdl_prologue();
  extern __shared__ float sh[];
  float* emb = sh;
  float* s1 = sh + dim;
  const int b = blockIdx.x, half = dim / 2;
  const float tv = static_cast<float>(t[b]);
  for (int j = threadIdx.x; j < half; j += blockDim.x) {
    const float a = tv / time_factor(factor, j, half);
    emb[j] = sinf(a);
    emb[j + half] = cosf(a);
  }
  __syncthreads();
  for (int n = threadIdx.x; n < dim; n += blockDim.x) {
    emb_out[static_cast<size_t>(b) * dim + n] = emb[n];
    float acc = b1[n];
    for (int k = 0; k < dim; ++k) acc = fmaf(w1[static_cast<size_t>(n) * dim + k], emb[k], acc);
    h1_out[static_cast<size_t>(b) * dim + n] = acc;
    s1[n] = acc / (1.f + expf(-acc));
  }
  __syncthreads();
  for (int n = threadIdx.x; n < dim; n += blockDim.x) {
    float acc = b2[n];
    for (int k = 0; k < dim; ++k) acc = fmaf(w2[static_cast<size_t>(n) * dim + k], s1[k], acc);
    temb[static_cast<size_t>(b) * dim + n] = acc;
    temb_silu[static_cast<size_t>(b) * dim + n] = acc / (1.f + expf(-acc));
  }
}

__device__ __forceinline__ float silu_grad(float z) {
  const float s = 1.f / (1.f + expf(-z));
  return s * (1.f + z * (1.f - s));
}

// dtemb[b][k] = silu'(temb[b][k]) * sum_j dtproj[b][j] * Wcat[j][k]     (grid B, block dim threads)
__global__ void temb_bwd_proj_kernel(const float* __restrict__ dtproj, int total, const float* __restrict__ wcat, int dim,
                                     const float* __restrict__ temb, float* __restrict__ dtemb) {
  pdl_prologue();
  const int b = blockIdx.x, k = threadIdx.x;
  if (k >= dim) return;
  float acc = 0.f;
  const float* dp = dtproj + static_cast<size_t>(b) * total;
  for (int j = 0; j < total; ++j) acc = fmaf(dp[j], wcat[static_cast<size_t>(j) * dim + k], acc);
  dtemb[static_cast<size_t>(b) * dim + k] = acc * silu_grad(temb[static_cast<size_t>(b) * dim + k]);
}

// dW[j][k] = sum_b dout[b][j] * in[b][k], db[j] = sum_b dout[b][j] (grid: rows j, block: dim threads).  The destination
// of row j is looked up in a table of (row offset, weight ptr, bias ptr) segments so that the concatenated
// t_emb_layers write straight into their separate gradient tensors.
struct LinSeg { int row0, rows; float* dw; float* db; };
constexpr int kMaxLinSegs = 40;
struct LinSegs { int n; LinSeg s[kMaxLinSegs]; };
__global__ void linear_wgrad_rows_kernel(const float* __restrict__ dout, int ldd, const float* __restrict__ in, int dim, int B,
                                         const __grid_constant__ LinSegs segs) {
  pdl_prologue();
  const int j = blockIdx.x, k = threadIdx.x;
  int si = 0;
  while (si + 1 < segs.n && j >= segs.s[si + 1].row0) ++si;
  const LinSeg sg = segs.s[si];
  const int r = j - sg.row0;
  if (r >= sg.rows) return;
  if (k < dim) {
    float acc = 0.f;
    for (int b = 0; b < B; ++b) acc = fmaf(dout[static_cast<size_t>(b) * ldd + j], in[static_cast<size_t>(b) * dim + k], acc);
    sg.dw[static_cast<size_t>(r) * dim + k] = acc;
  }
  if (k == 0 && sg.db) {
    float acc = 0.f;
    for (int b = 0; b < B; ++b) acc += dout[static_cast<size_t>(b) * ldd + j];
    sg.db[r] = acc;
  }
}

// ds1[b][k] = silu'(h1[b][k]) * sum_n dtemb[b][n] * W2[n][k]
__global__ void temb_bwd_hidden_kernel(const float* __restrict__ dtemb, const float* __restrict__ w2, int dim,
                                       const float* __restrict__ h1, float* __restrict__ dh1) {
  pdl_prologue();
  const int b = blockIdx.x, k = threadIdx.x;
  if (k >= dim) return;
  float acc = 0.f;
  for (int n = 0; n < dim; ++n) acc = fmaf(dtemb[static_cast<size_t>(b) * dim + n], w2[static_cast<size_t>(n) * dim + k], acc);
  dh1[static_cast<size_t>(b) * dim + k] = acc * silu_grad(h1[static_cast<size_t>(b) * dim + k]);
}

__global__ void silu_rows_kernel(const float* __restrict__ x, float* __restrict__ y, size_t n) {
  pdl_prologue();
  const size_t i = blockIdx.x * static_cast<size_t>(blockDim.x) + threadIdx.x;
  if (i < n) { const float v = x[i]; y[i] = v / (1.f + expf(-v)); }
}

__global__ void add_vectors_kernel(const float* __restrict__ a, const float* __restrict__ b, float* __restrict__ o, int n) {
  pdl_prologue();
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) o[i] = a[i] + b[i];
}

// ------------------------------------------------------------------------------------------------ Adam
// torch.optim.Adam(lr, betas, eps), no weight decay / amsgrad (train_ddpm.py:151): one launch over the flat buffers.
__global__ void adam_kernel(float* __restrict__ p, const float* __restrict__ g, float* __restrict__ m, float* __restrict__ v,
                            size_t n, float lr, float beta1, float beta2, float eps, float bc1, float bc2_sqrt, float gscale) {
  pdl_prologue();
  const float step_size = lr / bc1;
  for (size_t i = blockIdx.x * static_cast<size_t>(blockDim.x) + threadIdx.x; i < n; i += static_cast<size_t>(gridDim.x) * blockDim.x) {
    const float gr = g[i] * gscale;
    const float mi = __fadd_rn(m[i], __fmul_rn(__fsub_rn(gr, m[i]), 1.f - beta1));         // exp_avg.lerp_(grad, 1-beta1)
    const float vi = __fadd_rn(__fmul_rn(v[i], beta2), __fmul_rn(__fmul_rn(1.f - beta2, gr), gr));  // mul_(b2).addcmul_(g,g,1-b2)
    const float denom = __fadd_rn(__fdiv_rn(__fsqrt_rn(vi), bc2_sqrt), eps);
    p[i] = __fadd_rn(p[i], __fmul_rn(-step_size, __fdiv_rn(mi, denom)));
    m[i] = mi;
    v[i] = vi;
  }
}

// ------------------------------------------------------------------------------------------------ attention D
// D[bh][tok] = sum_d dO[b][tok][h*hd+d] * O[b][tok][h*hd+d]  (the softmax-backward row term); one warp per (b, tok)
__global__ void attn_rowdot_kernel(const __nv_bfloat16* __restrict__ o, const __nv_bfloat16* __restrict__ d_o, int ldo, int ldd,
                                   int B, int ntok, int heads, int hd, float* __restrict__ D) {
  pdl_prologue();
  const size_t row = (blockIdx.x * static_cast<size_t>(blockDim.x) + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (row >= static_cast<size_t>(B) * ntok) return;
  const int b = static_cast<int>(row / ntok), tok = static_cast<int>(row % ntok);
  for (int h = 0; h < heads; ++h) {
    float acc = 0.f;
    for (int d = lane * 2; d < hd; d += 64) {
      const float2 a = unpack_bf16(*reinterpret_cast<const uint32_t*>(o + row * ldo + h * hd + d));
      const float2 g = unpack_bf16(*reinterpret_cast<const uint32_t*>(d_o + row * ldd + h * hd + d));
      acc += a.x * g.x + a.y * g.y;
    }
#pragma unroll
    for (int s = 16; s > 0; s >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, s);
    if (lane == 0) D[(static_cast<size_t>(b) * heads + h) * ntok + tok] = acc;
  }
}

}  // namespace

// scratch: >= 1024 doubles
int mse_loss_grad(const float* pred, const float* target, float* dpred, size_t n, float grad_scale, float* loss, void* scratch,
                  cudaStream_t st) {
  const int blocks = 1024;
  double* part = static_cast<double*>(scratch);
  ProfScope prof(kProfScheduler, st, 12.0 * n);
  launch_k(mse_partial_kernel, blocks, 256, 0, st, pred, target, dpred, n, 2.0f * grad_scale / static_cast<float>(n), part);
  WC_LAUNCH_CHECK();
  launch_k(mse_finish_kernel, 1, 32, 0, st, part, blocks, n, loss);
  WC_LAUNCH_CHECK();
  return 0;
}

size_t boundary_wgrad_scratch_bytes() { return static_cast<size_t>(4 * 148) * (64 * 28 + 4) * sizeof(float); }
// sign +1: conv_in (dw [64][3][3][3], dbias [64]); sign -1: conv_out (dw [3][64][3][3], dbias [3])
int boundary_wgrad(const __nv_bfloat16* wide, int ldw, const float* narrow, int B, int H, int W, int sign, float* dw,
                   float* dbias, void* scratch, cudaStream_t st) {
  const int blocks = 4 * 148;
  float* part = static_cast<float*>(scratch);
  ProfScope prof(kProfBoundaryConv, st, 2.0 * 27 * 64 * static_cast<double>(B) * H * W);
  launch_k(boundary_wgrad_kernel, blocks, kBwThreads, 0, st, wide, ldw, narrow, B, H, W, sign, part);
  WC_LAUNCH_CHECK();
  launch_k(boundary_wgrad_finish_kernel, (64 * 28 + 3 + 127) / 128, 128, 0, st, part, blocks, sign, dw, dbias);
  WC_LAUNCH_CHECK();
  return 0;
}

int add_vectors(const float* a, const float* b, float* o, int n, cudaStream_t st) {
  launch_k(add_vectors_kernel, (n + 255) / 256, 256, 0, st, a, b, o, n);
  WC_LAUNCH_CHECK();
  return 0;
}

int flip_transpose_3x3(const float* w, float* wt, int Co, int Ci, cudaStream_t st) {
  launch_k(flip_transpose_3x3_kernel, (Co * Ci * 9 + 127) / 128, 128, 0, st, w, wt, Co, Ci);
  WC_LAUNCH_CHECK();
  return 0;
}

int temb_mlp_train(const long long* t, int Bt, int dim, const float* w1, const float* b1, const float* w2, const float* b2,
                   float* emb, float* h1, float* temb, float* temb_silu, cudaStream_t st) {
  launch_k(temb_mlp_train_kernel, Bt, 128, 2 * dim * sizeof(float), st, t, dim, time_factor_table(dim / 2), w1, b1, w2, b2, emb, h1, temb, temb_silu);
  WC_LAUNCH_CHECK();
  return 0;
}

// Backward of: temb = W2 silu(W1 emb + b1) + b2 ; tproj = Wcat silu(temb) + bcat.
// segs: destination of the concatenated projection's gradient rows (t_emb_layers.*.1.{weight,bias}).
int temb_backward(const float* dtproj, int total, int B, int dim, const float* wcat, const float* w2, const float* emb,
                  const float* h1, const float* temb, const float* temb_silu, int nsegs, const int* seg_row0,
                  const int* seg_rows, float* const* seg_dw, float* const* seg_db, float* dw1, float* db1, float* dw2,
                  float* db2, float* scratch /* 3*B*dim floats */, cudaStream_t st) {
  WC_REQUIRE(dim <= 1024 && nsegs <= kMaxLinSegs, "temb_backward: unsupported sizes");
  float* dtemb = scratch;
  float* dh1 = scratch + static_cast<size_t>(B) * dim;
  float* s1 = scratch + 2 * static_cast<size_t>(B) * dim;
  LinSegs segs;
  segs.n = nsegs;
  for (int i = 0; i < nsegs; ++i) segs.s[i] = {seg_row0[i], seg_rows[i], seg_dw[i], seg_db[i]};
  launch_k(linear_wgrad_rows_kernel, total, dim, 0, st, dtproj, total, temb_silu, dim, B, segs);
  WC_LAUNCH_CHECK();
  launch_k(temb_bwd_proj_kernel, B, dim, 0, st, dtproj, total, wcat, dim, temb, dtemb);
  WC_LAUNCH_CHECK();
  const size_t n = static_cast<size_t>(B) * dim;
  launch_k(silu_rows_kernel, static_cast<int>((n + 255) / 256), 256, 0, st, h1, s1, n);
  WC_LAUNCH_CHECK();
  LinSegs one;
  one.n = 1;
  one.s[0] = {0, dim, dw2, db2};
  launch_k(linear_wgrad_rows_kernel, dim, dim, 0, st, dtemb, dim, s1, dim, B, one);
  WC_LAUNCH_CHECK();
  launch_k(temb_bwd_hidden_kernel, B, dim, 0, st, dtemb, w2, dim, h1, dh1);
  WC_LAUNCH_CHECK();
  one.s[0] = {0, dim, dw1, db1};
  launch_k(linear_wgrad_rows_kernel, dim, dim, 0, st, dh1, dim, emb, dim, B, one);
  WC_LAUNCH_CHECK();
  return 0;
}

int adam_step(float* p, const float* g, float* m, float* v, size_t n, float lr, float beta1, float beta2, float eps, int step,
              float grad_scale, cudaStream_t st) {
  const double bc1 = 1.0 - pow(static_cast<double>(beta1), step);
  const double bc2 = 1.0 - pow(static_cast<double>(beta2), step);
  ProfScope prof(kProfScheduler, st, 28.0 * n);
  launch_k(adam_kernel, 4 * num_sms(), 256, 0, st, p, g, m, v, n, lr, beta1, beta2, eps, static_cast<float>(bc1),
                                             static_cast<float>(sqrt(bc2)), grad_scale);
  WC_LAUNCH_CHECK();
  return 0;
}

int attn_rowdot(const __nv_bfloat16* o, const __nv_bfloat16* d_o, int ldo, int ldd, int B, int ntok, int heads, int hd, float* D,
                cudaStream_t st) {
  WC_REQUIRE(hd % 2 == 0, "attn_rowdot: head_dim must be even");
  const size_t rows = static_cast<size_t>(B) * ntok;
  ProfScope prof(kProfOther, st, 4.0 * rows * heads * hd);
  launch_k(attn_rowdot_kernel, static_cast<int>((rows * 32 + 255) / 256), 256, 0, st, o, d_o, ldo, ldd, B, ntok, heads, hd, D);
  WC_LAUNCH_CHECK();
  return 0;
}

}  // namespace wc
