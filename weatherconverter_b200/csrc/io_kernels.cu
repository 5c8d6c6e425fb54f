// Image / label edges of the sampling loop as byte-exact GPU kernels:
//   * sample_ddpm.py:47-51     clamp(-1,1) -> (x+1)/2 -> torchvision make_grid -> ToPILImage (mul(255).byte())
//   * sample_integrated.py:32-37  postprocess: x*std + mean -> *255 -> clamp(0,255) -> uint8
//   * seg_model/datasets/acdc.py:135-138 encode_target (labelIds -> trainIds LUT) after the NEAREST resize + centre crop
//     of seg_model/inference.py:75-82 (ExtResize(just_label) + ExtCenterCrop)
//   * translation.py:138-145   Resize(BILINEAR) + CenterCrop + ToTensor + x*2-1 of the source image; the resize follows
//     Pillow's two-pass fixed-point resampling (Resample.c: 8-bit coefficients with 22 fractional bits, horizontal pass
//     then vertical pass, intermediate uint8) with coefficient tables computed on the host in double precision
//   * seg_model/inference.py:79-80  ExtToTensor + ExtNormalize(mean, std)
// All results are bit-identical to the reference's CPU path (tests/test_gpu_io.py).
#include "wc_host.h"

namespace wc {

namespace {

__device__ __forceinline__ uint8_t to_byte_trunc(float v) {  // Tensor.byte() of a value already inside [0, 255]
  return static_cast<uint8_t>(static_cast<int>(v));
}

// out: HWC uint8 grid [Hg][Wg][3] (the PIL image ToPILImage builds from the CHW float grid)
__global__ void ddpm_grid_u8_kernel(const float* __restrict__ x, uint8_t* __restrict__ out, int B, int H, int W, int xmaps, int ymaps,
                                    int pad, int Hg, int Wg) {
  pdl_prologue();
  const size_t total = static_cast<size_t>(Hg) * Wg;
  for (size_t i = blockIdx.x * static_cast<size_t>(blockDim.x) + threadIdx.x; i < total; i += static_cast<size_t>(gridDim.x) * blockDim.x) {
    const int gx = static_cast<int>(i % Wg), gy = static_cast<int>(i / Wg);
    uint8_t r = 0, g = 0, b = 0;  // pad_value 0
    int k = -1, py = 0, px = 0;
    if (B == 1 && pad == 0) { k = 0; py = gy; px = gx; }
    else {
      const int cy = gy / (H + pad), cx = gx / (W + pad);
      py = gy - cy * (H + pad) - pad; px = gx - cx * (W + pad) - pad;
      if (cy < ymaps && cx < xmaps && py >= 0 && px >= 0 && cy * xmaps + cx < B) k = cy * xmaps + cx;
    }
    if (k >= 0) {
      const float* p = x + (static_cast<size_t>(k) * 3) * H * W + static_cast<size_t>(py) * W + px;
      float v[3];
#pragma unroll
      for (int c = 0; c < 3; ++c) {
        float t = p[static_cast<size_t>(c) * H * W];
        t = fminf(fmaxf(t, -1.f), 1.f);            // torch.clamp(xt, -1, 1)
        t = __fdiv_rn(__fadd_rn(t, 1.f), 2.f);     // (ims + 1) / 2
        v[c] = __fmul_rn(t, 255.f);                // pic.mul(255)
      }
      r = to_byte_trunc(v[0]); g = to_byte_trunc(v[1]); b = to_byte_trunc(v[2]);
    }
    out[i * 3 + 0] = r; out[i * 3 + 1] = g; out[i * 3 + 2] = b;
  }
}

// NCHW float -> NCHW uint8: (x*std + mean) * 255, clamp(0, 255), truncate
__global__ void postprocess_u8_kernel(const float* __restrict__ x, uint8_t* __restrict__ out, size_t n, int HW, float m0, float m1, float m2,
                                      float s0, float s1, float s2) {
  pdl_prologue();
  for (size_t i = blockIdx.x * static_cast<size_t>(blockDim.x) + threadIdx.x; i < n; i += static_cast<size_t>(gridDim.x) * blockDim.x) {
    const int c = static_cast<int>((i / HW) % 3);
    const float mean = c == 0 ? m0 : (c == 1 ? m1 : m2), std = c == 0 ? s0 : (c == 1 ? s1 : s2);
    float v = __fadd_rn(__fmul_rn(x[i], std), mean);
    v = __fmul_rn(v, 255.f);
    v = fminf(fmaxf(v, 0.f), 255.f);
    out[i] = to_byte_trunc(v);
  }
}

// label ids (uint8 [Hs][Ws]) -> NEAREST resize via host index tables -> centre crop -> LUT -> int64 [Hc][Wc]
__global__ void label_encode_kernel(const uint8_t* __restrict__ lab, int Ws, const int* __restrict__ ytab, const int* __restrict__ xtab,
                                    int top, int left, int Hc, int Wc, const long long* __restrict__ lut, int nlut,
                                    long long* __restrict__ out) {
  pdl_prologue();
  const int total = Hc * Wc;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < total; i += gridDim.x * blockDim.x) {
    const int y = i / Wc, x = i % Wc;
    const int sy = ytab[top + y], sx = xtab[left + x];
    const int id = lab[static_cast<size_t>(sy) * Ws + sx];
    out[i] = id < nlut ? lut[id] : 255;
  }
}

// One pass of Pillow's 8-bit resampling along x: out[y][xx][c] = clip8((sum_k in[y][xmin+k][c] * kk[xx][k] + 2^21) >> 22)
__global__ void resample_h_kernel(const uint8_t* __restrict__ in, uint8_t* __restrict__ out, int H, int Win, int Wout, int C,
                                  const int* __restrict__ bounds, const int* __restrict__ kk, int ksize) {
  pdl_prologue();
  const size_t total = static_cast<size_t>(H) * Wout * C;
  for (size_t i = blockIdx.x * static_cast<size_t>(blockDim.x) + threadIdx.x; i < total; i += static_cast<size_t>(gridDim.x) * blockDim.x) {
    const int c = static_cast<int>(i % C), xx = static_cast<int>((i / C) % Wout);
    const size_t y = i / (static_cast<size_t>(C) * Wout);
    const int xmin = bounds[2 * xx], xmax = bounds[2 * xx + 1];
    int ss = 1 << 21;
    for (int k = 0; k < xmax; ++k) ss += static_cast<int>(in[(y * Win + xmin + k) * C + c]) * kk[xx * ksize + k];
    ss >>= 22;
    out[i] = static_cast<uint8_t>(ss < 0 ? 0 : (ss > 255 ? 255 : ss));
  }
}
// ... and along y
__global__ void resample_v_kernel(const uint8_t* __restrict__ in, uint8_t* __restrict__ out, int Hout, int W, int C,
                                  const int* __restrict__ bounds, const int* __restrict__ kk, int ksize) {
  pdl_prologue();
  const size_t total = static_cast<size_t>(Hout) * W * C;
  for (size_t i = blockIdx.x * static_cast<size_t>(blockDim.x) + threadIdx.x; i < total; i += static_cast<size_t>(gridDim.x) * blockDim.x) {
    const size_t col = i % (static_cast<size_t>(W) * C);
    const int yy = static_cast<int>(i / (static_cast<size_t>(W) * C));
    const int ymin = bounds[2 * yy], ymax = bounds[2 * yy + 1];
    int ss = 1 << 21;
    for (int k = 0; k < ymax; ++k) ss += static_cast<int>(in[static_cast<size_t>(ymin + k) * W * C + col]) * kk[yy * ksize + k];
    ss >>= 22;
    out[i] = static_cast<uint8_t>(ss < 0 ? 0 : (ss > 255 ? 255 : ss));
  }
}

// HWC uint8 -> centre crop -> CHW float: to_tensor (v/255) then either x*2-1 (mode 0, translation.py:143) or
// (x - mean)/std (mode 1, ExtNormalize)
__global__ void u8_to_tensor_kernel(const uint8_t* __restrict__ in, int Win, int top, int left, int Hc, int Wc, int mode, float m0, float m1,
                                    float m2, float s0, float s1, float s2, float* __restrict__ out) {
  pdl_prologue();
  const size_t total = static_cast<size_t>(3) * Hc * Wc;
  for (size_t i = blockIdx.x * static_cast<size_t>(blockDim.x) + threadIdx.x; i < total; i += static_cast<size_t>(gridDim.x) * blockDim.x) {
    const int x = static_cast<int>(i % Wc), y = static_cast<int>((i / Wc) % Hc), c = static_cast<int>(i / (static_cast<size_t>(Wc) * Hc));
    const float v = __fdiv_rn(static_cast<float>(in[(static_cast<size_t>(top + y) * Win + left + x) * 3 + c]), 255.f);
    float o;
    if (mode == 0) {
      o = __fsub_rn(__fmul_rn(v, 2.0f), 1.0f);
    } else {
      const float mean = c == 0 ? m0 : (c == 1 ? m1 : m2), std = c == 0 ? s0 : (c == 1 ? s1 : s2);
      o = __fdiv_rn(__fsub_rn(v, mean), std);
    }
    out[i] = o;
  }
}

inline int nblocks(size_t n) { return static_cast<int>(n == 0 ? 1 : (n + 255) / 256 > 148 * 32 ? 148 * 32 : (n + 255) / 256); }

}  // namespace

int ddpm_grid_u8(const float* x, uint8_t* out, int B, int H, int W, int nrow, int pad, cudaStream_t st) {
  WC_REQUIRE(B >= 1 && nrow >= 1, "ddpm_grid_u8: bad arguments");
  int xmaps = nrow < B ? nrow : B, ymaps = (B + xmaps - 1) / xmaps, Hg, Wg;
  if (B == 1) { pad = 0; Hg = H; Wg = W; }   // make_grid returns a single image unpadded
  else { Hg = (H + pad) * ymaps + pad; Wg = (W + pad) * xmaps + pad; }
  ProfScope prof(kProfOther, st, 12.0 * B * H * W + 3.0 * Hg * Wg);
  launch_k(ddpm_grid_u8_kernel, nblocks(static_cast<size_t>(Hg) * Wg), 256, 0, st, x, out, B, H, W, xmaps, ymaps, pad, Hg, Wg);
  WC_LAUNCH_CHECK();
  return 0;
}
int postprocess_u8(const float* x, uint8_t* out, int B, int H, int W, const float* mean3, const float* std3, cudaStream_t st) {
  const size_t n = static_cast<size_t>(B) * 3 * H * W;
  ProfScope prof(kProfOther, st, 5.0 * n);
  launch_k(postprocess_u8_kernel, nblocks(n), 256, 0, st, x, out, n, H * W, mean3[0], mean3[1], mean3[2], std3[0], std3[1], std3[2]);
  WC_LAUNCH_CHECK();
  return 0;
}
int label_encode(const uint8_t* lab, int Ws, const int* ytab, const int* xtab, int top, int left, int Hc, int Wc, const long long* lut,
                 int nlut, long long* out, cudaStream_t st) {
  launch_k(label_encode_kernel, nblocks(static_cast<size_t>(Hc) * Wc), 256, 0, st, lab, Ws, ytab, xtab, top, left, Hc, Wc, lut, nlut, out);
  WC_LAUNCH_CHECK();
  return 0;
}
int resample_u8(const uint8_t* in, uint8_t* tmp, uint8_t* out, int Hin, int Win, int Hout, int Wout, int C, const int* bounds_h,
                const int* kk_h, int ksize_h, const int* bounds_v, const int* kk_v, int ksize_v, cudaStream_t st) {
  launch_k(resample_h_kernel, nblocks(static_cast<size_t>(Hin) * Wout * C), 256, 0, st, in, tmp, Hin, Win, Wout, C, bounds_h, kk_h, ksize_h);
  WC_LAUNCH_CHECK();
  launch_k(resample_v_kernel, nblocks(static_cast<size_t>(Hout) * Wout * C), 256, 0, st, tmp, out, Hout, Wout, C, bounds_v, kk_v, ksize_v);
  WC_LAUNCH_CHECK();
  return 0;
}
int u8_to_tensor(const uint8_t* in, int Win, int top, int left, int Hc, int Wc, int mode, const float* mean3, const float* std3, float* out,
                 cudaStream_t st) {
  const float one[3] = {1.f, 1.f, 1.f}, zero[3] = {0.f, 0.f, 0.f};
  const float* m = mean3 ? mean3 : zero;
  const float* s = std3 ? std3 : one;
  launch_k(u8_to_tensor_kernel, nblocks(static_cast<size_t>(3) * Hc * Wc), 256, 0, st, in, Win, top, left, Hc, Wc, mode, m[0], m[1], m[2], s[0], s[1],
                                                                                s[2], out);
  WC_LAUNCH_CHECK();
  return 0;
}

}  // namespace wc
