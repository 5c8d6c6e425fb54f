// Weight gradient of a convolution / linear layer on tcgen05 (sm_100a):
//
//   dW[n, tap, c] = sum_{pixel} dY[pixel, n] * X_tap[pixel + (dy,dx), c]
//
// i.e. a GEMM whose contraction runs over PIXELS.  Both operands are NHWC activations (channels contiguous), so both
// are MN-major UMMA operands: the same TMA boxes the forward implicit GEMM uses (64 channels x a tile of pixels,
// 128-byte swizzle, zero fill outside the image) are consumed through shared-memory descriptors with the MN-major bit
// set in the instruction descriptor -- no transposes anywhere.
//   M = 128 output channels (two 64-channel atoms of dY), N = up to four 64-channel "B atoms", each atom = one
//   (tap, input-channel block) pair of X, K = 64 pixels per pipeline stage (four tcgen05.mma of K = 16).
// Work unit = (M tile, group of <= 4 B atoms, pixel split); persistent CTAs walk the units; fp32 partial sums per
// split go to a workspace and a second kernel reduces the splits in a fixed order (deterministic) and scatters into
// the PyTorch weight layout.  Reference op replaced: autograd's weight gradient of every nn.Conv2d /
// nn.ConvTranspose2d / nn.Linear of the UNet (loss.backward() at diffusion_model/train_ddpm.py:108).
#pragma once
#include <vector>

#include "wc_host.h"

namespace wc {

constexpr int kWgMaxTaps = 20;
constexpr int kWgMaxMaps = 5;
constexpr int kWgPixels = 64;  // pixels (K) per pipeline stage
constexpr size_t kWgradPartialBytes = 96ull << 20;  // workspace every wgrad launch may use for its split partials

struct WgradTapDev {
  int8_t map, dy, dx, pad_;
  int32_t nkb;   // 64-channel blocks of this tap
  int32_t koff;  // column offset of the tap in the packed [N][Ktot] layout
};

struct WgradArgs {
  int B, H, W;
  int tb, th, tw;  // pixel tile, tb*th*tw == 64
  int N;           // output channels (rows of dW)
  int ntaps, natoms, ngroups, nsplit, ptiles, m_tiles, ktot;
  WgradTapDev taps[kWgMaxTaps];
  float* partial;  // [nsplit][m_tiles*128][ktot]
};

struct WgradMaps {
  CUtensorMap dy;
  CUtensorMap x[kWgMaxMaps];
};

// Where the gradient of one tap goes: element (n, c) of the tap -> dst[((d0*D1 + d1)*KH + ky)*KW + kx] with
// (d0, d1) = (n, c0 + c) for Conv2d / Linear weights and (c0 + c, n) for ConvTranspose2d weights (transpose = 1).
struct WgradDst {
  float* dst;
  int koff, C, KH, KW, ky, kx, transpose, D1, c0;
};

struct WgradUnpackArgs {
  int N, ktot, nsplit, mpad, ntaps;
  const float* partial;
  WgradDst taps[kWgMaxTaps];
};

struct WgradPlan {
  WgradMaps maps;
  WgradArgs args;
  WgradUnpackArgs unpack;
  int grid = 0;
  double flops = 0;
};

// A strided spatial view of an NHWC activation: element (b, y, x, c) = act(b, y*ys + y0, x*xs + x0, c).
struct ViewSpec {
  Act act;
  int y0 = 0, ys = 1, x0 = 0, xs = 1;
};

struct WgradTap {
  ViewSpec x;      // input of the tap (view dims must equal the dY grid)
  int dy = 0, dx = 0;
  float* dst = nullptr;  // fp32 gradient tensor (PyTorch layout)
  int KH = 1, KW = 1, ky = 0, kx = 0, transpose = 0, D1 = 0, c0 = 0;
};

// dY: [B,H,W,N] view (the grid the pixel tiles walk).  partial: workspace of kWgradPartialBytes.
int build_wgrad(WgradPlan* plan, const ViewSpec& dY, int B, int H, int W, int N, const std::vector<WgradTap>& taps,
                float* partial);
int wgrad_launch(const WgradPlan& plan, cudaStream_t st);

struct WgradOp {
  std::vector<WgradPlan> plans;
  double flops = 0;
  int run(cudaStream_t st) const {
    for (const auto& p : plans)
      if (int e = wgrad_launch(p, st)) return e;
    return 0;
  }
};

// Weight gradient of y = conv(x, W[Cout,Cin,K,K]) (+ fused 1x1 conv of x2 with W2[Cout,Cin2,1,1]) given dy on the
// output grid; same geometry rules as build_conv (stride 1 "same" convs with dilation, or stride 2 via phase views).
int build_conv_wgrad(WgradOp* op, const Act& x, const Act& dy, int K, int stride, int pad, int dil, float* dw,
                     const Act* x2, float* dw2, float* partial);
// Weight gradient of y = conv_transpose2d(x, W[Cin,Cout,K,K], stride 2, pad) given dy [B,2H,2W,Cout].
int build_convT_wgrad(WgradOp* op, const Act& x, const Act& dy, int K, int pad, float* dw, float* partial);

}  // namespace wc
