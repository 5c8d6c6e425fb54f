// Swift-SRGAN generator forward as a static launch plan (reference: srgan_model/models.py:5-92,
// srgan_model/inference.py:35-39).  Depthwise 3x3 + pointwise pairs are composed into dense 3x3 convolutions
// (W[o,c,ky,kx] = pw[o,c]*dw[c,ky,kx], eval BatchNorm folded) and run on the tcgen05 implicit-GEMM path with
// PReLU / residual epilogues; PixelShuffle(2) is four output-phase convolutions writing straight into the 2x
// larger tensor; the first (3-channel, 9x9) and last (depthwise 9x9 + 64->3 + tanh) layers are CUDA-core kernels
// that also convert from / to the reference's NCHW fp32 layout.
#include "plan.cuh"
#include "../../include/wc_b200.h"

namespace wc {
int conv_small_cin(const float* x, const float* w, const float* bias, const float* scale, const float* shift,
                   __nv_bfloat16* y, int B, int Cin, int H, int W, int Cout, int K, int stride, int pad, int ldy,
                   int relu, cudaStream_t st, const float* prelu = nullptr);
int bn_fold(const float* g, const float* b, const float* mean, const float* var, float eps, float* scale, float* shift,
            int n, int n_pad, cudaStream_t st);
int compose_sep(const float* dw, const float* pw, const float* dwb, const float* pwb, float* w, float* bias, int Co, int Ci,
                int KK, cudaStream_t st);
int srgan_initial(const float* x, const float* dw, const float* dwb, const float* pw, const float* pwb, const float* slope,
                  __nv_bfloat16* y, int B, int H, int W, int ldy, cudaStream_t st);
int srgan_final(const __nv_bfloat16* x, const float* dw, const float* dwb, const float* pw, const float* pwb, float* y, int B,
                int H, int W, int ldx, cudaStream_t st);
int gather_stride(const float* src, float* dst, int n, int mul, int off, cudaStream_t st);
int phase_major_rows(const float* src, float* dst, int cb, int len, int rep_src, cudaStream_t st);
int srgan_final_compose(const float* dw, const float* dwb, const float* pw, const float* pwb, float* wq, float* bias3, cudaStream_t st);
int srgan_final_combine(const float* t, const float* bias3, float* y, int B, int H, int W, cudaStream_t st);
}  // namespace wc

struct wc_srgan {
  int num_blocks = 16, upscale = 4, nc = 64;
  wc::ParamTable params;
  std::unique_ptr<wc::DeviceArena> arena;
  int B = 0, H = 0, W = 0;
  void* ws = nullptr;
  size_t ws_bytes = 0;
  wc::OpList ops;
  double flops = 0;
  size_t ws_needed = 0;
  const float* x_in = nullptr;
  float* y_out = nullptr;
};

namespace wc {
namespace {

int build(wc_srgan* net, bool dry, void* ws, size_t ws_bytes, cudaStream_t st) {
  const int B = net->B, H = net->H, W = net->W, NCH = net->nc;
  Bump bump = dry ? Bump::dry() : Bump(ws, ws_bytes);
  DeviceArena* arena = net->arena.get();
  int err = 0;
  auto P = [&](const std::string& n) { return net->params.get(n, &err); };
  auto Popt = [&](const std::string& n) -> const float* {
    auto it = net->params.ptr.find(n);
    return it == net->params.ptr.end() ? nullptr : it->second;
  };
  auto push = [&](std::function<int(cudaStream_t)> f) { if (!dry) net->ops.push_back(std::move(f)); };
  auto fa = [&](size_t n) { return static_cast<float*>(arena->alloc(n * sizeof(float))); };

  // composed dense weights of a SeperableConv2d (+ optional folded BN): returns (weights, bias/shift, scale)
  struct Composed { float* w; float* bias; float* scale; };
  auto compose = [&](const std::string& p, int ci, int co, int K, const std::string& bn) -> Composed {
    Composed c{nullptr, nullptr, nullptr};
    c.w = fa(static_cast<size_t>(co) * ci * K * K);
    c.bias = fa(co);
    const float *dw = P(p + ".depthwise.weight"), *pw = P(p + ".pointwise.weight");
    const float *dwb = Popt(p + ".depthwise.bias"), *pwb = Popt(p + ".pointwise.bias");
    if (!c.w || !c.bias || err) { if (!err) err = 1; return c; }
    if (int e = compose_sep(dw, pw, dwb, pwb, c.w, c.bias, co, ci, K * K, st)) { err = e; return c; }
    if (!bn.empty()) {   // conv has no bias when followed by BN (models.py:30): bias := BN shift, weights scaled
      c.scale = fa(co);
      float* shift = fa(co);
      const float *g = P(bn + ".weight"), *b = P(bn + ".bias"), *m = P(bn + ".running_mean"), *v = P(bn + ".running_var");
      if (!c.scale || !shift || err) { if (!err) err = 1; return c; }
      if (int e = bn_fold(g, b, m, v, 1e-5f, c.scale, shift, co, co, st)) { err = e; return c; }
      c.bias = shift;
    }
    return c;
  };
  auto conv3 = [&](const Act& x, const Composed& c, int co, const float* prelu, const Act* res, const Act& out) {
    if (dry || err) return;
    WeightSrc w; w.w = c.w; w.d0 = co; w.d1 = x.C; w.KH = w.KW = 3; w.scale = c.scale;
    ConvGeom g; g.K = 3; g.pad = 1;
    Epilogue ep; ep.bias = c.bias; ep.prelu = prelu; ep.res = res;
    OutSpec os; os.mode = kOutNHWC; os.out = out;
    auto op = std::make_shared<ConvOp>();
    if (int e = build_conv(op.get(), arena, x, w, g, co, nullptr, nullptr, ep, os, st)) { err = e; return; }
    // algorithmic FLOPs = the reference's separable form (depthwise 3x3 + pointwise), not the composed dense conv
    op->flops = 2.0 * x.pixels() * (9.0 * x.C + static_cast<double>(x.C) * co);
    for (auto& pl : op->plans) pl.flops = op->flops / op->plans.size();
    net->flops += op->flops;
    net->ops.push_back([op](cudaStream_t s) { return op->run(s); });
  };

  // ---- initial: SeperableConv2d(3->64, k9) + PReLU, NCHW fp32 in
  Act initial = make_act(bump, B, H, W, NCH);
  if (!dry) {
    const float *dw = P("initial.cnn.depthwise.weight"), *dwb = Popt("initial.cnn.depthwise.bias");
    const float *pw = P("initial.cnn.pointwise.weight"), *pwb = Popt("initial.cnn.pointwise.bias");
    const float* slope = P("initial.act.weight");
    if (err) return err;
    WC_REQUIRE(NCH == 64, "srgan_initial is written for 64 hidden channels");
    wc_srgan* n = net;
    push([=](cudaStream_t s) { return srgan_initial(n->x_in, dw, dwb, pw, pwb, slope, initial.ptr, B, H, W, initial.ld, s); });
    net->flops += 2.0 * B * H * W * (243.0 + 3.0 * NCH);   // separable, as the reference runs it
  }
  // ---- residual trunk
  Act cur = initial;
  for (int i = 0; i < net->num_blocks; ++i) {
    const std::string p = "residual." + std::to_string(i);
    Act t1 = make_act(bump, B, H, W, NCH), out = make_act(bump, B, H, W, NCH);
    if (!dry) {
      Composed c1 = compose(p + ".block1.cnn", NCH, NCH, 3, p + ".block1.bn");
      Composed c2 = compose(p + ".block2.cnn", NCH, NCH, 3, p + ".block2.bn");
      const float* slope = P(p + ".block1.act.weight");
      if (err) return err;
      conv3(cur, c1, NCH, slope, nullptr, t1);
      conv3(t1, c2, NCH, nullptr, &cur, out);
    }
    cur = out;
  }
  Act trunk = make_act(bump, B, H, W, NCH);
  if (!dry) {
    Composed c = compose("convblock.cnn", NCH, NCH, 3, "convblock.bn");
    if (err) return err;
    conv3(cur, c, NCH, nullptr, &initial, trunk);
  }
  cur = trunk;
  // ---- upsamplers: SeperableConv2d(64->256, k3) + PixelShuffle(2) + PReLU(64)
  int h = H, w = W;
  for (int u = 0; u < net->upscale / 2; ++u) {
    const std::string p = "upsampler." + std::to_string(u);
    Act up = make_act(bump, B, 2 * h, 2 * w, NCH);
    if (!dry) {
      Composed c = compose(p + ".conv", NCH, NCH * 4, 3, "");
      const float* slope = P(p + ".act.weight");
      if (err) return err;
      static int merged = -1;   // WC_SRGAN_UP_MERGED=0: one launch per PixelShuffle phase (round 1)
      if (merged < 0) {
        const char* e = getenv("WC_SRGAN_UP_MERGED");
        merged = e ? atoi(e) : 1;
      }
      if (merged && NCH % 32 == 0) {
        // ONE launch for the four phases: weight rows, bias and PReLU slopes re-ordered phase-major (row q*64 + c = conv channel 4c + q),
        // N tile 128 so that the row-segment mode applies; every 32-channel chunk is stored through the strided map of its phase
        float* wpm = fa(static_cast<size_t>(NCH) * 4 * NCH * 9);
        float* bpm = fa(NCH * 4);
        float* spm = fa(NCH * 4);
        if (!wpm || !bpm || !spm) return 1;
        if (int e = phase_major_rows(c.w, wpm, NCH, NCH * 9, 0, st)) return e;
        if (int e = phase_major_rows(c.bias, bpm, NCH, 1, 0, st)) return e;
        if (int e = phase_major_rows(slope, spm, NCH, 1, 1, st)) return e;
        WeightSrc ws_; ws_.w = wpm; ws_.d0 = NCH * 4; ws_.d1 = NCH; ws_.KH = ws_.KW = 3;
        ConvGeom g; g.K = 3; g.pad = 1;
        Epilogue ep; ep.bias = bpm; ep.prelu = spm;
        OutSpec os; os.mode = kOutNHWC; os.out = up; os.up = 2; os.phases = 4; os.bn_max = 128;
        auto op = std::make_shared<ConvOp>();
        if (int e = build_conv(op.get(), arena, cur, ws_, g, NCH * 4, nullptr, nullptr, ep, os, st)) return e;
        op->flops = 2.0 * cur.pixels() * (9.0 * NCH + static_cast<double>(NCH) * NCH * 4);
        for (auto& pl : op->plans) pl.flops = op->flops / op->plans.size();
        net->flops += op->flops;
        net->ops.push_back([op](cudaStream_t s) { return op->run(s); });
      } else
      for (int q = 0; q < 4; ++q) {   // PixelShuffle: out[c, 2y+i, 2x+j] = conv[4c + 2i + j, y, x]
        float* bq = fa(NCH);
        if (!bq) return 1;
        if (int e = gather_stride(c.bias, bq, NCH, 4, q, st)) return e;
        WeightSrc ws_; ws_.w = c.w; ws_.d0 = NCH * 4; ws_.d1 = NCH; ws_.KH = ws_.KW = 3; ws_.n_out = NCH; ws_.row_mul = 4; ws_.row_off = q;
        ConvGeom g; g.K = 3; g.pad = 1;
        Epilogue ep; ep.bias = bq; ep.prelu = slope;
        OutSpec os; os.mode = kOutNHWC; os.out = up; os.up = 2; os.py = q / 2; os.px = q % 2;
        auto op = std::make_shared<ConvOp>();
        if (int e = build_conv(op.get(), arena, cur, ws_, g, NCH, nullptr, nullptr, ep, os, st)) return e;
        op->flops = 2.0 * cur.pixels() * (9.0 * NCH / 4.0 + static_cast<double>(NCH) * NCH);   // this phase's share
        for (auto& pl : op->plans) pl.flops = op->flops / op->plans.size();
        net->flops += op->flops;
        net->ops.push_back([op](cudaStream_t s) { return op->run(s); });
      }
    }
    cur = up; h *= 2; w *= 2;
  }
  // ---- final: depthwise 9x9 + pointwise 64->3 + (tanh + 1)/2, NCHW fp32 out
  // WC_SRGAN_FINAL_TC (default 1): on the tensor core - a 1x9 horizontal convolution with N = (kernel row, output) = 27 columns
  // into fp32 planes (igemm), then the vertical shift-add + bias + tanh (seg_kernels.cu: srgan_final_combine); 0: CUDA-core kernel
  static int final_tc = -1;
  if (final_tc < 0) {
    const char* e = getenv("WC_SRGAN_FINAL_TC");
    final_tc = e ? atoi(e) : 1;
  }
  const bool tc = final_tc && NCH == 64 && w % 4 == 0;
  float* tplanes = tc ? static_cast<float*>(bump.take(static_cast<size_t>(B) * 27 * h * w * sizeof(float))) : nullptr;
  if (!dry) {
    const float *dw = P("final_conv.depthwise.weight"), *dwb = Popt("final_conv.depthwise.bias");
    const float *pw = P("final_conv.pointwise.weight"), *pwb = Popt("final_conv.pointwise.bias");
    if (err) return err;
    wc_srgan* n = net;
    const Act last = cur;
    const int hh = h, ww = w;
    const double fl = 2.0 * B * hh * ww * (81.0 * NCH + 3.0 * NCH);   // algorithmic: the reference's separable form
    if (tc) {
      float* wq = fa(32 * 64 * 9);
      float* bias3 = fa(4);
      if (!wq || !bias3) return 1;
      if (int e = srgan_final_compose(dw, dwb, pw, pwb, wq, bias3, st)) return e;
      WeightSrc ws_; ws_.w = wq; ws_.d0 = 32; ws_.d1 = NCH; ws_.KH = 1; ws_.KW = 9;
      Epilogue ep;
      OutSpec os; os.mode = kOutNCHWf32; os.out_f32 = tplanes; os.n_store = 27;
      auto op = std::make_shared<ConvOp>();
      if (int e = build_conv_hrow(op.get(), arena, last, ws_, 32, ep, os, st)) return e;
      op->flops = fl;
      for (auto& pl : op->plans) pl.flops = fl / op->plans.size();
      net->ops.push_back([op](cudaStream_t s) { return op->run(s); });
      push([=](cudaStream_t s) { return srgan_final_combine(tplanes, bias3, n->y_out, B, hh, ww, s); });
    } else {
      push([=](cudaStream_t s) { return srgan_final(last.ptr, dw, dwb, pw, pwb, n->y_out, B, hh, ww, last.ld, s); });
    }
    net->flops += fl;
  }
  if (dry) net->ws_needed = bump.used() + 4096;
  else if (bump.overflow()) return fail("SRGAN workspace too small");
  return err;
}

}  // namespace
}  // namespace wc

using namespace wc;

extern "C" {

int wc_srgan_create(wc_srgan** out, int num_blocks, int upscale, int n_params, const char* const* names,
                    const float* const* ptrs, void* stream) {
  (void)stream;
  WC_REQUIRE(out && names && ptrs, "null argument");
  WC_REQUIRE(upscale == 2 || upscale == 4 || upscale == 8, "upscale must be 2, 4 or 8");
  auto net = std::make_unique<wc_srgan>();
  net->num_blocks = num_blocks; net->upscale = upscale;
  for (int i = 0; i < n_params; ++i) net->params.ptr[names[i]] = ptrs[i];
  *out = net.release();
  return 0;
}
void wc_srgan_destroy(wc_srgan* net) { delete net; }
size_t wc_srgan_workspace_bytes(const wc_srgan* net_c, int batch, int h, int w) {
  wc_srgan* net = const_cast<wc_srgan*>(net_c);
  const int sB = net->B, sH = net->H, sW = net->W;
  net->B = batch; net->H = h; net->W = w;
  size_t need = 0;
  if (build(net, true, nullptr, 0, nullptr) == 0) need = net->ws_needed;
  net->B = sB; net->H = sH; net->W = sW;
  return need;
}
int wc_srgan_forward(wc_srgan* net, const float* x, float* y, int batch, int h, int w, void* workspace,
                     size_t workspace_bytes, void* stream) {
  WC_REQUIRE(net && x && y && workspace, "null argument");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  if (net->B != batch || net->H != h || net->W != w || net->ws != workspace || net->ws_bytes != workspace_bytes) {
    net->ops.clear();
    net->arena = std::make_unique<DeviceArena>();
    net->B = batch; net->H = h; net->W = w; net->ws = workspace; net->ws_bytes = workspace_bytes;
    net->flops = 0;
    if (int e = build(net, false, workspace, workspace_bytes, st)) {
      net->B = 0;
      net->ops.clear();
      return e;
    }
  }
  net->x_in = x; net->y_out = y;
  for (auto& op : net->ops)
    if (int e = op(st)) return e;
  return 0;
}
double wc_srgan_flops(const wc_srgan* net) { return net ? net->flops : 0.0; }
int wc_srgan_launches(const wc_srgan* net) { return net ? static_cast<int>(net->ops.size()) : 0; }

}  // extern "C"
