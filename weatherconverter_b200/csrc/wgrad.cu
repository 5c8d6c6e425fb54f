// See wgrad.cuh for the design.
#include "wgrad.cuh"

#include <algorithm>

#include "igemm.cuh"
#include "wc_ptx.cuh"

namespace wc {

namespace {

constexpr int kThreads = 192;                 // warp 0 TMA, warp 1 MMA (+TMEM alloc), warps 2-5 epilogue
constexpr uint32_t kAtomBytes = kWgPixels * 128;  // one 64-channel x 64-pixel box
constexpr uint32_t kAStage = 2 * kAtomBytes;      // dY: two 64-channel atoms (M = 128)
constexpr uint32_t kStageBytes = 6 * kAtomBytes;  // + up to four B atoms (N <= 256)
constexpr int kStages = 4;
constexpr uint32_t kSmemBytes = kStages * kStageBytes + 1024 + 256;
constexpr uint32_t kTmemCols = 512;

// Shared-memory descriptor of an MN-major operand made of 64-element (128-byte) atoms: rows of 128 bytes = one K
// index each (here: one pixel), 8-row groups `sbo` bytes apart, consecutive 64-element atoms along M/N `lbo` bytes
// apart; 128-byte swizzle as written by TMA.
__device__ __forceinline__ uint64_t umma_smem_desc_mn(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
  uint64_t d = 0;
  d |= static_cast<uint64_t>((saddr & 0x3FFFF) >> 4);
  d |= static_cast<uint64_t>((lbo_bytes >> 4) & 0x3FFF) << 16;
  d |= static_cast<uint64_t>((sbo_bytes >> 4) & 0x3FFF) << 32;
  d |= 1ull << 46;
  d |= 2ull << 61;  // SWIZZLE_128B
  return d;
}

struct AtomRef {
  int map, dy, dx, c0, col;  // col: column offset in the packed [N][ktot] layout
};

__device__ __forceinline__ AtomRef atom_ref(const WgradArgs& p, int atom) {
  AtomRef r{0, 0, 0, 0, 0};
  int a = atom;
  for (int t = 0; t < p.ntaps; ++t) {
    const WgradTapDev tp = p.taps[t];
    if (a < tp.nkb) {
      r.map = tp.map; r.dy = tp.dy; r.dx = tp.dx; r.c0 = a * 64; r.col = tp.koff + a * 64;
      return r;
    }
    a -= tp.nkb;
  }
  return r;
}

__global__ void __launch_bounds__(kThreads, 1)
wgrad_kernel(const __grid_constant__ WgradMaps maps, const __grid_constant__ WgradArgs p) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t bar_base = smem_base + kStages * kStageBytes;
  auto full_bar = [&](int s) { return bar_base + 8u * s; };
  auto empty_bar = [&](int s) { return bar_base + 8u * (kStages + s); };
  auto tfull_bar = [&](int a) { return bar_base + 8u * (2 * kStages + a); };
  auto tempty_bar = [&](int a) { return bar_base + 8u * (2 * kStages + 2 + a); };
  const uint32_t tmem_slot = bar_base + 8u * (2 * kStages + 4);
  volatile uint32_t* tmem_slot_ptr = reinterpret_cast<volatile uint32_t*>(smem_raw + (tmem_slot - smem_u32(smem_raw)));

  const int warp = __shfl_sync(0xffffffffu, threadIdx.x >> 5, 0), lane = threadIdx.x & 31;   // warp index, provably uniform
  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&maps.dy);
    for (int i = 0; i < kWgMaxMaps; ++i) tma_prefetch_desc(&maps.x[i]);
  }
  if (warp == 1) {
    if (lane == 0) {
      for (int s = 0; s < kStages; ++s) { mbar_init(full_bar(s), 1); mbar_init(empty_bar(s), 1); }
      for (int a = 0; a < 2; ++a) { mbar_init(tfull_bar(a), 1); mbar_init(tempty_bar(a), 128); }
      fence_barrier_init();
    }
    __syncwarp();
    tmem_alloc(tmem_slot, kTmemCols);
    tmem_relinquish();
  }
  pdl_launch_dependents();   // programmatic dependent launch: the set-up above overlaps the previous kernel's tail
  pdl_wait();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot_ptr;

  const int tiles_x = (p.W + p.tw - 1) / p.tw, tiles_y = (p.H + p.th - 1) / p.th;
  const int units = p.m_tiles * p.ngroups * p.nsplit;
  auto decode_unit = [&](int u, int& mt, int& ag, int& sp) {
    sp = u % p.nsplit;
    ag = (u / p.nsplit) % p.ngroups;
    mt = u / (p.nsplit * p.ngroups);
  };
  auto split_range = [&](int sp, int& b, int& e) {
    b = static_cast<int>(static_cast<long long>(p.ptiles) * sp / p.nsplit);
    e = static_cast<int>(static_cast<long long>(p.ptiles) * (sp + 1) / p.nsplit);
  };

  if (warp == 0) {
    if (lane == 0) {
      // ===================== TMA producer =====================
      int stage = 0;
      uint32_t phase = 0;
      for (int u = blockIdx.x; u < units; u += gridDim.x) {
        int mt, ag, sp;
        decode_unit(u, mt, ag, sp);
        const int m0 = mt * 128;
        const int na = min(4, p.natoms - ag * 4);
        AtomRef at[4];
        for (int j = 0; j < na; ++j) at[j] = atom_ref(p, ag * 4 + j);
        const int nA = (m0 + 64 < p.N) ? 2 : 1;
        const uint32_t tx = static_cast<uint32_t>(nA + na) * kAtomBytes;
        int pb, pe;
        split_range(sp, pb, pe);
        for (int pt = pb; pt < pe; ++pt) {
          const int x0 = (pt % tiles_x) * p.tw, y0 = ((pt / tiles_x) % tiles_y) * p.th, b0 = (pt / (tiles_x * tiles_y)) * p.tb;
          mbar_wait(empty_bar(stage), phase ^ 1u);
          const uint32_t sa = smem_base + stage * kStageBytes;
          mbar_arrive_expect_tx(full_bar(stage), tx);
          for (int i = 0; i < nA; ++i) tma_load_4d(sa + i * kAtomBytes, &maps.dy, full_bar(stage), m0 + 64 * i, x0, y0, b0);
          for (int j = 0; j < na; ++j)
            tma_load_4d(sa + kAStage + j * kAtomBytes, &maps.x[at[j].map], full_bar(stage), at[j].c0, x0 + at[j].dx, y0 + at[j].dy, b0);
          if (++stage == kStages) { stage = 0; phase ^= 1u; }
        }
      }
    }
  } else if (warp == 1) {
    {
      // ===================== MMA issuer =====================
      // Warp-uniform control flow, one elected lane issues (see igemm.cu).
      const uint32_t tmem_base = __shfl_sync(0xffffffffu, *tmem_slot_ptr, 0);
      const uint64_t d0 = umma_smem_desc_mn(smem_base, kAtomBytes, 1024);
      const uint32_t dhi = umma_desc_hi(d0), a_lo0 = umma_desc_lo(d0), b_lo0 = a_lo0 + (kAStage >> 4);
      int stage = 0;
      uint32_t phase = 0;
      int it = 0;
      for (int u = blockIdx.x; u < units; u += gridDim.x, ++it) {
        int mt, ag, sp;
        decode_unit(u, mt, ag, sp);
        const int na = min(4, p.natoms - ag * 4);
        const uint32_t idesc = umma_idesc_bf16(128, static_cast<uint32_t>(64 * na), 1, 1);
        const int a = it & 1;
        const uint32_t aphase = (it >> 1) & 1u;
        mbar_wait(tempty_bar(a), aphase ^ 1u);
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + static_cast<uint32_t>(a) * 256u;
        int pb, pe;
        split_range(sp, pb, pe);
        for (int pt = pb; pt < pe; ++pt) {
          mbar_wait(full_bar(stage), phase);
          tc_fence_after();
          if (elect_one_sync()) {
            const uint32_t a_lo = a_lo0 + stage * (kStageBytes >> 4), b_lo = b_lo0 + stage * (kStageBytes >> 4);
#pragma unroll
            for (int k = 0; k < kWgPixels / 16; ++k)  // 16 pixels = two 8-row groups = 2048 bytes
              umma_bf16(d_tmem, umma_desc_join(a_lo + 128u * k, dhi), umma_desc_join(b_lo + 128u * k, dhi), idesc,
                        (pt > pb || k > 0) ? 1u : 0u);
            umma_commit(empty_bar(stage));
          }
          __syncwarp();
          if (++stage == kStages) { stage = 0; phase ^= 1u; }
        }
        if (elect_one_sync()) umma_commit(tfull_bar(a));
        __syncwarp();
      }
    }
  } else {
    // ===================== epilogue warps (2..5): TMEM -> fp32 partials =====================
    const int quad = warp & 3;
    const int row = quad * 32 + lane;
    const size_t mpad = static_cast<size_t>(p.m_tiles) * 128;
    int it = 0;
    for (int u = blockIdx.x; u < units; u += gridDim.x, ++it) {
      int mt, ag, sp;
      decode_unit(u, mt, ag, sp);
      const int na = min(4, p.natoms - ag * 4);
      int col[4];
      for (int j = 0; j < na; ++j) col[j] = atom_ref(p, ag * 4 + j).col;
      const int a = it & 1;
      const uint32_t aphase = (it >> 1) & 1u;
      const int m = mt * 128 + row;
      const bool valid = m < p.N;
      float* prow = p.partial + (static_cast<size_t>(sp) * mpad + m) * p.ktot;
      const uint32_t tacc = tmem_base + static_cast<uint32_t>(a) * 256u + (static_cast<uint32_t>(quad * 32) << 16);
      mbar_wait(tfull_bar(a), aphase);
      tc_fence_after();
      for (int j = 0; j < na; ++j) {
#pragma unroll
        for (int h = 0; h < 2; ++h) {
          uint32_t r[32];
          tmem_ld32(tacc + j * 64 + h * 32, r);
          tmem_wait_ld();
          if (valid) {
            float4* dst = reinterpret_cast<float4*>(prow + col[j] + h * 32);
#pragma unroll
            for (int q = 0; q < 8; ++q)
              dst[q] = make_float4(__uint_as_float(r[4 * q]), __uint_as_float(r[4 * q + 1]), __uint_as_float(r[4 * q + 2]),
                                   __uint_as_float(r[4 * q + 3]));
          }
        }
      }
      tc_fence_before();
      mbar_arrive(tempty_bar(a));
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, kTmemCols);
  }
}

// Sum the split partials in a fixed order and scatter into the PyTorch weight-gradient layout.
__global__ void wgrad_unpack_kernel(const __grid_constant__ WgradUnpackArgs p) {
  pdl_prologue();
  const size_t total = static_cast<size_t>(p.N) * p.ktot;
  for (size_t i = blockIdx.x * static_cast<size_t>(blockDim.x) + threadIdx.x; i < total;
       i += static_cast<size_t>(gridDim.x) * blockDim.x) {
    const int k = static_cast<int>(i % p.ktot), n = static_cast<int>(i / p.ktot);
    int t = 0;
    while (t + 1 < p.ntaps && k >= p.taps[t + 1].koff) ++t;
    const WgradDst d = p.taps[t];
    const int c = k - d.koff;
    if (c >= d.C) continue;
    float s = 0.f;
    for (int sp = 0; sp < p.nsplit; ++sp) s += p.partial[(static_cast<size_t>(sp) * p.mpad + n) * p.ktot + k];
    const size_t d0 = d.transpose ? (d.c0 + c) : n, d1 = d.transpose ? n : (d.c0 + c);
    d.dst[((d0 * d.D1 + d1) * d.KH + d.ky) * d.KW + d.kx] = s;
  }
}

void pick_tile64(int H, int W, int* tb, int* th, int* tw) {
  auto p2 = [](int v) { int p = 1; while (p * 2 <= v) p *= 2; return p; };
  int w = p2(W);
  if (w < W && w * 2 <= kWgPixels) w *= 2;
  if (w > kWgPixels) w = kWgPixels;
  const int rem = kWgPixels / w;
  int h = p2(H);
  if (h < H && h * 2 <= rem) h *= 2;
  if (h > rem) h = rem;
  *tw = w; *th = h; *tb = kWgPixels / (w * h);
}

}  // namespace

int build_wgrad(WgradPlan* plan, const ViewSpec& dY, int B, int H, int W, int N, const std::vector<WgradTap>& taps,
                float* partial) {
  WC_REQUIRE(!taps.empty() && static_cast<int>(taps.size()) <= kWgMaxTaps, "wgrad: tap count out of range");
  WC_REQUIRE(partial != nullptr, "wgrad: null partial workspace");
  WgradArgs& a = plan->args;
  a = WgradArgs{};
  a.B = B; a.H = H; a.W = W; a.N = N;
  pick_tile64(H, W, &a.tb, &a.th, &a.tw);
  WC_REQUIRE(dY.act.ld % 8 == 0, "wgrad: pixel stride must be a multiple of 8 elements");
  if (int e = igemm_make_amap(&plan->maps.dy, dY.act, a.tb, a.th, a.tw, dY.y0, dY.ys, dY.x0, dY.xs, H, W)) return e;
  struct Key { const void* p; int y0, ys, x0, xs, C; };
  std::vector<Key> keys;
  int koff = 0, natoms = 0;
  double macs = 0;
  a.ntaps = static_cast<int>(taps.size());
  plan->unpack = WgradUnpackArgs{};
  for (size_t i = 0; i < taps.size(); ++i) {
    const WgradTap& t = taps[i];
    WC_REQUIRE(t.x.act.ld % 8 == 0, "wgrad: pixel stride must be a multiple of 8 elements");
    int mi = -1;
    for (size_t k = 0; k < keys.size(); ++k)
      if (keys[k].p == t.x.act.ptr && keys[k].y0 == t.x.y0 && keys[k].ys == t.x.ys && keys[k].x0 == t.x.x0 &&
          keys[k].xs == t.x.xs && keys[k].C == t.x.act.C)
        mi = static_cast<int>(k);
    if (mi < 0) {
      WC_REQUIRE(static_cast<int>(keys.size()) < kWgMaxMaps, "wgrad: too many distinct input views");
      mi = static_cast<int>(keys.size());
      keys.push_back({t.x.act.ptr, t.x.y0, t.x.ys, t.x.x0, t.x.xs, t.x.act.C});
      if (int e = igemm_make_amap(&plan->maps.x[mi], t.x.act, a.tb, a.th, a.tw, t.x.y0, t.x.ys, t.x.x0, t.x.xs, H, W)) return e;
    }
    const int nkb = (t.x.act.C + 63) / 64;
    a.taps[i].map = static_cast<int8_t>(mi);
    a.taps[i].dy = static_cast<int8_t>(t.dy);
    a.taps[i].dx = static_cast<int8_t>(t.dx);
    a.taps[i].nkb = nkb;
    a.taps[i].koff = koff;
    WgradDst& d = plan->unpack.taps[i];
    d.dst = t.dst; d.koff = koff; d.C = t.x.act.C; d.KH = t.KH; d.KW = t.KW; d.ky = t.ky; d.kx = t.kx;
    d.transpose = t.transpose; d.D1 = t.D1; d.c0 = t.c0;
    koff += nkb * 64;
    natoms += nkb;
    macs += static_cast<double>(t.x.act.C) * N;
  }
  for (size_t k = keys.size(); k < kWgMaxMaps; ++k) plan->maps.x[k] = plan->maps.x[0];
  a.ktot = koff;
  a.natoms = natoms;
  a.ngroups = (natoms + 3) / 4;
  a.m_tiles = (N + 127) / 128;
  a.ptiles = ((W + a.tw - 1) / a.tw) * ((H + a.th - 1) / a.th) * ((B + a.tb - 1) / a.tb);
  const long units_mn = static_cast<long>(a.m_tiles) * a.ngroups;
  const size_t slice = static_cast<size_t>(a.m_tiles) * 128 * a.ktot * sizeof(float);
  WC_REQUIRE(slice <= kWgradPartialBytes, "wgrad: weight too large for the partial workspace");
  long nsplit = (2L * num_sms() + units_mn - 1) / units_mn;
  nsplit = std::min<long>(nsplit, a.ptiles);
  nsplit = std::min<long>(nsplit, static_cast<long>(kWgradPartialBytes / slice));
  a.nsplit = static_cast<int>(std::max<long>(nsplit, 1));
  a.partial = partial;
  plan->unpack.N = N; plan->unpack.ktot = a.ktot; plan->unpack.nsplit = a.nsplit; plan->unpack.mpad = a.m_tiles * 128;
  plan->unpack.ntaps = a.ntaps; plan->unpack.partial = partial;
  plan->grid = static_cast<int>(std::min<long>(units_mn * a.nsplit, num_sms()));
  plan->flops = 2.0 * macs * static_cast<double>(B) * H * W;
  return 0;
}

int build_conv_wgrad(WgradOp* op, const Act& x, const Act& dy, int K, int stride, int pad, int dil, float* dw,
                     const Act* x2, float* dw2, float* partial) {
  op->plans.clear();
  op->plans.emplace_back();
  WgradPlan& plan = op->plans.back();
  std::vector<WgradTap> taps;
  ViewSpec dyv; dyv.act = dy;
  const int N = dy.C;
  if (stride == 1) {
    WC_REQUIRE(2 * pad == dil * (K - 1), "stride-1 convolutions must be 'same' sized");
    WC_REQUIRE(dy.H == x.H && dy.W == x.W && dy.B == x.B, "wgrad: grid mismatch");
    for (int ky = 0; ky < K; ++ky)
      for (int kx = 0; kx < K; ++kx) {
        WgradTap t; t.x.act = x; t.dy = ky * dil - pad; t.dx = kx * dil - pad;
        t.dst = dw; t.KH = t.KW = K; t.ky = ky; t.kx = kx; t.D1 = x.C;
        taps.push_back(t);
      }
    if (x2) {
      WC_REQUIRE(x2->H == x.H && x2->W == x.W && x2->B == x.B && dw2, "wgrad: fused 1x1 input mismatch");
      WgradTap t; t.x.act = *x2; t.dst = dw2; t.D1 = x2->C;
      taps.push_back(t);
    }
  } else {
    WC_REQUIRE(stride == 2 && dil == 1 && !x2, "only stride 1 or 2 (undilated, unfused) supported");
    WC_REQUIRE(dy.H * 2 == x.H && dy.W * 2 == x.W && dy.B == x.B, "wgrad: grid mismatch");
    for (int ky = 0; ky < K; ++ky)
      for (int kx = 0; kx < K; ++kx) {
        const int oy = ky - pad, ox = kx - pad;
        const int py = ((oy % 2) + 2) % 2, px = ((ox % 2) + 2) % 2;
        WgradTap t; t.x.act = x; t.x.y0 = py; t.x.ys = 2; t.x.x0 = px; t.x.xs = 2;
        t.dy = (oy - py) / 2; t.dx = (ox - px) / 2;
        t.dst = dw; t.KH = t.KW = K; t.ky = ky; t.kx = kx; t.D1 = x.C;
        taps.push_back(t);
      }
  }
  if (int e = build_wgrad(&plan, dyv, dy.B, dy.H, dy.W, N, taps, partial)) return e;
  op->flops = plan.flops;
  return 0;
}

int build_convT_wgrad(WgradOp* op, const Act& x, const Act& dy, int K, int pad, float* dw, float* partial) {
  // out[2y+qy, 2x+qx][co] = sum_{ky = (qy+pad) mod 2, +2..} in[y + (qy+pad-ky)/2, ...][ci] * W[ci][co][ky][kx]
  WC_REQUIRE(dy.H == 2 * x.H && dy.W == 2 * x.W && dy.B == x.B, "wgrad: grid mismatch");
  op->plans.clear();
  op->flops = 0;
  for (int qy = 0; qy < 2; ++qy)
    for (int qx = 0; qx < 2; ++qx) {
      std::vector<WgradTap> taps;
      for (int ky = (qy + pad) % 2; ky < K; ky += 2)
        for (int kx = (qx + pad) % 2; kx < K; kx += 2) {
          WgradTap t; t.x.act = x; t.dy = (qy + pad - ky) / 2; t.dx = (qx + pad - kx) / 2;
          t.dst = dw; t.KH = t.KW = K; t.ky = ky; t.kx = kx; t.transpose = 1; t.D1 = dy.C;
          taps.push_back(t);
        }
      if (taps.empty()) continue;
      ViewSpec dyv; dyv.act = dy; dyv.y0 = qy; dyv.ys = 2; dyv.x0 = qx; dyv.xs = 2;
      op->plans.emplace_back();
      if (int e = build_wgrad(&op->plans.back(), dyv, x.B, x.H, x.W, dy.C, taps, partial)) return e;
      op->flops += op->plans.back().flops;
    }
  return 0;
}

int wgrad_launch(const WgradPlan& plan, cudaStream_t st) {
  static bool attr_set = false;
  if (!attr_set) {
    WC_CHECK_CUDA(cudaFuncSetAttribute(wgrad_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemBytes));
    attr_set = true;
  }
  {
    ProfScope prof(kProfIgemm, st, plan.flops);
    prof.note(plan.args.B * plan.args.H * plan.args.W, plan.args.N, plan.args.ktot, plan.args.nsplit, 1000 + plan.args.ntaps, plan.grid);
    launch_k<1>(wgrad_kernel, plan.grid, kThreads, kSmemBytes, st, plan.maps, plan.args);
    WC_LAUNCH_CHECK();
  }
  const size_t total = static_cast<size_t>(plan.unpack.N) * plan.unpack.ktot;
  const int blocks = static_cast<int>(std::min<size_t>((total + 255) / 256, 4096));
  ProfScope prof(kProfOther, st, 0);
  launch_k(wgrad_unpack_kernel, blocks, 256, 0, st, plan.unpack);
  WC_LAUNCH_CHECK();
  return 0;
}

}  // namespace wc
