"""Op-level GPU parity of the DeepLabV3+ backward building blocks against torch autograd (fp32, TF32 off)."""
import math

import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu


def _dev():
    if not torch.cuda.is_available():
        pytest.skip("needs CUDA")
    return torch.device("cuda")


def _bf(x):
    return x.to(torch.bfloat16).float()


def _rel(a, b):
    return float((a - b).norm() / (b.norm() + 1e-30))


@pytest.mark.parametrize("Cin,Cout,Ho,Wo,K,stride,dil", [
    (256, 64, 16, 32, 1, 1, 1), (64, 64, 16, 32, 3, 1, 1), (128, 128, 8, 16, 3, 2, 1), (256, 512, 8, 16, 1, 2, 1),
    (512, 512, 16, 32, 3, 1, 2), (2048, 256, 16, 32, 3, 1, 12), (304, 256, 16, 32, 3, 1, 1)])
def test_conv_dgrad(Cin, Cout, Ho, Wo, K, stride, dil):
    from weatherconverter_b200 import ops
    from weatherconverter_b200._lib import check, lib, ptr, stream_ptr
    dev = _dev()
    g = torch.Generator().manual_seed(Cin + Cout + K)
    B = 2
    H, W = Ho * stride, Wo * stride
    w = (torch.randn(Cout, Cin, K, K, generator=g) / math.sqrt(Cin * K * K)).to(dev)
    dz = torch.randn(B, Cout, Ho, Wo, generator=g).to(dev)
    mask = torch.randn(B, Cin, H, W, generator=g).to(dev)
    res = torch.randn(B, Cin, H, W, generator=g).to(dev)
    pad = dil * (K - 1) // 2 if stride == 1 else (1 if K == 3 else 0)
    ref = torch.nn.grad.conv2d_input((B, Cin, H, W), _bf(w), _bf(dz), stride=stride, padding=pad, dilation=dil)
    ref = (ref + _bf(res)) * (_bf(mask) > 0)
    dzh, mh, rh = ops.to_nhwc_bf16(dz), ops.to_nhwc_bf16(mask), ops.to_nhwc_bf16(res)
    dx = rh.clone() if stride == 2 else torch.empty(B, H, W, Cin, device=dev, dtype=torch.bfloat16)
    # stride 2: the plan initialises dx with the residual and accumulates in place (only covered phases are written)
    check(lib().wc_conv2d_dgrad(ptr(dzh), B, Ho, Wo, Cout, ptr(w), Cin, K, stride, dil, ptr(dx) if stride == 2 else ptr(rh),
                                ptr(mh), ptr(dx), stream_ptr()))
    got = ops.to_nchw_f32(dx)
    if stride == 2 and K == 1:   # uncovered phases keep the (unmasked) initial value in this op-level call
        ref2 = torch.nn.grad.conv2d_input((B, Cin, H, W), _bf(w), _bf(dz), stride=2, padding=0)
        cov = torch.zeros_like(ref2, dtype=torch.bool); cov[:, :, ::2, ::2] = True
        ref = torch.where(cov, (ref2 + _bf(res)) * (_bf(mask) > 0), _bf(res))
    assert _rel(got, ref) < 6e-3


def test_maxpool_fwd_bwd():
    from weatherconverter_b200 import ops
    from weatherconverter_b200._lib import check, lib, ptr, stream_ptr
    dev = _dev()
    g = torch.Generator().manual_seed(3)
    B, Cc, H, W = 2, 64, 32, 64
    x = F.relu(torch.randn(B, Cc, H, W, generator=g)).to(dev)
    xb = _bf(x).requires_grad_(True)
    y_ref = F.max_pool2d(xb, 3, 2, 1)
    dy = torch.randn(B, Cc, H // 2, W // 2, generator=g).to(dev)
    y_ref.backward(_bf(dy))
    ref_dx = xb.grad * (xb > 0)
    xh = ops.to_nhwc_bf16(x)
    y = torch.empty(B, H // 2, W // 2, Cc, device=dev, dtype=torch.bfloat16)
    idx = torch.empty(B, H // 2, W // 2, Cc, device=dev, dtype=torch.uint8)
    check(lib().wc_maxpool3x3s2(ptr(xh), ptr(y), ptr(idx), B, H, W, Cc, stream_ptr()))
    assert torch.equal(ops.to_nchw_f32(y), y_ref.detach())
    dx = torch.empty_like(xh)
    dyh = ops.to_nhwc_bf16(dy)
    check(lib().wc_maxpool3x3s2_bwd(ptr(dyh), ptr(idx), ptr(xh), ptr(dx), B, H, W, Cc, stream_ptr()))
    assert _rel(ops.to_nchw_f32(dx), ref_dx) < 5e-3


@pytest.mark.parametrize("Hi,Wi,Ho,Wo,C", [(16, 32, 64, 128, 256), (4, 8, 16, 32, 64)])
def test_bilinear_fwd_bwd(Hi, Wi, Ho, Wo, C):
    from weatherconverter_b200 import ops
    from weatherconverter_b200._lib import check, lib, ptr, stream_ptr
    dev = _dev()
    g = torch.Generator().manual_seed(Hi)
    B = 2
    x = torch.randn(B, C, Hi, Wi, generator=g).to(dev)
    xb = _bf(x).requires_grad_(True)
    y_ref = F.interpolate(xb, size=(Ho, Wo), mode="bilinear", align_corners=False)
    dy = torch.randn(B, C, Ho, Wo, generator=g).to(dev)
    y_ref.backward(_bf(dy))
    y = torch.empty(B, Ho, Wo, C, device=dev, dtype=torch.bfloat16)
    xh = ops.to_nhwc_bf16(x)
    check(lib().wc_bilinear(ptr(xh), ptr(y), B, Hi, Wi, Ho, Wo, C, stream_ptr()))
    assert _rel(ops.to_nchw_f32(y), y_ref.detach()) < 4e-3
    dx = torch.empty(B, Hi, Wi, C, device=dev, dtype=torch.bfloat16)
    dyh = ops.to_nhwc_bf16(dy)
    check(lib().wc_bilinear_bwd(ptr(dyh), None, ptr(dx), B, Hi, Wi, Ho, Wo, C, stream_ptr()))
    assert _rel(ops.to_nchw_f32(dx), xb.grad) < 4e-3


def test_loss_head():
    from weatherconverter_b200 import ops
    from weatherconverter_b200._lib import check, lib, ptr, stream_ptr
    dev = _dev()
    g = torch.Generator().manual_seed(5)
    B, h, w, H, W = 2, 16, 32, 64, 128
    lo = (3 * torch.randn(B, 19, h, w, generator=g)).to(dev).requires_grad_(True)
    lab = torch.randint(0, 19, (B, H, W), generator=g)
    lab[torch.rand(B, H, W, generator=g) < 0.05] = 255
    lab = lab.to(dev)
    hi = F.interpolate(lo, size=(H, W), mode="bilinear", align_corners=False)
    losses = torch.stack([F.cross_entropy(hi[b:b + 1], lab[b:b + 1], ignore_index=255) for b in range(B)])
    losses.sum().backward()
    pred = torch.empty(B, H, W, dtype=torch.long, device=dev)
    dhi = torch.empty(B, 19, H, W, device=dev)
    loss = torch.empty(B, device=dev)
    nv = torch.empty(B, dtype=torch.int32, device=dev)
    dlo = torch.empty(B, h, w, 32, device=dev, dtype=torch.bfloat16)
    lhi = torch.empty(B, 19, H, W, device=dev)
    lo_d = lo.detach()
    check(lib().wc_seg_loss_head(ptr(lo_d), ptr(lab), ptr(nv), ptr(pred), ptr(dhi), ptr(loss), ptr(lhi), ptr(dlo),
                                 B, h, w, H, W, stream_ptr()))
    assert (lhi - hi.detach()).abs().max() < 1e-4
    assert (pred == hi.argmax(1)).float().mean() > 0.9999
    assert (loss - losses.detach()).abs().max() < 1e-4
    got = ops.to_nchw_f32(dlo)
    assert _rel(got[:, :19], lo.grad) < 5e-3
    assert float(got[:, 19:].abs().max()) == 0.0
    # fused path (no full-resolution d-logit tensor): same pred / loss / low-resolution gradient from one kernel
    pred2 = torch.empty_like(pred); loss2 = torch.empty_like(loss)
    dlo2 = torch.full_like(dlo, float("nan"))
    check(lib().wc_seg_loss_head(ptr(lo_d), ptr(lab), ptr(nv), ptr(pred2), None, ptr(loss2), None, ptr(dlo2),
                                 B, h, w, H, W, stream_ptr()))
    assert torch.equal(pred2, pred)
    assert (loss2 - loss).abs().max() < 1e-5
    got2 = ops.to_nchw_f32(dlo2)
    assert _rel(got2[:, :19], lo.grad) < 5e-3
    assert _rel(got2[:, :19], got[:, :19]) < 4e-3          # both are bf16 roundings of the same fp32 sums (different summation order)
    assert float(got2[:, 19:].abs().max()) == 0.0


def test_conv1_dgrad():
    from weatherconverter_b200 import ops
    from weatherconverter_b200._lib import check, lib, ptr, stream_ptr
    dev = _dev()
    g = torch.Generator().manual_seed(8)
    B, H, W = 2, 64, 128
    w = (torch.randn(64, 3, 7, 7, generator=g) / math.sqrt(147)).to(dev)
    sc = (1 + 0.1 * torch.randn(64, generator=g)).to(dev)
    dz = torch.randn(B, 64, H // 2, W // 2, generator=g).to(dev)
    ref = torch.nn.grad.conv2d_input((B, 3, H, W), w * sc[:, None, None, None], _bf(dz), stride=2, padding=3)
    dx = torch.empty(B, 3, H, W, device=dev)
    dzh = ops.to_nhwc_bf16(dz)
    check(lib().wc_conv1_dgrad(ptr(dzh), ptr(w), ptr(sc), ptr(dx), B, H, W, stream_ptr()))
    assert _rel(dx, ref) < 1e-4
