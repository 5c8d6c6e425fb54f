"""Op-level Python wrappers over the C ABI (torch tensors in, torch tensors out).

PyTorch is used only to own device memory and streams; all arithmetic happens inside libwc_b200.so.
Activation tensors at this level are NHWC bf16 (``[B,H,W,C]``, contiguous or channel-sliced views).
"""
import torch

from . import _lib
from ._lib import check, lib, ptr, require_cuda, stream_ptr


def to_nhwc_bf16(x: torch.Tensor) -> torch.Tensor:
    """[B,C,H,W] fp32 (reference layout) -> [B,H,W,C] bf16."""
    require_cuda(x)
    x = x.contiguous().float()
    B, Cc, H, W = x.shape
    y = torch.empty(B, H, W, Cc, device=x.device, dtype=torch.bfloat16)
    check(lib().wc_nchw_f32_to_nhwc_bf16(ptr(x), ptr(y), B, Cc, H * W, Cc, stream_ptr()))
    return y


def to_nchw_f32(x: torch.Tensor) -> torch.Tensor:
    """[B,H,W,C] bf16 -> [B,C,H,W] fp32."""
    require_cuda(x)
    B, H, W, Cc = x.shape
    assert x.stride(3) == 1 and x.stride(1) == W * x.stride(2) and x.stride(0) == H * x.stride(1)
    y = torch.empty(B, Cc, H, W, device=x.device, dtype=torch.float32)
    check(lib().wc_nhwc_bf16_to_nchw_f32(ptr(x), ptr(y), B, Cc, H * W, x.stride(2), stream_ptr()))
    return y


def groupnorm_silu(x, gamma, beta, silu=True, eps=1e-5):
    require_cuda(x, gamma, beta)
    B, H, W, Cc = x.shape
    y = torch.empty(B, H, W, Cc, device=x.device, dtype=torch.bfloat16)
    ws = torch.empty(lib().wc_groupnorm_workspace_bytes(B), device=x.device, dtype=torch.uint8)
    check(lib().wc_groupnorm_silu(ptr(x), ptr(y), B, H * W, Cc, x.stride(2), Cc, ptr(gamma), ptr(beta), eps,
                                  1 if silu else 0, ptr(ws), stream_ptr()))
    return y


def conv2d(x, weight, bias=None, stride=1, pad=None, dil=1, transposed=False, rowbias=None, residual=None,
           x2=None, weight2=None, relu=False):
    """x [B,H,W,Cin] bf16; weight fp32 PyTorch layout; returns [B,Ho,Wo,Cout] bf16."""
    require_cuda(x, weight)
    B, H, W, Cin = x.shape
    K = weight.shape[-1]
    Cout = weight.shape[1] if transposed else weight.shape[0]
    if pad is None:
        pad = dil * (K - 1) // 2
    Ho, Wo = (2 * H, 2 * W) if transposed else ((H // 2, W // 2) if stride == 2 else (H, W))
    y = torch.empty(B, Ho, Wo, Cout, device=x.device, dtype=torch.bfloat16)
    weight = weight.contiguous().float()
    if weight2 is not None:
        weight2 = weight2.contiguous().float()
    check(lib().wc_conv2d(ptr(x), B, H, W, Cin, x.stride(2), ptr(weight), ptr(bias), Cout, K, stride, pad, dil,
                          1 if transposed else 0, ptr(rowbias), ptr(residual),
                          residual.stride(2) if residual is not None else 0,
                          ptr(x2), x2.shape[3] if x2 is not None else 0, x2.stride(2) if x2 is not None else 0,
                          ptr(weight2), 1 if relu else 0, ptr(y), Cout, stream_ptr()))
    return y


def conv_in(x, weight, bias=None, scale=None, shift=None, stride=1, pad=None, relu=False):
    """Boundary conv from the reference layout: x [B,3,H,W] fp32 -> [B,Ho,Wo,Cout] bf16."""
    require_cuda(x, weight)
    B, Cc, H, W = x.shape
    Cout, _, K, _ = weight.shape
    if pad is None:
        pad = K // 2
    Ho, Wo = (H + 2 * pad - K) // stride + 1, (W + 2 * pad - K) // stride + 1
    y = torch.empty(B, Ho, Wo, Cout, device=x.device, dtype=torch.bfloat16)
    x, weight = x.contiguous(), weight.contiguous()    # named: must outlive the raw-pointer call
    check(lib().wc_conv_in(ptr(x), ptr(weight), ptr(bias), ptr(scale), ptr(shift), ptr(y),
                           B, H, W, Cout, K, stride, pad, Cout, 1 if relu else 0, stream_ptr()))
    return y


def conv_out(x, weight, bias=None, tanh_out=False):
    """Boundary conv to the reference layout: x [B,H,W,Cin] bf16 -> [B,3,H,W] fp32."""
    require_cuda(x, weight)
    B, H, W, Cin = x.shape
    K = weight.shape[-1]
    y = torch.empty(B, 3, H, W, device=x.device, dtype=torch.float32)
    weight = weight.contiguous()
    check(lib().wc_conv_out(ptr(x), ptr(weight), ptr(bias), ptr(y), B, H, W, Cin, K, x.stride(2),
                            1 if tanh_out else 0, stream_ptr()))
    return y


def attention(q, k, vt):
    """q,k [B,heads,N,hd] bf16; vt [B,heads,hd,N] bf16 -> [B,N,heads*hd] bf16."""
    require_cuda(q, k, vt)
    B, h, N, hd = q.shape
    out = torch.empty(B, N, h * hd, device=q.device, dtype=torch.bfloat16)
    q, k, vt = q.contiguous(), k.contiguous(), vt.contiguous()
    check(lib().wc_attention(ptr(q), ptr(k), ptr(vt), ptr(out), B, h, N, hd,
                             h * hd, stream_ptr()))
    return out


def ddpm_step(xt, eps, z, beta, sqrt_one_minus_acp, sqrt_alpha, sigma, want_parts=False):
    """Fused posterior update; returns x_{t-1} (and (mean, sigma*z) when want_parts)."""
    require_cuda(xt, eps, z)
    xt, eps = xt.contiguous(), eps.contiguous()
    B = xt.shape[0]
    n = xt[0].numel()
    out = torch.empty_like(xt)
    mean = torch.empty_like(xt) if want_parts else None
    sigz = torch.empty_like(xt) if (want_parts and z is not None) else None
    z = z.contiguous() if z is not None else None
    check(lib().wc_ddpm_step(ptr(xt), ptr(eps), ptr(z), ptr(out), ptr(mean),
                             ptr(sigz), n, B, float(beta), float(sqrt_one_minus_acp), float(sqrt_alpha), float(sigma),
                             stream_ptr()))
    return (out, mean, sigz) if want_parts else out


def launch_count() -> int:
    return int(lib().wc_launch_count())
