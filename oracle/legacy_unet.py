"""Oracle: legacy "old model" UNet forward + the integrated sampler (TEST INFRASTRUCTURE, see oracle/__init__.py).

Functional fp32 restatement of diffusion_model/models/old_modules.py of the reference:
  SelfAttention :73-94   ResidualBlock :126-160   DownBlock :163-195   UpBlock :198-227
  UNet.__init__ :244-281  sinusoidal_embedding :283-307  UNet.forward :309-360
and of the loop of diffusion_model/sample_integrated.py:40-67 (batched t, noise-variance conditioning,
sample_prev_timestep2).
"""
import math

import torch
import torch.nn.functional as F

DOWN = [("down1", 64, 32), ("down2", 32, 64), ("down3", 64, 96), ("down4", 96, 128)]
UP = [("up1", 256, 128, 128), ("up2", 128, 96, 96), ("up3", 96, 64, 64), ("up4", 64, 32, 32)]
ATTN = {"attn_down3": (64, 32), "attn_down4": (96, 16), "attn_bottleneck": (256, 8), "attn_up1": (128, 16), "attn_up2": (96, 32)}
DEPTH = 3


def sinusoidal_embedding(x):                                                     # old_modules.py:283-307
    freqs = torch.exp(torch.linspace(math.log(1.0), math.log(1000.0), 16))
    ang = 2.0 * math.pi * freqs
    return torch.cat([torch.sin(ang * x), torch.cos(ang * x)], dim=3).permute(0, 3, 1, 2)


def _resblock(sd, p, x, residual):                                               # :126-160
    res = F.conv2d(x, sd[p + ".res.weight"]) if residual else x
    h = F.batch_norm(x, sd[p + ".double_conv.0.running_mean"], sd[p + ".double_conv.0.running_var"],
                     sd[p + ".double_conv.0.weight"], sd[p + ".double_conv.0.bias"], False, 0.0, 1e-5)
    h = F.silu(F.conv2d(h, sd[p + ".double_conv.1.weight"], padding=1))
    h = F.conv2d(h, sd[p + ".double_conv.3.weight"], padding=1)
    return h + res


def _attn(sd, p, x, channels, size):                                             # :73-94
    B = x.shape[0]
    t = x.view(B, channels, size * size).swapaxes(1, 2)
    ln = F.layer_norm(t, (channels,), sd[p + ".ln.weight"], sd[p + ".ln.bias"])
    hd = channels // 4
    qkv = F.linear(ln, sd[p + ".mha.in_proj_weight"], sd[p + ".mha.in_proj_bias"])
    q, k, v = qkv.split(channels, dim=-1)
    q = q.view(B, -1, 4, hd).transpose(1, 2) / math.sqrt(hd)
    k = k.view(B, -1, 4, hd).transpose(1, 2)
    v = v.view(B, -1, 4, hd).transpose(1, 2)
    o = (torch.softmax(q @ k.transpose(-2, -1), dim=-1) @ v).transpose(1, 2).reshape(B, -1, channels)
    av = F.linear(o, sd[p + ".mha.out_proj.weight"], sd[p + ".mha.out_proj.bias"]) + t
    ff = F.layer_norm(av, (channels,), sd[p + ".ff_self.0.weight"], sd[p + ".ff_self.0.bias"])
    ff = F.linear(F.gelu(F.linear(ff, sd[p + ".ff_self.1.weight"], sd[p + ".ff_self.1.bias"])),
                  sd[p + ".ff_self.3.weight"], sd[p + ".ff_self.3.bias"])
    return (ff + av).swapaxes(2, 1).reshape(B, channels, size, size)


def legacy_unet_forward(sd, x, t):
    """x [B,3,128,128]; t [B,1,1,1] float (noise variance 1 - alpha_bar_t).  Returns [B,3,128,128]."""
    size = x.shape[-1]
    h = F.conv2d(x, sd["pre_conv.weight"], padding=1)                            # :311
    emb = F.interpolate(sinusoidal_embedding(t), size=(size, size), mode="nearest")   # :313-315
    h = torch.cat([h, emb], dim=1)
    skips = []
    for name, cin, cout in DOWN:                                                 # :321-333
        if name == "down3":
            h = _attn(sd, "attn_down3", h, 64, size // 4)
        if name == "down4":
            h = _attn(sd, "attn_down4", h, 96, size // 8)
        sk = []
        for i in range(DEPTH):
            h = _resblock(sd, f"{name}.residual_blocks.{i}", h, residual=(i == 0))
            sk.append(h)
        h = F.avg_pool2d(h, 2)
        skips.append(sk)
    h = _resblock(sd, "bottleneck1", h, True)                                    # :336-340
    h = _attn(sd, "attn_bottleneck", h, 256, size // 16)
    h = _resblock(sd, "bottleneck2", h, True)
    for name, cin, cout, cskip in UP:                                            # :343-355
        sk = skips.pop()
        h = F.interpolate(h, scale_factor=2, mode="bilinear")
        for i in range(DEPTH):
            h = torch.cat([h, sk.pop()], dim=1)
            h = _resblock(sd, f"{name}.residual_blocks.{i}", h, True)
        if name == "up1":
            h = _attn(sd, "attn_up1", h, 128, size // 8)
        if name == "up2":
            h = _attn(sd, "attn_up2", h, 96, size // 4)
    return F.conv2d(h, sd["output.weight"], padding=1)                           # :357


def legacy_param_spec():
    f32, i64 = torch.float32, torch.int64
    spec = {"pre_conv.weight": ((32, 3, 3, 3), f32)}

    def rb(p, cin, cout):
        spec[p + ".res.weight"] = ((cout, cin, 1, 1), f32)
        for k, sh in (("weight", (cin,)), ("bias", (cin,)), ("running_mean", (cin,)), ("running_var", (cin,))):
            spec[f"{p}.double_conv.0.{k}"] = (sh, f32)
        spec[p + ".double_conv.0.num_batches_tracked"] = ((), i64)
        spec[p + ".double_conv.1.weight"] = ((cout, cin, 3, 3), f32)
        spec[p + ".double_conv.3.weight"] = ((cout, cout, 3, 3), f32)

    def attn(p, c):
        spec[p + ".mha.in_proj_weight"] = ((3 * c, c), f32); spec[p + ".mha.in_proj_bias"] = ((3 * c,), f32)
        spec[p + ".mha.out_proj.weight"] = ((c, c), f32); spec[p + ".mha.out_proj.bias"] = ((c,), f32)
        spec[p + ".ln.weight"] = ((c,), f32); spec[p + ".ln.bias"] = ((c,), f32)
        spec[p + ".ff_self.0.weight"] = ((c,), f32); spec[p + ".ff_self.0.bias"] = ((c,), f32)
        spec[p + ".ff_self.1.weight"] = ((c, c), f32); spec[p + ".ff_self.1.bias"] = ((c,), f32)
        spec[p + ".ff_self.3.weight"] = ((c, c), f32); spec[p + ".ff_self.3.bias"] = ((c,), f32)

    # registration order of old_modules.UNet.__init__ (:244-281)
    for i in range(DEPTH):
        rb(f"down1.residual_blocks.{i}", 64 if i == 0 else 32, 32)
    for i in range(DEPTH):
        rb(f"down2.residual_blocks.{i}", 32 if i == 0 else 64, 64)
    attn("attn_down3", 64)
    for i in range(DEPTH):
        rb(f"down3.residual_blocks.{i}", 64 if i == 0 else 96, 96)
    attn("attn_down4", 96)
    for i in range(DEPTH):
        rb(f"down4.residual_blocks.{i}", 96 if i == 0 else 128, 128)
    rb("bottleneck1", 128, 256)
    attn("attn_bottleneck", 256)
    rb("bottleneck2", 256, 256)
    for name, cin, cout, cskip in UP:
        for i in range(DEPTH):
            rb(f"{name}.residual_blocks.{i}", (cin if i == 0 else cout) + cskip, cout)
        if name == "up1":
            attn("attn_up1", 128)
        if name == "up2":
            attn("attn_up2", 96)
    spec["output.weight"] = ((3, 32, 3, 3), f32)
    return spec
