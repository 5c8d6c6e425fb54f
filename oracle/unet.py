"""Oracle: attention UNet forward (TEST INFRASTRUCTURE, see oracle/__init__.py).

Functional fp32 restatement of diffusion_model/models/unet_base.py of the reference, driven by a
reference-format ``state_dict`` (key names of SURVEY.md section 8b):
  get_time_embedding :7-30      DownBlock.forward :131-164   MidBlock.forward :228-268
  UpBlock.forward :336-369      Unet.__init__ :378-449       Unet.forward :451-488
The multi-head attention follows torch.nn.MultiheadAttention(batch_first=True) semantics
(packed in-proj [3C,C], q scaled by 1/sqrt(hd), softmax over keys, out-proj), SURVEY Appendix A.
"""
import math
import torch
import torch.nn.functional as F


def get_time_embedding(time_steps, temb_dim):                                     # unet_base.py:7-30
    factor = 10000 ** (torch.arange(0, temb_dim // 2, dtype=torch.float32) / (temb_dim // 2))
    t_emb = time_steps[:, None].repeat(1, temb_dim // 2) / factor
    return torch.cat([torch.sin(t_emb), torch.cos(t_emb)], dim=-1)


def _mha(sd, p, x, num_heads):
    """x: [B, N, C] -> [B, N, C]; nn.MultiheadAttention(C, heads, batch_first=True)(x, x, x)[0]."""
    B, N, C = x.shape
    hd = C // num_heads
    qkv = F.linear(x, sd[p + ".in_proj_weight"], sd[p + ".in_proj_bias"])
    q, k, v = qkv.split(C, dim=-1)
    q = q.view(B, N, num_heads, hd).transpose(1, 2) / math.sqrt(hd)
    k = k.view(B, N, num_heads, hd).transpose(1, 2)
    v = v.view(B, N, num_heads, hd).transpose(1, 2)
    att = torch.softmax(q @ k.transpose(-2, -1), dim=-1)
    o = (att @ v).transpose(1, 2).reshape(B, N, C)
    return F.linear(o, sd[p + ".out_proj.weight"], sd[p + ".out_proj.bias"])


def _resnet(sd, p, i, x, t_emb):
    """One ResNet sub-layer (unet_base.py:146-150): GN-SiLU-Conv3x3, + t-emb, GN-SiLU-Conv3x3, + Conv1x1(x)."""
    h = F.group_norm(x, 8, sd[f"{p}.resnet_conv_first.{i}.0.weight"], sd[f"{p}.resnet_conv_first.{i}.0.bias"], 1e-5)
    h = F.conv2d(F.silu(h), sd[f"{p}.resnet_conv_first.{i}.2.weight"], sd[f"{p}.resnet_conv_first.{i}.2.bias"], padding=1)
    te = F.linear(F.silu(t_emb), sd[f"{p}.t_emb_layers.{i}.1.weight"], sd[f"{p}.t_emb_layers.{i}.1.bias"])
    h = h + te[:, :, None, None]
    h = F.group_norm(h, 8, sd[f"{p}.resnet_conv_second.{i}.0.weight"], sd[f"{p}.resnet_conv_second.{i}.0.bias"], 1e-5)
    h = F.conv2d(F.silu(h), sd[f"{p}.resnet_conv_second.{i}.2.weight"], sd[f"{p}.resnet_conv_second.{i}.2.bias"], padding=1)
    return h + F.conv2d(x, sd[f"{p}.residual_input_conv.{i}.weight"], sd[f"{p}.residual_input_conv.{i}.bias"])


def _attn(sd, p, i, x, num_heads):
    """Attention sub-layer (unet_base.py:153-161)."""
    B, C, H, W = x.shape
    a = x.reshape(B, C, H * W)
    a = F.group_norm(a, 8, sd[f"{p}.attention_norms.{i}.weight"], sd[f"{p}.attention_norms.{i}.bias"], 1e-5)
    a = _mha(sd, f"{p}.attentions.{i}", a.transpose(1, 2), num_heads)
    return x + a.transpose(1, 2).reshape(B, C, H, W)


def unet_forward(sd, cfg, x, t, taps=None):
    """cfg: object/dict with the ModelConfig fields (diffusion_model/config/models.py:21-33).
    x [B,3,H,W] fp32, t int tensor [1] or [B].  Returns noise_pred [B,3,H,W]."""
    g = (lambda k: cfg[k]) if isinstance(cfg, dict) else (lambda k: getattr(cfg, k))
    dc, mc, ds = g("down_channels"), g("mid_channels"), g("down_sample")
    nh, ims, ar = g("num_heads"), g("im_size"), g("attn_resolutions")
    out = F.conv2d(x, sd["conv_in.weight"], sd["conv_in.bias"], padding=1)       # :455
    t_emb = get_time_embedding(torch.as_tensor(t).long().reshape(-1), g("time_emb_dim"))  # :461
    t_emb = F.linear(t_emb, sd["t_proj.0.weight"], sd["t_proj.0.bias"])
    t_emb = F.linear(F.silu(t_emb), sd["t_proj.2.weight"], sd["t_proj.2.bias"])  # :462
    if taps is not None:
        taps["conv_in"] = out
    skips = []
    nlev = len(dc) - 1
    for i in range(nlev):                                                        # :466-468
        skips.append(out)
        p = f"downs.{i}"
        use_attn = (ims // (2 ** i)) in ar                                       # :404-405
        for l in range(g("num_down_layers")):
            out = _resnet(sd, p, l, out, t_emb)
            if use_attn:
                out = _attn(sd, p, l, out, nh)
        if ds[i]:
            out = F.conv2d(out, sd[f"{p}.down_sample_conv.weight"], sd[f"{p}.down_sample_conv.bias"], stride=2, padding=1)
        if taps is not None:
            taps[p] = out
    for i in range(len(mc) - 1):                                                 # :473-474, MidBlock :228-268
        p = f"mids.{i}"
        out = _resnet(sd, p, 0, out, t_emb)
        for l in range(g("num_mid_layers")):
            out = _attn(sd, p, l, out, nh)
            out = _resnet(sd, p, l + 1, out, t_emb)
        if taps is not None:
            taps[p] = out
    for j, i in enumerate(reversed(range(nlev))):                                # :478-480, UpBlock :336-369
        p = f"ups.{j}"
        use_attn = (ims // (2 ** i)) in ar                                       # :434-435
        if ds[i]:
            out = F.conv_transpose2d(out, sd[f"{p}.up_sample_conv.weight"], sd[f"{p}.up_sample_conv.bias"], stride=2, padding=1)
        out = torch.cat([out, skips.pop()], dim=1)                               # :349
        for l in range(g("num_up_layers")):
            out = _resnet(sd, p, l, out, t_emb)
            if use_attn:
                out = _attn(sd, p, l, out, nh)
        if taps is not None:
            taps[p] = out
    out = F.group_norm(out, 8, sd["norm_out.weight"], sd["norm_out.bias"], 1e-5)  # :483
    return F.conv2d(F.silu(out), sd["conv_out.weight"], sd["conv_out.bias"], padding=1)  # :484-485


def unet_param_spec(cfg):
    """name -> (shape, dtype) for the reference Unet(cfg).state_dict() (382 tensors for config.yaml)."""
    g = (lambda k: cfg[k]) if isinstance(cfg, dict) else (lambda k: getattr(cfg, k))
    dc, mc, ds = g("down_channels"), g("mid_channels"), g("down_sample")
    T, ims, ar, imc = g("time_emb_dim"), g("im_size"), g("attn_resolutions"), g("im_channels")
    f32 = torch.float32
    spec = {}

    def lin(p, o, i):
        spec[p + ".weight"] = ((o, i), f32); spec[p + ".bias"] = ((o,), f32)

    def conv(p, o, i, k):
        spec[p + ".weight"] = ((o, i, k, k), f32); spec[p + ".bias"] = ((o,), f32)

    def norm(p, c):
        spec[p + ".weight"] = ((c,), f32); spec[p + ".bias"] = ((c,), f32)

    def block(p, cin, cout, nres, nattn, use_attn):
        for l in range(nres):
            ci = cin if l == 0 else cout
            norm(f"{p}.resnet_conv_first.{l}.0", ci); conv(f"{p}.resnet_conv_first.{l}.2", cout, ci, 3)
            lin(f"{p}.t_emb_layers.{l}.1", cout, T)
            norm(f"{p}.resnet_conv_second.{l}.0", cout); conv(f"{p}.resnet_conv_second.{l}.2", cout, cout, 3)
            conv(f"{p}.residual_input_conv.{l}", cout, ci, 1)
        if use_attn:
            for l in range(nattn):
                norm(f"{p}.attention_norms.{l}", cout)
                spec[f"{p}.attentions.{l}.in_proj_weight"] = ((3 * cout, cout), f32)
                spec[f"{p}.attentions.{l}.in_proj_bias"] = ((3 * cout,), f32)
                lin(f"{p}.attentions.{l}.out_proj", cout, cout)

    lin("t_proj.0", T, T); lin("t_proj.2", T, T)
    conv("conv_in", dc[0], imc, 3)
    nlev = len(dc) - 1
    for i in range(nlev):
        ua = (ims // 2 ** i) in ar
        block(f"downs.{i}", dc[i], dc[i + 1], g("num_down_layers"), g("num_down_layers"), ua)
        if ds[i]:
            conv(f"downs.{i}.down_sample_conv", dc[i + 1], dc[i + 1], 4)
    for i in range(len(mc) - 1):
        block(f"mids.{i}", mc[i], mc[i + 1], g("num_mid_layers") + 1, g("num_mid_layers"), True)
    for j, i in enumerate(reversed(range(nlev))):
        ua = (ims // 2 ** i) in ar
        cin, cout = dc[i] * 2, (dc[i - 1] if i != 0 else dc[0])
        block(f"ups.{j}", cin, cout, g("num_up_layers"), g("num_up_layers"), ua)
        if ds[i]:
            spec[f"ups.{j}.up_sample_conv.weight"] = ((cin // 2, cin // 2, 4, 4), f32)
            spec[f"ups.{j}.up_sample_conv.bias"] = ((cin // 2,), f32)
    norm("norm_out", dc[0]); conv("conv_out", imc, dc[0], 3)
    return spec


DEFAULT_MODEL_CONFIG = dict(   # diffusion_model/config/config.yaml:16-28
    name="ddpm", im_channels=3, im_size=128, down_channels=[64, 128, 256, 512, 768],
    mid_channels=[768, 768, 512], down_sample=[True, True, True, False], time_emb_dim=128,
    num_down_layers=2, num_mid_layers=2, num_up_layers=2, num_heads=4, attn_resolutions=[8, 16, 32, 64])
