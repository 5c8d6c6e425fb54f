"""Oracle: Swift-SRGAN generator forward (TEST INFRASTRUCTURE, see oracle/__init__.py).

Functional fp32 restatement of srgan_model/models.py of the reference (eval-mode BatchNorm):
  SeperableConv2d :5-21  ConvBlock :24-35  UpsampleBlock :38-48  ResidualBlock :51-62
  Generator.forward :87-92 ;  srgan_model/inference.py:35-39 wraps it in no_grad.
"""
import torch
import torch.nn.functional as F


def _sep(sd, p, x, k, bias):
    c = x.shape[1]
    x = F.conv2d(x, sd[p + ".depthwise.weight"], sd.get(p + ".depthwise.bias") if bias else None,
                 padding=k // 2, groups=c)
    return F.conv2d(x, sd[p + ".pointwise.weight"], sd.get(p + ".pointwise.bias") if bias else None)


def _bn(sd, p, x):
    return F.batch_norm(x, sd[p + ".running_mean"], sd[p + ".running_var"], sd[p + ".weight"], sd[p + ".bias"],
                        False, 0.0, 1e-5)


def _convblock(sd, p, x, k, use_act, use_bn):
    x = _sep(sd, p + ".cnn", x, k, bias=not use_bn)
    if use_bn:
        x = _bn(sd, p + ".bn", x)
    return F.prelu(x, sd[p + ".act.weight"]) if use_act else x


def generator_forward(sd, x, num_blocks=16, upscale_factor=4):
    initial = _convblock(sd, "initial", x, 9, True, False)
    h = initial
    for i in range(num_blocks):
        r = _convblock(sd, f"residual.{i}.block1", h, 3, True, True)
        r = _convblock(sd, f"residual.{i}.block2", r, 3, False, True)
        h = r + h
    h = _convblock(sd, "convblock", h, 3, False, True) + initial
    for i in range(upscale_factor // 2):
        h = _sep(sd, f"upsampler.{i}.conv", h, 3, True)
        h = F.prelu(F.pixel_shuffle(h, 2), sd[f"upsampler.{i}.act.weight"])
    return (torch.tanh(_sep(sd, "final_conv", h, 9, True)) + 1) / 2


def srgan_param_spec(in_channels=3, nc=64, num_blocks=16, upscale_factor=4):
    """name -> (shape, dtype) of Generator(...).state_dict() (283 tensors for the defaults)."""
    f32, i64 = torch.float32, torch.int64
    spec = {}

    def sep(p, ci, co, k, bias):
        spec[p + ".depthwise.weight"] = ((ci, 1, k, k), f32)
        if bias:
            spec[p + ".depthwise.bias"] = ((ci,), f32)
        spec[p + ".pointwise.weight"] = ((co, ci, 1, 1), f32)
        if bias:
            spec[p + ".pointwise.bias"] = ((co,), f32)

    def bn(p, c):
        spec[p + ".weight"] = ((c,), f32); spec[p + ".bias"] = ((c,), f32)
        spec[p + ".running_mean"] = ((c,), f32); spec[p + ".running_var"] = ((c,), f32)
        spec[p + ".num_batches_tracked"] = ((), i64)

    def convblock(p, ci, co, k, use_bn, prelu=True):
        sep(p + ".cnn", ci, co, k, not use_bn)
        if use_bn:
            bn(p + ".bn", co)
        spec[p + ".act.weight"] = ((co,), f32)      # PReLU exists even when use_act=False (models.py:32)

    convblock("initial", in_channels, nc, 9, False)
    for i in range(num_blocks):
        convblock(f"residual.{i}.block1", nc, nc, 3, True)
        convblock(f"residual.{i}.block2", nc, nc, 3, True)
    convblock("convblock", nc, nc, 3, True)
    for i in range(upscale_factor // 2):
        sep(f"upsampler.{i}.conv", nc, nc * 4, 3, True)
        spec[f"upsampler.{i}.act.weight"] = ((nc,), f32)
    sep("final_conv", nc, in_channels, 9, True)
    return spec
