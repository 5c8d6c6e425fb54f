"""Deterministic synthetic weights shared by the golden generator, the oracle and the CUDA path.

The reference ships no checkpoints, so every parity run uses weights that
are a pure function of (parameter name, shape, base seed).  Values are drawn per tensor from a
generator seeded with crc32(name) ^ seed, so they do not depend on module construction order.
"""
import zlib
import torch


def _gen(name: str, seed: int) -> torch.Generator:
    g = torch.Generator(device="cpu")
    g.manual_seed((zlib.crc32(name.encode()) ^ (seed * 2654435761)) & 0x7FFFFFFF)
    return g


def synth_tensor(name: str, shape, seed: int, like_dtype=torch.float32) -> torch.Tensor:
    shape = tuple(shape)
    g = _gen(name, seed)
    leaf = name.rsplit(".", 1)[-1]
    if like_dtype in (torch.int64, torch.int32):          # num_batches_tracked
        return torch.zeros(shape, dtype=like_dtype)
    if leaf == "running_mean":
        return 0.1 * torch.randn(shape, generator=g)
    if leaf == "running_var":
        return 0.5 + torch.rand(shape, generator=g)
    if leaf in ("in_proj_weight",):
        fan_in = shape[1]
        return torch.randn(shape, generator=g) / fan_in ** 0.5
    if leaf in ("in_proj_bias", "bias"):
        if len(shape) == 1:
            return 0.05 * torch.randn(shape, generator=g)
    if leaf == "weight":
        if len(shape) == 1:
            # norm scale (GroupNorm/BatchNorm) or PReLU slope: keep near 1 / 0.25
            if ".act." in name or name.endswith("act.weight"):
                return 0.25 + 0.05 * torch.randn(shape, generator=g)
            return 1.0 + 0.1 * torch.randn(shape, generator=g)
        fan_in = 1
        for s in shape[1:]:
            fan_in *= s
        if "up_sample_conv" in name:                      # ConvTranspose2d: [Cin, Cout, kh, kw]
            fan_in = shape[0] * shape[2] * shape[3] / 4.0
        return torch.randn(shape, generator=g) / max(fan_in, 1) ** 0.5
    return 0.05 * torch.randn(shape, generator=g)


def synth_state_dict(spec, seed: int):
    """spec: mapping name -> tensor (or (shape, dtype)); returns a new state dict."""
    out = {}
    for name, v in spec.items():
        if isinstance(v, torch.Tensor):
            shape, dt = v.shape, v.dtype
        else:
            shape, dt = v
        out[name] = synth_tensor(name, shape, seed, dt).to(dt)
    return out
