"""Swift-SRGAN Generator with the reference's constructor / forward contract (srgan_model/models.py:65-92) running
on the C plan in csrc/srgan.cu.  state_dict keys/shapes/order equal the reference's (283 tensors for the defaults).
Inference only (the reference wraps it in no_grad, srgan_model/inference.py:35-39); eval-mode BatchNorm."""
import ctypes as C
import math

import torch
import torch.nn as nn

from .. import _lib
from .._lib import check, lib, ptr, stream_ptr
from .._params import register_dotted


def param_spec(in_channels=3, nc=64, num_blocks=16, upscale_factor=4):
    """Ordered {name: (shape, kind)}, kind in {'w','b','prelu','bn_w','bn_b','mean','var','count'}."""
    spec = {}

    def sep(p, ci, co, k, bias):
        spec[p + ".depthwise.weight"] = ((ci, 1, k, k), "w")
        if bias:
            spec[p + ".depthwise.bias"] = ((ci,), "b")
        spec[p + ".pointwise.weight"] = ((co, ci, 1, 1), "w")
        if bias:
            spec[p + ".pointwise.bias"] = ((co,), "b")

    def convblock(p, ci, co, k, use_bn):
        sep(p + ".cnn", ci, co, k, not use_bn)
        if use_bn:
            spec[p + ".bn.weight"] = ((co,), "bn_w"); spec[p + ".bn.bias"] = ((co,), "bn_b")
            spec[p + ".bn.running_mean"] = ((co,), "mean"); spec[p + ".bn.running_var"] = ((co,), "var")
            spec[p + ".bn.num_batches_tracked"] = ((), "count")
        spec[p + ".act.weight"] = ((co,), "prelu")

    convblock("initial", in_channels, nc, 9, False)
    for i in range(num_blocks):
        convblock(f"residual.{i}.block1", nc, nc, 3, True)
        convblock(f"residual.{i}.block2", nc, nc, 3, True)
    convblock("convblock", nc, nc, 3, True)
    for i in range(upscale_factor // 2):
        sep(f"upsampler.{i}.conv", nc, nc * 4, 3, True)
        spec[f"upsampler.{i}.act.weight"] = ((nc,), "prelu")
    sep("final_conv", nc, in_channels, 9, True)
    return spec


class Generator(nn.Module):
    """Swift-SRGAN Generator (in_channels, num_channels, num_blocks, upscale_factor) -> super-resolved image in [0,1]."""

    def __init__(self, in_channels: int = 3, num_channels: int = 64, num_blocks: int = 16, upscale_factor: int = 4):
        super().__init__()
        if in_channels != 3 or num_channels != 64:
            raise NotImplementedError("the B200 SRGAN plan is built for in_channels=3, num_channels=64 (the reference's use)")
        self.num_blocks, self.upscale_factor = num_blocks, upscale_factor
        for name, (shape, kind) in param_spec(in_channels, num_channels, num_blocks, upscale_factor).items():
            if kind == "w":
                fan_in = shape[1] * shape[2] * shape[3]
                register_dotted(self, name, torch.empty(shape).uniform_(-1 / math.sqrt(fan_in), 1 / math.sqrt(fan_in)))
            elif kind == "b":
                register_dotted(self, name, torch.zeros(shape))
            elif kind == "prelu":
                register_dotted(self, name, torch.full(shape, 0.25))
            elif kind == "bn_w":
                register_dotted(self, name, torch.ones(shape))
            elif kind == "bn_b":
                register_dotted(self, name, torch.zeros(shape))
            elif kind == "mean":
                register_dotted(self, name, torch.zeros(shape), buffer=True)
            elif kind == "var":
                register_dotted(self, name, torch.ones(shape), buffer=True)
            else:
                register_dotted(self, name, torch.zeros(shape, dtype=torch.long), buffer=True)
        self._handle, self._key, self._ws, self._keep = None, None, {}, None

    def _tensors(self):
        sd = dict(self.named_parameters())
        sd.update({n: b for n, b in self.named_buffers() if b.dtype == torch.float32})
        return sd

    def _destroy(self):
        if self._handle is not None:
            lib().wc_srgan_destroy(self._handle)
            self._handle = None

    def __del__(self):
        try:
            self._destroy()
        except Exception:
            pass

    def _ensure(self, device):
        ts = self._tensors()
        key = (str(device), tuple((t.data_ptr(), t._version) for t in ts.values()))
        if self._handle is not None and key == self._key:
            return
        self._destroy()
        for n, t in ts.items():
            if t.device != device or not t.is_contiguous():
                raise RuntimeError(f"SRGAN parameter {n} must be contiguous on {device}; call .to(device)")
        n = len(ts)
        names = (C.c_char_p * n)(*[k.encode() for k in ts])
        ptrs = (C.c_void_p * n)(*[t.data_ptr() for t in ts.values()])
        h = C.c_void_p()
        check(lib().wc_srgan_create(C.byref(h), self.num_blocks, self.upscale_factor, n, names, ptrs, stream_ptr()))
        self._handle, self._key, self._keep, self._ws = h, key, list(ts.values()), {}

    def forward(self, x, out=None):
        _lib.require_cuda(x)
        if self.training:
            raise RuntimeError("the B200 SRGAN path implements eval-mode BatchNorm only; call .eval()")
        x = x.contiguous().float()
        B, _, h, w = x.shape
        self._ensure(x.device)
        k = (B, h, w)
        if k not in self._ws:
            nbytes = lib().wc_srgan_workspace_bytes(self._handle, B, h, w)
            if nbytes == 0:
                check(1)
            self._ws = {k: torch.empty(nbytes, dtype=torch.uint8, device=x.device)}
        ws = self._ws[k]
        if out is None:
            out = torch.empty(B, 3, h * self.upscale_factor, w * self.upscale_factor, device=x.device)
        check(lib().wc_srgan_forward(self._handle, ptr(x), ptr(out), B, h, w, ptr(ws), ws.numel(), stream_ptr()))
        self._last = x
        return out

    def flops(self):
        return float(lib().wc_srgan_flops(self._handle)) if self._handle is not None else 0.0
