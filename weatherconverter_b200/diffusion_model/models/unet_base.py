"""Unet with the reference's interface (diffusion_model/models/unet_base.py:372-488) running on libwc_b200.so.

``Unet(model_config).forward(x, t)`` takes the reference's tensors (x [B,3,H,W] fp32 NCHW, t int tensor [1] or
[B] / int / tuple) and returns noise_pred [B,3,H,W] fp32.  ``state_dict()`` has exactly the reference's key
names and shapes (382 tensors for config.yaml), so reference checkpoints load with ``load_state_dict``.

The module only owns parameters and device buffers; the layer graph, weight re-packing (bf16, K-major) and all
arithmetic live in the C library (csrc/unet.cu).  Inference only: outputs carry no autograd graph.
"""
import ctypes as C
import math

import torch
import torch.nn as nn

from ... import _lib
from ..._lib import UnetConfigStruct, check, lib, ptr, stream_ptr


def _cfg_get(cfg, k):
    return cfg[k] if isinstance(cfg, dict) else getattr(cfg, k)


def param_spec(cfg):
    """Ordered {name: shape} of the reference Unet's parameters for a ModelConfig
    (structure of unet_base.py:378-449: t_proj, conv_in, downs, mids, ups, norm_out, conv_out)."""
    dc, mc, ds = list(_cfg_get(cfg, "down_channels")), list(_cfg_get(cfg, "mid_channels")), list(_cfg_get(cfg, "down_sample"))
    T, ims, imc = _cfg_get(cfg, "time_emb_dim"), _cfg_get(cfg, "im_size"), _cfg_get(cfg, "im_channels")
    attn_res = list(_cfg_get(cfg, "attn_resolutions"))
    spec = {}

    def wb(prefix, wshape):
        spec[prefix + ".weight"] = tuple(wshape)
        spec[prefix + ".bias"] = (wshape[0],)

    def stage(prefix, cin, cout, n_res, n_attn):
        for l in range(n_res):
            ci = cin if l == 0 else cout
            wb(f"{prefix}.resnet_conv_first.{l}.0", (ci,))
            wb(f"{prefix}.resnet_conv_first.{l}.2", (cout, ci, 3, 3))
        for l in range(n_res):
            wb(f"{prefix}.t_emb_layers.{l}.1", (cout, T))
        for l in range(n_res):
            wb(f"{prefix}.resnet_conv_second.{l}.0", (cout,))
            wb(f"{prefix}.resnet_conv_second.{l}.2", (cout, cout, 3, 3))
        for l in range(n_attn):
            wb(f"{prefix}.attention_norms.{l}", (cout,))
        for l in range(n_attn):
            spec[f"{prefix}.attentions.{l}.in_proj_weight"] = (3 * cout, cout)
            spec[f"{prefix}.attentions.{l}.in_proj_bias"] = (3 * cout,)
            wb(f"{prefix}.attentions.{l}.out_proj", (cout, cout))
        for l in range(n_res):
            wb(f"{prefix}.residual_input_conv.{l}", (cout, cin if l == 0 else cout, 1, 1))

    wb("t_proj.0", (T, T))
    wb("t_proj.2", (T, T))
    wb("conv_in", (dc[0], imc, 3, 3))
    levels = len(dc) - 1
    for i in range(levels):
        attn = (ims // 2 ** i) in attn_res
        stage(f"downs.{i}", dc[i], dc[i + 1], _cfg_get(cfg, "num_down_layers"), _cfg_get(cfg, "num_down_layers") if attn else 0)
        if ds[i]:
            wb(f"downs.{i}.down_sample_conv", (dc[i + 1], dc[i + 1], 4, 4))
    for i in range(len(mc) - 1):
        stage(f"mids.{i}", mc[i], mc[i + 1], _cfg_get(cfg, "num_mid_layers") + 1, _cfg_get(cfg, "num_mid_layers"))
    for j, i in enumerate(reversed(range(levels))):
        attn = (ims // 2 ** i) in attn_res
        stage(f"ups.{j}", 2 * dc[i], dc[i - 1] if i != 0 else dc[0], _cfg_get(cfg, "num_up_layers"),
              _cfg_get(cfg, "num_up_layers") if attn else 0)
        if ds[i]:
            spec[f"ups.{j}.up_sample_conv.weight"] = (dc[i], dc[i], 4, 4)   # ConvTranspose2d: [Cin, Cout, 4, 4]
            spec[f"ups.{j}.up_sample_conv.bias"] = (dc[i],)
    wb("norm_out", (dc[0],))
    wb("conv_out", (imc, dc[0], 3, 3))
    return spec


class _Node(nn.Module):
    """Anonymous container: gives parameters the reference's dotted state_dict names."""


def _init_param(name, shape):
    leaf = name.rsplit(".", 1)[-1]
    if len(shape) == 1:
        if leaf == "weight":          # GroupNorm scale
            return torch.ones(shape)
        if "norm" in name or ".0.bias" in name and "resnet_conv" in name:
            return torch.zeros(shape)
        if leaf in ("in_proj_bias",) or name.endswith("out_proj.bias"):
            return torch.zeros(shape)
        return None                   # conv / linear bias: filled with its weight's fan-in below
    if leaf == "in_proj_weight":
        bound = math.sqrt(6.0 / (shape[0] + shape[1]))    # xavier_uniform (nn.MultiheadAttention default)
    else:
        fan_in = 1
        for s in shape[1:]:
            fan_in *= s
        if "up_sample_conv" in name:
            fan_in = shape[0] * shape[2] * shape[3]
        bound = 1.0 / math.sqrt(fan_in)                   # kaiming_uniform(a=sqrt(5)) (nn.Conv2d / nn.Linear default)
    return torch.empty(shape).uniform_(-bound, bound)


class Unet(nn.Module):
    r"""Unet comprising down blocks, mid blocks and up blocks (same constructor/forward contract as the reference)."""

    def __init__(self, model_config):
        super().__init__()
        g = lambda k: _cfg_get(model_config, k)  # noqa: E731
        self.down_channels = list(g("down_channels"))
        self.mid_channels = list(g("mid_channels"))
        self.t_emb_dim = g("time_emb_dim")
        self.down_sample = list(g("down_sample"))
        self.num_down_layers = g("num_down_layers")
        self.num_mid_layers = g("num_mid_layers")
        self.num_up_layers = g("num_up_layers")
        self.attn_resolutions = list(g("attn_resolutions"))
        self.num_heads = g("num_heads")
        self.im_size = g("im_size")
        self.im_channels = g("im_channels")
        assert self.mid_channels[0] == self.down_channels[-1]
        assert self.mid_channels[-1] == self.down_channels[-2]
        assert len(self.down_sample) == len(self.down_channels) - 1
        self._spec = param_spec(model_config)
        for name, shape in self._spec.items():
            init = _init_param(name, shape)
            if init is None:   # bias of a conv / linear: U(-1/sqrt(fan_in), 1/sqrt(fan_in)) of its weight
                wshape = self._spec[name[:-4] + "weight"]
                fan_in = 1
                for s in wshape[1:]:
                    fan_in *= s
                init = torch.empty(shape).uniform_(-1.0 / math.sqrt(fan_in), 1.0 / math.sqrt(fan_in))
            self._register(name, nn.Parameter(init))
        self._handle = None
        self._handle_key = None
        self._workspaces = {}
        self._keepalive = None
        self._weights_epoch = 0      # bumped by DenoisingTrainer.optimizer_step (in-place kernel updates of the weights)

    def _register(self, dotted, param):
        node = self
        parts = dotted.split(".")
        for p in parts[:-1]:
            if p not in node._modules:
                node.add_module(p, _Node())
            node = node._modules[p]
        node.register_parameter(parts[-1], param)

    # ---- C handle management ---------------------------------------------------------------------------
    def _config_struct(self):
        c = UnetConfigStruct()
        c.im_channels, c.im_size, c.time_emb_dim = self.im_channels, self.im_size, self.t_emb_dim
        c.num_down_layers, c.num_mid_layers, c.num_up_layers = self.num_down_layers, self.num_mid_layers, self.num_up_layers
        c.num_heads = self.num_heads
        c.n_down_channels = len(self.down_channels)
        c.n_mid_channels = len(self.mid_channels)
        c.n_attn_resolutions = len(self.attn_resolutions)
        for i, v in enumerate(self.down_channels):
            c.down_channels[i] = v
        for i, v in enumerate(self.mid_channels):
            c.mid_channels[i] = v
        for i, v in enumerate(self.down_sample):
            c.down_sample[i] = 1 if v else 0
        for i, v in enumerate(self.attn_resolutions):
            c.attn_resolutions[i] = v
        return c

    def _destroy(self):
        if self._handle is not None:
            lib().wc_unet_destroy(self._handle)
            self._handle = None

    def __del__(self):
        try:
            self._destroy()
        except Exception:
            pass

    def _ensure_handle(self, device):
        params = dict(self.named_parameters())
        key = (str(device), self._weights_epoch, tuple((p.data_ptr(), p._version) for p in params.values()))
        if self._handle is not None and key == self._handle_key:
            return
        self._destroy()
        names, tensors = [], []
        for n, p in params.items():
            if p.device != device or p.dtype != torch.float32 or not p.is_contiguous():
                raise RuntimeError(f"Unet parameter {n} must be a contiguous fp32 tensor on {device}; call .to(device)")
            names.append(n.encode())
            tensors.append(p.detach())
        n = len(names)
        c_names = (C.c_char_p * n)(*names)
        c_ptrs = (C.c_void_p * n)(*[t.data_ptr() for t in tensors])
        c_numels = (C.c_int64 * n)(*[t.numel() for t in tensors])
        handle = C.c_void_p()
        cfg = self._config_struct()
        set_time_factor_table(self.t_emb_dim, device)
        check(lib().wc_unet_create(C.byref(handle), C.byref(cfg), n, c_names, c_ptrs, c_numels, stream_ptr()))
        self._handle, self._handle_key, self._keepalive = handle, key, tensors
        self._workspaces = {}

    def _workspace(self, B, H, W, device):
        k = (B, H, W)
        ws = self._workspaces.get(k)
        if ws is None:
            nbytes = lib().wc_unet_workspace_bytes(self._handle, B, H, W)
            if nbytes == 0:
                check(1)
            self._workspaces = {}        # one live binding at a time: the C plan is bound to one workspace
            ws = torch.empty(nbytes, dtype=torch.uint8, device=device)
            self._workspaces[k] = ws
        return ws

    # ---- reference API ---------------------------------------------------------------------------------
    def forward(self, x, t, out=None):
        _lib.require_cuda(x)
        x = x.contiguous().float()
        B, Cc, H, W = x.shape
        if Cc != self.im_channels:
            raise RuntimeError(f"expected {self.im_channels} input channels, got {Cc}")
        t = torch.as_tensor(t).long().to(x.device).reshape(-1)     # reference :461
        if t.numel() not in (1, B):
            raise RuntimeError("t must have 1 or batch entries")
        self._ensure_handle(x.device)
        ws = self._workspace(B, H, W, x.device)
        if out is None:
            out = torch.empty_like(x)
        check(lib().wc_unet_forward(self._handle, ptr(x), ptr(t), t.numel(), ptr(out), B, H, W, ptr(ws), ws.numel(),
                                    stream_ptr()))
        self._last_io = (x, t)   # keep inputs alive until the stream has consumed them
        return out

    def flops_per_forward(self) -> float:
        """Algorithmic FLOPs (2*MAC) of the last bound forward."""
        return float(lib().wc_unet_flops(self._handle)) if self._handle is not None else 0.0

    def launches_per_forward(self) -> int:
        return int(lib().wc_unet_launches(self._handle)) if self._handle is not None else 0


_FACTOR_SET = {}


def set_time_factor_table(temb_dim, device):
    """Hand the reference's factor table (reference :22-24, evaluated by torch on the host exactly as written there) to the
    library once per (device, dim): every time-embedding kernel then divides t by the reference's own fp32 factors."""
    key = (str(device), temb_dim)
    if _FACTOR_SET.get("key") == key:
        return
    half = temb_dim // 2
    factor = 10000 ** (torch.arange(start=0, end=half, dtype=torch.float32) / half)
    arr = (C.c_float * half)(*factor.tolist())
    with torch.cuda.device(device):
        check(lib().wc_set_time_factor_table(arr, half))
    _FACTOR_SET["key"] = key


def get_time_embedding(time_steps, temb_dim):
    """Sinusoidal embedding (reference :7-30) as its own kernel (csrc/direct.cu: time_embedding_kernel - the same code
    the UNet plan runs inside temb_mlp_kernel): [B] int tensor on the GPU -> [B, temb_dim] fp32."""
    assert temb_dim % 2 == 0, "time embedding dimension must be divisible by 2"
    _lib.require_cuda(time_steps)
    t = time_steps.long().contiguous().reshape(-1)
    set_time_factor_table(temb_dim, t.device)
    out = torch.empty(t.numel(), temb_dim, device=t.device)
    check(lib().wc_time_embedding(ptr(t), t.numel(), temb_dim, ptr(out), stream_ptr()))
    return out
