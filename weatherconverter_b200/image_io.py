"""Image / label edges of the sampling loop on the GPU, byte-exact against the reference's host path.

What the reference does on the CPU with PIL / numpy / torchvision right before and after the hot path:
  * seg_model/inference.py:75-82,103 + acdc.py:135-138: label NEAREST resize to (540, 960), centre crop 512, labelIds ->
    trainIds;  ExtToTensor + ExtNormalize of the image              -> ``encode_label`` / ``normalize_image``
  * translation.py:138-145: Resize(128, BILINEAR) + CenterCrop(128) + ToTensor + x*2-1        -> ``diffusion_input``
  * sample_ddpm.py:47-51: clamp, (x+1)/2, make_grid, ToPILImage                                -> ``ddpm_grid_uint8``
  * sample_integrated.py:32-37: postprocess                                                    -> ``postprocess_uint8``
The arithmetic runs in csrc/io_kernels.cu; this module only builds the small index / coefficient tables exactly as
Pillow's C code does (double precision on the host) and owns the device buffers.
"""
import ctypes as C
import math

import numpy as np
import torch

from . import _lib
from ._lib import check, lib, ptr, stream_ptr

# acdc.py:21-57: (id, train_id) of the 35 Cityscapes classes in list order; id_to_train_id = [c.train_id for c in classes]
ID_TO_TRAIN_ID = (255, 255, 255, 255, 255, 255, 255, 0, 1, 255, 255, 2, 3, 4, 255, 255, 255, 5, 255, 6, 7, 8, 9, 10, 11, 12,
                  13, 14, 15, 255, 255, 16, 17, 18, 255)
_PRECISION_BITS = 32 - 8 - 2


def resized_size(h, w, size):
    """torchvision F.resize output size: (h, w) tuple as given; int = smaller edge, long edge int(size * long / short)."""
    if isinstance(size, (tuple, list)):
        return int(size[0]), int(size[1])
    if w <= h:
        return int(size * h / w), int(size)
    return int(size), int(size * w / h)


def center_crop_offsets(h, w, ch, cw):
    """torchvision F.center_crop: top/left = int(round((dim - crop) / 2.0)) (Python's round-half-even)."""
    return int(round((h - ch) / 2.0)), int(round((w - cw) / 2.0))


def nearest_index_table(in_size, out_size):
    """Pillow Geometry.c ImagingScaleAffine: xo = a*0.5; for each x: xin = (int)xo; xo += a  (incremental double sum)."""
    a = float(in_size) / float(out_size)
    xo = a * 0.5
    tab = np.empty(out_size, dtype=np.int32)
    for x in range(out_size):
        xin = -1 if xo < 0.0 else int(xo)
        tab[x] = min(max(xin, 0), in_size - 1)
        xo += a
    return tab


def bilinear_coeffs(in_size, out_size):
    """Pillow Resample.c precompute_coeffs + normalize_coeffs_8bpc for the BILINEAR (triangle) filter, support 1.0."""
    scale = float(in_size) / float(out_size)
    filterscale = max(scale, 1.0)
    support = 1.0 * filterscale
    ksize = int(math.ceil(support)) * 2 + 1
    bounds = np.zeros(2 * out_size, dtype=np.int32)
    kk = np.zeros((out_size, ksize), dtype=np.int32)
    ss = 1.0 / filterscale
    for xx in range(out_size):
        center = (xx + 0.5) * scale
        xmin = max(int(center - support + 0.5), 0)
        xmax = min(int(center + support + 0.5), in_size) - xmin
        w = []
        for x in range(xmax):
            v = (x + xmin - center + 0.5) * ss
            v = -v if v < 0.0 else v
            w.append(1.0 - v if v < 1.0 else 0.0)
        ww = 0.0
        for v in w:
            ww += v
        for x in range(xmax):
            c = w[x] / ww if ww != 0.0 else w[x]
            kk[xx, x] = int(-0.5 + c * (1 << _PRECISION_BITS)) if c < 0 else int(0.5 + c * (1 << _PRECISION_BITS))
        bounds[2 * xx], bounds[2 * xx + 1] = xmin, xmax
    return bounds, kk.reshape(-1), ksize


def _f3(v):
    return (C.c_float * 3)(*[float(x) for x in v])


def encode_label(label_ids: torch.Tensor, resize=(1080 // 2, 1920 // 2), crop=(512, 512)) -> torch.Tensor:
    """label_ids uint8 [Hs,Ws] (gt_labelIds PNG) on the GPU -> encoded int64 [1,Hc,Wc] (seg_model/inference.py:75-78,103-104)."""
    _lib.require_cuda(label_ids)
    assert label_ids.dtype == torch.uint8 and label_ids.dim() == 2
    label_ids = label_ids.contiguous()
    Hs, Ws = label_ids.shape
    Hr, Wr = resized_size(Hs, Ws, resize)
    top, left = center_crop_offsets(Hr, Wr, crop[0], crop[1])
    dev = label_ids.device
    ytab = torch.from_numpy(nearest_index_table(Hs, Hr)).to(dev)
    xtab = torch.from_numpy(nearest_index_table(Ws, Wr)).to(dev)
    lut = torch.tensor(ID_TO_TRAIN_ID, dtype=torch.int64, device=dev)
    out = torch.empty(1, crop[0], crop[1], dtype=torch.int64, device=dev)
    check(lib().wc_label_encode(ptr(label_ids), Ws, ptr(ytab), ptr(xtab), top, left, crop[0], crop[1], ptr(lut), lut.numel(), ptr(out),
                                stream_ptr()))
    return out


def resize_bilinear_u8(img: torch.Tensor, out_h: int, out_w: int) -> torch.Tensor:
    """PIL ``Image.resize((out_w, out_h), BILINEAR)`` of an HWC uint8 image on the GPU (bit-identical)."""
    _lib.require_cuda(img)
    assert img.dtype == torch.uint8 and img.dim() == 3
    img = img.contiguous()
    H, W, Cc = img.shape
    dev = img.device
    bh, kh, ksh = bilinear_coeffs(W, out_w)
    bv, kv, ksv = bilinear_coeffs(H, out_h)
    t = [torch.from_numpy(a).to(dev) for a in (bh, kh, bv, kv)]
    tmp = torch.empty(H, out_w, Cc, dtype=torch.uint8, device=dev)
    out = torch.empty(out_h, out_w, Cc, dtype=torch.uint8, device=dev)
    check(lib().wc_resample_u8(ptr(img), ptr(tmp), ptr(out), H, W, out_h, out_w, Cc, ptr(t[0]), ptr(t[1]), ksh, ptr(t[2]), ptr(t[3]), ksv,
                               stream_ptr()))
    return out


def diffusion_input(img: torch.Tensor, image_size=128) -> torch.Tensor:
    """translation.py:138-145: Resize(image_size, BILINEAR) + CenterCrop + ToTensor + x*2-1; img HWC uint8 -> [1,3,S,S] fp32."""
    H, W, _ = img.shape
    Hr, Wr = resized_size(H, W, image_size)
    r = resize_bilinear_u8(img, Hr, Wr) if (Hr, Wr) != (H, W) else img.contiguous()
    top, left = center_crop_offsets(Hr, Wr, image_size, image_size)
    out = torch.empty(1, 3, image_size, image_size, dtype=torch.float32, device=img.device)
    check(lib().wc_u8_to_tensor(ptr(r), Wr, top, left, image_size, image_size, 0, None, None, ptr(out), stream_ptr()))
    return out


def normalize_image(img: torch.Tensor, mean=(0.485, 0.456, 0.406), std=(0.229, 0.224, 0.225)) -> torch.Tensor:
    """ExtToTensor + ExtNormalize (seg_model/inference.py:79-80): img HWC uint8 -> [1,3,H,W] fp32."""
    _lib.require_cuda(img)
    img = img.contiguous()
    H, W, _ = img.shape
    out = torch.empty(1, 3, H, W, dtype=torch.float32, device=img.device)
    check(lib().wc_u8_to_tensor(ptr(img), W, 0, 0, H, W, 1, _f3(mean), _f3(std), ptr(out), stream_ptr()))
    return out


def ddpm_grid_uint8(xt: torch.Tensor, nrow: int, padding: int = 2) -> torch.Tensor:
    """sample_ddpm.py:47-51 on the GPU: returns the HWC uint8 array of the PIL image the reference saves."""
    _lib.require_cuda(xt)
    xt = xt.contiguous().float()
    B, Cc, H, W = xt.shape
    assert Cc == 3
    xmaps = min(nrow, B)
    ymaps = int(math.ceil(float(B) / xmaps))
    Hg, Wg = (H, W) if B == 1 else ((H + padding) * ymaps + padding, (W + padding) * xmaps + padding)
    out = torch.empty(Hg, Wg, 3, dtype=torch.uint8, device=xt.device)
    check(lib().wc_ddpm_grid_u8(ptr(xt), ptr(out), B, H, W, nrow, padding, stream_ptr()))
    return out


def postprocess_uint8(xt: torch.Tensor, mean=(0.4865, 0.4998, 0.4323), std=(0.2326, 0.2276, 0.2659)) -> torch.Tensor:
    """sample_integrated.py:32-37 on the GPU: [B,3,H,W] fp32 -> uint8 (same layout)."""
    _lib.require_cuda(xt)
    xt = xt.contiguous().float()
    B, Cc, H, W = xt.shape
    out = torch.empty(B, 3, H, W, dtype=torch.uint8, device=xt.device)
    check(lib().wc_postprocess_u8(ptr(xt), ptr(out), B, H, W, _f3(mean), _f3(std), stream_ptr()))
    return out
