"""Oracle: image / label edges of the loop in numpy (TEST INFRASTRUCTURE, see oracle/__init__.py).

Restates, for byte-exact comparison with the GPU kernels of csrc/io_kernels.cu:
  * seg_model/inference.py:75-82,103 (ExtResize(just_label) NEAREST + ExtCenterCrop + ExtToTensor + ExtNormalize) and
    seg_model/datasets/acdc.py:21-57,135-138 (id_to_train_id LUT, encode_target)
  * translation.py:138-145 (Resize BILINEAR + CenterCrop + ToTensor + x*2-1); the resize is Pillow's two-pass 8-bit
    resampling (third-party dependency: Pillow, src/libImaging/Resample.c precompute_coeffs / normalize_coeffs_8bpc /
    ImagingResampleHorizontal_8bpc / Vertical_8bpc; NEAREST = Geometry.c ImagingScaleAffine), restated from its published
    algorithm and pinned against the Pillow build in this image by tests/golden/make_golden.py:g_io
  * sample_ddpm.py:47-51 (clamp, (x+1)/2, torchvision make_grid, ToPILImage) and sample_integrated.py:32-37 (postprocess)
"""
import math

import numpy as np

# acdc.py:21-57, train ids in list order (the last entry is 'license plate', id -1)
ID_TO_TRAIN_ID = np.array([255, 255, 255, 255, 255, 255, 255, 0, 1, 255, 255, 2, 3, 4, 255, 255, 255, 5, 255, 6, 7, 8, 9, 10, 11,
                           12, 13, 14, 15, 255, 255, 16, 17, 18, 255], dtype=np.int64)
PRECISION_BITS = 32 - 8 - 2


def resized_size(h, w, size):                                    # torchvision F.resize
    if isinstance(size, (tuple, list)):
        return int(size[0]), int(size[1])
    return (int(size * h / w), int(size)) if w <= h else (int(size), int(size * w / h))


def center_crop(a, ch, cw):                                      # torchvision F.center_crop
    h, w = a.shape[:2]
    top, left = int(round((h - ch) / 2.0)), int(round((w - cw) / 2.0))
    return a[top:top + ch, left:left + cw]


def nearest_resize(a, oh, ow):                                   # Geometry.c ImagingScaleAffine
    def tab(n_in, n_out):
        step = float(n_in) / float(n_out)
        xo, out = step * 0.5, []
        for _ in range(n_out):
            out.append(min(max(int(xo) if xo >= 0.0 else -1, 0), n_in - 1))
            xo += step
        return np.array(out)
    return a[tab(a.shape[0], oh)][:, tab(a.shape[1], ow)]


def _coeffs(n_in, n_out):                                        # Resample.c precompute_coeffs + normalize_coeffs_8bpc
    scale = float(n_in) / float(n_out)
    fs = max(scale, 1.0)
    support = fs
    res = []
    for xx in range(n_out):
        center = (xx + 0.5) * scale
        xmin = max(int(center - support + 0.5), 0)
        xmax = min(int(center + support + 0.5), n_in) - xmin
        w = [max(0.0, 1.0 - abs((x + xmin - center + 0.5) / fs)) for x in range(xmax)]
        ww = 0.0
        for v in w:
            ww += v
        k = [int(0.5 + (v / ww if ww != 0.0 else v) * (1 << PRECISION_BITS)) for v in w]
        res.append((xmin, np.array(k, dtype=np.int64)))
    return res


def bilinear_resize_u8(img, oh, ow):                             # ImagingResampleHorizontal_8bpc then Vertical_8bpc
    H, W, Cc = img.shape
    tmp = np.zeros((H, ow, Cc), np.uint8)
    for xx, (xmin, k) in enumerate(_coeffs(W, ow)):
        ss = (img[:, xmin:xmin + len(k), :].astype(np.int64) * k[None, :, None]).sum(1) + (1 << (PRECISION_BITS - 1))
        tmp[:, xx, :] = np.clip(ss >> PRECISION_BITS, 0, 255)
    out = np.zeros((oh, ow, Cc), np.uint8)
    for yy, (ymin, k) in enumerate(_coeffs(H, oh)):
        ss = (tmp[ymin:ymin + len(k)].astype(np.int64) * k[:, None, None]).sum(0) + (1 << (PRECISION_BITS - 1))
        out[yy] = np.clip(ss >> PRECISION_BITS, 0, 255)
    return out


def encode_label(label_ids, resize=(540, 960), crop=(512, 512)):          # inference.py:75-78,103 + acdc.py:135-138
    oh, ow = resized_size(label_ids.shape[0], label_ids.shape[1], resize)
    return ID_TO_TRAIN_ID[center_crop(nearest_resize(label_ids, oh, ow), crop[0], crop[1])][None]


def diffusion_input(img, image_size=128):                                  # translation.py:138-145
    oh, ow = resized_size(img.shape[0], img.shape[1], image_size)
    r = bilinear_resize_u8(img, oh, ow) if (oh, ow) != img.shape[:2] else img
    c = center_crop(r, image_size, image_size).astype(np.float32) / np.float32(255.0)
    return (c.transpose(2, 0, 1) * np.float32(2.0) - np.float32(1.0))[None]


def normalize_image(img, mean=(0.485, 0.456, 0.406), std=(0.229, 0.224, 0.225)):   # inference.py:79-80
    t = img.astype(np.float32).transpose(2, 0, 1) / np.float32(255.0)
    m, s = np.array(mean, np.float32)[:, None, None], np.array(std, np.float32)[:, None, None]
    return ((t - m) / s)[None]


def ddpm_grid_uint8(xt, nrow, padding=2):                                  # sample_ddpm.py:47-51 (+ torchvision make_grid)
    ims = (np.clip(xt.astype(np.float32), -1.0, 1.0) + np.float32(1.0)) / np.float32(2.0)
    B, Cc, H, W = ims.shape
    if B == 1:
        grid = ims[0]
    else:
        xmaps = min(nrow, B)
        ymaps = int(math.ceil(float(B) / xmaps))
        grid = np.zeros((Cc, (H + padding) * ymaps + padding, (W + padding) * xmaps + padding), np.float32)
        for k in range(B):
            y, x = k // xmaps, k % xmaps
            grid[:, y * (H + padding) + padding:y * (H + padding) + padding + H, x * (W + padding) + padding:x * (W + padding) + padding + W] = ims[k]
    return (grid * np.float32(255.0)).astype(np.uint8).transpose(1, 2, 0)   # ToPILImage: mul(255).byte(), HWC


def postprocess_uint8(xt, mean=(0.4865, 0.4998, 0.4323), std=(0.2326, 0.2276, 0.2659)):    # sample_integrated.py:32-37
    m, s = np.array(mean, np.float32)[None, :, None, None], np.array(std, np.float32)[None, :, None, None]
    images = xt.astype(np.float32) * s + m
    return np.clip(images * np.float32(255.0), 0, 255).astype(np.uint8)
