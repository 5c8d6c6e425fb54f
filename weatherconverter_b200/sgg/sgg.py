"""Semantic-gradient guidance with the reference's entry point (sgg/sgg.py:9-24).

``apply_gsg(seg_model, mu, sigma, sr_xt, gt, _lambda)`` = segmentation forward + CE + input gradient
(csrc/seg.cu), then ONE fused kernel for ``avg_pool2d(4)`` -> de-normalise -> channel L2 (float64) ->
``mu + lambda*sigma*mag + sigma`` (csrc/elementwise.cu: sgg_update_kernel).  The reference returns float64
(numpy promotion); this returns fp32 rounded from the same float64 arithmetic (the repaired driver casts
back to fp32 anyway).  ``apply_lcg`` raises as shipped and is a later-round item (section 8f-4).
"""
import torch

from .. import _lib
from .._lib import check, lib, ptr, stream_ptr
from ..seg_model.inference import infer_batch


def apply_gsg_batch(seg_model, mu, sigma, sr_xt, gt, _lambda, pool=None, return_aux=False, fused_pool=True):
    """Batched GSG: every image uses its own loss/gradient (vmap of the reference's B = 1 semantics)."""
    _lib.require_cuda(mu, sigma, sr_xt, gt)
    B, _, h, w = mu.shape
    if pool is None:
        pool = sr_xt.shape[-1] // w
    # avg_pool2d(grad, pool) (sgg.py:18) is folded into the segmentor's stem data-gradient kernel when possible
    fused = fused_pool and pool in (2, 4, 6, 8)
    out = infer_batch(seg_model, sr_xt, gt, grad_pool=pool if fused else 1)
    mu, sigma = mu.contiguous().float(), sigma.contiguous().float()
    xt = torch.empty_like(mu)
    check(lib().wc_sgg_update(ptr(out["grad"]), ptr(mu), ptr(sigma), ptr(xt), None, B, h, w, 1 if fused else pool,
                              float(_lambda), stream_ptr()))
    return (xt, out) if return_aux else xt


def apply_gsg(seg_model, mu, sigma, sr_xt, gt, _lambda):
    return apply_gsg_batch(seg_model, mu, sigma, sr_xt, gt, _lambda)


def apply_lcg(seg_model, mu, sigma, sr_xt, gt, _lambda):
    raise NotImplementedError("apply_lcg raises in the reference as shipped (sgg.py:41,58 shape mismatch, SURVEY D3); "
                              "the repaired 19-way local guidance is a later-round item (SURVEY 8f-4)")
