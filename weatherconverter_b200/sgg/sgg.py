"""Semantic-gradient guidance with the reference's entry point (sgg/sgg.py:9-24).

``apply_gsg(seg_model, mu, sigma, sr_xt, gt, _lambda)`` = segmentation forward + CE + input gradient
(csrc/seg.cu), then ONE fused kernel for ``avg_pool2d(4)`` -> de-normalise -> channel L2 (float64) ->
``mu + lambda*sigma*mag + sigma`` (csrc/elementwise.cu: sgg_update_kernel).  The reference returns float64
(numpy promotion); this returns fp32 rounded from the same float64 arithmetic (the repaired driver casts
back to fp32 anyway).  ``apply_lcg`` raises as shipped; here it runs with the documented repair of its final masked sum.
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


def apply_lcg(seg_model, mu, sigma, sr_xt, gt, _lambda, num_classes=19, pool=None):
    """Local class guidance (reference sgg.py:27-60).  The per-class body is the reference's: mask the image and
    the labels with mc = (gt == c), take the segmentation-loss gradient, pool, magnitude.  The shipped final sum
    multiplies 128x128 tensors by 512x512 masks and raises; the repair (documented in DESIGN.md section 4) average-pools the class masks to the latent resolution and leaves pixels without a class
    unguided:  xt = mu + sigma + lambda * sigma * sum_c avg_pool(mc)_c * |g4_c|.
    All classes of all images run as ONE batch of B*19 masked images through the segmentor plan."""
    _lib.require_cuda(mu, sigma, sr_xt, gt)
    B, _, h, w = mu.shape
    Hs, Ws = sr_xt.shape[-2:]
    if pool is None:
        pool = Ws // w
    sr_xt, gt = sr_xt.contiguous().float(), gt.contiguous().long()
    xm = torch.empty(B * num_classes, 3, Hs, Ws, device=mu.device)
    gm = torch.empty(B * num_classes, Hs, Ws, dtype=torch.long, device=mu.device)
    check(lib().wc_lcg_prepare(ptr(sr_xt), ptr(gt), ptr(xm), ptr(gm), B, num_classes, Hs, Ws, stream_ptr()))
    g4 = infer_batch(seg_model, xm, gm, grad_pool=pool)["grad"]
    mu, sigma = mu.contiguous().float(), sigma.contiguous().float()
    xt = torch.empty_like(mu)
    check(lib().wc_lcg_combine(ptr(g4), ptr(gt), ptr(mu), ptr(sigma), ptr(xt), B, num_classes, h, w, pool, float(_lambda),
                               stream_ptr()))
    return xt
