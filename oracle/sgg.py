"""Oracle: semantic-gradient guidance + repaired translation driver
(TEST INFRASTRUCTURE, see oracle/__init__.py).

  compute_gradient_magnitude   seg_model/inference.py:36-53 (float64 on host, B=1 semantics)
  apply_gsg                    sgg/sgg.py:9-24
  sample_with_sgg (repaired)   translation.py:46-97 with the minimal repairs of SURVEY.md 8c:
       no_grad outside infer, line :90 deleted, xt = mu at i == 0, guided xt cast back to fp32,
       GSG on every step, B > 1 = per-image loop of the B = 1 reference.
"""
import numpy as np
import torch
import torch.nn.functional as F

from . import deeplab, srgan, unet

STD = np.array([0.229, 0.224, 0.225])


def compute_gradient_magnitude(g):
    """g [1,3,h,w] fp32 -> [h,w] float64 (inference.py:39-43, denormalize=True, norm=False)."""
    gn = g.squeeze(0).cpu().numpy() * STD[:, None, None]
    return torch.from_numpy(np.sqrt(np.sum(gn ** 2, axis=0)))


def apply_gsg(seg_sd, mu, sigma, sr_xt, gt, lam, backbone="resnet50", pool=4):
    """Per-image GSG (sgg.py:9-24); returns (xt fp64->fp32, pred, grad)."""
    outs, preds, grads = [], [], []
    for b in range(mu.shape[0]):
        pred, g, _ = deeplab.infer(seg_sd, sr_xt[b:b + 1], gt[b:b + 1], backbone)
        g4 = F.avg_pool2d(g, pool, pool) if pool > 1 else g                       # sgg.py:18
        mag = compute_gradient_magnitude(g4)                                      # sgg.py:19
        mu_hat = mu[b:b + 1] + lam * sigma[b:b + 1] * mag                         # sgg.py:21 (fp64)
        outs.append((mu_hat + sigma[b:b + 1]).float())                            # sgg.py:22
        preds.append(pred); grads.append(g)
    return torch.cat(outs), torch.cat(preds), torch.cat(grads)


def apply_lcg(seg_sd, mu, sigma, sr_xt, gt, lam, backbone="resnet50", pool=4, num_classes=19):
    """Repaired local class guidance (sgg/sgg.py:27-60; the shipped version raises).

    Per class c (sgg.py:39-54, verbatim): mc = (gt == c); gradient of CE(seg(sr_xt * mc), gt * mc) w.r.t. the masked
    input; avg_pool2d(4); float64 magnitude; xt_c = mu + lambda * sigma * mag_c + sigma.
    Repair of the final masked sum (sgg.py:58 multiplies 128x128 tensors by 512x512 masks): the class masks are
    average-pooled to the latent resolution, mc_lat = avg_pool2d(mc, pool), and pixels not covered by any class
    (label 255) keep the unguided value:
        xt = sum_c mc_lat_c * xt_c + (1 - sum_c mc_lat_c) * (mu + sigma) = mu + sigma + lambda * sigma * sum_c mc_lat_c * mag_c
    Returns xt (fp32, like the repaired driver casts it)."""
    outs = []
    for b in range(mu.shape[0]):
        acc = torch.zeros(mu.shape[2:], dtype=torch.float64)
        for c in range(num_classes):
            mc = (gt[b:b + 1] == c).long().unsqueeze(1)                              # sgg.py:41
            x_m = sr_xt[b:b + 1] * mc                                                # sgg.py:44
            gt_m = gt[b:b + 1] * mc.squeeze(1)                                       # sgg.py:45
            _, g, _ = deeplab.infer(seg_sd, x_m, gt_m, backbone)                     # sgg.py:47
            mag = compute_gradient_magnitude(F.avg_pool2d(g, pool, pool))            # sgg.py:49-50
            mc_lat = F.avg_pool2d(mc.float(), pool, pool)[0, 0].double()
            acc = acc + mc_lat * mag
        outs.append(((mu[b:b + 1] + lam * sigma[b:b + 1] * acc) + sigma[b:b + 1]).float())
    return torch.cat(outs)


def sample_with_sgg(unet_sd, unet_cfg, sched, seg_sd, srgan_sd, x0, gt, noise, t_fwd, zs,
                    lam=60.0, n_steps=500, backbone="resnet50", guidance=True, record=None, mode="gsg"):
    """Repaired translation driver (translation.py:46-97 + SURVEY.md 8c repairs).
    x0 [B,3,h,w] in [-1,1]; gt [B,4h,4w] int64; noise like x0; t_fwd [B] int64 (translation.py:63);
    zs: list of n_steps tensors like x0, zs[i] is the z drawn at reverse step i (scheduler.py:110).
    mode: "gsg" = GSG on every step (repaired default); "alternate" = the shipped schedule :84-87 (LCG on even, GSG on odd i).
    Returns sr_x0 [B,3,4h,4w] in [0,1]."""
    with torch.no_grad():
        xt = sched.add_noise2(x0, noise, t_fwd)                                   # :65
        for i in reversed(range(n_steps)):                                        # :70
            eps = unet.unet_forward(unet_sd, unet_cfg, xt, torch.tensor([i]))    # :74
            mu, sigma, _ = sched.sample_prev_timestep(xt, eps, i, z=zs[i])        # :78
            if i == 0:
                xt = mu                                                           # repair of D2
            elif guidance:
                sr_xt = srgan.generator_forward(srgan_sd, xt)                     # :81
                with torch.enable_grad():
                    if mode == "lcg" or (mode == "alternate" and i % 2 == 0):
                        xt = apply_lcg(seg_sd, mu, sigma, sr_xt, gt, lam, backbone)     # :84-85 (even steps)
                    else:
                        xt, _, _ = apply_gsg(seg_sd, mu, sigma, sr_xt, gt, lam, backbone)   # :86-87
            else:
                xt = mu + sigma                                                   # :90 (reference quirk D1)
            if record is not None:
                record.append(xt.clone())
        return srgan.generator_forward(srgan_sd, xt)                              # :95
