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


def sample_with_sgg(unet_sd, unet_cfg, sched, seg_sd, srgan_sd, x0, gt, noise, t_fwd, zs,
                    lam=60.0, n_steps=500, backbone="resnet50", guidance=True, record=None):
    """Repaired translation driver (translation.py:46-97 + SURVEY.md 8c repairs).
    x0 [B,3,h,w] in [-1,1]; gt [B,4h,4w] int64; noise like x0; t_fwd [B] int64 (translation.py:63);
    zs: list of n_steps tensors like x0, zs[i] is the z drawn at reverse step i (scheduler.py:110).
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
                    xt, _, _ = apply_gsg(seg_sd, mu, sigma, sr_xt, gt, lam, backbone)   # :87 (every step)
            else:
                xt = mu + sigma                                                   # :90 (reference quirk D1)
            if record is not None:
                record.append(xt.clone())
        return srgan.generator_forward(srgan_sd, xt)                              # :95
