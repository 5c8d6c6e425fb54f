"""Oracle: one denoising-loss training step (TEST INFRASTRUCTURE, see oracle/__init__.py).

fp32 restatement of the step body of diffusion_model/train_ddpm.py:95-114 of the reference:
  noisy = scheduler.add_noise(im, noise, t) (:105) ; pred = model(noisy, t) (:106) ; loss = MSELoss(pred, noise) (:108)
  loss.backward() (:109) ; Adam(lr).step() (:113, optimizer built at :151 with torch defaults betas (0.9, 0.999), eps 1e-8)
The network is oracle.unet.unet_forward; gradients come from torch autograd over that fp32 restatement; Adam is
restated explicitly (torch.optim.Adam's single-tensor update, no weight decay / amsgrad).
"""
import torch

from .scheduler import OracleScheduler
from .unet import unet_forward


def adam_update(p, g, m, v, step, lr=1e-4, beta1=0.9, beta2=0.999, eps=1e-8):
    """torch.optim.Adam step for one tensor (in place on p, m, v)."""
    m.lerp_(g, 1 - beta1)
    v.mul_(beta2).addcmul_(g, g, value=1 - beta2)
    bc1 = 1 - beta1 ** step
    bc2 = 1 - beta2 ** step
    denom = (v.sqrt() / (bc2 ** 0.5)).add_(eps)
    p.addcdiv_(m, denom, value=-(lr / bc1))


def train_step(sd, cfg, images, noise, t, sched: OracleScheduler, lr=1e-4, adam_state=None, step=1):
    """Returns (loss, grads dict, new state dict).  sd is not modified."""
    params = {k: v.clone().requires_grad_(True) for k, v in sd.items()}
    noisy = sched.add_noise(images, noise, t)
    pred = unet_forward(params, cfg, noisy, t)
    loss = torch.nn.functional.mse_loss(pred, noise)
    loss.backward()
    grads = {k: p.grad.detach() for k, p in params.items()}
    new_sd = {}
    if adam_state is None:
        adam_state = {k: (torch.zeros_like(v), torch.zeros_like(v)) for k, v in sd.items()}
    for k, v in sd.items():
        p = v.clone()
        m, vv = adam_state[k]
        adam_update(p, grads[k], m, vv, step, lr)
        new_sd[k] = p
    return loss.detach(), grads, new_sd
