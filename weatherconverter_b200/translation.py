"""SGG-guided translation driver with the reference's entry point (translation.py:46-97).

``sample_with_sgg(input_tensor, diff_model, diff_scheduler, seg_model, gt, srgan_model)`` keeps the reference's
signature and constants (LAMBDA = 60, N = 500).  The driver as shipped cannot finish (D1 the guided x_t
is overwritten, D2 ``mu + None`` at i = 0, D3 LCG shape error on the 2nd iteration, D7 float64 promotion, D8 autograd
graph kept across steps), so the default here is the REPAIRED loop of SURVEY.md 8c: no autograd graph, global
guidance (GSG) on every step i > 0, ``x_t = mu`` at i = 0, guided x_t kept (fp32).  ``reference_quirks=True``
reproduces D1 verbatim (guidance computed, then discarded) for step-level comparison with the shipped code.
Batches are independent chains: image b is guided by its own loss/gradient (vmap of the reference's B = 1 call).
"""
import torch

from .sgg.sgg import apply_gsg_batch
from .srgan_model.inference import inference as srgan_inference

device = torch.device('cuda' if torch.cuda.is_available() else 'cpu')

LAMBDA = 60.0
N_STEPS = 500


def load_input_image(path: str, image_size: int = 128) -> torch.Tensor:
    """translation.py:131,138-145: open the RGB image (host, PIL decode) and run Resize(BILINEAR) + CenterCrop + ToTensor +
    x*2-1 on the GPU, bit-identical to the reference's torchvision/PIL transform (weatherconverter_b200/image_io.py)."""
    import numpy as np
    from PIL import Image
    from . import image_io
    img = torch.from_numpy(np.array(Image.open(path).convert("RGB"))).to(device)
    return image_io.diffusion_input(img, image_size)


@torch.no_grad()
def sample_with_sgg(input_tensor, diff_model, diff_scheduler, seg_model, gt, srgan_model, *, n_steps=N_STEPS,
                    lam=LAMBDA, noise=None, t_forward=None, step_noise=None, device_noise=False, generator=None,
                    guidance=True, reference_quirks=False, record=None):
    """input_tensor [B,3,h,w] in [-1,1]; gt [B,4h,4w] int64 trainIds (255 = ignore).  Returns sr_x0 [B,3,4h,4w].
    Injection points for parity runs: ``t_forward`` [B] (reference :63 draws randint(0, N)), ``noise`` like the input
    (:64), ``step_noise`` [N, B,3,h,w] or callable i -> z (scheduler.py:110)."""
    x0 = input_tensor.to(device).float().contiguous()
    gt = gt.to(device)
    B = x0.shape[0]
    t = t_forward if t_forward is not None else torch.randint(0, n_steps, (B,))            # reference :63
    noise = noise if noise is not None else torch.randn_like(x0)                            # reference :64
    xt = diff_scheduler.add_noise2(x0, noise.to(device), t.to(device))                      # reference :65
    eps = torch.empty_like(xt)
    for i in reversed(range(n_steps)):                                                      # reference :70
        diff_model(xt, torch.as_tensor(i).unsqueeze(0).to(device), out=eps)                # reference :74
        if i == 0:
            xt = diff_scheduler.step(xt, eps, 0)                                            # repair of D2: x_0 = mu
        else:
            if step_noise is not None:
                z = (step_noise(i) if callable(step_noise) else step_noise[i]).to(device)
            elif device_noise:
                z = torch.randn(xt.shape, device=xt.device, generator=generator)
            else:
                z = None
            mu, sigma, _ = diff_scheduler.sample_prev_timestep(xt, eps, i, z=z)             # reference :78
            if guidance:
                sr_xt = srgan_inference(srgan_model, xt)                                    # reference :81
                guided = apply_gsg_batch(seg_model, mu, sigma, sr_xt, gt, lam)              # reference :87
                xt = (mu + sigma) if reference_quirks else guided                           # reference :90 (D1)
            else:
                xt = mu + sigma
        if record is not None:
            record.append(xt.clone())
    return srgan_inference(srgan_model, xt)                                                 # reference :95-97
