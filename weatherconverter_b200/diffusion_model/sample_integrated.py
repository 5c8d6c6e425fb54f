"""Sampling with the legacy UNet, the reference's entry points (diffusion_model/sample_integrated.py:32-75).

``sample`` keeps the reference signature and side effect (writes a PNG grid); ``sample_tensor`` is the same reverse
loop (batched t, noise-variance conditioning, sample_prev_timestep2) returning the final tensor, with injected noise
for parity tests.
"""
import os
from datetime import datetime

import torch

from .config.models import Config, DiffusionConfig, ModelConfig, TrainingConfig
from .models.old_modules import UNet
from .sample_ddpm import load_config, load_scheduler  # noqa: F401  (same helpers as the reference re-exports)
from .scheduler.linear_noise_scheduler import LinearNoiseScheduler

device = torch.device('cuda' if torch.cuda.is_available() else 'cpu')


def postprocess(xt, mean=[0.4865, 0.4998, 0.4323], std=[0.2326, 0.2276, 0.2659]):
    """Reference :32-37: de-normalise, scale to 0..255, clamp, uint8, to the host — one byte-exact kernel on the GPU
    (raises for CPU tensors: there is no CPU path)."""
    from ..image_io import postprocess_uint8
    return postprocess_uint8(xt, mean, std).cpu()


@torch.no_grad()
def sample_tensor(model: UNet, scheduler: LinearNoiseScheduler, shape=None, num_timesteps=None, xT=None, noise=None,
                  record=None):
    """Reverse loop of sample_integrated.py:52-65.  noise: optional [T, *shape] tensor, noise[i] is the z of step i."""
    T = num_timesteps if num_timesteps is not None else scheduler.num_timesteps
    if xT is None:
        xT = torch.randn(shape).to(device)                                        # :52 (CPU generator, then H2D)
    xt = xT.to(device).float().contiguous()
    eps = torch.empty_like(xt)
    var_table = scheduler.one_minus_cum_prod.to(xt.device)
    for i in reversed(range(T)):
        t = torch.full((xt.size(0),), i, dtype=torch.long).to(xt.device)          # :57
        model(xt, var_table[t].view(-1, 1, 1, 1), out=eps)                        # :60
        z = noise[i].to(xt.device) if (noise is not None and i != 0) else None
        mean, sigma, _ = scheduler.sample_prev_timestep2(xt, eps, t, z=z)         # :62
        xt = mean + sigma if i != 0 else mean                                     # :64
        if record is not None:
            record.append(xt.clone())
    return xt


def sample(model: UNet, scheduler: LinearNoiseScheduler, train_config: TrainingConfig, model_config: ModelConfig,
           diffusion_config: DiffusionConfig, save_path: str = 'diffusion_model_v2/outputs/samples'):
    xt = sample_tensor(model, scheduler,
                       (train_config.sample_size, model_config.im_channels, model_config.im_size, model_config.im_size),
                       num_timesteps=diffusion_config.num_timesteps)
    images = postprocess(xt)
    import torchvision
    from torchvision.utils import make_grid
    grid = make_grid(images, nrow=train_config.num_grid_rows)
    img = torchvision.transforms.ToPILImage()(grid)
    os.makedirs(save_path, exist_ok=True)
    now = datetime.now()
    img.save(os.path.join(save_path, f'old_x_{now.hour}:{now.minute}:{now.second}.png'))
    img.close()


def load_model(model_path: str) -> torch.nn.Module:
    model = UNet().to(device)
    checkpoint = torch.load(model_path, map_location=device)
    model.load_state_dict(checkpoint['model_state_dict'])
    model.eval()
    return model


def infer(config: Config):
    checkpoint_path = os.path.join(config.folders.checkpoints, 'old_model/1000-checkpoint.ckpt')
    model = load_model(checkpoint_path)
    scheduler = load_scheduler(config.diffusion)
    with torch.no_grad():
        sample(model, scheduler, config.training, config.model, config.diffusion, os.path.join(config.folders.samples, 'old_model'))


if __name__ == '__main__':
    infer(load_config('diffusion_model/config/config.yaml'))
