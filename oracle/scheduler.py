"""Oracle: linear-beta DDPM scheduler (TEST INFRASTRUCTURE, see oracle/__init__.py).

Follows diffusion_model/scheduler/linear_noise_scheduler.py of the reference:
  tables            :16-21
  add_noise2        :30-35      add_noise :37-61
  sample_prev_timestep2 :63-77  sample_prev_timestep :79-116
The only extension is an injected ``z`` (the reference draws it from the CPU generator, :76/:110).
"""
import torch


class OracleScheduler:
    def __init__(self, num_timesteps, beta_start, beta_end):
        self.num_timesteps = num_timesteps
        self.betas = torch.linspace(beta_start, beta_end, num_timesteps)          # :16
        self.alphas = 1. - self.betas                                             # :17
        self.alpha_cum_prod = torch.cumprod(self.alphas, dim=0)                   # :18
        self.sqrt_alpha_cum_prod = torch.sqrt(self.alpha_cum_prod)                # :19
        self.one_minus_cum_prod = 1 - self.alpha_cum_prod                         # :20
        self.sqrt_one_minus_alpha_cum_prod = torch.sqrt(1 - self.alpha_cum_prod)  # :21

    def add_noise(self, original, noise, t):                                      # :37-61 / :30-35
        a = self.sqrt_alpha_cum_prod[t].reshape(-1, 1, 1, 1)
        b = self.sqrt_one_minus_alpha_cum_prod[t].reshape(-1, 1, 1, 1)
        return a * original + b * noise

    add_noise2 = add_noise

    def sample_prev_timestep(self, xt, noise_pred, t, z=None):                    # :79-116
        t = int(t)
        mean = xt - (self.betas[t] * noise_pred) / self.sqrt_one_minus_alpha_cum_prod[t]   # :96-98
        mean = mean / torch.sqrt(self.alphas[t])                                  # :100
        if t == 0:
            return mean, None, None                                               # :102-103
        variance = (1 - self.alpha_cum_prod[t - 1]) / (1.0 - self.alpha_cum_prod[t])       # :107
        variance = variance * self.betas[t]                                       # :108
        sigma = variance ** 0.5                                                   # :109
        if z is None:
            z = torch.randn(xt.shape)                                             # :110
        return mean, sigma * z, None

    def sample_prev_timestep2(self, xt, noise_pred, t, z=None):                   # :63-77
        beta = self.betas[t].view(-1, 1, 1, 1)
        alpha = self.alphas[t].view(-1, 1, 1, 1)
        s = self.sqrt_one_minus_alpha_cum_prod[t].view(-1, 1, 1, 1)
        mean = xt - ((beta * noise_pred) / s)
        mean = mean / torch.sqrt(alpha)
        if torch.all(t == 0):
            return mean, None, None
        sigma = beta ** 0.5
        if z is None:
            z = torch.randn(xt.shape)
        return mean, sigma * z, None
