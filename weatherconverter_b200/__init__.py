"""B200-native DDPM reverse sampling + semantic-gradient guidance (drop-in for the hot path of
xXCoffeeColaXc/WeatherConverter).  Sub-packages mirror the reference's module paths:

    weatherconverter_b200.diffusion_model.models.unet_base.Unet
    weatherconverter_b200.diffusion_model.scheduler.linear_noise_scheduler.LinearNoiseScheduler
    weatherconverter_b200.diffusion_model.sample_ddpm.sample / load_model / load_scheduler

All arithmetic runs in libwc_b200.so (hand-written sm_100a CUDA, C ABI in include/wc_b200.h).
"""
__version__ = "0.1.0"
