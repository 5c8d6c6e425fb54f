from .linear_noise_scheduler import LinearNoiseScheduler  # noqa: F401
