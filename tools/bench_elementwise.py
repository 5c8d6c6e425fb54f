"""Bandwidth-bound kernels at the C3 shapes (for ncu captures): GroupNorm+SiLU, fused DDPM posterior update, guidance update."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from weatherconverter_b200 import ops
from weatherconverter_b200._lib import check, lib, ptr, stream_ptr
from weatherconverter_b200.diffusion_model.scheduler.linear_noise_scheduler import LinearNoiseScheduler
dev = torch.device("cuda")
B, h, w = 32, 64, 128
x = torch.randn(B, h, w, 64, device=dev).bfloat16()
gamma, beta = torch.ones(64, device=dev), torch.zeros(64, device=dev)
x2 = torch.randn(B, h, w, 128, device=dev).bfloat16()
g2, b2 = torch.ones(128, device=dev), torch.zeros(128, device=dev)
s = LinearNoiseScheduler(1000, 1e-4, 0.02)
xt, eps, z = (torch.randn(B, 3, h, w, device=dev) for _ in range(3))
grad = torch.randn(B, 3, h, w, device=dev) * 1e-4
mu, sig = torch.randn(B, 3, h, w, device=dev), 0.1 * torch.randn(B, 3, h, w, device=dev)
out = torch.empty_like(mu)
def run():
    ops.groupnorm_silu(x, gamma, beta)
    ops.groupnorm_silu(x2, g2, b2)
    s.step(xt, eps, 499, z=z)
    check(lib().wc_sgg_update(ptr(grad), ptr(mu), ptr(sig), ptr(out), None, B, h, w, 1, 60.0, stream_ptr()))
for _ in range(3):
    run()
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record(); run(); e1.record(); torch.cuda.synchronize()
print(f"one pass (2 GroupNorm, ddpm_step, sgg_update): {e0.elapsed_time(e1)*1e3:.1f} us")
