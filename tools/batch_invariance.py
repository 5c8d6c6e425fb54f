"""Diagnostic: error of the UNet forward vs the fp32 oracle as a function of the batch size and of the image's position in the
batch (the chains are independent, so any dependence is a bug or a precision difference between kernel paths)."""
import os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle.unet import DEFAULT_MODEL_CONFIG, unet_forward
from oracle.weights import synth_state_dict
from weatherconverter_b200.diffusion_model.models.unet_base import Unet, param_spec

dev = torch.device("cuda")
ims, H, W = int(sys.argv[1]), int(sys.argv[2]), int(sys.argv[3])
cfg = dict(DEFAULT_MODEL_CONFIG); cfg["im_size"] = ims
sd = synth_state_dict({k: (v, torch.float32) for k, v in param_spec(cfg).items()}, 3455)
m = Unet(cfg).to(dev).eval(); m.load_state_dict(sd)
g = torch.Generator().manual_seed(1234)
x = torch.randn(1, 3, H, W, generator=g)
t = torch.tensor([637])
with torch.no_grad():
    ref = unet_forward(sd, cfg, x, t)
others = torch.randn(32, 3, H, W, generator=g)
y1 = m(x.to(dev), t.to(dev)).cpu()
print(f"B=1: rel {float((y1-ref).norm()/ref.norm()):.3e}")
Bs = [int(v) for v in sys.argv[4].split(',')] if len(sys.argv) > 4 else [2, 3, 4, 8, 16, 32]
for B in Bs:
    for pos in sorted({0, B // 2, B - 1}):
        xb = others[:B].clone(); xb[pos] = x[0]
        yb = m(xb.to(dev), t.to(dev))[pos:pos + 1].cpu()
        print(f"B={B} pos={pos}: rel vs oracle {float((yb-ref).norm()/ref.norm()):.3e}  vs B=1 {float((yb-y1).norm()/y1.norm()):.3e}")
