"""Sweep of the igemm epilogue with a residual input over batch sizes / shapes (op-level, vs torch fp32 on the GPU)."""
import math, os, sys
import torch
import torch.nn.functional as F
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from weatherconverter_b200 import ops
torch.backends.cudnn.allow_tf32 = False
torch.backends.cuda.matmul.allow_tf32 = False
dev = torch.device("cuda")
bf = lambda x: x.to(torch.bfloat16).float()
bad = 0
SHAPES = ((768, 768, 16, 32, 1), (512, 512, 16, 32, 1), (256, 256, 64, 128, 1), (512, 512, 32, 64, 1), (64, 64, 64, 128, 1),
          (128, 128, 64, 128, 1), (64, 64, 64, 128, 3), (256, 256, 16, 32, 3))
BATCHES = (1, 2, 3, 4, 5, 8, 16, 32)
if os.environ.get("SWEEP_SHORT"):
    SHAPES, BATCHES = ((768, 768, 16, 32, 1), (256, 256, 64, 128, 1), (512, 512, 32, 64, 1)), (3, 5, 16, 32)
for (Cin, Cout, H, W, K) in SHAPES:
    for B in BATCHES:
        if B * H * W * max(Cin, Cout) > 3e8:
            continue
        g = torch.Generator().manual_seed(Cin + B)
        x = torch.randn(B, Cin, H, W, generator=g).to(dev)
        w = (torch.randn(Cout, Cin, K, K, generator=g) / math.sqrt(Cin * K * K)).to(dev)
        b = (0.1 * torch.randn(Cout, generator=g)).to(dev)
        res = torch.randn(B, Cout, H, W, generator=g).to(dev)
        ref = F.conv2d(bf(x), bf(w), b, padding=K // 2) + bf(res)
        worst = 0.0
        for rep in range(3):
            y = ops.to_nchw_f32(ops.conv2d(ops.to_nhwc_bf16(x), w, b, residual=ops.to_nhwc_bf16(res)))
            per_img = [(float((y[i] - ref[i]).norm() / ref[i].norm())) for i in range(B)]
            worst = max(worst, max(per_img))
        flag = "" if worst < 5e-3 else "   <-- BAD"
        bad += worst >= 5e-3
        print(f"Cin {Cin} Cout {Cout} {H}x{W} K{K} B={B}: worst per-image rel {worst:.2e}{flag}")
print("BAD cases:", bad)
