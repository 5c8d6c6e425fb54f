"""Runs each tensor-core op repeatedly on fixed inputs and reports launches whose output differs bit-wise from the first."""
import sys, os, math
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from weatherconverter_b200 import ops
dev = torch.device("cuda")
g = torch.Generator().manual_seed(0)
R = int(os.environ.get("REPS", "30"))

def check(name, fn):
    ref = fn()
    ref = [r.clone() for r in (ref if isinstance(ref, tuple) else (ref,))]
    bad = 0
    for _ in range(R):
        out = fn()
        out = out if isinstance(out, tuple) else (out,)
        if not all(torch.equal(a, b) for a, b in zip(out, ref)):
            bad += 1
    print(f"{name}: {bad}/{R} runs differ", flush=True)

for (B, h, N, hd) in [(2, 4, 2048, 16), (2, 4, 2048, 32), (2, 4, 512, 64), (2, 4, 128, 128), (2, 4, 32, 192), (2, 4, 128, 16), (2, 4, 32, 32),
                      (4, 4, 8192, 16)]:
    q = torch.randn(B, h, N, hd, generator=g).to(dev).bfloat16()
    k = torch.randn(B, h, N, hd, generator=g).to(dev).bfloat16()
    v = torch.randn(B, h, N, hd, generator=g).to(dev).bfloat16()
    vt = v.transpose(2, 3).contiguous()
    check(f"attention_lse N{N} hd{hd}", lambda: ops.attention_lse(q, k, vt))
    o, lse = ops.attention_lse(q, k, vt)
    do = torch.randn(B, N, h * hd, generator=g).to(dev).bfloat16()
    check(f"attention_bwd N{N} hd{hd}", lambda: ops.attention_bwd(q, k, v, o, do, lse))
for (B, Cin, Cout, H, W, K) in [(2, 64, 64, 32, 64, 3), (2, 128, 256, 16, 32, 3), (2, 64, 192, 32, 64, 1), (2, 768, 768, 4, 8, 3)]:
    x = torch.randn(B, H, W, Cin, generator=g).to(dev).bfloat16()
    w = (torch.randn(Cout, Cin, K, K, generator=g) / math.sqrt(Cin * K * K)).to(dev)
    b = torch.zeros(Cout, device=dev)
    check(f"conv2d {Cin}->{Cout} {H}x{W} k{K}", lambda: ops.conv2d(x, w, b))
    dy = torch.randn(B, H, W, Cout, generator=g).to(dev).bfloat16()
    check(f"wgrad {Cin}->{Cout} {H}x{W} k{K}", lambda: ops.conv2d_wgrad(x, dy, K))
