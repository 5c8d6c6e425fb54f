"""Micro-benchmark of the attention backward kernels (CUDA events, warm, one shape per line): KV pass + Q pass + rowdot."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from weatherconverter_b200 import ops

shapes = [(16, 4, 8192, 64), (16, 4, 8192, 16), (16, 4, 2048, 128), (16, 4, 2048, 32), (16, 4, 512, 192)]
if len(sys.argv) > 1:
    shapes = [tuple(int(v) for v in a.split(",")) for a in sys.argv[1:]]
dev = torch.device("cuda")
for (B, h, N, hd) in shapes:
    q = torch.randn(B, h, N, hd, device=dev).bfloat16()
    k = torch.randn(B, h, N, hd, device=dev).bfloat16()
    v = torch.randn(B, h, N, hd, device=dev).bfloat16()
    d_o = torch.randn(B, N, h * hd, device=dev).bfloat16()
    o, lse = ops.attention_lse(q, k, v.transpose(2, 3).contiguous())
    for _ in range(2):
        ops.attention_bwd(q, k, v, o, d_o, lse)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    it = 3
    e0.record()
    for _ in range(it):
        ops.attention_bwd(q, k, v, o, d_o, lse)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / it
    fl = 14.0 * B * h * N * N * hd
    print(f"bwd B{B} h{h} N{N} hd{hd}: {ms:.3f} ms  {fl / ms / 1e9:.1f} TFLOP/s (7 GEMMs)  exps/s {2*B*h*N*N/ms/1e9:.2f} T", flush=True)
