import sys, os, math
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from weatherconverter_b200 import ops
dev = torch.device("cuda")
B, h, N, hd = 32, 4, 8192, 16
q = (torch.randn(B, h, N, hd, device=dev) * (math.log2(math.e) / 4)).bfloat16()
k = torch.randn(B, h, N, hd, device=dev).bfloat16()
vt = torch.randn(B, h, hd, N, device=dev).bfloat16()
for _ in range(2): ops.attention(q, k, vt, scale=-1.0)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(5): ops.attention(q, k, vt, scale=-1.0)
e1.record(); torch.cuda.synchronize()
print(os.environ.get("WC_ATTN_SMALL4_POLY"), f"{e0.elapsed_time(e1)/5:.3f} ms")
