import sys, os, math
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
dev = torch.device("cuda")
tr = torch.zeros(3 * 2000 * 2, dtype=torch.int64, device=dev)
os.environ["WC_IGEMM_TRACE"] = hex(tr.data_ptr())
from weatherconverter_b200 import ops
B, Cin, Cout, H, W, K = 32, 64, 64, 64, 128, 3
x = torch.randn(B, H, W, Cin, device=dev).bfloat16()
w = torch.randn(Cout, Cin, K, K, device=dev) / math.sqrt(Cin * K * K)
for _ in range(3):
    tr.zero_()
    ops.conv2d(x, w, None)
torch.cuda.synchronize()
t = tr.cpu().view(3, 2000, 2)
t0 = min(int(t[r, 0, 1]) for r in range(3) if int(t[r, 0, 1]) > 0)
for r, name in enumerate(("producer", "mma", "epilogue")):
    ev = [(int(a), int(b) - t0) for a, b in t[r].tolist() if b > 0]
    print(name, len(ev), "events; first 40:", " ".join(f"{a}@{b}" for a, b in ev[:40]))
    if len(ev) > 60:
        print("   ... mid:", " ".join(f"{a}@{b}" for a, b in ev[len(ev)//2:len(ev)//2+24]))
