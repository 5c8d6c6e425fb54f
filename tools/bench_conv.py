"""Micro-benchmark of single igemm convolutions (per-launch CUDA events via wc_profile_detail)."""
import sys, os, ctypes as C, math
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from weatherconverter_b200 import ops, _lib
dev = torch.device("cuda")
# B, Cin, Cout, H, W, K, with_res
cases = [(32, 64, 64, 64, 128, 3, 1), (32, 64, 64, 64, 128, 3, 0), (32, 64, 256, 64, 128, 1, 1), (32, 256, 64, 64, 128, 1, 0),
         (16, 64, 64, 128, 256, 3, 0), (16, 128, 128, 128, 256, 3, 0), (16, 256, 256, 64, 128, 3, 0), (16, 768, 768, 16, 32, 3, 0)]
if len(sys.argv) > 1:
    cases = [tuple(int(v) for v in a.split(",")) for a in sys.argv[1:]]
lib = _lib.lib()
for (B, Cin, Cout, H, W, K, with_res) in cases:
    x = torch.randn(B, H, W, Cin, device=dev).bfloat16()
    w = torch.randn(Cout, Cin, K, K, device=dev) / math.sqrt(Cin * K * K)
    b = torch.zeros(Cout, device=dev)
    res = torch.randn(B, H, W, Cout, device=dev).bfloat16() if with_res else None
    for _ in range(2):
        ops.conv2d(x, w, b, residual=res)
    torch.cuda.synchronize()
    lib.wc_profile_begin()
    for _ in range(5):
        ops.conv2d(x, w, b, residual=res)
    cap = 256
    cls, ms, work, info = (C.c_int * cap)(), (C.c_double * cap)(), (C.c_double * cap)(), (C.c_int * (cap * 6))()
    n = lib.wc_profile_detail(cap, cls, ms, work, info)
    m8, c8, w8 = (C.c_double * 8)(), (C.c_longlong * 8)(), (C.c_double * 8)()
    lib.wc_profile_end(m8, c8, w8)
    ts = [ms[i] for i in range(n) if cls[i] == 0]
    t = min(ts)
    fl = 2.0 * B * H * W * Cin * Cout * K * K
    byts = B * H * W * 2.0 * (Cin + Cout * (2 if with_res else 1))
    print(f"B{B} {Cin}->{Cout} {H}x{W} k{K} res{with_res}: {t*1e3:.1f} us  {fl/t/1e9:.0f} TFLOP/s  {byts/t/1e6:.0f} GB/s (algorithmic)  BN={info[3]} row3={info[4]//100}", flush=True)
