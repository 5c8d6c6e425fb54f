"""Micro-benchmark of the CUDA-core boundary kernels at C3 sizes (CUDA events, warm)."""
import sys, os, math
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from weatherconverter_b200 import ops
from weatherconverter_b200._lib import check, lib, ptr, stream_ptr
dev = torch.device("cuda")
B = 32

def timeit(name, fn, it=5):
    for _ in range(2): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(it): fn()
    e1.record(); torch.cuda.synchronize()
    print(f"{name}: {e0.elapsed_time(e1)/it:.3f} ms", flush=True)

x = torch.rand(B, 3, 256, 512, device=dev)
w7 = torch.randn(64, 3, 7, 7, device=dev) / 12
sc, sh = torch.ones(64, device=dev), torch.zeros(64, device=dev)
timeit("seg conv1 7x7 s2 (256x512 -> 128x256x64)", lambda: ops.conv_in(x, w7, None, sc, sh, stride=2, pad=3, relu=True))
xl = torch.rand(B, 3, 64, 128, device=dev)
w9 = torch.randn(64, 3, 9, 9, device=dev) / 15
timeit("srgan initial 9x9 (64x128)", lambda: ops.conv_in(xl, w9, sc, None, None, stride=1, pad=4))
w3 = torch.randn(64, 3, 3, 3, device=dev) / 5
timeit("unet conv_in 3x3 (64x128)", lambda: ops.conv_in(xl, w3, sc))
dz = torch.randn(B, 128, 256, 64, device=dev).bfloat16()
dx = torch.empty(B, 3, 256, 512, device=dev)
timeit("conv1 dgrad (128x256x64 -> 256x512x3)", lambda: check(lib().wc_conv1_dgrad(ptr(dz), ptr(w7), ptr(sc), ptr(dx), B, 256, 512, stream_ptr())))
h = torch.randn(B, 64, 128, 64, device=dev).bfloat16()
wo = torch.randn(3, 64, 3, 3, device=dev) / 24
timeit("unet conv_out 3x3 (64x128)", lambda: ops.conv_out(h, wo, None))
# maxpool fwd/bwd
c1 = torch.rand(B, 128, 256, 64, device=dev).bfloat16()
p1 = torch.empty(B, 64, 128, 64, device=dev, dtype=torch.bfloat16); idx = torch.empty(B, 64, 128, 64, device=dev, dtype=torch.uint8)
timeit("maxpool fwd", lambda: check(lib().wc_maxpool3x3s2(ptr(c1), ptr(p1), ptr(idx), B, 128, 256, 64, stream_ptr())))
dc1 = torch.empty_like(c1)
timeit("maxpool bwd", lambda: check(lib().wc_maxpool3x3s2_bwd(ptr(p1), ptr(idx), ptr(c1), ptr(dc1), B, 128, 256, 64, stream_ptr())))
a = torch.randn(B, 16, 32, 256, device=dev).bfloat16(); up = torch.empty(B, 64, 128, 256, device=dev, dtype=torch.bfloat16)
timeit("bilinear fwd 16x32->64x128 x256", lambda: check(lib().wc_bilinear(ptr(a), ptr(up), B, 16, 32, 64, 128, 256, stream_ptr())))
da = torch.empty_like(a)
timeit("bilinear bwd", lambda: check(lib().wc_bilinear_bwd(ptr(up), None, ptr(da), B, 16, 32, 64, 128, 256, stream_ptr())))
lo = torch.randn(B, 19, 64, 128, device=dev); lab = torch.randint(0, 19, (B, 256, 512), device=dev)
pred = torch.empty(B, 256, 512, dtype=torch.long, device=dev); dhi = torch.empty(B, 19, 256, 512, device=dev)
loss = torch.empty(B, device=dev); nv = torch.empty(B, dtype=torch.int32, device=dev); dlo = torch.empty(B, 64, 128, 32, device=dev, dtype=torch.bfloat16)
timeit("loss head (fwd+adjoint)", lambda: check(lib().wc_seg_loss_head(ptr(lo), ptr(lab), ptr(nv), ptr(pred), ptr(dhi), ptr(loss), None, ptr(dlo), B, 64, 128, 256, 512, stream_ptr())))
# srgan full forward
from weatherconverter_b200.srgan_model.models import Generator
G = Generator().to(dev).eval()
timeit("srgan full forward (64x128 -> 256x512)", lambda: G(xl))
import ctypes as C
l = lib(); l.wc_profile_begin(); G(xl)
ms, cnt, wk = (C.c_double * 8)(), (C.c_longlong * 8)(), (C.c_double * 8)()
l.wc_profile_end(ms, cnt, wk)
print("srgan classes ms:", [round(v, 3) for v in ms], [int(c) for c in cnt])
