"""Per-launch table of every non-igemm kernel class for one eager step of a bench workload (CUDA-event timing via wc_profile_detail;
info = the six integers each launcher passes to ProfScope::note, e.g. attention: B*heads, tokens, head_dim)."""
import sys, os, ctypes as C, collections
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench
from weatherconverter_b200 import _lib
wl_name = sys.argv[1] if len(sys.argv) > 1 else "c3"
dev = torch.device("cuda")
wl = bench.GpuWorkload(dict(bench.WORKLOADS[wl_name]), dev, 0)
z = torch.randn_like(wl.x0)
xt = wl.x0.clone()
for k in range(3):
    xt = wl.step(xt, k, z)
torch.cuda.synchronize()
lib = _lib.lib()
lib.wc_profile_begin()
xt = wl.step(xt, 3, z)
cap = 4096
cls, ms, work, info = (C.c_int * cap)(), (C.c_double * cap)(), (C.c_double * cap)(), (C.c_int * (cap * 6))()
n = lib.wc_profile_detail(cap, cls, ms, work, info)
m8, c8, w8 = (C.c_double * 8)(), (C.c_longlong * 8)(), (C.c_double * 8)()
lib.wc_profile_end(m8, c8, w8)
groups = collections.defaultdict(lambda: [0, 0.0])
for i in range(min(n, cap)):
    if cls[i] != 0:
        key = (cls[i],) + tuple(info[i * 6 + j] for j in range(4))
        groups[key][0] += 1; groups[key][1] += ms[i] * 1e3
print("class  info(4)                      n    total us   us each")
for key, g in sorted(groups.items(), key=lambda kv: -kv[1][1])[:40]:
    print(f"{key[0]:5d}  {str(key[1:]):28s} {g[0]:3d} {g[1]:10.1f} {g[1]/g[0]:9.1f}")
