"""Per-launch igemm table for one step of a bench workload (CUDA-event timing through wc_profile_detail)."""
import sys, os, ctypes as C, collections
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench
from weatherconverter_b200 import _lib
wl_name = sys.argv[1] if len(sys.argv) > 1 else "c3"
dev = torch.device("cuda")
spec = dict(bench.WORKLOADS[wl_name])
wl = bench.GpuWorkload(spec, dev, 0)
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
rows = []
for i in range(min(n, cap)):
    if cls[i] == 0:
        M, N, K, BN, taps, grid = [info[i * 6 + j] for j in range(6)]
        rows.append((i, ms[i] * 1e3, work[i] / (ms[i] * 1e-3) / 1e12, M, N, K, BN, taps, grid, work[i] / 1e9))
tot = sum(r[1] for r in rows)
print(f"{len(rows)} igemm launches, {tot/1e3:.2f} ms, {sum(r[9] for r in rows)/tot*1e3:.0f} TFLOP/s avg")
groups = collections.defaultdict(lambda: [0, 0.0, 0.0])
for r in rows:
    key = (r[3], r[4], r[5], r[6], r[7])
    g = groups[key]; g[0] += 1; g[1] += r[1]; g[2] += r[9]
print(f"{'us':>9} {'share':>6} {'n':>3} {'TF/s':>6} {'floor us':>9} {'x floor':>7}  M N K BN taps   (floor = max(FLOPs / 1370 TF/s, bytes / 6553 GB/s) per launch)")
fl_tot = 0.0
for key, g in sorted(groups.items(), key=lambda kv: -kv[1][1])[:60]:
    M, N, K, BN, taps = key
    t = max(1, taps % 100)
    fl_ = taps // 1000
    key = key + (('lean ' if fl_ & 1 else '') + ('deep ' if fl_ & 32 else '') + ('mask ' if fl_ & 2 else '') + ('res ' if fl_ & 4 else '') + ('qkv/f32 ' if fl_ & 8 else '') + ('no-tma-store' if fl_ & 16 else ''),)
    fl = max(g[2] / g[0] * 1e9 / 1370e12, 2.0 * (M * K / t + N * K + M * N) / 6553e9) * 1e6 * g[0]
    fl_tot += fl
    print(f"{g[1]:9.1f} {100*g[1]/tot:5.1f}% {g[0]:3d} {g[2]/g[1]*1e3:6.0f} {fl:9.1f} {g[1]/fl:7.2f}  {key}")
print(f"sum of floors (listed): {fl_tot/1e3:.2f} ms")
