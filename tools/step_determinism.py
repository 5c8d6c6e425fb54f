"""Run-to-run determinism of one full reverse step of a bench workload at its benchmarked batch: the same step from the same state,
R times, must give bit-identical x_(t-1) (the round-2 race in the igemm epilogue showed up exactly here: batch 32 was not
reproducible).  usage: python tools/step_determinism.py [c3|c2|ref] [R]"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench
name = sys.argv[1] if len(sys.argv) > 1 else "c3"
R = int(sys.argv[2]) if len(sys.argv) > 2 else 6
dev = torch.device("cuda")
wl = bench.GpuWorkload(dict(bench.WORKLOADS[name]), dev, 0)
z = torch.randn_like(wl.x0)
x = wl.x0.clone()
for k in range(2):
    x = wl.step(x, k, z)
torch.cuda.synchronize()
ref = wl.step(x.clone(), 2, z).clone()
bad = 0
for r in range(R):
    out = wl.step(x.clone(), 2, z)
    if not torch.equal(out, ref):
        bad += 1
        print(f"run {r}: max-abs diff {float((out - ref).abs().max()):.3e}")
print(f"{name}: batch {wl.x0.shape[0]}, {R} repeats of one reverse step, {bad} differ bit-wise; finite: {bool(torch.isfinite(ref).all())}")
sys.exit(1 if bad else 0)
