"""Training-step results must not depend on the previous contents of freshly allocated device memory: pollute the caching
allocator's pool with different byte patterns, build a trainer, run steps, compare."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import torch
from oracle.unet import DEFAULT_MODEL_CONFIG
from weatherconverter_b200.diffusion_model.train_ddpm import DenoisingTrainer
from test_gpu_train import _build
dev = torch.device("cuda")
cfg = dict(DEFAULT_MODEL_CONFIG); cfg["im_size"] = 32
g = torch.Generator().manual_seed(3)
images = (torch.rand(2, 3, 32, 64, generator=g) * 2 - 1).to(dev)
noise = torch.randn(2, 3, 32, 64, generator=g).to(dev)
t = torch.tensor([100, 600])

def pollute(kind):
    torch.cuda.empty_cache()
    bufs = []
    for _ in range(6):
        x = torch.empty(1 << 28, dtype=torch.float32, device=dev)   # 1 GiB each
        if kind == "nan": x.fill_(float("nan"))
        elif kind == "zero": x.zero_()
        elif kind == "big": x.fill_(3.0e38)
        else: x.normal_()
        bufs.append(x)
    # a spread of small blocks too (small-pool allocations)
    small = [torch.full((n,), float("nan") if kind == "nan" else 1e30, device=dev) for n in (64, 256, 1024, 4096, 65536, 1 << 20) for _ in range(8)]
    del bufs, small

res = {}
for kind in ["zero", "nan", "rand", "big"]:
    pollute(kind)
    _, model, sched = _build(cfg, 7, dev)
    tr = DenoisingTrainer(model, sched, lr=1e-4)
    ls = [float(tr.step(images, noise=noise, t=t)) for _ in range(3)]
    res[kind] = (ls, tr.flat_params.clone())
    print(kind, ls, "finite params:", bool(torch.isfinite(tr.flat_params).all()), flush=True)
    del tr, model
for kind in ["nan", "rand", "big"]:
    print(kind, "== zero:", torch.equal(res[kind][1], res["zero"][1]), "max diff", float((res[kind][1] - res["zero"][1]).abs().max()))
