"""2-GPU data-parallel parity of the training step: after one step, the weights of a 2-rank run (each rank
half of the batch, NCCL all-reduce of the gradients) must equal a 1-process run on the concatenated batch.
Launch:  python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 tools/ddp_check.py"""
import os
import sys

import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle.unet import DEFAULT_MODEL_CONFIG  # noqa: E402
from oracle.weights import synth_state_dict  # noqa: E402
from weatherconverter_b200.diffusion_model.models.unet_base import Unet, param_spec  # noqa: E402
from weatherconverter_b200.diffusion_model.scheduler.linear_noise_scheduler import LinearNoiseScheduler  # noqa: E402
from weatherconverter_b200.diffusion_model.train_ddpm import DenoisingTrainer  # noqa: E402


def main():
    rank, world, lr_ = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(lr_)
    dev = torch.device("cuda", lr_)
    dist.init_process_group("nccl", device_id=dev)
    cfg = dict(DEFAULT_MODEL_CONFIG); cfg["im_size"] = 64
    sd = synth_state_dict({k: (v, torch.float32) for k, v in param_spec(cfg).items()}, 11)
    g = torch.Generator().manual_seed(5)
    Bg = 2 * world
    images = torch.rand(Bg, 3, 64, 64, generator=g) * 2 - 1
    noise = torch.randn(Bg, 3, 64, 64, generator=g)
    t = torch.randint(0, 1000, (Bg,), generator=g)
    sched = LinearNoiseScheduler(1000, 1e-4, 0.02)

    def run(trainer, lo, hi, steps=2):
        for _ in range(steps):
            loss = trainer.step(images[lo:hi].to(dev), noise=noise[lo:hi].to(dev), t=t[lo:hi])
        torch.cuda.synchronize()
        return float(loss)

    # every rank deliberately starts from DIFFERENT weights: the trainer must broadcast rank 0's at construction (torch DDP's
    # contract) instead of relying on identical seeding
    sd_rank = sd if rank == 0 else synth_state_dict({k: (v, torch.float32) for k, v in param_spec(cfg).items()}, 11 + 97 * rank)
    model = Unet(cfg).to(dev); model.load_state_dict(sd_rank)
    tr = DenoisingTrainer(model, sched, lr=1e-4, bucket_bytes=32 << 20)
    ref0 = tr.flat_params.clone()
    dist.broadcast(ref0, 0)
    assert torch.equal(ref0, tr.flat_params), f"rank {rank}: parameters were not broadcast from rank 0"
    # default noise / timestep draws come from per-rank generators: the ranks must NOT draw the same t for their different images
    t_draw = torch.randint(0, 1000, (4,), generator=tr._gen_cpu)
    gathered = [torch.empty(4, dtype=torch.int64, device=dev) for _ in range(world)]
    dist.all_gather(gathered, t_draw.to(dev))
    assert world == 1 or not torch.equal(gathered[0], gathered[1]), "ranks drew identical timesteps"
    per = Bg // world
    loss = run(tr, rank * per, (rank + 1) * per)
    nb = len(tr._buckets)
    ok = True
    if rank == 0:
        model1 = Unet(cfg).to(dev); model1.load_state_dict(sd)
        tr1 = DenoisingTrainer(model1, sched, lr=1e-4, data_parallel=False)   # single-process reference on the concatenated batch
        loss1 = run(tr1, 0, Bg)
        dp, d1 = tr.flat_params, tr1.flat_params
        upd = (d1 - torch.cat([v.flatten() for v in []]) if False else None)
        diff = (dp - d1).abs()
        # Adam's step is ~lr per element; weights of the two runs may differ only by fp32 summation order effects
        frac = float((diff > 2e-5).float().mean())
        print(f"ddp_check: world {world}, {nb} buckets, rank-0 loss {loss:.6f}, full-batch loss {loss1:.6f}, "
              f"max |dw| {float(diff.max()):.3e}, fraction of weights differing by > 0.2 lr: {frac:.2e}")
        ok = frac < 2e-3
    flag = torch.tensor([1 if ok else 0], device=dev)
    dist.broadcast(flag, 0)
    # all ranks must hold identical replicas after the step
    ref = tr.flat_params.clone()
    dist.broadcast(ref, 0)
    same = bool(torch.equal(ref, tr.flat_params))
    if not same:
        print(f"rank {rank}: replica differs from rank 0")
    dist.destroy_process_group()
    sys.exit(0 if (int(flag) == 1 and same) else 1)


if __name__ == "__main__":
    main()
