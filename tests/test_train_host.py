"""CPU tests of the data-parallel host logic of the training step (bucket layout, overlap schedule, all-reduce over gloo
with world size 2).  The CUDA plan itself is covered by the -m gpu tests."""
import os

import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from weatherconverter_b200.diffusion_model.models.unet_base import param_spec
from weatherconverter_b200.diffusion_model.train_ddpm import _backward_group, backward_with_buckets, make_buckets
from oracle.unet import DEFAULT_MODEL_CONFIG


def _layout():
    names = list(param_spec(DEFAULT_MODEL_CONFIG))
    numel = {k: int(torch.Size(v).numel()) for k, v in param_spec(DEFAULT_MODEL_CONFIG).items()}
    order = sorted(range(len(names)), key=lambda i: (_backward_group(names[i], 4, 2, 4), i))
    slices, groups, off = {}, [], 0
    for i in order:
        off = (off + 3) & ~3
        slices[names[i]] = (off, numel[names[i]])
        groups.append(_backward_group(names[i], 4, 2, 4))
        off += numel[names[i]]
    return [names[i] for i in order], slices, groups, off


def test_backward_groups_and_buckets():
    names, slices, groups, used = _layout()
    assert len(names) == 382 and groups == sorted(groups)
    assert names[0].startswith(("norm_out", "conv_out")) and "t_emb_layers" in names[-1] or names[-1].startswith("t_proj")
    # synthetic ready-op marks that follow the backward order
    ready = [10 * (g + 1) for g in groups]
    n_ops = max(ready)
    buckets = make_buckets(names, slices, groups, ready, n_ops, used, 64 << 20)
    assert buckets[0][0] == 0 and buckets[-1][1] == used and buckets[-1][2] == n_ops
    for (a, b, r), (a2, b2, r2) in zip(buckets, buckets[1:]):
        assert b == a2 and r <= r2 and (b - a) * 4 >= (64 << 20)
    # every parameter lies in exactly one bucket, and that bucket closes no earlier than the parameter is ready
    for i, n in enumerate(names):
        off, cnt = slices[n]
        owner = [bk for bk in buckets if bk[0] <= off and off + cnt <= bk[1]]
        assert len(owner) == 1 and owner[0][2] >= ready[i], n


def _worker(rank, world, port, q):
    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    n_ops, used = 12, 1000
    buckets = [(0, 300, 4), (300, 700, 9), (700, 1000, 12)]
    flat = torch.zeros(used)
    done = []

    def run_ops(a, b):      # op k finalises elements [k*used/n_ops, (k+1)*used/n_ops) with a rank-dependent value
        for k in range(a, b):
            lo, hi = k * used // n_ops, (k + 1) * used // n_ops
            flat[lo:hi] = (rank + 1) * (k + 1)
            done.append(k)
    backward_with_buckets(run_ops, n_ops, flat, buckets, world)
    expect = torch.zeros(used)
    for k in range(n_ops):
        expect[k * used // n_ops:(k + 1) * used // n_ops] = sum((r + 1) * (k + 1) for r in range(world))
    q.put((rank, bool(torch.equal(flat, expect)), done == list(range(n_ops))))
    dist.destroy_process_group()


def test_bucketed_allreduce_gloo_world2():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + os.getpid() % 500
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in procs]
    for p in procs:
        p.join(timeout=60)
    assert all(ok and ordered for _, ok, ordered in res), res
