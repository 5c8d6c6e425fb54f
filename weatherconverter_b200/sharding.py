"""Batch sharding for independent sampling chains: one process per GPU, no data-path collective.

Each image is its own Markov chain, so rank r simply owns a contiguous slice of the global batch and the RNG stream
of every image is keyed by its GLOBAL index: results are identical for any number of GPUs.
"""
import torch


def shard_range(global_batch: int, world_size: int, rank: int):
    """Contiguous [start, stop) slice of the global batch owned by `rank` (sizes differ by at most one)."""
    if not (0 <= rank < world_size):
        raise ValueError("rank out of range")
    base, rem = divmod(global_batch, world_size)
    start = rank * base + min(rank, rem)
    return start, start + base + (1 if rank < rem else 0)


def image_generator(global_index: int, seed: int = 1234) -> torch.Generator:
    """CPU generator of one image's chain (reference draws all noise from the CPU generator)."""
    g = torch.Generator()
    g.manual_seed(seed * 1000003 + global_index)
    return g


def initial_noise(shape_chw, start: int, stop: int, seed: int = 1234) -> torch.Tensor:
    """x_T for images [start, stop): stack of per-image draws (independent of world size)."""
    return torch.stack([torch.randn(shape_chw, generator=image_generator(i, seed)) for i in range(start, stop)])


def gather_images(local: torch.Tensor, world_size: int, global_batch: int = None):
    """Host-side gather of finished images to rank 0 (the only collective; after the timed region).  Shards may be uneven
    (shard_range: sizes differ by at most one): every rank pads its shard to the largest shard size, rank 0 slices the
    padding off again.  ``global_batch`` defaults to world_size * len(local) (even shards)."""
    import torch.distributed as dist
    if world_size == 1:
        return [local]
    rank = dist.get_rank()
    if global_batch is None:
        global_batch = world_size * local.shape[0]
    sizes = [b - a for a, b in (shard_range(global_batch, world_size, r) for r in range(world_size))]
    if local.shape[0] != sizes[rank]:
        raise ValueError(f"rank {rank} holds {local.shape[0]} images, shard_range says {sizes[rank]}")
    biggest = max(sizes)
    padded = local
    if local.shape[0] < biggest:
        padded = torch.cat([local, local.new_zeros((biggest - local.shape[0],) + tuple(local.shape[1:]))])
    out = [torch.empty_like(padded) for _ in range(world_size)] if rank == 0 else None
    dist.gather(padded.contiguous(), out, dst=0)
    return [o[:n] for o, n in zip(out, sizes)] if rank == 0 else None
