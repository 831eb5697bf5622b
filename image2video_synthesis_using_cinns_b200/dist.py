"""Multi-GPU: batch sharding + one all-gather of the finished frames.

The sampling path never mixes samples (GroupNorm / InstanceNorm are per sample, the embedder's BatchNorm runs
in eval mode, ActNorm is a fixed affine), so N GPUs = N independent shards of the start-frame batch with the
weights replicated.  The only exchange is the gather of finished frames; there is no data-path collective
inside the model and therefore nothing to fuse a kernel with.

To keep N-GPU outputs identical to the 1-GPU run, the residual is drawn ONCE for the global batch on the CPU
generator (quirk Q5, get_model.py:59) by every rank and then sliced.
"""
from __future__ import annotations

import torch
import torch.distributed as dist


def shard_bounds(n: int, world: int, rank: int):
    """Contiguous near-equal split of n rows: the first n % world ranks get one extra row."""
    base, extra = divmod(n, world)
    start = rank * base + min(rank, extra)
    return start, start + base + (1 if rank < extra else 0)


def sharded_sample(sample_fn, x_0, residual, cond=None, group=None):
    """Run ``sample_fn(x_shard, residual_shard, cond_shard) -> (b, T, C, H, W)`` on this rank's rows of the
    global batch and all-gather the result: every rank returns the full (B, T, C, H, W) tensor in the global
    row order.  Uneven shards are padded to the largest shard for the collective and trimmed afterwards."""
    world = dist.get_world_size(group) if dist.is_initialized() else 1
    rank = dist.get_rank(group) if dist.is_initialized() else 0
    n = x_0.shape[0]
    lo, hi = shard_bounds(n, world, rank)
    out = sample_fn(x_0[lo:hi], residual[lo:hi], None if cond is None else cond[lo:hi])
    if world == 1:
        return out
    max_rows = shard_bounds(n, world, 0)[1]
    if max_rows == 0:
        return out
    pad = out
    if out.shape[0] != max_rows:
        pad = out.new_zeros((max_rows,) + tuple(out.shape[1:]))
        pad[: out.shape[0]] = out
    pad = pad.contiguous()
    gathered = pad.new_empty((world * max_rows,) + tuple(out.shape[1:]))
    if dist.get_backend(group) == "nccl":
        dist.all_gather_into_tensor(gathered, pad, group=group)
    else:   # gloo (CPU tests)
        parts = list(gathered.chunk(world, dim=0))
        dist.all_gather(parts, pad, group=group)
    rows = []
    for r in range(world):
        a, b = shard_bounds(n, world, r)
        rows.append(gathered[r * max_rows: r * max_rows + (b - a)])
    return torch.cat(rows, dim=0)


class ShardedModel:
    """``Model`` whose ``forward`` splits the start-frame batch over the process group."""

    def __init__(self, model, group=None):
        self.model, self.group = model, group

    @torch.no_grad()
    def sample(self, x_0, cond=None, residual=None):
        if residual is None:
            residual = torch.randn(x_0.size(0), self.model.z_dim)      # global draw, CPU RNG (Q5)
        fn = lambda x, r, c: self.model.sample(x, c, residual=r)
        return sharded_sample(fn, x_0, residual, cond, self.group)

    def forward(self, x_0, cond=None):
        return self.sample(x_0, cond)[: self.model.vid_length]          # quirk Q1, get_model.py:75

    __call__ = forward
