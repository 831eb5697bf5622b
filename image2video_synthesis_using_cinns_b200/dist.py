"""Multi-GPU: batch sharding + one all-gather of the finished frames.

The sampling path never mixes samples (GroupNorm / InstanceNorm are per sample, the embedder's BatchNorm runs
in eval mode, ActNorm is a fixed affine), so N GPUs = N independent shards of the start-frame batch with the
weights replicated.  The only exchange is the gather of finished frames; there is no data-path collective
inside the model and therefore nothing to fuse a kernel with.

To keep N-GPU outputs identical to the 1-GPU run, the residual is drawn ONCE for the global batch on rank 0's CPU
generator (quirk Q5, get_model.py:59), broadcast, and then sliced -- ranks seeded differently (the usual
seed + rank) therefore still render one consistent global batch.

``FrameGather`` issues the collective on a side stream so that it overlaps the next batch's compute, and the
``uint8`` mode gathers the denormalised GIF pixels (utils/auxiliaries.py:15-22) instead of fp32 frames: 4x fewer
bytes over NVLink, with the clip maximum all-reduced first so every rank quantises with the same scale.
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


def global_residual(n, z_dim, group=None, device=None):
    """(n, z_dim) residual for the GLOBAL batch: drawn on rank 0's CPU generator (Q5) and broadcast, so the ranks agree
    whatever their own RNG state is.  `device`: where the broadcast buffer lives (NCCL needs a CUDA tensor)."""
    world = dist.get_world_size(group) if dist.is_initialized() else 1
    if world == 1:
        return torch.randn(n, z_dim)
    rank = dist.get_rank(group)
    res = torch.randn(n, z_dim) if rank == 0 else torch.empty(n, z_dim)
    src = dist.get_global_rank(group, 0) if group is not None else 0
    if dist.get_backend(group) == "nccl":
        buf = res.to(device if device is not None else torch.device("cuda", torch.cuda.current_device()))
        dist.broadcast(buf, src=src, group=group)
        return buf
    dist.broadcast(res, src=src, group=group)
    return res


class FrameGather:
    """All-gather of finished frames on a side stream: ``submit(frames)`` returns at once, the transfer runs behind the
    caller's next kernels, ``wait()`` makes the caller's stream own the gathered tensor.  Two output buffers alternate, so
    one gather may be in flight while the previous result is still being read.

    ``mode="nccl"`` (default): one ``all_gather_into_tensor`` on the side stream (its kernel finds its SMs between the
    persistent conv kernels of the next step: measured +0.7 / +2 ms per step at N = 2 / 4).
    ``mode="p2p"`` (EXPERIMENTAL): the output buffers live in symmetric memory (``torch.distributed._symmetric_memory``)
    and every rank WRITES its shard straight into its peers' buffers with peer-to-peer copies -- NVLink through the copy
    engines, no SM taken from the conv kernels -- bracketed by two device-side barriers on the side stream.  It passes
    ``tools/check_gather.py`` (also ``--stress``) and a bench run at N = 2 (52.8 ms per step against 53.3 with NCCL,
    ``profiles/r02_bench_n2_p2p_gather_experimental.json``), but ONE earlier bench run hung with it for reasons not yet
    understood (it had followed another torchrun job in the same shell): not used by default."""

    def __init__(self, device, group=None, mode="nccl"):
        self.device, self.group = torch.device(device), group
        self.stream = torch.cuda.Stream(device=self.device)
        self._bufs, self._k, self._pending = [None, None], 0, None
        self.mode = mode
        self._sym = [None, None]          # (handle, [peer views]) per buffer
        self._n = 0

    def _p2p_buffers(self, k, shape, dtype):
        import torch.distributed._symmetric_memory as symm
        world = dist.get_world_size(self.group)
        buf = symm.empty(*shape, dtype=dtype, device=self.device)
        hdl = symm.rendezvous(buf, self.group if self.group is not None else dist.group.WORLD)
        peers = [hdl.get_buffer(r, shape, dtype) for r in range(world)]
        self._bufs[k], self._sym[k] = buf, (hdl, peers)

    def submit(self, frames):
        world, rank = dist.get_world_size(self.group), dist.get_rank(self.group)
        frames = frames.contiguous()
        k = self._k
        rows = frames.shape[0]
        shape = (world * rows,) + tuple(frames.shape[1:])
        if self._bufs[k] is None or tuple(self._bufs[k].shape) != shape or self._bufs[k].dtype != frames.dtype:
            if self.mode == "p2p":
                try:
                    self._p2p_buffers(k, shape, frames.dtype)
                except Exception as e:          # no symmetric memory on this system / group: NCCL path
                    self.mode, self.fallback_reason = "nccl", repr(e)
            if self.mode != "p2p":
                self._bufs[k], self._sym[k] = torch.empty(shape, dtype=frames.dtype, device=self.device), None
        self.stream.wait_stream(torch.cuda.current_stream(self.device))     # frames are complete on the caller's stream
        with torch.cuda.stream(self.stream):
            if self.mode == "p2p" and self._sym[k] is not None:
                hdl, peers = self._sym[k]
                hdl.barrier(channel=0)            # every rank is past the consumers of this buffer's previous content
                for i in range(world):
                    r = (rank + i) % world        # start with the local copy, then walk the peers in a rotated order
                    peers[r][rank * rows:(rank + 1) * rows].copy_(frames, non_blocking=True)
                hdl.barrier(channel=1)            # all shards have landed in this rank's buffer
            else:
                dist.all_gather_into_tensor(self._bufs[k], frames, group=self.group)
        frames.record_stream(self.stream)                                   # the allocator must not recycle it early
        self._pending = self._bufs[k]
        self._k ^= 1
        self._n += 1
        return self._pending

    def wait(self):
        """Block the caller's STREAM (not the host) until the last submitted gather is complete; returns its result."""
        torch.cuda.current_stream(self.device).wait_stream(self.stream)
        return self._pending


class HostFrameSink:
    """Device -> pinned-host read-back of finished frames on a copy stream: ``put(frames)`` returns at once, the copy runs
    behind the caller's next kernels (two pinned buffers alternate), ``wait()`` blocks the HOST until the last copy has
    landed and returns that buffer.  This is how a serving loop keeps the PCIe read-back of batch i under the compute of
    batch i + 1; ``bench.py``'s end-to-end leg uses it."""

    def __init__(self, device):
        self.device = torch.device(device)
        self.stream = torch.cuda.Stream(device=self.device)
        self._bufs, self._k, self._last = [None, None], 0, None
        self._done = [None, None]

    def put(self, frames):
        k = self._k
        if self._bufs[k] is None or self._bufs[k].shape != frames.shape or self._bufs[k].dtype != frames.dtype:
            self._bufs[k] = torch.empty(frames.shape, dtype=frames.dtype).pin_memory()
        if self._done[k] is not None:
            self._done[k].synchronize()                                    # the buffer's previous copy (two puts ago)
        self.stream.wait_stream(torch.cuda.current_stream(self.device))    # frames are complete on the caller's stream
        with torch.cuda.stream(self.stream):
            self._bufs[k].copy_(frames, non_blocking=True)
            ev = torch.cuda.Event()
            ev.record(self.stream)
        frames.record_stream(self.stream)
        self._done[k], self._last = ev, k
        self._k ^= 1
        return self._bufs[k]

    def wait(self):
        if self._last is None:
            return None
        self._done[self._last].synchronize()
        return self._bufs[self._last]


class ShardedModel:
    """``Model`` whose ``forward`` splits the start-frame batch over the process group."""

    def __init__(self, model, group=None):
        self.model, self.group = model, group

    @torch.no_grad()
    def sample(self, x_0, cond=None, residual=None):
        if residual is None:
            residual = global_residual(x_0.size(0), self.model.z_dim, self.group, getattr(self.model, "device", None))
        fn = lambda x, r, c: self.model.sample(x, c, residual=r)
        return sharded_sample(fn, x_0, residual, cond, self.group)

    @torch.no_grad()
    def sample_u8(self, x_0, cond=None, residual=None):
        """Same shards, but every rank returns the uint8 videos (B, T, H, W, 3) = trunc(255 * denorm / max) with the
        maximum taken over the GLOBAL batch (utils/auxiliaries.py:15-22): one scalar MAX all-reduce + a gather of
        1 byte per colour sample instead of 4."""
        from . import cli
        if residual is None:
            residual = global_residual(x_0.size(0), self.model.z_dim, self.group, getattr(self.model, "device", None))

        def fn(x, r, c):
            seq = self.model.sample(x, c, residual=r)
            mx = cli.frames_max(seq) if seq.shape[0] > 0 else torch.zeros(1, device=seq.device)
            if dist.is_initialized() and dist.get_world_size(self.group) > 1:
                dist.all_reduce(mx, op=dist.ReduceOp.MAX, group=self.group)
            if seq.shape[0] == 0:
                return torch.empty((0, seq.shape[1]) + tuple(seq.shape[3:]) + (3,), dtype=torch.uint8, device=seq.device)
            return cli.frames_to_u8(seq, mx, "video")
        return sharded_sample(fn, x_0, residual, cond, self.group)

    def forward(self, x_0, cond=None):
        return self.sample(x_0, cond)[: self.model.vid_length]          # quirk Q1, get_model.py:75

    __call__ = forward
