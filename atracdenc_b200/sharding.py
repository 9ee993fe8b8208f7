"""
Multi-GPU plumbing for the encode path.  The path shards by STREAM (independent encoder instances,
SURVEY.md §8e): rank r owns a contiguous range of streams and there is no data-path collective.
torch.distributed is used only at the edges: an optional scatter of PCM shards / gather of
bitstream shards, barriers, and the max-over-ranks of device times for reporting.
Backend "nccl" on GPUs, "gloo" in the CPU tests.
"""
from __future__ import annotations

import torch
import torch.distributed as dist


def shard_range(n_streams: int, rank: int, world: int) -> tuple[int, int]:
    """Contiguous [lo, hi) range of streams owned by `rank`; sizes differ by at most one."""
    base, extra = divmod(n_streams, world)
    lo = rank * base + min(rank, extra)
    return lo, lo + base + (1 if rank < extra else 0)


def max_over_ranks(value: float, device: torch.device | str = "cpu") -> float:
    """Timing reduction required by the bench contract: a step is as slow as its slowest rank."""
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return float(value)
    t = torch.tensor([value], dtype=torch.float64, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t[0])


def gather_units(local_units: torch.Tensor, n_streams: int, dst: int = 0):
    """Gathers per-rank bitstream shards [S_r][F][U][B] (uint8) onto `dst` in stream order.
    Returns the full [S][F][U][B] tensor on dst, None elsewhere."""
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return local_units
    world, rank = dist.get_world_size(), dist.get_rank()
    shapes = [shard_range(n_streams, r, world) for r in range(world)]
    tail = tuple(local_units.shape[1:])
    if rank == dst:
        bufs = [torch.empty((hi - lo,) + tail, dtype=local_units.dtype, device=local_units.device) for lo, hi in shapes]
    else:
        bufs = None
    # shards may differ in size by one stream: point-to-point instead of a fixed-size gather; the receives are posted as
    # ONE batch (a grouped NCCL call) so that the seven transfers run concurrently instead of one after the other
    if rank == dst:
        bufs[dst].copy_(local_units)
        ops = [dist.P2POp(dist.irecv, bufs[r], r) for r in range(world) if r != dst]
        for q in dist.batch_isend_irecv(ops):
            q.wait()
        return torch.cat(bufs, dim=0)
    for q in dist.batch_isend_irecv([dist.P2POp(dist.isend, local_units.contiguous(), dst)]):
        q.wait()
    return None


def scatter_pcm(full_pcm: torch.Tensor | None, n_streams: int, per_stream_shape, dtype, device, src: int = 0):
    """Scatters PCM shards [S_r][...] from `src` (which holds [S][...]); never broadcasts the whole batch."""
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return full_pcm
    world, rank = dist.get_world_size(), dist.get_rank()
    lo, hi = shard_range(n_streams, rank, world)
    mine = torch.empty((hi - lo,) + tuple(per_stream_shape), dtype=dtype, device=device)
    if rank == src:
        ops = []
        for r in range(world):
            a, b = shard_range(n_streams, r, world)
            if r == src:
                mine.copy_(full_pcm[a:b])
            else:
                ops.append(dist.P2POp(dist.isend, full_pcm[a:b].contiguous(), r))
        for q in dist.batch_isend_irecv(ops):
            q.wait()
    else:
        for q in dist.batch_isend_irecv([dist.P2POp(dist.irecv, mine, src)]):
            q.wait()
    return mine
