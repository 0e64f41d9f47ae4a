"""Multi-GPU driver of the path: shard, transform locally, (optionally) gather.  No collective on the hot path.

Every (channel, epoch) signal and every analysis frequency is independent in the reference
(`base.py:399-406` has no cross-signal term, `mneutils.py:39` maps epochs independently), so one process
per GPU owns a contiguous block of signals - or, when there are fewer signals than ranks, a contiguous
block of frequencies (each rank then repeats the cheap forward transform).  The only communication is the
optional end-of-run `all_gather` of results over `torch.distributed` (NCCL on GPUs, gloo in the CPU tests).
"""
from typing import Callable, Optional, Tuple

import numpy as np


def shard_range(n_items: int, rank: int, world: int) -> Tuple[int, int]:
    """Contiguous block [lo, hi) of `n_items` for `rank`; block sizes differ by at most one."""
    if world <= 0 or not (0 <= rank < world):
        raise ValueError("bad rank/world")
    base, rem = divmod(int(n_items), world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def shard_axis(n_signals: int, n_freqs: int, world: int) -> str:
    """'signals' when every rank can get at least one signal, else 'freqs'."""
    return "signals" if n_signals >= world or n_freqs < world else "freqs"


def gather_blocks(local, axis: int, counts, group=None):
    """all_gather of unequal blocks along `axis` (torch tensors): pad to the largest block, gather, trim."""
    import torch
    import torch.distributed as dist
    world = dist.get_world_size(group)
    mx = max(counts)
    pad_shape = list(local.shape)
    pad_shape[axis] = mx
    buf = torch.zeros(pad_shape, dtype=local.dtype, device=local.device)
    buf.narrow(axis, 0, local.shape[axis]).copy_(local)
    outs = [torch.empty_like(buf) for _ in range(world)]
    dist.all_gather(outs, buf, group=group)
    return torch.cat([o.narrow(axis, 0, c) for o, c in zip(outs, counts)], dim=axis)


def distributed_transform(local_fn: Callable, signals, freqs, *, gather: bool = True, group=None,
                          rank: Optional[int] = None, world: Optional[int] = None):
    """Run `local_fn(signals_block, freqs_block) -> [s, f, n]` on this rank's shard.

    `local_fn` is the single-GPU call, e.g. `lambda x, f: wavelet.power(torch.as_tensor(x).cuda(), f, reuse=False)`.
    Returns the full `[S, F, N]` result on every rank when `gather` (torch tensor), else
    `(block, axis_name, (lo, hi))`.
    """
    import torch
    import torch.distributed as dist
    if world is None:
        world = dist.get_world_size(group) if dist.is_initialized() else 1
    if rank is None:
        rank = dist.get_rank(group) if dist.is_initialized() else 0
    if not hasattr(signals, "shape"):   # numpy arrays, torch tensors (host or device resident) and array-likes with shape + slicing pass through
        signals = np.asarray(signals)
    freqs = np.asarray(freqs, dtype=np.float64)
    S, F = signals.shape[0], len(freqs)
    axis_name = shard_axis(S, F, world)
    if axis_name == "signals":
        lo, hi = shard_range(S, rank, world)
        block = local_fn(signals[lo:hi], freqs) if hi > lo else None
        counts = [shard_range(S, r, world)[1] - shard_range(S, r, world)[0] for r in range(world)]
        axis = 0
    else:
        lo, hi = shard_range(F, rank, world)
        # the reference needs >= 2 frequencies to build a bank (base.py:272); shards of one borrow a neighbour
        flo = lo if hi - lo >= 2 or F < 2 else max(0, min(lo, F - 2))
        fhi = max(hi, flo + 2) if F >= 2 else hi
        full = local_fn(signals, freqs[flo:fhi]) if hi > lo else None
        block = None if full is None else full[:, lo - flo: hi - flo]
        counts = [shard_range(F, r, world)[1] - shard_range(F, r, world)[0] for r in range(world)]
        axis = 1
    if block is not None and not torch.is_tensor(block):
        block = torch.as_tensor(np.asarray(block))
    if not gather or world == 1:
        return block if (gather and world == 1) else (block, axis_name, (lo, hi))
    if block is None:   # more ranks than items: contribute an empty block of the right shape
        ref_shape = [S, F, signals.shape[1]]
        ref_shape[axis] = 0
        block = torch.zeros(ref_shape, dtype=torch.float64)
    return gather_blocks(block, axis, counts, group)
