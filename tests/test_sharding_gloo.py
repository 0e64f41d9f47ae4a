"""World-size-2 (and 3) gloo test of the multi-GPU driver's host logic: the shards tile the work exactly, the
local transform is called on the right blocks, and the gathered result equals the single-process one.
The local transform here is the CPU oracle (test infrastructure) standing in for the per-GPU CUDA call."""
import os
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "oracle")]

from ninwavelets_b200.sharding import shard_range, shard_axis, distributed_transform  # noqa: E402


def test_shard_ranges_tile_exactly():
    for n in (0, 1, 5, 64, 306 * 200, 7):
        for world in (1, 2, 3, 4, 8):
            blocks = [shard_range(n, r, world) for r in range(world)]
            assert blocks[0][0] == 0 and blocks[-1][1] == n
            assert all(a[1] == b[0] for a, b in zip(blocks, blocks[1:]))
            sizes = [b - a for a, b in blocks]
            assert max(sizes) - min(sizes) <= 1
    assert shard_axis(64, 100, 8) == "signals" and shard_axis(1, 100, 8) == "freqs" and shard_axis(1, 2, 8) == "signals"


def _local(x, f):
    import cwt_oracle as orc
    fam = orc.Family("morse", sfreq=1000.0)
    return np.stack([orc.power(fam, xi, f) for xi in x]) if len(x) else np.zeros((0, len(f), x.shape[1]))


def _worker(rank, world, port, S, F, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        rng = np.random.default_rng(7)
        x = rng.standard_normal((S, 300))
        freqs = np.arange(1, F + 1.0)
        out = distributed_transform(_local, x, freqs, gather=True)
        ref = _local(x, freqs)
        err = float(np.abs(out.numpy() - ref).max() / np.abs(ref).max())
        q.put((rank, tuple(out.shape), err))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world,S,F", [(2, 5, 6), (2, 1, 7), (3, 4, 5)])
def test_distributed_transform_gloo(world, S, F):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + (os.getpid() % 2000) + world * 3 + S
    procs = [ctx.Process(target=_worker, args=(r, world, port, S, F, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in procs]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    for rank, shape, err in res:
        assert shape == (S, F, 300), (rank, shape)
        assert err < 1e-12, (rank, err)


def test_distributed_transform_takes_resident_tensors_and_proxies():
    """The driver must not pull its input through numpy: a torch tensor (device resident on a GPU box) or any object with
    `.shape` and slicing (bench.py hands it a view of the rank's resident block) is sliced as it is."""
    rng = np.random.default_rng(3)
    x = rng.standard_normal((5, 300))
    fr = np.arange(2.0, 7.0)
    seen = []

    def local(xs, f):
        seen.append(type(xs))
        return _local(np.asarray(xs), f)

    for rank in range(2):
        blk, axis, (lo, hi) = distributed_transform(local, torch.as_tensor(x), fr, gather=False, rank=rank, world=2)
        assert axis == "signals" and np.allclose(blk.numpy(), _local(x[lo:hi], fr))
    assert all(t is torch.Tensor for t in seen)

    class Proxy:
        shape = x.shape

        def __getitem__(self, sl):
            return x[sl]
    blk, axis, (lo, hi) = distributed_transform(_local, Proxy(), fr, gather=False, rank=1, world=2)
    assert (lo, hi) == shard_range(5, 1, 2) and np.allclose(blk.numpy(), _local(x[lo:hi], fr))
