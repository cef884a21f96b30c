"""N>1 host logic on CPU: env sharding and the gradient pmean over a world_size-2 gloo group."""
import os

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from vitmarl_b200 import parallel


def test_shard_envs_partitions_exactly():
    for E, G in ((65536, 8), (4096, 1), (10, 3), (7, 8)):
        spans = [parallel.shard_envs(E, G, r) for r in range(G)]
        assert spans[0][0] == 0 and spans[-1][1] == E
        assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
        sizes = [hi - lo for lo, hi in spans]
        assert max(sizes) - min(sizes) <= 1
    assert parallel.shard_envs(65536, 8, 3) == (24576, 32768)
    with pytest.raises(ValueError):
        parallel.shard_envs(8, 2, 2)


def _worker(rank, world, port, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    shapes = [torch.Size((3, 5)), torch.Size((7,)), torch.Size((2, 2, 2))]
    # two buckets in backward-completion order (last tensor first), double-buffered flat gradient
    red = parallel.GradAllReducer(shapes, device="cpu", bucket_ranges=[(2, 3), (0, 2)])
    ok = True
    for rnd in range(3):
        for i, g in enumerate(red.grads()):
            g.copy_(torch.full(g.shape, float((rank + 1) * (i + 1) + rnd)))
        red.allreduce_mean(async_op=True)
        reduced = red.wait()
        ok = ok and all(torch.allclose(g, torch.full(g.shape, (1 + 2) / 2 * (i + 1) + rnd)) for i, g in enumerate(reduced))
        prev = red.flat
        red.swap()                                   # the next "backward" writes the other buffer
        ok = ok and red.flat.data_ptr() != prev.data_ptr()
    # sharded rollout bookkeeping: every rank owns a disjoint env range, union is everything
    lo, hi = parallel.shard_envs(13, world, rank)
    owned = torch.zeros(13)
    owned[lo:hi] = 1
    dist.all_reduce(owned)
    ok = ok and bool((owned == 1).all())
    out[rank] = ok
    dist.destroy_process_group()


def test_grad_pmean_world2_gloo():
    mgr = mp.Manager()
    out = mgr.dict()
    port = 29500 + (os.getpid() % 2000)
    mp.spawn(_worker, args=(2, port, out), nprocs=2, join=True)
    assert out[0] and out[1]
