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


def test_bucket_coalescing():
    """14 per-block buckets of a ViT-Tiny table coalesced into 1 / 2 / 14 collectives: spans stay contiguous, cover the table
    once, and every group waits on the event of its last bucket."""
    from vitmarl_b200 import vit
    cfg = vit.VIT_TINY_8
    shapes = [torch.Size(s) for s in _tiny_shapes(cfg)]
    ranges = vit.ViTEncoder(cfg).bucket_param_ranges()
    for n, want in ((0, 14), (14, 14), (1, 1), (2, 2), (3, 3)):
        red = parallel.GradAllReducer(shapes, device="cpu", bucket_ranges=ranges, n_groups=n, double_buffer=False)
        assert len(red.groups) == want
        assert sum(hi - lo for lo, hi, _ in red.groups) == red.flat.numel()
        assert red.groups[-1][2] == 13 and [g[2] for g in red.groups] == sorted(g[2] for g in red.groups)
    two = parallel.GradAllReducer(shapes, device="cpu", bucket_ranges=ranges, n_groups=2, double_buffer=False).groups
    assert two[0][1] == red.flat.numel() and two[1][0] == 0 and two[0][0] == two[1][1]      # [tail of the table], then [head]


def _tiny_shapes(cfg):
    D, K, T, H = cfg.dim, cfg.patch_dim, cfg.tokens, cfg.mlp_dim
    out = [(D, K), (D,), (T, D)]
    for _ in range(cfg.depth):
        out += [(D,), (D,), (3 * D, D), (3 * D,), (D, D), (D,), (D,), (D,), (H, D), (H,), (D, H), (D,)]
    return out + [(D,), (D,)]


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
