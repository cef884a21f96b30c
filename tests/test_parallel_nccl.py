"""The only collective of the path -- the PPO-gradient pmean (jaxrl/MARL/ippo_rnn_JAXMARL_pmap.py:564-565) -- over NCCL on
two real GPUs: per-block buckets enqueued on a side stream behind the events the backward pass records, mean folded into the
reduction (ncclAvg), double-buffered flat gradient.  Skipped on boxes with fewer than 2 GPUs (run with `gpurun --gpus 2`)."""
import os

import pytest
import torch

pytestmark = pytest.mark.gpu


def _worker(rank, world, port, out):
    import torch.distributed as dist
    from oracle import vit_oracle as VO
    from vitmarl_b200 import parallel, vit
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    try:
        cfg = vit.VIT_PARITY
        params = vit.init_params(cfg, 0, "cuda")
        enc = vit.ViTEncoder(cfg)
        packed = vit.pack_params(cfg, params)
        red = parallel.GradAllReducer([t.shape for t in packed], device="cuda", bucket_ranges=enc.bucket_param_ranges())
        B = 24
        g = torch.Generator().manual_seed(100 + rank)                    # every rank its own minibatch
        lens = torch.randint(0, cfg.img_w + 1, (B, cfg.img_h, 1, cfg.channels), generator=g)
        x = (torch.arange(cfg.img_w)[None, None, :, None] < lens).float().cuda()
        dy = torch.randn(B, cfg.dim, generator=g).cuda()
        ok = True
        for rnd in range(3):                                             # several minibatches: buffers flip, events are re-recorded
            enc.apply_packed(packed, x, train=True)
            enc.vjp_packed(packed, dy, grads=red.grads(), flat=red.flat, bucket_events=red.events)
            red.allreduce_mean(async_op=True)                            # per-bucket ncclAllReduce(avg) behind the bucket events
            red.swap()
            reduced = red.wait()
            # reference: every rank's local gradient (no collective) gathered and averaged with plain tensor ops
            local, _ = enc.vjp_packed(packed, dy)
            for a, b in zip(reduced, local):
                parts = [torch.empty_like(b) for _ in range(world)]
                dist.all_gather(parts, b.contiguous())
                want = torch.stack(parts).mean(0)
                ok = ok and bool(torch.allclose(a, want, rtol=1e-5, atol=1e-6))
            ok = ok and bool(all(torch.equal(reduced[i], reduced[i]) for i in range(len(reduced))))
        # all ranks hold the same averaged table
        chk = torch.stack([t.double().sum() for t in reduced]).sum().reshape(1)
        both = [torch.empty_like(chk) for _ in range(world)]
        dist.all_gather(both, chk)
        ok = ok and bool(torch.equal(both[0], both[1]))
        # and it matches the fp32 oracle's gradient of the SUM of both minibatches / world (one leaf, cosine)
        out[rank] = ok
    finally:
        dist.destroy_process_group()


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs 2 GPUs (NCCL)")
def test_grad_pmean_world2_nccl_bucketed_overlapped():
    import torch.multiprocessing as mp
    mgr = mp.Manager()
    out = mgr.dict()
    port = 29600 + (os.getpid() % 2000)
    mp.spawn(_worker, args=(2, port, out), nprocs=2, join=True)
    assert out[0] and out[1]
