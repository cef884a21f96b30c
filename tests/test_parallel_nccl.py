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
        bad = []
        for rnd in range(3):                                             # several minibatches: buffers flip, events are re-recorded
            enc.apply_packed(packed, x, train=True)
            enc.vjp_packed(packed, dy, grads=red.grads(), flat=red.flat, bucket_events=red.events)
            red.allreduce_mean(async_op=True)                            # per-bucket ncclAllReduce(avg) behind the bucket events
            red.swap()
            reduced = red.wait()
            # reference: every rank's local gradient (no collective) gathered and averaged with plain tensor ops.  The weight
            # gradients are fp32 split-K reductions (red.add: summation order varies run to run), hence a norm-wise tolerance.
            local, _ = enc.vjp_packed(packed, dy)
            for i, (a, b) in enumerate(zip(reduced, local)):
                parts = [torch.empty_like(b) for _ in range(world)]
                dist.all_gather(parts, b.contiguous())
                want = torch.stack(parts).mean(0)
                err = float((a - want).norm() / (want.norm() + 1e-30))
                if not err <= 1e-4:
                    bad.append((rnd, i, err, float(want.norm())))
        # all ranks hold the same averaged table, bit for bit (same reduction on every rank)
        chk = torch.stack([t.double().sum() for t in reduced]).sum().reshape(1)
        both = [torch.empty_like(chk) for _ in range(world)]
        dist.all_gather(both, chk)
        if not torch.equal(both[0], both[1]):
            bad.append(("ranks differ", float(both[0]), float(both[1])))
        out[rank] = bad
    finally:
        dist.destroy_process_group()


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs 2 GPUs (NCCL)")
def test_grad_pmean_world2_nccl_bucketed_overlapped():
    import torch.multiprocessing as mp
    mgr = mp.Manager()
    out = mgr.dict()
    port = 29600 + (os.getpid() % 2000)
    mp.spawn(_worker, args=(2, port, out), nprocs=2, join=True)
    assert out[0] == [] and out[1] == [], (out[0][:8], out[1][:8])
