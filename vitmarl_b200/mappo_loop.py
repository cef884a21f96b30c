"""BASELINE configs[4] shape: one full rollout + update iteration with the environments sharded over the GPUs of a box.

    rollout : S steps of  order-book step -> render -> ViT encode  on E envs per GPU (``RolloutEncoder.step``); the rendered
              observation of every step is kept for the update AS THE PATCH MATRIX the kernel wrote (S x E x 16 KB, no
              re-layout), together with the per-agent trade reductions the reward functions need (computed inside the step
              kernel -- the [T,8] trade log never goes to HBM);
    update  : ``epochs`` passes over the S*E stored observations in ``minibatches`` minibatches
              (ippo_rnn_JAXMARL.py:470-475: value_and_grad per minibatch).  A minibatch = ViT forward+backward over its images
              in micro-batches whose gradients ACCUMULATE in the flat fp32 table inside the library (no eager add/copy
              kernels), then the pmean of that table (ippo_rnn_JAXMARL_pmap.py:564-565) as NCCL all-reduce(avg) on a side
              stream, hung behind the CUDA events the last micro-batch's backward records (``pmean_groups``: 2 = the table in
              two halves, the first reduced while the lower blocks are still in their backward pass -- measured best on 8 GPUs
              for this 21.5 MB table, see bench.py extra / DESIGN.md; 0 = one per transformer block; 1 = one collective).

The policy head, PPO loss and optimiser are boundary-only rows of SURVEY.md 8a (A13-A14): dL/d(encoding) is synthetic and the
parameters are not changed, but every encoder FLOP, every byte and every collective of the iteration is there."""
from __future__ import annotations

import torch

from . import jaxob, parallel, rollout, synth, vit
from .config import World_EnvironmentConfig

__all__ = ["MappoLoop"]


class MappoLoop:
    def __init__(self, envs_per_gpu: int = 8192, rollout_steps: int = 128, epochs: int = 4, minibatches: int = 16, micro: int = 8192,
                 msgs_per_step: int = 13, vit_cfg: vit.ViTConfig = vit.VIT_TINY_8, rank: int = 0, world: int = 1,
                 agent_ids=(-100, -101), pmean_groups: int = 2):
        self.E, self.S, self.M = envs_per_gpu, rollout_steps, msgs_per_step
        self.epochs, self.minibatches = epochs, minibatches
        self.world, self.rank = world, rank
        self.cfg, self.vcfg = World_EnvironmentConfig(), vit_cfg
        self.agent_ids = list(agent_ids)
        E, S, M, c = self.E, self.S, self.M, vit_cfg
        seed = 1234 + 1000 * rank
        l2 = synth.make_l2_books(E, seed)
        init = torch.from_numpy(synth.init_msgs_from_l2_batched(l2)).cuda()
        self.asks0, self.bids0, _ = jaxob.scan_through_entire_array(self.cfg, None, init, (jaxob.init_orderside(100, E), jaxob.init_orderside(100, E), None))
        stream = synth.MessageStream(E, seed)
        self.blocks = [torch.from_numpy(stream.next(M)).cuda() for _ in range(8)]       # cycled (the books keep evolving)
        params = vit.init_params(c, 0, "cuda")
        self.eng = rollout.RolloutEncoder(self.cfg, c, params, E, M)
        self.enc = vit.ViTEncoder(c)                                                    # training path (own workspace)
        self.packed = vit.pack_params(c, params)
        self.red = parallel.GradAllReducer([t.shape for t in self.packed], device="cuda", bucket_ranges=self.enc.bucket_param_ranges(),
                                           n_groups=pmean_groups)
        self.traj = torch.empty((S, E, c.tokens, c.patch_dim), dtype=torch.bfloat16, device="cuda")    # patch matrices, as rendered
        self.feats = torch.empty((S, E, c.dim), dtype=torch.float32, device="cuda")
        self.stats = torch.empty((S, E, len(self.agent_ids), 8), dtype=torch.int32, device="cuda")
        self.total = S * E
        self.mb_size = self.total // minibatches
        self.micro = min(micro, self.mb_size)
        self.dy = torch.randn(self.micro, c.dim, device="cuda")
        self.flat_traj = self.traj.view(self.total, c.tokens, c.patch_dim)
        self.eng.reset(self.asks0.clone(), self.bids0.clone())

    # ---------------------------------------------------------------------------------------------------------------
    def rollout(self):
        for t in range(self.S):
            f = self.eng.step(self.blocks[t % len(self.blocks)], stat_agent_ids=self.agent_ids, keep_trades=False)
            self.feats[t].copy_(f, non_blocking=True)
            self.traj[t].copy_(self.eng.last.image, non_blocking=True)
            self.stats[t].copy_(self.eng.last.trade_stats, non_blocking=True)

    def minibatch(self, mb: int):
        lo0, hi0 = mb * self.mb_size, (mb + 1) * self.mb_size
        red = self.red
        first = True
        for lo in range(lo0, hi0, self.micro):
            x = self.flat_traj[lo:lo + self.micro]
            last = lo + self.micro >= hi0
            self.enc.apply_packed(self.packed, x, train=True, patches=True)
            self.enc.vjp_packed(self.packed, self.dy[: x.shape[0]], grads=red.grads(), flat=red.flat, accumulate=not first,
                                bucket_events=red.events if last else None)
            first = False
        red.allreduce_mean(async_op=True)       # per-block buckets, each behind its event of the last micro-batch
        red.swap()                              # the next minibatch's backward writes the other buffer

    def update(self):
        for _ in range(self.epochs):
            for mb in range(self.minibatches):
                self.minibatch(mb)
                # the optimiser step consumes the reduced table here, and the next minibatch's forward needs the new
                # parameters: the compute stream waits for the tail of the reduction (a stream dependency, not a host sync).
                # What overlaps is each block's all-reduce with the backward of the blocks below it.
                self.red.wait()
        self.eng.update_params(packed=self.packed)          # the rollout engine re-folds after an update

    def warmup(self):
        for t in range(3):
            self.eng.step(self.blocks[t], stat_agent_ids=self.agent_ids, keep_trades=False)
        self.traj[0].copy_(self.eng.last.image)
        x = self.flat_traj[: self.micro]
        self.enc.apply_packed(self.packed, x, train=True, patches=True)
        self.enc.vjp_packed(self.packed, self.dy, grads=self.red.grads(), flat=self.red.flat, bucket_events=self.red.events)
        self.red.allreduce_mean(async_op=True)
        self.red.swap()
        self.red.wait()

    def run(self, dist=None):
        """-> dict with rollout / update / iteration seconds (max over ranks) and the derived rates."""
        def sync():
            if dist is not None:
                dist.barrier()
            torch.cuda.synchronize()
        self.warmup()
        sync()
        ev = [torch.cuda.Event(enable_timing=True) for _ in range(3)]
        ev[0].record(); self.rollout(); ev[1].record(); self.update(); ev[2].record()
        sync()
        t_roll, t_upd = ev[0].elapsed_time(ev[1]) * 1e-3, ev[1].elapsed_time(ev[2]) * 1e-3
        t_all = t_roll + t_upd
        if dist is not None:
            tt = torch.tensor([t_roll, t_upd, t_all], device="cuda", dtype=torch.float64)
            dist.all_reduce(tt, op=dist.ReduceOp.MAX)
            t_roll, t_upd, t_all = [float(v) for v in tt]
        W, E, S = self.world, self.E, self.S
        return {"metric": "mappo_iteration_env_steps_per_sec", "value": W * E * S / t_all, "unit": "env-steps/s", "n_gpus": W,
                "envs_total": W * E, "rollout_steps": S, "epochs": self.epochs, "minibatches": self.minibatches, "micro_batch": self.micro,
                "rollout_s": t_roll, "update_s": t_upd, "rollout_env_steps_per_sec": W * E * S / t_roll,
                "update_images_per_sec": W * self.epochs * self.total / t_upd,
                "collectives": self.epochs * self.minibatches * len(self.red.groups) if W > 1 else 0,
                "grad_bytes_per_minibatch": self.red.flat.numel() * 4,
                "note": "encoder + env + NCCL only: policy head / PPO loss / optimiser are boundary-only rows (synthetic dL/d(encoding))"}
