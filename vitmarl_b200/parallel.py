"""Multi-GPU plumbing (one process per GPU, torch.distributed over NCCL/NVLink).

The reference is data-parallel only: environments are sharded over devices
(``jaxrl/MARL/ippo_rnn_JAXMARL_pmap.py:290-330``), parameters replicated (``:279``) and the ONLY
collective is ``jax.lax.pmean`` of the loss and the gradient pytree once per minibatch
(``:564-565``).  Here: ``shard_envs`` gives the contiguous env range of a rank (no data-path
collective on the rollout-and-encode path), ``GradAllReducer`` is the pmean of the packed fp32
gradient table: the gradients are produced directly into ONE flat buffer (views per tensor), so the
psum is a single ncclAllReduce launch per minibatch on a side stream, overlappable with the next
backward's first kernels."""
from __future__ import annotations

from typing import List, Sequence, Tuple

import torch
import torch.distributed as dist

__all__ = ["shard_envs", "GradAllReducer", "flat_views"]


def shard_envs(num_envs: int, world_size: int, rank: int) -> Tuple[int, int]:
    """Contiguous range [lo, hi) of environments owned by `rank` (GPU g owns envs [g*E/G, (g+1)*E/G),
    remainder spread over the first ranks)."""
    if world_size <= 0 or not (0 <= rank < world_size):
        raise ValueError("bad rank / world size")
    base, rem = divmod(num_envs, world_size)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def flat_views(shapes: Sequence[torch.Size], dtype=torch.float32, device="cuda"):
    """One flat buffer + per-tensor views (16-byte aligned segments)."""
    offs, total = [], 0
    for s in shapes:
        offs.append(total)
        n = int(torch.Size(s).numel())
        total += (n + 3) // 4 * 4
    flat = torch.zeros(total, dtype=dtype, device=device)
    views = [flat[o:o + int(torch.Size(s).numel())].view(s) for o, s in zip(offs, shapes)]
    return flat, views


class GradAllReducer:
    """pmean of a gradient table (``jax.lax.pmean(grads, 'device_batch')``, pmap trainer :565)."""

    def __init__(self, shapes: Sequence[torch.Size], device="cuda", group=None):
        self.group = group
        self.flat, self.views = flat_views(shapes, torch.float32, device)
        self.world = dist.get_world_size(group) if dist.is_initialized() else 1
        self.stream = torch.cuda.Stream(device=device) if torch.device(device).type == "cuda" else None
        self._work = None

    def grads(self) -> List[torch.Tensor]:
        """Views to hand to ``ViTEncoder.vjp_packed(..., grads=...)`` so the backward writes in place."""
        return self.views

    def allreduce_mean(self, async_op: bool = False):
        if self.world == 1:
            return None
        if self.stream is not None:
            self.stream.wait_stream(torch.cuda.current_stream())
            with torch.cuda.stream(self.stream):
                self._work = dist.all_reduce(self.flat, op=dist.ReduceOp.SUM, group=self.group, async_op=True)
        else:
            self._work = dist.all_reduce(self.flat, op=dist.ReduceOp.SUM, group=self.group, async_op=True)
        if not async_op:
            self.wait()
        return self._work

    def wait(self):
        if self._work is not None:
            self._work.wait()
            self._work = None
            if self.stream is not None:
                torch.cuda.current_stream().wait_stream(self.stream)
            self.flat.mul_(1.0 / self.world)
