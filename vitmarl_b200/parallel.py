"""Multi-GPU plumbing (one process per GPU, torch.distributed over NCCL/NVLink).

The reference is data-parallel only: environments are sharded over devices
(``jaxrl/MARL/ippo_rnn_JAXMARL_pmap.py:290-330``), parameters replicated (``:279``) and the ONLY
collective is ``jax.lax.pmean`` of the loss and the gradient pytree once per minibatch
(``:564-565``).  Here: ``shard_envs`` gives the contiguous env range of a rank (no data-path
collective on the rollout-and-encode path), ``GradAllReducer`` is the pmean of the packed fp32
gradient table: the gradients are produced directly into ONE flat buffer (views per tensor) cut into per-block
buckets; each bucket's ncclAllReduce(avg) is enqueued on a side stream behind the CUDA event the backward pass
records when that block's gradients are final, so the reduction of block l overlaps the backward of blocks l-1..0."""
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
    """pmean of a gradient table (``jax.lax.pmean(grads, 'device_batch')``, pmap trainer :565), bucketed and overlapped.

    * The gradients are produced directly into ONE flat fp32 buffer (per-tensor views, 16-byte aligned segments): no
      gather / scatter copies, one memset per backward pass.
    * ``bucket_ranges`` (index ranges into the table, in the order the backward pass completes them --
      ``ViTEncoder.bucket_param_ranges()``) cut the flat buffer into per-block buckets.  ``ViTEncoder.vjp_packed(...,
      bucket_events=red.events)`` records an event per bucket on the compute stream; :meth:`allreduce_mean` enqueues one
      ``ncclAllReduce`` per bucket on a side stream behind its event, so block l's reduction runs over NVLink while
      blocks l-1 .. 0 are still in their backward pass.
    * The mean is folded into the collective (``ReduceOp.AVG`` on NCCL): no separate scaling pass.
    * DOUBLE BUFFERED: :meth:`swap` flips to the second flat buffer, so the next minibatch's backward may start writing
      while the previous buffer is still being reduced / consumed by the optimiser (without it the next backward's memset
      would race with the in-flight all-reduce).  ``wait()`` makes the compute stream wait for the reduction of the buffer
      it was started on and returns that buffer's views."""

    def __init__(self, shapes: Sequence[torch.Size], device="cuda", group=None, bucket_ranges=None, double_buffer: bool = True,
                 n_groups: int = 0):
        self.group = group
        dev = torch.device(device)
        self.cuda = dev.type == "cuda"
        nbuf = 2 if double_buffer else 1
        self._flat, self._views = [], []
        for _ in range(nbuf):
            f, v = flat_views(shapes, torch.float32, device)
            self._flat.append(f); self._views.append(v)
        self._cur = 0
        self.world = dist.get_world_size(group) if dist.is_initialized() else 1
        self.stream = torch.cuda.Stream(device=device) if self.cuda else None
        # element offsets of every tensor in the flat buffer
        offs, total = [], 0
        for s in shapes:
            offs.append(total)
            total += (int(torch.Size(s).numel()) + 3) // 4 * 4
        offs.append(total)
        self._offs = offs
        if bucket_ranges is None:
            bucket_ranges = [(0, len(shapes))]
        self.bucket_ranges = [(int(a), int(b)) for a, b in bucket_ranges]
        self.bucket_spans = [(offs[a], offs[b]) for a, b in self.bucket_ranges]      # element spans in the flat buffer
        # Collectives actually issued: consecutive buckets (in completion order) coalesced into `n_groups` all-reduces of roughly
        # equal size, each hung behind the event of its LAST bucket (0 = one all-reduce per bucket).  Measured on B200/NVLink
        # (bench.py extra, profiles/): the 21.5 MB ViT-Tiny table reduces in ~70 us, while every collective that overlaps the
        # backward costs the persistent compute kernels a straggler wave -- so few, large groups beat per-block buckets here.
        self.groups = self._coalesce(n_groups)
        self.events = None
        if self.cuda:
            self.events = [torch.cuda.Event() for _ in self.bucket_ranges]
            for e in self.events:
                e.record()                       # materialises the cudaEvent_t (torch creates it lazily)
        self._work = []
        self._pending = None
        # NCCL reduces with ncclAvg; gloo (CPU tests) has no AVG -> sum, then scale
        self._avg = self.cuda and dist.is_initialized() and dist.get_backend(group) == "nccl"

    def _coalesce(self, n_groups: int):
        spans = self.bucket_spans
        if n_groups <= 0 or n_groups >= len(spans):
            return [(lo, hi, b) for b, (lo, hi) in enumerate(spans)]
        total = sum(hi - lo for lo, hi in spans)
        groups, lo_g, hi_g, acc, made = [], None, None, 0, 0
        for b, (lo, hi) in enumerate(spans):
            lo_g = lo if lo_g is None else min(lo_g, lo)
            hi_g = hi if hi_g is None else max(hi_g, hi)
            acc += hi - lo
            last = b == len(spans) - 1
            if last or (made < n_groups - 1 and acc >= total * (made + 1) / n_groups):
                groups.append((lo_g, hi_g, b))
                lo_g = hi_g = None
                made += 1
        covered = sum(hi - lo for lo, hi, _ in groups)
        if covered != total:          # buckets of a group must be contiguous in the flat buffer (they are, in table order)
            raise ValueError("bucket_ranges are not contiguous in completion order; cannot coalesce")
        return groups

    # ---- buffers -------------------------------------------------------------------------------------------------
    @property
    def flat(self) -> torch.Tensor:
        return self._flat[self._cur]

    @property
    def views(self) -> List[torch.Tensor]:
        return self._views[self._cur]

    def grads(self) -> List[torch.Tensor]:
        """Views to hand to ``ViTEncoder.vjp_packed(..., grads=red.grads(), flat=red.flat, bucket_events=red.events)``."""
        return self._views[self._cur]

    def swap(self):
        """Flip to the other flat buffer (double buffering): the next backward writes there."""
        self._cur = (self._cur + 1) % len(self._flat)

    # ---- the collective ------------------------------------------------------------------------------------------
    def allreduce_mean(self, async_op: bool = False, use_events: bool = True):
        """Enqueue the per-bucket all-reduces of the CURRENT buffer.  With ``use_events`` (and a preceding
        ``vjp_packed(..., bucket_events=self.events)``) bucket b starts as soon as its event has fired; otherwise the side
        stream waits for everything enqueued so far on the compute stream."""
        if self.world == 1:
            return None
        flat = self._flat[self._cur]
        self._pending = self._cur
        op = dist.ReduceOp.AVG if self._avg else dist.ReduceOp.SUM
        if self.stream is not None:
            if not use_events:
                self.stream.wait_stream(torch.cuda.current_stream())
            with torch.cuda.stream(self.stream):
                for lo, hi, b in self.groups:
                    if hi <= lo:
                        continue
                    if use_events:
                        self.stream.wait_event(self.events[b])
                    self._work.append(dist.all_reduce(flat[lo:hi], op=op, group=self.group, async_op=True))
        else:
            for lo, hi, _ in self.groups:
                if hi > lo:
                    self._work.append(dist.all_reduce(flat[lo:hi], op=op, group=self.group, async_op=True))
        if not async_op:
            return self.wait()
        return self._work

    def wait(self) -> List[torch.Tensor]:
        """Order the compute stream after the in-flight reduction; returns the reduced buffer's views."""
        idx = self._cur if self._pending is None else self._pending
        if self._work:
            for w in self._work:
                w.wait()
            self._work = []
            if self.stream is not None:
                torch.cuda.current_stream().wait_stream(self.stream)
            if not self._avg:
                self._flat[idx].mul_(1.0 / self.world)
        self._pending = None
        return self._views[idx]
