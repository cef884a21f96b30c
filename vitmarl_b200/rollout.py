"""Batched rollout-and-encode step: the public call a user of this framework makes.

    engine = RolloutEncoder(cfg, vit_cfg, params, E, M)
    engine.reset(asks, bids)                       # device-resident book state
    feats = engine.step(msgs_device)               # [E, D] fp32 ViT encoding of the new LOB image
    feats_host = engine.step_host(msgs_pinned)     # same with host buffers (H2D / D2H inside)

One step = fused order-book kernel (scan + forward fill + mid price + vision tensor + raster,
``marl_env.py:377-393,466-467,637-662``) followed by the ViT encoder forward on the rendered
images.  Environments are independent, so multi-GPU runs shard ``E`` across ranks with no
collective on this path (``ippo_rnn_JAXMARL_pmap.py:290-330``)."""
from __future__ import annotations

import torch

from . import env as venv
from . import vit as vvit
from .config import World_EnvironmentConfig

__all__ = ["RolloutEncoder", "kernels_per_step"]


def kernels_per_step(vit_cfg: vvit.ViTConfig) -> int:
    """Launches of this library's kernels in one rollout step.
    ViT-Tiny (D=192, 3 heads): 1 env-step (which renders the patch matrix itself: no patchify) + patch-embed GEMM
    + L x (fused attention block + fused MLP block) + final LN/pool (the 2 parameter-fold launches run only on the engine's
    first step).  Other shapes run the unfused sequence: (1 + 4 L) GEMMs + 2 L LayerNorm + L attention."""
    L = vit_cfg.depth
    if vit_cfg.dim == 192 and vit_cfg.heads == 3 and vit_cfg.mlp_dim == 768 and vit_cfg.tokens == 64:
        return 1 + 1 + 2 * L + 1
    return 1 + (1 + 4 * L) + 2 * L + L + 1


class RolloutEncoder:
    def __init__(self, cfg: World_EnvironmentConfig, vit_cfg: vvit.ViTConfig, params, E: int, M: int,
                 n_levels: int = 10, device="cuda"):
        self.cfg, self.vit_cfg, self.E, self.M, self.n_levels = cfg, vit_cfg, E, M, n_levels
        self.device = torch.device(device)
        self.encoder = vvit.ViTEncoder(vit_cfg)
        self.packed = vvit.pack_params(vit_cfg, params)
        self.params_version = 0          # generation counter of `packed`'s contents (bumped by update_params)
        # step outputs allocated once, rewritten in place every step; TWO sets so that the host-buffer path can copy step i's
        # results out while step i+1 is already writing the other set
        self._bufs = [venv.StepBuffers(), venv.StepBuffers()]
        self._feat = [None, None]
        self.state = None
        self.last = None
        # staging for the host-buffer path (double buffered: messages in, encoding + vision tensor out)
        self._host = None
        self._host_i = 0

    def reset(self, asks: torch.Tensor, bids: torch.Tensor):
        self.state = venv.reset(self.cfg, asks, bids, self.M)
        return self.state

    def update_params(self, params=None, packed=None):
        """After an optimiser step: hand in the new pytree (re-packed here) or an already packed table whose contents
        changed in place; the next step folds the parameters again."""
        if params is not None:
            self.packed = vvit.pack_params(self.vit_cfg, params)
        elif packed is not None:
            self.packed = packed
        self.params_version += 1

    def step(self, msgs: torch.Tensor, stat_agent_ids=None, keep_trades: bool = True, slot: int = 0) -> torch.Tensor:
        """Two library calls, 1 + 1 + 2L + 1 kernel launches, no allocation and no torch kernel: the env-step outputs live in
        buffers that are rewritten in place (``self.last`` and the returned encoding are valid until the next step that uses
        the same ``slot``)."""
        c = self.vit_cfg
        # the order-book kernel renders the raster directly as the encoder's patch matrix (no patchify pass in between)
        self.state, out = venv.step(self.cfg, self.state, msgs, n_levels=self.n_levels, want_obs=True,
                                    image_hw=(c.img_h, c.img_w), image_dtype=torch.bfloat16, inplace=True, image_patch=c.patch,
                                    stat_agent_ids=stat_agent_ids, keep_trades=keep_trades, buffers=self._bufs[slot])
        self.last = out
        if self._feat[slot] is None:
            self._feat[slot] = torch.empty((self.E, c.dim), dtype=torch.float32, device=self.device)
        # folded parameters are reused until update_params() bumps the version
        return self.encoder.apply_packed(self.packed, out.image, params_version=self.params_version, patches=True, out=self._feat[slot])

    def last_image(self) -> torch.Tensor:
        """The last step's raster as the reference-shaped [E,H,W,2] image (a view-permutation of the patch matrix)."""
        c = self.vit_cfg
        p, gh, gw = c.patch, c.img_h // c.patch, c.img_w // c.patch
        return self.last.image.view(self.E, gh, gw, p, p, c.channels).permute(0, 1, 3, 2, 4, 5).reshape(self.E, c.img_h, c.img_w, c.channels)

    def _host_setup(self):
        E, c = self.E, self.vit_cfg
        ev = lambda: [torch.cuda.Event(), torch.cuda.Event()]
        self._host = {
            "msgs": [torch.empty((E, self.M, 8), dtype=torch.int32, device=self.device) for _ in range(2)],
            "feat": [torch.empty((E, c.dim), dtype=torch.float32).pin_memory() for _ in range(2)],
            "obs": [torch.empty((E, self.n_levels, 3, 2), dtype=torch.float32).pin_memory() for _ in range(2)],
            "h2d": torch.cuda.Stream(self.device), "d2h": torch.cuda.Stream(self.device),
            "h2d_done": ev(), "step_done": ev(), "d2h_done": ev(), "used": [False, False],
        }

    def step_host(self, msgs_pinned: torch.Tensor):
        """Host-buffer entry: H2D of the step's messages (pinned host memory), the step, D2H of the encoding and the vision
        tensor -> (feats_host, vision_obs_host), valid after ``wait_host()`` / a device synchronise and until the call after
        next.  The copies run on two side streams against double-buffered staging, so the input copy of step i+1 and the
        result copy of step i-1 overlap the kernels of step i (every step's copies are still issued by, and ordered against,
        that step; a caller whose next messages depend on this step's results simply synchronises in between)."""
        if self._host is None:
            self._host_setup()
        h = self._host
        k = self._host_i & 1
        self._host_i += 1
        cur = torch.cuda.current_stream(self.device)
        if h["used"][k]:
            h["h2d"].wait_event(h["step_done"][k])        # the step two calls ago has read this message buffer
        with torch.cuda.stream(h["h2d"]):
            h["msgs"][k].copy_(msgs_pinned, non_blocking=True)
            h["h2d_done"][k].record(h["h2d"])
        cur.wait_event(h["h2d_done"][k])
        if h["used"][k]:
            cur.wait_event(h["d2h_done"][k])              # ... and its results have left this slot's device buffers
        feats = self.step(h["msgs"][k], slot=k)
        obs = self.last.vision_obs
        h["step_done"][k].record(cur)
        with torch.cuda.stream(h["d2h"]):
            h["d2h"].wait_event(h["step_done"][k])
            h["feat"][k].copy_(feats, non_blocking=True)
            h["obs"][k].copy_(obs, non_blocking=True)
            h["d2h_done"][k].record(h["d2h"])
        h["used"][k] = True
        return h["feat"][k], h["obs"][k]

    def wait_host(self):
        """Make the current stream wait for every outstanding result copy of ``step_host`` (call before reading the host
        buffers without a full device synchronise, and before closing a timed region)."""
        if self._host is not None:
            cur = torch.cuda.current_stream(self.device)
            for k in range(2):
                if self._host["used"][k]:
                    cur.wait_event(self._host["d2h_done"][k])

    @property
    def h2d_bytes_per_step(self) -> int:
        return self.E * self.M * 8 * 4

    @property
    def d2h_bytes_per_step(self) -> int:
        return self.E * self.vit_cfg.dim * 4 + self.E * self.n_levels * 6 * 4
