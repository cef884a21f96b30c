"""``ActorCriticRNN`` with the vision encoder wired in -- the ``# FIXME: APPLY VISION`` slot of the reference trainers
(``gymnax_exchange/jaxrl/MARL/ippo_rnn_JAXMARL.py:75-115``; call sites ``:317, :408, :425``).

Same call shape as the reference module::

    hidden, logits, value = net.apply(params, hidden, (obs, dones))      # train_state.apply_fn(params, h, ac_in)

* ``obs``: ``[S, B, F]`` time-major vector observation (what the reference passes), or a tuple ``(vec_obs [S,B,F] | None,
  image)`` where ``image`` is ``[S, B, H, W, C]`` (NHWC like ``VisionAgent.__call__``, ``networks/vision_agent.py:17``) or
  the patch matrix ``[S, B, T, P*P*C]`` the fused env step renders.  The image goes through the ViT encoder and its ``[S*B, D]``
  output is concatenated to the vector observation IN FRONT OF the first Dense ("the encoder output replaces / extends the
  first Dense input") -- as a two-source GEMM, no concatenated copy is materialised.
* ``dones``: ``[S, B]`` bool; ``hidden``: ``[B, GRU_HIDDEN_DIM]``.  Returns the new hidden state, the Categorical logits
  ``[S, B, A]`` (the reference wraps them in ``distrax.Categorical``; sampling stays with the caller's PRNG) and the value
  ``[S, B]``.
* ``params``: flax-named pytree ``{'params': {'vit': <ViT pytree>, 'Dense_0', 'ScannedRNN_0': {'GRUCell_0': {ir,iz,in,hr,hz,hn}},
  'Dense_1', 'Dense_2' (actor), 'Dense_3', 'Dense_4' (critic)}}`` -- kernels ``[in, out]``.

Every FLOP runs in libvitmarl_b200.so (ViT: tcgen05 kernels; head: fp32 CUDA-core kernels, csrc/policy_head.cu); torch only
owns the buffers.  Forward only: the PPO loss / optimiser stay with the caller (boundary-only rows A13-A14)."""
from __future__ import annotations

import math
from typing import Dict, Optional, Tuple

import torch

from . import _capi
from . import vit as vvit

__all__ = ["ActorCriticRNN", "init_head_params", "calculate_gae"]

ACT_NONE, ACT_RELU = 0, 1


def _orthogonal(shape, scale, gen):
    a = torch.randn(shape, generator=gen)
    q, r = torch.linalg.qr(a if shape[0] >= shape[1] else a.t())
    q = q * torch.sign(torch.diagonal(r))
    return (q if shape[0] >= shape[1] else q.t()) * scale


def init_head_params(in_dim: int, action_dim: int, fc_dim: int = 128, gru_dim: int = 128, seed: int = 0, device="cuda") -> Dict:
    """Initialisers of the reference module (orthogonal(sqrt 2) / orthogonal(2) / orthogonal(0.01) / orthogonal(1), zero biases,
    ippo_rnn_JAXMARL.py:85-113; flax GRUCell defaults: lecun-normal input kernels, orthogonal recurrent kernels)."""
    g = torch.Generator().manual_seed(seed)
    d = lambda t: t.to(device)
    z = lambda n: torch.zeros(n, device=device)
    H = fc_dim                                              # GRUCell(features=ins.shape[1]): the carry is FC_DIM_SIZE wide
    gru = {}
    for k in ("ir", "iz", "in"):
        gru[k] = {"kernel": d(torch.randn((fc_dim, H), generator=g) / math.sqrt(fc_dim)), "bias": z(H)}
    for k in ("hr", "hz"):
        gru[k] = {"kernel": d(_orthogonal((H, H), 1.0, g))}
    gru["hn"] = {"kernel": d(_orthogonal((H, H), 1.0, g)), "bias": z(H)}
    return {"Dense_0": {"kernel": d(_orthogonal((in_dim, fc_dim), math.sqrt(2.0), g)), "bias": z(fc_dim)},
            "ScannedRNN_0": {"GRUCell_0": gru},
            "Dense_1": {"kernel": d(_orthogonal((H, gru_dim), 2.0, g)), "bias": z(gru_dim)},
            "Dense_2": {"kernel": d(_orthogonal((gru_dim, action_dim), 0.01, g)), "bias": z(action_dim)},
            "Dense_3": {"kernel": d(_orthogonal((H, fc_dim), 2.0, g)), "bias": z(fc_dim)},
            "Dense_4": {"kernel": d(_orthogonal((fc_dim, 1), 1.0, g)), "bias": z(1)}}


def _dense(x0: torch.Tensor, W: torch.Tensor, b: Optional[torch.Tensor], act: int, x1: Optional[torch.Tensor] = None) -> torch.Tensor:
    R, K0 = x0.shape
    K1 = 0 if x1 is None else x1.shape[1]
    N = W.shape[1]
    if W.shape[0] != K0 + K1:
        raise _capi.VitmarlError(_capi.EINVAL, f"Dense kernel has {W.shape[0]} inputs, got {K0}+{K1}")
    y = torch.empty((R, N), dtype=torch.float32, device=x0.device)
    rc = _capi.lib().vitmarl_dense_f32(torch.cuda.current_stream().cuda_stream, R, K0, K1, N, x0.data_ptr(), x0.stride(0),
                                       None if x1 is None else x1.data_ptr(), 0 if x1 is None else x1.stride(0),
                                       W.data_ptr(), None if b is None else b.data_ptr(), act, y.data_ptr(), N)
    _capi.check(rc)
    return y


class ActorCriticRNN:
    def __init__(self, action_dim: int, config: Dict, vit_cfg: Optional[vvit.ViTConfig] = None):
        self.action_dim = action_dim
        self.config = config
        self.vit_cfg = vit_cfg
        self.encoder = vvit.ViTEncoder(vit_cfg) if vit_cfg is not None else None
        self._gru_cache = None

    @staticmethod
    def initialize_carry(batch_size: int, hidden_size: int, device="cuda") -> torch.Tensor:
        """ScannedRNN.initialize_carry (ippo_rnn_JAXMARL.py:68-72): zeros."""
        return torch.zeros((batch_size, hidden_size), dtype=torch.float32, device=device)

    def init(self, seed: int, vec_dim: int, device="cuda") -> Dict:
        in_dim = vec_dim + (self.vit_cfg.dim if self.vit_cfg is not None else 0)
        p = init_head_params(in_dim, self.action_dim, self.config.get("FC_DIM_SIZE", 128), self.config.get("GRU_HIDDEN_DIM", 128), seed, device)
        if self.vit_cfg is not None:
            p["vit"] = vvit.init_params(self.vit_cfg, seed, device)
        return {"params": p}

    # ---------------------------------------------------------------------------------------------------------------
    def encode(self, params: Dict, image: torch.Tensor, *, packed=None, params_version: Optional[int] = None) -> torch.Tensor:
        """image [R,H,W,C] or patch matrix [R,T,P*P*C] -> [R,D] through the ViT encoder."""
        c = self.vit_cfg
        patches = image.dim() == 3
        if packed is None:
            packed = vvit.pack_params(c, params["vit"])
        return self.encoder.apply_packed(packed, image if not patches else image.contiguous(), patches=patches, params_version=params_version)

    def apply(self, variables: Dict, hidden: torch.Tensor, x, *, packed_vit=None, params_version: Optional[int] = None
              ) -> Tuple[torch.Tensor, torch.Tensor, torch.Tensor]:
        p = variables["params"]
        obs, dones = x
        vec, image = (obs if isinstance(obs, tuple) else (obs, None))
        if image is not None and self.encoder is None:
            raise _capi.VitmarlError(_capi.EINVAL, "an image observation needs a vit_cfg")
        ref = vec if vec is not None else image
        if not ref.is_cuda:
            raise _capi.VitmarlError(_capi.ENODEVICE, "ActorCriticRNN.apply needs CUDA tensors (there is no CPU fallback)")
        S, B = ref.shape[0], ref.shape[1]
        f32 = lambda t: t.to(torch.float32).contiguous()
        enc = None
        if image is not None:
            enc = self.encode(p, image.reshape((S * B,) + tuple(image.shape[2:])), packed=packed_vit, params_version=params_version)
        if vec is not None:
            x0, x1 = f32(vec.reshape(S * B, -1)), enc
        else:
            x0, x1 = enc, None
        emb = _dense(x0, f32(p["Dense_0"]["kernel"]), f32(p["Dense_0"]["bias"]), ACT_RELU, x1)       # [S*B, FC]
        g = p["ScannedRNN_0"]["GRUCell_0"]
        Wi = torch.cat([f32(g[k]["kernel"]) for k in ("ir", "iz", "in")], dim=1)                      # [FC, 3H]
        bi = torch.cat([f32(g[k]["bias"]) for k in ("ir", "iz", "in")])
        Wh = torch.cat([f32(g[k]["kernel"]) for k in ("hr", "hz", "hn")], dim=1)                      # [H, 3H]
        bhn = f32(g["hn"]["bias"])
        H = Wh.shape[0]
        gi = _dense(emb, Wi, bi, ACT_NONE)                                                            # all time steps at once
        h = f32(hidden).clone()
        ys = torch.empty((S, B, H), dtype=torch.float32, device=h.device)
        rs = dones.reshape(S, B).to(torch.uint8).contiguous()
        lib, st = _capi.lib(), torch.cuda.current_stream().cuda_stream
        for t in range(S):                                                                           # nn.scan over time
            gh = _dense(h, Wh, None, ACT_NONE)
            _capi.check(lib.vitmarl_gru_cell_f32(st, B, H, gi[t * B:(t + 1) * B].data_ptr(), gh.data_ptr(), bhn.data_ptr(), h.data_ptr(),
                                                 rs[t].data_ptr(), ys[t].data_ptr()))
            h = ys[t]
        y = ys.reshape(S * B, H)
        a = _dense(y, f32(p["Dense_1"]["kernel"]), f32(p["Dense_1"]["bias"]), ACT_RELU)
        logits = _dense(a, f32(p["Dense_2"]["kernel"]), f32(p["Dense_2"]["bias"]), ACT_NONE)
        c = _dense(y, f32(p["Dense_3"]["kernel"]), f32(p["Dense_3"]["bias"]), ACT_RELU)
        value = _dense(c, f32(p["Dense_4"]["kernel"]), f32(p["Dense_4"]["bias"]), ACT_NONE)
        return h.clone(), logits.reshape(S, B, -1), value.reshape(S, B)


def calculate_gae(gamma: float, gae_lambda: float, reward: torch.Tensor, value: torch.Tensor, done: torch.Tensor, last_val: torch.Tensor):
    """``_calculate_gae(gamma, gae_lambda, traj_batch, last_val)`` (ippo_rnn_JAXMARL.py:372-394) -> (advantages, targets), all
    ``[S, B]`` fp32, ``done`` = ``traj_batch.global_done``.  One kernel launch for the whole reverse scan."""
    if not reward.is_cuda:
        raise _capi.VitmarlError(_capi.ENODEVICE, "calculate_gae needs CUDA tensors (there is no CPU fallback)")
    S, B = reward.shape
    r, v = reward.to(torch.float32).contiguous(), value.to(torch.float32).contiguous()
    d, lv = done.to(torch.uint8).contiguous(), last_val.to(torch.float32).contiguous()
    adv, tgt = torch.empty_like(r), torch.empty_like(r)
    _capi.check(_capi.lib().vitmarl_gae_f32(torch.cuda.current_stream().cuda_stream, S, B, float(gamma), float(gae_lambda), r.data_ptr(),
                                            v.data_ptr(), d.data_ptr(), lv.data_ptr(), adv.data_ptr(), tgt.data_ptr()))
    return adv, tgt
