"""ORACLE (test infrastructure, NOT product code): plain fp32 PyTorch restatement of ``ActorCriticRNN.__call__`` and
``ScannedRNN`` (gymnax_exchange/jaxrl/MARL/ippo_rnn_JAXMARL.py:48-115) on the flax-named parameter pytree, with
``flax.linen.GRUCell``'s equations (r/z/n gates; ``hr``, ``hz`` without bias, ``hn`` with bias; new_h = (1-z) n + z h).
The first Dense sees concat(vector obs, encoding) when an encoding is supplied (the ``# FIXME: APPLY VISION`` slot).
Parity unpinned by reference-held vectors (the reference has no tests; JAX/flax are not installable here)."""
import torch


def gru_cell(g, h, x):
    dense = lambda p, v: v @ p["kernel"] + (p["bias"] if "bias" in p else 0.0)
    r = torch.sigmoid(dense(g["ir"], x) + dense(g["hr"], h))
    z = torch.sigmoid(dense(g["iz"], x) + dense(g["hz"], h))
    n = torch.tanh(dense(g["in"], x) + r * dense(g["hn"], h))
    return (1.0 - z) * n + z * h


def actor_critic(p, hidden, vec, dones, enc=None):
    """vec [S,B,F] or None, enc [S,B,D] or None, dones [S,B] bool, hidden [B,H] -> (hidden, logits [S,B,A], value [S,B])"""
    x = vec if enc is None else (enc if vec is None else torch.cat([vec, enc], dim=-1))
    emb = torch.relu(x @ p["Dense_0"]["kernel"] + p["Dense_0"]["bias"])
    g = p["ScannedRNN_0"]["GRUCell_0"]
    h, ys = hidden, []
    for t in range(emb.shape[0]):
        h = torch.where(dones[t][:, None], torch.zeros_like(h), h)
        h = gru_cell(g, h, emb[t])
        ys.append(h)
    y = torch.stack(ys)
    a = torch.relu(y @ p["Dense_1"]["kernel"] + p["Dense_1"]["bias"])
    logits = a @ p["Dense_2"]["kernel"] + p["Dense_2"]["bias"]
    c = torch.relu(y @ p["Dense_3"]["kernel"] + p["Dense_3"]["bias"])
    value = (c @ p["Dense_4"]["kernel"] + p["Dense_4"]["bias"]).squeeze(-1)
    return h, logits, value


def calculate_gae(gamma, gae_lambda, reward, value, done, last_val):
    """NumPy float32 restatement of `_calculate_gae` (ippo_rnn_JAXMARL.py:372-394): reverse scan, the reference's order of operations.
    reward / value [S,B] float32, done [S,B] bool, last_val [B] -> advantages, targets."""
    import numpy as np
    f = np.float32
    reward, value, last_val = reward.astype(f), value.astype(f), last_val.astype(f)
    S = reward.shape[0]
    gae, next_value = np.zeros_like(last_val), last_val
    adv = np.empty_like(reward)
    gamma, lam = f(gamma), f(gae_lambda)
    for t in range(S - 1, -1, -1):
        nd = f(1) - done[t].astype(f)
        delta = reward[t] + gamma * next_value * nd - value[t]
        gae = delta + gamma * lam * nd * gae
        adv[t] = gae
        next_value = value[t]
    return adv, adv + value
