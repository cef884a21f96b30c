"""A13 wiring: ``ActorCriticRNN.apply(params, hidden, (obs, dones))`` with the ViT encoder in the `# FIXME: APPLY VISION` slot
(ippo_rnn_JAXMARL.py:75-115) against the fp32 oracle.  Head alone (fp32 CUDA-core kernels): 1e-5; with the bf16 tensor-core
encoder in front: logits / value within 2e-2 of their range (the encoder's activation tolerance)."""
import pytest
import torch

pytestmark = pytest.mark.gpu

from oracle import policy_oracle as PO      # noqa: E402
from oracle import vit_oracle as VO         # noqa: E402
from vitmarl_b200 import actor_critic, vit  # noqa: E402


def _jitter(p, seed):
    g = torch.Generator().manual_seed(seed)
    return VO.tree_map(lambda t: t + 0.05 * torch.randn(t.shape, generator=g).to(t.device), p)


@pytest.mark.parametrize("S,B,F", [(1, 300, 23), (5, 70, 610), (3, 64, 1)])
def test_head_alone_matches_fp32_oracle(S, B, F):
    net = actor_critic.ActorCriticRNN(9, {"FC_DIM_SIZE": 128, "GRU_HIDDEN_DIM": 128})
    v = net.init(0, F)
    v = {"params": _jitter(v["params"], 1)}
    g = torch.Generator().manual_seed(2)
    obs = torch.randn(S, B, F, generator=g).cuda()
    dones = (torch.rand(S, B, generator=g) < 0.3).cuda()
    h0 = torch.randn(B, 128, generator=g).cuda()
    h, logits, value = net.apply(v, h0, (obs, dones))
    hr, lr, vr = PO.actor_critic(v["params"], h0, obs, dones)
    assert logits.shape == (S, B, 9) and value.shape == (S, B) and h.shape == (B, 128)
    for a, b in ((h, hr), (logits, lr), (value, vr)):
        assert (a - b).abs().max().item() <= 1e-5 * max(1.0, b.abs().max().item())


def test_vision_wiring_image_and_patch_matrix():
    cfg = vit.VIT_PARITY
    net = actor_critic.ActorCriticRNN(5, {"FC_DIM_SIZE": 128, "GRU_HIDDEN_DIM": 128}, vit_cfg=cfg)
    v = net.init(3, 17)
    v = {"params": _jitter(v["params"], 4)}
    S, B = 2, 40
    g = torch.Generator().manual_seed(5)
    lens = torch.randint(0, cfg.img_w + 1, (S, B, cfg.img_h, 1, cfg.channels), generator=g)
    img = (torch.arange(cfg.img_w)[None, None, None, :, None] < lens).float().cuda()
    vec = torch.randn(S, B, 17, generator=g).cuda()
    dones = (torch.rand(S, B, generator=g) < 0.2).cuda()
    h0 = net.initialize_carry(B, 128)
    h, logits, value = net.apply(v, h0, ((vec, img), dones))
    enc_ref = VO.vit_forward(cfg, v["params"]["vit"], img.reshape(S * B, cfg.img_h, cfg.img_w, cfg.channels)).reshape(S, B, cfg.dim)
    hr, lr, vr = PO.actor_critic(v["params"], h0, vec, dones, enc=enc_ref)
    for a, b in ((h, hr), (logits, lr), (value, vr)):
        assert (a - b).abs().max().item() <= 2e-2 * max(b.abs().max().item(), 1e-3)
    # the same call on the patch matrix the fused env step renders (no patchify pass) gives the same result bit for bit
    p = cfg.patch
    pat = img.to(torch.bfloat16).reshape(S, B, cfg.img_h // p, p, cfg.img_w // p, p, cfg.channels).permute(0, 1, 2, 4, 3, 5, 6)
    pat = pat.reshape(S, B, cfg.tokens, p * p * cfg.channels).contiguous()
    h2, logits2, value2 = net.apply(v, h0, ((vec, pat), dones))
    assert torch.equal(logits, logits2) and torch.equal(value, value2) and torch.equal(h, h2)
    # image only (the encoder output REPLACES the vector observation)
    v2 = net.init(6, 0)
    h3, logits3, _ = net.apply(v2, h0, ((None, img), dones))
    _, lr3, _ = PO.actor_critic(v2["params"], h0, None, dones,
                                enc=VO.vit_forward(cfg, v2["params"]["vit"], img.reshape(S * B, cfg.img_h, cfg.img_w, cfg.channels)).reshape(S, B, cfg.dim))
    assert (logits3 - lr3).abs().max().item() <= 2e-2 * max(lr3.abs().max().item(), 1e-3)


def test_gae_bit_exact_vs_numpy_scan():
    """`_calculate_gae` (ippo_rnn_JAXMARL.py:372-394) for NUM_STEPS x (NUM_ENVS * agents) columns in one launch."""
    import numpy as np
    rng = np.random.default_rng(0)
    S, B = 128, 777
    reward = rng.normal(size=(S, B)).astype(np.float32)
    value = rng.normal(size=(S, B)).astype(np.float32)
    done = rng.random((S, B)) < 0.05
    last_val = rng.normal(size=(B,)).astype(np.float32)
    adv_w, tgt_w = PO.calculate_gae(0.99, 0.95, reward, value, done, last_val)
    adv, tgt = actor_critic.calculate_gae(0.99, 0.95, torch.from_numpy(reward).cuda(), torch.from_numpy(value).cuda(),
                                          torch.from_numpy(done).cuda(), torch.from_numpy(last_val).cuda())
    assert adv.cpu().numpy().tobytes() == adv_w.tobytes() and tgt.cpu().numpy().tobytes() == tgt_w.tobytes()
