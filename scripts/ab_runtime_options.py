"""A/B of per-call options under the SUSTAINED rollout loop (power-capped clocks), interleaved on one box:
usage: python scripts/ab_runtime_options.py   -> env-steps/s for attn_flags / pdl settings, 3 rounds of ~1.5 s each."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from vitmarl_b200 import jaxob, rollout, synth, vit
from vitmarl_b200.config import World_EnvironmentConfig
E, M = 4096, 13
cfg = World_EnvironmentConfig(); vcfg = vit.VIT_TINY_8
l2 = synth.make_l2_books(E, 5)
init = torch.from_numpy(synth.init_msgs_from_l2_batched(l2)).cuda()
a, b, _ = jaxob.scan_through_entire_array(cfg, None, init, (jaxob.init_orderside(cfg.nOrders, E), jaxob.init_orderside(cfg.nOrders, E), None))
eng = rollout.RolloutEncoder(cfg, vcfg, vit.init_params(vcfg, 0, "cuda"), E, M)
stream = synth.MessageStream(E, 7)
msgs = [torch.from_numpy(stream.next(M)).cuda() for _ in range(16)]
def run(steps):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(steps): eng.step(msgs[i % 16])
    e1.record(); torch.cuda.synchronize()
    return E * steps / (e0.elapsed_time(e1) * 1e-3)
settings = [("flags4 pdl1", 4, 1), ("flags0 pdl1", 0, 1), ("flags5 pdl1", 5, 1), ("flags6 pdl1", 6, 1), ("flags4 pdl0", 4, 0)]
eng.reset(a.clone(), b.clone()); run(300)        # heat up: the comparison is about the power-capped steady state
res = {s[0]: [] for s in settings}
for rnd in range(3):
    for name, fl, pdl in settings:
        eng.encoder.options.attn_flags = fl; eng.encoder.options.pdl = pdl
        eng.reset(a.clone(), b.clone()); run(20)
        res[name].append(run(450))
for k, v in res.items(): print(f"{k:14s}", " ".join(f"{x/1e6:.4f}" for x in v), "M env-steps/s")
