"""Small end-to-end invocation of every kernel for compute-sanitizer (memcheck) runs."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from vitmarl_b200 import jaxob, synth, vit, rollout, env as venv
from vitmarl_b200.config import World_EnvironmentConfig
cfg = World_EnvironmentConfig()
E, M = 33, 13
l2 = synth.make_l2_books(E, 5)
init = torch.from_numpy(synth.init_msgs_from_l2_batched(l2)).cuda()
a, b, t = jaxob.scan_through_entire_array(cfg, None, init, (jaxob.init_orderside(100, E), jaxob.init_orderside(100, E), None))
vcfg = vit.ViTConfig(64, 64, 2, 8, 192, 2, 3, 768)
params = vit.init_params(vcfg, 0, "cuda")
eng = rollout.RolloutEncoder(cfg, vcfg, params, E, M)
eng.reset(a, b)
stream = synth.MessageStream(E, 5)
for _ in range(2):
    y = eng.step(torch.from_numpy(stream.next(M)).cuda(), stat_agent_ids=[-100, 1000001], keep_trades=True)
enc = vit.ViTEncoder(vcfg)
x = eng.last_image()
enc.apply({"params": params}, x, train=True)
g = enc.vjp({"params": params}, torch.randn(E, 192, device="cuda"), want_dx=True)
ct = torch.zeros((E, 2), dtype=torch.int32, device="cuda")
jaxob.getCancelMsgs(eng.state.ask_raw_orders, -2, 4, -1, ct); jaxob.get_agent_trades(eng.state.trades, 1000001)
# unfused training path, policy head, GAE
enc.options.fused = 0
enc.apply({"params": params}, x, train=True); enc.vjp({"params": params}, torch.randn(E, 192, device="cuda"))
from vitmarl_b200 import actor_critic
net = actor_critic.ActorCriticRNN(5, {"FC_DIM_SIZE": 128, "GRU_HIDDEN_DIM": 128}, vit_cfg=vcfg)
v = net.init(0, 7)
net.apply(v, net.initialize_carry(E, 128), ((torch.randn(1, E, 7, device="cuda"), eng.last.image[None]), torch.zeros(1, E, dtype=torch.bool, device="cuda")))
actor_critic.calculate_gae(0.99, 0.95, torch.randn(9, E, device="cuda"), torch.randn(9, E, device="cuda"), torch.zeros(9, E, dtype=torch.bool, device="cuda"), torch.randn(E, device="cuda"))
torch.cuda.synchronize()
print("sanitize run ok", float(y.abs().sum()))
