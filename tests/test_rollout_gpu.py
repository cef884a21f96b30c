"""Public rollout-and-encode call (vitmarl_b200/rollout.py): the host-buffer entry `step_host` -- H2D of the messages, the
step, D2H of the encoding and the vision tensor, pipelined over two side streams and double-buffered staging -- must return,
step for step, exactly what the device-resident `step` returns (same kernels, same inputs: bit-identical)."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

from vitmarl_b200 import jaxob, rollout, synth, vit  # noqa: E402
from vitmarl_b200.config import World_EnvironmentConfig  # noqa: E402


def _engine(E, M, cfg, vcfg, params):
    l2 = synth.make_l2_books(E, 5)
    init = torch.from_numpy(synth.init_msgs_from_l2_batched(l2)).cuda()
    a, b, _ = jaxob.scan_through_entire_array(cfg, None, init, (jaxob.init_orderside(cfg.nOrders, E), jaxob.init_orderside(cfg.nOrders, E), None))
    eng = rollout.RolloutEncoder(cfg, vcfg, params, E, M)
    eng.reset(a, b)
    return eng


@pytest.mark.parametrize("E,M,steps", [(70, 13, 7), (33, 5, 4)])
def test_step_host_pipeline_equals_device_steps(E, M, steps):
    cfg = World_EnvironmentConfig()
    vcfg = vit.ViTConfig(64, 64, 2, 8, 192, 2, 3, 768)
    params = vit.init_params(vcfg, 0, "cuda")
    stream = synth.MessageStream(E, 11)
    msgs = [torch.from_numpy(stream.next(M)) for _ in range(steps)]
    dev = _engine(E, M, cfg, vcfg, params)
    want = []
    for m in msgs:
        f = dev.step(m.cuda())
        want.append((f.clone(), dev.last.vision_obs.clone()))
    host = _engine(E, M, cfg, vcfg, params)
    pinned = [m.pin_memory() for m in msgs]
    got = []
    for i, m in enumerate(pinned):
        got.append(host.step_host(m))             # no synchronise between calls: the copies of neighbouring steps overlap
        if i % 2 == 1 or i == steps - 1:          # a staging slot is valid until the call after next: read both before reuse
            host.wait_host()
            torch.cuda.synchronize()
            for j in range(max(0, i - 1), i + 1):
                got[j] = (got[j][0].clone(), got[j][1].clone())
    for (f, o), (fw, ow) in zip(got, want):
        assert torch.equal(f, fw.cpu()) and torch.equal(o, ow.cpu())
    # the device-resident books of both engines went through the same steps
    assert torch.equal(host.state.ask_raw_orders, dev.state.ask_raw_orders) and torch.equal(host.state.trades, dev.state.trades)
