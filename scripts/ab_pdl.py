"""A/B on one box: end-to-end rollout step with and without programmatic dependent launch of the fused blocks."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from vitmarl_b200 import _capi, rollout, synth, vit
from vitmarl_b200.config import World_EnvironmentConfig

def main():
    lib = _capi.lib()
    E, M = 4096, 13
    cfg = World_EnvironmentConfig()
    vcfg = vit.VIT_TINY_8
    params = vit.init_params(vcfg, 0, "cuda")
    l2 = synth.make_l2_books(E, 1234)
    eng = rollout.RolloutEncoder(cfg, vcfg, params, E, M)
    from vitmarl_b200 import jaxob
    init = torch.from_numpy(synth.init_msgs_from_l2_batched(l2)).cuda()
    a, b, _ = jaxob.scan_through_entire_array(cfg, None, init, (jaxob.init_orderside(100, E), jaxob.init_orderside(100, E), None))
    stream = synth.MessageStream(E, 1234)
    msgs = [torch.from_numpy(stream.next(M)).cuda() for _ in range(8)]
    for flags in (4, 4 | 0x100, 4, 4 | 0x100):
        eng.encoder.options.pdl = 0 if (flags & 0x100) else 1      # per-call option
        eng.reset(a.clone(), b.clone())
        for i in range(5): eng.step(msgs[i % 8])
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for i in range(200): eng.step(msgs[i % 8])
        e1.record(); torch.cuda.synchronize()
        print("PDL" if not (flags & 0x100) else "no PDL", f"{e0.elapsed_time(e1)/200:.4f} ms/step", flush=True)

if __name__ == "__main__":
    main()
