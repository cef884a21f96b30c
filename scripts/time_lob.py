"""Ad-hoc kernel timing for the order-book kernels (CUDA events, L2 flushed between launches)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from vitmarl_b200 import jaxob, synth, env as venv
from vitmarl_b200.config import World_EnvironmentConfig
CFG = World_EnvironmentConfig()
HBM = 6548.2

def run(E, M, N=100, fused=False, image=False, iters=20, quiet=False, patch=None):
    l2 = synth.make_l2_books(E, 7)
    init = torch.from_numpy(synth.init_msgs_from_l2_batched(l2)).cuda()
    a, b, t = jaxob.scan_through_entire_array(CFG, None, init, (jaxob.init_orderside(N, E), jaxob.init_orderside(N, E), None))
    stream = synth.MessageStream(E, 7)
    msgs = torch.from_numpy(stream.next(M)).cuda()
    flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device="cuda")
    state = venv.reset(CFG, a.clone(), b.clone(), M)
    ts = []
    for i in range(iters + 3):
        flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        if fused:
            venv.step(CFG, state, msgs, image_hw=(64, 64) if image else None, inplace=False, image_patch=patch)
        else:
            jaxob.scan_through_entire_array_save_bidask(CFG, None, msgs, (a, b, None), M)
        e1.record(); torch.cuda.synchronize()
        if i >= 3: ts.append(e0.elapsed_time(e1) * 1e-3)
    t = float(np.median(ts))
    byts = E * (4 * N * 24 + 100 * 32 + M * 32 + 2 * M * 8)
    if quiet:
        return t, byts
    print(f"E={E} M={M} N={N} fused={fused} image={image} patch={patch}: {t*1e6:.1f} us  {E*M/t/1e9:.3f} Gmsg/s  {E/t/1e6:.2f} Menv-steps/s  "
          f"alg {byts/t/1e9:.0f} GB/s = {byts/t/1e9/HBM:.3f} of HBM peak")

if __name__ == "__main__":
    if "--one" in sys.argv:     # --one E M [N]: a single configuration (used under ncu)
        a = [int(x) for x in sys.argv[sys.argv.index("--one") + 1:]]
        run(a[0], a[1], N=a[2] if len(a) > 2 else 100, iters=2); sys.exit(0)
    if "--sweep" in sys.argv:   # BASELINE configs[2]: capacity 10..100 rows/side, 1k..64k envs, M = 100 data messages per launch
        out = ["| N (rows/side) | envs | us/launch | G msgs/s | algorithmic GB/s | of HBM (6548 GB/s) |", "|---|---|---|---|---|---|"]
        for N in (10, 20, 50, 100):
            for E in (1024, 4096, 16384, 65536):
                t, b = run(E, 100, N=N, iters=10, quiet=True)
                out.append(f"| {N} | {E} | {t*1e6:.1f} | {E*100/t/1e9:.2f} | {b/t/1e9:.0f} | {b/t/1e9/HBM:.3f} |")
                print(out[-1], flush=True)
        for E in (4096, 16384, 65536):       # the MAPPO step shape (M = 13) for comparison
            t, b = run(E, 13, iters=10, quiet=True)
            out.append(f"| 100 (M=13) | {E} | {t*1e6:.1f} | {E*13/t/1e9:.2f} | {b/t/1e9:.0f} | {b/t/1e9/HBM:.3f} |")
        os.makedirs("gpurun_out", exist_ok=True)
        open("gpurun_out/lob_sweep.md", "w").write("\n".join(out) + "\n")
        sys.exit(0)
    if "--quick" in sys.argv:
        run(4096, 13, iters=2); run(4096, 13, fused=True, image=True, iters=2); run(16384, 100, iters=2); sys.exit(0)
    run(4096, 13); run(4096, 13, fused=True); run(4096, 13, fused=True, image=True); run(4096, 13, fused=True, image=True, patch=8)
    run(4096, 113); run(16384, 13); run(65536, 13); run(65536, 100); run(16384, 100, N=50); run(16384, 100, N=10)
