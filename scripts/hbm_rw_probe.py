"""HBM probe with library kernels (measurement only): pure-write, pure-read and copy bandwidth on this GPU -- the denominator
MEASURED_PEAKS.json gives is a COPY (half read, half write); write-heavy kernels of this repo plateau near the pure-write figure."""
import torch
n = 1 << 30                                   # 2 GiB of bf16
a = torch.empty(n, dtype=torch.bfloat16, device="cuda"); b = torch.empty_like(a)
def t(fn, reps=10):
    for _ in range(3): fn()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps * 1e-3
w = t(lambda: a.zero_())
r = t(lambda: a.view(torch.int32).max())
c = t(lambda: b.copy_(a))
print(f"pure write {2*n/w/1e9:.0f} GB/s   pure read {2*n/r/1e9:.0f} GB/s   copy (read+write bytes) {4*n/c/1e9:.0f} GB/s")
