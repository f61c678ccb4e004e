"""Pinned H2D ceiling of the box vs the e2e leg of bench.py (164 MB per 8-slide step)."""
import torch, time, json
n = 8 * 10000 * 512
h = [torch.randn(n).pin_memory() for _ in range(2)]
d = [torch.empty(n, device="cuda") for _ in range(2)]
s = torch.cuda.Stream()
for it in range(3):
    d[0].copy_(h[0], non_blocking=True)
torch.cuda.synchronize()
t0 = time.perf_counter()
with torch.cuda.stream(s):
    for it in range(20):
        d[it & 1].copy_(h[it & 1], non_blocking=True)
s.synchronize()
dt = time.perf_counter() - t0
gbs = 20 * n * 4 / dt / 1e9
print(json.dumps({"pinned_h2d_GBps": round(gbs, 2), "slides_per_s_ceiling": round(gbs * 1e9 / (10000 * 512 * 4), 1)}))
