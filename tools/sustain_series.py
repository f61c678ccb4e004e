"""ms per 16-slide forward step over a long back-to-back run (one CUDA event per 20 steps): how the power-capped part settles.
python tools/sustain_series.py [steps]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench
from snuffy_b200 import snuffy
dev = torch.device("cuda", 0)
torch.cuda.set_device(0)
steps = int(sys.argv[1]) if len(sys.argv) > 1 else 3000
model, _ = bench.build_model(dev)
model.eval()
xs = [torch.randn(16, 10000, 512, device=dev) for _ in range(2)]
graphs = []
with torch.no_grad():
    for x in xs:
        graphs.append(bench.capture(lambda x=x: snuffy.forward_bags(model, x)))
torch.cuda.synchronize()
ev = [torch.cuda.Event(enable_timing=True) for _ in range(steps // 20 + 1)]
ev[0].record()
for i in range(steps):
    graphs[i & 1].replay()
    if (i + 1) % 20 == 0:
        ev[(i + 1) // 20].record()
torch.cuda.synchronize()
ms = [ev[k].elapsed_time(ev[k + 1]) / 20 for k in range(len(ev) - 1)]
print("ms/step per 20-step window:", " ".join(f"{m:.3f}" for m in ms))
print(f"first 3 windows {sum(ms[:3]) / 3:.3f}  windows 20-40 {sum(ms[20:40]) / 20:.3f}  last 20 {sum(ms[-20:]) / 20:.3f}  min {min(ms):.3f} max {max(ms):.3f}")
