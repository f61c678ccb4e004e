"""A few train steps at cfg2 (one bag per step): target for the ncu launch list of the backward pass; prints the CUDA-event
time per step and the host-side issue time per step (a step that is host-bound shows issue ~= event time)."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench
from snuffy_b200 import dp
dev = torch.device("cuda", 0)
torch.cuda.set_device(0)
steps = int(os.environ.get("STEPS", 3))
ms, n, loss, _, _ = bench.train_throughput(dev, 1, steps=steps, warm=2, graph=False)
print("train ms/step", ms / n, "loss", loss)
if os.environ.get("PROF_ONLY"):
    sys.exit(0)
model, _ = bench.build_model(dev)
for l in model.b_classifier.encoder.layers:
    l.return_attn = False
tr = dp.DataParallelTrainer(model, lr=2e-4)
x = torch.randn(1, 10000, 512, device=dev); y = torch.ones(1, 1, device=dev)
for _ in range(3):
    tr.train_step(x, y)
torch.cuda.synchronize()
t0 = time.perf_counter()
for _ in range(10):
    tr.train_step(x, y)
t1 = time.perf_counter()
torch.cuda.synchronize()
t2 = time.perf_counter()
print("host issue ms/step", (t1 - t0) / 10 * 1e3, "wall ms/step", (t2 - t0) / 10 * 1e3)
