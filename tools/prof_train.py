"""A few train steps at cfg2 (one bag per step): target for the ncu launch list of the backward pass; prints the CUDA-event
time per step (eager launches, then graph replays) and the host-side issue time per eager step."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench
from snuffy_b200 import dp
dev = torch.device("cuda", 0)
torch.cuda.set_device(0)
steps = int(os.environ.get("STEPS", 3))
xs = [torch.randn(1, 10000, 512, device=dev) for _ in range(8)]
y = torch.ones(1, 1, device=dev)


def run(graph, steps, warm=2):
    model, _ = bench.build_model(dev)
    tr = dp.DataParallelTrainer(model, lr=2e-4, betas=(0.5, 0.9), weight_decay=5e-3, cuda_graph=graph)
    for i in range(warm):
        tr.train_step(xs[i & 7], y)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0 = time.perf_counter()
    e0.record()
    for i in range(steps):
        loss = tr.train_step(xs[i & 7], y)
    e1.record()
    t1 = time.perf_counter()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / steps, (t1 - t0) / steps * 1e3, float(loss)


ms, issue, loss = run(False, steps)
print("eager train ms/step", ms, "host issue ms/step", issue, "loss", loss)
if os.environ.get("PROF_ONLY"):
    sys.exit(0)
ms, issue, loss = run(True, 50, warm=3)
print("graph-replayed train ms/step", ms, "host issue ms/step", issue, "loss", loss)
