"""A few train steps at cfg2 (one bag per step): target for the ncu launch list of the backward pass."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench
dev = torch.device("cuda", 0)
torch.cuda.set_device(0)
ms, steps, loss, _ = bench.train_throughput(dev, 1, steps=int(os.environ.get("STEPS", 3)), warm=2)
print("train ms/step", ms / steps, "loss", loss)
