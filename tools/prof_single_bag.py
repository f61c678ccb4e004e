import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench
dev = torch.device("cuda", 0)
model, _ = bench.build_model(dev)
for l in model.b_classifier.encoder.layers:
    l.return_attn = False
x = torch.randn(1, 10000, 512, device=dev)
with torch.no_grad():
    for _ in range(4):
        model(x)
torch.cuda.synchronize()
