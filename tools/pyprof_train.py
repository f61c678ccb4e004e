import os, sys, cProfile, pstats
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench
from snuffy_b200 import dp
dev = torch.device("cuda", 0)
model, _ = bench.build_model(dev)
for l in model.b_classifier.encoder.layers:
    l.return_attn = False
tr = dp.DataParallelTrainer(model, lr=2e-4)
x = torch.randn(1, 10000, 512, device=dev); y = torch.ones(1, 1, device=dev)
for _ in range(3):
    tr.train_step(x, y)
torch.cuda.synchronize()
pr = cProfile.Profile()
pr.enable()
for _ in range(20):
    tr.train_step(x, y)
pr.disable()
torch.cuda.synchronize()
st = pstats.Stats(pr)
st.sort_stats("tottime").print_stats(28)
