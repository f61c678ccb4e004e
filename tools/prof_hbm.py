"""The HBM-bound kernels of the path alone at the bench shape (16 slides of 10000 x 512): target for ncu --set full."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from snuffy_b200 import ops
B, n, d = 16, 10000, 512
dev = torch.device("cuda")
g = torch.Generator(device=dev).manual_seed(0)
x = [torch.randn(B * n, d, device=dev, generator=g) for _ in range(2)]
w = torch.randn(1, d, device=dev, generator=g); b = torch.zeros(1, device=dev)
gam, bet = torch.ones(d, device=dev), torch.zeros(d, device=dev)
wh = torch.randn(1, d, device=dev, generator=g)
for i in range(4):
    xi = x[i & 1]
    ops.scores(xi, w, b)                                             # scores_kernel
    ops.scores_ln_planes(xi, w, b)                                   # ln_rows_kernel with the fused scorer (layer 0)
    ops.ln_rows(xi, None, None, want_planes=True, affine=False)      # ln_rows_kernel (layers >= 1)
    ops.ln_mean_head(xi.view(B, n, d), gam, bet, wh, b)              # ln_mean_head_reg_kernel
torch.cuda.synchronize()
