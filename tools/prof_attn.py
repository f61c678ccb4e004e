"""Run only the tensor-core attention kernel at cfg2 (8 slides) a few times: target for `ncu -k regex:attn_tc`."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from snuffy_b200 import ops

B, n, d, h, ks = int(os.environ.get("B", 8)), 10000, 512, 8, 200
DROP = float(os.environ.get("DROP", 0))       # > 0: the training form (row statistics + keep bits written)
dev = torch.device("cuda")
g = torch.Generator(device=dev).manual_seed(0)
x = torch.randn(B * n, d, device=dev, generator=g)
w = torch.randn(2 * d, d, device=dev, generator=g) * 0.05
gam, bet = torch.ones(d, device=dev), torch.zeros(d, device=dev)
_, up, _ = ops.ln_rows(x, gam, bet, want_planes=True)
wp = ops.weight_planes(w)
_, _, qvp = ops.gemm_tc(up, wp, M=B * n, N=2 * d, K=d, passes=3, want_out=False, want_planes=True)
kp = torch.randn(B * ks, d, device=dev, generator=g)
kw = dict(want_probs=False, want_stats=True, dropout_p=DROP, seed=1, offset=2, want_mask=not os.environ.get("NOMASK")) if DROP > 0 else dict(want_probs=False)
if os.environ.get("STATS"):
    kw = dict(want_probs=False, want_stats=True)
for _ in range(3):
    ops.sparse_attn_tc(qvp, kp, B, n, ks, h, d, **kw)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(10):
    ops.sparse_attn_tc(qvp, kp, B, n, ks, h, d, **kw)
e1.record(); torch.cuda.synchronize()
print("attn_tc ms", e0.elapsed_time(e1) / 10, "B", B, kw)
