import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from snuffy_b200 import ops
def run(B, n, ks, h, d):
    rs = np.random.RandomState(5)
    qv = torch.from_numpy(rs.standard_normal((B * n, 2 * d)).astype(np.float32)).cuda()
    kp = torch.from_numpy(rs.standard_normal((B * ks, d)).astype(np.float32)).cuda()
    _, planes, _ = ops.ln_rows(qv, None, None, apply_ln=False, want_planes=True, zero_planes=True)
    o1, _, _ = ops.sparse_attn_tc(planes, kp, B, n, ks, h, d, want_probs=False)
    o2, _, _ = ops.sparse_attn(qv[:, :d], qv[:, d:], kp, B, n, ks, h, want_probs=False)
    torch.cuda.synchronize()
    print(f"B={B} n={n} ks={ks} h={h} d={d} dk={d//h}: maxerr {(o1-o2).abs().max().item():.3g} omax {o2.abs().max().item():.3g}", flush=True)
for args in [(1, 600, 392, 8, 768)]:
    run(*args)
