import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from snuffy_b200 import ops
def run(B, n, ks, h, d, scale, vscale=None):
    rs = np.random.RandomState(5)
    qv_np = (rs.standard_normal((B * n, 2 * d)) * scale).astype(np.float32)
    if vscale is not None:
        qv_np[:, d:] *= vscale / scale
    qv = torch.from_numpy(qv_np).cuda()
    kp = torch.from_numpy((rs.standard_normal((B * ks, d)) * scale).astype(np.float32)).cuda()
    _, planes, _ = ops.ln_rows(qv, None, None, apply_ln=False, want_planes=True, zero_planes=True)
    o1, p1, s1 = ops.sparse_attn_tc(planes, kp, B, n, ks, h, d, want_probs=True, want_stats=True)
    o2, p2, s2 = ops.sparse_attn(qv[:, :d], qv[:, d:], kp, B, n, ks, h, want_probs=True, want_stats=True)
    bad = (~torch.isfinite(o1))
    print(f"B={B} n={n} ks={ks} h={h} d={d} scale={scale} vscale={vscale}: o finite {torch.isfinite(o1).all().item()} bad rows {bad.any(-1).sum().item()}/{o1.shape[0]} "
          f"bad cols {bad.any(0).sum().item()}/{d} maxerr {(torch.nan_to_num(o1)-o2).abs().max().item():.3g} omax {o2.abs().max().item():.3g}")
run(3, 200, 40, 2, 128, 10)
run(1, 256, 40, 2, 128, 10)
run(1, 200, 40, 2, 128, 10)
run(1, 256, 40, 2, 128, 10, vscale=1)
run(1, 256, 40, 2, 128, 1, vscale=10)
run(1, 256, 128, 2, 128, 10)
run(1, 256, 32, 1, 64, 10)
