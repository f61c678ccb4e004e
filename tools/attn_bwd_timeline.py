"""Per-phase timeline of the fused attention backward (CTA 0, row warp 2) from a -DATTN_DEBUG_TIMING variant library:
SNUFFY_B200_LIB=tools/variants/libdbgb.so python tools/attn_bwd_timeline.py      (B bags of cfg2 shape; B = 1 is the training step)"""
import ctypes, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from snuffy_b200 import ops, _lib
B, n, d, h, ks = int(os.environ.get("B", 1)), 10000, 512, 8, 200
dev = torch.device("cuda")
g = torch.Generator(device=dev).manual_seed(0)
qv = torch.randn(B * n, 2 * d, device=dev, generator=g)
_, qvp, _ = ops.ln_rows(qv, None, None, apply_ln=False, want_planes=True, zero_planes=True)
kp = torch.randn(B * ks, d, device=dev, generator=g)
d_o = torch.randn(B * ks, d, device=dev, generator=g)
P = float(os.environ.get("P", 0.1))
_, _, stats, mask = ops.sparse_attn_tc(qvp, kp, B, n, ks, h, d, want_probs=False, want_stats=True, dropout_p=P, seed=3, offset=7,
                                       want_mask=True)
for _ in range(3):
    ops.sparse_attn_bwd_fused(qvp, kp, d_o, stats, B, n, ks, h, d, P, mask)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(10):
    ops.sparse_attn_bwd_fused(qvp, kp, d_o, stats, B, n, ks, h, d, P, mask)
e1.record(); torch.cuda.synchronize()
print("fused attention backward ms per call", e0.elapsed_time(e1) / 10)
raw = ctypes.CDLL(_lib.LIB_PATH)
if hasattr(raw, "snuffy_attn_bwd_debug_read"):
    buf = (ctypes.c_longlong * (64 * 8))()
    assert raw.snuffy_attn_bwd_debug_read(buf) == 0
    a = np.array(buf[:], dtype=np.int64).reshape(64, 8)
    t0 = a[0, 0]
    names = ["S ready", "P~ stored", "dV ready", "delta done", "G ready", "dS stored", "dQ ready", "dQ out"]
    for t in range(int(os.environ.get("T1", 8))):
        if a[t, 0] > 0:
            print(f"tile {t:2d} | " + " ".join(f"{names[k]}={a[t, k] - t0}" for k in range(8)))
