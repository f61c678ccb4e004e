"""Per-phase timeline of the tensor-core attention kernel (CTA 0) from the -DATTN_DEBUG_TIMING variant library.
SNUFFY_B200_LIB=tools/variants/libdbg.so python tools/attn_timeline.py"""
import ctypes, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from snuffy_b200 import ops, _lib
B, n, d, h, ks = 8, 10000, 512, 8, int(os.environ.get("KS", 200))
dev = torch.device("cuda")
g = torch.Generator(device=dev).manual_seed(0)
qv = torch.randn(B * n, 2 * d, device=dev, generator=g)
_, qvp, _ = ops.ln_rows(qv, None, None, apply_ln=False, want_planes=True, zero_planes=True)
kp = torch.randn(B * ks, d, device=dev, generator=g)
for _ in range(3):
    ops.sparse_attn_tc(qvp, kp, B, n, ks, h, d, want_probs=False)
torch.cuda.synchronize()
raw = ctypes.CDLL(_lib.LIB_PATH)
buf = (ctypes.c_longlong * (4 * 64 * 8))()
assert raw.snuffy_attn_debug_read(buf) == 0
a = np.array(buf[:], dtype=np.int64).reshape(4, 64, 8)
t0 = a[1, 0, 0]
names = {0: ["wait s_full", "s_full", "S in regs", "sum done", "wait p_empty", "p_empty", "P stored"],
         1: ["wait q_full", "q_full", "s_empty", "MMA1 issued", "wait p_full", "p_full", "v_full", "MMA2 issued"],
         2: ["wait q_empty", "q_empty", "v_empty", "o_full (prev item)", "O staged", "O written", "K split"]}
for t in range(int(os.environ.get("T0", 2)), int(os.environ.get("T1", 14))):
    print(f"tile {t:2d} | softmax " + " ".join(f"{names[0][k]}={a[0,t,k]-t0}" for k in range(7)))
    print(f"        | mma     " + " ".join(f"{names[1][k]}={a[1,t,k]-t0}" for k in range(8)))
    print(f"        | prod    " + " ".join(f"{names[2][k]}={a[2,t,k]-t0}" for k in range(7)))
    if a[3, t, 0] > 0:
        print(f"        | item    " + " ".join(f"{n_}={a[3,t,k]-t0}" for k, n_ in enumerate(["ksplit start", "ksplit loop done", "fence done", "O copy start", "O copy done"])))
