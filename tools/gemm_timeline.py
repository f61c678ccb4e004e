"""Per-tile timeline of the tcgen05 GEMM (CTA 0) from the -DGEMM_DEBUG_TIMING variant library:
tools/build_variant.sh gdbg snuffy_b200/csrc/gemm_tc.cu -DGEMM_DEBUG_TIMING
SNUFFY_B200_LIB=tools/variants/libgdbg.so SHAPE=ffn_up python tools/gemm_timeline.py"""
import ctypes, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from snuffy_b200 import ops, _lib
dev = torch.device("cuda")
g = torch.Generator(device=dev).manual_seed(0)
R = int(os.environ.get("ROWS", 160000))
shape = os.environ.get("SHAPE", "ffn_up")
N, K, kw = {"ffn_up": (2048, 512, dict(want_out=False, want_planes=True, act="relu")),
            "ffn_up_train": (2048, 512, dict(want_out=False, want_planes=True, act="relu", drop=(0.1, 1, 2))),
            "ffn_up_train_preact": (2048, 512, dict(want_out=False, want_preact=True, want_planes=True, act="relu", drop=(0.1, 1, 2))),
            "relugrad": (2048, 512, None),
            "qv": (1024, 512, dict(want_out=False, want_planes=True)),
            "ffn_down": (512, 2048, dict())}[shape]
x = torch.randn(R, K, device=dev, generator=g)
a = ops.ln_rows(x, None, None, apply_ln=False, want_planes=True)[1]
w = ops.weight_planes(torch.randn(N, K, device=dev, generator=g) * 0.05)
bias = torch.randn(N, device=dev, generator=g)
resid = torch.randn(R, N, device=dev, generator=g) if shape == "ffn_down" else None
if shape == "relugrad":
    act_planes = ops.ln_rows(torch.relu(torch.randn(R, N, device=dev, generator=g)), None, None, apply_ln=False, want_planes=True)[1]
for _ in range(3):
    if shape == "relugrad":
        ops.gemm_tc_relugrad(a, w, act_planes, 0.1, M=R, N=N, K=K, want_colsum=True)
    else:
        ops.gemm_tc(a, w, M=R, N=N, K=K, passes=int(os.environ.get("PASSES", 3)), bias=bias, resid=resid, **kw)
torch.cuda.synchronize()
raw = ctypes.CDLL(_lib.LIB_PATH)
buf = (ctypes.c_longlong * (3 * 64 * 4))()
assert raw.snuffy_gemm_debug_read(buf) == 0
t = np.array(buf[:], dtype=np.int64).reshape(3, 64, 4)
t0 = t[1, 0, 0]
print(f"{shape}: R={R} N={N} K={K}; clocks relative to the MMA warp's first item")
print("item | mma: wait tempty  issue+retire | epilogue: wait tfull  work | period")
for i in range(2, 14):
    m, e = t[1, i], t[2, i]
    print(f"{i:4d} | {m[1]-m[0]:8d} {m[2]-m[1]:8d} | {e[1]-e[0]:8d} {e[2]-e[1]:8d} | {t[1, i, 0] - t[1, i-1, 0]:8d}")
