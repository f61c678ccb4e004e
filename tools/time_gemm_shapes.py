"""CUDA-event times of the dense products of one cfg2 training step (one bag: 10 000 rows), each alone, 20 back-to-back launches.
SNUFFY_B200_TAIL_SPLIT=0 python tools/time_gemm_shapes.py for the whole-tile schedule."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from snuffy_b200 import ops

dev = torch.device("cuda")
g = torch.Generator(device=dev).manual_seed(0)
R = int(os.environ.get("ROWS", 10000))


def planes(rows, k):
    x = torch.randn(rows, k, device=dev, generator=g)
    return ops.ln_rows(x, None, None, apply_ln=False, want_planes=True)[1]


def timeit(fn, n=20):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n):
        fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n * 1e3


for name, N, K, kw in [("Q|V  [R,512]x[1024,512]", 1024, 512, dict(want_planes=True)),
                       ("FFN-up [R,512]x[2048,512]", 2048, 512, dict(want_out=False, want_preact=True, want_planes=True, act="relu")),
                       ("FFN-down [R,2048]x[512,2048]", 512, 2048, dict()),
                       ("du1 [R,1024]x[512,1024]", 512, 1024, dict()),
                       ("keys [200,512]x[512,512]", 512, 512, dict(rows=200))]:
    rows = kw.pop("rows", R)
    a = planes(rows, K)
    w = ops.weight_planes(torch.randn(N, K, device=dev, generator=g) * 0.05)
    us = timeit(lambda: ops.gemm_tc(a, w, M=rows, N=N, K=K, passes=3, **kw))
    flops = 2.0 * rows * N * K * 3
    print(f"{name:34s} {us:8.1f} us  {flops / us * 1e-6:7.1f} TFLOP/s (3-pass)  tiles {((rows + 127) // 128) * (N // ops._block_n(N))}")
