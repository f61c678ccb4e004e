import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from snuffy_b200 import ops
def timeit(fn, iters=20):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters * 1e3
B, h, N, ks, d = 1, 8, 10000, 200, 512
dk = d // h
qv = torch.randn(B * N, 2 * d, device="cuda")
q = qv[:, :d]
kp = torch.randn(B * ks, d, device="cuda")
S = torch.empty(B * h * N, ks, device="cuda")
f = lambda: ops.gemm_f32_batched(q, kp, S, M=N, N=ks, K=dk, lda=2 * d, ldb=d, ldc=ks, alpha=0.125, nb_outer=B, nb_inner=h,
                                 sa=(N * 2 * d, dk), sb=(ks * d, dk), sc=(h * N * ks, N * ks), ksplit=1)
print("S batched (8 heads, K=64): us", timeit(f))
qc = q.contiguous()
f2 = lambda: ops.gemm_f32_batched(qc, kp, S, M=N, N=ks, K=dk, lda=d, ldb=d, ldc=ks, alpha=0.125, nb_outer=B, nb_inner=h,
                                  sa=(N * d, dk), sb=(ks * d, dk), sc=(h * N * ks, N * ks), ksplit=1)
print("S batched contiguous q: us", timeit(f2))
a = torch.randn(N, 64, device="cuda"); b = torch.randn(ks, 64, device="cuda")
f3 = lambda: ops.gemm_f32(a, b, M=N, N=ks, K=64)
print("single head dense [10000x64]x[200x64]^T: us", timeit(f3))
a = torch.randn(N, 512, device="cuda"); b = torch.randn(1024, 512, device="cuda")
f4 = lambda: ops.gemm_f32(a, b, M=N, N=1024, K=512)
t = timeit(f4); print("dense 10000x1024x512: us", t, "TFLOP/s", 2 * N * 1024 * 512 / t / 1e6)
Pd = torch.randn(B * h * N, ks, device="cuda"); dO = torch.randn(B * ks, d, device="cuda"); dv = torch.empty(B * N, 2 * d, device="cuda")
f5 = lambda: ops.gemm_f32_batched(Pd, dO, dv, M=N, N=dk, K=ks, lda=ks, ldb=d, ldc=2 * d, b_kc=False, nb_outer=B, nb_inner=h,
                                  sa=(h * N * ks, N * ks), sb=(ks * d, dk), sc=(N * 2 * d, dk), ksplit=1)
print("dV batched (N=64, K=200): us", timeit(f5))
