"""Times the dX product with the fused activation backward (snuffy_gemm_tc_actgrad) against its unfused pieces at the cfg2
FFN shape [10000 x 2048, K = 512]:   python tools/time_actgrad.py"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from snuffy_b200 import ops  # noqa: E402

M, N, K = 10000, 2048, 512
g = torch.Generator(device="cuda").manual_seed(0)
a = torch.randn(M, K, device="cuda", generator=g)
w = torch.randn(N, K, device="cuda", generator=g) / K ** 0.5
gate = torch.randn(M, N, device="cuda", generator=g)
_, ap, _ = ops.ln_rows(a, None, None, apply_ln=False, want_planes=True)
bp = ops.weight_planes(w)


def timed(fn, iters=20):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters * 1e3


for act in ("relu", "gelu"):
    for drop in ((0.0, 0, 0), (0.1, 7, 3)):
        print(act, "dropout", drop[0])
        print("  plain product, fp32 out                 %7.1f us" % timed(lambda: ops.gemm_tc(ap, bp, M=M, N=N, K=K)))
        print("  plain product, fp32 out + planes        %7.1f us" % timed(lambda: ops.gemm_tc(ap, bp, M=M, N=N, K=K, want_planes=True)))
        print("  fused, fp32 out only                    %7.1f us" % timed(
            lambda: ops.gemm_tc_actgrad(ap, bp, gate, act, M=M, N=N, K=K, drop=drop, want_planes=False)))
        print("  fused, planes only                      %7.1f us" % timed(
            lambda: ops.gemm_tc_actgrad(ap, bp, gate, act, M=M, N=N, K=K, drop=drop, want_out=False)))
        print("  fused, fp32 out + planes                %7.1f us" % timed(
            lambda: ops.gemm_tc_actgrad(ap, bp, gate, act, M=M, N=N, K=K, drop=drop)))

        def unfused():
            out, _, _ = ops.gemm_tc(ap, bp, M=M, N=N, K=K)
            dh, _ = ops.act_bwd(gate, out, act, drop, want_dh=True, want_a=False)
            ops.ln_rows(dh, None, None, apply_ln=False, want_planes=True)
        print("  unfused: product, act_bwd, planes       %7.1f us" % timed(unfused))
