"""Randomised shape sweep of the main kernels against fp64 references (bug hunt; not part of the test suite).
usage: python tools/stress.py [seconds]"""
import os, sys, time, math, traceback
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np, torch
from snuffy_b200 import ops
budget = float(sys.argv[1]) if len(sys.argv) > 1 else 60.0
rs = np.random.RandomState(int(os.environ.get("SEED", 0)))
fails, runs = [], 0


def rel(a, b):
    return float((a - b).abs().max() / (b.abs().max() + 1e-12))


def t_gemm_tc():
    M, N, K = int(rs.randint(1, 700)), 4 * int(rs.randint(1, 300)), 8 * int(rs.randint(1, 130))
    a = torch.randn(M, K, device="cuda"); b = torch.randn(N, K, device="cuda") / math.sqrt(K)
    bias = torch.randn(N, device="cuda")
    _, ap, _ = ops.ln_rows(a, None, None, apply_ln=False, want_planes=True)
    out, _, _ = ops.gemm_tc(ap, ops.weight_planes(b), M=M, N=N, K=K, passes=3, bias=bias)
    ref = a.double() @ b.double().t() + bias.double()
    return rel(out.double(), ref) < 3e-5, (M, N, K)


def t_splitk():
    R, M, N = int(rs.randint(1, 5000)), 4 * int(rs.randint(1, 200)), 4 * int(rs.randint(1, 200))
    dy = torch.randn(R, M, device="cuda"); x = torch.randn(R, N, device="cuda")
    out = ops.gemm_tc_splitk(ops.planes_t(dy, 128), ops.planes_t(x, ops._block_n(N)), M=M, N=N, K=R)
    return rel(out.double(), dy.double().t() @ x.double()) < 3e-5, (R, M, N)


def t_attn():
    dk = int(rs.choice([32, 64, 96, 128])); h = int(rs.randint(1, 9)); d = dk * h
    B = int(rs.randint(1, 4)); n = int(rs.randint(1, 900)); ks = int(rs.randint(1, 600))
    if not ops.sparse_attn_tc_supported(B, n, ks, h, d):
        return True, None
    qv = torch.randn(B * n, 2 * d, device="cuda"); kp = torch.randn(B * ks, d, device="cuda")
    _, planes, _ = ops.ln_rows(qv, None, None, apply_ln=False, want_planes=True, zero_planes=True)
    o1, p1, _ = ops.sparse_attn_tc(planes, kp, B, n, ks, h, d, want_probs=True)
    q = qv[:, :d].double().view(B, n, h, dk).transpose(1, 2); v = qv[:, d:].double().view(B, n, h, dk).transpose(1, 2)
    k = kp.double().view(B, ks, h, dk).transpose(1, 2)
    p = (q @ k.transpose(-2, -1) / math.sqrt(dk)).softmax(-1)
    o = (p.transpose(-2, -1) @ v).transpose(1, 2).reshape(B * ks, d)
    ok = (p1.double() - p).abs().max() < 3e-5 and rel(o1.double(), o) < 1e-4
    return bool(ok), (B, n, ks, h, d)


def t_select():
    B, N, C = int(rs.randint(1, 4)), int(rs.randint(1, 3000)), int(rs.randint(1, 4))
    K = int(rs.randint(1, min(N, 1100) + 1))
    c = torch.from_numpy(rs.randint(-50, 50, size=(B, N, C)).astype(np.float32)).cuda()     # heavy ties
    flags = torch.zeros(B, N, dtype=torch.uint8, device="cuda")
    idx = ops.select_topk(c, K, flags).cpu().numpy()
    cn = c.cpu().numpy()
    for b in range(B):
        for j in range(C):
            order = np.lexsort((np.arange(N), -cn[b, :, j]))[:K]          # descending score, ties -> lower index
            if not np.array_equal(order, idx[b, j]):
                return False, (B, N, C, K)
    avail = N - int(flags.sum(1).max().item())                       # un-flagged rows in the fullest bag
    kr = int(rs.randint(0, min(avail, 1100) + 1))
    if kr:
        r = ops.select_random(flags, kr, 7, runs).cpu().numpy()
        fl = flags.cpu().numpy()
        for b in range(B):
            if len(set(r[b].tolist())) != kr or fl[b][r[b]].any() or r[b].min() < 0 or r[b].max() >= N:
                return False, ("random", B, N, kr)
    return True, None


def t_ln():
    rows, d = int(rs.randint(1, 400)), 8 * int(rs.randint(1, 200))
    x = torch.randn(rows, d, device="cuda") * 3 + 1; g = torch.randn(d, device="cuda"); b = torch.randn(d, device="cuda")
    out, _, st = ops.ln_rows(x, g, b, want_f32=True, want_stats=True)
    ref = torch.nn.functional.layer_norm(x.double(), (d,), g.double(), b.double())
    ok = (out.double() - ref).abs().max() < 2e-5
    dy = torch.randn(rows, d, device="cuda")
    dx, dg, db = ops.ln_rows_bwd(x, st, g, dy=dy)
    x64 = x.double().requires_grad_(True); g64 = g.double().requires_grad_(True)
    (torch.nn.functional.layer_norm(x64, (d,), g64, b.double()) * dy.double()).sum().backward()
    ok = ok and rel(dx.double(), x64.grad) < 1e-4 and rel(dg.double(), g64.grad) < 1e-4
    return bool(ok), (rows, d)


def t_blockdiag():
    """K-windowed product against a block-diagonal B, 128- and 256-row B chunks, aligned or not to the 32-wide k-blocks."""
    h = int(rs.randint(1, 9)); gn = 4 * int(rs.randint(1, 70)); gk = 8 * int(rs.randint(1, 30)); M = int(rs.randint(1, 600))
    N, K = h * gn, h * gk
    a = torch.randn(M, K, device="cuda"); b = torch.zeros(N, K, device="cuda")
    for j in range(h):
        b[j * gn:(j + 1) * gn, j * gk:(j + 1) * gk] = torch.randn(gn, gk, device="cuda") / math.sqrt(gk)
    _, ap, _ = ops.ln_rows(a, None, None, apply_ln=False, want_planes=True, zero_planes=True)
    bp = ops.weight_planes(b) if runs % 2 else ops.planes_t(b.t().contiguous(), 128)
    out = ops.gemm_tc_blockdiag(ap, 0, bp, M=M, N=N, K=K, group_n=gn, group_k=gk)
    return rel(out.double(), a.double() @ b.double().t()) < 3e-5, (M, h, gn, gk)


def t_splitk_blockdiag():
    h = int(rs.randint(1, 9)); dm = 4 * int(rs.randint(1, 70)); dn = 4 * int(rs.randint(1, 40)); K = int(rs.randint(1, 4000))
    M, N = h * dm, h * dn
    a = torch.randn(K, M, device="cuda"); b = torch.randn(K, N, device="cuda")
    out = ops.gemm_tc_splitk_blockdiag(ops.planes_t(a, 128), ops.planes_t(b, 128), M=M, N=N, K=K, diag_m=dm, diag_n=dn).double()
    ref = a.double().t() @ b.double()
    ok = True
    for j in range(h):
        ok = ok and rel(out[j * dm:(j + 1) * dm, j * dn:(j + 1) * dn], ref[j * dm:(j + 1) * dm, j * dn:(j + 1) * dn]) < 3e-5
    return bool(ok), (h, dm, dn, K)


def t_actgrad():
    M, N, K = int(rs.randint(1, 700)), 4 * int(rs.randint(1, 300)), 8 * int(rs.randint(1, 130))
    act = ["relu", "gelu", "leakyrelu", "selu"][runs % 4]; p = float(rs.choice([0.0, 0.2]))
    a = torch.randn(M, K, device="cuda"); b = torch.randn(N, K, device="cuda") / math.sqrt(K); gate = torch.randn(M, N, device="cuda")
    _, ap, _ = ops.ln_rows(a, None, None, apply_ln=False, want_planes=True)
    bp = ops.weight_planes(b)
    drop = (p, 5, int(rs.randint(1, 1000)))
    out, _ = ops.gemm_tc_actgrad(ap, bp, gate, act, M=M, N=N, K=K, drop=drop, want_planes=False)
    prod, _, _ = ops.gemm_tc(ap, bp, M=M, N=N, K=K)
    want, _ = ops.act_bwd(gate, prod, act, drop, want_dh=True, want_a=False) if N * M % 4 == 0 else (None, None)
    return (True if want is None else bool(torch.allclose(out, want, rtol=1e-6, atol=1e-7))), (M, N, K, act, p)


def t_attn_bwd():
    dk = int(rs.choice([16, 32, 64, 96])); h = int(rs.randint(1, 9)); d = dk * h
    B = int(rs.randint(1, 3)); n = int(rs.randint(1, 600)); ks = int(rs.randint(2, 300)); p = float(rs.choice([0.0, 0.15]))   # ks >= 2: see t_attn_bwd_fused
    if not ops.sparse_attn_bwd_tc_supported(B, n, ks, h, d):
        return True, None
    qv = torch.randn(B * n, 2 * d, device="cuda"); kp = torch.randn(B * ks, d, device="cuda"); d_o = torch.randn(B * ks, d, device="cuda")
    drop = (p, 3, 17)
    _, _, stats = ops.sparse_attn(qv[:, :d], qv[:, d:], kp, B, n, ks, h, want_probs=False, want_stats=True)
    _, qvp, _ = ops.ln_rows(qv, None, None, apply_ln=False, want_planes=True, zero_planes=True)
    dq, dv, dkp, _ = ops.sparse_attn_bwd_tc(qvp, qv, kp, d_o, stats, B, n, ks, h, d, drop)
    rq, rv, rkp, _ = ops.sparse_attn_bwd(qv[:, :d], qv[:, d:], kp, d_o, stats, B, n, ks, h, drop)
    # one key per head: P = 1 and the true dQ, dKp are exactly 0 — compare against the scale of the non-degenerate gradient
    floor = 1e-2 * float(rv.abs().max())
    close = lambda a, b: float((a.double() - b.double()).abs().max()) <= 1e-4 * max(float(b.abs().max()), floor)
    ok = close(dq, rq) and close(dv, rv) and close(dkp, rkp)
    return bool(ok), (B, n, ks, h, d, p)


def t_attn_bwd_fused():
    """Training forward (statistics + keep bits) and the one-kernel backward vs the fp32 SIMT backward of the same draw."""
    dk = int(rs.choice([32, 64, 96, 128])); h = int(rs.randint(1, 9)); d = dk * h
    # ks >= 2: with ONE key P is identically 1 and the true dQ / dKp are exactly 0; the kernel's P (from the saved statistics)
    # is 1 +- 1 ulp, which leaves rounding noise where the reference has zeros
    B = int(rs.randint(1, 4)); n = int(rs.randint(1, 900)); ks = int(rs.randint(2, 225)); p = float(rs.choice([0.0, 0.15, 0.5]))
    if not ops.sparse_attn_bwd_fused_supported(B, n, ks, h, d) or not ops.sparse_attn_tc_supported(B, n, ks, h, d):
        return True, None
    qv = torch.randn(B * n, 2 * d, device="cuda"); kp = torch.randn(B * ks, d, device="cuda"); d_o = torch.randn(B * ks, d, device="cuda")
    _, qvp, _ = ops.ln_rows(qv, None, None, apply_ln=False, want_planes=True, zero_planes=True)
    o, _, stats, mask = ops.sparse_attn_tc(qvp, kp, B, n, ks, h, d, want_probs=False, want_stats=True, dropout_p=p, seed=3,
                                           offset=runs, want_mask=True)
    o_ref, _, _ = ops.sparse_attn(qv[:, :d], qv[:, d:], kp, B, n, ks, h, want_probs=False, dropout_p=p, seed=3, offset=runs)
    dq, dv, dkp, _ = ops.sparse_attn_bwd_fused(qvp, kp, d_o, stats, B, n, ks, h, d, p, mask)
    rq, rv, rkp, _ = ops.sparse_attn_bwd(qv[:, :d], qv[:, d:], kp, d_o, stats, B, n, ks, h, (p, 3, runs))
    floor = 1e-2 * float(rv.abs().max())
    close = lambda a, b: float((a.double() - b.double()).abs().max()) <= 1e-4 * max(float(b.abs().max()), floor)
    ok = close(o, o_ref) and close(dq, rq) and close(dv, rv) and close(dkp, rkp)
    return bool(ok), (B, n, ks, h, d, p)


tests = [t_gemm_tc, t_splitk, t_attn, t_select, t_ln, t_blockdiag, t_splitk_blockdiag, t_actgrad, t_attn_bwd, t_attn_bwd_fused,
         t_attn, t_attn_bwd_fused]
if os.environ.get("ONLY"):
    tests = [t for t in tests if t.__name__ in os.environ["ONLY"].split(",")]
t0 = time.time()
while time.time() - t0 < budget:
    fn = tests[runs % len(tests)]
    runs += 1
    try:
        ok, info = fn()
        torch.cuda.synchronize()
        if not ok:
            fails.append((fn.__name__, info)); print("FAIL", fn.__name__, info, flush=True)
    except Exception as e:
        fails.append((fn.__name__, repr(e)[:200])); print("EXC", fn.__name__, repr(e)[:300], flush=True)
        traceback.print_exc()
        if "illegal" in repr(e).lower() or "launch failure" in repr(e).lower():
            break
import collections
print(f"stress: {runs} runs, {len(fails)} failures", dict(collections.Counter(f[0] for f in fails)))
