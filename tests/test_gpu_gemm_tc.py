"""tcgen05 split-bf16 GEMM (csrc/gemm_tc.cu) against float64 numpy and against the fp32 SIMT GEMM."""
import numpy as np
import pytest
import torch

from helpers import decode_planes

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def ops():
    from snuffy_b200 import ops as _ops
    return _ops


def _operands(ops, M, N, K, seed):
    rs = np.random.RandomState(seed)
    a = rs.standard_normal((M, K)).astype(np.float32)
    b = (rs.standard_normal((N, K)) / np.sqrt(K)).astype(np.float32)
    ad, bd = torch.from_numpy(a).cuda(), torch.from_numpy(b).cuda()
    _, ap, _ = ops.ln_rows(ad, None, None, apply_ln=False, want_planes=True)
    return a, b, ad, bd, ap, ops.weight_planes(bd)


@pytest.mark.parametrize("M,N,K", [(128, 128, 32), (128, 256, 64), (256, 512, 512), (300, 384, 96), (1000, 2048, 512),
                                   (1000, 512, 2048), (10000, 1024, 512), (777, 48, 48), (64, 32, 32)])
def test_gemm_tc_three_pass_matches_fp32(ops, M, N, K):
    a, b, ad, bd, ap, bp = _operands(ops, M, N, K, M + N + K)
    out, _, _ = ops.gemm_tc(ap, bp, M=M, N=N, K=K, passes=3)
    ref = a.astype(np.float64) @ b.astype(np.float64).T
    err = np.abs(out.cpu().numpy() - ref).max()
    # 3-pass split keeps ~16 mantissa bits per operand: error ~ 2^-16 * |a||b| * sqrt(K) terms, far below 1e-4
    assert err < 3e-5 * max(1.0, np.abs(ref).max()), err
    simt = ops.gemm_f32(ad, bd, M=M, N=N, K=K).cpu().numpy()
    assert np.abs(out.cpu().numpy() - simt).max() < 3e-5 * max(1.0, np.abs(ref).max())


@pytest.mark.parametrize("M,N,K", [(10000, 512, 2048), (1000, 512, 2048), (300, 256, 1536), (19000, 512, 3072), (200, 768, 3072)])
def test_gemm_tc_tail_split_long_k(ops, M, N, K):
    """K >= 1536: the tiles of the last, partial wave are cut into K slices that hand their accumulators over to the slice
    owning the epilogue (bias, residual through row_map, planes).  Against fp64, bit-identical from launch to launch, and the
    hand-over buffers re-armed for the next call."""
    a, b, ad, bd, ap, bp = _operands(ops, M, N, K, M + N + K + 7)
    rs = np.random.RandomState(5)
    bias = torch.from_numpy(rs.standard_normal(N).astype(np.float32)).cuda()
    resid = torch.from_numpy(rs.standard_normal((M, N)).astype(np.float32)).cuda()
    outs = []
    for _ in range(3):
        out, _, planes = ops.gemm_tc(ap, bp, M=M, N=N, K=K, passes=3, bias=bias, resid=resid, want_planes=True)
        outs.append(out)
    ref = a.astype(np.float64) @ b.astype(np.float64).T + bias.cpu().numpy() + resid.cpu().numpy()
    assert np.abs(outs[0].cpu().numpy() - ref).max() < 3e-5 * max(1.0, np.abs(ref).max())
    assert torch.equal(outs[0], outs[1]) and torch.equal(outs[0], outs[2])
    hi, lo = decode_planes(planes, M, N)
    want = ref - resid.cpu().numpy()                          # the planes carry the activated value, before the residual
    assert np.abs((hi + lo)[:, :N] - want).max() < 3e-5 * max(1.0, np.abs(want).max())


def test_gemm_tc_single_pass_is_bf16(ops):
    M, N, K = 256, 256, 128
    a, b, ad, bd, ap, bp = _operands(ops, M, N, K, 1)
    out, _, _ = ops.gemm_tc(ap, bp, M=M, N=N, K=K, passes=1)
    r = lambda t: torch.from_numpy(t).to(torch.bfloat16).to(torch.float64).numpy()
    ref = r(a) @ r(b).T
    assert np.abs(out.cpu().numpy() - ref).max() < 1e-4


@pytest.mark.parametrize("act", ["none", "relu", "gelu"])
def test_gemm_tc_epilogue(ops, act):
    from oracle import snuffy_oracle as so
    M, N, K, ks = 700, 256, 64, 11
    a, b, ad, bd, ap, bp = _operands(ops, M, N, K, 3)
    rs = np.random.RandomState(9)
    bias = rs.standard_normal(N).astype(np.float32)
    resid = rs.standard_normal((M, N)).astype(np.float32)
    alt = rs.standard_normal((ks, N)).astype(np.float32)
    sel = rs.permutation(M)[:ks]
    rm = -np.ones(M, np.int32)
    rm[sel] = np.arange(ks)
    t = lambda v, dt=torch.float32: torch.as_tensor(v, dtype=dt).cuda()
    out, pre, planes = ops.gemm_tc(ap, bp, M=M, N=N, K=K, passes=3, bias=t(bias), act=act, resid=t(resid),
                                   row_map=t(rm, torch.int32), resid_alt=t(alt), want_out=True, want_preact=True,
                                   want_planes=True)
    z = a.astype(np.float64) @ b.astype(np.float64).T + bias
    fn = (lambda v: v) if act == "none" else so.activation_fn(act)
    r = resid.astype(np.float64).copy()
    r[sel] = alt
    assert np.abs(pre.cpu().numpy() - z).max() < 3e-5
    assert np.abs(out.cpu().numpy() - (fn(z) + r)).max() < 3e-5
    hi, lo = decode_planes(planes, M, N)
    assert np.abs(hi + lo - fn(z)).max() < 6e-5         # activated value, split again for the next GEMM


def test_gemm_tc_chained_like_ffn(ops):
    """planes_out of GEMM 1 feed GEMM 2 as the A operand (FFN up -> down), K padding included."""
    M, d, dff = 500, 48, 192
    rs = np.random.RandomState(5)
    x = rs.standard_normal((M, d)).astype(np.float32)
    w1 = (rs.standard_normal((dff, d)) / np.sqrt(d)).astype(np.float32)
    w2 = (rs.standard_normal((d, dff)) / np.sqrt(dff)).astype(np.float32)
    t = lambda v: torch.from_numpy(v).cuda()
    _, xp, _ = ops.ln_rows(t(x), None, None, apply_ln=False, want_planes=True)
    _, _, hp = ops.gemm_tc(xp, ops.weight_planes(t(w1)), M=M, N=dff, K=d, act="relu", want_out=False, want_planes=True)
    out, _, _ = ops.gemm_tc(hp, ops.weight_planes(t(w2)), M=M, N=d, K=dff, resid=t(x))
    ref = x + np.maximum(x.astype(np.float64) @ w1.T.astype(np.float64), 0) @ w2.T.astype(np.float64)
    assert np.abs(out.cpu().numpy() - ref).max() < 5e-5


# ------------------------------------------------------------------ backward operands: transposed planes + split-K
@pytest.mark.parametrize("R,C,rc", [(1000, 96, 128), (333, 512, 256), (64, 40, 128)])
def test_planes_t_modes(ops, R, C, rc):
    import torch.nn.functional as F
    from helpers import decode_planes
    g = torch.Generator(device="cuda").manual_seed(R)
    x = torch.randn(R, C, device="cuda", generator=g)
    # mode 0: plain transpose
    hi, lo = decode_planes(ops.planes_t(x, rc), C, R)
    assert not (hi + lo)[:, R:].any()                                # k padding is zero
    hi, lo = hi[:, :R], lo[:, :R]
    ref = x.double().cpu().numpy().T
    assert np.abs(hi + lo - ref).max() < 2e-5 * max(1.0, np.abs(ref).max())
    # mode 1: LayerNorm from saved statistics, rows redirected through row_map
    gamma, beta = torch.randn(C, device="cuda", generator=g), torch.randn(C, device="cuda", generator=g)
    alt = torch.randn(5, C, device="cuda", generator=g)
    idx = torch.tensor([[3, 17, R - 1, 0, 9]], device="cuda")
    row_map = ops.build_row_map(idx, R)
    y = x.clone(); y[idx[0]] = alt
    _, _, stats = ops.ln_rows(x, gamma, beta, row_map=row_map, alt=alt, want_f32=True, want_stats=True)
    hi, lo = decode_planes(ops.planes_t(x, rc, mode=1, stats=stats, gamma=gamma, beta=beta, row_map=row_map, alt=alt), C, R)
    ref = F.layer_norm(y.double(), (C,), gamma.double(), beta.double()).cpu().numpy().T
    assert np.abs((hi + lo)[:, :R] - ref).max() < 2e-5 * max(1.0, np.abs(ref).max()) + 2e-6
    # mode 2: activation of a saved pre-activation
    hi, lo = decode_planes(ops.planes_t(x, rc, mode=2, act="gelu"), C, R)
    ref = F.gelu(x.double()).cpu().numpy().T
    assert np.abs((hi + lo)[:, :R] - ref).max() < 2e-5 * max(1.0, np.abs(ref).max())


@pytest.mark.parametrize("M,N,K", [(512, 2048, 10000), (1024, 512, 3000), (128, 256, 40), (200, 64, 777)])
def test_gemm_tc_splitk_weight_gradient(ops, M, N, K):
    """dW[M, N] = dY^T X with dY [K, M], X [K, N]: transposed planes + split-K tcgen05 GEMM vs fp64."""
    g = torch.Generator(device="cuda").manual_seed(K)
    dy = torch.randn(K, M, device="cuda", generator=g)
    x = torch.randn(K, N, device="cuda", generator=g)
    out = ops.gemm_tc_splitk(ops.planes_t(dy, 128), ops.planes_t(x, ops.lib.snuffy_gemm_tc_block_n(N)), M=M, N=N, K=K)
    ref = dy.double().t() @ x.double()
    assert float((out.double() - ref).abs().max() / ref.abs().max()) < 2e-5


@pytest.mark.parametrize("M,N,R", [(512, 2048, 10000), (2048, 512, 10000), (1024, 512, 3000), (128, 256, 40), (256, 384, 777),
                                   (768, 3072, 6001), (128, 128, 7)])
@pytest.mark.parametrize("passes", [3, 1])
def test_gemm_tc_splitk_rows_weight_gradient(ops, M, N, R, passes):
    """dW[M, N] = dY^T X from the ROW planes of dY [R, M] and X [R, N] (MN-major descriptors, no transposed copies) vs fp64
    and vs the transposed-planes form; R not a multiple of 16 leaves garbage rows in the planes on purpose."""
    g = torch.Generator(device="cuda").manual_seed(R)
    dy = torch.randn(R, M, device="cuda", generator=g)
    x = torch.randn(R, N, device="cuda", generator=g)
    assert ops.gemm_tc_splitk_rows_supported(M, N)
    junk = torch.full((64 << 20,), float("nan"), device="cuda")            # recycled by the allocator as the planes' padding
    del junk
    _, dyp, _ = ops.ln_rows(dy, None, None, apply_ln=False, want_planes=True)
    _, xp, _ = ops.ln_rows(x, None, None, apply_ln=False, want_planes=True)
    out = ops.gemm_tc_splitk_rows(dyp, xp, M=M, N=N, R=R, passes=passes)
    ref = dy.double().t() @ x.double()
    tol = 2e-5 if passes == 3 else 2e-2
    assert float((out.double() - ref).abs().max() / ref.abs().max()) < tol
    if passes == 3:
        old = ops.gemm_tc_splitk(ops.planes_t(dy, 128), ops.planes_t(x, ops.lib.snuffy_gemm_tc_block_n(N)), M=M, N=N, K=R)
        assert float((out - old).abs().max() / ref.abs().max()) < 1e-5


def test_weight_planes_t_dx_product(ops):
    g = torch.Generator(device="cuda").manual_seed(3)
    dy = torch.randn(700, 256, device="cuda", generator=g)
    w = torch.randn(256, 96, device="cuda", generator=g)                     # nn.Linear.weight [out, in]
    _, ap, _ = ops.ln_rows(dy, None, None, apply_ln=False, want_planes=True)
    out, _, _ = ops.gemm_tc(ap, ops.weight_planes_t(w), M=700, N=96, K=256, passes=3)
    ref = dy.double() @ w.double()
    assert float((out.double() - ref).abs().max() / ref.abs().max()) < 2e-5


@pytest.mark.parametrize("act", ["relu", "gelu", "leakyrelu", "selu"])
@pytest.mark.parametrize("M,N,K,p", [(1000, 2048, 512, 0.0), (333, 96, 64, 0.25), (128, 256, 32, 0.1)])
def test_gemm_tc_actgrad_epilogue(ops, act, M, N, K, p):
    """(A B^T) * act'(gate) * dropout mask in the epilogue == product, then the stand-alone activation backward kernel; the
    operand planes it emits decode to the same matrix."""
    a, b, ad, bd, ap, bp = _operands(ops, M, N, K, M + N + K + 1)
    rs = np.random.RandomState(N)
    gate = torch.from_numpy(rs.standard_normal((M, N)).astype(np.float32)).cuda()
    drop = (p, 11, 3)
    out, planes = ops.gemm_tc_actgrad(ap, bp, gate, act, M=M, N=N, K=K, drop=drop)
    prod, _, _ = ops.gemm_tc(ap, bp, M=M, N=N, K=K)
    want, _ = ops.act_bwd(gate, prod, act, drop, want_dh=True, want_a=False)
    assert torch.allclose(out, want, rtol=1e-6, atol=1e-7), (out - want).abs().max()
    hi, lo = decode_planes(planes, M, N)
    assert np.allclose((hi + lo)[:, :N], out.cpu().numpy().astype(np.float64), rtol=2e-5, atol=1e-6)
    if p > 0:
        zero = (out == 0).float().mean().item()
        assert abs(zero - p) < 0.05 or act == "relu"
    # the bias gradient from the same epilogue, with neither fp32 nor plane output
    _, _, cs = ops.gemm_tc_actgrad(ap, bp, gate, act, M=M, N=N, K=K, drop=drop, want_out=False, want_planes=False,
                                   want_colsum=True)
    ref = out.double().sum(0)
    assert float((cs.double() - ref).abs().max()) < 1e-5 * max(1.0, float(out.abs().sum(0).max()))


@pytest.mark.parametrize("M,N,K,p", [(1000, 2048, 512, 0.1), (333, 96, 64, 0.25), (10000, 2048, 512, 0.0)])
def test_gemm_tc_relugrad_reads_the_gate_off_the_forward_planes(ops, M, N, K, p):
    """ReLU backward gated by the forward's own dropout(relu(h)) planes == the product gated by the saved pre-activation and the
    re-drawn dropout mask (snuffy_gemm_tc_actgrad), planes and column sums included."""
    rs = np.random.RandomState(M + N)
    x = torch.from_numpy(rs.standard_normal((M, K)).astype(np.float32)).cuda()
    w1 = torch.from_numpy((rs.standard_normal((N, K)) / np.sqrt(K)).astype(np.float32)).cuda()
    drop = (p, 11, 3)
    _, xp, _ = ops.ln_rows(x, None, None, apply_ln=False, want_planes=True)
    _, h_pre, a_planes = ops.gemm_tc(xp, ops.weight_planes(w1), M=M, N=N, K=K, act="relu", want_out=False, want_preact=True,
                                     want_planes=True, drop=drop)
    a, b, ad, bd, ap, bp = _operands(ops, M, N, K, 99)                       # the incoming gradient product dY W2
    want, want_planes, want_cs = ops.gemm_tc_actgrad(ap, bp, h_pre, "relu", M=M, N=N, K=K, drop=drop, want_colsum=True)
    out, planes, cs = ops.gemm_tc_relugrad(ap, bp, a_planes, p, M=M, N=N, K=K, want_out=True, want_colsum=True)
    assert torch.allclose(out, want, rtol=1e-6, atol=1e-7), float((out - want).abs().max())
    assert torch.equal(planes.buf.view(torch.int16), want_planes.buf.view(torch.int16)) or \
        np.allclose(sum(decode_planes(planes, M, N)), sum(decode_planes(want_planes, M, N)), rtol=1e-6, atol=1e-7)
    assert float((cs - want_cs).abs().max()) < 1e-5 * max(1.0, float(want.abs().sum(0).max()))


@pytest.mark.parametrize("M,h,gn,gk", [(1000, 8, 200, 64), (300, 4, 24, 32), (257, 2, 40, 32), (500, 8, 64, 200), (384, 4, 16, 8),
                                       (130, 1, 96, 64)])
def test_gemm_tc_blockdiag_skips_only_zero_blocks(ops, M, h, gn, gk):
    """Against a block-diagonal B the K-windowed product equals the dense one (the skipped k-blocks only meet zeros),
    for windows that are and are not aligned to the 32-wide k-blocks, with 128- and 256-row B chunks."""
    N, K = h * gn, h * gk
    if N % 4 or K % 8:
        pytest.skip("shape not served")
    rs = np.random.RandomState(M + N + K)
    a = rs.standard_normal((M, K)).astype(np.float32)
    b = np.zeros((N, K), dtype=np.float32)
    for j in range(h):
        b[j * gn:(j + 1) * gn, j * gk:(j + 1) * gk] = rs.standard_normal((gn, gk)).astype(np.float32) / np.sqrt(gk)
    ad, bd = torch.from_numpy(a).cuda(), torch.from_numpy(b).cuda()
    _, ap, _ = ops.ln_rows(ad, None, None, apply_ln=False, want_planes=True, zero_planes=True)
    ref = a.astype(np.float64) @ b.astype(np.float64).T
    tol = 3e-5 * max(1.0, np.abs(ref).max())
    for bp in (ops.weight_planes(bd), ops.planes_t(bd.t().contiguous(), 128)):
        out = ops.gemm_tc_blockdiag(ap, 0, bp, M=M, N=N, K=K, group_n=gn, group_k=gk)
        assert np.abs(out.cpu().numpy() - ref).max() < tol
    wide = torch.empty(M, N + 8, device="cuda")                      # strided output, like dQ | dV
    ops.gemm_tc_blockdiag(ap, 0, ops.weight_planes(bd), M=M, N=N, K=K, group_n=gn, group_k=gk, out=wide, ldc=N + 8)
    assert np.abs(wide[:, :N].cpu().numpy() - ref).max() < tol


@pytest.mark.parametrize("h,dm,dn,K", [(8, 200, 64, 10000), (4, 24, 32, 700), (2, 40, 32, 257), (8, 100, 96, 1500), (1, 64, 64, 300)])
def test_gemm_tc_splitk_blockdiag_computes_every_diagonal_block(ops, h, dm, dn, K):
    """Tile skipping of the block-diagonal-output split-K product: every block with row // dm == col // dn is exact."""
    M, N = h * dm, h * dn
    rs = np.random.RandomState(M + N + K)
    a = rs.standard_normal((K, M)).astype(np.float32)          # contraction runs over rows, like dS^T Q
    b = rs.standard_normal((K, N)).astype(np.float32)
    ad, bd = torch.from_numpy(a).cuda(), torch.from_numpy(b).cuda()
    out = ops.gemm_tc_splitk_blockdiag(ops.planes_t(ad, 128), ops.planes_t(bd, 128), M=M, N=N, K=K, diag_m=dm, diag_n=dn)
    ref = a.astype(np.float64).T @ b.astype(np.float64)
    got = out.cpu().numpy()
    for j in range(h):
        blk = (slice(j * dm, (j + 1) * dm), slice(j * dn, (j + 1) * dn))
        assert np.abs(got[blk] - ref[blk]).max() < 3e-5 * max(1.0, np.abs(ref[blk]).max()), j
