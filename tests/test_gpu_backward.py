"""Gradient parity of the drop-in modules (train.py:259 ``loss.backward()``) against autograd of the reference's op
sequence on the CPU in float64 (oracle/torch_port.py extended with explicit dropout masks), with the caller's loss
(train.py:828-846: w * BCE(bag) + (1 - w) * BCE(max instance)).

Tolerance: every parameter gradient within 2e-3 of its own max-abs (fp32 SIMT backward vs an fp64 reference; the forward
runs in both precision modes).  Kernel-level tests cover the backward primitives one by one.
"""
import math

import numpy as np
import pytest
import torch
import torch.nn.functional as F

from helpers import build_snuffy, load_golden, load_params, set_precision, snuffy_inputs, force_selections

pytestmark = pytest.mark.gpu

_ACT = {"relu": F.relu, "gelu": F.gelu, "leakyrelu": lambda t: F.leaky_relu(t, 0.01), "selu": F.selu}


def ref_forward(x, P, heads, depth, act, sels, masks=None):
    """The reference's op sequence (snuffy.py:126-157, 160-205, 224-225, 68-86) for x [B, N, d] float64 with explicit
    selections sels[l] [B, Ksel] and optional dropout keep-scale masks per layer (attn, enc1, ff, enc2)."""
    B, n, d = x.shape
    dk = d // heads
    c = F.linear(x, P["i_classifier.fc.0.weight"], P["i_classifier.fc.0.bias"])
    bidx = torch.arange(B)[:, None]
    for l in range(depth):
        pre = f"b_classifier.encoder.layers.{l}."
        lin = lambda t, k: F.linear(t, P[pre + k + ".weight"], P[pre + k + ".bias"])
        sel = sels[l]
        m = masks[l] if masks is not None else {}
        xs = x[bidx, sel]                                                  # raw rows [B, K, d]
        u = F.layer_norm(x, (d,), P[pre + "sublayer.0.norm.weight"], P[pre + "sublayer.0.norm.bias"])
        q = lin(u, "self_attn.linears.0").view(B, n, heads, dk).transpose(1, 2)
        kp = lin(xs, "self_attn.linears.1").view(B, -1, heads, dk).transpose(1, 2)
        v = lin(u, "self_attn.linears.2").view(B, n, heads, dk).transpose(1, 2)
        p = (q @ kp.transpose(-2, -1) / math.sqrt(dk)).softmax(-1)
        if "attn" in m:
            p = p * m["attn"]
        o = (p.transpose(-2, -1) @ v).transpose(1, 2).reshape(B, -1, d)
        z = lin(o, "self_attn.linears.3")
        if "enc1" in m:
            z = z * m["enc1"]
        y = x.clone()
        y[bidx, sel] = xs + z
        u2 = F.layer_norm(y, (d,), P[pre + "sublayer.1.norm.weight"], P[pre + "sublayer.1.norm.bias"])
        hdn = _ACT[act](lin(u2, "feed_forward.w_1"))
        if "ff" in m:
            hdn = hdn * m["ff"]
        f = lin(hdn, "feed_forward.w_2")
        if "enc2" in m:
            f = f * m["enc2"]
        x = y + f
    zf = F.layer_norm(x, (d,), P["b_classifier.encoder.norm.weight"], P["b_classifier.encoder.norm.bias"])
    bag = F.linear(zf.mean(1), P["b_classifier.linear.weight"], P["b_classifier.linear.bias"])
    return c, bag


def mil_loss(classes, bag, label, w=0.5):
    mx = classes.max(dim=1)[0]                                             # train.py:831-834
    return w * F.binary_cross_entropy_with_logits(bag, label) + (1 - w) * F.binary_cross_entropy_with_logits(mx, label)


def _rel(a, b):
    return float((a - b).abs().max() / (b.abs().max() + 1e-12))


def _compare_grads(model, P64, tol=2e-3):
    worst = ("", 0.0)
    for name, p in model.named_parameters():
        ref = P64[name].grad
        assert p.grad is not None, f"{name} received no gradient"
        got = p.grad.detach().double().cpu()
        if ref.abs().max() < 1e-10:                                         # key-projection bias: identically 0 (App. B-16)
            assert got.abs().max() < 1e-6, (name, got.abs().max())
            continue
        r = _rel(got, ref)
        if r > worst[1]:
            worst = (name, r)
    assert worst[1] < tol, worst


def _params64(params):
    return {k: torch.from_numpy(np.asarray(v)).double().requires_grad_(True) for k, v in params.items()}


@pytest.mark.parametrize("precision", ["fp32", "bf16x3"])
@pytest.mark.parametrize("name", ["bin_tiny_relu", "bin_rand_gelu", "bin_short_leaky", "bin_k201_selu", "bin_cfg1"])
def test_binary_gradients_match_reference_autograd(name, precision):
    from snuffy_b200 import snuffy
    z, c = load_golden(name)
    params, x = snuffy_inputs(c)
    model = load_params(build_snuffy(snuffy, c), params)
    set_precision(model, precision)
    sels = z["ref32_sel"]
    force_selections(model, sels)
    label = torch.ones(1, 1)
    classes, bag, _ = model(torch.from_numpy(x).cuda())
    loss = mil_loss(classes, bag, label.cuda())
    loss.backward()

    P64 = _params64(params)
    sel_t = [torch.as_tensor(np.asarray(s), dtype=torch.int64).view(1, -1) for s in sels]
    c64, bag64 = ref_forward(torch.from_numpy(x).double(), P64, c["heads"], c["depth"], c["act"], sel_t)
    loss64 = mil_loss(c64, bag64, label.double())
    loss64.backward()
    assert abs(float(loss) - float(loss64)) < 1e-4
    _compare_grads(model, P64)


def test_batched_bags_and_input_gradient():
    """forward_bags on B = 3 bags, depth 2: parameter grads sum over bags; the bag tensor itself gets a gradient."""
    from snuffy_b200 import snuffy
    z, c = load_golden("bin_rand_gelu")
    params, _ = snuffy_inputs(c)
    model = load_params(build_snuffy(snuffy, c), params)
    set_precision(model, "fp32")
    rs = np.random.RandomState(3)
    B, n, d = 3, c["n"], c["d"]
    x = rs.standard_normal((B, n, d)).astype(np.float32)
    ksel = len(z["ref32_sel"][0])
    sels = [np.stack([rs.permutation(n)[:ksel] for _ in range(B)]) for _ in range(c["depth"])]
    force_selections(model, sels)
    xg = torch.from_numpy(x).cuda().requires_grad_(True)
    classes, bag, _ = snuffy.forward_bags(model, xg)
    label = torch.tensor([[1.0], [0.0], [1.0]])
    loss = mil_loss(classes, bag, label.cuda())
    loss.backward()

    P64 = _params64(params)
    x64 = torch.from_numpy(x).double().requires_grad_(True)
    c64, bag64 = ref_forward(x64, P64, c["heads"], c["depth"], c["act"], [torch.from_numpy(s) for s in sels])
    mil_loss(c64, bag64, label.double()).backward()
    _compare_grads(model, P64)
    assert _rel(xg.grad.double().cpu(), x64.grad) < 2e-3


def test_train_mode_dropout_gradients():
    """Train mode with the reference's dropouts (attention 0.1 default, encoder, FFN): masks are regenerated from the
    (seed, offset) counters, read back through the same kernels, and fed to the float64 reference."""
    from snuffy_b200 import engine, ops, snuffy
    z, c = load_golden("bin_tiny_relu")
    params, x = snuffy_inputs(c)
    model = load_params(build_snuffy(snuffy, c, ff_dropout=0.2, enc_dropout=0.15), params)
    model.train()
    set_precision(model, "fp32")
    sels = z["ref32_sel"]
    force_selections(model, sels)
    torch.manual_seed(77)
    seed = torch.initial_seed()
    off0 = engine._RANDOM._offset if engine._RANDOM._seed == seed else 0
    n, d, h = c["n"], c["d"], c["heads"]
    ksel = len(sels[0])
    classes, bag, _ = model(torch.from_numpy(x).cuda())
    label = torch.zeros(1, 1)
    mil_loss(classes, bag, label.cuda()).backward()

    def keep(shape, p, offset):                     # keep-scale mask of a flat-indexed tensor
        ones = torch.ones(int(np.prod(shape)), device="cuda")
        return ops.act_bwd(None, ones, drop=(p, seed, offset))[0].view(*shape).double().cpu()

    def keep_attn(p, offset):                       # [1, h, n, ksel], index = ((b*h + j)*N + n)*Ksel + key
        s = torch.zeros(h * n, ksel, device="cuda")
        st = torch.tensor([0.0, 1.0], device="cuda").repeat(h * n, 1).contiguous()
        pd = torch.empty_like(s)
        ops.attn_rows_bwd(s, st, n, 0, 1.0, (p, seed, offset), pd=pd)
        return pd.view(1, h, n, ksel).double().cpu()

    masks = [{"attn": keep_attn(0.1, off0 + 1), "enc1": keep((1, ksel, d), 0.15, off0 + 2),
              "ff": keep((1, n, 4 * d), 0.2, off0 + 3), "enc2": keep((1, n, d), 0.15, off0 + 4)}]
    for m in masks[0].values():
        assert 0.0 < float((m == 0).double().mean()) < 0.5
    P64 = _params64(params)
    sel_t = [torch.as_tensor(np.asarray(s), dtype=torch.int64).view(1, -1) for s in sels]
    c64, bag64 = ref_forward(torch.from_numpy(x).double(), P64, h, c["depth"], c["act"], sel_t, masks)
    assert (bag.detach().double().cpu() - bag64.detach()).abs().max() < 1e-4
    mil_loss(c64, bag64, label.double()).backward()
    _compare_grads(model, P64)


def test_multiclass_gradients():
    from snuffy_b200 import snuffy_multiclass
    z, c = load_golden("mc_b1_c2")
    params, x = snuffy_inputs(c)
    model = load_params(build_snuffy(snuffy_multiclass, c, multiclass=True), params)
    set_precision(model, "fp32")
    sels = z["ref32_sel"]
    force_selections(model, sels)
    classes, bag, _ = model(torch.from_numpy(x).cuda())
    label = torch.zeros(x.shape[0], c["C"])
    label[:, 0] = 1
    mil_loss(classes, bag, label.cuda()).backward()
    P64 = _params64(params)
    sel_t = [torch.as_tensor(np.asarray(s), dtype=torch.int64).view(x.shape[0], -1) for s in sels]
    c64, bag64 = ref_forward(torch.from_numpy(x).double(), P64, c["heads"], c["depth"], c["act"], sel_t)
    mil_loss(c64, bag64, label.double()).backward()
    _compare_grads(model, P64)


# ------------------------------------------------------------------ backward primitives
@pytest.fixture(scope="module")
def ops():
    from snuffy_b200 import ops as _ops
    return _ops


@pytest.mark.parametrize("R,M,N", [(1000, 48, 40), (10000, 512, 128), (7, 5, 3), (4100, 200, 64)])
def test_matmul_tn_split_k(ops, R, M, N):
    g = torch.Generator(device="cuda").manual_seed(R)
    a = torch.randn(R, M, device="cuda", generator=g)
    b = torch.randn(R, N, device="cuda", generator=g)
    ref = a.double().t() @ b.double()
    assert _rel(ops.matmul_tn(a, b).double(), ref) < 1e-5


def test_matmul_nn_nt(ops):
    g = torch.Generator(device="cuda").manual_seed(5)
    a = torch.randn(300, 70, device="cuda", generator=g)
    w = torch.randn(70, 90, device="cuda", generator=g)
    r = torch.randn(300, 90, device="cuda", generator=g)
    assert _rel(ops.matmul_nn(a, w, resid=r).double(), a.double() @ w.double() + r.double()) < 1e-5
    assert _rel(ops.matmul_nt(a, w.t().contiguous()).double(), a.double() @ w.double()) < 1e-5


def test_gemm_batched_head_slices(ops):
    B, h, n, k, dk = 2, 3, 50, 17, 8
    d = h * dk
    g = torch.Generator(device="cuda").manual_seed(9)
    q = torch.randn(B * n, 2 * d, device="cuda", generator=g)[:, :d]          # row-strided view like Q of Q|V
    kp = torch.randn(B * k, d, device="cuda", generator=g)
    s = torch.empty(B * h * n, k, device="cuda")
    ops.gemm_f32_batched(q, kp, s, M=n, N=k, K=dk, lda=q.stride(0), ldb=d, ldc=k, alpha=0.5, nb_outer=B, nb_inner=h,
                         sa=(n * q.stride(0), dk), sb=(k * d, dk), sc=(h * n * k, n * k), ksplit=1)
    ref = 0.5 * torch.einsum("bnhc,bkhc->bhnk", q.reshape(B, n, h, dk).double(), kp.view(B, k, h, dk).double())
    assert _rel(s.view(B, h, n, k).double(), ref) < 1e-5


@pytest.mark.parametrize("rows,d", [(100, 32), (1000, 512), (33, 768)])
def test_ln_rows_bwd(ops, rows, d):
    g = torch.Generator(device="cuda").manual_seed(rows)
    x = torch.randn(rows, d, device="cuda", generator=g) * 3 + 1
    gamma = torch.randn(d, device="cuda", generator=g)
    beta = torch.randn(d, device="cuda", generator=g)
    dy = torch.randn(rows, d, device="cuda", generator=g)
    add = torch.randn(rows, d, device="cuda", generator=g)
    _, _, stats = ops.ln_rows(x, gamma, beta, want_f32=True, want_stats=True)
    dx, dg, db = ops.ln_rows_bwd(x, stats, gamma, dy=dy, add=add)
    x64 = x.double().requires_grad_(True)
    g64, b64 = gamma.double().requires_grad_(True), beta.double().requires_grad_(True)
    (F.layer_norm(x64, (d,), g64, b64) * dy.double()).sum().backward()
    assert _rel(dx.double(), x64.grad + add.double()) < 1e-4
    assert _rel(dg.double(), g64.grad) < 1e-4 and _rel(db.double(), b64.grad) < 1e-4
    # broadcast upstream (mean-pool): every row of bag b receives dyb[b] / N
    per = rows // 2 if rows % 2 == 0 else rows
    nb = rows // per
    dyb = torch.randn(nb, d, device="cuda", generator=g)
    dx2, dg2, _ = ops.ln_rows_bwd(x, stats, gamma, dy_bcast=dyb, rows_per_bag=per, bscale=1.0 / per)
    x64.grad = None; g64.grad = None
    (F.layer_norm(x64, (d,), g64, b64).view(nb, per, d).mean(1) * dyb.double()).sum().backward()
    assert _rel(dx2.double(), x64.grad) < 1e-4 and _rel(dg2.double(), g64.grad) < 1e-4


@pytest.mark.parametrize("act", ["relu", "gelu", "leakyrelu", "selu", "tanh", "none"])
def test_act_bwd(ops, act):
    g = torch.Generator(device="cuda").manual_seed(1)
    h = torch.randn(64, 128, device="cuda", generator=g)
    da = torch.randn(64, 128, device="cuda", generator=g)
    dh, a = ops.act_bwd(h, da, act, want_dh=True, want_a=True)
    fn = dict(_ACT, tanh=torch.tanh, none=lambda t: t)[act]
    h64 = h.double().requires_grad_(True)
    out = fn(h64)
    (out * da.double()).sum().backward()
    assert _rel(a.double(), out.detach()) < 1e-5 and _rel(dh.double(), h64.grad) < 1e-5


def test_colsum_and_scatter_add(ops):
    g = torch.Generator(device="cuda").manual_seed(2)
    x = torch.randn(5000, 96, device="cuda", generator=g)
    w = torch.randn(5000, 3, device="cuda", generator=g)
    assert _rel(ops.colsum(x, w).double(), w.double().t() @ x.double()) < 1e-5
    assert _rel(ops.colsum(x).double(), x.double().sum(0, keepdim=True)) < 1e-5
    assert _rel(ops.colsum(w[:, :1].contiguous()).double(), w[:, :1].double().sum(0, keepdim=True)) < 1e-5
    dx = torch.zeros(2, 40, 8, device="cuda")
    idx = torch.tensor([[3, 7, 39], [0, 1, 2]], device="cuda")
    src = torch.randn(6, 8, device="cuda", generator=g)
    ops.scatter_add_rows(dx, idx, src)
    ref = torch.zeros(2, 40, 8, device="cuda")
    ref[torch.arange(2)[:, None], idx] += src.view(2, 3, 8)
    assert torch.equal(dx, ref)


@pytest.mark.parametrize("B,n,ks,h,d", [(1, 300, 40, 2, 32), (2, 130, 200, 8, 64), (1, 257, 33, 1, 48)])
def test_sparse_attention_backward(ops, B, n, ks, h, d):
    g = torch.Generator(device="cuda").manual_seed(n)
    qv = torch.randn(B * n, 2 * d, device="cuda", generator=g)
    kp = torch.randn(B * ks, d, device="cuda", generator=g)
    d_o = torch.randn(B * ks, d, device="cuda", generator=g)
    q, v = qv[:, :d], qv[:, d:]
    _, _, stats = ops.sparse_attn(q, v, kp, B, n, ks, h, want_probs=False, want_stats=True)
    dq, dv, dkp, _ = ops.sparse_attn_bwd(q, v, kp, d_o, stats, B, n, ks, h)
    dk = d // h
    q64 = q.double().clone().requires_grad_(True)
    v64 = v.double().clone().requires_grad_(True)
    k64 = kp.double().requires_grad_(True)
    qh = q64.view(B, n, h, dk).transpose(1, 2)
    vh = v64.view(B, n, h, dk).transpose(1, 2)
    kh = k64.view(B, ks, h, dk).transpose(1, 2)
    p = (qh @ kh.transpose(-2, -1) / math.sqrt(dk)).softmax(-1)
    o = (p.transpose(-2, -1) @ vh).transpose(1, 2).reshape(B * ks, d)
    (o * d_o.double()).sum().backward()
    assert _rel(dq.double(), q64.grad) < 1e-4
    assert _rel(dv.double(), v64.grad) < 1e-4
    assert _rel(dkp.double(), k64.grad) < 1e-4


@pytest.mark.parametrize("name", ["ds_c1", "ds_c3_linear", "ds_c2_v"])
def test_dsmil_gradients(name):
    """dsmil.MILNet (dsmil.py:72-106) under autograd: logits + instance max loss, all parameters."""
    from helpers import build_dsmil
    from snuffy_b200 import dsmil
    z, c = load_golden(name)
    model, params, x = build_dsmil(dsmil, c)
    model = model.cuda().eval()
    xg = torch.from_numpy(x).cuda().requires_grad_(True)
    classes, bag, attn = model(xg)
    label = torch.zeros(1, c["C"]); label[0, 0] = 1
    loss = (0.5 * F.binary_cross_entropy_with_logits(bag, label.cuda())
            + 0.5 * F.binary_cross_entropy_with_logits(classes.max(0, keepdim=True)[0], label.cuda()))
    loss.backward()

    P = _params64(params)
    x64 = torch.from_numpy(x).double().requires_grad_(True)
    lin = lambda t, k: F.linear(t, P[k + ".weight"], P[k + ".bias"])
    c64 = lin(x64, "i_classifier.fc.0")
    q = (lambda t: torch.tanh(lin(F.relu(lin(t, "b_classifier.q.0")), "b_classifier.q.2"))) if c["nonlinear"] \
        else (lambda t: lin(t, "b_classifier.q"))
    v = F.relu(lin(x64, "b_classifier.v.1")) if c["passing_v"] else x64
    Q = q(x64)
    crit = c64.argmax(0)
    qm = q(x64[crit])
    A = (Q @ qm.t() / torch.sqrt(torch.tensor(128.0, dtype=torch.float32)).double()).softmax(0)
    Bm = A.t() @ v
    bag64 = F.conv1d(Bm.view(1, c["C"], -1), P["b_classifier.fcc.weight"], P["b_classifier.fcc.bias"]).view(1, -1)
    loss64 = (0.5 * F.binary_cross_entropy_with_logits(bag64, label.double())
              + 0.5 * F.binary_cross_entropy_with_logits(c64.max(0, keepdim=True)[0], label.double()))
    loss64.backward()
    assert abs(float(loss) - float(loss64)) < 1e-4
    _compare_grads(model, P)
    assert _rel(xg.grad.double().cpu(), x64.grad) < 2e-3


def test_tensor_core_backward_matches_fp32_backward_at_scale():
    """cfg2-shaped layer (d = 512, 8 heads, K = 200) on 3000 patches, train mode with dropout: the tcgen05 backward
    (transposed split-bf16 planes + split-K) against the exact-fp32 SIMT backward of the same forward."""
    from snuffy_b200 import engine, snuffy
    c = dict(n=3000, d=512, heads=8, K=200, r=0.5, depth=1, act="relu", wseed=0, xseed=7)
    params, x = snuffy_inputs(c)
    grads = {}
    for precision in ("fp32", "bf16x3"):
        model = load_params(build_snuffy(snuffy, c, ff_dropout=0.1, enc_dropout=0.1), params)
        model.train()
        set_precision(model, precision)
        rs = np.random.RandomState(1)
        force_selections(model, [rs.permutation(c["n"])[:200][None]])
        torch.manual_seed(5)
        engine._RANDOM._seed = None                  # same (seed, offset) counters -> same dropout masks in both runs
        classes, bag, _ = model(torch.from_numpy(x).cuda())
        mil_loss(classes, bag, torch.ones(1, 1).cuda()).backward()
        grads[precision] = {k: p.grad.detach().double() for k, p in model.named_parameters()}
    for k, ref in grads["fp32"].items():
        if ref.abs().max() < 1e-9:
            continue
        # the two runs also differ in the FORWARD precision (split-bf16 vs fp32), which the softmax Jacobian amplifies
        # for the tiny query/key-projection gradients (measured 2.3e-3 on linears.0.weight, <= 4e-4 elsewhere)
        assert _rel(grads["bf16x3"][k], ref) < 5e-3, (k, _rel(grads["bf16x3"][k], ref))


@pytest.mark.parametrize("r", [0.0, 0.5])
def test_second_stream_and_row_plane_gradients_change_nothing(r):
    """The off-chain work on the second stream (selection, key path, weight / bias gradients), the row-plane weight gradients and
    the ReLU gate from planes against the single-stream, transposed-copy, saved-pre-activation forms: same draws, same loss,
    gradients equal up to the summation order of the weight-gradient products, over several repetitions (a missing stream
    dependency would show as run-to-run differences)."""
    from snuffy_b200 import backward, engine, snuffy
    c = dict(n=3000, d=512, heads=8, K=200, r=r, depth=2, act="relu", wseed=0, xseed=7)
    params, x = snuffy_inputs(c)
    xs = torch.from_numpy(x).cuda()
    flags = [(engine, "FWD_SIDE_STREAM"), (backward, "DW_SIDE_STREAM"), (engine, "DW_BY_ROWS")]
    saved = [getattr(m, k) for m, k in flags]
    runs = {}
    try:
        for mode in (True, False, True, True):
            for m, k in flags:
                setattr(m, k, mode)
            model = load_params(build_snuffy(snuffy, c, ff_dropout=0.1, enc_dropout=0.1), params)
            model.train()
            set_precision(model, "bf16x3")
            torch.manual_seed(5)
            engine._RANDOM._seed = None              # same (seed, offset) counters -> same masks and random patches
            classes, bag, _ = model(xs)
            loss = mil_loss(classes, bag, torch.ones(1, 1).cuda())
            loss.backward()
            torch.cuda.synchronize()
            g = {k: p.grad.detach().clone() for k, p in model.named_parameters()}
            runs.setdefault(mode, []).append((float(loss), g))
    finally:
        for (m, k), v in zip(flags, saved):
            setattr(m, k, v)
    on, off = runs[True], runs[False][0]
    for loss, g in on[1:]:                           # the two-stream form is deterministic from run to run
        assert loss == on[0][0]
        for k in g:
            assert torch.equal(g[k], on[0][1][k]), k
    assert abs(on[0][0] - off[0]) < 1e-6
    for k, ref in off[1].items():
        if ref.abs().max() < 1e-9:
            continue
        assert _rel(on[0][1][k].double(), ref.double()) < 2e-5, (k, _rel(on[0][1][k].double(), ref.double()))


@pytest.mark.parametrize("B,n,ks,h,d,p", [(1, 300, 40, 2, 64, 0.0), (1, 1000, 200, 8, 512, 0.0), (2, 256, 24, 4, 128, 0.0),
                                          (1, 777, 100, 8, 768, 0.2), (3, 200, 40, 4, 128, 0.1), (1, 500, 256, 4, 256, 0.1),
                                          (2, 384, 8, 4, 64, 0.3)])
def test_sparse_attention_backward_tensor_core(ops, B, n, ks, h, d, p):
    """All heads of a bag through dense tcgen05 GEMMs against head-block operands == the head-batched SIMT backward
    (and fp64 autograd when there is no dropout)."""
    assert ops.sparse_attn_bwd_tc_supported(B, n, ks, h, d)
    g = torch.Generator(device="cuda").manual_seed(n)
    qv = torch.randn(B * n, 2 * d, device="cuda", generator=g)
    kp = torch.randn(B * ks, d, device="cuda", generator=g)
    d_o = torch.randn(B * ks, d, device="cuda", generator=g)
    q, v = qv[:, :d], qv[:, d:]
    drop = (p, 5, 9)
    _, _, stats = ops.sparse_attn(q, v, kp, B, n, ks, h, want_probs=False, want_stats=True)
    _, qvp, _ = ops.ln_rows(qv, None, None, apply_ln=False, want_planes=True, zero_planes=True)
    dq, dv, dkp, dqv = ops.sparse_attn_bwd_tc(qvp, qv, kp, d_o, stats, B, n, ks, h, d, drop)
    rq, rv, rkp, _ = ops.sparse_attn_bwd(q, v, kp, d_o, stats, B, n, ks, h, drop)
    assert _rel(dq.double(), rq.double()) < 1e-4 and _rel(dv.double(), rv.double()) < 1e-4
    assert _rel(dkp.double(), rkp.double()) < 1e-4
    assert torch.equal(dqv[:, :d], dq) and torch.equal(dqv[:, d:], dv)


@pytest.mark.parametrize("B,n,ks,h,d,p", [(1, 1000, 200, 8, 512, 0.0), (2, 300, 64, 2, 64, 0.0), (1, 1000, 200, 8, 512, 0.1),
                                          (3, 200, 40, 4, 128, 0.1), (1, 500, 208, 4, 256, 0.2), (2, 384, 8, 4, 128, 0.3),
                                          (1, 130, 100, 2, 192, 0.0), (2, 1000, 50, 1, 128, 0.0)])
def test_sparse_attention_backward_fused(ops, B, n, ks, h, d, p):
    """The one-kernel tcgen05 attention backward (csrc/attn_bwd_tc.cu) == the fp32 SIMT backward of the same forward (same
    saved statistics, same dropout draw): dQ, dV, dKp within 1e-4 of each tensor's max-abs.  Shapes cover bags that start
    inside a 128-row tile, head sizes 32 / 64 / 96 / 128 (stacked and plain operands) and padded key counts."""
    assert ops.sparse_attn_bwd_fused_supported(B, n, ks, h, d)
    g = torch.Generator(device="cuda").manual_seed(n + ks)
    qv = torch.randn(B * n, 2 * d, device="cuda", generator=g)
    kp = torch.randn(B * ks, d, device="cuda", generator=g)
    d_o = torch.randn(B * ks, d, device="cuda", generator=g)
    q, v = qv[:, :d], qv[:, d:]
    drop = (p, 5, 9)
    _, qvp, _ = ops.ln_rows(qv, None, None, apply_ln=False, want_planes=True, zero_planes=True)
    # the tensor-core forward leaves the statistics and, with dropout, the keep bits of its draw for the backward
    _, _, stats, mask = ops.sparse_attn_tc(qvp, kp, B, n, ks, h, d, want_probs=False, want_stats=True, dropout_p=p, seed=5,
                                           offset=9, want_mask=True)
    assert (mask is None) == (p == 0)
    dq, dv, dkp, dqv = ops.sparse_attn_bwd_fused(qvp, kp, d_o, stats, B, n, ks, h, d, p, mask)
    rq, rv, rkp, _ = ops.sparse_attn_bwd(q, v, kp, d_o, stats, B, n, ks, h, drop)
    torch.cuda.synchronize()
    assert _rel(dv.double(), rv.double()) < 1e-4, _rel(dv.double(), rv.double())
    assert _rel(dq.double(), rq.double()) < 1e-4, _rel(dq.double(), rq.double())
    assert _rel(dkp.double(), rkp.double()) < 1e-4, _rel(dkp.double(), rkp.double())
    assert not ops.sparse_attn_bwd_fused_supported(1, 1000, 256, 8, 512)          # Ksel > 224: the GEMM formulation serves it


@pytest.mark.gpu
@pytest.mark.parametrize("d,dff", [(512, 2048), (256, 520), (768, 3072), (96, 200)])
def test_weight_planes_batch_matches_single_conversions(ops, d, dff):
    """One launch for every derived operand of a training step = the per-operand conversions, byte for byte
    (padding rows / k included), and LayerWeights.prepare_train = prepare + prepare_backward."""
    from snuffy_b200 import engine
    g = torch.Generator(device="cuda").manual_seed(d)
    mk = lambda *s: torch.randn(*s, device="cuda", generator=g)
    ps = [mk(d, d), mk(d), mk(d, d), mk(d), mk(d, d), mk(d), mk(d, d), mk(d), mk(dff, d), mk(dff), mk(d, dff), mk(d),
          mk(d), mk(d), mk(d), mk(d)]
    a, b = engine.LayerWeights(*ps), engine.LayerWeights(*ps)
    a.prepare("bf16x3"); a.prepare_backward()
    served = b.prepare_train()
    assert served == (d % ops._block_n(2 * d) == 0 and d % 32 == 0)
    if not served:
        return
    torch.cuda.synchronize()
    for name in ("wqv_planes", "w1_planes", "w2_planes", "wk_planes", "wo_planes", "wqvt_planes", "w1t_planes", "w2t_planes",
                 "wkt_planes", "wot_planes"):
        pa, pb = getattr(a, name), getattr(b, name)
        assert (pa.rows, pa.K, pa.rc, pa.stride) == (pb.rows, pb.K, pb.rc, pb.stride), name
        assert torch.equal(pa.buf.view(torch.int16), pb.buf.view(torch.int16)), name
    assert torch.equal(a.bqv, b.bqv)
