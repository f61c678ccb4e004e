"""The stand-alone class surface: `attention`, `MultiHeadedAttention.forward`,
`PositionwiseFeedForward.forward`, `SublayerConnection.forward` (both modes) and `IClassifier.forward` called directly, vs
outputs of the unmodified reference (tests/golden/pieces.npz from oracle/make_golden.py --only pieces; reference
snuffy.py:44-54, 100-110, 160-168, 183-205, 224-225 and dsmil.py:39-50).  Tolerance 1e-4 absolute (fp32 kernels).
Gradients through the stand-alone forms are checked against fp64 autograd of the same op sequence in PyTorch."""
import os

import numpy as np
import pytest
import torch

from conftest import GOLDEN

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def z():
    return np.load(os.path.join(GOLDEN, "pieces.npz"))


def _t(a, grad=False):
    t = torch.from_numpy(np.asarray(a)).cuda()
    return t.requires_grad_(True) if grad else t


def _load(mod, z, prefix):
    sd = {k: torch.from_numpy(z[prefix + k]) for k in mod.state_dict().keys()}
    mod.load_state_dict(sd, strict=True)
    return mod.cuda().eval()


def test_attention_function(z):
    from snuffy_b200 import snuffy
    with torch.no_grad():
        o, p = snuffy.attention(_t(z["att_q"]), _t(z["att_k"]), _t(z["att_v"]))
    assert o.shape == z["att_out"].shape and p.shape == z["att_p"].shape
    assert np.abs(o.cpu().numpy() - z["att_out"]).max() < 1e-4
    assert np.abs(p.cpu().numpy() - z["att_p"]).max() < 1e-5


def test_multi_headed_attention_forward(z):
    from snuffy_b200 import snuffy, snuffy_multiclass
    for mod in (snuffy, snuffy_multiclass):
        mha = _load(mod.MultiHeadedAttention(4, 64), z, "mha_")
        with torch.no_grad():
            out, attn = mha(_t(z["mha_query"]), _t(z["mha_key"]), _t(z["mha_query"]))
        assert np.abs(out.cpu().numpy() - z["mha_out"]).max() < 1e-4
        assert np.abs(attn.cpu().numpy() - z["mha_attn"]).max() < 1e-5
        assert mha.attn is attn


@pytest.mark.parametrize("act", ["relu", "gelu", "leakyrelu", "selu"])
def test_positionwise_feed_forward(z, act):
    from snuffy_b200 import snuffy
    ffn = _load(snuffy.PositionwiseFeedForward(64, 256, act, 0.0), z, f"ffn_{act}_")
    with torch.no_grad():
        out = ffn(_t(z["ffn_x"]))
    assert np.abs(out.cpu().numpy() - z[f"ffn_{act}_out"]).max() < 1e-4


def test_sublayer_connection_both_modes(z):
    from snuffy_b200 import snuffy
    sc = _load(snuffy.SublayerConnection(64, 0.0), z, "sc_")
    ffn = _load(snuffy.PositionwiseFeedForward(64, 256, "selu", 0.0), z, "ffn_selu_")
    mha = _load(snuffy.MultiHeadedAttention(4, 64), z, "mha_")
    x = _t(z["sc_x"])
    top, rnd = _t(z["sc_top"]), _t(z["sc_rnd"])
    with torch.no_grad():
        ff = sc(x, ffn, None, None, None, 'ff')                                   # snuffy.py:157 call form
        keys = x[:, torch.cat((top, rnd))]
        a_out, a_p = sc(x, lambda u: mha(u, keys, u), None, top, rnd, 'attn')      # snuffy.py:148-150 call form
        b_out, b_p = sc(x, lambda u: mha(u, keys[:, :5], u), None, top, None, 'attn')
    assert np.abs(ff.cpu().numpy() - z["sc_ff_out"]).max() < 1e-4
    assert np.abs(a_out.cpu().numpy() - z["sc_attn_out"]).max() < 1e-4
    assert np.abs(a_p.cpu().numpy() - z["sc_attn_p"]).max() < 1e-5
    assert np.abs(b_out.cpu().numpy() - z["sc_attn_out_norand"]).max() < 1e-4
    assert np.abs(b_p.cpu().numpy() - z["sc_attn_p_norand"]).max() < 1e-5
    # the multiclass call form (snuffy_multiclass.py:103-113): batched float indices + five extra positionals
    from snuffy_b200 import snuffy_multiclass
    scm = _load(snuffy_multiclass.SublayerConnection(64, 0.0), z, "sc_")
    with torch.no_grad():
        m_out, _ = scm(x, lambda u: mha(u, keys, u), None, top.float()[None], rnd.float()[None], 1, 64, 1, 8, 'attn')
        m_ff = scm(x, ffn, None, None, None, 1, 64, 1, 8, 'ff')
    assert torch.equal(m_out, a_out) and torch.equal(m_ff, ff)
    with pytest.raises(ValueError):
        sc(x, ffn, None, None, None, 'bogus')


def _backbone():
    return torch.nn.Sequential(torch.nn.Conv2d(3, 8, 3, padding=1), torch.nn.ReLU(), torch.nn.AdaptiveAvgPool2d(2))


@pytest.mark.parametrize("which,ncls", [("snuffy", 1), ("dsmil", 3)])
def test_iclassifier_over_a_conv_backbone(z, which, ncls):
    """roi.py:177,324 / compute_feats.py:242 call form: IClassifier(backbone, feature_size, classes)(images) -> (feats, c).
    The backbone is the caller's torch module (out of scope); flatten + the Linear instance scorer run in the library."""
    import importlib
    mod = importlib.import_module(f"snuffy_b200.{which}")
    ic = _load(mod.IClassifier(_backbone(), 32, ncls), z, f"ic_{which}_")
    with torch.no_grad():
        feats, c = ic(_t(z["ic_imgs"]))
    assert feats.shape == (12, 32) and c.shape == (12, ncls)
    assert np.abs(feats.cpu().numpy() - z[f"ic_{which}_feats"]).max() < 1e-4
    assert np.abs(c.cpu().numpy() - z[f"ic_{which}_c"]).max() < 1e-4
    # differentiable end to end (the backbone through PyTorch's autograd, the scorer through the library's backward)
    ic.train()
    feats, c = ic(_t(z["ic_imgs"]))
    c.sum().backward()
    ref = z[f"ic_{which}_feats"].astype(np.float64).sum(0)
    assert np.abs(ic.fc.weight.grad.cpu().numpy() - ref[None]).max() < 1e-3
    assert ic.feature_extractor[0].weight.grad is not None


def _ref_attention(q, k, v):
    s = torch.matmul(q, k.transpose(-2, -1)) / (q.shape[-1] ** 0.5)
    p = s.softmax(-1)
    return torch.matmul(p.transpose(-2, -1), v), p


def test_gradients_through_the_stand_alone_forms(z):
    """autograd through attention / MultiHeadedAttention / PositionwiseFeedForward / SublayerConnection called directly."""
    from snuffy_b200 import snuffy
    # attention()
    q, k, v = _t(z["att_q"], True), _t(z["att_k"], True), _t(z["att_v"], True)
    w = torch.from_numpy(np.random.RandomState(0).standard_normal(z["att_out"].shape).astype(np.float32)).cuda()
    o, _ = snuffy.attention(q, k, v)
    (o * w).sum().backward()
    q64, k64, v64 = (torch.from_numpy(z[n]).double().requires_grad_(True) for n in ("att_q", "att_k", "att_v"))
    o64, _ = _ref_attention(q64, k64, v64)
    (o64 * w.cpu().double()).sum().backward()
    for got, ref in ((q, q64), (k, k64), (v, v64)):
        assert (got.grad.cpu().double() - ref.grad).abs().max() <= 2e-3 * ref.grad.abs().max()
    # SublayerConnection('ff') around the FFN, and ('attn') around MultiHeadedAttention
    sc = _load(snuffy.SublayerConnection(64, 0.0), z, "sc_").train()
    ffn = _load(snuffy.PositionwiseFeedForward(64, 256, "gelu", 0.0), z, "ffn_gelu_").train()
    mha = _load(snuffy.MultiHeadedAttention(4, 64, 0.0), z, "mha_").train()
    x = _t(z["sc_x"], True)
    top, rnd = _t(z["sc_top"]), _t(z["sc_rnd"])
    sel = torch.cat((top, rnd))
    y = sc(x, ffn, None, None, None, 'ff')
    a_out, _ = sc(y, lambda u: mha(u, y[:, sel], u), None, top, rnd, 'attn')
    a_out.square().sum().backward()

    import copy
    mods = [copy.deepcopy(m).cpu().double() for m in (sc, ffn, mha)]
    sc64, ffn64, mha64 = mods
    x64 = torch.from_numpy(z["sc_x"]).double().requires_grad_(True)
    n64 = torch.nn.LayerNorm(64).double()
    n64.load_state_dict(sc64.norm.state_dict())
    act = torch.nn.GELU()
    y64 = x64 + torch.nn.functional.linear(act(torch.nn.functional.linear(n64(x64), ffn64.w_1.weight, ffn64.w_1.bias)),
                                           ffn64.w_2.weight, ffn64.w_2.bias)
    u = n64(y64)
    lin = mha64.linears
    sel_c = sel.cpu()
    def heads(t):
        return t.view(1, -1, 4, 16).transpose(1, 2)
    o64, _ = _ref_attention(heads(lin[0](u)), heads(lin[1](y64[:, sel_c])), heads(lin[2](u)))
    out64 = y64[:, sel_c] + lin[3](o64.transpose(1, 2).reshape(1, -1, 64))
    out64.square().sum().backward()
    assert (x.grad.cpu().double() - x64.grad).abs().max() <= 2e-3 * x64.grad.abs().max()
    pairs = [(sc.norm.weight, n64.weight), (sc.norm.bias, n64.bias), (ffn.w_1.weight, ffn64.w_1.weight),
             (ffn.w_2.bias, ffn64.w_2.bias)] + [(a.weight, b.weight) for a, b in zip(mha.linears, lin)]
    for got, ref in pairs:
        assert (got.grad.cpu().double() - ref.grad).abs().max() <= 2e-3 * max(ref.grad.abs().max(), 1e-6)
