"""The oracle restatement vs the reference's own outputs (tests/golden, made by
oracle/make_golden.py from the unmodified reference).  CPU only."""
import json
import os

import numpy as np
import pytest

from oracle import snuffy_oracle as so
from oracle.params import make_bag, make_dsmil_params, make_snuffy_params
from conftest import GOLDEN

BIN = ["bin_tiny_relu", "bin_rand_gelu", "bin_short_leaky", "bin_k201_selu", "bin_cfg1", "bin_cfg2", "bin_cfg2_rand",
       "bin_cfg4_small"]
MC = ["mc_b1_c2", "mc_b3_c3", "mc_c1_r0", "mc_cfg3s", "mc_cfg3"]
DS = ["ds_c1", "ds_c3_linear", "ds_c2_v", "ds_cfg2"]


def load(name):
    z = np.load(os.path.join(GOLDEN, name + ".npz"))
    return z, json.loads(str(z["config"]))


def cfg_of(c):
    return so.SnuffyConfig(d=c["d"], heads=c["heads"], big_lambda=c["K"], random_patch_share=c["r"],
                           depth=c["depth"], activation=c["act"], num_classes=c.get("C", 1))


def inputs(c):
    params = make_snuffy_params(c["d"], c["depth"], c.get("C", 1), 4, c["wseed"], realistic=c.get("realistic", False))
    x = make_bag(c["n"], c["d"], c["xseed"], c.get("B", 1))
    return params, x


@pytest.mark.parametrize("name", BIN)
def test_binary_fp64_matches_reference(name):
    z, c = load(name)
    params, x = inputs(c)
    assert abs(float(z["x_checksum"]) - x.astype(np.float64).sum()) < 1e-9
    if "x" in z:
        assert np.array_equal(z["x"], x)
    cfg = cfg_of(c)
    # selection parity: T from the scores with the documented tie rule, R replayed from NumPy's global RNG
    np.random.seed(c["npseed"])
    out = so.snuffy_forward(x, params, cfg, sampler=so.numpy_global_rng_sampler, keep_layers=True)
    assert np.array_equal(np.stack(out["selections"]), z["ref64_sel"])
    assert np.abs(out["classes"] - z["ref64_classes"]).max() < 1e-12
    assert np.abs(out["bag"] - z["ref64_bag"]).max() < 1e-11
    rows = z["sub_rows"]
    for l, xl in enumerate(out["layers"]):
        ref = z["ref64_layers_rows"][l][0]
        assert np.abs(xl[rows] - ref).max() <= 1e-9 * max(1.0, np.abs(ref).max())
        assert abs(xl.sum() - z["ref64_layers_sum"][l]) <= 1e-8 * max(1.0, abs(z["ref64_layers_sum"][l]))
    if "ref64_last_sel_rows" in z:                      # `lean` fixtures keep the large arrays of the fp32 run only
        last = out["layers"][-1][out["selections"][-1]]
        assert np.abs(last - z["ref64_last_sel_rows"]).max() <= 1e-9 * max(1.0, np.abs(last).max())
    if "ref64_attn" in z:
        assert np.abs(out["attn"] - z["ref64_attn"]).max() < 1e-9
    elif "ref64_attn_rows" in z:
        assert np.abs(out["attn"][..., rows, :] - z["ref64_attn_rows"]).max() < 1e-9
    # |S| follows python-double rounding (K=200, r=0.7 -> 61+140 = 201)
    assert out["selections"][0].shape[0] == min(c["n"], cfg.k_top) + cfg.k_rand(c["n"])


@pytest.mark.parametrize("name", BIN)
def test_binary_fp32_oracle_within_bar(name):
    """An fp32 evaluation of the restatement sits within the 1e-4 bar of the fp32 reference."""
    z, c = load(name)
    params, x = inputs(c)
    out = so.snuffy_forward(x, params, cfg_of(c), selections=list(z["ref32_sel"]), dtype=np.float32)
    assert np.abs(out["classes"] - z["ref32_classes"]).max() < 1e-4
    assert np.abs(out["bag"] - z["ref32_bag"]).max() < 1e-4


@pytest.mark.parametrize("name", MC)
def test_multiclass_fp64_matches_reference(name):
    z, c = load(name)
    params, x = inputs(c)
    np.random.seed(c["npseed"])
    out = so.snuffy_multiclass_forward(x, params, cfg_of(c), sampler=so.numpy_global_rng_sampler)
    assert np.array_equal(np.stack(out["selections"]), z["ref64_sel"])
    assert np.abs(out["classes"] - z["ref64_classes"]).max() < 1e-12
    assert np.abs(out["bag"] - z["ref64_bag"]).max() < 1e-10
    if "ref64_attn" in z:
        assert np.abs(out["attn"] - z["ref64_attn"]).max() < 1e-8


@pytest.mark.parametrize("name", DS)
def test_dsmil_fp64_matches_reference(name):
    z, c = load(name)
    params = make_dsmil_params(c["d"], c["C"], c["nonlinear"], c["passing_v"], c["wseed"])
    x = make_bag(c["n"], c["d"], c["xseed"], 1)[0]
    out = so.dsmil_forward(x, params, c["nonlinear"], c["passing_v"])
    assert np.abs(out["classes"] - z["ref64_classes"]).max() < 1e-12
    assert np.abs(out["bag"] - z["ref64_bag"]).max() < 1e-9
    assert np.abs(out["B"] - z["ref64_B"]).max() < 1e-9
    if z["ref64_attn"].shape == out["attn"].shape:
        assert np.abs(out["attn"] - z["ref64_attn"]).max() < 1e-9


def _torch_params(params):
    import torch
    return {k: torch.from_numpy(v) for k, v in params.items()}


def test_port_reproduces_cfg4_big():
    """N = 50 000, K = 1024, r = 0.5: the CPU port replays the reference's fp32 run (same ops, same NumPy stream)."""
    import torch
    from oracle import torch_port
    z, c = load("bin_cfg4_big")
    params, x = inputs(c)
    assert abs(float(z["x_checksum"]) - x.astype(np.float64).sum()) < 1e-9
    np.random.seed(c["npseed"])
    with torch.no_grad():
        cls, bag, attn = torch_port.forward(torch.from_numpy(x), _torch_params(params), c["heads"], c["K"], c["r"], c["depth"],
                                            c["act"])
    assert np.abs(cls.numpy() - z["ref32_classes"]).max() < 1e-6
    assert np.abs(bag.numpy() - z["ref32_bag"]).max() < 1e-6
    assert np.abs(attn.numpy()[..., z["sub_rows"], :] - z["ref32_attn_rows"]).max() < 1e-6
    assert z["ref32_sel"].shape == (1, 1024) and len(set(z["ref32_sel"][0].tolist())) == 1024


def test_port_reproduces_bench_batch_and_packed_bags():
    """bin_cfg2_b16 (the 16 bags of the bench step) and the variable-length fixtures: one port forward per bag."""
    import torch
    from oracle import torch_port
    z, c = load("bin_cfg2_b16")
    tp = _torch_params(make_snuffy_params(c["d"], c["depth"], 1, 4, c["wseed"], realistic=True))
    for b in (0, 7, 15):
        x = make_bag(c["n"], c["d"], c["xseed"] + b, 1)
        assert abs(float(z["x_checksum"][b]) - x.astype(np.float64).sum()) < 1e-9
        with torch.no_grad():
            cls, bag, _ = torch_port.forward(torch.from_numpy(x), tp, c["heads"], c["K"], c["r"], c["depth"], c["act"],
                                             selections=list(z["ref32_sel"][b]))
        assert np.abs(cls.numpy()[0] - z["ref32_classes"][b]).max() < 1e-6
        assert np.abs(bag.numpy()[0] - z["ref32_bag"][b]).max() < 1e-6
    for name in ("bin_cfg4_packed", "bin_cfg4_packed_deep"):
        z, c = load(name)
        tp = _torch_params(make_snuffy_params(c["d"], c["depth"], 1, 4, c["wseed"], realistic=c.get("realistic", False)))
        for b, n in enumerate(c["lens"]):
            x = make_bag(n, c["d"], c["xseed"] + b, 1)
            with torch.no_grad():
                cls, bag, _ = torch_port.forward(torch.from_numpy(x), tp, c["heads"], c["K"], c["r"], c["depth"], c["act"],
                                                 selections=list(z[f"ref32_sel_{b}"]))
            assert np.abs(cls.numpy()[0] - z[f"ref32_classes_{b}"]).max() < 1e-6
            assert np.abs(bag.numpy()[0] - z["ref32_bag"][b]).max() < 1e-5


def test_pieces_oracle():
    """Stand-alone attention / FFN / SublayerConnection fixtures vs the numpy restatement."""
    z = np.load(os.path.join(GOLDEN, "pieces.npz"))
    q, k, v = (z[n].astype(np.float64) for n in ("att_q", "att_k", "att_v"))
    s = np.einsum("bhnd,bhkd->bhnk", q, k) / np.sqrt(q.shape[-1])
    p = np.exp(s - s.max(-1, keepdims=True)); p /= p.sum(-1, keepdims=True)
    assert np.abs(p - z["att_p"]).max() < 1e-5
    assert np.abs(np.einsum("bhnk,bhnd->bhkd", p, v) - z["att_out"]).max() < 1e-4
    x = z["sc_x"].astype(np.float64)
    u = so.layer_norm(x, z["sc_norm.weight"].astype(np.float64), z["sc_norm.bias"].astype(np.float64))
    h = so.activation_fn("selu")(u @ z["ffn_selu_w_1.weight"].T.astype(np.float64) + z["ffn_selu_w_1.bias"])
    ff = x + h @ z["ffn_selu_w_2.weight"].T.astype(np.float64) + z["ffn_selu_w_2.bias"]
    assert np.abs(ff - z["sc_ff_out"]).max() < 1e-4


def test_loss_glue():
    z = np.load(os.path.join(GOLDEN, "loss_glue.npz"))
    for i in range(len(z["w"])):
        loss, pred = so.mil_loss(z["c"][i], z["bag"][i], np.array([[z["y"][i]]]), float(z["w"][i]), z["pos_weight"])
        assert abs(loss - z["loss"][i]) < 1e-12
        assert np.abs(pred - z["pred"][i]).max() < 1e-12


def test_selection_edge_cases():
    cfg = so.SnuffyConfig(d=4, heads=1, big_lambda=200, random_patch_share=0.7)
    assert cfg.k_top == 61 and cfg.k_rand(10_000) == 140          # App. B-7
    assert cfg.k_rand(61) == 0 and cfg.k_rand(100) == 39
    # ties: lower index first
    c = np.array([1.0, 3.0, 3.0, 2.0, 3.0])
    assert so.select_top(c, 3).tolist() == [1, 2, 4]
    assert so.remaining_ascending(6, np.array([4, 1])).tolist() == [0, 2, 3, 5]
    with pytest.raises(KeyError):
        so.activation_fn("swish")
    with pytest.raises(ValueError):
        so.snuffy_forward(np.zeros((2, 4, 4)), {}, cfg)
