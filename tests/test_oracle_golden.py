"""The oracle restatement vs the reference's own outputs (tests/golden, made by
oracle/make_golden.py from the unmodified reference).  CPU only."""
import json
import os

import numpy as np
import pytest

from oracle import snuffy_oracle as so
from oracle.params import make_bag, make_dsmil_params, make_snuffy_params
from conftest import GOLDEN

BIN = ["bin_tiny_relu", "bin_rand_gelu", "bin_short_leaky", "bin_k201_selu", "bin_cfg1", "bin_cfg2", "bin_cfg2_rand"]
MC = ["mc_b1_c2", "mc_b3_c3", "mc_c1_r0", "mc_cfg3s"]
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
    last = out["layers"][-1][out["selections"][-1]]
    assert np.abs(last - z["ref64_last_sel_rows"]).max() <= 1e-9 * max(1.0, np.abs(last).max())
    if "ref64_attn" in z:
        assert np.abs(out["attn"] - z["ref64_attn"]).max() < 1e-9
    else:
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
