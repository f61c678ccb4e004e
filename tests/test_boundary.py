"""CPU-side checks of the drop-in boundary: the C-ABI library loads without a GPU and exports
every symbol include/snuffy_b200.h declares; the nn.Modules keep the reference's class names, constructor
signatures, attribute names and state_dict keys; host logic (|S| rounding, error behaviour).  No compute."""
import copy
import ctypes
import inspect
import os
import re
import sys

import numpy as np
import pytest
import torch

from conftest import ROOT
from helpers import build_snuffy, load_golden, snuffy_inputs
from oracle.params import make_dsmil_params, make_snuffy_params


def header_functions():
    src = open(os.path.join(ROOT, "include", "snuffy_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return re.findall(r"\b(snuffy_[a-z0-9_]+)\s*\(", src)


def test_library_loads_and_exports_every_declared_symbol():
    from snuffy_b200 import _lib
    declared = set(header_functions())
    assert len(declared) >= 20
    raw = ctypes.CDLL(_lib.LIB_PATH)
    for name in declared:
        assert hasattr(raw, name), f"{name} declared in include/snuffy_b200.h but not exported"
    assert declared == set(_lib.EXPORTED_SYMBOLS), declared ^ set(_lib.EXPORTED_SYMBOLS)
    assert _lib.lib.snuffy_version() >= 100
    assert isinstance(_lib.last_error(), str)


def test_library_has_no_driver_or_torch_dependency():
    """The .so must load on a box without libcuda (this container) and carry no torch types in its ABI."""
    import subprocess
    from snuffy_b200 import _lib
    out = subprocess.run(["ldd", _lib.LIB_PATH], capture_output=True, text=True).stdout
    assert "libcuda.so" not in out and "libtorch" not in out and "libc10" not in out


def test_argument_validation_without_a_gpu():
    """Entry points validate before launching: bad arguments return non-zero and set the error string."""
    from snuffy_b200._lib import SnuffyLibraryError, check, last_error, lib
    assert lib.snuffy_scores_fwd(None, None, None, None, 10, 0, 1, None) != 0
    assert "bad shape" in last_error()
    assert lib.snuffy_select_topk(None, 1, 10, 1, 5, None, None, None) != 0
    with pytest.raises(SnuffyLibraryError):
        check(lib.snuffy_gemm_tc(None, 0, None, 0, 1, 1, 1, 3, None, 0, None, 0, None, None, None, 0, None, None, 0, 0.0,
                                 0, 0, None), "snuffy_gemm_tc")
    assert lib.snuffy_sparse_attn_workspace(1, 100, 10, 3, 64) == -1        # d % h != 0
    assert lib.snuffy_gemm_tc_block_n(2048) == 256 and lib.snuffy_gemm_tc_block_n(384) == 128
    assert lib.snuffy_plane_elems(10000, 512, 128) == 79 * 16 * 128 * 32
    # structs and handle sizes the bindings mirror
    from snuffy_b200 import _lib
    assert lib.snuffy_plane_job_bytes() == ctypes.sizeof(_lib.PlaneJob) == 80
    assert lib.snuffy_comm_handle_bytes() == 64 and lib.snuffy_comm_counter_bytes() >= 512
    assert lib.snuffy_weight_planes_batch(None, 0, None) != 0 and "jobs" in last_error()
    assert lib.snuffy_peer_allreduce(None, None, None, 0, 2, 16, None) != 0
    assert lib.snuffy_gemm_tc_splitk_rows(None, 0, None, 0, 128, 128, 16, 3, 0, None, None, 0, None) != 0


SNUFFY_CLASSES = ["FCLayer", "IClassifier", "BClassifier", "Encoder", "SublayerConnection", "EncoderLayer",
                  "MultiHeadedAttention", "PositionwiseFeedForward", "MILNet"]


@pytest.mark.parametrize("modname", ["snuffy", "snuffy_multiclass"])
def test_module_surface_matches_reference(modname):
    import importlib
    mod = importlib.import_module(f"snuffy_b200.{modname}")
    for cls in SNUFFY_CLASSES:
        assert inspect.isclass(getattr(mod, cls)), cls
    assert callable(mod.clones) and callable(mod.attention) and isinstance(mod.device, torch.device)
    sig = lambda c: list(inspect.signature(c.__init__).parameters)[1:]
    assert sig(mod.FCLayer) == ["in_size", "out_size"]
    assert sig(mod.IClassifier) == ["feature_extractor", "feature_size", "output_class"]
    assert sig(mod.MultiHeadedAttention) == ["h", "d_model", "dropout"]
    assert sig(mod.PositionwiseFeedForward) == ["d_model", "d_ff", "activation", "dropout"]
    assert sig(mod.Encoder) == ["layer", "N"]
    assert sig(mod.BClassifier) == ["encoder", "num_classes", "input_size"]
    assert sig(mod.MILNet) == ["i_classifier", "b_classifier"]
    if modname == "snuffy":
        assert sig(mod.EncoderLayer) == ["size", "self_attn", "feed_forward", "dropout", "big_lambda",
                                         "random_patch_share"]
    else:
        assert sig(mod.EncoderLayer) == ["size", "self_attn", "feed_forward", "num_class", "dropout", "big_lambda",
                                         "random_patch_share"]
    assert inspect.signature(mod.MultiHeadedAttention.__init__).parameters["dropout"].default == 0.1
    assert inspect.signature(mod.PositionwiseFeedForward.__init__).parameters["dropout"].default == 0.1


def test_state_dict_keys_and_checkpoint_round_trip():
    from snuffy_b200 import dsmil, snuffy, snuffy_multiclass
    for mod, multi, C in ((snuffy, False, 1), (snuffy_multiclass, True, 3)):
        c = dict(d=32, heads=2, K=8, r=0.5, depth=2, act="gelu", C=C)
        model = build_snuffy(mod, c, multiclass=multi)
        params = make_snuffy_params(32, 2, C)
        assert sorted(model.state_dict().keys()) == sorted(params.keys())
        model.load_state_dict({k: torch.from_numpy(v) for k, v in params.items()}, strict=True)
        for k, v in model.state_dict().items():
            assert tuple(v.shape) == params[k].shape, k
        # utils.py:69-120 re-initialises through .apply on nn.Linear children; parameters must be reachable there
        linears = [m for m in model.modules() if isinstance(m, torch.nn.Linear)]
        assert len(linears) == 1 + 2 * 6 + 1
        # callers deep-copy pieces (train.py:865,880-881) and read layer.size (snuffy.py:80)
        clone = copy.deepcopy(model)
        assert clone.b_classifier.encoder.layers[0].size == 32
        assert all(torch.equal(a, b) for a, b in zip(model.state_dict().values(), clone.state_dict().values()))
        assert clone.i_classifier.fc[0].weight.data_ptr() != model.i_classifier.fc[0].weight.data_ptr()
    pd = make_dsmil_params(64, 2, True, True)
    dm = dsmil.MILNet(dsmil.FCLayer(64, 2), dsmil.BClassifier(64, 2, 0.0, True, True))
    assert sorted(dm.state_dict().keys()) == sorted(pd.keys())
    assert dm.b_classifier.fcc.weight.shape == (2, 2, 64)
    pl = make_dsmil_params(48, 3, False, False)
    dl = dsmil.MILNet(dsmil.FCLayer(48, 3), dsmil.BClassifier(48, 3, 0.0, False, False))
    assert sorted(dl.state_dict().keys()) == sorted(pl.keys())


def test_dropin_shims_bind_the_reference_module_names():
    sys.path.insert(0, os.path.join(ROOT, "dropin"))
    try:
        for name in ("snuffy", "snuffy_multiclass", "dsmil"):
            sys.modules.pop(name, None)
        import dsmil
        import snuffy
        import snuffy_multiclass
        import snuffy_b200
        assert snuffy.MILNet is snuffy_b200.snuffy.MILNet
        assert snuffy_multiclass.EncoderLayer is snuffy_b200.snuffy_multiclass.EncoderLayer
        assert dsmil.BClassifier is snuffy_b200.dsmil.BClassifier
    finally:
        sys.path.remove(os.path.join(ROOT, "dropin"))
        for name in ("snuffy", "snuffy_multiclass", "dsmil"):
            sys.modules.pop(name, None)


def test_selection_sizes_follow_python_double_rounding():
    from snuffy_b200 import engine
    assert engine.k_top_of(200, 0.7) == 61 and engine.k_rand_of(200, 0.7, 10_000) == 140     # 201, App. B-7
    assert engine.k_top_of(200, 0.0) == 200 and engine.k_rand_of(200, 0.0, 10_000) == 0
    assert engine.k_rand_of(200, 0.7, 61) == 0 and engine.k_rand_of(200, 0.7, 100) == 39
    assert engine.k_top_of(32, 0.3) == 23


def test_no_cpu_fallback_and_reference_error_behaviour():
    from snuffy_b200 import snuffy
    _, c = load_golden("bin_tiny_relu")
    params, x = snuffy_inputs(c)
    model = build_snuffy(snuffy, c)
    with pytest.raises(RuntimeError, match="CUDA device only"):
        model(torch.from_numpy(x))
    with pytest.raises(KeyError):
        snuffy.PositionwiseFeedForward(32, 128, "swish")          # dict lookup, snuffy.py:216-222
    with pytest.raises(AssertionError):
        snuffy.MultiHeadedAttention(5, 32)                          # snuffy.py:176
    with pytest.raises(ValueError):
        os.environ["SNUFFY_B200_PRECISION"] = "fp8"
        try:
            from snuffy_b200 import engine
            engine.default_precision()
        finally:
            del os.environ["SNUFFY_B200_PRECISION"]


def test_product_package_never_imports_the_oracle():
    pkg = os.path.join(ROOT, "snuffy_b200")
    for fn in os.listdir(pkg):
        if fn.endswith(".py"):
            src = open(os.path.join(pkg, fn)).read()
            assert not re.search(r"^\s*(from|import)\s+oracle", src, flags=re.M), fn


def test_torch_port_matches_golden():
    """oracle/torch_port.py (the timed CPU baseline) reproduces the reference's fp32 outputs."""
    from oracle import torch_port
    for name in ("bin_tiny_relu", "bin_rand_gelu", "bin_k201_selu", "bin_cfg1"):
        z, c = load_golden(name)
        params, x = snuffy_inputs(c)
        tp = {k: torch.from_numpy(v) for k, v in params.items()}
        np.random.seed(c["npseed"])
        with torch.no_grad():
            cls, bag, attn = torch_port.forward(torch.from_numpy(x), tp, c["heads"], c["K"], c["r"], c["depth"], c["act"])
        assert np.abs(cls.numpy() - z["ref32_classes"]).max() < 1e-5
        assert np.abs(bag.numpy() - z["ref32_bag"]).max() < 1e-4
        if "ref32_attn" in z:
            assert attn.shape == z["ref32_attn"].shape
