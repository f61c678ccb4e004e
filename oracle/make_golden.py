"""Generate tests/golden/*.npz by running the UNMODIFIED reference modules.

Run in the build container only (``/root/reference`` does not exist on the GPU
box):   python oracle/make_golden.py [--ref /root/reference]

For every case the fixture stores the config, the seeds that regenerate the
inputs/weights through ``oracle/params.py``, and the reference's own outputs in
fp64 (``ref64_*`` — pins the oracle restatement) and fp32 (``ref32_*`` — what the
CUDA path is compared against at the 1e-4 bar).  The random-patch indices the
reference drew from NumPy's global RNG are captured with a forward pre-hook on
``EncoderLayer.sublayer[0]`` (its positional args carry T and R: snuffy.py:148-150),
so nothing in the reference is edited or monkey-patched.

TEST INFRASTRUCTURE — see oracle/snuffy_oracle.py header.
"""
from __future__ import annotations

import argparse
import copy
import json
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
from oracle.params import make_bag, make_dsmil_params, make_snuffy_params  # noqa: E402

OUT = os.path.join(os.path.dirname(HERE), "tests", "golden")

# name -> config.  `store_x`: keep the input in the file (tiny cases only).
BINARY_CASES = {
    "bin_tiny_relu":   dict(n=64, d=32, heads=2, K=8, r=0.0, depth=1, act="relu", wseed=1, xseed=11, npseed=5),
    "bin_rand_gelu":   dict(n=96, d=32, heads=4, K=16, r=0.5, depth=2, act="gelu", wseed=2, xseed=12, npseed=6),
    "bin_short_leaky": dict(n=10, d=32, heads=1, K=32, r=0.3, depth=1, act="leakyrelu", wseed=3, xseed=13, npseed=7),
    "bin_k201_selu":   dict(n=300, d=64, heads=8, K=200, r=0.7, depth=3, act="selu", wseed=4, xseed=14, npseed=8),
    "bin_cfg1":        dict(n=256, d=384, heads=1, K=32, r=0.0, depth=1, act="relu", wseed=0, xseed=1234, npseed=0),
    "bin_cfg2":        dict(n=10000, d=512, heads=8, K=200, r=0.0, depth=1, act="relu", wseed=0, xseed=1234, npseed=0,
                            realistic=True),
    "bin_cfg2_rand":   dict(n=10000, d=512, heads=8, K=200, r=0.5, depth=1, act="relu", wseed=0, xseed=1235, npseed=1,
                            realistic=True),
}
MULTI_CASES = {
    "mc_b1_c2":   dict(n=80, d=32, heads=2, K=12, r=0.5, depth=2, act="relu", C=2, B=1, wseed=5, xseed=15, npseed=9),
    "mc_b3_c3":   dict(n=120, d=48, heads=4, K=20, r=0.25, depth=1, act="gelu", C=3, B=3, wseed=6, xseed=16, npseed=10),
    "mc_c1_r0":   dict(n=100, d=32, heads=1, K=40, r=0.0, depth=1, act="relu", C=1, B=1, wseed=7, xseed=17, npseed=11),
    "mc_cfg3s":   dict(n=600, d=96, heads=8, K=40, r=0.5, depth=4, act="relu", C=2, B=1, wseed=8, xseed=18, npseed=12),
}
DSMIL_CASES = {
    "ds_c1":        dict(n=100, d=64, C=1, nonlinear=True, passing_v=False, wseed=9, xseed=19),
    "ds_c3_linear": dict(n=150, d=48, C=3, nonlinear=False, passing_v=False, wseed=10, xseed=20),
    "ds_c2_v":      dict(n=90, d=32, C=2, nonlinear=True, passing_v=True, wseed=11, xseed=21),
    "ds_cfg2":      dict(n=10000, d=512, C=1, nonlinear=True, passing_v=False, wseed=12, xseed=22),
}


def _build_snuffy(mod, cfg, multiclass):
    """Same constructor chain as train.py:862-890 / 924-952."""
    d, C = cfg["d"], cfg.get("C", 1)
    i_cls = mod.FCLayer(in_size=d, out_size=C)
    attn = mod.MultiHeadedAttention(cfg["heads"], d)
    ff = mod.PositionwiseFeedForward(d, d * 4, cfg["act"], 0.0)
    if multiclass:
        layer = mod.EncoderLayer(d, copy.deepcopy(attn), copy.deepcopy(ff), C, 0.0, cfg["K"], cfg["r"])
    else:
        layer = mod.EncoderLayer(d, copy.deepcopy(attn), copy.deepcopy(ff), 0.0, cfg["K"], cfg["r"])
    b_cls = mod.BClassifier(mod.Encoder(layer, cfg["depth"]), C, d)
    return mod.MILNet(i_cls, b_cls)


def _load(model, params, dtype):
    sd = {k: torch.from_numpy(v).to(dtype) for k, v in params.items()}
    model.load_state_dict(sd, strict=True)
    return model.to(dtype).eval()


def _run_snuffy(mod, cfg, multiclass, dtype, sub_rows):
    params = make_snuffy_params(cfg["d"], cfg["depth"], cfg.get("C", 1), 4, cfg["wseed"],
                                realistic=cfg.get("realistic", False))
    x = make_bag(cfg["n"], cfg["d"], cfg["xseed"], cfg.get("B", 1))
    model = _load(_build_snuffy(mod, cfg, multiclass), params, dtype)
    sel, layers = [], []

    def pre_hook(_m, args):
        top, rnd = args[3], args[4]
        top = top.detach().cpu().numpy()
        if multiclass:
            s = np.concatenate([top, rnd.detach().cpu().numpy()], axis=1)
        else:
            s = top.reshape(-1) if rnd is None else np.concatenate([top.reshape(-1), rnd.cpu().numpy()])
        sel.append(s.astype(np.int64))

    def layer_hook(_m, _a, out):
        layers.append(out[0].detach().numpy().copy())

    hooks = []
    for layer in model.b_classifier.encoder.layers:
        hooks.append(layer.sublayer[0].register_forward_pre_hook(pre_hook))
        hooks.append(layer.register_forward_hook(layer_hook))
    np.random.seed(cfg["npseed"])
    with torch.no_grad():
        classes, bag, attn = model(torch.from_numpy(x).to(dtype))
    for h in hooks:
        h.remove()
    out = {
        "classes": classes.numpy(), "bag": bag.numpy(),
        "sel": np.stack(sel),                       # [depth, Ksel] or [depth, B, Ksel]
        "layers_rows": np.stack([l[:, sub_rows, :] for l in layers]),   # sampled rows of each layer output
        "layers_absmax": np.array([np.abs(l).max() for l in layers]),
        "layers_sum": np.array([l.astype(np.float64).sum() for l in layers]),
    }
    # the selected rows are where attention writes: keep all of them for the last layer
    last_sel = sel[-1] if not multiclass else sel[-1][0]
    out["last_sel_rows"] = layers[-1][0][last_sel]
    a = attn.numpy()
    if a.size <= 200_000:
        out["attn"] = a
    else:                                           # strided sample of query rows
        out["attn_rows"] = a[..., sub_rows, :]
    return x, out


def _sub_rows(n):
    return np.unique(np.linspace(0, n - 1, num=min(n, 16)).astype(np.int64))


def gen_snuffy(ref_dir, cases, multiclass):
    sys.path.insert(0, ref_dir)
    import importlib
    mod = importlib.import_module("snuffy_multiclass" if multiclass else "snuffy")
    mod.device = torch.device("cpu")   # module attribute only decides where index tensors go (App. B-5)
    for name, cfg in cases.items():
        rows = _sub_rows(cfg["n"])
        x, o64 = _run_snuffy(mod, cfg, multiclass, torch.float64, rows)
        _, o32 = _run_snuffy(mod, cfg, multiclass, torch.float32, rows)
        assert np.array_equal(o64["sel"], o32["sel"]), f"{name}: fp32/fp64 selections differ (tie?)"
        blob = {"config": np.array(json.dumps(cfg)), "sub_rows": rows}
        if x.size <= 20_000:
            blob["x"] = x
        blob["x_checksum"] = np.array(x.astype(np.float64).sum())
        for k, v in o64.items():
            blob["ref64_" + k] = v
        for k, v in o32.items():
            blob["ref32_" + k] = v
        np.savez_compressed(os.path.join(OUT, name + ".npz"), **blob)
        print(f"{name}: bag64={o64['bag'].ravel()} bag32={o32['bag'].ravel()} sel={o64['sel'].shape}")


def gen_dsmil(ref_dir):
    sys.path.insert(0, ref_dir)
    import importlib
    mod = importlib.import_module("dsmil")
    for name, cfg in DSMIL_CASES.items():
        params = make_dsmil_params(cfg["d"], cfg["C"], cfg["nonlinear"], cfg["passing_v"], cfg["wseed"])
        x = make_bag(cfg["n"], cfg["d"], cfg["xseed"], 1)[0]
        blob = {"config": np.array(json.dumps(cfg)), "x_checksum": np.array(x.astype(np.float64).sum())}
        if x.size <= 20_000:
            blob["x"] = x
        for tag, dtype in (("ref64_", torch.float64), ("ref32_", torch.float32)):
            model = mod.MILNet(mod.FCLayer(cfg["d"], cfg["C"]),
                               mod.BClassifier(cfg["d"], cfg["C"], 0.0, cfg["nonlinear"], cfg["passing_v"]))
            model = _load(model, params, dtype)
            with torch.no_grad():
                classes, bag, a = model(torch.from_numpy(x).to(dtype))
                _, _, bmat = model.b_classifier(torch.from_numpy(x).to(dtype), classes)
            blob[tag + "classes"] = classes.numpy()
            blob[tag + "bag"] = bag.numpy()
            blob[tag + "attn"] = a.numpy() if a.numel() <= 200_000 else a.numpy()[_sub_rows(cfg["n"])]
            blob[tag + "B"] = bmat.numpy()
        np.savez_compressed(os.path.join(OUT, name + ".npz"), **blob)
        print(f"{name}: bag={blob['ref64_bag'].ravel()}")


def gen_loss(ref_dir):
    """Loss glue of train.py:828-846 evaluated with torch (BCEWithLogitsLoss) on fixed logits."""
    rs = np.random.RandomState(3)
    cases = []
    for w in (0.5, 0.2, 1.0):
        for y in (0.0, 1.0):
            c = rs.standard_normal((1, 50, 1))
            bag = rs.standard_normal((1, 1))
            crit = torch.nn.BCEWithLogitsLoss(torch.tensor([1.7], dtype=torch.float64))
            mx, _ = torch.max(torch.from_numpy(c), 1)
            yl = torch.tensor([[y]], dtype=torch.float64)
            loss = w * crit(torch.from_numpy(bag).view(1, -1), yl) + (1 - w) * crit(mx.view(1, -1), yl)
            pred = (1 - w) * torch.sigmoid(mx) + w * torch.sigmoid(torch.from_numpy(bag))
            cases.append(dict(w=w, y=y, c=c, bag=bag, loss=float(loss), pred=pred.numpy().ravel()))
    np.savez_compressed(os.path.join(OUT, "loss_glue.npz"),
                        w=np.array([c["w"] for c in cases]), y=np.array([c["y"] for c in cases]),
                        c=np.stack([c["c"] for c in cases]), bag=np.stack([c["bag"] for c in cases]),
                        loss=np.array([c["loss"] for c in cases]), pred=np.stack([c["pred"] for c in cases]),
                        pos_weight=np.array([1.7]))
    print("loss_glue: ", [round(c["loss"], 6) for c in cases])


def _import_reference_train(ref_dir):
    """The caller module itself, with stand-ins for the logging / plotting / WSI packages it imports but this path never
    touches: nothing of the reference is edited."""
    import types

    class _Any(types.ModuleType):
        def __getattr__(self, k):
            if k.startswith("__"):
                raise AttributeError(k)
            return type(k, (), {})
    for name in ["lightly", "lightly.utils", "lightly.utils.scheduler", "multiresolutionimageinterface", "skimage",
                 "skimage.measure", "matplotlib", "matplotlib.pyplot", "wandb"]:
        sys.modules.setdefault(name, _Any(name))
    os.environ.setdefault("WANDB_MODE", "disabled")
    if ref_dir not in sys.path:
        sys.path.insert(0, ref_dir)
    import train
    return train


def gen_patch_outputs(ref_dir):
    """Patch probabilities from the reference's Snuffy._run_model (train.py:913-916 over 828-846), detection tuples by the
    expression of train.py:342-345 and the filter `mp_thresholding` (train.py:138-141) itself."""
    import re
    train = _import_reference_train(ref_dir)
    rs = np.random.RandomState(21)
    n = 300
    logits = (rs.standard_normal((1, n, 1)) * 3).astype(np.float32)
    logits[0, :6, 0] = [0.0, -0.0, 40.0, -40.0, 100.0, -100.0]
    bag = rs.standard_normal((1, 1)).astype(np.float32)
    names = [f"patch_{rs.randint(0, 400)}_{rs.randint(0, 400)}.jpeg" for _ in range(n)]
    trainer = object.__new__(train.Snuffy)
    trainer.milnet = lambda x: (torch.from_numpy(logits), torch.from_numpy(bag), None)
    trainer.criterion = torch.nn.BCEWithLogitsLoss()
    trainer.single_weight_parameter = torch.tensor(0.5)
    pred, loss, att = trainer._run_model(torch.zeros(1, n, 4), torch.tensor([[1.0]]))
    probs32 = att.numpy()
    reg = r'[^\d]*(\d+)[^\d]*(\d+)[^\d]*'                                                   # train.py:313
    pos = [tuple(map(int, re.search(reg, p).group(1, 2))) for p in names]                      # train.py:314-320
    dets = [(float(prob), position[0] * 512 + 256, position[1] * 512 + 256)
            for position, prob in zip(pos, probs32.squeeze())]                                 # train.py:342-345
    thresholds = [0.5, float(np.sort(probs32.ravel())[n // 3]), 0.0, 1.0]
    kept = [train.mp_thresholding((dets, t, "slide"))[1] for t in thresholds]
    np.savez_compressed(os.path.join(OUT, "patch_outputs.npz"), logits=logits, bag=bag, names=np.array(names),
                        probs32=probs32, pred=np.asarray(pred), positions=np.array(pos, dtype=np.int64),
                        detections=np.array(dets, dtype=np.float64), thresholds=np.array(thresholds, dtype=np.float64),
                        **{f"kept{i}": np.array(k, dtype=np.float64).reshape(-1, 3) for i, k in enumerate(kept)})
    print("patch_outputs: kept", [len(k) for k in kept])


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--ref", default=os.environ.get("SNUFFY_REF", "/root/reference"))
    ap.add_argument("--only", default=None, help="generate one group: patch_outputs")
    args = ap.parse_args()
    os.makedirs(OUT, exist_ok=True)
    torch.set_num_threads(os.cpu_count() or 1)
    if args.only == "patch_outputs":
        gen_patch_outputs(args.ref)
        return
    gen_snuffy(args.ref, BINARY_CASES, multiclass=False)
    gen_snuffy(args.ref, MULTI_CASES, multiclass=True)
    gen_dsmil(args.ref)
    gen_loss(args.ref)
    gen_patch_outputs(args.ref)


if __name__ == "__main__":
    main()
