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
# BASELINE.json configs[2] / configs[3] at their full shapes (`lean`: the fp64 run keeps only the small outputs)
BIG_BINARY_CASES = {
    "bin_cfg4_small": dict(n=1000, d=512, heads=8, K=64, r=0.5, depth=1, act="relu", wseed=0, xseed=4001, npseed=41,
                           realistic=True, lean=True),
    "bin_cfg4_big":   dict(n=50000, d=512, heads=8, K=1024, r=0.5, depth=1, act="relu", wseed=0, xseed=4002, npseed=42,
                           realistic=True, lean=True),
}
BIG_MULTI_CASES = {
    "mc_cfg3":    dict(n=6000, d=768, heads=8, K=200, r=0.5, depth=4, act="relu", C=2, B=1, wseed=8, xseed=3001, npseed=31,
                       realistic=True, lean=True),
}
# the bench workload: 16 different cfg2 bags (one reference forward each); `forward_bags` runs them as one batch
B16_CASE = dict(n=10000, d=512, heads=8, K=200, r=0.0, depth=1, act="relu", wseed=0, xseed=5000, npseed=0, realistic=True,
                bags=16)
# configs[3]: variable-length bags, one reference forward each; `forward_packed` runs them as one packed tensor
PACKED_CASES = {
    "bin_cfg4_packed":      dict(d=512, heads=8, K=256, r=0.5, depth=1, act="relu", wseed=0, xseed=6000, npseed=61,
                                 realistic=True, lens=[1000, 3037, 7211, 1500, 20000]),
    "bin_cfg4_packed_deep": dict(d=128, heads=4, K=64, r=0.25, depth=2, act="gelu", wseed=3, xseed=6100, npseed=62,
                                 lens=[300, 77, 1000, 128, 513, 64]),
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
        lean_skip = ("last_sel_rows", "attn", "attn_rows") if cfg.get("lean") else ()
        for k, v in o64.items():
            if k not in lean_skip:
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


def _ref_binary_forward(mod, model, x, npseed):
    """One reference forward of a [1, n, d] bag: (classes, bag, selections per layer, layer outputs)."""
    sel, layers = [], []

    def pre_hook(_m, args):
        top, rnd = args[3], args[4]
        top = top.detach().cpu().numpy().reshape(-1)
        sel.append((top if rnd is None else np.concatenate([top, rnd.cpu().numpy()])).astype(np.int64))

    hooks = []
    for layer in model.b_classifier.encoder.layers:
        hooks.append(layer.sublayer[0].register_forward_pre_hook(pre_hook))
        hooks.append(layer.register_forward_hook(lambda _m, _a, out: layers.append(out[0].detach().numpy().copy())))
    np.random.seed(npseed)
    with torch.no_grad():
        classes, bag, _ = model(torch.from_numpy(x))
    for h in hooks:
        h.remove()
    return classes.numpy(), bag.numpy(), np.stack(sel), layers


def gen_b16(ref_dir):
    """bin_cfg2_b16: the bench batch — 16 different cfg2 bags, one unmodified-reference fp32 forward each."""
    sys.path.insert(0, ref_dir)
    import importlib
    mod = importlib.import_module("snuffy")
    mod.device = torch.device("cpu")
    cfg = B16_CASE
    params = make_snuffy_params(cfg["d"], cfg["depth"], 1, 4, cfg["wseed"], realistic=True)
    model = _load(_build_snuffy(mod, cfg, False), params, torch.float32)
    rows = _sub_rows(cfg["n"])
    out = dict(classes=[], bag=[], sel=[], rows=[], sel_rows=[], xsum=[])
    for b in range(cfg["bags"]):
        x = make_bag(cfg["n"], cfg["d"], cfg["xseed"] + b, 1)
        classes, bag, sel, layers = _ref_binary_forward(mod, model, x, cfg["npseed"] + b)
        out["classes"].append(classes[0]); out["bag"].append(bag[0]); out["sel"].append(sel)
        out["rows"].append(layers[-1][0][rows]); out["sel_rows"].append(layers[-1][0][sel[-1][::8]])
        out["xsum"].append(x.astype(np.float64).sum())
    np.savez_compressed(os.path.join(OUT, "bin_cfg2_b16.npz"), config=np.array(json.dumps(cfg)), sub_rows=rows,
                        x_checksum=np.array(out["xsum"]), ref32_classes=np.stack(out["classes"]),
                        ref32_bag=np.stack(out["bag"]), ref32_sel=np.stack(out["sel"]),
                        ref32_layers_rows=np.stack(out["rows"]), ref32_sel_rows=np.stack(out["sel_rows"]))
    print("bin_cfg2_b16: bag", np.stack(out["bag"]).ravel()[:4], "...")


def gen_packed(ref_dir):
    """Variable-length bags (configs[3]): one unmodified-reference fp32 forward per bag, with the rows it selected."""
    sys.path.insert(0, ref_dir)
    import importlib
    mod = importlib.import_module("snuffy")
    mod.device = torch.device("cpu")
    for name, cfg in PACKED_CASES.items():
        params = make_snuffy_params(cfg["d"], cfg["depth"], 1, 4, cfg["wseed"], realistic=cfg.get("realistic", False))
        model = _load(_build_snuffy(mod, cfg, False), params, torch.float32)
        blob = dict(config=np.array(json.dumps(cfg)))
        bags, xsum = [], []
        for b, n in enumerate(cfg["lens"]):
            x = make_bag(n, cfg["d"], cfg["xseed"] + b, 1)
            classes, bag, sel, layers = _ref_binary_forward(mod, model, x, cfg["npseed"] + b)
            blob[f"ref32_classes_{b}"] = classes[0]
            blob[f"ref32_sel_{b}"] = sel                           # [depth, Ksel] LOCAL rows
            blob[f"ref32_sel_rows_{b}"] = layers[-1][0][sel[-1][::8]]
            bags.append(bag[0]); xsum.append(x.astype(np.float64).sum())
        blob["ref32_bag"] = np.stack(bags)
        blob["x_checksum"] = np.array(xsum)
        np.savez_compressed(os.path.join(OUT, name + ".npz"), **blob)
        print(f"{name}: bag", np.stack(bags).ravel())


def gen_pieces(ref_dir):
    """The stand-alone class surface of snuffy.py (attention 160-168, MultiHeadedAttention 183-205, PositionwiseFeedForward
    224-225, SublayerConnection 100-110, IClassifier 51-54) and dsmil.IClassifier (dsmil.py:39-50), each called directly."""
    sys.path.insert(0, ref_dir)
    import importlib
    mod = importlib.import_module("snuffy")
    mod.device = torch.device("cpu")
    dsm = importlib.import_module("dsmil")
    rs = np.random.RandomState(77)
    f32 = lambda *shape: rs.standard_normal(shape).astype(np.float32)
    blob = {}
    # attention(query, key, value)
    q, k, v = f32(2, 3, 50, 16), f32(2, 3, 7, 16), f32(2, 3, 50, 16)
    o, p = mod.attention(torch.from_numpy(q), torch.from_numpy(k), torch.from_numpy(v))
    blob.update(att_q=q, att_k=k, att_v=v, att_out=o.numpy(), att_p=p.numpy())
    # MultiHeadedAttention(h=4, d=64).eval()(query, key, value)
    d, h = 64, 4
    mha = mod.MultiHeadedAttention(h, d).eval()
    mha_w = {n_: (f32(*t.shape) * (0.15 if t.dim() == 2 else 0.1)) for n_, t in mha.state_dict().items()}
    mha.load_state_dict({n_: torch.from_numpy(t) for n_, t in mha_w.items()})
    xq, xk = f32(2, 40, d), f32(2, 9, d)
    with torch.no_grad():
        mo, mp = mha(torch.from_numpy(xq), torch.from_numpy(xk), torch.from_numpy(xq))
    blob.update({"mha_" + n_: t for n_, t in mha_w.items()})
    blob.update(mha_query=xq, mha_key=xk, mha_out=mo.numpy(), mha_attn=mp.numpy())
    # PositionwiseFeedForward, all four activations
    xf = f32(2, 40, d)
    blob["ffn_x"] = xf
    for act in ("relu", "gelu", "leakyrelu", "selu"):
        ffn = mod.PositionwiseFeedForward(d, 4 * d, act, 0.0).eval()
        fw = {n_: (f32(*t.shape) * (0.1 if t.dim() == 2 else 0.1)) for n_, t in ffn.state_dict().items()}
        ffn.load_state_dict({n_: torch.from_numpy(t) for n_, t in fw.items()})
        with torch.no_grad():
            blob[f"ffn_{act}_out"] = ffn(torch.from_numpy(xf)).numpy()
        blob.update({f"ffn_{act}_" + n_: t for n_, t in fw.items()})
    # SublayerConnection, both modes (the FFN / MHA above as the sublayer)
    sc = mod.SublayerConnection(d, 0.0).eval()
    sw = {"norm.weight": (1 + 0.1 * f32(d)), "norm.bias": 0.1 * f32(d)}
    sc.load_state_dict({n_: torch.from_numpy(t) for n_, t in sw.items()})
    blob.update({"sc_" + n_: t for n_, t in sw.items()})
    xs_ = f32(1, 40, d)
    top, rnd = np.array([5, 1, 33, 20, 8], dtype=np.int64), np.array([0, 39, 17], dtype=np.int64)
    with torch.no_grad():
        ff_out = sc(torch.from_numpy(xs_), ffn, None, None, None, 'ff')
        xt = torch.from_numpy(xs_)
        keys = torch.index_select(xt, 1, torch.from_numpy(np.concatenate([top, rnd])))
        a_out, a_p = sc(xt, lambda u: mha(u, keys, u), None, torch.from_numpy(top), torch.from_numpy(rnd), 'attn')
        a_out2, a_p2 = sc(xt, lambda u: mha(u, keys[:, :5], u), None, torch.from_numpy(top), None, 'attn')
    blob.update(sc_x=xs_, sc_top=top, sc_rnd=rnd, sc_ff_out=ff_out.numpy(), sc_attn_out=a_out.numpy(), sc_attn_p=a_p.numpy(),
                sc_attn_out_norand=a_out2.numpy(), sc_attn_p_norand=a_p2.numpy())
    # IClassifier over a small conv backbone (roi.py:177,324 / compute_feats.py:242 call form)
    def backbone():
        return torch.nn.Sequential(torch.nn.Conv2d(3, 8, 3, padding=1), torch.nn.ReLU(), torch.nn.AdaptiveAvgPool2d(2))
    imgs = f32(12, 3, 16, 16)
    blob["ic_imgs"] = imgs
    for tag, m_, ncls in (("snuffy", mod, 1), ("dsmil", dsm, 3)):
        ic = m_.IClassifier(backbone(), 32, ncls).eval()
        iw = {n_: f32(*t.shape) * 0.2 for n_, t in ic.state_dict().items()}
        ic.load_state_dict({n_: torch.from_numpy(t) for n_, t in iw.items()})
        with torch.no_grad():
            feats, c = ic(torch.from_numpy(imgs))
        blob.update({f"ic_{tag}_" + n_: t for n_, t in iw.items()})
        blob[f"ic_{tag}_feats"], blob[f"ic_{tag}_c"] = feats.numpy(), c.numpy()
    np.savez_compressed(os.path.join(OUT, "pieces.npz"), **blob)
    print("pieces:", len(blob), "arrays")


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--ref", default=os.environ.get("SNUFFY_REF", "/root/reference"))
    ap.add_argument("--only", default=None, help="generate one group: patch_outputs | big | b16 | packed | pieces")
    args = ap.parse_args()
    os.makedirs(OUT, exist_ok=True)
    torch.set_num_threads(os.cpu_count() or 1)
    groups = {"patch_outputs": lambda: gen_patch_outputs(args.ref), "b16": lambda: gen_b16(args.ref),
              "packed": lambda: gen_packed(args.ref), "pieces": lambda: gen_pieces(args.ref),
              "big": lambda: (gen_snuffy(args.ref, BIG_BINARY_CASES, multiclass=False),
                              gen_snuffy(args.ref, BIG_MULTI_CASES, multiclass=True))}
    if args.only is not None:
        groups[args.only]()
        return
    gen_snuffy(args.ref, BINARY_CASES, multiclass=False)
    gen_snuffy(args.ref, MULTI_CASES, multiclass=True)
    gen_dsmil(args.ref)
    gen_loss(args.ref)
    gen_patch_outputs(args.ref)
    for g in ("big", "b16", "packed", "pieces"):
        groups[g]()


if __name__ == "__main__":
    main()
