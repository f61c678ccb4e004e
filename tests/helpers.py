"""Shared helpers for the parity tests: build the drop-in modules exactly as train.py:862-890 / 924-952 does
and load the deterministic synthetic weights of oracle/params.py."""
import copy
import json
import os

import numpy as np
import torch

from conftest import GOLDEN
from oracle import snuffy_oracle as so
from oracle.params import make_bag, make_dsmil_params, make_snuffy_params


def load_golden(name):
    z = np.load(os.path.join(GOLDEN, name + ".npz"))
    return z, json.loads(str(z["config"]))


def oracle_cfg(c):
    return so.SnuffyConfig(d=c["d"], heads=c["heads"], big_lambda=c["K"], random_patch_share=c["r"],
                           depth=c["depth"], activation=c["act"], num_classes=c.get("C", 1))


def snuffy_inputs(c):
    params = make_snuffy_params(c["d"], c["depth"], c.get("C", 1), 4, c["wseed"], realistic=c.get("realistic", False))
    x = make_bag(c["n"], c["d"], c["xseed"], c.get("B", 1))
    return params, x


def build_snuffy(mod, c, multiclass=False, ff_dropout=0.0, enc_dropout=0.0):
    d, C = c["d"], c.get("C", 1)
    i_cls = mod.FCLayer(in_size=d, out_size=C)
    attn = mod.MultiHeadedAttention(c["heads"], d)
    ff = mod.PositionwiseFeedForward(d, d * 4, c["act"], ff_dropout)
    if multiclass:
        layer = mod.EncoderLayer(d, copy.deepcopy(attn), copy.deepcopy(ff), C, enc_dropout, c["K"], c["r"])
    else:
        layer = mod.EncoderLayer(d, copy.deepcopy(attn), copy.deepcopy(ff), enc_dropout, c["K"], c["r"])
    b_cls = mod.BClassifier(mod.Encoder(layer, c["depth"]), C, d)
    return mod.MILNet(i_cls, b_cls)


def load_params(model, params, device="cuda"):
    model.load_state_dict({k: torch.from_numpy(np.asarray(v)) for k, v in params.items()}, strict=True)
    return model.to(device).eval()


def set_precision(model, precision):
    for layer in model.b_classifier.encoder.layers:
        layer.precision = precision


def force_selections(model, sels):
    """sels: [depth, Ksel] or [depth, B, Ksel] array of the reference's selected rows."""
    for layer, s in zip(model.b_classifier.encoder.layers, sels):
        layer.forced_selection = None if s is None else torch.as_tensor(np.asarray(s), dtype=torch.int64)


def build_dsmil(mod, c):
    params = make_dsmil_params(c["d"], c["C"], c["nonlinear"], c["passing_v"], c["wseed"])
    model = mod.MILNet(mod.FCLayer(c["d"], c["C"]), mod.BClassifier(c["d"], c["C"], 0.0, c["nonlinear"], c["passing_v"]))
    model.load_state_dict({k: torch.from_numpy(v) for k, v in params.items()}, strict=True)
    return model, params, make_bag(c["n"], c["d"], c["xseed"], 1)[0]


def decode_planes(p, rows, K):
    """snuffy_b200.ops.Planes -> (hi, lo) float64 [rows, Kpad], following the layout in csrc/common.cuh."""
    buf = p.buf.float().cpu().numpy().astype(np.float64)
    rc = p.rc
    kbn = (K + 31) // 32
    rt = (rows + rc - 1) // rc
    out = []
    for plane in range(2):
        a = buf[plane * p.stride:(plane + 1) * p.stride].reshape(rt, kbn, 4, rc, 8)
        out.append(a.transpose(0, 3, 1, 2, 4).reshape(rt * rc, kbn * 32)[:rows])
    return out
