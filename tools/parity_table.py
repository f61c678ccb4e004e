"""Measured parity margins of the CUDA path against the reference's own fp32 outputs (tests/golden/*.npz), both precisions."""
import os, sys, json
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np, torch
from helpers import build_snuffy, force_selections, load_golden, load_params, set_precision, snuffy_inputs
from snuffy_b200 import snuffy, snuffy_multiclass
for name in ["bin_tiny_relu", "bin_rand_gelu", "bin_short_leaky", "bin_k201_selu", "bin_cfg1", "bin_cfg2", "bin_cfg2_rand",
             "bin_cfg4_small", "bin_cfg4_big", "mc_b1_c2", "mc_b3_c3", "mc_c1_r0", "mc_cfg3s", "mc_cfg3"]:
    z, c = load_golden(name)
    mc = name.startswith("mc")
    mod = snuffy_multiclass if mc else snuffy
    params, x = snuffy_inputs(c)
    for precision in ("fp32", "bf16x3"):
        model = load_params(build_snuffy(mod, c, multiclass=mc), params)
        set_precision(model, precision)
        force_selections(model, z["ref32_sel"])
        with torch.no_grad():
            classes, bag, attn = model(torch.from_numpy(x).cuda())
        e_cls = float(np.abs(classes.cpu().numpy() - z["ref32_classes"]).max())
        e_bag = float(np.abs(bag.cpu().numpy() - z["ref32_bag"]).max())
        if "ref32_attn" in z:
            e_att = float(np.abs(attn.cpu().numpy() - z["ref32_attn"]).max())
        else:
            rows = torch.from_numpy(z["sub_rows"]).cuda()
            e_att = float(np.abs(attn[..., rows, :].cpu().numpy() - z["ref32_attn_rows"]).max())
        print(json.dumps({"fixture": name, "N": c["n"], "d": c["d"], "depth": c["depth"], "precision": precision,
                          "err_classes": e_cls, "err_bag": e_bag, "err_attn": e_att}), flush=True)

# the bench batch: 16 different cfg2 bags through forward_bags (fused scorer + shared normalised planes, A not materialised)
from oracle.params import make_bag, make_snuffy_params
z, c = load_golden("bin_cfg2_b16")
params = make_snuffy_params(c["d"], c["depth"], 1, 4, c["wseed"], realistic=True)
x = torch.from_numpy(np.concatenate([make_bag(c["n"], c["d"], c["xseed"] + b, 1) for b in range(c["bags"])])).cuda()
for precision in ("fp32", "bf16x3"):
    model = load_params(build_snuffy(snuffy, c), params)
    set_precision(model, precision)
    for layer in model.b_classifier.encoder.layers:
        layer.return_attn = False
    with torch.no_grad():
        classes, bag, _ = snuffy.forward_bags(model, x)
    print(json.dumps({"fixture": "bin_cfg2_b16 (forward_bags, 16 bags)", "N": c["n"], "d": c["d"], "depth": 1, "precision": precision,
                      "err_classes": float(np.abs(classes.cpu().numpy() - z["ref32_classes"]).max()),
                      "err_bag": float(np.abs(bag.cpu().numpy() - z["ref32_bag"]).max())}), flush=True)
