"""Measured parity margins of the CUDA path against the reference's own fp32 outputs (tests/golden/*.npz), both precisions."""
import os, sys, json
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np, torch
from helpers import build_snuffy, force_selections, load_golden, load_params, set_precision, snuffy_inputs
from snuffy_b200 import snuffy, snuffy_multiclass
for name in ["bin_tiny_relu", "bin_rand_gelu", "bin_short_leaky", "bin_k201_selu", "bin_cfg1", "bin_cfg2", "bin_cfg2_rand",
             "mc_b1_c2", "mc_b3_c3", "mc_c1_r0", "mc_cfg3s"]:
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
        a = attn.cpu().numpy()
        e_att = float(np.abs(a - z["ref32_attn"]).max()) if "ref32_attn" in z else float(np.abs(a[..., z["sub_rows"], :] - z["ref32_attn_rows"]).max())
        print(json.dumps({"fixture": name, "N": c["n"], "d": c["d"], "depth": c["depth"], "precision": precision,
                          "err_classes": e_cls, "err_bag": e_bag, "err_attn": e_att}), flush=True)
