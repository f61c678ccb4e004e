#!/usr/bin/env python
"""Runs the reference's OWN caller — train.py's `Snuffy(args)` / `SnuffyMulticlass(args)` trainers, `.valid()`, `.train()`,
`.valid()` — on synthetic bags, either against the drop-in modules (`--impl ours`: `dropin/` first on sys.path, so train.py's
`import snuffy` / `import snuffy_multiclass` bind snuffy_b200) or against the reference's own modules (`--impl reference`).
train.py is imported unmodified from `--ref`; only packages it imports but this path never uses (wandb, lightly, skimage,
matplotlib, ASAP's multiresolutionimageinterface) are replaced by inert stand-ins.  Prints one JSON line.

TEST INFRASTRUCTURE (run as a subprocess by tests/test_train_py_swap.py so that sys.path / sys.modules stay isolated).
"""
import argparse
import json
import os
import sys
import types

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)


class _Inert(types.ModuleType):
    """Module stand-in: any attribute is a callable that accepts anything and returns None (wandb.log, wandb.init ...)."""

    def __getattr__(self, k):
        if k.startswith("__"):
            raise AttributeError(k)

        def _noop(*a, **kw):
            return None
        _noop.__name__ = k
        return _noop


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--ref", required=True)
    ap.add_argument("--impl", choices=["ours", "reference"], required=True)
    ap.add_argument("--arch", choices=["snuffy", "snuffy_multiclass"], default="snuffy")
    ap.add_argument("--soft_average", type=int, default=0)
    a = ap.parse_args()

    for name in ["lightly", "lightly.utils", "lightly.utils.scheduler", "multiresolutionimageinterface", "skimage",
                 "skimage.measure", "matplotlib", "matplotlib.pyplot", "wandb"]:
        sys.modules.setdefault(name, _Inert(name))
    os.environ.setdefault("WANDB_MODE", "disabled")
    sys.path.insert(0, a.ref)
    if a.impl == "ours":
        sys.path.insert(0, os.path.join(ROOT, "dropin"))
        sys.path.insert(1, ROOT)
    import numpy as np
    import torch
    import train                                                   # the reference's caller, unmodified

    multiclass = a.arch == "snuffy_multiclass"
    C, d = (2, 64) if multiclass else (1, 64)
    argv = ["--arch", a.arch, "--feats_size", str(d), "--num_heads", "4", "--big_lambda", "16", "--random_patch_share", "0.25",
            "--depth", "2", "--num_classes", str(C), "--encoder_dropout", "0.0", "--lr", "0.001", "--num_epochs", "2",
            "--soft_average", str(a.soft_average), "--dataset", "tcga" if multiclass else "camelyon16"]
    args = train.get_args_parser().parse_args(argv)
    args = train.validate_args(args)                                # exactly what train.main() does (train.py:1005-1011)
    import ast
    args.betas = ast.literal_eval("".join(args.betas))
    args.weight_init__weight_init_i__weight_init_b = ast.literal_eval("".join(args.weight_init__weight_init_i__weight_init_b))

    torch.manual_seed(0)
    np.random.seed(0)
    trainer = (train.SnuffyMulticlass if multiclass else train.Snuffy)(args)
    mod = train.snuffy_multiclass if multiclass else train.snuffy
    # identical weights in both arms through the state_dict boundary (init draws differ between CPU and CUDA generators)
    sys.path.insert(0, ROOT)
    from oracle.params import make_snuffy_params
    sd = {k: torch.from_numpy(v) for k, v in make_snuffy_params(d, 2, C, 4, seed=21).items()}
    trainer.milnet.load_state_dict(sd, strict=True)
    for m in trainer.milnet.modules():
        if isinstance(m, torch.nn.Dropout):
            m.p = 0.0                                               # train.py leaves the attention dropout at 0.1 (App. B-6): RNG-free compare
    if a.impl == "ours":
        for layer in trainer.milnet.b_classifier.encoder.layers:
            layer.random_mode = "numpy"                             # replay the reference's NumPy global-RNG draws

    rs = np.random.RandomState(5)
    lens = [150, 97, 203, 128, 64, 177, 111, 190]
    feats = [rs.standard_normal((n, d)).astype(np.float32) for n in lens]
    if multiclass:
        labels = [np.eye(C, dtype=np.float32)[i % C] for i in range(len(lens))]
    else:
        labels = [np.array([float(i & 1)], dtype=np.float32) for i in range(len(lens))]
    for f, l in zip(feats, labels):                                  # make the task learnable: shift positive bags
        f[:8] += 1.5 * (l[0] if not multiclass else l.argmax())

    def pick(res, keys):
        return {k: (np.asarray(res[k], dtype=np.float64).tolist() if res.get(k) is not None else None) for k in keys}

    out = {"impl": a.impl, "arch": a.arch, "module_file": os.path.abspath(mod.__file__), "device": train.device,
           "soft_average": bool(args.soft_average)}
    np.random.seed(1)
    out["valid0"] = pick(trainer.valid((labels, feats, None, None, None)), ["epoch_valid_loss", "epoch_valid_accuracy"])
    np.random.seed(2)
    p0 = [p.detach().clone() for p in trainer.milnet.parameters()]
    out["train1"] = pick(trainer.train((labels, feats, None, None), 1), ["epoch_train_loss", "epoch_train_accuracy"])
    out["params_moved"] = int(sum(float((p.detach() - q).abs().max()) > 0 for p, q in zip(trainer.milnet.parameters(), p0)))
    out["params_total"] = len(p0)
    np.random.seed(3)
    out["valid1"] = pick(trainer.valid((labels, feats, None, None, None)), ["epoch_valid_loss", "epoch_valid_accuracy"])
    out["single_weight_parameter"] = float(trainer.single_weight_parameter)
    if a.impl == "ours":
        from snuffy_b200._lib import lib
        out["library_launches"] = int(lib.snuffy_launch_count())
    sys.stdout.write("\nSWAP_RESULT " + json.dumps(out) + "\n")


if __name__ == "__main__":
    main()
