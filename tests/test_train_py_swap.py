"""The drop-in claim, executed: the reference's own `train.py` (unmodified; `Snuffy(args)` / `SnuffyMulticlass(args)`,
`.valid()`, `.train()`, `.valid()`: train.py:223-293, 295-411, 797-982) runs against `dropin/` on the GPU and produces the same
numbers as against the reference's own modules on the CPU, from identical weights (loaded through the state_dict boundary),
identical synthetic bags and the same NumPy random stream (sklearn's shuffle + np.random.choice, replayed by
`random_mode = "numpy"`), dropout forced to 0 in both arms.

The reference is looked up at $SNUFFY_REF, /root/reference, baseline/_ref (tools/stage_reference.py stages it there so that it
travels to the GPU box); the tests skip when none is present.  Tolerances: eval-mode losses 1e-4 (the parity bar); the loss
of a TRAINING epoch 2e-3 (8 AdamW steps amplify last-bit gradient differences), final eval loss 5e-3."""
import json
import os
import subprocess
import sys

import pytest

from conftest import ROOT

DRIVER = os.path.join(ROOT, "tests", "train_py_swap_driver.py")


def reference_dir():
    for cand in (os.environ.get("SNUFFY_REF"), "/root/reference", os.path.join(ROOT, "baseline", "_ref")):
        if cand and os.path.isfile(os.path.join(cand, "train.py")) and os.path.isfile(os.path.join(cand, "snuffy.py")):
            return cand
    return None


def run_driver(impl, arch, soft_average=0, hide_gpu=False):
    env = dict(os.environ, WANDB_MODE="disabled")
    if hide_gpu:
        env["CUDA_VISIBLE_DEVICES"] = ""
    res = subprocess.run([sys.executable, DRIVER, "--ref", reference_dir(), "--impl", impl, "--arch", arch, "--soft_average",
                          str(soft_average)], capture_output=True, text=True, env=env, timeout=900, cwd=ROOT)
    lines = [l for l in res.stdout.splitlines() if l.startswith("SWAP_RESULT ")]
    assert res.returncode == 0 and lines, (res.stdout[-2000:], res.stderr[-4000:])
    return json.loads(lines[-1][len("SWAP_RESULT "):])


needs_ref = pytest.mark.skipif(reference_dir() is None, reason="reference not present (SNUFFY_REF, /root/reference, baseline/_ref)")


@needs_ref
@pytest.mark.parametrize("arch", ["snuffy", "snuffy_multiclass"])
def test_reference_arm_of_the_driver_runs_on_the_cpu(arch):
    """CPU-only sanity of the harness itself: train.py + the reference's own modules, GPUs hidden."""
    r = run_driver("reference", arch, hide_gpu=True)
    assert r["device"] == "cpu" and "snuffy_b200" not in r["module_file"]
    assert r["params_moved"] == r["params_total"]
    assert all(v == v for v in (r["valid0"]["epoch_valid_loss"], r["train1"]["epoch_train_loss"], r["valid1"]["epoch_valid_loss"]))


@needs_ref
@pytest.mark.gpu
@pytest.mark.parametrize("arch,soft_average", [("snuffy", 0), ("snuffy", 1), ("snuffy_multiclass", 0)])
def test_train_py_runs_unchanged_on_the_dropin_and_matches_the_reference(arch, soft_average):
    ours = run_driver("ours", arch, soft_average)
    ref = run_driver("reference", arch, soft_average, hide_gpu=True)
    assert ours["device"] == "cuda" and os.sep + "dropin" + os.sep in ours["module_file"]       # train.py bound the drop-in
    assert ours["library_launches"] > 100                       # the CUDA library did the work (no fallback exists)
    assert ours["params_moved"] == ours["params_total"] == ref["params_total"]
    assert abs(ours["valid0"]["epoch_valid_loss"] - ref["valid0"]["epoch_valid_loss"]) < 1e-4, (ours["valid0"], ref["valid0"])
    assert ours["valid0"]["epoch_valid_accuracy"] == ref["valid0"]["epoch_valid_accuracy"]
    assert abs(ours["train1"]["epoch_train_loss"] - ref["train1"]["epoch_train_loss"]) < 2e-3, (ours["train1"], ref["train1"])
    assert abs(ours["valid1"]["epoch_valid_loss"] - ref["valid1"]["epoch_valid_loss"]) < 5e-3, (ours["valid1"], ref["valid1"])
    assert abs(ours["single_weight_parameter"] - ref["single_weight_parameter"]) < 1e-4
    if soft_average:
        assert ours["single_weight_parameter"] != 0.5
