#!/usr/bin/env python
"""Stage the UNMODIFIED reference files the hot path's caller needs into baseline/_ref/ (git-ignored; it travels to the GPU
box with the gpurun snapshot, where /root/reference does not exist).

The reference has no setup.py / pyproject, so the `pip install --target baseline/_ref /root/reference` of the bench contract
does not apply; its seven pure-Python modules on this path are staged as they are (byte-identical copies, verified below):
  train.py (the caller: Trainer / SmallWeightTrainer / Snuffy / SnuffyMulticlass), utils.py, froc.py, metrics.py (its imports),
  snuffy.py, snuffy_multiclass.py, dsmil.py (the modules the drop-in replaces; used by the reference arms).
Used by tests/test_train_py_swap.py, bench.py --impl reference and bench.py's gpu_eager_reference leg.  Nothing under
snuffy_b200/ reads this directory.
"""
import filecmp
import os
import shutil
import sys

FILES = ["train.py", "utils.py", "froc.py", "metrics.py", "snuffy.py", "snuffy_multiclass.py", "dsmil.py"]
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def stage(ref_dir: str = None, quiet: bool = False) -> bool:
    ref_dir = ref_dir or os.environ.get("SNUFFY_REF", "/root/reference")
    if not os.path.isfile(os.path.join(ref_dir, "train.py")):
        if not quiet:
            print(f"[stage_reference] no reference at {ref_dir}; nothing staged")
        return False
    out = os.path.join(ROOT, "baseline", "_ref")
    os.makedirs(out, exist_ok=True)
    for f in FILES:
        dst = os.path.join(out, f)
        if not (os.path.exists(dst) and filecmp.cmp(os.path.join(ref_dir, f), dst, shallow=False)):
            shutil.copyfile(os.path.join(ref_dir, f), dst)
        assert filecmp.cmp(os.path.join(ref_dir, f), dst, shallow=False)
    if not quiet:
        print(f"[stage_reference] {len(FILES)} files -> {out}")
    return True


if __name__ == "__main__":
    stage(sys.argv[1] if len(sys.argv) > 1 else None)
