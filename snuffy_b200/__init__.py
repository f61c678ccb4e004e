"""snuffy_b200 — B200-native (sm_100a) implementation of the Snuffy / DSMIL MIL-aggregator hot path.

Importing the package loads ``libsnuffy_b200.so`` (built by ``./build.sh`` / ``__graft_entry__.build()``);
there is no CPU or PyTorch fallback.  Drop-in modules: :mod:`snuffy_b200.snuffy`,
:mod:`snuffy_b200.snuffy_multiclass`, :mod:`snuffy_b200.dsmil` (also importable under the reference's own
names by putting ``dropin/`` first on ``sys.path``).
"""
from . import _lib  # noqa: F401  (fails loudly if the native library is missing)
from . import dp, dsmil, engine, ops, patch_outputs, snuffy, snuffy_multiclass, store  # noqa: F401

from .dp import invalidate_weight_caches  # noqa: F401,E402  (call after editing parameters through `.data` / raw pointers)

__all__ = ["snuffy", "snuffy_multiclass", "dsmil", "ops", "engine", "dp", "store", "invalidate_weight_caches"]
__version__ = "0.1.0"
