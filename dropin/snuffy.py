"""`import snuffy` shim: the reference module name bound to the B200-native implementation."""
from snuffy_b200.snuffy import *  # noqa: F401,F403
from snuffy_b200.snuffy import device  # noqa: F401
