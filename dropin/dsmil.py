"""`import dsmil` shim: the reference module name bound to the B200-native implementation."""
from snuffy_b200.dsmil import *  # noqa: F401,F403
