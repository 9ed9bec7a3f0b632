"""Runtime package of the B200-native diffusion hot path (ctypes binding + op wrappers + U-Net engine)."""
from ._lib import B200Error, LIB_PATH, lib  # noqa: F401
