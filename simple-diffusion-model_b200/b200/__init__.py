"""Runtime package of the B200-native diffusion hot path (ctypes binding + op wrappers + U-Net engine)."""
from ._lib import B200Error, LIB_PATH, lib  # noqa: F401


def set_deterministic(on=True):
    """Deterministic mode of the C library (include/sdm_b200.h: b2_set_deterministic): GroupNorm statistics without fp atomics,
    no split-K on the forward kernels -- an image's result no longer depends on batch size, sharding or timing."""
    lib().b2_set_deterministic(1 if on else 0)
