"""Runtime package of the B200-native diffusion hot path (ctypes binding + op wrappers + U-Net engine)."""
from ._lib import B200Error, LIB_PATH, lib  # noqa: F401


_DETERMINISTIC = None


def set_deterministic(on=True):
    """Deterministic mode of the C library (include/sdm_b200.h: b2_set_deterministic): GroupNorm statistics without fp atomics,
    no split-K on the forward kernels -- an image's result no longer depends on batch size, sharding or timing."""
    global _DETERMINISTIC
    _DETERMINISTIC = bool(on)
    lib().b2_set_deterministic(1 if on else 0)


def is_deterministic():
    """Deterministic mode requested through set_deterministic() or SDM_B200_DETERMINISTIC=1 (the library's own default)."""
    import os
    if _DETERMINISTIC is not None:
        return _DETERMINISTIC
    return os.environ.get("SDM_B200_DETERMINISTIC", "0") == "1"


def set_option(name, value):
    """Kernel-selection switches of the C library (b2_set_option): "halo", "swap_ab" -- for A/B measurements and tests."""
    import ctypes
    if lib().b2_set_option(ctypes.cast(ctypes.c_char_p(name.encode()), ctypes.c_void_p), int(value)) != 0:
        raise B200Error(lib().b2_last_error().decode())
