"""ctypes binding of libsdm_b200.so (C ABI declared in include/sdm_b200.h).

There is deliberately no fallback: if the shared library is missing or a tensor is not on a CUDA device the
call raises.  Build the library with `make` at the repo root (or `python -c "import __graft_entry__ as g; g.build()"`).
"""
import ctypes
import os

import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
# SDM_B200_LIB: an alternative build of the same library (kernel A/B runs from tools/); the product path is the in-tree one
LIB_PATH = os.environ.get("SDM_B200_LIB") or os.path.join(os.path.dirname(_HERE), "lib", "libsdm_b200.so")

_P, _I, _L, _F, _U, _D = (ctypes.c_void_p, ctypes.c_int, ctypes.c_longlong, ctypes.c_float, ctypes.c_ulonglong,
                          ctypes.c_double)

# name -> argument ctypes, in the order of include/sdm_b200.h
SIGNATURES = {
    "b2_set_workspace": [_P, _L],
    "b2_set_deterministic": [_I],
    "b2_zero": [_P, _L, _P],
    "b2_set_option": [_P, _I],
    "b2_conv2d_nhwc": [_I, _P, _I, _I, _I, _I, _L, _P, _P, _I, _P, _L, _I, _P, _L, _P, _I, _I, _I, _P],
    "b2_conv2d_nhwc_colsum": [_I, _P, _I, _I, _I, _I, _L, _P, _P, _I, _P, _L, _I, _P, _L, _P, _I, _I, _I, _P, _P, _L, _P, _P],
    "b2_conv2d_nhwc_dual": [_I, _P, _I, _I, _I, _I, _L, _P, _P, _I, _P, _L, _P, _L, _I, _P],
    "b2_adagn_bwd_fused": [_P, _L, _P, _L, _P, _P, _P, _P, _L, _P, _P, _L, _P, _P, _P, _L, _P, _I, _I, _I, _I, _F, _I, _I, _P],
    "b2_conv3x3_first": [_P, _P, _P, _P, _L, _I, _I, _I, _I, _I, _I, _I, _P],
    "b2_conv3x3_last": [_P, _L, _P, _P, _P, _I, _I, _I, _I, _I, _I, _I, _P],
    "b2_gemm_nt": [_P, _L, _L, _L, _P, _L, _L, _L, _P, _L, _L, _L, _I, _I, _I, _I, _I, _P, _F, _I, _P, _L, _I, _I, _P],
    "b2_gemm_nt_bmn": [_P, _L, _L, _L, _P, _L, _L, _L, _P, _L, _L, _L, _I, _I, _I, _I, _I, _F, _P, _L, _I, _P],
    "b2_attn_scores_softmax": [_P, _P, _L, _L, _L, _P, _L, _I, _I, _I, _I, _F, _P, _I, _P],
    "b2_attn_scores_bwd": [_P, _L, _L, _L, _P, _L, _L, _L, _P, _P, _P, _L, _I, _I, _I, _I, _F, _I, _P],
    "b2_rowdot": [_P, _L, _P, _L, _L, _P, _L, _I, _I, _I, _P],
    "b2_conv2d_wgrad": [_I, _P, _I, _I, _I, _I, _L, _P, _I, _L, _P, _I, _P],
    "b2_conv2d_wgrad_batch": [_I, _P, _I, _P],
    "b2_gemm_tn": [_P, _L, _L, _L, _P, _L, _L, _L, _P, _L, _L, _L, _I, _I, _I, _I, _I, _F, _I, _I, _P],
    "b2_nchw_to_nhwc_pad": [_P, _P, _I, _I, _I, _I, _I, _I, _P],
    "b2_nhwc_to_nchw": [_P, _L, _P, _I, _I, _I, _I, _I, _P],
    "b2_space_to_depth2": [_P, _L, _P, _I, _I, _I, _I, _I, _P],
    "b2_pack_weight": [_I, _P, _P, _I, _I, _I, _I, _P],
    "b2_gn_stats": [_P, _L, _P, _I, _I, _I, _I, _I, _I, _P],
    "b2_adagn_apply": [_P, _L, _P, _P, _P, _P, _L, _P, _L, _P, _L, _I, _I, _I, _I, _F, _I, _I, _P],
    "b2_adagn_bwd": [_P, _L, _P, _L, _P, _P, _P, _P, _L, _P, _P, _L, _P, _P, _P, _L, _P, _I, _I, _I, _I, _F, _I, _P],
    "b2_act": [_I, _P, _L, _P, _L, _P, _L, _P, _L, _I, _I, _P],
    "b2_f32_act": [_I, _P, _P, _P, _L, _P],
    "b2_softmax_query_axis_bwd": [_P, _P, _P, _I, _I, _I, _L, _F, _I, _P],
    "b2_add": [_P, _L, _P, _L, _P, _L, _L, _I, _I, _P],
    "b2_unpack_weight_grad": [_I, _P, _P, _I, _I, _I, _I, _P],
    "b2_softmax_query_axis": [_P, _P, _I, _I, _I, _L, _I, _P],
    "b2_transpose_batched": [_P, _L, _L, _L, _P, _L, _L, _L, _I, _I, _I, _I, _I, _P],
    "b2_sinusoid_embedding": [_P, _P, _I, _I, _P],
    "b2_small_gemm": [_P, _L, _I, _P, _L, _I, _P, _L, _I, _I, _I, _P, _I, _I, _P],
    "b2_philox_normal": [_P, _L, _U, _U, _L, _P],
    "b2_qsample": [_P, _P, _P, _P, _I, _P, _I, _I, _L, _P],
    "b2_qsample_philox": [_P, _P, _P, _P, _I, _P, _I, _I, _L, _U, _U, _P, _L, _P],
    "b2_mse_loss_grad_philox": [_P, _P, _P, _L, _F, _U, _U, _P, _L, _P],
    "b2_ddim_step": [_P, _P, _P, _P, _P, _L, _F, _F, _F, _F, _F, _I, _P],
    "b2_ddpm_step": [_P, _P, _P, _P, _L, _F, _F, _F, _I, _U, _U, _L, _P],
    "b2_cold_step": [_P, _P, _P, _P, _L, _F, _F, _F, _F, _P],
    "b2_adam_flat": [_P, _P, _P, _P, _L, _D, _D, _F, _F, _F, _F, _P, _P],
    "b2_adam_flat_graph": [_P, _P, _P, _P, _L, _D, _D, _F, _P, _P, _I, _P],
    "b2_adam_advance": [_P, _D, _D, _P],
    "b2_adam_flat_g16": [_P, _P, _P, _P, _L, _D, _D, _F, _F, _F, _F, _P, _P, _P],
    "b2_cast_f32_bf16": [_P, _P, _L, _I, _P],
    "b2_pack_weight_multi": [_P, _I, _L, _I, _P],
    "b2_transpose_weight_cl_multi": [_P, _I, _L, _P],
    "b2_transpose_weight_cl": [_P, _P, _I, _I, _P],
    "b2_transpose_linear_weight": [_P, _P, _I, _I, _P],
    "b2_pack_convt_bf16": [_P, _P, _P, _I, _I, _P],
    "b2_area_resample": [_P, _P, _L, _I, _I, _I, _I, _P],
    "b2_u8_to_image": [_P, _P, _P, _I, _I, _I, _I, _P],
    "b2_flip_images": [_P, _P, _P, _I, _I, _I, _I, _P],
    "b2_image_to_u8": [_P, _P, _I, _I, _I, _I, _F, _F, _P],
    "b2_image_grid_u8": [_P, _P, _I, _I, _I, _I, _I, _I, _I, _F, _F, _P],
    "b2_mse_loss_grad": [_P, _P, _P, _P, _L, _F, _P],
}

_lib = None


class B200Error(RuntimeError):
    pass


def lib():
    """Loads the shared library once; raises (never falls back) when it is absent."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise B200Error(f"{LIB_PATH} not found: build the sm_100a extension first (`make`); there is no CPU fallback")
        handle = ctypes.CDLL(LIB_PATH)
        handle.b2_last_error.restype = ctypes.c_char_p
        handle.b2_version.restype = _I
        for name, args in SIGNATURES.items():
            try:
                fn = getattr(handle, name)
            except AttributeError:
                if os.environ.get("SDM_B200_LIB"):       # an older A/B build (tools/): entry points added since are simply absent
                    continue
                raise B200Error(f"{LIB_PATH} does not export {name}: rebuild the extension (`make`)")
            fn.argtypes = args
            fn.restype = _I
        _lib = handle
    return _lib


LAUNCHES = 0          # kernels launched through the C ABI since import (bench.py reports the delta as gpu_launches)
_HOOK = None          # optional (name, args) -> context manager, used by bench.py to time one kernel family with CUDA events


def set_launch_hook(hook):
    global _HOOK
    _HOOK = hook


_WORKSPACE = {}       # device index -> zeroed split-K workspace registered with the library (b2_set_workspace)
WORKSPACE_BYTES = 128 << 20
_WORKSPACE_USERS = ("b2_conv2d_nhwc", "b2_conv2d_nhwc_colsum", "b2_conv2d_nhwc_dual", "b2_gemm_nt", "b2_gemm_nt_bmn", "b2_conv2d_wgrad", "b2_conv2d_wgrad_batch", "b2_gemm_tn")


def _ensure_workspace():
    """The split-K workspace belongs to the caller (PyTorch's allocator), one per device, registered on first use."""
    dev = torch.cuda.current_device()
    if dev not in _WORKSPACE and os.environ.get("SDM_B200_SPLIT_K", "1") != "0":
        if torch.cuda.is_current_stream_capturing():
            return                              # never allocate inside a capture; eager warm-up registers it first
        buf = torch.zeros(WORKSPACE_BYTES, dtype=torch.uint8, device=f"cuda:{dev}")
        rc = lib().b2_set_workspace(buf.data_ptr(), WORKSPACE_BYTES)
        if rc != 0:
            raise B200Error(f"b2_set_workspace: {lib().b2_last_error().decode()}")
        _WORKSPACE[dev] = buf


def call(name, *args):
    global LAUNCHES
    handle = lib()
    if name in _WORKSPACE_USERS and torch.cuda.current_device() not in _WORKSPACE:
        _ensure_workspace()
    if name != "b2_zero":          # a memset node, not a kernel: not part of the gpu_launches claim
        LAUNCHES += 1
    if _HOOK is not None:
        with _HOOK(name, args):
            rc = getattr(handle, name)(*args)
    else:
        rc = getattr(handle, name)(*args)
    if rc != 0:
        raise B200Error(f"{name}: {handle.b2_last_error().decode()}")


def ptr(t):
    """Device pointer of a CUDA tensor (None -> NULL)."""
    if t is None:
        return None
    if not t.is_cuda:
        raise B200Error("sdm_b200 kernels need CUDA tensors: there is no CPU fallback")
    return t.data_ptr()


def stream():
    return torch.cuda.current_stream().cuda_stream
