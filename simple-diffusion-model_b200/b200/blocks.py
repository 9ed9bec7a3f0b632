"""Standalone execution of single blocks (the reference exposes every class of models/custom_layers.py on its own).
NCHW fp32 CUDA tensors in and out; inside, the block runs on the same engine methods as the full U-Net."""
import torch

from . import ops
from ._lib import B200Error
from .engine import UNetEngine
from .train_engine import UNetTrainEngine


class _Host:
    """Minimal stand-in for a U_Net so a UNetEngine can serve one block."""

    def __init__(self, module, precision):
        self.module = module
        self.precision = precision
        self.cond_emb = None

    def modules(self):
        return self.module.modules()

    def parameters(self):
        return self.module.parameters()

    def named_parameters(self):
        return self.module.named_parameters()


def _precision():
    import os
    return os.environ.get("SDM_B200_PRECISION", "bf16")


def swish_standalone(x):
    if not x.is_cuda:
        raise B200Error("Swish needs a CUDA tensor: this build has no CPU path")
    from ._lib import call, ptr, stream
    xf = x.contiguous().float()
    out = torch.empty_like(xf)
    call("b2_f32_act", 0, None, ptr(xf), ptr(out), xf.numel(), stream())
    return out.to(x.dtype)


def run_block_train(module, kind, x, dout, precision):
    """Forward + backward of one block (used by the op-level tests): returns (out, dx, {param name: grad})."""
    host = _Host(module, precision)
    eng = UNetTrainEngine(host)
    code = ops.TF32 if precision == "tf32" else ops.BF16
    dt = ops.TORCH_DTYPE[code]
    eng.layout = None
    lay = eng.grad_layout(x.device)
    lay.flat.zero_()
    xh = x.permute(0, 2, 3, 1).contiguous().to(dt)
    if kind == "attention":
        saved = {}
        out = eng.attention(module, xh, save=saved)
        dh = dout.permute(0, 2, 3, 1).contiguous().to(dt)
        dx = eng._bwd_attention(module, xh, saved, dh)
    else:
        raise B200Error(f"run_block_train: unsupported block kind {kind}")
    grads = {n: lay.view(p).clone() for n, p in module.named_parameters() if id(p) in lay.views}
    return ops.nhwc_to_nchw(out), ops.nhwc_to_nchw(dx), grads
