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


def run_standalone(module, kind, x, *args):
    """Forward of ONE block of models/custom_layers.py on its own (the reference exposes every class; inside U_Net the
    blocks never run this way).  NCHW fp32 CUDA in / out, inference only (no autograd graph is recorded); channel counts
    must fill whole 128-byte K blocks (multiples of 64 in bf16 mode, 32 in tf32 mode)."""
    from ._lib import call, ptr, stream
    if not x.is_cuda:
        raise B200Error(f"{type(module).__name__}.forward needs CUDA tensors: this build has no CPU path")
    precision = _precision()
    eng = getattr(module, "_b2_engine", None)
    if eng is None or eng.net.precision != precision:
        eng = UNetEngine(_Host(module, precision))
        object.__setattr__(module, "_b2_engine", eng)
    code = ops.TF32 if precision == "tf32" else ops.BF16
    dt = ops.TORCH_DTYPE[code]
    with torch.no_grad():
        if kind == "embedding":
            eng.net.cond_emb = module
            cond = args[0] if args else None
            return eng.embedding(x, cond.to(x.device) if cond is not None else None)
        if x.shape[1] % ops.K_ALIGN[code]:
            raise B200Error(f"standalone {type(module).__name__}: {x.shape[1]} channels is not a multiple of {ops.K_ALIGN[code]}")
        xh = x.detach().float().permute(0, 2, 3, 1).contiguous().to(dt)
        n = xh.shape[0]
        emb = args[0] if args else None
        ctx = {"emb": None, "stats_i": 0}
        mods = eng._adagn_modules()
        if emb is not None and mods:
            emb = emb.detach().float().reshape(-1, emb.shape[-1]).contiguous().to(x.device)
            w_all, b_all, off, total = eng._adagn_table()
            be = emb.shape[0]
            if be not in (1, n):
                raise B200Error(f"embedding batch {be} does not broadcast over image batch {n}")
            s_all = torch.empty((be, total), dtype=torch.float32, device=x.device)
            ops.small_gemm(emb, w_all, be, total, emb.shape[1], emb.shape[1], w_all.shape[1], s_all, total, bias=b_all)
            groups = max(m.group_norm.num_groups for m in mods)
            ctx.update(emb=emb, s_all=s_all, adagn_off=off, s_bstride=(total if be == n else 0),
                       stats=torch.zeros((len(off), n, groups, 2), dtype=torch.float32, device=x.device))
        if kind == "adagn":
            if ctx["emb"] is None:
                raise B200Error("AdaGN.forward needs the embedding")
            gn = module.group_norm
            stats = ctx["stats"][0]
            _, hh, ww, c = xh.shape
            call("b2_gn_stats", ptr(xh), c, ptr(stats), n, hh * ww, c, gn.num_groups, 0, code, stream())
            y = ops.adagn_apply(xh, stats, gn.weight, gn.bias, ctx["s_all"], ctx["s_bstride"], groups=gn.num_groups, eps=gn.eps)
        elif kind == "conv_block":
            y = eng.conv_block(module, xh, ctx)
        elif kind == "residual_block":
            if not isinstance(module.shortcut, torch.nn.Identity):
                raise B200Error("ResidualBlock with a 1x1 shortcut is unreachable from U_Net and not implemented (SURVEY Q5)")
            y = eng.residual_block(module, xh, ctx)
        elif kind == "attention":
            y = eng.attention(module, xh)
        elif kind in ("upsample", "downsample"):
            y = eng.resample(module, xh)
        elif kind == "unet_block":
            y = eng.unet_block(module, xh, ctx, None)
        else:
            raise B200Error(f"run_standalone: unknown block kind {kind}")
        return ops.nhwc_to_nchw(y)


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
