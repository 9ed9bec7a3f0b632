"""Fused loss: mean squared error with its gradient produced in the same pass (train_diffusion.py:350)."""
import torch

from ._lib import call, ptr, stream


class _MSE(torch.autograd.Function):
    @staticmethod
    def forward(ctx, pred, target):
        p = pred.contiguous().float()
        t = target.contiguous().float()
        grad = torch.empty_like(p)
        loss = torch.empty((), dtype=torch.float32, device=p.device)
        call("b2_mse_loss_grad", ptr(p), ptr(t), ptr(grad), ptr(loss), p.numel(), 1.0, stream())
        ctx.save_for_backward(grad)
        return loss

    @staticmethod
    def backward(ctx, g):
        (grad,) = ctx.saved_tensors
        return grad * g, None


class _MSEPhilox(torch.autograd.Function):
    @staticmethod
    def forward(ctx, pred, noise):
        loss, grad = mse_loss_philox(pred, noise)
        ctx.save_for_backward(grad)
        return loss

    @staticmethod
    def backward(ctx, g):
        (grad,) = ctx.saved_tensors
        return grad * g, None


def mse_loss(pred, target):
    """F.mse_loss(pred, target) (mean reduction) on CUDA tensors: one kernel yields the loss and d(loss)/d(pred).
    `target` may be a degraders.PhiloxNoise: the eps that `forward_philox` drew, re-generated inside the kernel."""
    if not torch.is_tensor(target):
        return _MSEPhilox.apply(pred, target)
    return _MSE.apply(pred, target)


def mse_loss_philox(pred, noise):
    """F.mse_loss(pred, eps) where eps is the in-kernel Philox draw of `degrader.forward_philox(..., noise)`: the target is
    re-generated inside the loss kernel, never stored.  Returns (loss, d loss / d pred); no autograd node."""
    p = pred.contiguous().float()
    grad = torch.empty_like(p)
    loss = torch.empty((), dtype=torch.float32, device=p.device)
    call("b2_mse_loss_grad_philox", ptr(p), ptr(grad), ptr(loss), p.numel(), 1.0, noise.seed, noise.offset, ptr(noise.offset_dev),
         noise.first_elem, stream())
    return loss, grad


def area_resize(x, size):
    """F.interpolate(x, size=size, mode="area") on an fp32 NCHW CUDA tensor (the SR trainers' low-resolution
    conditioning image, train_SR_diffusion.py:321-328): one kernel, adaptive-average-pool semantics."""
    if not x.is_cuda:
        from ._lib import B200Error
        raise B200Error("area_resize needs a CUDA tensor: this build has no CPU path")
    oh, ow = (size, size) if isinstance(size, int) else size
    xc = x.contiguous().float()
    n, c, h, w = xc.shape
    out = torch.empty((n, c, oh, ow), dtype=torch.float32, device=x.device)
    call("b2_area_resample", ptr(xc), ptr(out), n * c, h, w, oh, ow, stream())
    return out
