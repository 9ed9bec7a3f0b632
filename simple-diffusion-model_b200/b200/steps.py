"""One optimisation step of each reference trainer, as the sequence of fused kernels it becomes here.  The entry points
(train_*.py) and bench.py call these; they keep the reference's order of operations:
    noise / t are drawn by the caller -> q-sample -> U-Net -> MSE -> backward -> (all-reduce) -> Adam.
"""
import torch

from .functional import mse_loss


def _noised(degrader, x0, t, eps):
    """q(x_t | x_0): eps is a tensor (the reference's torch.randn_like draw) or a degraders.PhiloxNoise (drawn in-kernel)."""
    if torch.is_tensor(eps) or eps is None:
        return degrader(x0, t, eps)
    return degrader.forward_philox(x0, t, eps)


def eps_prediction_step(net, degrader, optimizer, x0, t, eps, labels=None, cond_img=None):
    """train_diffusion.py:333-364 (target = eps) and train_doodle_diffusion.py:304-320 (cond image concatenated)."""
    x_t = _noised(degrader, x0, t, eps)
    inp = torch.cat((x_t, cond_img), dim=1) if cond_img is not None else x_t
    pred = net(inp, t, labels)
    loss = mse_loss(pred, eps)
    optimizer.zero_grad(set_to_none=True)
    loss.backward()
    optimizer.step()
    return loss


def x0_prediction_step(net, degrader, optimizer, x0, t, eps, labels=None, cond_img=None, target=None):
    """train_noise_cold_diffusion.py:330-352 (target = x0) / train_SR_diffusion.py:366-380 (target = x0 - lr image)."""
    x_t = _noised(degrader, x0, t, eps)
    inp = torch.cat((x_t, cond_img), dim=1) if cond_img is not None else x_t
    pred = net(inp, t, labels)
    loss = mse_loss(pred, x0 if target is None else target)
    optimizer.zero_grad(set_to_none=True)
    loss.backward()
    optimizer.step()
    return loss


def scaled_step(net, degrader, optimizer, scaler, x0, t, eps, labels=None, cond_img=None, target=None, kind="eps"):
    """The reference's mixed-precision step (train_diffusion.py:130, 333-364): `scaler.scale(loss).backward();
    scaler.step(optim); scaler.update()` with a torch GradScaler.  The reference needs the scaler because its autocast dtype
    is fp16; this build computes in bf16 (fp32 range), so scaling is numerically a no-op -- but a training loop written against
    the reference keeps working unchanged: the scaled gradient flows through the hand-written backward into the flat fp32
    gradient buffer, `scaler.step` un-scales it in place, skips the update when it finds an inf / nan, and calls FusedAdam.
    (Do not combine with DataParallel.attach_optimizer: bucket-wise updates would run before the un-scale.)"""
    x_t = _noised(degrader, x0, t, eps)
    inp = torch.cat((x_t, cond_img), dim=1) if cond_img is not None else x_t
    pred = net(inp, t, labels)
    want = eps if kind == "eps" else (x0 if target is None else target)
    loss = mse_loss(pred, want)
    optimizer.zero_grad(set_to_none=True)
    scaler.scale(loss).backward()
    scaler.step(optimizer)
    scaler.update()
    return loss
