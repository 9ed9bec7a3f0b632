"""One optimisation step of each reference trainer, as the sequence of fused kernels it becomes here.  The entry points
(train_*.py) and bench.py call these; they keep the reference's order of operations:
    noise / t are drawn by the caller -> q-sample -> U-Net -> MSE -> backward -> (all-reduce) -> Adam.
"""
import torch

from .functional import mse_loss


def eps_prediction_step(net, degrader, optimizer, x0, t, eps, labels=None, cond_img=None):
    """train_diffusion.py:333-364 (target = eps) and train_doodle_diffusion.py:304-320 (cond image concatenated)."""
    x_t = degrader(x0, t, eps)
    inp = torch.cat((x_t, cond_img), dim=1) if cond_img is not None else x_t
    pred = net(inp, t, labels)
    loss = mse_loss(pred, eps)
    optimizer.zero_grad(set_to_none=True)
    loss.backward()
    optimizer.step()
    return loss


def x0_prediction_step(net, degrader, optimizer, x0, t, eps, labels=None, cond_img=None, target=None):
    """train_noise_cold_diffusion.py:330-352 (target = x0) / train_SR_diffusion.py:366-380 (target = x0 - lr image)."""
    x_t = degrader(x0, t, eps)
    inp = torch.cat((x_t, cond_img), dim=1) if cond_img is not None else x_t
    pred = net(inp, t, labels)
    loss = mse_loss(pred, x0 if target is None else target)
    optimizer.zero_grad(set_to_none=True)
    loss.backward()
    optimizer.step()
    return loss
