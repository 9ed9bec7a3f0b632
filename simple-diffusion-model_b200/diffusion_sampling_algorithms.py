"""Reverse-process samplers with the reference's names and signatures (diffusion_sampling_algorithms.py:5-217).

Per step the reference launches 8-15 ATen kernels plus a host->device copy of `t` and (for the linear table) device
gathers; here every per-step scalar is precomputed on the host (it depends on the schedule only), all timesteps live in
one device tensor, and each update is ONE fused kernel (b2_ddim_step / b2_ddpm_step / b2_cold_step).

RNG: with `REFERENCE_RNG = True` (default) the samplers draw from torch's generator exactly where the reference does
(DDIM draws a `randn_like` per step although eta = 0), so generator state and outputs match the reference seed for seed.
Setting `PHILOX_SEED` to an int switches DDPM to in-kernel Philox keyed on global element indices (sharded == unsharded).
"""
import torch

from b200._lib import B200Error, call, ptr, stream
from utils.utils import printProgressBar

REFERENCE_RNG = True
PHILOX_SEED = None
_SHARD = None           # (lo, hi, total): this process samples images [lo, hi) of a `total`-image job (generate_* under torchrun)


def set_shard(lo=None, hi=None, total=None):
    """Declares that the x_t handed to the samplers holds images [lo, hi) of a `total`-image job.  Per-step noise is then
    drawn for the WHOLE job from the (identically seeded) generator and sliced -- or, with PHILOX_SEED, keyed on the global
    element index -- so that the union of the shards equals the single-process result.  `set_shard()` clears it."""
    global _SHARD
    _SHARD = None if lo is None or total is None or (lo == 0 and hi == total) else (int(lo), int(hi), int(total))


def _step_noise(x_t):
    """`torch.randn(x_t.shape)` of the reference (:43, :129), shard-consistent (see set_shard)."""
    if _SHARD is None or x_t.shape[0] != _SHARD[1] - _SHARD[0]:
        return torch.randn(x_t.shape, device=x_t.device)
    lo, hi, total = _SHARD
    return torch.randn((total,) + tuple(x_t.shape[1:]), device=x_t.device)[lo:hi].contiguous()


def _first_elem(x_t):
    if _SHARD is None or x_t.shape[0] != _SHARD[1] - _SHARD[0]:
        return 0
    return _SHARD[0] * (x_t.numel() // x_t.shape[0])


def skip_schedule(min_noise, max_noise, step_size):
    """max, max-step, ... plus min_noise if the stride missed it (reference :79-82, :164-168).  Pure Python ints."""
    steps = list(range(max_noise, min_noise - 1, -step_size))
    if min_noise not in steps:
        steps = steps + [min_noise]
    return steps


def _prep(x_t, cond_img, device):
    if not x_t.is_cuda:
        raise B200Error("samplers need CUDA tensors: this build has no CPU path")
    x_t = x_t.contiguous().float()
    if cond_img is not None:
        cond_img = cond_img.to(x_t.device).float()
    return x_t, cond_img


def _net_input(x_t, cond_img):
    return torch.cat((x_t, cond_img), dim=1) if cond_img is not None else x_t


def _f(t):
    return float(t.reshape(-1)[0])


def _eval(diffusion_net, x, t, labels):
    """One network evaluation.  A U_Net with CUDA graphs enabled hands out the graph's own output buffer (no per-step clone):
    every caller below consumes it with the update kernel -- stream-ordered before the next replay overwrites it."""
    graphed = getattr(diffusion_net, "_graphed", None)
    if graphed is not None and x.is_cuda and not torch.is_grad_enabled():
        return graphed(x, t, labels)
    return diffusion_net(x, t, labels).contiguous()


def ddpm_sampling(diffusion_net, noise_degradation, x_t, min_noise=1, max_noise=1_000, cond_img=None, labels_tensor=None,
                  device="cpu", log=print):
    x_t, cond_img = _prep(x_t, cond_img, device)
    diffusion_net.eval()
    steps = list(range(max_noise, min_noise - 1, -1))
    ts = torch.tensor(steps, device=x_t.device, dtype=torch.int64)
    n_elem = x_t.numel()
    with torch.no_grad():
        for i, step in enumerate(steps):
            beta, alpha, abar = noise_degradation.host_params(step)
            eps_hat = _eval(diffusion_net, _net_input(x_t, cond_img), ts[i:i + 1], labels_tensor)
            sigma = beta ** 0.5
            scale_1 = 1 / (alpha ** 0.5)
            scale_2 = (1 - alpha) / ((1 - abar) ** 0.5)
            z, use_philox = None, 0
            if step > 1:
                if PHILOX_SEED is not None:
                    use_philox = 1
                else:
                    z = _step_noise(x_t)
            out = torch.empty_like(x_t)
            call("b2_ddpm_step", ptr(x_t), ptr(eps_hat), ptr(z), ptr(out), n_elem, _f(scale_1), _f(scale_2), _f(sigma),
                 use_philox, int(PHILOX_SEED or 0), int(step), _first_elem(x_t), stream())
            x_t = out
            printProgressBar(iteration=max_noise - step, total=max_noise - min_noise, prefix='Iterations:',
                             suffix='Complete', length=50, log=log)
    return x_t


def ddim_sampling(diffusion_net, noise_degradation, x_t, min_noise=1, max_noise=1_000, cond_img=None, labels_tensor=None,
                  ddim_step_size=10, device="cpu", log=print):
    x_t, cond_img = _prep(x_t, cond_img, device)
    diffusion_net.eval()
    steps = skip_schedule(min_noise, max_noise, ddim_step_size)
    ts = torch.tensor(steps, device=x_t.device, dtype=torch.int64)
    eta = 0.0
    n_elem = x_t.numel()
    x0_approx = None
    with torch.no_grad():
        for count, step in enumerate(steps):
            eps_hat = _eval(diffusion_net, _net_input(x_t, cond_img), ts[count:count + 1], labels_tensor)
            abar_t = noise_degradation.host_params(step)[2]
            c_scale = 1 / abar_t ** 0.5
            c_s = (1 - abar_t) ** 0.5
            last = count == len(steps) - 1
            if not last:
                abar_n = noise_degradation.host_params(steps[count + 1])[2]
                sigma = eta * (((1 - abar_n) / (1 - abar_t)) ** 0.5 * (1 - (abar_t / abar_n)) ** 0.5)
                noise = _step_noise(x_t) if REFERENCE_RNG else None           # drawn by the reference even though eta = 0
                out = torch.empty_like(x_t)
                call("b2_ddim_step", ptr(x_t), ptr(eps_hat), ptr(noise) if _f(sigma) != 0.0 else None, ptr(out), None, n_elem,
                     _f(c_scale), _f(c_s), _f(abar_n ** 0.5), _f((1 - abar_n - sigma ** 2) ** 0.5), _f(sigma), 0, stream())
                x_t = out
                printProgressBar(iteration=max_noise - step, total=max_noise - min_noise, prefix='Iterations:',
                                 suffix='Complete', length=50, log=log)
            else:
                x0_approx = torch.empty_like(x_t)
                call("b2_ddim_step", ptr(x_t), ptr(eps_hat), None, None, ptr(x0_approx), n_elem, _f(c_scale), _f(c_s), 0.0, 0.0,
                     0.0, 1, stream())
    if steps[-1] == 1:
        return x0_approx
    return x_t


def cold_diffusion_sampling(diffusion_net, noise_degradation, x_t, noise, min_noise=1, max_noise=1_000, cond_img=None,
                            labels_tensor=None, skip_step_size=10, device="cpu", log=print):
    x_t, cond_img = _prep(x_t, cond_img, device)
    noise = noise.to(x_t.device).contiguous().float()
    diffusion_net.eval()
    steps = skip_schedule(min_noise, max_noise, skip_step_size)
    ts = torch.tensor(steps, device=x_t.device, dtype=torch.int64)
    n_elem = x_t.numel()
    x0_hat = None
    with torch.no_grad():
        for count, step in enumerate(steps):
            x0_hat = _eval(diffusion_net, _net_input(x_t, cond_img), ts[count:count + 1], labels_tensor)
            if count < len(steps) - 1:
                abar_t = noise_degradation.host_params(step)[2]
                abar_n = noise_degradation.host_params(steps[count + 1])[2]
                out = torch.empty_like(x_t)
                call("b2_cold_step", ptr(x_t), ptr(x0_hat), ptr(noise), ptr(out), n_elem, _f(abar_t ** 0.5),
                     _f((1 - abar_t) ** 0.5), _f(abar_n ** 0.5), _f((1 - abar_n) ** 0.5), stream())
                x_t = out
                printProgressBar(iteration=max_noise - step, total=max_noise - min_noise, prefix='Iterations:',
                                 suffix='Complete', length=50, log=log)
    if getattr(diffusion_net, "_graphed", None) is not None:
        x0_hat = x0_hat.clone()                  # the caller keeps it: detach it from the graph's output buffer
    return x0_hat
