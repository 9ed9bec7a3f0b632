"""Forward diffusion process q(x_t | x_0) with the reference's classes and signatures (degraders.py:10-104).

`get_timestep_params` keeps returning small torch tensors (it is schedule bookkeeping); the image-sized arithmetic of
`forward` is ONE fused sm_100a kernel (b2_qsample) instead of ~6 ATen launches.  `host_params` is an addition used by
the samplers: the per-step scalars are pure functions of the schedule, so they are computed once on the host in fp32
with exactly the reference's expressions instead of synchronising the GPU every step.
"""
import math

import torch
import torch.nn as nn

from b200._lib import B200Error, call, ptr, stream


def _qsample(img, steps, eps, abar_table, max_step):
    if not img.is_cuda:
        raise B200Error("degrader.forward needs CUDA tensors: this build has no CPU path")
    if abar_table is not None:
        _check_steps(steps, max_step)
    if eps is None:
        eps = torch.randn_like(img)
    img_c = img.contiguous().float()
    eps_c = eps.contiguous().float()
    steps_c = steps.to(device=img.device, dtype=torch.int64).contiguous()
    out = torch.empty_like(img_c)
    n = img_c.shape[0]
    call("b2_qsample", ptr(img_c), ptr(eps_c), ptr(out), ptr(steps_c), steps_c.numel(), ptr(abar_table), int(max_step), n,
         img_c.numel() // n, stream())
    return out


def _check_steps(steps, max_step):
    """The reference's `torch.gather` raises on t < 0 or t > max_noise_step (degraders.py:44-49).  Timesteps that live on the
    host are validated here; device-resident ones (the trainer's randint) are checked inside the kernel, which poisons the
    image with NaN instead of reading outside the table."""
    if torch.is_tensor(steps) and steps.is_cuda:
        return
    idx = torch.as_tensor(steps, dtype=torch.int64).reshape(-1)
    if idx.numel() and (int(idx.min()) < 0 or int(idx.max()) > int(max_step)):
        raise IndexError(f"timestep out of range [0, {int(max_step)}]: {idx.tolist()}")


class PhiloxNoise:
    """Where the in-kernel eps of `forward_philox` comes from: Philox4x32-10 keyed on (seed, offset, first_elem + element
    index).  `offset` is the draw number (the optimisation step); `offset_dev` (a 1-element fp32 CUDA tensor, e.g. the fused
    optimiser's device-side step count) overrides it so that CUDA-graph replays draw fresh noise; `first_elem` is the global
    index of this process's first element (rank * N * C * H * W under data parallelism)."""

    def __init__(self, seed, offset=0, offset_dev=None, first_elem=0):
        self.seed, self.offset, self.offset_dev, self.first_elem = int(seed), int(offset), offset_dev, int(first_elem)


def _qsample_philox(img, steps, noise, abar_table, max_step, eps_out=None):
    if not img.is_cuda:
        raise B200Error("degrader.forward_philox needs CUDA tensors: this build has no CPU path")
    if abar_table is not None:
        _check_steps(steps, max_step)
    img_c = img.contiguous().float()
    steps_c = steps.to(device=img.device, dtype=torch.int64).contiguous()
    out = torch.empty_like(img_c)
    n = img_c.shape[0]
    call("b2_qsample_philox", ptr(img_c), ptr(out), ptr(eps_out), ptr(steps_c), steps_c.numel(), ptr(abar_table), int(max_step),
         n, img_c.numel() // n, noise.seed, noise.offset, ptr(noise.offset_dev), noise.first_elem, stream())
    return out


class NoiseDegradation(nn.Module):
    """Linear beta schedule with T+1 table entries indexed directly by t (degraders.py:26-42)."""

    def __init__(self, beta_1, beta_T, max_noise_step, device="cpu"):
        super().__init__()
        self.beta_1 = beta_1
        self.beta_T = beta_T
        self.max_noise_step = max_noise_step
        # Built on the host (deterministic, equal to the CPU reference bit for bit), then moved.
        beta = torch.linspace(start=self.beta_1, end=self.beta_T, steps=int(self.max_noise_step + 1))
        self._host_beta = beta
        self._host_alpha = 1 - beta
        self._host_abar = torch.cumprod(self._host_alpha, dim=0)
        self.beta = self._host_beta.to(device)
        self.alpha = self._host_alpha.to(device)
        self.alpha_cumulative_prod = self._host_abar.to(device)

    def get_timestep_params(self, step):
        step = step.to(self.beta.device)
        return self.beta[step], self.alpha[step], self.alpha_cumulative_prod[step]

    def host_params(self, step):
        """(beta, alpha, alpha_bar) of integer step(s) as fp32 CPU tensors."""
        _check_steps(step, self.max_noise_step)
        idx = torch.as_tensor(step, dtype=torch.int64).reshape(-1)
        return self._host_beta[idx], self._host_alpha[idx], self._host_abar[idx]

    def forward(self, img, steps, eps=None):
        if self.alpha_cumulative_prod.device != img.device:
            self.beta, self.alpha = self.beta.to(img.device), self.alpha.to(img.device)
            self.alpha_cumulative_prod = self.alpha_cumulative_prod.to(img.device)
        return _qsample(img, steps, eps, self.alpha_cumulative_prod, self.max_noise_step)

    def forward_philox(self, img, steps, noise, eps_out=None):
        """`forward` with eps drawn inside the kernel (PhiloxNoise) instead of by a separate `randn_like` launch."""
        if self.alpha_cumulative_prod.device != img.device:
            self.beta, self.alpha = self.beta.to(img.device), self.alpha.to(img.device)
            self.alpha_cumulative_prod = self.alpha_cumulative_prod.to(img.device)
        return _qsample_philox(img, steps, noise, self.alpha_cumulative_prod, self.max_noise_step, eps_out)


class CosineNoiseDegradation(nn.Module):
    """Cosine schedule evaluated in closed form from t, beta clipped to [0.001, 0.999] (degraders.py:63-94)."""

    def __init__(self, max_noise_step=1000):
        super().__init__()
        self.max_noise_step = max_noise_step
        self.offset = 0.008

    def compute_alpha_bar(self, steps):
        def f(s):
            return torch.cos((((s / self.max_noise_step) + self.offset) / (1 + self.offset)) * (math.pi / 2)) ** 2
        return f(steps) / f(torch.zeros_like(steps))

    def get_timestep_params(self, steps):
        alpha_bar = self.compute_alpha_bar(steps)
        beta = torch.clip(1 - (alpha_bar / self.compute_alpha_bar(steps - 1)), min=0.001, max=0.999)
        return beta, 1 - beta, alpha_bar

    def host_params(self, step):
        return self.get_timestep_params(torch.as_tensor(step, dtype=torch.int64).reshape(-1).cpu())

    def forward(self, img, steps, eps=None):
        return _qsample(img, steps, eps, None, self.max_noise_step)

    def forward_philox(self, img, steps, noise, eps_out=None):
        """`forward` with eps drawn inside the kernel (PhiloxNoise) instead of by a separate `randn_like` launch."""
        return _qsample_philox(img, steps, noise, None, self.max_noise_step, eps_out)
