"""CPU oracle for the diffusion hot path -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.

A plain-PyTorch fp32, CPU, functional restatement of the reference algorithm (Vinmwaura/Simple-Diffusion-Model).
Only `tests/`, `__graft_entry__.smoke()` and `bench.py`'s cpu_baseline / `--impl reference` legs may import this
module, and only as the checker / CPU baseline; the product path (simple-diffusion-model_b200/) never does.

Parity pin: the reference publishes no golden vectors for this path (its only test checks an output shape,
tests/test_u_net_model.py:21-23), so this restatement is pinned against outputs of the reference itself,
imported unmodified from /root/reference by `tests/golden/make_golden.py` (committed with the fixtures it wrote,
`tests/golden/*.pt`); `tests/test_oracle_golden.py` replays them.

Each function cites the reference file:line it follows.  Weights are consumed as a reference-format
``state_dict`` (same keys/shapes as models/U_Net.py), activations are NCHW fp32.
"""
import math

import torch
import torch.nn.functional as F


# ----------------------------------------------------------------------------------------------- model pieces
def swish(x):
    """models/custom_layers.py:18-20."""
    return x * torch.sigmoid(x)


def _lin(sd, key, x):
    return F.linear(x, sd[key + ".weight"], sd[key + ".bias"])


def sinusoid(t, time_dim):
    """models/custom_layers.py:84-90: [sin(t f_k), cos(t f_k)], f_k = exp(-k ln(1e4)/(half-1))."""
    half = time_dim // 2
    k = torch.arange(half, dtype=torch.float32, device=t.device)   # device-agnostic: the GPU tests also run this restatement through torch eager
    freq = torch.exp(k * -(math.log(10_000) / (half - 1)))
    arg = t[:, None] * freq[None, :]
    return torch.cat((arg.sin(), arg.cos()), dim=1)


def _mlp4(sd, prefix, x):
    """Linear/Swish x3 + Linear (models/custom_layers.py:59-77)."""
    for i in (0, 2, 4):
        x = swish(_lin(sd, f"{prefix}.{i}", x))
    return _lin(sd, f"{prefix}.6", x)


def cond_embedding(sd, t, cond):
    """models/custom_layers.py:82-98."""
    time_dim = sd["cond_emb.time_layer.0.weight"].shape[0]
    emb = _mlp4(sd, "cond_emb.time_layer", sinusoid(t, time_dim))
    if "cond_emb.cond_layer.0.weight" in sd:
        emb = emb + _mlp4(sd, "cond_emb.cond_layer", cond)
    return emb


def adagn(sd, prefix, x, emb, groups=32):
    """models/custom_layers.py:35-45.  The shift re-uses the *scale* Linear: out = s*GN(x) + s."""
    gn = F.group_norm(x, groups, sd[prefix + ".group_norm.weight"], sd[prefix + ".group_norm.bias"], eps=1e-5)
    s = _lin(sd, prefix + ".y_scale", emb)[:, :, None, None]
    return s * gn + s


def conv_block(sd, prefix, x, emb, act=True):
    """models/custom_layers.py:240-245: Conv3x3 -> Swish -> AdaGN (post-activation norm)."""
    x = F.conv2d(x, sd[prefix + ".conv_layer.0.weight"], sd[prefix + ".conv_layer.0.bias"], padding=1)
    if act:
        x = swish(x)
    if emb is not None and (prefix + ".adagn.y_scale.weight") in sd:
        x = adagn(sd, prefix + ".adagn", x, emb)
    return x


def residual_block(sd, prefix, x, emb):
    """models/custom_layers.py:282-287 (shortcut is Identity whenever reached from U_Net)."""
    h = conv_block(sd, prefix + ".conv_block_1", x, emb)
    h = conv_block(sd, prefix + ".conv_block_2", h, emb)
    if (prefix + ".shortcut.weight") in sd:
        x = F.conv2d(x, sd[prefix + ".shortcut.weight"], sd[prefix + ".shortcut.bias"])
    return h + x


def attention_block(sd, prefix, x, heads):
    """models/custom_layers.py:127-163.  Softmax runs over the QUERY axis; the GroupNorm is never applied."""
    n, c, hh, ww = x.shape
    seq = x.reshape(n, c, hh * ww).permute(0, 2, 1)
    w_p = sd[prefix + ".projection.weight"]
    d_k = w_p.shape[0] // (3 * heads)
    qkv = F.linear(seq, w_p, sd[prefix + ".projection.bias"]).reshape(n, hh * ww, heads, 3 * d_k)
    q, k, v = qkv[..., :d_k], qkv[..., d_k:2 * d_k], qkv[..., 2 * d_k:]
    logits = torch.einsum("bihd,bjhd->bijh", q, k) * (d_k ** -0.5)
    p = torch.softmax(logits, dim=1)
    o = torch.einsum("bijh,bjhd->bihd", p, v).reshape(n, hh * ww, heads * d_k)
    o = _lin(sd, prefix + ".output", o) + seq
    return o.permute(0, 2, 1).reshape(n, c, hh, ww)


def _count(sd, prefix):
    idx = set()
    for key in sd:
        if key.startswith(prefix):
            idx.add(int(key[len(prefix):].split(".")[0]))
    return len(idx)


def unet_block(sd, prefix, x, emb, heads, up):
    """models/custom_layers.py:336-341 + sampler at :169-207."""
    for i in range(_count(sd, prefix + ".res_layers.")):
        x = residual_block(sd, f"{prefix}.res_layers.{i}", x, emb)
        if f"{prefix}.attn_layers.{i}.projection.weight" in sd:
            x = attention_block(sd, f"{prefix}.attn_layers.{i}", x, heads)
    w, b = sd[prefix + ".out_layer.conv_layer.0.weight"], sd[prefix + ".out_layer.conv_layer.0.bias"]
    if up:
        x = F.conv_transpose2d(x, w, b, stride=2, padding=1)
    else:
        x = F.conv2d(x, w, b, stride=2, padding=1)
    return swish(x)


def unet_forward(sd, x, t, cond=None, heads=1, image_recon=False):
    """models/U_Net.py:147-173."""
    emb = cond_embedding(sd, t, cond) if "cond_emb.time_layer.0.weight" in sd else None
    x = conv_block(sd, "in_layer.0", x, None)
    x = conv_block(sd, "in_layer.1", x, None)
    skips = []
    n_layers = _count(sd, "down_layers.")
    for i in range(n_layers):
        x = unet_block(sd, f"down_layers.{i}", x, emb, heads, up=False)
        skips.append(x)
    x = conv_block(sd, "middle_layer.0", x, None)
    x = conv_block(sd, "middle_layer.1", x, None)
    for i in range(n_layers):
        x = torch.cat((x, skips.pop()), dim=1)
        x = unet_block(sd, f"up_layers.{i}", x, emb, heads, up=True)
    x = conv_block(sd, "out_layers.0", x, None)
    x = conv_block(sd, "out_layers.1", x, None, act=False)
    return torch.tanh(x) if image_recon else x


# ----------------------------------------------------------------------------------------------- schedules
def linear_tables(beta_1, beta_t, max_step):
    """degraders.py:26-42: T+1 entries, indexed directly by t."""
    beta = torch.linspace(start=beta_1, end=beta_t, steps=int(max_step + 1))
    alpha = 1 - beta
    return beta, alpha, torch.cumprod(alpha, dim=0)


def cosine_alpha_bar(steps, max_step, offset=0.008):
    """degraders.py:70-82."""
    f = lambda s: torch.cos((((s / max_step) + offset) / (1 + offset)) * (math.pi / 2)) ** 2
    return f(steps) / f(torch.zeros_like(steps))


def timestep_params(sched, steps):
    """degraders.py:44-49 (linear) / :84-94 (cosine).  sched = ("linear", b1, bT, T) | ("cosine", T)."""
    if sched[0] == "linear":
        beta, alpha, abar = linear_tables(*sched[1:])
        return beta[steps], alpha[steps], abar[steps]
    t_max = sched[1]
    abar = cosine_alpha_bar(steps, t_max)
    beta = torch.clip(1 - abar / cosine_alpha_bar(steps - 1, t_max), min=0.001, max=0.999)
    return beta, 1 - beta, abar


def q_sample(sched, img, steps, eps):
    """degraders.py:51-59 / :96-104."""
    abar = timestep_params(sched, steps)[2][:, None, None, None]
    return abar ** 0.5 * img + (1 - abar) ** 0.5 * eps


def skip_schedule(min_noise, max_noise, step):
    """diffusion_sampling_algorithms.py:79-82 / :164-168."""
    steps = list(range(max_noise, min_noise - 1, -step))
    if min_noise not in steps:
        steps.append(min_noise)
    return steps


# ----------------------------------------------------------------------------------------------- sampler updates
def ddpm_update(sched, x_t, eps_hat, step, z):
    """diffusion_sampling_algorithms.py:48-55; z must be 0 when step == 1 (:42-45)."""
    beta, alpha, abar = timestep_params(sched, torch.tensor([step]))
    return (1 / alpha ** 0.5) * (x_t - ((1 - alpha) / (1 - abar) ** 0.5) * eps_hat) + beta ** 0.5 * z


def ddim_x0(sched, x_t, eps_hat, step):
    """diffusion_sampling_algorithms.py:107-112."""
    abar = timestep_params(sched, torch.tensor([step]))[2]
    return (1 / abar ** 0.5) * (x_t - (1 - abar) ** 0.5 * eps_hat)


def ddim_update(sched, x_t, eps_hat, step, step_next, eps, eta=0.0):
    """diffusion_sampling_algorithms.py:107-136."""
    abar = timestep_params(sched, torch.tensor([step]))[2]
    abar_n = timestep_params(sched, torch.tensor([step_next]))[2]
    x0 = ddim_x0(sched, x_t, eps_hat, step)
    sigma = eta * (((1 - abar_n) / (1 - abar)) ** 0.5 * (1 - abar / abar_n) ** 0.5)
    return abar_n ** 0.5 * x0 + (1 - abar_n - sigma ** 2) ** 0.5 * eps_hat + sigma * eps


def cold_update(sched, x_t, x0_hat, step, step_next, noise):
    """diffusion_sampling_algorithms.py:193-208."""
    return x_t - q_sample(sched, x0_hat, torch.tensor([step]), noise) + q_sample(sched, x0_hat, torch.tensor([step_next]), noise)


def ddim_sample(net, sched, x_t, min_noise, max_noise, step, cond_img=None, labels=None, draw=torch.randn_like):
    """diffusion_sampling_algorithms.py:66-148 with `net(x, t, labels)` any callable."""
    steps = skip_schedule(min_noise, max_noise, step)
    x0 = None
    for i, s in enumerate(steps):
        t = torch.tensor([s])
        inp = torch.cat((x_t, cond_img), dim=1) if cond_img is not None else x_t
        eps_hat = net(inp, t, labels)
        x0 = ddim_x0(sched, x_t, eps_hat, s)
        if i < len(steps) - 1:
            x_t = ddim_update(sched, x_t, eps_hat, s, steps[i + 1], draw(x0))
    return x0 if steps[-1] == 1 else x_t


def ddpm_sample(net, sched, x_t, min_noise, max_noise, cond_img=None, labels=None, draw=torch.randn):
    """diffusion_sampling_algorithms.py:5-64."""
    for s in range(max_noise, min_noise - 1, -1):
        t = torch.tensor([s])
        inp = torch.cat((x_t, cond_img), dim=1) if cond_img is not None else x_t
        eps_hat = net(inp, t, labels)
        z = draw(x_t.shape) if s > 1 else 0
        x_t = ddpm_update(sched, x_t, eps_hat, s, z)
    return x_t


def cold_sample(net, sched, x_t, noise, min_noise, max_noise, step, cond_img=None, labels=None):
    """diffusion_sampling_algorithms.py:150-217."""
    steps = skip_schedule(min_noise, max_noise, step)
    x0 = None
    for i, s in enumerate(steps):
        t = torch.tensor([s])
        inp = torch.cat((x_t, cond_img), dim=1) if cond_img is not None else x_t
        x0 = net(inp, t, labels)
        if i < len(steps) - 1:
            x_t = cold_update(sched, x_t, x0, s, steps[i + 1], noise)
    return x0


# ----------------------------------------------------------------------------------------------- train step
def train_step_loss(sd, sched, x0, t, eps, cond=None, heads=1, image_recon=False, target="eps"):
    """train_diffusion.py:336-350 (target eps) / train_noise_cold_diffusion.py:340 (target x0)."""
    x_t = q_sample(sched, x0, t, eps)
    pred = unet_forward(sd, x_t, t, cond, heads=heads, image_recon=image_recon)
    return F.mse_loss(pred, eps if target == "eps" else x0)


def adam_update(p, g, m, v, step, lr, b1=0.5, b2=0.999, eps=1e-8):
    """torch.optim.Adam as configured at train_diffusion.py:214-218 (betas (0.5, 0.999), no weight decay)."""
    m = b1 * m + (1 - b1) * g
    v = b2 * v + (1 - b2) * g * g
    mhat = m / (1 - b1 ** step)
    vhat = v / (1 - b2 ** step)
    return p - lr * mhat / (vhat.sqrt() + eps), m, v
