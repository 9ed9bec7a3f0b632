"""Generates tests/golden/*.pt by running the UNMODIFIED reference from /root/reference (read-only).

Run in the build container only:  PYTHONDONTWRITEBYTECODE=1 python tests/golden/make_golden.py
The GPU box has no /root/reference; tests read the committed fixtures.
"""
import os
import sys

os.environ.setdefault("PYTHONDONTWRITEBYTECODE", "1")
sys.dont_write_bytecode = True
HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, "/root/reference")
sys.path.insert(1, ROOT)

import torch  # noqa: E402
import torch.nn.functional as F  # noqa: E402

from models.U_Net import U_Net  # noqa: E402  (reference)
from degraders import NoiseDegradation, CosineNoiseDegradation  # noqa: E402  (reference)
import diffusion_sampling_algorithms as ref_samplers  # noqa: E402  (reference)
from oracle.weights import synth_state_dict  # noqa: E402

torch.set_num_threads(8)

UNET_CASES = {
    # name: (U_Net kwargs, N, H, W, schedule, target)
    "tiny": (dict(num_resnet_blocks=1, time_dim=32, num_layers=2, attn_layers=[1], min_channel=32, max_channel=64), 2, 16, 16),
    "tiny_cond": (dict(num_resnet_blocks=2, in_channel=6, time_dim=32, cond_dim=5, num_layers=2, attn_layers=[0, 1], num_heads=2,
                       dim_per_head=16, min_channel=32, max_channel=64, image_recon=True), 2, 16, 16),
    "gpu_small": (dict(num_resnet_blocks=1, time_dim=64, num_layers=2, attn_layers=[1], min_channel=128, max_channel=256), 2, 32, 32),
    "gpu_cond": (dict(num_resnet_blocks=2, in_channel=6, time_dim=64, cond_dim=10, num_layers=3, attn_layers=[1, 2], num_heads=2,
                      dim_per_head=64, min_channel=128, max_channel=256, image_recon=True), 3, 32, 32),
    "default64": (dict(), 2, 64, 64),
    # BASELINE configs[2] / configs[3] shapes on the class-default net: S = 1024 attention (multi-tile softmax fix-up inside the
    # full net) with labels at 128x128; 6-channel image_recon (SR) net at 256x256 (S = 4096 attention).  Gradients are kept as
    # 1024 strided samples per tensor to bound the fixture size.
    "default128_cond": (dict(cond_dim=10), 2, 128, 128),
    "default256_sr": (dict(in_channel=6, image_recon=True), 1, 256, 256),
}
SAMPLES = {"default128_cond": 1024, "default256_sr": 1024}


def unet_case(name, kwargs, n, h, w):
    torch.manual_seed(0)
    net = U_Net(**kwargs)
    shapes = {k: tuple(v.shape) for k, v in net.state_dict().items()}
    seed = 1234
    net.load_state_dict(synth_state_dict(shapes, seed))
    g = torch.Generator().manual_seed(99)
    in_ch = kwargs.get("in_channel", 3)
    x = torch.rand((n, in_ch, h, w), generator=g) * 2 - 1
    t = torch.randint(1, 1000, (n,), generator=g)
    cond = None
    if kwargs.get("cond_dim"):
        cond = (torch.rand((n, kwargs["cond_dim"]), generator=g) > 0.7).float()
    target = torch.randn((n, kwargs.get("out_channel", 3), h, w), generator=g)
    net.train()
    out = net(x, t, cond)
    loss = F.mse_loss(out, target)
    loss.backward()
    grads, no_grad = {}, []
    for pname, p in net.named_parameters():
        if p.grad is None:
            no_grad.append(pname)
            continue
        gflat = p.grad.detach().flatten()
        # large tensors: `ns` evenly strided samples (index i * numel // ns) instead of the whole gradient
        ns = SAMPLES.get(name, 4096)
        idx = (torch.arange(ns, dtype=torch.int64) * gflat.numel()) // ns if gflat.numel() > ns else None
        grads[pname] = {"norm": float(gflat.norm()), "sum": float(gflat.sum()),
                        "head": gflat[:64].clone(),
                        "sample": gflat[idx].clone() if idx is not None else None,
                        "full": p.grad.detach().clone() if gflat.numel() <= ns else None}
    # batch-1 timestep broadcast (samplers pass t of shape [1])
    net.eval()
    with torch.no_grad():
        out_t1 = net(x, t[:1], cond[0] if cond is not None else None)
    fx = dict(kwargs=kwargs, shapes=shapes, seed=seed, x=x, t=t, cond=cond, target=target, out=out.detach(),
              loss=float(loss), grads=grads, no_grad=sorted(no_grad), out_t1=out_t1, n_samples=SAMPLES.get(name, 4096))
    torch.save(fx, os.path.join(HERE, f"unet_{name}.pt"))
    print(name, "params", sum(int(torch.tensor(s).prod()) for s in shapes.values()), "loss", float(loss), "no-grad params", len(no_grad))


def schedule_case():
    lin = NoiseDegradation(5e-3, 9e-3, 1000)
    steps = torch.arange(0, 1001)
    b, a, ab = lin.get_timestep_params(steps)
    cos = CosineNoiseDegradation(1000)
    steps_c = torch.arange(1, 1001)
    cb, ca, cab = cos.get_timestep_params(steps_c)
    g = torch.Generator().manual_seed(5)
    img = torch.rand((4, 3, 8, 8), generator=g) * 2 - 1
    eps = torch.randn((4, 3, 8, 8), generator=g)
    t = torch.tensor([1, 250, 999, 1000])
    fx = dict(lin_args=(5e-3, 9e-3, 1000), lin_beta=b, lin_alpha=a, lin_abar=ab, cos_beta=cb, cos_alpha=ca, cos_abar=cab,
              img=img, eps=eps, t=t, lin_q=lin(img, t, eps), cos_q=cos(img, t, eps), cos_q_t1=cos(img, t[1:2], eps))
    torch.save(fx, os.path.join(HERE, "schedules.pt"))
    print("schedules ok")


class StubNet(torch.nn.Module):
    """A deterministic stand-in for the U-Net so sampler arithmetic is pinned independently of the network."""

    def __init__(self):
        super().__init__()
        self.calls = []

    def forward(self, x, t, labels=None):
        self.calls.append(int(t[0]))
        base = x[:, :3]
        out = 0.3 * torch.tanh(base) + 0.001 * t.float()[:, None, None, None] / 10.0
        if x.shape[1] > 3:
            out = out + 0.05 * x[:, 3:6]
        return out


def sampler_case():
    fx = {}
    quiet = lambda *a, **k: None
    g = torch.Generator().manual_seed(11)
    x_T = torch.randn((3, 3, 8, 8), generator=g)
    cond_img = torch.rand((3, 3, 8, 8), generator=g) * 2 - 1
    lin = NoiseDegradation(5e-3, 9e-3, 1000)
    cos = CosineNoiseDegradation(1000)
    for sched_name, sched in (("linear", lin), ("cosine", cos)):
        for (mn, mx, step) in ((1, 1000, 20), (1, 1000, 100), (250, 1000, 37), (1, 60, 7)):
            net = StubNet()
            torch.manual_seed(2024)
            out = ref_samplers.ddim_sampling(net, sched, x_T.clone(), min_noise=mn, max_noise=mx, ddim_step_size=step, log=quiet)
            rng_after = torch.rand(1).item()
            fx[f"ddim_{sched_name}_{mn}_{mx}_{step}"] = dict(out=out, calls=net.calls, rng_after=rng_after)
            net = StubNet()
            out = ref_samplers.cold_diffusion_sampling(net, sched, x_T.clone(), x_T.clone(), min_noise=mn, max_noise=mx,
                                                       skip_step_size=step, log=quiet)
            fx[f"cold_{sched_name}_{mn}_{mx}_{step}"] = dict(out=out, calls=net.calls)
        net = StubNet()
        torch.manual_seed(2025)
        out = ref_samplers.ddpm_sampling(net, sched, x_T.clone(), min_noise=1, max_noise=40, log=quiet)
        fx[f"ddpm_{sched_name}_1_40"] = dict(out=out, calls=net.calls, rng_after=torch.rand(1).item())
        net = StubNet()
        torch.manual_seed(2026)
        out = ref_samplers.ddim_sampling(net, sched, x_T.clone(), min_noise=1, max_noise=1000, ddim_step_size=50,
                                         cond_img=cond_img, log=quiet)
        fx[f"ddim_cond_{sched_name}"] = dict(out=out, calls=net.calls)
    fx["x_T"] = x_T
    fx["cond_img"] = cond_img
    torch.save(fx, os.path.join(HERE, "samplers.pt"))
    print("samplers ok", len(fx))


if __name__ == "__main__":
    which = sys.argv[1:] or ["schedules", "samplers"] + list(UNET_CASES)
    if "schedules" in which:
        schedule_case()
    if "samplers" in which:
        sampler_case()
    for name, (kw, n, h, w) in UNET_CASES.items():
        if name in which:
            unet_case(name, kw, n, h, w)
