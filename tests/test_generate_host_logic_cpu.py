"""Host logic of the generation entry points (b200/generator.py: argument handling, seeding and x_T draw, the per-model
ensemble loop, checkpoint loading, hand-offs between models, SR up-sampling + delta) against tensors RETURNED BY THE
UNMODIFIED REFERENCE entry points on the same exported model folders (tests/golden/generate_runs.pt, written by
tests/golden/make_golden_generate.py).

This runs on CPU: the device-bound pieces underneath the host logic -- the kernels behind the samplers, the degraders and
the U_Net forward -- are swapped for the CPU oracle, so what is compared is exactly the code path of b200/generator.py.
The U_Net objects are still CONSTRUCTED by this repo's models.U_Net (their random initialisation consumes the global RNG
between the x_T draw and the DDPM noise draws, as the reference's constructor does), so equal outputs also prove that a
seed reproduces the reference's images."""
import json
import os

import pytest
import torch
import torch.nn.functional as F

from conftest import load_golden, rel_l2
from oracle import diffusion_oracle as orc
from oracle.weights import synth_state_dict


class _OracleDegrader:
    def __init__(self, sched):
        self.sched = sched

    def __call__(self, img, steps, eps=None):
        return orc.q_sample(self.sched, img, steps, eps)


class _OracleNet:
    def __init__(self, net, heads, recon):
        self.sd = {k: v.detach().clone() for k, v in net.state_dict().items()}
        self.heads, self.recon = heads, recon

    def __call__(self, x, t, labels=None):
        with torch.no_grad():
            return orc.unet_forward(self.sd, x, t, labels, heads=self.heads, image_recon=self.recon)


@pytest.fixture
def cpu_host_logic(monkeypatch):
    import diffusion_sampling_algorithms as S
    from b200 import generator as G
    real_load = G._load_net

    def degrader(model_dict, args, device):
        if model_dict["noise_scheduler"].upper() == "LINEAR":
            return _OracleDegrader(("linear", model_dict["beta_1"], model_dict["beta_T"], args["max_T"]))
        return _OracleDegrader(("cosine", args["max_T"]))

    def load_net(model_dict, folder, device):
        net = real_load(model_dict, folder, device)          # this repo's U_Net: constructed (RNG!), checkpoint loaded
        return _OracleNet(net, model_dict["attn_heads"], model_dict["image_recon"])

    monkeypatch.setattr(G, "_device", lambda: torch.device("cpu"))
    monkeypatch.setattr(G, "_degrader", degrader)
    monkeypatch.setattr(G, "_load_net", load_net)
    monkeypatch.setattr(G, "area_resize", lambda x, size: F.interpolate(x, size=size, mode="area"))
    monkeypatch.setattr(S, "ddpm_sampling", lambda diffusion_net, noise_degradation, x_t, min_noise, max_noise, cond_img,
                        labels_tensor, device, log: orc.ddpm_sample(diffusion_net, noise_degradation.sched, x_t, min_noise,
                                                                    max_noise, cond_img, labels_tensor))
    monkeypatch.setattr(S, "ddim_sampling", lambda diffusion_net, noise_degradation, x_t, min_noise, max_noise, cond_img,
                        labels_tensor, ddim_step_size, device, log: orc.ddim_sample(
                            diffusion_net, noise_degradation.sched, x_t, min_noise, max_noise, ddim_step_size, cond_img,
                            labels_tensor))
    monkeypatch.setattr(S, "cold_diffusion_sampling", lambda diffusion_net, noise_degradation, x_t, noise, min_noise,
                        max_noise, cond_img, labels_tensor, skip_step_size, device, log: orc.cold_sample(
                            diffusion_net, noise_degradation.sched, x_t, noise, min_noise, max_noise, skip_step_size, cond_img,
                            labels_tensor))
    return G


@pytest.mark.parametrize("name", ["ddim_ensemble", "ddpm_labels_cond_img", "ddpm_ensemble", "cold_ensemble", "sr_ensemble"])
def test_generate_entry_points_match_reference_runs(name, cpu_host_logic, tmp_path):
    fx = load_golden("generate_runs.pt")[name]
    for m, shapes in zip(fx["models"], fx["shapes"]):
        torch.save({"model": synth_state_dict(shapes, m["seed"])}, tmp_path / m["model_name"])
    (tmp_path / "config.json").write_text(json.dumps({"models": fx["models"]}))
    import generate_images_cold_diffusion, generate_images_diffusion, generate_sr_images_diffusion
    entry = {"generate_images_diffusion": generate_images_diffusion.generate_images_diffusion,
             "generate_images_cold_diffusion": generate_images_cold_diffusion.generate_images_cold_diffusion,
             "generate_sr_images_diffusion": generate_sr_images_diffusion.generate_sr_images_diffusion}[fx["entry"]]
    args = ["-c", str(tmp_path / "config.json")] + list(fx["argv"])
    got = entry(args, log=lambda *a, **k: None, save_locally=False, **fx["inputs"])
    assert got.shape == fx["result"].shape
    err = rel_l2(got, fx["result"])
    print(f"{name}: rel_l2 vs the reference entry point = {err:.2e}")
    assert err < 1e-6
