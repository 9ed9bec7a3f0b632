"""End-to-end runs of the reference-named entry points on a tiny network with the synthetic dataset: a few training
steps through each train_* script (checkpoints written in the reference's format), then generation from the written
checkpoint through each generate_* function, including the base -> super-resolution cascade hand-off."""
import json
import os

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

NET = dict(in_channel=3, out_channel=3, num_layers=2, num_resnet_block=1, attn_layers=[1], attn_heads=1,
           attn_dim_per_head=None, time_dim=64, cond_dim=None, min_channel=64, max_channel=128, img_recon=False)


def _train_cfg(tmp, dataset, **over):
    cfg = dict(dataset_path=dataset, use_conditional=False, out_dir=str(tmp), checkpoint_steps=2, lr_steps=2, max_epoch=1,
               plot_img_count=2, flip_imgs=True, model_checkpoint=None, load_diffusion_optim=False, config_checkpoint=None,
               diffusion_lr=2e-4, batch_size=4, noise_scheduler="LINEAR", beta1=5e-3, betaT=9e-3, diffusion_alg="DDIM",
               skip_step=250, min_noise_step=1, max_noise_step=1000, max_actual_noise_step=1000, num_workers=0, **NET)
    cfg.update(over)
    path = os.path.join(tmp, "train.json")
    with open(path, "w") as f:
        json.dump(cfg, f)
    return path


def _export(tmp, ckpt_name, cfg_over=None, **model_over):
    model = dict(model_name=ckpt_name, img_C=3, img_H=32, img_W=32, in_channel=3, out_channel=3, num_layers=2,
                 num_resnet_block=1, attn_layers=[1], attn_heads=1, attn_dim_per_head=None, time_dim=64, cond_dim=None,
                 min_channel=64, max_channel=128, image_recon=False, max_noise=1000, min_noise=1, noise_scheduler="LINEAR",
                 beta_1=5e-3, beta_T=9e-3)
    model.update(model_over)
    path = os.path.join(tmp, "checkpoint", "config.json")
    with open(path, "w") as f:
        json.dump({"models": [model]}, f)
    return path


def test_train_and_generate_base(tmp_path):
    import train_diffusion
    from generate_images_diffusion import generate_images_diffusion
    tmp = str(tmp_path)
    out = train_diffusion.main(["-c", _train_cfg(tmp, "synthetic:12x3x32x32"), "--max-steps", "3"])
    assert out["global_steps"] == 3 and np.isfinite(out["loss"])
    assert abs(out["optimizer"].param_groups[0]["lr"] - 1e-4) < 1e-12          # halved once at global step 2
    for name in ("diffusion_0.pt", "config_0.pt", "diffusion_2.pt", "diffusion_3.pt"):
        assert os.path.isfile(os.path.join(tmp, "checkpoint", name)), name
    assert os.path.isfile(os.path.join(tmp, "plots", "diffusion_plot_0.jpg"))
    ck = torch.load(os.path.join(tmp, "checkpoint", "diffusion_3.pt"), map_location="cpu", weights_only=False)
    assert set(ck.keys()) == {"model", "optimizer"}
    assert ck["model"]["down_layers.0.res_layers.0.conv_block_1.conv_layer.0.weight"].shape == (64, 64, 3, 3)
    cfg = _export(tmp, "diffusion_3.pt")
    imgs = generate_images_diffusion(["-c", cfg, "-n", "3", "-s", "7", "--diff_alg", "ddim", "--ddim_step_size", "250"],
                                     log=lambda *a, **k: None, save_locally=False)
    assert imgs.shape == (3, 3, 32, 32) and torch.isfinite(imgs).all()
    again = generate_images_diffusion(["-c", cfg, "-n", "3", "-s", "7", "--diff_alg", "ddim", "--ddim_step_size", "250"],
                                      log=lambda *a, **k: None, save_locally=False)
    assert float((imgs - again).abs().max()) < 1e-2 * float(imgs.abs().max())   # same seed -> same images (bf16 atomics aside)
    with pytest.raises(ValueError):
        generate_images_diffusion(["-c", cfg, "-n", "0"], save_locally=False)


def test_train_cold_and_generate(tmp_path):
    import train_noise_cold_diffusion
    from generate_images_cold_diffusion import generate_images_cold_diffusion
    tmp = str(tmp_path)
    out = train_noise_cold_diffusion.main(["-c", _train_cfg(tmp, "synthetic:8x3x32x32:4", use_conditional=True, cond_dim=4,
                                                             noise_scheduler="COSINE", img_recon=True), "--max-steps", "2"])
    assert out["global_steps"] == 2 and np.isfinite(out["loss"])
    assert os.path.isfile(os.path.join(tmp, "labels.txt"))
    cfg = _export(tmp, "diffusion_2.pt", cond_dim=4, image_recon=True, noise_scheduler="COSINE")
    imgs = generate_images_cold_diffusion(["-c", cfg, "-n", "2", "-s", "1", "--cold_step_size", "250", "-l", "1", "0", "0", "1"],
                                          log=lambda *a, **k: None, save_locally=False)
    assert imgs.shape == (2, 3, 32, 32) and torch.isfinite(imgs).all() and float(imgs.abs().max()) <= 1.0
    with pytest.raises(ValueError):
        generate_images_cold_diffusion(["-c", cfg, "-n", "2"], save_locally=False)     # labels missing


def test_train_sr_and_cascade(tmp_path):
    import train_SR_diffusion
    from generate_sr_images_diffusion import generate_sr_images_diffusion
    tmp = str(tmp_path)
    out = train_SR_diffusion.main(["-c", _train_cfg(tmp, "synthetic:8x3x64x64", in_channel=6, img_recon=True,
                                                     noise_scheduler="COSINE", lr_dim=16, sr_dim=64, cond_t=250, skip_step=500),
                                   "--max-steps", "2"])
    assert out["global_steps"] == 2 and np.isfinite(out["loss"])
    cfg = _export(tmp, "diffusion_2.pt", img_H=64, img_W=64, in_channel=6, image_recon=True, noise_scheduler="COSINE", cond_t=250)
    lr_img = (np.random.RandomState(0).rand(16, 16, 3) * 255).astype(np.uint8)        # the cascade hand-off format (HWC BGR)
    sr = generate_sr_images_diffusion(["-c", cfg, "-s", "3", "--cold_step_size", "500"], lr_img=lr_img,
                                      log=lambda *a, **k: None, save_locally=False)
    assert sr.shape == (1, 3, 64, 64) and torch.isfinite(sr).all()
    with pytest.raises(ValueError):
        generate_sr_images_diffusion(["-c", cfg], lr_img="not an array", save_locally=False)


def test_train_doodle(tmp_path):
    import train_doodle_diffusion
    tmp = str(tmp_path)
    out = train_doodle_diffusion.main(["-c", _train_cfg(tmp, "synthetic:8x3x32x32:img", in_channel=6, diffusion_alg="DDPM",
                                                         max_noise_step=20, max_actual_noise_step=20, skip_step=5,
                                                         checkpoint_steps=100), "--max-steps", "2"])
    assert out["global_steps"] == 2 and np.isfinite(out["loss"])


def test_area_resize_matches_torch():
    import torch.nn.functional as F
    from b200.functional import area_resize
    x = torch.randn((2, 3, 64, 64), device="cuda")
    down = area_resize(x, (16, 16))
    assert torch.allclose(down, F.interpolate(x, size=(16, 16), mode="area"), atol=1e-6)
    assert torch.equal(area_resize(down, (64, 64)), F.interpolate(down, size=(64, 64), mode="area"))
    assert torch.allclose(area_resize(x, (24, 40)), F.interpolate(x, size=(24, 40), mode="area"), atol=1e-6)
