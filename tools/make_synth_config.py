"""Writes a training JSON (reference schema, create_diffusion_config.py) for the synthetic dataset: smoke runs of the train_*.py
entry points without image files.   python tools/make_synth_config.py out.json [--out-dir DIR] [--size 64] [--batch 8]"""
import argparse
import json

ap = argparse.ArgumentParser()
ap.add_argument("path")
ap.add_argument("--out-dir", default="/tmp/sdm_b200_train")
ap.add_argument("--size", type=int, default=64)
ap.add_argument("--batch", type=int, default=8)
ap.add_argument("--images", type=int, default=64)
ap.add_argument("--small", action="store_true", help="2-level 64/128-channel net instead of the creator defaults")
a = ap.parse_args()
net = dict(num_layers=2, num_resnet_block=1, attn_layers=[1], min_channel=64, max_channel=128, time_dim=64) if a.small else \
    dict(num_layers=4, num_resnet_block=1, attn_layers=[2, 3], min_channel=128, max_channel=512, time_dim=512)
cfg = dict(dataset_path=f"synthetic:{a.images}x3x{a.size}x{a.size}", use_conditional=False, cond_dim=None, out_dir=a.out_dir,
           checkpoint_steps=1000, lr_steps=100000, max_epoch=1, plot_img_count=0, flip_imgs=True, model_checkpoint=None,
           load_diffusion_optim=False, config_checkpoint=None, diffusion_lr=2e-5, batch_size=a.batch, noise_scheduler="LINEAR",
           beta1=5e-3, betaT=9e-3, diffusion_alg="DDIM", skip_step=100, min_noise_step=1, max_noise_step=1000,
           max_actual_noise_step=1000, in_channel=3, out_channel=3, attn_heads=1, attn_dim_per_head=None, img_recon=False,
           num_workers=0, **net)
json.dump(cfg, open(a.path, "w"), indent=1)
print(a.path)
