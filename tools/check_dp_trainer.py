"""Multi-GPU check of the training entry point (run under torchrun on N GPUs): `train_diffusion.main` on the synthetic dataset
for a few steps with batch-sharded data parallelism, uint8 input pipeline, in-kernel Philox noise and the CUDA-graph step.
Asserts: finite loss; every rank ends with bit-identical weights (the all-reduced gradients and the fused Adam agree); the
ranks drew DIFFERENT timesteps (rank-dependent seeds) -- checked through the per-rank torch generator state."""
import json
import os
import sys
import tempfile

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "simple-diffusion-model_b200")); sys.path.insert(1, ROOT)
import torch
import torch.distributed as dist

world, rank, local = int(os.environ["WORLD_SIZE"]), int(os.environ["RANK"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)
import train_diffusion

tmp = tempfile.mkdtemp(prefix=f"dp_trainer_r{rank}_")
cfg = dict(dataset_path="synthetic:64x3x32x32", use_conditional=False, out_dir=tmp, checkpoint_steps=100, lr_steps=100, max_epoch=1,
           plot_img_count=0, flip_imgs=True, model_checkpoint=None, load_diffusion_optim=False, config_checkpoint=None,
           diffusion_lr=2e-4, batch_size=4, noise_scheduler="LINEAR", beta1=5e-3, betaT=9e-3, diffusion_alg="DDIM", skip_step=250,
           min_noise_step=1, max_noise_step=1000, max_actual_noise_step=1000, num_workers=0, in_channel=3, out_channel=3,
           num_layers=2, num_resnet_block=1, attn_layers=[1], attn_heads=1, attn_dim_per_head=None, time_dim=64, cond_dim=None,
           min_channel=128, max_channel=256, img_recon=False, seed=11, philox_noise=True)
path = os.path.join(tmp, "train.json")
json.dump(cfg, open(path, "w"))
out = train_diffusion.main(["-c", path, "--max-steps", "4"])
assert out["global_steps"] == 4 and out["loss"] == out["loss"], out["loss"]
net = out["net"]
flat = net.engine().layout.params_flat
digest = torch.stack([flat.double().sum(), flat.double().abs().sum(), flat[::997].double().pow(2).sum()])
all_d = [torch.empty_like(digest) for _ in range(world)]
dist.all_gather(all_d, digest)
same = all(torch.equal(all_d[0], d) for d in all_d)
# the next draw of each rank's CUDA generator differs when the ranks were seeded base + rank
probe = torch.randint(0, 1 << 30, (4,), device=dev).double()
all_p = [torch.empty_like(probe) for _ in range(world)]
dist.all_gather(all_p, probe)
distinct = len({tuple(p.tolist()) for p in all_p}) == world
print(f"rank {rank}: loss {out['loss']:.5f}; weights identical across ranks: {same}; per-rank random streams distinct: {distinct}", flush=True)
assert same and distinct
dist.barrier()
dist.destroy_process_group()
