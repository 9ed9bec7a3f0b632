"""Multi-GPU check (run under torchrun on N GPUs): data-parallel gradients (per-rank shard, bucketed NCCL all-reduce
overlapped with backward, 1/world scaling) equal the single-process gradient of the whole batch; sharded DDIM sampling
equals unsharded sampling bit for bit."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "simple-diffusion-model_b200")); sys.path.insert(1, ROOT)
import torch
import torch.distributed as dist

world, rank, local = int(os.environ["WORLD_SIZE"]), int(os.environ["RANK"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)
from b200.functional import mse_loss
from b200.parallel import DataParallel, shard_range
from degraders import NoiseDegradation
import diffusion_sampling_algorithms as S
from models.U_Net import U_Net

kw = dict(num_resnet_blocks=1, num_layers=2, attn_layers=[1], min_channel=128, max_channel=256)
torch.manual_seed(0)
ref = U_Net(**kw).to(dev).train().set_precision("tf32")
net = U_Net(**kw).to(dev).train().set_precision("tf32")
net.load_state_dict(ref.state_dict())
dp = DataParallel(net, bucket_bytes=8 << 20, device=dev)
g = torch.Generator(device="cpu").manual_seed(7)
n_total = 4 * world
x = (torch.rand((n_total, 3, 32, 32), generator=g) * 2 - 1).to(dev)
t = torch.randint(1, 1000, (n_total,), generator=g).to(dev)
tgt = torch.randn((n_total, 3, 32, 32), generator=g).to(dev)
lo, hi = shard_range(n_total, rank, world)
loss = mse_loss(net(x[lo:hi], t[lo:hi]), tgt[lo:hi])
loss.backward()
got = dp.layout.flat.clone() * dp.grad_scale
# single-process reference on the whole batch (every rank computes it; no collective involved)
loss_ref = mse_loss(ref(x, t), tgt)
loss_ref.backward()
want = ref.engine().layout.flat
err = float((got - want).norm() / want.norm())
print(f"rank {rank}: DP gradient vs full-batch gradient rel-L2 = {err:.3e}; buckets launched = {len(dp.launched)}", flush=True)
assert err < 2e-3, err
# sharded sampling == unsharded sampling (no collective in the path)
net.eval()
deg = NoiseDegradation(5e-3, 9e-3, 1000, device=dev)
x_T = torch.randn((2 * world, 3, 32, 32), generator=g).to(dev)
full = S.ddim_sampling(net, deg, x_T, ddim_step_size=250, device=dev, log=lambda *a, **k: None)
lo, hi = shard_range(2 * world, rank, world)
mine = S.ddim_sampling(net, deg, x_T[lo:hi].clone(), ddim_step_size=250, device=dev, log=lambda *a, **k: None)
same = torch.equal(mine, full[lo:hi])
rel = float((mine - full[lo:hi]).norm() / full[lo:hi].norm())
print(f"rank {rank}: sharded DDIM vs unsharded: bitwise {same}, rel-L2 {rel:.2e} (GroupNorm sums use fp32 atomics: order-dependent in the last bits)", flush=True)
assert rel < 1e-3, rel
# deterministic mode (no fp atomics, no batch-dependent split-K): the shard IS the slice of the unsharded result, bit for bit
import b200
b200.set_deterministic(True)
full_d = S.ddim_sampling(net, deg, x_T, ddim_step_size=250, device=dev, log=lambda *a, **k: None)
mine_d = S.ddim_sampling(net, deg, x_T[lo:hi].clone(), ddim_step_size=250, device=dev, log=lambda *a, **k: None)
b200.set_deterministic(False)
print(f"rank {rank}: deterministic mode: sharded DDIM vs unsharded bitwise {torch.equal(mine_d, full_d[lo:hi])}", flush=True)
assert torch.equal(mine_d, full_d[lo:hi])
dist.barrier()
dist.destroy_process_group()
