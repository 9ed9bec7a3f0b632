import os, sys, cProfile, pstats, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "simple-diffusion-model_b200")); sys.path.insert(1, ROOT)
import torch
from b200.optim import FusedAdam
from b200.parallel import DataParallel
from b200.steps import eps_prediction_step
from degraders import NoiseDegradation
from models.U_Net import U_Net
dev = torch.device("cuda:0")
torch.manual_seed(0)
net = U_Net().to(dev).train().set_precision("bf16")
dp = DataParallel(net, device=dev)
opt = FusedAdam(net.parameters(), lr=2e-5, betas=(0.5, 0.999))
deg = NoiseDegradation(5e-3, 9e-3, 1000, device=dev)
n = int(sys.argv[1]) if len(sys.argv) > 1 else 8
x0 = torch.rand((n, 3, 64, 64), device=dev) * 2 - 1
def step():
    eps = torch.randn_like(x0); t = torch.randint(1, 1000, (n,), device=dev)
    return eps_prediction_step(net, deg, opt, x0, t, eps)
for _ in range(3): step()
torch.cuda.synchronize()
t0 = time.perf_counter(); step(); t1 = time.perf_counter(); torch.cuda.synchronize(); t2 = time.perf_counter()
print(f"host enqueue {1e3*(t1-t0):.1f} ms, until GPU idle {1e3*(t2-t0):.1f} ms")
pr = cProfile.Profile(); pr.enable(); step(); pr.disable(); torch.cuda.synchronize()
pstats.Stats(pr).sort_stats("cumulative").print_stats(28)
