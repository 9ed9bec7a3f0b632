"""Launches the memory-bound kernels once each at bench-sized shapes, for an `ncu --set full` capture
(dram__bytes_read/write + gpu__time_duration -> achieved HBM GB/s per kernel)."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "simple-diffusion-model_b200"))
import torch  # noqa: E402
from b200 import ops  # noqa: E402
from b200._lib import call, ptr, stream  # noqa: E402

dev, bf = "cuda", torch.bfloat16
flush = torch.empty(512 << 20, dtype=torch.uint8, device=dev)


def cold():
    flush.zero_()          # evict L2 so the next launch streams from HBM
    torch.cuda.synchronize()


for n, c, hw in ((256, 1024, 16), (256, 128, 64), (32, 512, 32), (32, 128, 128)):
    p = hw * hw
    y = torch.randn((n, hw, hw, c), device=dev, dtype=bf)
    res = torch.randn((n, hw, hw, c), device=dev, dtype=bf)
    out = torch.empty_like(y)
    stats = torch.stack([torch.zeros((n, 32), device=dev), torch.full((n, 32), float(p * c // 32), device=dev)], dim=-1).contiguous()
    gamma, beta, s = torch.ones(c, device=dev), torch.zeros(c, device=dev), torch.randn((n, c), device=dev)
    cold(); ops.adagn_apply(y, stats, gamma, beta, s, c, out=out)
    cold(); ops.adagn_apply(y, stats, gamma, beta, s, c, out=out, residual=res)
    cold(); ops.adagn_apply(y, stats, gamma, beta, s, c, out=out, pre_swish=True)
    work = torch.zeros((2 * n * c,), device=dev)
    ds, dg, db, dbias = torch.zeros((n, c), device=dev), torch.zeros(c, device=dev), torch.zeros(c, device=dev), torch.zeros(c, device=dev)
    cold()
    call("b2_adagn_bwd", ptr(res), c, ptr(y), c, ptr(stats), ptr(gamma), ptr(beta), ptr(s), c, ptr(work), ptr(ds), c, ptr(dg), ptr(db),
         ptr(out), c, ptr(dbias), n, p, c, 32, 1e-5, 0, stream())
    cold(); ops.act(0, None, y, out, None, n * p, c, 0, c, c, 0)
    cold(); ops.act(1, res, y, out, dbias, n * p, c, c, c, c, 0)
    del y, res, out
npar = 128 << 20
bufs = [torch.randn(npar, device=dev).abs_() for _ in range(4)]
sh = torch.empty(npar, dtype=bf, device=dev)
cold()
call("b2_adam_flat", ptr(bufs[0]), ptr(bufs[1]), ptr(bufs[2]), ptr(bufs[3]), npar, 0.5, 0.999, 1e-8, 1e-4, 1.0, 1.0, ptr(sh), stream())
x = [torch.randn((256, 3, 64, 64), device=dev) for _ in range(3)]
t = torch.randint(1, 1000, (256,), device=dev)
cold(); call("b2_qsample", ptr(x[0]), ptr(x[1]), ptr(x[2]), ptr(t), 256, None, 1000, 256, 3 * 64 * 64, stream())
cold(); call("b2_ddim_step", ptr(x[0]), ptr(x[1]), None, ptr(x[2]), None, x[0].numel(), 1.1, 0.3, 0.9, 0.2, 0.0, 0, stream())
torch.cuda.synchronize()
print("done")
