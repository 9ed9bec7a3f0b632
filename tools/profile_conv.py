"""Launches the dominant tcgen05 kernel configurations a few times each, for `ncu --set full` captures:
    ncu --set full --clock-control none --import-source on -k regex:gemm -o out python tools/profile_conv.py [--batch 32]
Kernel launch order per shape: forward conv (bias + Swish + GroupNorm sums), then the weight gradient."""
import argparse
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "simple-diffusion-model_b200"))
import torch  # noqa: E402
from b200 import ops  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--batch", type=int, default=256)
ap.add_argument("--shapes", default="1024x16,512x32,128x64")
ap.add_argument("--wgrad", action="store_true")
ap.add_argument("--iters", type=int, default=2)
args = ap.parse_args()
torch.manual_seed(0)
dev = "cuda:0"
n = args.batch
for spec in args.shapes.split(","):
    c, hw = (int(v) for v in spec.split("x"))
    x = (torch.randn((n, hw, hw, c), device=dev) * 0.5).bfloat16()
    dz = (torch.randn((n, hw, hw, c), device=dev) * 0.5).bfloat16()
    w = (torch.randn((c, 9 * c), device=dev) * 0.02).bfloat16()
    bias = torch.zeros(c, device=dev)
    stats = torch.zeros((n, 32, 2), device=dev)
    gw = torch.zeros((c, 9 * c), device=dev)
    for _ in range(args.iters):
        y = ops.conv2d(0, x, w, bias, c, act=1, gn_stats=stats, groups=32)
        if args.wgrad:
            ops.conv2d_wgrad(0, x, dz, c, gw)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(args.iters):
        y = ops.conv2d(0, x, w, bias, c, act=1, gn_stats=stats, groups=32)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / args.iters
    print(f"conv3x3 N{n} {hw}x{hw} C{c}: {ms:.3f} ms {2.0 * n * hw * hw * c * c * 9 / ms * 1e-9:.1f} TFLOP/s", flush=True)
