"""Runs the dominant tcgen05 kernel configurations a few times (for `ncu --set full`): the 3x3 convs that carry most
of the U-Net's FLOPs at the bench batch size."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "simple-diffusion-model_b200"))
import torch  # noqa: E402
from b200 import ops  # noqa: E402

torch.manual_seed(0)
dev = "cuda:0"
for (n, hw, c) in ((256, 16, 1024), (256, 32, 512), (256, 64, 128)):
    x = (torch.randn((n, hw, hw, c), device=dev) * 0.5).bfloat16()
    w = (torch.randn((c, 9 * c), device=dev) * 0.02).bfloat16()
    bias = torch.zeros(c, device=dev)
    stats = torch.zeros((n, 32, 2), device=dev)
    for _ in range(3):
        y = ops.conv2d(0, x, w, bias, c, act=1, gn_stats=stats, groups=32)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(5):
        y = ops.conv2d(0, x, w, bias, c, act=1, gn_stats=stats, groups=32)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 5
    print(f"conv3x3 N{n} {hw}x{hw} C{c}: {ms:.3f} ms {2.0 * n * hw * hw * c * c * 9 / ms * 1e-9:.1f} TFLOP/s", flush=True)
