"""Sweeps (N-tile width, split-K factor) of the tcgen05 implicit-GEMM kernel over the small / mid-size 3x3 conv shapes
(the deep U-Net levels) to calibrate the tiling heuristic in csrc/conv_api.cu.  python tools/sweep_tiling.py --batch 32"""
import argparse
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "simple-diffusion-model_b200"))
import torch  # noqa: E402
from b200 import ops  # noqa: E402
from bench_gemm_layers import timeit  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--batch", type=int, default=32)
ap.add_argument("--shapes", default="512x16,512x8,512x4,512x2,1024x2,1024x4,1024x8,256x32")
args = ap.parse_args()
n, dev = args.batch, "cuda"
for spec in args.shapes.split(","):
    c, hw = (int(v) for v in spec.split("x"))
    sets = []
    for _ in range(4):
        x = (torch.randn((n, hw, hw, c), device=dev) * 0.5).bfloat16()
        w = (torch.randn((c, 9 * c), device=dev) * 0.02).bfloat16()
        sets.append((x, w, torch.empty_like(x), torch.zeros(c, device=dev), torch.zeros((n, 32, 2), device=dev)))
    m_tiles = (n * hw * hw + 127) // 128
    line = [f"C{c} {hw}x{hw} N{n} (m_tiles {m_tiles}, k_iters {9 * c // 64}):"]
    best = None
    for bn in (256, 128, 64):
        for sp in (1, 2, 3, 4, 6, 8, 12, 16):
            tiles = m_tiles * (c // bn)
            if sp > 1 and (tiles * (sp - 1) >= 148 or (9 * c // 64) // sp < 2):
                continue
            os.environ["SDM_B200_FORCE_TILING"] = f"{bn},{sp}"
            t = timeit(lambda x, w, y, b, st: ops.conv2d(0, x, w, b, c, act=1, out=y, gn_stats=st, groups=32), sets) * 1e3
            line.append(f"{bn}/{sp}={t:.1f}")
            if best is None or t < best[0]:
                best = (t, bn, sp)
    os.environ.pop("SDM_B200_FORCE_TILING", None)
    t = timeit(lambda x, w, y, b, st: ops.conv2d(0, x, w, b, c, act=1, out=y, gn_stats=st, groups=32), sets) * 1e3
    print(" ".join(line), f"| best {best[1]}/{best[2]}={best[0]:.1f} us | heuristic={t:.1f} us", flush=True)
