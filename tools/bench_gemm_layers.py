"""Per-layer throughput of the tcgen05 kernels at the 3x3/s1 conv shapes of the class-default U-Net (which carry 92 %
of its FLOPs): forward (bias + Swish + GroupNorm sums epilogue), data gradient, weight gradient.  CUDA-graph timing,
rotating buffers.   python tools/bench_gemm_layers.py --batch 256 --fwd-only | --batch 32"""
import argparse
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "simple-diffusion-model_b200"))
import torch  # noqa: E402
from b200 import ops  # noqa: E402

# (channels, spatial size at a 64x64 input, number of such convs in the net)
SHAPES = [(128, 64, 14), (256, 32, 10), (512, 16, 10), (512, 8, 10), (512, 4, 10), (512, 2, 2), (1024, 2, 10), (1024, 4, 10),
          (1024, 8, 10), (1024, 16, 10), (512, 32, 10)]


def timeit(launch, sets, iters=10):
    for a in sets:
        launch(*a)
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        for i in range(iters):
            launch(*sets[i % len(sets)])
    g.replay()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    g.replay()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--batch", type=int, default=32)
    ap.add_argument("--fwd-only", action="store_true")
    ap.add_argument("--wgrad-only", action="store_true")
    ap.add_argument("--scale", type=int, default=1, help="spatial scale (2 = 128x128 input)")
    ap.add_argument("--dgrad-only", action="store_true")
    ap.add_argument("--dgrad-mode", type=int, default=0, help="0: transposed weight copy (K-major B); 5: forward weights consumed MN-major")
    ap.add_argument("--json", default=None)
    args = ap.parse_args()
    n, dev = args.batch, "cuda"
    rows, tot = [], {"fwd": 0.0, "dgrad": 0.0, "wgrad": 0.0}
    flops_tot = 0.0
    for c, hw, count in SHAPES:
        hw *= args.scale
        flops = 2.0 * n * hw * hw * c * c * 9
        act_bytes = n * hw * hw * c * 2
        reps = max(2, min(6, (200 << 20) // max(act_bytes + c * c * 18, 1) + 1))

        def mk():
            x = (torch.randn((n, hw, hw, c), device=dev) * 0.5).bfloat16()
            dz = (torch.randn((n, hw, hw, c), device=dev) * 0.5).bfloat16()
            w = (torch.randn((c, 9 * c), device=dev) * 0.02).bfloat16()
            y = torch.empty_like(x)
            gw = torch.zeros((c, 9 * c), device=dev)
            return x, dz, w, y, gw, torch.zeros(c, device=dev), torch.zeros((n, 32, 2), device=dev)

        sets = [mk() for _ in range(reps)]
        res = {"C": c, "hw": hw, "count": count, "gflop": flops / 1e9}
        if not args.wgrad_only and not args.dgrad_only:
            t = timeit(lambda x, dz, w, y, gw, b, st: ops.conv2d(0, x, w, b, c, act=1, out=y, gn_stats=st, groups=32), sets)
            res["fwd_us"], res["fwd_tflops"] = t * 1e3, flops / t / 1e9
            tot["fwd"] += t * count
        if not args.fwd_only:
            if not args.wgrad_only:
                t = timeit(lambda x, dz, w, y, gw, b, st: ops.conv2d(args.dgrad_mode, dz, w, None, c, act=0, out=y), sets)
                res["dgrad_us"], res["dgrad_tflops"] = t * 1e3, flops / t / 1e9
                tot["dgrad"] += t * count
            if not args.dgrad_only:
                t = timeit(lambda x, dz, w, y, gw, b, st: ops.conv2d_wgrad(0, x, dz, c, gw), sets)
                res["wgrad_us"], res["wgrad_tflops"] = t * 1e3, flops / t / 1e9
                tot["wgrad"] += t * count
        flops_tot += flops * count
        rows.append(res)
        print("  ".join(f"{k}={v:.1f}" if isinstance(v, float) else f"{k}={v}" for k, v in res.items()), flush=True)
    for k, v in tot.items():
        if v > 0:
            print(f"{k}: {v:.3f} ms for all 3x3/s1 convs of one pass at batch {n} -> {flops_tot / v / 1e9:.1f} TFLOP/s")
    if args.json:
        json.dump({"batch": n, "rows": rows, "total_ms": tot, "flops": flops_tot}, open(args.json, "w"), indent=1)


if __name__ == "__main__":
    main()
