"""AdaGN backward (b2_adagn_bwd: reduce + apply passes) at the training shapes of BASELINE configs[2] (128x128, batch 32):
time per layer and achieved HBM bandwidth on the ALGORITHMIC 6 bytes / element (dout + z read once, dz written once).
Rotating buffer sets larger than L2, CUDA-graph timed.  Environment knobs are read by the library at first use, so A/B runs
are separate processes:
    SDM_B200_BWD_L2_CHUNK_MB=0|32|64|96   SDM_B200_BWD_U=2|4      python tools/bench_adagn_bwd.py [--batch 32] [--json out]
"""
import argparse
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "simple-diffusion-model_b200"))
sys.path.insert(1, ROOT)
import torch  # noqa: E402

from b200._lib import call, ptr, stream  # noqa: E402

PEAK = 6552.0
if os.path.exists(os.path.join(ROOT, "MEASURED_PEAKS.json")):
    PEAK = float(json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"])


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--batch", type=int, default=32)
    ap.add_argument("--json", default=None)
    a = ap.parse_args()
    n, dev, bf = a.batch, "cuda", torch.bfloat16
    # (C, H): the AdaGN layers of the class-default net at 128x128 and how many of the 100 have that shape
    layers = [(128, 128, 10), (256, 64, 10), (512, 32, 10), (512, 16, 10), (512, 8, 10), (1024, 4, 10), (1024, 8, 10),
              (1024, 16, 10), (1024, 32, 10), (512, 64, 10)]
    rows, total_ms, total_bytes = [], 0.0, 0
    for c, hw, count in layers:
        p = hw * hw
        elems = n * p * c
        reps = max(2, min(8, (384 << 20) // (elems * 6) + 1))
        sets = []
        for _ in range(reps):
            dout = torch.randn((n, hw, hw, c), device=dev, dtype=bf)
            z = torch.randn((n, hw, hw, c), device=dev, dtype=bf)
            dz = torch.empty_like(z)
            y = torch.nn.functional.silu(z.float())
            g = y.reshape(n, p, 32, c // 32)
            stats = torch.stack([g.sum(dim=(1, 3)), (g * g).sum(dim=(1, 3))], dim=-1).contiguous()
            sets.append((dout, z, dz, stats, torch.ones(c, device=dev), torch.zeros(c, device=dev), torch.randn((n, c), device=dev),
                         torch.zeros((2 * n * c,), device=dev), torch.zeros((n, c), device=dev), torch.zeros(c, device=dev),
                         torch.zeros(c, device=dev), torch.zeros(c, device=dev)))
            del y, g

        def launch(dout, z, dz, st, ga, be, s, work, ds, dg, db, dbias):
            call("b2_adagn_bwd", ptr(dout), c, ptr(z), c, ptr(st), ptr(ga), ptr(be), ptr(s), c, ptr(work), ptr(ds), c, ptr(dg), ptr(db),
                 ptr(dz), c, ptr(dbias), n, p, c, 32, 1e-5, 0, stream())

        for s_ in sets:
            launch(*s_)
        torch.cuda.synchronize()
        iters = 4 * reps
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g):
            for i in range(iters):
                launch(*sets[i % reps])
        g.replay()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        g.replay()
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / iters
        gbs = elems * 6 / ms / 1e6
        rows.append({"C": c, "HW": hw, "tensor_MB": elems * 2 / 2 ** 20, "us": ms * 1e3, "GBps_on_6B": gbs, "frac": gbs / PEAK, "layers": count})
        total_ms += ms * count
        total_bytes += elems * 6 * count
        print(f"C{c:5d} {hw:3d}x{hw:<3d} {elems * 2 / 2 ** 20:7.1f} MB/tensor {ms * 1e3:8.1f} us  {gbs:7.0f} GB/s on 6 B/elem  {gbs / PEAK * 100:5.1f} %", flush=True)
        del sets
        torch.cuda.empty_cache()
    agg = total_bytes / total_ms / 1e6
    print(f"all 100 layers: {total_ms:.3f} ms per backward pass, {agg:.0f} GB/s on {total_bytes / 1e6:.0f} MB algorithmic = {agg / PEAK * 100:.1f} % of {PEAK:.0f}"
          f"   [chunk={os.environ.get('SDM_B200_BWD_L2_CHUNK_MB', 'default')} U={os.environ.get('SDM_B200_BWD_U', 'auto')}]", flush=True)
    if a.json:
        json.dump({"batch": n, "hbm_peak_gbs": PEAK, "rows": rows, "total_ms": total_ms, "aggregate_GBps_on_6B": agg,
                   "env": {k: v for k, v in os.environ.items() if k.startswith("SDM_B200_")}}, open(a.json, "w"), indent=1)


if __name__ == "__main__":
    main()
