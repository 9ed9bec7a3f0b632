"""Diagnostic: b2_conv2d_nhwc mode 5 (data gradient from the forward weights, MN-major B) against mode 0 on the transposed copy."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "simple-diffusion-model_b200"))
import torch  # noqa: E402


def main():
    from b200 import ops
    cfgs = [(3, 128, 256, 8, 8), (4, 256, 256, 16, 16), (40, 256, 512, 32, 32), (40, 128, 512, 32, 32), (10, 128, 128, 64, 64),
            (40, 512, 512, 32, 32), (2, 256, 128, 64, 64)]
    for cfg in cfgs:
        n, cin, cout, h, w = cfg
        g = torch.Generator(device="cuda").manual_seed(7)
        wt = torch.randn((cout, cin, 3, 3), device="cuda", generator=g) * (1.0 / (cout * 9) ** 0.5)
        dy = torch.randn((n, cout, h, w), device="cuda", generator=g).permute(0, 2, 3, 1).contiguous().to(torch.bfloat16)
        w_fwd = ops.pack_weight(0, wt, cout, cin, cin, 0)
        w_tr = ops.pack_weight(1, wt, cout, cin, cout, 0)
        want = ops.conv2d(0, dy, w_tr, None, cin, act=0).float()
        got = ops.conv2d(5, dy, w_fwd, None, cin, act=0).float()
        torch.cuda.synchronize()
        bad = ~torch.isfinite(got)
        diff = (got - want).abs()
        diff[bad] = 0
        wrong = (diff > 1e-2) | bad
        idx = wrong.nonzero()
        chans = sorted(set(idx[:, 3].tolist()))
        rows = sorted(set((idx[:, 0] * h * w + idx[:, 1] * w + idx[:, 2]).tolist()))
        print(cfg, "nonfinite", int(bad.sum()), "wrong", int(wrong.sum()), "of", got.numel(), "max diff", float(diff.max()),
              "| channels", chans[:6], "..", chans[-3:] if chans else [], len(chans), "| pixel rows", rows[:6], "..", rows[-3:] if rows else [], len(rows),
              flush=True)


if __name__ == "__main__":
    main()
