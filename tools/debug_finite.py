"""Finds the first non-finite value in a DDIM run of the class-default net (debug aid)."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "simple-diffusion-model_b200"))
import torch  # noqa: E402
import diffusion_sampling_algorithms as S  # noqa: E402
from degraders import CosineNoiseDegradation  # noqa: E402
from models.U_Net import U_Net  # noqa: E402

torch.manual_seed(0)
net = U_Net().cuda().eval()
deg = CosineNoiseDegradation(1000)
n = int(os.environ.get("N", "8"))
x = torch.randn((n, 3, 64, 64)).cuda()
calls = []


class Probe:
    def eval(self):
        return self

    def __call__(self, xin, t, labels=None):
        out = net(xin, t, labels)
        calls.append((int(t[0]), float(xin.abs().max()), float(out.abs().max()), bool(torch.isfinite(out).all())))
        return out


out = S.ddim_sampling(Probe(), deg, x, min_noise=1, max_noise=1000, ddim_step_size=20, device="cuda", log=lambda *a, **k: None)
for c in calls[:6] + calls[-3:]:
    print("t=%d  |x|max=%.3e  |eps|max=%.3e finite=%s" % c)
print("final finite:", bool(torch.isfinite(out).all()), "first bad eval:", next((c[0] for c in calls if not c[3]), None))
if not all(c[3] for c in calls):
    # locate the first layer producing a non-finite activation at the failing step
    bad_t = next(c[0] for c in calls if not c[3])
    from b200 import ops
    import b200.engine as E
    orig = ops.conv2d
    seen = []

    def spy(mode, xx, *a, **k):
        y = orig(mode, xx, *a, **k)
        seen.append(("conv", mode, tuple(xx.shape), bool(torch.isfinite(xx.float()).all()), bool(torch.isfinite(y.float()).all()),
                     float(xx.float().abs().max())))
        return y
    ops.conv2d = spy
    calls.clear()
    try:
        S.ddim_sampling(Probe(), deg, x, min_noise=bad_t, max_noise=1000, ddim_step_size=20, device="cuda", log=lambda *a, **k: None)
    finally:
        ops.conv2d = orig
    per_eval = len(seen) // max(len(calls), 1)
    last = seen[-per_eval:]
    for i, s_ in enumerate(last):
        if not s_[4]:
            print("first non-finite conv output at conv #%d of the eval:" % i, s_, "previous:", last[i - 1] if i else None)
            break
    # attention internals at the failing step
    import b200.engine as E
    orig_attn = E.UNetEngine.attention
    report = []

    def attn_spy(self, blk, xx, out=None, save=None):
        sv = {}
        y = orig_attn(self, blk, xx, out=out, save=sv)
        f = lambda t_: (bool(torch.isfinite(t_.float()).all()), float(t_.float().abs().max()))
        report.append(dict(x=f(xx), qkv=f(sv["qkv"]), pt=f(sv["pt"][..., :xx.shape[1] * xx.shape[2]]), o=f(sv["o"]), y=f(y)))
        if save is not None:
            save.update(sv)
        return y
    E.UNetEngine.attention = attn_spy
    S.ddim_sampling(Probe(), deg, x, min_noise=bad_t, max_noise=1000, ddim_step_size=20, device="cuda", log=lambda *a, **k: None)
    n_attn = len(report) // 2 if len(report) > 30 else len(report)
    for i, r in enumerate(report[-30:]):
        if not r["y"][0] or i < 2:
            print("attention #%d" % i, r)
            if not r["y"][0]:
                break
