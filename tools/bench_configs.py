"""The measurement configurations of SURVEY.md section 8(d) that bench.py does not print (bench.py carries C2 and C3 at batch 32):

  C1  class-default U_Net, 64x64, linear beta 5e-3..9e-3, eps-prediction train step, N = 8, bf16 and tf32 (parity mode)
  C3  class-default U_Net cond_dim 10, 128x128, bf16, per-GPU N in {8, 16, 32}
  C4  super-resolution: in 6 / out 3 / tanh head, cosine schedule, 256x256, target x0 - lr (train_SR_diffusion.py:320-374)
      + cold sampling with skip 20 conditioned on the low-resolution image (generate_sr_images_diffusion.py:224)
  C5  cold diffusion: x0-prediction train step at 128x128, tanh head (train_noise_cold_diffusion.py:330-340)
      + cold_diffusion_sampling skip 20
each for the class-default net (610.7 M parameters) and, for C4/C5, also the config-creator default
(4 levels, 1 residual block, time_dim 512, attention on levels 2-3: 127.3 M parameters).

All train steps replay one CUDA graph (b200.graph.GraphedTrainStep), samplers replay the U_Net graph (net.cuda_graphs);
timing = CUDA events over the replays after warm-up.  One JSON line per measurement.

    python tools/bench_configs.py [--only C1,C4] [--steps 6]
"""
import argparse
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "simple-diffusion-model_b200"))
sys.path.insert(1, ROOT)
import torch  # noqa: E402

CREATOR = dict(num_resnet_blocks=1, num_layers=4, attn_layers=[2, 3], time_dim=512)
DEV = None


def timed(fn, steps, warmup):
    for _ in range(warmup):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / steps


def emit(**kw):
    print(json.dumps(kw), flush=True)


def guarded(fn, *a, **k):
    """One failing configuration must not hide the others' numbers."""
    try:
        fn(*a, **k)
    except Exception as e:  # noqa: BLE001
        emit(config=a[0], failed=f"{type(e).__name__}: {e}"[:400])
        torch.cuda.empty_cache()


def train_case(name, net_kw, size, n, kind, precision, steps, cosine=False, sr=False, labels_dim=0):
    from b200.flops import unet_forward_flops
    from b200.functional import area_resize
    from b200.graph import GraphedTrainStep
    from b200.optim import FusedAdam
    from degraders import CosineNoiseDegradation, NoiseDegradation
    from models.U_Net import U_Net
    torch.manual_seed(0)
    net = U_Net(**net_kw).to(DEV).train().set_precision(precision)
    # same set-up as the trainer (b200/trainer.py): flat parameters + bucket-wise Adam on a second stream under the backward pass
    from b200.parallel import DataParallel
    dp = DataParallel(net, device=DEV)
    opt = FusedAdam(net.parameters(), lr=2e-5, betas=(0.5, 0.999), grad_scale=dp.grad_scale, capturable=True)
    if os.environ.get("SDM_B200_BENCH_ADAM_OVERLAP", "1") == "1":
        dp.attach_optimizer(opt)
    deg = CosineNoiseDegradation(1000) if cosine else NoiseDegradation(5e-3, 9e-3, 1000, device=DEV)
    step = GraphedTrainStep(net, deg, opt, kind=kind)
    g = torch.Generator(device=DEV).manual_seed(1234)
    x0 = torch.rand((n, 3, size, size), device=DEV, generator=g) * 2 - 1
    labels = (torch.rand((n, labels_dim), device=DEV, generator=g) > 0.7).float() if labels_dim else None
    cond_img = target = None
    if sr:
        # train_SR_diffusion.py:321-347: lr = area-down then area-up of x0; the net sees (x_t of x0, x_t of lr at cond_t = 250)
        lr = area_resize(area_resize(x0, size // 4), size)
        cond_img = deg(lr, torch.full((n,), 250, device=DEV, dtype=torch.int64), torch.randn_like(lr))
        target = x0 - lr

    def one():
        eps = torch.randn_like(x0)
        t = torch.randint(1, 1000, (n,), device=DEV)
        return step(x0, t, eps, labels, cond_img, target)

    ms = timed(one, steps, 3)
    loss = float(one())
    in_ch = net_kw.get("in_channel", 3)
    flops = 3.0 * unet_forward_flops(net, size, size, batch=n, tensor_core_only=True)
    emit(config=name, leg="train", kind=kind, precision=precision, size=size, in_channel=in_ch, batch=n,
         params_m=round(sum(p.numel() for p in net.parameters()) / 1e6, 1), ms_per_step=round(ms, 3), img_per_s=round(n / ms * 1e3, 2),
         model_tflops=round(flops / ms / 1e9, 1), launches_per_step=step.launches_per_step, loss=loss,
         peak_mem_gb=round(torch.cuda.max_memory_allocated() / 2 ** 30, 2))
    del step, opt, net
    torch.cuda.empty_cache()
    torch.cuda.reset_peak_memory_stats()


def cold_case(name, net_kw, size, n, precision, reps, sr=False):
    from b200.flops import unet_forward_flops
    from degraders import CosineNoiseDegradation
    from diffusion_sampling_algorithms import cold_diffusion_sampling, skip_schedule
    from models.U_Net import U_Net
    torch.manual_seed(0)
    net = U_Net(**net_kw).to(DEV).eval().set_precision(precision).cuda_graphs(True)
    deg = CosineNoiseDegradation(1000)
    g = torch.Generator(device=DEV).manual_seed(1234)
    noise = torch.randn((n, 3, size, size), device=DEV, generator=g)
    x_start = torch.rand((n, 3, size, size), device=DEV, generator=g) * 2 - 1
    x_t = deg(x_start, torch.tensor([1000], device=DEV), noise)
    cond = torch.rand((n, 3, size, size), device=DEV, generator=g) * 2 - 1 if sr else None
    quiet = lambda *a, **k: None

    def one():
        return cold_diffusion_sampling(net, deg, x_t, noise, 1, 1000, cond_img=cond, skip_step_size=20, device=DEV, log=quiet)

    ms = timed(one, reps, 1)
    evals = len(skip_schedule(1, 1000, 20))
    flops = evals * unet_forward_flops(net, size, size, batch=n, tensor_core_only=True)
    emit(config=name, leg="cold_sampling_skip20", precision=precision, size=size, batch=n, evals=evals,
         ms_per_batch=round(ms, 2), img_per_s=round(n / ms * 1e3, 2), model_tflops=round(flops / ms / 1e9, 1),
         finite=bool(torch.isfinite(one()).all()))
    del net
    torch.cuda.empty_cache()


def main():
    global DEV
    ap = argparse.ArgumentParser()
    ap.add_argument("--only", default="C1,C3,C4,C5")
    ap.add_argument("--steps", type=int, default=6)
    args = ap.parse_args()
    DEV = torch.device("cuda", 0)
    torch.cuda.set_device(DEV)
    want = set(args.only.split(","))
    sr_kw = dict(in_channel=6, out_channel=3, image_recon=True)
    if "C1" in want:
        for prec in ("bf16", "tf32"):
            guarded(train_case, "C1 default 64x64 linear eps", {}, 64, 8, "eps", prec, args.steps)
    if "C1bf16" in want:          # A/B runs: one precision / one batch size
        guarded(train_case, "C1 default 64x64 linear eps", {}, 64, 8, "eps", "bf16", args.steps)
    if "C3b32" in want:
        guarded(train_case, "C3 default cond10 128x128", dict(cond_dim=10), 128, 32, "eps", "bf16", args.steps, labels_dim=10)
    if "C3" in want:
        for n in (8, 16, 32):
            guarded(train_case, "C3 default cond10 128x128", dict(cond_dim=10), 128, n, "eps", "bf16", args.steps, labels_dim=10)
    if "C4" in want:
        for tag, kw, n in (("creator", dict(CREATOR, **sr_kw), 8), ("default", sr_kw, 8)):
            guarded(train_case, f"C4 SR {tag} 256x256", kw, 256, n, "target", "bf16", args.steps, cosine=True, sr=True)
            guarded(cold_case, f"C4 SR {tag} 256x256", kw, 256, n, "bf16", 1, sr=True)
    if "C5" in want:
        for tag, kw in (("creator", dict(CREATOR, image_recon=True)), ("default", dict(image_recon=True))):
            guarded(train_case, f"C5 cold {tag} 128x128", kw, 128, 32, "x0", "bf16", args.steps, cosine=True)
            guarded(cold_case, f"C5 cold {tag} 128x128", kw, 128, 32, "bf16", 1)


if __name__ == "__main__":
    main()
