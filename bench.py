#!/usr/bin/env python
"""Headline benchmark: DDIM "50-step" sampling (51 U-Net evaluations, reference diffusion_sampling_algorithms.py:66-148)
of the class-default 64x64 U-Net under the cosine schedule, batch 256 per GPU (BASELINE.json configs[1]).

    python bench.py --gpus N --steps K --warmup W            # our arm (N > 1 under torchrun, one rank per GPU)
    python bench.py --impl reference --steps K --warmup W    # the reference algorithm on the host CPU (oracle port)

One "step" = one complete sampling of one batch (51 evaluations + 50 updates).  `value` is measured with x_T resident
in HBM; `e2e` goes through the public API from pinned host memory and back.  Sampling shards by image with no
collective, so N GPUs run N independent batches (weak scaling); timing is CUDA events, max over ranks.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(ROOT, "simple-diffusion-model_b200"))
sys.path.insert(1, ROOT)

import torch  # noqa: E402

METRIC = "ddim50_sampled_images_per_s"
UNIT = "img/s"
IMG, BATCH, STEP_SIZE, T_MAX = 64, 256, 20, 1000
TRAIN_IMG, TRAIN_BATCH, TRAIN_COND = 128, 32, 10        # BASELINE.json configs[2]: 128x128 label-conditioned, bf16, DP


def load_peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        with open(path) as f:
            p = json.load(f)
        return p, "measured"
    return {"hbm_gbs": 6650.0, "bf16_tflops": 1590.0, "bf16_tflops_sustained": 1400.0}, "fallback"


class ClockSampler:
    """Samples nvidia-smi clocks / throttle reasons while the timed region runs."""

    QUERY = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
             "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index, self.rows, self.proc = index, [], None

    def __enter__(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), f"--query-gpu={self.QUERY}",
                                          "--format=csv,noheader,nounits", "-lms", "200"], stdout=subprocess.PIPE, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None
        return self

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def __exit__(self, *exc):
        if self.proc is not None:
            self.proc.terminate()
            try:
                self.proc.wait(timeout=5)
            except Exception:
                self.proc.kill()

    def summary(self):
        sm = sorted(int(r[0]) for r in self.rows if r and r[0].isdigit())
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["unavailable"]}
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = [n for i, n in enumerate(names) if any(len(r) > 2 + i and r[2 + i] == "Active" for r in self.rows)]
        mx = max(int(r[1]) for r in self.rows if len(r) > 1 and r[1].isdigit())
        return {"sm_mhz": sm[len(sm) // 2], "sm_max_mhz": mx, "reasons": reasons, "samples": len(sm)}


class KernelTimer:
    """CUDA-event pairs around every launch of the tcgen05 kernel family, on the launching stream."""

    FAMILY = ("b2_conv2d_nhwc", "b2_gemm_nt", "b2_gemm_tn", "b2_attn_scores_softmax")

    def __init__(self):
        self.pairs = []

    def __call__(self, name, args):
        return _Span(self, name in self.FAMILY)

    def total_ms(self):
        return sum(a.elapsed_time(b) for a, b in self.pairs)


class _Span:
    def __init__(self, timer, active):
        self.timer, self.active = timer, active

    def __enter__(self):
        if self.active:
            self.a = torch.cuda.Event(enable_timing=True)
            self.b = torch.cuda.Event(enable_timing=True)
            self.a.record()

    def __exit__(self, *exc):
        if self.active:
            self.b.record()
            self.timer.pairs.append((self.a, self.b))


# ---------------------------------------------------------------------------------------------------- CPU (reference) arm
def cpu_reference_sample(threads, evals=3, batch=4):
    """The reference algorithm (oracle port, fp32, torch CPU) on a bounded sample: `evals` of the 51 DDIM evaluations
    at batch `batch`; per-evaluation cost does not depend on the timestep, so img/s = batch / (t * 51 / evals)."""
    from oracle import diffusion_oracle as orc
    from oracle.weights import synth_state_dict
    torch.set_num_threads(threads)
    state = getattr(cpu_reference_sample, "_state", None)
    if state is None:
        from models.U_Net import U_Net
        with torch.device("meta"):
            shapes = {k: tuple(v.shape) for k, v in U_Net().state_dict().items()}
        state = synth_state_dict(shapes, 0)
        cpu_reference_sample._state = state
    sched = ("cosine", T_MAX)
    x_t = torch.randn((batch, 3, IMG, IMG), generator=torch.Generator().manual_seed(1))
    steps = [1 + STEP_SIZE * (evals - 1 - i) for i in range(evals)]       # ..., 41, 21, 1
    net = lambda x, t, labels=None: orc.unet_forward(state, x, t, None)
    t0 = time.perf_counter()
    with torch.no_grad():
        orc.ddim_sample(net, sched, x_t, 1, steps[0], STEP_SIZE)
    dt = time.perf_counter() - t0
    return batch / (dt * 51.0 / evals), dt, f"{evals} of 51 DDIM evaluations at batch {batch} (class-default U_Net 64x64, fp32), scaled x51/{evals}"


def run_reference_arm(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    threads = os.cpu_count() or 1
    for _ in range(args.warmup):
        cpu_reference_sample(threads, evals=1, batch=2)
    vals, secs = [], 0.0
    sample = ""
    for _ in range(args.steps):
        v, dt, sample = cpu_reference_sample(threads)
        vals.append(v)
        secs += dt
    value = len(vals) / sum(1.0 / v for v in vals)
    line = {"impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": 1000.0 * BATCH / value, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": f"DDIM-50 (51 evals) cosine schedule, class-default U_Net {IMG}x{IMG}, batch {BATCH}/GPU"},
            "cpu_baseline": {"value": value, "unit": UNIT, "cores": threads, "kind": "port", "sample": sample},
            "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}, "gpu_launches": 0}
    print(json.dumps(line), flush=True)



# ---------------------------------------------------------------------------------------------------- train-step leg
def cpu_reference_train(threads, batch=2, img=64):
    """Bounded CPU sample of the reference train step (oracle port): forward + autograd backward + torch Adam at 64x64."""
    from oracle import diffusion_oracle as orc
    from oracle.weights import synth_state_dict
    from models.U_Net import U_Net
    torch.set_num_threads(threads)
    with torch.device("meta"):
        shapes = {k: tuple(v.shape) for k, v in U_Net().state_dict().items()}
    sd = {k: v.requires_grad_(True) for k, v in synth_state_dict(shapes, 0).items()}
    live = [v for k, v in sd.items() if ".y_shift." not in k and not (".attn_layers." in k and ".norm." in k)]
    opt = torch.optim.Adam(live, lr=2e-5, betas=(0.5, 0.999))
    g = torch.Generator().manual_seed(3)
    x0 = torch.rand((batch, 3, img, img), generator=g) * 2 - 1
    eps = torch.randn((batch, 3, img, img), generator=g)
    t = torch.randint(1, 1000, (batch,), generator=g)
    t0 = time.perf_counter()
    opt.zero_grad()
    loss = orc.train_step_loss(sd, ("linear", 5e-3, 9e-3, 1000), x0, t, eps)
    loss.backward()
    opt.step()
    dt = time.perf_counter() - t0
    return batch / dt, dt, f"1 eps-prediction step at batch {batch}, {img}x{img} (class-default U_Net, fp32, torch CPU autograd + Adam)"


def run_train_leg(args, dev, world, rank, barrier, max_over_ranks):
    """Secondary metric of BASELINE.json (configs[2]): eps-prediction DDPM train step (q-sample -> forward -> MSE ->
    backward -> gradient all-reduce -> Adam) of the label-conditioned class-default U-Net at 128x128, bf16, batch-sharded
    data parallel, the whole step replayed as one CUDA graph.  Returns the `train` object of the JSON line."""
    from b200.flops import unet_forward_flops
    from b200.graph import GraphedTrainStep
    from b200.optim import FusedAdam
    from b200.parallel import DataParallel
    from degraders import NoiseDegradation
    from models.U_Net import U_Net

    torch.manual_seed(0)
    net = U_Net(cond_dim=TRAIN_COND).to(dev).train().set_precision(args.precision)
    dp = DataParallel(net, device=dev)
    opt = FusedAdam(net.parameters(), lr=2e-5, betas=(0.5, 0.999), grad_scale=dp.grad_scale, capturable=True)
    dp.attach_optimizer(opt)                 # bucket-wise Adam on a second stream, underneath the backward pass
    deg = NoiseDegradation(5e-3, 9e-3, T_MAX, device=dev)
    step = GraphedTrainStep(net, deg, opt, kind="eps")
    n, s = args.train_batch, TRAIN_IMG
    gen = torch.Generator(device=dev).manual_seed(4321 + rank)
    x0_host = (torch.rand((n, 3, s, s)) * 2 - 1).pin_memory()
    lab_host = (torch.rand((n, TRAIN_COND)) > 0.7).float().pin_memory()
    x0 = x0_host.to(dev)
    labels = lab_host.to(dev)

    def one(h2d):
        if h2d:                                   # e2e: this step's batch arrives from pinned host memory
            x0.copy_(x0_host, non_blocking=True)
            labels.copy_(lab_host, non_blocking=True)
        eps = torch.randn(x0.shape, device=dev, generator=gen)
        t = torch.randint(1, T_MAX, (n,), device=dev, generator=gen)
        return step(x0, t, eps, labels)

    for _ in range(3):
        loss = one(False)
    barrier()
    k = max(args.steps * 3, 8)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(k):
        loss = one(False)
    e1.record()
    barrier()
    ms = max_over_ranks(e0.elapsed_time(e1)) / k
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(k):
        loss = one(True)
        loss_host = float(loss)                   # the reference reads the loss every step (train_diffusion.py:366)
    e1.record()
    barrier()
    ms_e2e = max_over_ranks(e0.elapsed_time(e1)) / k
    flops = 3.0 * unet_forward_flops(net, s, s, batch=n, tensor_core_only=True)
    out = {"metric": "train_images_per_s", "value": world * n / (ms / 1000.0), "unit": "img/s", "ms_per_step": ms,
           "e2e": {"value": world * n / (ms_e2e / 1000.0), "unit": "img/s", "ms_per_step": ms_e2e,
                   "h2d_bytes_per_step": x0_host.numel() * 4 + lab_host.numel() * 4, "d2h_bytes_per_step": 4},
           "config": {"workload": f"eps-prediction DDPM train step, class-default U_Net(cond_dim={TRAIN_COND}) {s}x{s}, "
                                  f"batch {n}/GPU, Adam(0.5, 0.999), data-parallel gradient all-reduce overlapped with backward, "
                                  f"CUDA-graph replay", "global_batch": world * n},
           "dtype": args.precision, "steps": k, "loss": loss_host, "model_tflops_per_gpu": flops / (ms / 1000.0) / 1e12,
           "flops_per_step_per_gpu": flops, "gpu_launches_per_step": step.launches_per_step,
           "peak_mem_gb": torch.cuda.max_memory_allocated() / 2 ** 30}
    del step, opt, dp, net
    torch.cuda.empty_cache()
    return out

# ---------------------------------------------------------------------------------------------------- our arm
def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--batch", type=int, default=BATCH)
    ap.add_argument("--precision", default="bf16", choices=["bf16", "tf32"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-train", action="store_true", help="skip the secondary (train-step) measurement")
    ap.add_argument("--no-graph", action="store_true", help="issue the sampler's U-Net evaluations eagerly (no CUDA graph)")
    ap.add_argument("--train-batch", type=int, default=TRAIN_BATCH)
    args = ap.parse_args()
    if args.impl == "reference":
        return run_reference_arm(args)
    if args.warmup < 3:
        args.warmup = 3

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=dev)

    import b200._lib as b2lib
    import diffusion_sampling_algorithms as S
    from b200.flops import unet_forward_flops
    from degraders import CosineNoiseDegradation
    from models.U_Net import U_Net

    torch.manual_seed(0)
    net = U_Net().to(dev).eval().set_precision(args.precision)
    deg = CosineNoiseDegradation(T_MAX)
    use_graph = not args.no_graph
    quiet = lambda *a, **k: None
    batch = args.batch
    n_evals = len(S.skip_schedule(1, T_MAX, STEP_SIZE))
    flops_step = unet_forward_flops(net, IMG, IMG, batch=batch, tensor_core_only=True) * n_evals

    gen = torch.Generator(device=dev).manual_seed(1234 + rank)
    x_dev = torch.randn((batch, 3, IMG, IMG), device=dev, generator=gen)
    x_host = torch.randn((batch, 3, IMG, IMG)).pin_memory()
    out_host = torch.empty((batch, 3, IMG, IMG)).pin_memory()

    def sample(x):
        return S.ddim_sampling(net, deg, x, min_noise=1, max_noise=T_MAX, ddim_step_size=STEP_SIZE, device=dev, log=quiet)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(ms):
        if world > 1:
            t = torch.tensor([ms], device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            return float(t)
        return ms

    net.cuda_graphs(use_graph)                 # each U-Net evaluation replays one captured graph (b200/graph.py)
    for _ in range(args.warmup):
        sample(x_dev)
    barrier()

    # ---- value: K steps, inputs resident in HBM
    launches0 = b2lib.LAUNCHES
    with ClockSampler(local) as clocks:
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        barrier()
        e0.record()
        for _ in range(args.steps):
            sample(x_dev)
        e1.record()
        barrier()
        ms_total = max_over_ranks(e0.elapsed_time(e1))
    launches = b2lib.LAUNCHES - launches0
    ms_step = ms_total / args.steps
    value = world * batch / (ms_step / 1000.0)

    # ---- e2e: public API from pinned host memory and back, copies inside the timed region
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(args.steps):
        xd = x_host.to(dev, non_blocking=True)
        out = sample(xd)
        out_host.copy_(out, non_blocking=True)
    e1.record()
    barrier()
    ms_e2e = max_over_ranks(e0.elapsed_time(e1)) / args.steps
    e2e_value = world * batch / (ms_e2e / 1000.0)
    finite = bool(torch.isfinite(out_host).all())

    # ---- roofline pass: the same step once more, issued eagerly so that every launch of the tcgen05 kernel family can be
    # bracketed by CUDA events on the launching stream (kernels inside a graph replay cannot be)
    net.cuda_graphs(False)
    sample(x_dev)
    timer = KernelTimer()
    barrier()
    b2lib.set_launch_hook(timer)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    sample(x_dev)
    e1.record()
    barrier()
    b2lib.set_launch_hook(None)
    ms_eager = e0.elapsed_time(e1)
    kern_ms = timer.total_ms()
    n_kern = len(timer.pairs)

    del net
    torch.cuda.empty_cache()
    train = None if args.no_train else run_train_leg(args, dev, world, rank, barrier, max_over_ranks)

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    peaks, peak_src = load_peaks()
    peak_tf = float(peaks.get("bf16_tflops_sustained", peaks["bf16_tflops"]))
    achieved_tf = flops_step / (kern_ms / 1000.0) / 1e12 if kern_ms > 0 else 0.0
    roofline = {"bound": "tensor", "kernel": "igemm_nt_kernel / gemm_tn_kernel (tcgen05 implicit GEMM: conv3x3/convT/linear/softmax(QK^T)/PV)",
                "achieved": achieved_tf, "peak": peak_tf, "unit": "TFLOP/s", "frac": achieved_tf / peak_tf,
                # dram__bytes_read.sum + dram__bytes_write.sum of one launch of the family's heaviest shape (conv3x3 N256 16x16
                # C1024, 1237 GFLOP, algorithmic bytes 287e6) from profiles/r01d_ncu_conv_b256.md (ncu --set full)
                "traffic": 253.9e6 if batch == BATCH else None, "traffic_unit": "bytes/launch (heaviest shape)",
                "peak_source": f"{peak_src} bf16_tflops_sustained (kernel timed inside a long step)",
                "launches_timed": n_kern, "avg_launch_ms": kern_ms / max(n_kern, 1),
                "kernel_share_of_step": kern_ms / ms_eager if ms_eager else None,
                "timed_in": "one extra sampling step issued eagerly (CUDA events around every launch of the family)",
                "eager_ms_per_step": ms_eager,
                "flops_per_step": flops_step, "precision": args.precision}
    line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "bf16" if args.precision == "bf16" else "tf32", "data": "synthetic",
            "config": {"workload": f"DDIM-50 (51 evals, ddim_step_size=20, T=1000) cosine schedule, class-default U_Net "
                                   f"(610.7M params, random init) {IMG}x{IMG} RGB, batch {batch}/GPU, sharded by image",
                       "global_batch": world * batch, "cuda_graph": use_graph, "l2": "working set per evaluation (1.2 GB weights + >250 MB activations per layer) exceeds the 126 MB L2"},
            "clocks": clocks.summary(),
            "e2e": {"value": e2e_value, "unit": UNIT, "ms_per_step": ms_e2e, "h2d_bytes_per_step": x_host.numel() * 4,
                    "d2h_bytes_per_step": out_host.numel() * 4, "finite_output": finite},
            "gpu_launches": launches, "roofline": roofline}
    if train is not None:
        line["train"] = train
    if world == 1 and not args.no_cpu_baseline:
        threads = os.cpu_count() or 1
        v, dt, sample_desc = cpu_reference_sample(threads)
        line["cpu_baseline"] = {"value": v, "unit": UNIT, "cores": threads, "kind": "port", "sample": sample_desc,
                                "seconds": dt}
        if train is not None:
            v, dt, sample_desc = cpu_reference_train(threads)
            train["cpu_baseline"] = {"value": v, "unit": "img/s", "cores": threads, "kind": "port", "sample": sample_desc,
                                     "seconds": dt}
    print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
