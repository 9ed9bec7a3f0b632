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
REF_DIR = os.path.join(ROOT, "baseline", "_ref")            # verbatim copy of the unmodified reference (__graft_entry__.install_reference)
_REF_ARM = any(a in ("reference", "torch-eager") or a.endswith("=reference") or a.endswith("=torch-eager") for a in sys.argv[1:]) \
    and os.path.isfile(os.path.join(REF_DIR, "models", "U_Net.py"))
if _REF_ARM:
    # the reference arms import the reference's OWN modules (models.U_Net, degraders, diffusion_sampling_algorithms): this repo's
    # same-named package must not be importable in that process -- none of our models, kernels or engine on that path
    sys.path.insert(0, REF_DIR)
    sys.path.insert(1, ROOT)
else:
    sys.path.insert(0, os.path.join(ROOT, "simple-diffusion-model_b200"))
    sys.path.insert(1, ROOT)

import torch  # noqa: E402

WORKLOAD = ("DDIM-50 (51 evals, ddim_step_size=20, T=1000) cosine schedule, class-default U_Net (610.7M params, random init) "
            "64x64 RGB, batch {batch}/GPU, sharded by image")
L2_NOTE = "inputs larger than L2: working set per evaluation (1.2 GB weights + >250 MB activations per layer) exceeds the 126 MB L2"
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


def load_traffic(batch):
    """roofline.traffic = dram__bytes_read.sum + dram__bytes_write.sum per launch of the family's heaviest shape, READ from the
    committed summary of an `ncu --set full` capture (profiles/roofline_traffic.json, written by tools/ncu_summary.py --traffic);
    null when no capture of this batch size is on file -- never a constant in this script."""
    path = os.path.join(ROOT, "profiles", "roofline_traffic.json")
    try:
        with open(path) as f:
            t = json.load(f)
        if int(t.get("batch", -1)) == int(batch):
            return {"traffic": float(t["dram_bytes_per_launch"]), "traffic_unit": "bytes/launch (heaviest shape)",
                    "traffic_kernel": t.get("kernel"), "traffic_algorithmic_bytes": t.get("algorithmic_bytes"),
                    "traffic_source": t.get("source")}
    except (OSError, ValueError, KeyError):
        pass
    return {"traffic": None}


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


REF_BATCH, REF_EVALS = 8, 3        # bounded sample of the reference arm: 3 of the 51 DDIM evaluations at batch 8


def ref_reference_sample(threads, evals=REF_EVALS, batch=REF_BATCH):
    """The UNMODIFIED reference (baseline/_ref: its U_Net, its CosineNoiseDegradation, its ddim_sampling) on the host CPU, all
    threads, on a bounded sample of the bench workload: the last `evals` steps of the DDIM-50 schedule (t = ..., 41, 21, 1) at
    batch `batch` through the reference's own public sampler.  Per-evaluation cost does not depend on t, so
    img/s = batch / (seconds * 51 / evals)."""
    from degraders import CosineNoiseDegradation          # reference modules (sys.path[0] == baseline/_ref)
    from diffusion_sampling_algorithms import ddim_sampling
    from models.U_Net import U_Net
    torch.set_num_threads(threads)
    net = getattr(ref_reference_sample, "_net", None)
    if net is None:
        torch.manual_seed(0)
        net = U_Net().eval()                              # class-default 610.7 M net, reference initialisation
        ref_reference_sample._net = net
    x_t = torch.randn((batch, 3, IMG, IMG), generator=torch.Generator().manual_seed(1))
    top = 1 + STEP_SIZE * (evals - 1)
    t0 = time.perf_counter()
    out = ddim_sampling(net, CosineNoiseDegradation(T_MAX), x_t, min_noise=1, max_noise=top, ddim_step_size=STEP_SIZE,
                        device="cpu", log=lambda *a, **k: None)
    dt = time.perf_counter() - t0
    assert tuple(out.shape) == (batch, 3, IMG, IMG)
    return batch / (dt * 51.0 / evals), dt, (f"unmodified reference (baseline/_ref) ddim_sampling: {evals} of 51 DDIM evaluations at batch {batch} "
                                             f"(class-default U_Net {IMG}x{IMG}, fp32, torch CPU), scaled x51/{evals}")


def run_reference_arm(args):
    """`--impl reference`: the reference's own CPU implementation of the path on the box's host cores.  kind "reference" when
    baseline/_ref is present (the normal case), else the oracle port."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    threads = os.cpu_count() or 1
    real = _REF_ARM
    sampler = ref_reference_sample if real else cpu_reference_sample
    for _ in range(max(args.warmup, 1)):
        sampler(threads, evals=1, batch=2)
    vals, secs, sample = [], [], ""
    for _ in range(args.steps):
        v, dt, sample = sampler(threads)
        vals.append(v)
        secs.append(dt)
    value = len(vals) / sum(1.0 / v for v in vals)
    line = {"impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
            "warmup": args.warmup,
            # one timed step = one bounded sample (see cpu_baseline.sample); the full 51-evaluation batch-256 sampling would take
            # ms_per_full_step on these cores
            "ms_per_step": 1000.0 * sum(secs) / len(secs), "ms_per_full_step": 1000.0 * BATCH / value,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            # the same workload, metric and unit as the b200 arm; what one timed step actually ran is in cpu_baseline.sample
            "config": {"workload": WORKLOAD.format(batch=BATCH), "global_batch": BATCH * max(args.gpus, 1), "l2": L2_NOTE},
            "cpu_baseline": {"value": value, "unit": UNIT, "cores": threads, "kind": "reference" if real else "port", "sample": sample},
            "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}, "gpu_launches": 0}
    print(json.dumps(line), flush=True)


# ---------------------------------------------------------------------------------------------------- torch eager on the B200
def run_torch_eager_arm(args):
    """`--impl torch-eager`: the UNMODIFIED reference modules (baseline/_ref) executed by PyTorch eager on the B200 -- cuDNN /
    cuBLAS / ATen, the only pre-existing Blackwell kernels for this path -- on the bench's two workloads:
    the reference's own `ddim_sampling` (51 evaluations, batch 256, 64x64, cosine) and the reference's train-step body
    (train_diffusion.py:310-366: randn_like, randint, q-sample, U_Net, mse_loss, GradScaler backward/step, Adam(0.5, 0.999)) on
    the label-conditioned net at 128x128.  Two precisions each: the reference's own mode on CUDA, autocast (its default
    dtype is fp16; bf16 is measured too because that is what this repo computes in), and fp32 with TF32 enabled.  CUDA
    events, 1 warm-up sampling / 3 warm-up steps.  Prints one JSON line."""
    import torch.nn.functional as F
    from degraders import CosineNoiseDegradation, NoiseDegradation      # reference modules
    from diffusion_sampling_algorithms import ddim_sampling
    from models.U_Net import U_Net
    dev = torch.device("cuda", 0)
    torch.cuda.set_device(dev)
    torch.backends.cuda.matmul.allow_tf32 = True
    torch.backends.cudnn.allow_tf32 = True
    quiet = lambda *a, **k: None
    out = {"impl": "torch-eager", "torch": torch.__version__, "cudnn": torch.backends.cudnn.version(),
           "note": "unmodified reference modules from baseline/_ref, PyTorch eager on the same B200; TF32 enabled for fp32"}

    def timed(fn, warm, reps):
        for _ in range(warm):
            fn()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(reps):
            fn()
        e1.record()
        torch.cuda.synchronize()
        return e0.elapsed_time(e1) / reps

    # ---- DDIM-50, batch 256, 64x64, cosine (BASELINE configs[1])
    torch.manual_seed(0)
    net = U_Net().to(dev).eval()
    deg = CosineNoiseDegradation(T_MAX)
    batch = args.batch
    x_T = torch.randn((batch, 3, IMG, IMG), device=dev)
    ddim = {}
    for mode in ("bf16_autocast", "fp16_autocast", "fp32_tf32"):
        def sample():
            if mode == "fp32_tf32":
                return ddim_sampling(net, deg, x_T, min_noise=1, max_noise=T_MAX, ddim_step_size=STEP_SIZE, device=dev, log=quiet)
            with torch.autocast("cuda", dtype=torch.bfloat16 if mode == "bf16_autocast" else torch.float16):
                return ddim_sampling(net, deg, x_T, min_noise=1, max_noise=T_MAX, ddim_step_size=STEP_SIZE, device=dev, log=quiet)
        try:
            # warm-up: a short schedule (3 evaluations) is enough to settle cuDNN's heuristics and the allocator
            with torch.autocast("cuda", dtype=torch.bfloat16, enabled=(mode == "bf16_autocast")):
                ddim_sampling(net, deg, x_T, min_noise=1, max_noise=41, ddim_step_size=STEP_SIZE, device=dev, log=quiet)
            ms = timed(sample, 0, max(1, min(args.steps, 2)))
            ddim[mode] = {"img_per_s": batch / (ms / 1000.0), "ms_per_step": ms}
        except Exception as e:      # noqa: BLE001  (an OOM here must not lose the other numbers)
            ddim[mode] = {"error": str(e)[:200]}
            torch.cuda.empty_cache()
    out["ddim50"] = {"workload": WORKLOAD.format(batch=batch), "unit": "img/s", **ddim}
    del net
    torch.cuda.empty_cache()

    # ---- train step, 128x128, cond_dim 10 (BASELINE configs[2])
    train = {}
    for mode in ("bf16_autocast", "fp16_autocast", "fp32_tf32"):
        n = args.train_batch
        while n >= 4:
            try:
                torch.manual_seed(0)
                net = U_Net(cond_dim=TRAIN_COND).to(dev).train()
                opt = torch.optim.Adam(net.parameters(), lr=2e-5, betas=(0.5, 0.999))
                scaler = torch.amp.GradScaler("cuda", enabled=(mode == "fp16_autocast"))
                degl = NoiseDegradation(5e-3, 9e-3, T_MAX, dev)
                x0 = torch.rand((n, 3, TRAIN_IMG, TRAIN_IMG), device=dev) * 2 - 1
                labels = (torch.rand((n, TRAIN_COND), device=dev) > 0.7).float()

                def step():
                    noise = torch.randn_like(x0)
                    opt.zero_grad()
                    t = torch.randint(low=1, high=T_MAX, size=(n,), device=dev)
                    with torch.autocast("cuda", dtype=torch.float16 if mode == "fp16_autocast" else torch.bfloat16,
                                        enabled=(mode != "fp32_tf32")):
                        x_t = degl(img=x0, steps=t, eps=noise)
                        loss = F.mse_loss(net(x_t, t, labels), noise)
                    scaler.scale(loss).backward()
                    scaler.step(opt)
                    scaler.update()
                    return loss.item()                       # the reference reads the loss every step (train_diffusion.py:366)

                ms = timed(step, 3, 5)
                train[mode] = {"img_per_s": n / (ms / 1000.0), "ms_per_step": ms, "batch": n,
                               "peak_mem_gb": torch.cuda.max_memory_allocated() / 2 ** 30}
                break
            except torch.OutOfMemoryError:
                n //= 2
                train[mode] = {"error": "out of memory", "batch_tried": n * 2}
            finally:
                net = opt = None
                torch.cuda.empty_cache()
                torch.cuda.reset_peak_memory_stats()
    out["train"] = {"workload": f"eps-prediction train step, class-default U_Net(cond_dim={TRAIN_COND}) {TRAIN_IMG}x{TRAIN_IMG}, "
                                f"Adam(0.5, 0.999), torch eager", "unit": "img/s", **train}
    print(json.dumps(out), flush=True)


def torch_eager_subprocess(args, timeout=900):
    """Runs `bench.py --impl torch-eager` in a fresh process (the reference's module names clash with this repo's) once this
    process has released the GPU; returns its JSON object, or {"unavailable": why}."""
    if not os.path.isfile(os.path.join(REF_DIR, "models", "U_Net.py")):
        return {"unavailable": "baseline/_ref missing (run __graft_entry__.build() where /root/reference exists)"}
    cmd = [sys.executable, os.path.abspath(__file__), "--impl", "torch-eager", "--steps", str(args.steps), "--batch", str(args.batch),
           "--train-batch", str(args.train_batch)]
    env = {k: v for k, v in os.environ.items() if k not in ("RANK", "LOCAL_RANK", "WORLD_SIZE", "MASTER_ADDR", "MASTER_PORT")}
    try:
        res = subprocess.run(cmd, capture_output=True, text=True, timeout=timeout, env=env)
    except subprocess.TimeoutExpired:
        return {"unavailable": f"torch-eager arm exceeded {timeout} s"}
    for ln in reversed(res.stdout.strip().splitlines()):
        if ln.startswith("{"):
            try:
                return json.loads(ln)
            except json.JSONDecodeError:
                break
    return {"unavailable": "torch-eager arm printed no JSON", "stderr_tail": res.stderr[-300:]}


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
    n, s = args.train_batch, TRAIN_IMG
    # eps is drawn inside the q-sample kernel and re-drawn inside the loss kernel (Philox keyed on the step count held in device
    # memory and the global element index): no RNG launch, no eps tensor (north_star; reference train_diffusion.py:310)
    step = GraphedTrainStep(net, deg, opt, kind="eps", philox_seed=4321, philox_first_elem=rank * n * 3 * s * s)
    gen = torch.Generator(device=dev).manual_seed(4321 + rank)
    x0_host = (torch.rand((n, 3, s, s)) * 2 - 1).pin_memory()
    lab_host = (torch.rand((n, TRAIN_COND)) > 0.7).float().pin_memory()
    x0 = x0_host.to(dev)
    labels = lab_host.to(dev)

    def one(h2d):
        if h2d:                                   # e2e: this step's batch arrives from pinned host memory
            x0.copy_(x0_host, non_blocking=True)
            labels.copy_(lab_host, non_blocking=True)
        t = torch.randint(1, T_MAX, (n,), device=dev, generator=gen)
        return step(x0, t, None, labels)

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
    ap.add_argument("--impl", default="b200", choices=["b200", "reference", "torch-eager"])
    ap.add_argument("--no-eager", action="store_true", help="skip the torch-eager-on-B200 comparison (a subprocess after the timed legs)")
    ap.add_argument("--batch", type=int, default=BATCH)
    ap.add_argument("--precision", default="bf16", choices=["bf16", "tf32"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-train", action="store_true", help="skip the secondary (train-step) measurement")
    ap.add_argument("--no-graph", action="store_true", help="issue the sampler's U-Net evaluations eagerly (no CUDA graph)")
    ap.add_argument("--train-batch", type=int, default=TRAIN_BATCH)
    args = ap.parse_args()
    if args.impl == "reference":
        return run_reference_arm(args)
    if args.impl == "torch-eager":
        return run_torch_eager_arm(args)
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
                **load_traffic(batch),
                "peak_source": f"{peak_src} bf16_tflops_sustained (kernel timed inside a long step)",
                "launches_timed": n_kern, "avg_launch_ms": kern_ms / max(n_kern, 1),
                "kernel_share_of_step": kern_ms / ms_eager if ms_eager else None,
                "timed_in": "one extra sampling step issued eagerly (CUDA events around every launch of the family)",
                "eager_ms_per_step": ms_eager,
                "flops_per_step": flops_step, "precision": args.precision}
    line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "bf16" if args.precision == "bf16" else "tf32", "data": "synthetic",
            "config": {"workload": WORKLOAD.format(batch=batch), "global_batch": world * batch, "l2": L2_NOTE},
            "execution": {"cuda_graph": use_graph},
            "clocks": clocks.summary(),
            "e2e": {"value": e2e_value, "unit": UNIT, "ms_per_step": ms_e2e, "h2d_bytes_per_step": x_host.numel() * 4,
                    "d2h_bytes_per_step": out_host.numel() * 4, "finite_output": finite},
            "gpu_launches": launches, "roofline": roofline}
    if train is not None:
        line["train"] = train
    if world == 1 and not args.no_eager:
        # the bar that matters on this hardware: the unmodified reference through PyTorch eager (cuDNN / cuBLAS) on the same B200
        eager = torch_eager_subprocess(args)
        line["torch_eager_b200"] = eager
        try:
            best = max(v["img_per_s"] for v in eager["ddim50"].values() if isinstance(v, dict) and "img_per_s" in v)
            line["torch_eager_b200"]["speedup_ddim50_vs_best_eager"] = value / best
            if train is not None:
                best_t = max(v["img_per_s"] for v in eager["train"].values() if isinstance(v, dict) and "img_per_s" in v)
                line["torch_eager_b200"]["speedup_train_vs_best_eager"] = train["value"] / best_t
        except (KeyError, ValueError, TypeError):
            pass
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
