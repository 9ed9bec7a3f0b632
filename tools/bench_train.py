"""Training-step throughput (secondary metric of BASELINE.json): class-default U_Net, eps-prediction DDPM step
(q-sample -> forward -> MSE -> backward -> gradient all-reduce -> Adam), per-GPU batch fixed (weak scaling).

    python tools/bench_train.py --batch 8 --size 64 [--cond-dim 10] [--precision bf16]          # 1 GPU
    python -m torch.distributed.run --nproc-per-node N ... tools/bench_train.py ...              # N GPUs, NCCL
"""
import argparse
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "simple-diffusion-model_b200"))
sys.path.insert(1, ROOT)
import torch  # noqa: E402


def run(args):
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=dev)
    import b200._lib as b2lib
    from b200.flops import unet_forward_flops
    from b200.optim import FusedAdam
    from b200.parallel import DataParallel
    from b200.steps import eps_prediction_step
    from degraders import NoiseDegradation
    from models.U_Net import U_Net

    torch.manual_seed(0)
    net = U_Net(cond_dim=args.cond_dim if args.cond_dim > 0 else None).to(dev).train().set_precision(args.precision)
    dp = DataParallel(net, device=dev)
    opt = FusedAdam(net.parameters(), lr=2e-5, betas=(0.5, 0.999), grad_scale=dp.grad_scale, capturable=args.graph)
    if os.environ.get("SDM_B200_OVERLAP_ADAM", "1") != "0":
        dp.attach_optimizer(opt)
    deg = NoiseDegradation(5e-3, 9e-3, 1000, device=dev)
    n, s = args.batch, args.size
    gen = torch.Generator(device=dev).manual_seed(1234 + rank)
    x0 = torch.rand((n, 3, s, s), device=dev, generator=gen) * 2 - 1
    labels = (torch.rand((n, args.cond_dim), device=dev, generator=gen) > 0.7).float() if args.cond_dim > 0 else None

    graphed = None
    if args.graph:
        from b200.graph import GraphedTrainStep
        graphed = GraphedTrainStep(net, deg, opt, kind="eps")

    def step():
        eps = torch.randn_like(x0)
        t = torch.randint(1, 1000, (n,), device=dev)
        if graphed is not None:
            return graphed(x0, t, eps, labels)
        return eps_prediction_step(net, deg, opt, x0, t, eps, labels)

    for _ in range(args.warmup):
        loss = step()
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    l0 = b2lib.LAUNCHES
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    import time
    wall = []
    for _ in range(args.steps):
        t0 = time.perf_counter()
        loss = step()
        wall.append(round(1e3 * (time.perf_counter() - t0), 1))
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / args.steps
    if world > 1:
        tt = torch.tensor([ms], device=dev)
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        ms = float(tt)
    if rank == 0:
        flops = 3.0 * unet_forward_flops(net, s, s, batch=n, tensor_core_only=True)
        print(json.dumps({"metric": "train_images_per_s", "value": world * n / (ms / 1000.0), "unit": "img/s", "n_gpus": world,
                          "ms_per_step": ms, "per_gpu_batch": n, "size": s, "precision": args.precision, "cond_dim": args.cond_dim,
                          "loss": float(loss), "model_tflops_per_gpu": flops / (ms / 1000.0) / 1e12,
                          "gpu_launches_per_step": (b2lib.LAUNCHES - l0) // args.steps, "scaling": "weak",
                          "graph": bool(args.graph), "peak_mem_gb": torch.cuda.max_memory_allocated() / 2 ** 30, "host_ms_per_step": wall}), flush=True)
    if world > 1:
        # the captured graph holds NCCL kernels: release it before the communicator (destroying the group first hangs)
        del graphed, opt, dp, net
        torch.cuda.synchronize()
        dist.destroy_process_group()


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--batch", type=int, default=8)
    ap.add_argument("--size", type=int, default=64)
    ap.add_argument("--cond-dim", type=int, default=0)
    ap.add_argument("--precision", default="bf16")
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--graph", action="store_true", help="replay the whole step as one CUDA graph (b200/graph.py)")
    run(ap.parse_args())
