"""Shared driver behind the four training entry points (reference train_diffusion.py, train_noise_cold_diffusion.py,
train_SR_diffusion.py, train_doodle_diffusion.py).  The reference repeats ~470 lines per script; they differ only in
how the network input / regression target are assembled and in which sampler draws the progress plots:

    flavour   input to the U-Net                          target        plots            reference step body
    base      x_t                                         eps           DDPM / DDIM      train_diffusion.py:295-366
    cold      x_t                                         x0            cold sampler     train_noise_cold_diffusion.py:290-356
    sr        cat(x_t, q(area_up(area_down(x0)), cond_t)) x0 - lr       cold sampler     train_SR_diffusion.py:302-388
    doodle    cat(x_t, conditioning image)                eps           DDPM / DDIM      train_doodle_diffusion.py:269-331

Config JSON keys, validation errors, log lines, checkpoint files ({"model","optimizer"} / {"starting_epoch",
"global_steps"[,"beta_1","beta_T"]}), the learning-rate halving and the order of the random draws follow the reference.
What changes is the execution: every step is the fused kernel sequence of b200/steps.py, replayed as one CUDA graph
(b200/graph.py), with FusedAdam and -- under torchrun -- batch-sharded data parallelism (b200/parallel.py).
"""
import argparse
import csv
import glob
import json
import logging
import os
import pathlib

import torch

from ._lib import B200Error
from .functional import area_resize
from .graph import GraphedTrainStep
from .optim import FusedAdam
from .parallel import DataParallel, shard_range

FLAVOURS = {
    "base": dict(project="Diffusion", kind="eps", sampler="ddpm_ddim"),
    "cold": dict(project="Noise-Cold-Diffusion", kind="x0", sampler="cold"),
    "sr": dict(project="SR-Diffusion", kind="target", sampler="cold"),
    "doodle": dict(project="Doodle-Diffusion", kind="eps", sampler="ddpm_ddim"),
}


def _parse(raw_args, description):
    parser = argparse.ArgumentParser(description=description)
    parser.add_argument("-c", "--config-path", help="File path to load json config file.", required=True, type=pathlib.Path)
    parser.add_argument("--device", help="Hardware device (this build runs on CUDA only).", choices=["cpu", "cuda"], type=str,
                        default="cuda")
    parser.add_argument("--max-steps", help="Stop after this many optimisation steps (addition: smoke tests).", type=int,
                        default=None)
    return vars(parser.parse_args(raw_args))


def _dataset(flavour, cfg, shuffle_seed=None, raw_uint8=False):
    """`shuffle_seed`: under torchrun every rank must map DistributedSampler's indices onto the SAME row order, so the one-off
    table shuffle of the labelled datasets is seeded (the reference is single-process and shuffles unseeded)."""
    path = cfg["dataset_path"]
    if path is None:
        raise ValueError("No dataset_path entered.")
    if isinstance(path, str) and path.startswith("synthetic:"):
        from custom_dataset.img_dataset import SyntheticImages
        return SyntheticImages(path, raw_uint8=raw_uint8)
    if flavour == "doodle":
        from custom_dataset.doodle_dataset import DoodleImgDataset
        return DoodleImgDataset(dataset_path=path, shuffle_seed=shuffle_seed, raw_uint8=raw_uint8)
    if cfg.get("use_conditional"):
        from custom_dataset.conditional_img_dataset import ConditionalImgDataset
        return ConditionalImgDataset(dataset_path=path, shuffle_seed=shuffle_seed, raw_uint8=raw_uint8)
    from custom_dataset.img_dataset import ImageDataset
    img_list = glob.glob(path)
    if len(img_list) <= 0:
        raise Exception("No dataset found!")
    return ImageDataset(img_paths=img_list, raw_uint8=raw_uint8)


def rank_seed(base_seed, rank):
    """Seed of torch's CPU and CUDA generators on `rank` (torch.manual_seed seeds both)."""
    return int(base_seed) + int(rank)


def _flip_per_image(x):
    """torchvision RandomHorizontalFlip(p=0.5) applied image by image, as the reference does
    (train_diffusion.py:312-314): one CPU-generator draw per image, flipped by one kernel on the device."""
    from .image_io import draw_flip_flags, flip_images
    return flip_images(x, draw_flip_flags(x.shape[0]))


def run_training(flavour, raw_args=None):
    import diffusion_sampling_algorithms as samplers
    from degraders import CosineNoiseDegradation, NoiseDegradation
    from diffusion_enums import DiffusionAlg, NoiseScheduler
    from models.U_Net import U_Net
    from utils.utils import load_checkpoint, plot_sampled_images, save_model

    spec = FLAVOURS[flavour]
    args = _parse(raw_args, f"Train {spec['project']} models.")
    if args["device"] != "cuda":
        raise B200Error("this build has no CPU path: run the reference implementation for --device cpu")
    with open(args["config_path"], "r") as f:
        cfg = json.load(f)

    starting_epoch, global_steps = 0, 0
    checkpoint_steps, lr_steps, max_epoch = cfg["checkpoint_steps"], cfg["lr_steps"], cfg["max_epoch"]
    plot_img_count, flip_imgs = cfg["plot_img_count"], cfg.get("flip_imgs", False)
    use_conditional = cfg.get("use_conditional", False) and flavour != "doodle"
    out_dir = cfg["out_dir"]

    if cfg["noise_scheduler"] == "LINEAR":
        scheduling, beta_1, beta_T = NoiseScheduler.LINEAR, cfg["beta1"], cfg["betaT"]
    elif cfg["noise_scheduler"] == "COSINE":
        scheduling, beta_1, beta_T = NoiseScheduler.COSINE, None, None
    else:
        raise ValueError("Invalid noise scheduler type.")
    diffusion_alg = None
    if spec["sampler"] == "ddpm_ddim":
        if cfg["diffusion_alg"] == "DDIM":
            diffusion_alg = DiffusionAlg.DDIM
        elif cfg["diffusion_alg"] == "DDPM":
            diffusion_alg = DiffusionAlg.DDPM
        else:
            raise ValueError("Invalid diffusion algorithm type.")
    min_t, max_t, max_actual_t, skip = cfg["min_noise_step"], cfg["max_noise_step"], cfg["max_actual_noise_step"], cfg["skip_step"]
    if max_actual_t < min_t or max_t < min_t or skip > max_actual_t or skip < 0 or min_t < 0:
        raise ValueError("Invalid step values entered!")
    if flavour == "sr":
        lr_dim, sr_dim, cond_t = cfg["lr_dim"], cfg["sr_dim"], cfg["cond_t"]
        if lr_dim > sr_dim:
            raise ValueError("Invalid low / super resolution dimensions.")

    # configuration is valid: only now touch the device / process group / file system
    os.makedirs(out_dir, exist_ok=True)
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    device = torch.device("cuda", local)
    if world > 1 and not torch.distributed.is_initialized():
        torch.distributed.init_process_group("nccl", device_id=device)
    # Random streams.  Single process: untouched, like the reference (seed it from outside for reproducibility).  Data
    # parallel: every rank needs its OWN stream for eps / t / flips -- otherwise the global batch repeats one (t, eps) draw
    # world-size times -- and the SAME dataset row order (rank_seed / shuffle_seed below); "seed" in the JSON is an addition.
    base_seed = cfg.get("seed")
    shuffle_seed = None
    if world > 1 or base_seed is not None:
        base_seed = int(base_seed if base_seed is not None else 0)
        torch.manual_seed(rank_seed(base_seed, rank))
        shuffle_seed = base_seed

    log = logging.getLogger(f"b200.{flavour}.{rank}")
    log.setLevel(logging.DEBUG)
    log.handlers.clear()
    log.propagate = False
    if rank == 0:
        fmt = logging.Formatter("%(asctime)s %(message)s")
        for h in (logging.FileHandler(os.path.join(out_dir, f"{spec['project']}.log")), logging.StreamHandler()):
            h.setFormatter(fmt)
            log.addHandler(h)
    else:
        log.addHandler(logging.NullHandler())

    # "gpu_input_pipeline" (addition, default on): the dataset hands out the decoded uint8 HWC bytes; normalisation to [-1, 1],
    # HWC -> CHW and the random flips are one kernel on the device, fed by a pinned double-buffered prefetcher
    # (b200/image_io.py) -- bit-identical to the reference's host-side numpy arithmetic
    raw_input = bool(cfg.get("gpu_input_pipeline", True))
    dataset = _dataset(flavour, cfg, shuffle_seed, raw_uint8=raw_input)
    batch_size = cfg["batch_size"]
    sampler = torch.utils.data.distributed.DistributedSampler(dataset, world, rank, shuffle=True) if world > 1 else None
    loader = torch.utils.data.DataLoader(dataset, batch_size=batch_size, num_workers=cfg.get("num_workers", 4),
                                         shuffle=sampler is None, sampler=sampler, pin_memory=True, drop_last=world > 1)
    plot_batch = next(iter(torch.utils.data.DataLoader(dataset, batch_size=max(plot_img_count, 1), num_workers=0, shuffle=False)))
    if raw_input:
        from .image_io import DeviceImageLoader, draw_flip_flags, u8_to_image
        as_float = lambda t: u8_to_image(t.to(device)) if torch.is_tensor(t) and t.dtype == torch.uint8 else t
        plot_batch = [as_float(t) for t in plot_batch] if isinstance(plot_batch, (list, tuple)) else as_float(plot_batch)
        # the flips ride in the conversion kernel; their per-image CPU draws happen when the batch is handed out, in order
        flips_here = flip_imgs and flavour != "doodle"
        batches = DeviceImageLoader(loader, device, flip_fn=draw_flip_flags if flips_here else None)
    else:
        batches, flips_here = loader, False
    plot_labels, plot_cond = None, None
    if flavour == "doodle":
        plot_imgs, plot_cond = plot_batch
    elif use_conditional:
        plot_imgs, plot_labels = plot_batch
        if rank == 0:
            with open(os.path.join(out_dir, "labels.txt"), "a") as f:
                csv.writer(f).writerows([dataset.get_labels()] + plot_labels.cpu().tolist())
    else:
        plot_imgs = plot_batch[0] if isinstance(plot_batch, (list, tuple)) else plot_batch

    net = U_Net(in_channel=cfg["in_channel"], out_channel=cfg["out_channel"], num_layers=cfg["num_layers"],
                num_resnet_blocks=cfg["num_resnet_block"], attn_layers=cfg["attn_layers"], num_heads=cfg["attn_heads"],
                dim_per_head=cfg["attn_dim_per_head"], time_dim=cfg["time_dim"], cond_dim=cfg["cond_dim"],
                min_channel=cfg["min_channel"], max_channel=cfg["max_channel"], image_recon=cfg["img_recon"])
    ckpt = None
    if cfg.get("model_checkpoint") is not None:
        ok, ckpt = load_checkpoint(cfg["model_checkpoint"])
        if not ok:
            raise Exception("An error occured while loading model checkpoint!")
        net.custom_load_state_dict(ckpt["model"])
    net = net.to(device).set_precision(cfg.get("precision", "bf16"))
    dp = DataParallel(net, device=device)            # flattens the parameters; a no-op collective-wise when world == 1
    # "grad_scaler": true (addition) runs the reference's GradScaler protocol (train_diffusion.py:130, 358-364) around the
    # eager step; it is numerically a no-op in bf16 and costs the CUDA-graph replay, hence off by default
    scaler = torch.amp.GradScaler("cuda") if cfg.get("grad_scaler") else None
    use_graph = os.environ.get("SDM_B200_CUDA_GRAPH", "1") != "0" and scaler is None
    optim = FusedAdam(net.parameters(), lr=cfg["diffusion_lr"], betas=(0.5, 0.999), grad_scale=dp.grad_scale, capturable=use_graph)
    if scaler is None:
        dp.attach_optimizer(optim)
    if ckpt is not None and cfg.get("load_diffusion_optim"):
        optim.load_state_dict(ckpt["optimizer"])
    if cfg.get("config_checkpoint") is not None:
        ok, cc = load_checkpoint(cfg["config_checkpoint"])
        if not ok:
            raise Exception("An error occured while loading config checkpoint!")
        if scheduling == NoiseScheduler.LINEAR:
            beta_1, beta_T = cc["beta_1"], cc["beta_T"]
        starting_epoch, global_steps = cc["starting_epoch"], cc["global_steps"]

    degrader = NoiseDegradation(beta_1, beta_T, max_t, device) if scheduling == NoiseScheduler.LINEAR else CosineNoiseDegradation(max_t)
    graphed = None                      # built at the first batch (the Philox stream needs the per-rank element count)
    # "philox_noise": <seed> (addition): eps is drawn inside the q-sample kernel -- and re-drawn inside the loss kernel for the
    # eps-prediction flavours -- from Philox keyed on (seed, optimisation step, global element index) instead of by a separate
    # torch.randn_like launch.  Off by default: the default keeps the reference's RNG stream.
    philox_seed = cfg.get("philox_noise")
    if philox_seed is True:
        philox_seed = int(cfg.get("seed") or 0)

    log.info("#" * 100)
    log.info("Train Parameters:")
    log.info(f"Max Epoch: {max_epoch:,}")
    log.info(f"Dataset Path: {cfg['dataset_path']}")
    log.info(f"Output Path: {out_dir}")
    log.info(f"Checkpoint Steps: {checkpoint_steps}")
    log.info(f"Batch size: {batch_size:,} (per GPU; {world} GPU(s))")
    log.info(f"Diffusion LR: {optim.param_groups[0]['lr']:.5f}")
    log.info(f"Using Conditional Info.: {use_conditional}")
    log.info(f"Image Augmentation (Random Horizontal Flip): {flip_imgs}")
    log.info("#" * 100)
    log.info("Model Parameters:")
    for key in ("in_channel", "out_channel", "num_layers", "num_resnet_block", "attn_layers", "attn_heads", "attn_dim_per_head",
                "time_dim", "cond_dim", "min_channel", "max_channel", "img_recon"):
        log.info(f"{key}: {cfg[key]}")
    log.info("#" * 100)
    log.info("Diffusion Parameters:")
    if scheduling == NoiseScheduler.LINEAR:
        log.info(f"Beta_1: {beta_1:,.5f}")
        log.info(f"Beta_T: {beta_T:,.5f}")
    log.info(f"Min Noise Step: {min_t:,}")
    log.info(f"Max Noise Step: {max_t:,}")
    log.info(f"Max Actual Noise Step: {max_actual_t:,}")
    log.info("#" * 100)

    def checkpoint(steps):
        if rank != 0:
            return
        state = {"starting_epoch": starting_epoch, "global_steps": steps}
        if scheduling == NoiseScheduler.LINEAR:
            state["beta_1"], state["beta_T"] = beta_1, beta_T
        save_model(model_net=state, file_name="config", dest_path=out_dir, checkpoint=True, steps=steps)
        save_model(model_net={"model": net.state_dict(), "optimizer": optim.state_dict()}, file_name="diffusion",
                   dest_path=out_dir, checkpoint=True, steps=steps)

    def lr_condition(x0):
        """Noised low-resolution conditioning of the SR flavour (train_SR_diffusion.py:320-364, generate_sr:170-178)."""
        lr = area_resize(area_resize(x0, (lr_dim, lr_dim)), (sr_dim, sr_dim))
        return lr

    def plot(steps, shape):
        if rank != 0 or plot_img_count <= 0:
            return
        n, c, h, w = shape
        labels = plot_labels.to(device) if plot_labels is not None else None
        cond = plot_cond.to(device) if plot_cond is not None else None
        imgs = plot_imgs.to(device)
        if spec["sampler"] == "ddpm_ddim":
            if max_actual_t < max_t:
                x_t = degrader(img=imgs, steps=torch.tensor([max_actual_t]), eps=torch.randn((imgs.shape[0], c, h, w), device=device))
            else:
                x_t = torch.randn((plot_img_count, c, h, w), device=device)
            if diffusion_alg == DiffusionAlg.DDPM:
                out = samplers.ddpm_sampling(net, degrader, x_t, min_noise=min_t, max_noise=max_actual_t, cond_img=cond,
                                             labels_tensor=labels, device=device, log=print)
            else:
                out = samplers.ddim_sampling(net, degrader, x_t, min_noise=min_t, max_noise=max_actual_t, cond_img=cond,
                                             labels_tensor=labels, ddim_step_size=skip, device=device, log=print)
        else:
            noise = torch.randn((imgs.shape[0], c, h, w), device=device)
            cond_in, base = None, 0.0
            if flavour == "sr":
                base = lr_condition(imgs)
                cond_in = degrader(img=base, steps=torch.tensor([cond_t]), eps=noise)
            if max_actual_t < max_t and flavour != "sr":      # the SR trainer always starts from pure noise (train_SR:441)
                x_t = degrader(img=imgs, steps=torch.tensor([max_actual_t]), eps=noise)
            else:
                x_t = 1 * noise
            out = samplers.cold_diffusion_sampling(net, degrader, x_t, noise, min_noise=min_t, max_noise=max_actual_t,
                                                   cond_img=cond_in, labels_tensor=labels, skip_step_size=skip, device=device,
                                                   log=print)
            out = out + base
        plot_sampled_images(sampled_imgs=out, file_name=f"diffusion_plot_{steps}", dest_path=out_dir)
        net.train()

    done = False
    for epoch in range(starting_epoch, max_epoch):
        if sampler is not None:
            sampler.set_epoch(epoch)
        total_loss, count = 0.0, 0
        for index, data in enumerate(batches):
            count += 1
            labels, cond_img, target = None, None, None
            if flavour == "doodle":
                x0, cond_img = data[0].to(device, non_blocking=True), data[1].to(device, non_blocking=True)
            elif use_conditional:
                x0, labels = data[0].to(device, non_blocking=True), data[1].to(device, non_blocking=True)
            else:
                x0 = (data[0] if isinstance(data, (list, tuple)) else data).to(device, non_blocking=True)
            n, c, h, w = x0.shape
            # order of the random draws: the base/doodle trainers draw eps before the flips, cold/SR after
            draw = (lambda like: None) if philox_seed is not None else torch.randn_like
            if flavour in ("base", "doodle"):
                eps = draw(x0)
                if flip_imgs and flavour == "base" and not flips_here:
                    x0 = _flip_per_image(x0)
            else:
                if flip_imgs and not flips_here:
                    x0 = _flip_per_image(x0)
                eps = draw(x0)
            first_elem = rank * batch_size * c * h * w        # this rank's slice of the global Philox stream
            if philox_seed is not None and not use_graph:
                from degraders import PhiloxNoise
                eps = PhiloxNoise(philox_seed, global_steps, None, first_elem)
            if flavour == "sr":
                lr = lr_condition(x0)
                target = x0 - lr
                if philox_seed is not None and use_graph:
                    cond_img = lr                              # noised at cond_t inside the captured step, with the same eps
                elif philox_seed is not None:
                    cond_img = degrader.forward_philox(lr, torch.tensor([cond_t]), eps)
                else:
                    cond_img = degrader(img=lr, steps=torch.tensor([cond_t]), eps=eps)
            t = torch.randint(low=min_t, high=max_actual_t, size=(n,), device=device)
            net.train()
            if use_graph and graphed is None:
                graphed = GraphedTrainStep(net, degrader, optim, kind=spec["kind"], philox_seed=philox_seed,
                                           philox_first_elem=first_elem, cond_t=cond_t if flavour == "sr" else None)
            if graphed is not None:
                loss = graphed(x0, t, eps, labels, cond_img, target)
            elif scaler is not None:
                from .steps import scaled_step
                loss = scaled_step(net, degrader, optim, scaler, x0, t, eps, labels, cond_img, target, kind=spec["kind"])
            else:
                from .steps import eps_prediction_step, x0_prediction_step
                if spec["kind"] == "eps":
                    loss = eps_prediction_step(net, degrader, optim, x0, t, eps, labels, cond_img)
                else:
                    loss = x0_prediction_step(net, degrader, optim, x0, t, eps, labels, cond_img, target)
            loss_value = float(loss)                          # host sync, as the reference's .item() (train_diffusion.py:366)
            if loss_value != loss_value:
                raise Exception("NaN encountered during training")
            total_loss += loss_value
            if global_steps % lr_steps == 0 and global_steps > 0:
                for group in optim.param_groups:
                    group["lr"] = group["lr"] * 0.5
            if global_steps % checkpoint_steps == 0 and global_steps >= 0:
                checkpoint(global_steps)
                plot(global_steps, (n, c, h, w))
            log.info("Cum. Steps: {:,} | Steps: {:,} / {:,} | Diffusion: {:.5f} | LR: {:.9f}".format(
                global_steps + 1, index + 1, len(loader), total_loss / count, optim.param_groups[0]["lr"]))
            global_steps += 1
            if args["max_steps"] is not None and global_steps >= args["max_steps"]:
                done = True
                break
        checkpoint(global_steps)
        log.info("Epoch: {:,} | Diffusion: {:.5f} | LR: {:.9f}".format(epoch, total_loss / max(count, 1), optim.param_groups[0]["lr"]))
        if done:
            break
    return {"global_steps": global_steps, "loss": total_loss / max(count, 1), "net": net, "optimizer": optim}
