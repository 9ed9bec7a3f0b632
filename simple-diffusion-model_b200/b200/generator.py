"""Shared body of the three generation entry points (reference generate_images_diffusion.py:35-270,
generate_images_cold_diffusion.py:24-205, generate_sr_images_diffusion.py:32-252): argument parsing, seeding, x_T draw,
the per-model loop over config["models"] (ensembles split the t-range across checkpoints; cascades chain base -> SR),
checkpoint loading and the sampler call.  Images are independent, so under torchrun each rank generates its shard of
`num_images` with no collective (`shard_range`)."""
import argparse
import json
import os
import pathlib
import uuid
from datetime import datetime

import numpy as np
import torch

from ._lib import B200Error
from .functional import area_resize

SUPPORTED_IMG_FORMATS = ["jpeg", "jpg", "png"]


def _shard(t):
    """Under torchrun every rank draws the same full batch (same seed) and keeps its contiguous shard of the images: the
    sampling loop needs no collective and the union of the shards equals the single-process result (the per-step DDPM noise
    is drawn for the whole job and sliced likewise, diffusion_sampling_algorithms.set_shard)."""
    world, rank = int(os.environ.get("WORLD_SIZE", "1")), int(os.environ.get("RANK", "0"))
    if world <= 1 or t is None:
        return t
    import diffusion_sampling_algorithms as S
    from .parallel import shard_range
    lo, hi = shard_range(t.shape[0], rank, world)
    S.set_shard(lo, hi, t.shape[0])          # the samplers' per-step draws (DDPM z) follow the same sharding
    return t[lo:hi].contiguous()


def _image_kind(path):
    with open(path, "rb") as f:
        head = f.read(16)
    if head.startswith(b"\x89PNG\r\n\x1a\n"):
        return "png"
    if head[:3] == b"\xff\xd8\xff":
        return "jpeg"
    return None


def _is_image(img):
    return isinstance(img, np.ndarray) or (torch.is_tensor(img) and img.dtype == torch.uint8)


def _to_tensor(img, device):
    """uint8-range HWC BGR image -> [-1, 1] CHW tensor (the cascade hand-off format, generate_sr:117-126).  A numpy array (cv2)
    takes the reference's host arithmetic; a CUDA uint8 tensor -- produced by `b200.image_io.image_to_u8` from the previous
    stage's samples -- is normalised by one kernel on the device (device-side cascade: no host round trip, same values).
    A leading batch dimension ([N, H, W, C]) is kept."""
    if torch.is_tensor(img):
        from .image_io import u8_to_image
        t = img.to(device)
        out = u8_to_image(t if t.dim() == 4 else t.unsqueeze(0))
        return out if t.dim() == 4 else out[0]
    t = torch.from_numpy((img.astype(float) - 127.5) / 127.5).float()
    return (t.permute(0, 3, 1, 2) if t.dim() == 4 else t.permute(2, 0, 1)).to(device)


def _common_args(description, step_flag, step_help):
    from diffusion_enums import DiffusionAlg
    p = argparse.ArgumentParser(description=description)
    p.add_argument("--device", choices=["cpu", "cuda"], type=str, default="cuda")
    p.add_argument("-c", "--config", help="File path to config file.", required=True, type=pathlib.Path)
    p.add_argument("-s", "--seed", help="Seed value for generating image(default: None).", type=int, default=None)
    p.add_argument("-d", "--dest_path", help="File path to save images generated (Default: ./plots).", type=pathlib.Path)
    p.add_argument(step_flag, help=step_help, default=10, type=int)
    p.add_argument("-T", "--max_T", help="Max T value for noise scheduling (In cases of Ensemble methods).", default=1_000, type=int)
    p.add_argument("-l", "--labels", nargs="*", help="Conditional Labels.", type=float, default=None)
    return p, DiffusionAlg


def _setup(args):
    if args["device"] != "cuda":
        raise B200Error("this build has no CPU path: run the reference implementation for --device cpu")
    if args["seed"] is not None:
        torch.manual_seed(args["seed"])
    if args["dest_path"] is None:
        out_dir = "./"
    else:
        if not args["dest_path"].exists():
            raise ValueError("Invalid destination path, kindly correct and ensure it exists!")
        out_dir = str(args["dest_path"])
    with open(args["config"], "r") as f:
        details = json.load(f)
    if "models" not in details or len(details["models"]) == 0:
        raise ValueError("Invalid/no model details in json, kindly correct and try again!")
    return out_dir, details["models"], os.path.split(args["config"])[0]


def _device():
    """One process per GPU: LOCAL_RANK selects the device under torchrun."""
    if "LOCAL_RANK" in os.environ:
        torch.cuda.set_device(int(os.environ["LOCAL_RANK"]))
    return torch.device("cuda", torch.cuda.current_device())


def _degrader(model_dict, args, device):
    from degraders import CosineNoiseDegradation, NoiseDegradation
    name = model_dict["noise_scheduler"].upper()
    if name == "LINEAR":
        return NoiseDegradation(model_dict["beta_1"], model_dict["beta_T"], args["max_T"], device)
    if name == "COSINE":
        return CosineNoiseDegradation(args["max_T"])
    raise ValueError("Invalid noise scheduler type.")


def _load_net(model_dict, folder, device):
    from models.U_Net import U_Net
    from utils.utils import load_checkpoint
    net = U_Net(in_channel=model_dict["in_channel"], out_channel=model_dict["out_channel"], num_layers=model_dict["num_layers"],
                num_resnet_blocks=model_dict["num_resnet_block"], attn_layers=model_dict["attn_layers"],
                num_heads=model_dict["attn_heads"], dim_per_head=model_dict["attn_dim_per_head"], time_dim=model_dict["time_dim"],
                cond_dim=model_dict["cond_dim"], min_channel=model_dict["min_channel"], max_channel=model_dict["max_channel"],
                image_recon=model_dict["image_recon"]).to(device)
    path = os.path.join(folder, model_dict["model_name"])
    if not os.path.isfile(path):
        raise FileNotFoundError("Invalid path for model in json file, kindly correct and try again!")
    ok, ckpt = load_checkpoint(path)
    if not ok:
        raise Exception("Failed to load model!")
    net.load_state_dict(ckpt["model"])
    return net.eval()


def _labels(model_dict, args, device):
    if model_dict["cond_dim"] is None:
        return None
    if args["labels"] is None or len(args["labels"]) != model_dict["cond_dim"]:
        raise ValueError("Invalid / No conditional labels passed!")
    return torch.tensor(args["labels"]).float().to(device)


def _finish(x, img_h, img_w, out_dir, save_locally, log):
    if not save_locally:
        return x
    from utils.utils import plot_sampled_images
    name = datetime.now().strftime("%d-%m-%Y %H:%M:%S") + "_" + f"({img_h},{img_w})" + "_" + uuid.uuid4().hex
    if int(os.environ.get("WORLD_SIZE", "1")) > 1:
        name += f"_rank{os.environ.get('RANK', '0')}"
    plot_sampled_images(sampled_imgs=x, file_name=name, dest_path=out_dir, log=log)
    return None


def generate_images_diffusion(raw_args=None, log=print, cond_img=None, save_locally=True):
    import diffusion_sampling_algorithms as S
    p, DiffusionAlg = _common_args("Generate Images using Diffusion models.", "--ddim_step_size",
                                   "Number of steps to skip when using ddim.")
    p.add_argument("-n", "--num_images", help="Number of images to generate(default=1).", default=1, type=int)
    p.add_argument("--diff_alg", default="ddpm", choices=[a.name.lower() for a in DiffusionAlg])
    p.add_argument("--cond_img_path", help="File path to conditional image e.g Doodle image.", type=pathlib.Path, default=None)
    args = vars(p.parse_args(raw_args))
    if args["num_images"] <= 0:
        raise ValueError("Invalid image numbers, should be greater than 0!")
    if args["diff_alg"] == "ddim" and (args["ddim_step_size"] < 0 or args["ddim_step_size"] > args["max_T"]):
        raise ValueError("Invalid step size for DDIM!")
    out_dir, models, folder = _setup(args)
    device = _device()
    if args["cond_img_path"] is not None:
        if not os.path.isfile(args["cond_img_path"]):
            raise FileNotFoundError("Invalid path for conditional image, kindly correct and try again!")
        if _image_kind(args["cond_img_path"]) not in SUPPORTED_IMG_FORMATS:
            raise ValueError("Image format is not supported!")
        import cv2
        cond_img = cv2.imread(str(args["cond_img_path"]))
    if cond_img is not None:
        if not _is_image(cond_img):
            raise ValueError("Unsupported conditional image.")
        cond_img = _to_tensor(cond_img, device)
        cond_img = cond_img if cond_img.dim() == 4 else cond_img.unsqueeze(0).repeat(args["num_images"], 1, 1, 1)
    x_t, img_h, img_w = None, None, None
    for model_dict in models:
        if x_t is None:                                  # X_T ~ N(0, I) once; later ensemble members continue from x_t
            img_h, img_w = model_dict["img_H"], model_dict["img_W"]
            x_t = _shard(1 * torch.randn((args["num_images"], model_dict["img_C"], img_h, img_w), device=device))
            cond_img = _shard(cond_img)
        labels = _labels(model_dict, args, device)
        degrader = _degrader(model_dict, args, device)
        net = _load_net(model_dict, folder, device)
        if args["diff_alg"] == "ddpm":
            x_t = S.ddpm_sampling(diffusion_net=net, noise_degradation=degrader, x_t=x_t, min_noise=model_dict["min_noise"],
                                  max_noise=model_dict["max_noise"], cond_img=cond_img, labels_tensor=labels, device=device, log=log)
        elif args["diff_alg"] == "ddim":
            x_t = S.ddim_sampling(diffusion_net=net, noise_degradation=degrader, x_t=x_t, min_noise=model_dict["min_noise"],
                                  max_noise=model_dict["max_noise"], cond_img=cond_img, labels_tensor=labels,
                                  ddim_step_size=args["ddim_step_size"], device=device, log=log)
        else:
            raise ValueError("Invalid Diffusion Algorithm type.")
    return _finish(x_t, img_h, img_w, out_dir, save_locally, log)


def generate_images_cold_diffusion(raw_args=None, log=print, save_locally=True):
    import diffusion_sampling_algorithms as S
    p, _ = _common_args("Generate Images using Cold Diffusion models.", "--cold_step_size",
                        "Number of steps to skip when using cold diffusion.")
    p.add_argument("-n", "--num_images", help="Number of images to generate(default=1).", default=1, type=int)
    args = vars(p.parse_args(raw_args))
    if args["num_images"] <= 0:
        raise ValueError("Invalid image numbers, should be greater than 0!")
    if args["cold_step_size"] < 0 or args["cold_step_size"] > args["max_T"]:
        raise ValueError("Invalid step size for Cold Diffusion!")
    out_dir, models, folder = _setup(args)
    device = _device()
    noise, x0_approx, img_h, img_w = None, None, None, None
    for model_dict in models:
        degrader = _degrader(model_dict, args, device)
        if noise is None:
            img_h, img_w = model_dict["img_H"], model_dict["img_W"]
            noise = _shard(torch.randn((args["num_images"], model_dict["img_C"], img_h, img_w), device=device))
            x_t = 1 * noise
        else:                                            # ensemble hand-off: re-noise the estimate with the SAME noise
            x_t = degrader(img=x0_approx, steps=torch.tensor([model_dict["max_noise"]]), eps=noise)
        labels = _labels(model_dict, args, device)
        net = _load_net(model_dict, folder, device)
        x0_approx = S.cold_diffusion_sampling(diffusion_net=net, noise_degradation=degrader, x_t=x_t, noise=noise,
                                              min_noise=model_dict["min_noise"], max_noise=model_dict["max_noise"], cond_img=None,
                                              labels_tensor=labels, skip_step_size=args["cold_step_size"], device=device, log=log)
    return _finish(x0_approx, img_h, img_w, out_dir, save_locally, log)


def generate_sr_images_diffusion(raw_args=None, lr_img=None, log=print, save_locally=True):
    import diffusion_sampling_algorithms as S
    p, _ = _common_args("Generate Super-Resolution Images using Diffusion models.", "--cold_step_size",
                        "Number of steps to skip when using cold diffusion.")
    p.add_argument("--lr_img_path", help="File path to low resolution image.", type=pathlib.Path, default=None)
    args = vars(p.parse_args(raw_args))
    if args["cold_step_size"] < 0 or args["cold_step_size"] > args["max_T"]:
        raise ValueError("Invalid step size for Cold Diffusion!")
    out_dir, models, folder = _setup(args)
    device = _device()
    if lr_img is not None:
        if not _is_image(lr_img):
            raise ValueError("Invalid low resolution image passed!")
    else:
        path = args["lr_img_path"]
        if path is None or not os.path.isfile(path) or _image_kind(path) not in SUPPORTED_IMG_FORMATS:
            raise ValueError("Invalid/No path for low resolution image or unsupported image.")
        import cv2
        lr_img = cv2.imread(str(path))
    lr = _to_tensor(lr_img, device)
    lr = lr if lr.dim() == 4 else lr.unsqueeze(0)
    noise, delta, upsampled, cond_in, img_h, img_w = None, None, None, None, None, None
    for model_dict in models:
        degrader = _degrader(model_dict, args, device)
        if noise is None:
            img_h, img_w = model_dict["img_H"], model_dict["img_W"]
            if img_h < lr.shape[2] or img_w < lr.shape[3]:
                raise ValueError("Invalid shapes for High Resolution and Low Resolution images.")
            noise = torch.randn((lr.shape[0], model_dict["img_C"], img_h, img_w), device=device)
            x_t = 1 * noise
            upsampled = area_resize(lr, (img_h, img_w))
            cond_in = degrader(img=upsampled, steps=torch.tensor([model_dict["cond_t"]]), eps=noise)
        else:
            x_t = degrader(img=delta, steps=torch.tensor([model_dict["max_noise"]]), eps=noise)
        labels = _labels(model_dict, args, device)
        net = _load_net(model_dict, folder, device)
        delta = S.cold_diffusion_sampling(diffusion_net=net, noise_degradation=degrader, x_t=x_t, noise=noise,
                                          min_noise=model_dict["min_noise"], max_noise=model_dict["max_noise"], cond_img=cond_in,
                                          labels_tensor=labels, skip_step_size=args["cold_step_size"], device=device, log=log)
    return _finish(upsampled + delta, img_h, img_w, out_dir, save_locally, log)
