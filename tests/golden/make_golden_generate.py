"""Generates tests/golden/generate_runs.pt by calling the UNMODIFIED reference generation entry points on CPU
(`generate_images_diffusion`, `generate_images_cold_diffusion`, `generate_sr_images_diffusion` with `save_locally=False`)
on tiny exported model folders: ensembles that split the timestep range over two checkpoints, a label + condition-image
model under DDPM, a two-stage super-resolution ensemble.  The fixture holds the model descriptions (weights are
re-derived from seeds), the command lines, the numpy inputs and the returned tensors; tests/test_generate_host_logic_cpu.py
runs this repo's entry points on the same folders (numerics by the oracle, host logic by b200/generator.py) and must
return the same tensors -- which also pins the RNG consumption of everything between `torch.manual_seed` and the
samplers' own draws (x_T first, then the U_Net constructors).

Run in the build container only:  PYTHONDONTWRITEBYTECODE=1 python tests/golden/make_golden_generate.py
"""
import json
import os
import sys
import tempfile

os.environ.setdefault("PYTHONDONTWRITEBYTECODE", "1")
sys.dont_write_bytecode = True
HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, "/root/reference")
sys.path.insert(1, ROOT)

import numpy as np  # noqa: E402
import torch  # noqa: E402

from models.U_Net import U_Net  # noqa: E402  (reference)
from generate_images_diffusion import generate_images_diffusion  # noqa: E402  (reference)
from generate_images_cold_diffusion import generate_images_cold_diffusion  # noqa: E402  (reference)
from generate_sr_images_diffusion import generate_sr_images_diffusion  # noqa: E402  (reference)
from oracle.weights import synth_state_dict  # noqa: E402

torch.set_num_threads(8)
NET = dict(num_resnet_block=1, time_dim=32, num_layers=2, attn_layers=[1], attn_heads=1, attn_dim_per_head=None,
           min_channel=32, max_channel=64, cond_dim=None, in_channel=3, out_channel=3, image_recon=False)


def model(name, seed, size, lo, hi, sched, **over):
    d = dict(NET, model_name=name, seed=seed, img_C=3, img_H=size, img_W=size, min_noise=lo, max_noise=hi, noise_scheduler=sched,
             beta_1=5e-3, beta_T=9e-3)
    d.update(over)
    return d


rng = np.random.RandomState(3)
CASES = {
    # name: (entry point, models, command line (config / device appended), extra keyword inputs)
    "ddim_ensemble": ("generate_images_diffusion",
                      [model("a.pt", 101, 16, 11, 20, "LINEAR"), model("b.pt", 102, 16, 1, 10, "LINEAR")],
                      ["-n", "2", "-s", "7", "--diff_alg", "ddim", "--ddim_step_size", "3", "-T", "20"], {}),
    "ddpm_labels_cond_img": ("generate_images_diffusion",
                             [model("a.pt", 103, 16, 1, 12, "COSINE", in_channel=6, cond_dim=3)],
                             ["-n", "2", "-s", "11", "--diff_alg", "ddpm", "-T", "12", "-l", "1", "0", "1"],
                             {"cond_img": rng.randint(0, 256, (16, 16, 3)).astype(np.uint8)}),
    "ddpm_ensemble": ("generate_images_diffusion",
                      [model("a.pt", 104, 16, 7, 12, "COSINE"), model("b.pt", 105, 16, 1, 6, "COSINE")],
                      ["-n", "3", "-s", "13", "--diff_alg", "ddpm", "-T", "12"], {}),
    "cold_ensemble": ("generate_images_cold_diffusion",
                      [model("a.pt", 106, 16, 11, 20, "COSINE", image_recon=True),
                       model("b.pt", 107, 16, 1, 10, "COSINE", image_recon=True)],
                      ["-n", "2", "-s", "5", "--cold_step_size", "3", "-T", "20"], {}),
    "sr_ensemble": ("generate_sr_images_diffusion",
                    [model("a.pt", 108, 32, 11, 20, "COSINE", in_channel=6, image_recon=True, cond_t=5),
                     model("b.pt", 109, 32, 1, 10, "LINEAR", in_channel=6, image_recon=True, cond_t=5)],
                    ["-s", "3", "--cold_step_size", "3", "-T", "20"],
                    {"lr_img": rng.randint(0, 256, (8, 8, 3)).astype(np.uint8)}),
}
ENTRY = {"generate_images_diffusion": generate_images_diffusion, "generate_images_cold_diffusion": generate_images_cold_diffusion,
         "generate_sr_images_diffusion": generate_sr_images_diffusion}


def unet_kwargs(m):
    return dict(in_channel=m["in_channel"], out_channel=m["out_channel"], num_layers=m["num_layers"],
                num_resnet_blocks=m["num_resnet_block"], attn_layers=m["attn_layers"], num_heads=m["attn_heads"],
                dim_per_head=m["attn_dim_per_head"], time_dim=m["time_dim"], cond_dim=m["cond_dim"], min_channel=m["min_channel"],
                max_channel=m["max_channel"], image_recon=m["image_recon"])


def write_folder(folder, models):
    """An exported model folder (export_models.py layout): config.json + one {"model": state_dict} file per entry."""
    shapes = []
    for m in models:
        sh = {k: tuple(v.shape) for k, v in U_Net(**unet_kwargs(m)).state_dict().items()}
        shapes.append(sh)
        torch.save({"model": synth_state_dict(sh, m["seed"])}, os.path.join(folder, m["model_name"]))
    with open(os.path.join(folder, "config.json"), "w") as f:
        json.dump({"models": models}, f)
    return shapes


def main():
    quiet = lambda *a, **k: None
    scratch = os.path.join(ROOT, "gpurun_out")
    os.makedirs(scratch, exist_ok=True)
    out = {}
    for name, (entry, models, argv, inputs) in CASES.items():
        with tempfile.TemporaryDirectory(dir=scratch) as folder:
            shapes = write_folder(folder, models)
            args = ["-c", os.path.join(folder, "config.json"), "--device", "cpu"] + argv
            result = ENTRY[entry](args, log=quiet, save_locally=False, **inputs)
        out[name] = dict(entry=entry, models=models, shapes=shapes, argv=argv, inputs=inputs, result=result.detach().clone())
        print(name, tuple(result.shape), float(result.abs().mean()))
    torch.save(out, os.path.join(HERE, "generate_runs.pt"))
    print("wrote generate_runs.pt", os.path.getsize(os.path.join(HERE, "generate_runs.pt")), "bytes")


if __name__ == "__main__":
    main()
