"""Generates tests/golden/train_runs.pt by running the UNMODIFIED reference trainers from /root/reference on CPU:
`train_diffusion.main()`, `train_noise_cold_diffusion.main()` and `train_SR_diffusion.main()`, four optimisation steps each
of a tiny U_Net on four synthetic PNG images -- and `train_doodle_diffusion.main()`, whose dataset class imports `tinydb`:
that package is absent from this image, so a minimal read-only stand-in (same `TinyDB(path).table(name).all()` surface over
the same JSON file format) is registered in sys.modules for the run.

Nothing of the reference is modified: the script only wraps, at run time, the functions the trainers call so that it can
RECORD what flowed through them -- the (image, timestep, eps) triples given to the noise degrader, the target and value of
every F.mse_loss -- and afterwards reads the checkpoints the trainers wrote after every step.  The fixture pins the
train-step bodies, the optimiser configuration and the learning-rate schedule (A19) of the real scripts;
tests/test_oracle_train_golden.py replays the recorded inputs through the oracle.

Run in the build container only:  PYTHONDONTWRITEBYTECODE=1 python tests/golden/make_golden_train.py
"""
import importlib
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

import cv2  # noqa: E402
import numpy as np  # noqa: E402
import torch  # noqa: E402
import torch.nn.functional as F  # noqa: E402

import types  # noqa: E402

try:
    import tinydb  # noqa: F401,E402
except ImportError:                                  # stand-in for the one call chain custom_dataset/doodle_dataset.py:19-28 uses
    class _Table(list):
        def all(self):
            return list(self)

    class _TinyDB:
        def __init__(self, path):
            with open(path) as f:
                self._tables = json.load(f)

        def table(self, name):
            return _Table(self._tables.get(name, {}).values())

    _mod = types.ModuleType("tinydb")
    _mod.TinyDB = _TinyDB
    sys.modules["tinydb"] = _mod

import degraders as ref_degraders  # noqa: E402  (reference)
from models.U_Net import U_Net  # noqa: E402  (reference)
from oracle.weights import synth_state_dict  # noqa: E402

torch.set_num_threads(8)
SAMPLES = 256       # evenly spaced elements kept per parameter tensor and checkpoint

NET = dict(num_resnet_blocks=1, time_dim=32, num_layers=2, attn_layers=[1], min_channel=32, max_channel=64)
RUNS = {
    # name: (module, U_Net kwargs, image size, config overrides)
    "base": ("train_diffusion", dict(NET), 16,
             dict(noise_scheduler="LINEAR", beta1=5e-3, betaT=9e-3, diffusion_alg="DDIM")),
    "cold": ("train_noise_cold_diffusion", dict(NET, image_recon=True), 16, dict(noise_scheduler="COSINE")),
    "sr": ("train_SR_diffusion", dict(NET, in_channel=6, image_recon=True), 32,
           dict(noise_scheduler="COSINE", lr_dim=8, sr_dim=32, cond_t=5)),
    "doodle": ("train_doodle_diffusion", dict(NET, in_channel=6), 16,
               dict(noise_scheduler="LINEAR", beta1=5e-3, betaT=9e-3, diffusion_alg="DDPM")),
}


def sample(t):
    flat = t.detach().float().flatten()
    idx = (torch.arange(SAMPLES, dtype=torch.int64) * flat.numel()) // SAMPLES
    return flat[idx].clone()


def run(name, module, net_kw, size, overrides, work, keep_ckpt=None):
    seed = 4321
    shapes = {k: tuple(v.shape) for k, v in U_Net(**net_kw).state_dict().items()}
    init = os.path.join(work, f"{name}_init.pt")
    torch.save({"model": synth_state_dict(shapes, seed)}, init)
    rng = np.random.RandomState(7)
    img_dir = os.path.join(work, f"{name}_imgs")
    os.makedirs(img_dir)
    for i in range(4):
        cv2.imwrite(os.path.join(img_dir, f"{i}.png"), rng.randint(0, 256, (size, size, 3)).astype(np.uint8))
    dataset_path = os.path.join(img_dir, "*.png")
    if name == "doodle":                             # TinyDB file: Data rows {filename, <label>: condition image path}
        for i in range(4):
            cv2.imwrite(os.path.join(img_dir, f"c{i}.png"), rng.randint(0, 256, (size, size, 3)).astype(np.uint8))
        dataset_path = os.path.join(work, "doodle_db.json")
        with open(dataset_path, "w") as f:
            json.dump({"Labels": {"1": {"labels": ["doodle"]}},
                       "Data": {str(i + 1): {"filename": os.path.join(img_dir, f"{i}.png"),
                                             "doodle": os.path.join(img_dir, f"c{i}.png")} for i in range(4)}}, f)
    out_dir = os.path.join(work, f"{name}_out")
    cfg = dict(dataset_path=dataset_path, out_dir=out_dir, checkpoint_steps=1, lr_steps=2, max_epoch=2,
               plot_img_count=1, use_conditional=False, flip_imgs=False, model_checkpoint=init, config_checkpoint=None,
               load_diffusion_optim=False, diffusion_lr=2e-4, batch_size=2, min_noise_step=1, max_noise_step=20,
               max_actual_noise_step=20, skip_step=5, in_channel=net_kw.get("in_channel", 3), out_channel=3,
               num_layers=net_kw["num_layers"], num_resnet_block=net_kw["num_resnet_blocks"], attn_layers=net_kw["attn_layers"],
               attn_heads=1, attn_dim_per_head=None, time_dim=net_kw["time_dim"], cond_dim=None,
               min_channel=net_kw["min_channel"], max_channel=net_kw["max_channel"], img_recon=net_kw.get("image_recon", False))
    cfg.update(overrides)
    cfg_path = os.path.join(work, f"{name}.json")
    with open(cfg_path, "w") as f:
        json.dump(cfg, f)

    # ---- recorders around the functions the trainer calls (the trainer itself is untouched)
    pending, net_inputs, steps = [], [], []
    orig_fwd = {cls: cls.forward for cls in (ref_degraders.NoiseDegradation, ref_degraders.CosineNoiseDegradation)}
    orig_mse, orig_zero, orig_net = F.mse_loss, torch.optim.Adam.zero_grad, U_Net.forward

    def net_forward(self, x, t=None, cond=None):
        if torch.is_grad_enabled() and self.training:
            net_inputs.append(dict(x=x.detach().clone(), t=t.detach().clone(), cond=cond))
        return orig_net(self, x, t, cond)

    def make_fwd(orig):
        def fwd(self, img, steps, eps=None):
            out = orig(self, img, steps, eps)
            if torch.is_grad_enabled() and eps is not None:
                pending.append(dict(img=img.detach().clone(), steps=steps.detach().clone(), eps=eps.detach().clone(),
                                    out=out.detach().clone()))
            return out
        return fwd

    def mse(pred, target, *a, **k):
        loss = orig_mse(pred, target, *a, **k)
        if torch.is_grad_enabled() and pred.requires_grad:
            steps.append(dict(degrader_calls=list(pending), net_input=net_inputs[-1], target=target.detach().clone(),
                              pred=pred.detach().clone(), loss=float(loss.detach())))
        return loss

    def zero_grad(self, *a, **k):
        pending.clear()                              # a train step starts here; plot-time degrader calls are dropped
        return orig_zero(self, *a, **k)

    for cls, orig in orig_fwd.items():
        cls.forward = make_fwd(orig)
    F.mse_loss = mse
    torch.optim.Adam.zero_grad = zero_grad
    U_Net.forward = net_forward
    argv = sys.argv
    try:
        sys.argv = [module, "-c", cfg_path, "--device", "cpu"]
        torch.manual_seed(99)
        importlib.import_module(module).main()
    finally:
        sys.argv = argv
        for cls, orig in orig_fwd.items():
            cls.forward = orig
        F.mse_loss = orig_mse
        torch.optim.Adam.zero_grad = orig_zero
        U_Net.forward = orig_net

    assert len(steps) == 4, len(steps)
    ckpts = []
    for k in range(4):
        ck = torch.load(os.path.join(out_dir, "checkpoint", f"diffusion_{k}.pt"), map_location="cpu", weights_only=False)
        opt = ck["optimizer"]
        ckpts.append(dict(weights={n: sample(v) for n, v in ck["model"].items()},
                          lr=opt["param_groups"][0]["lr"], betas=tuple(opt["param_groups"][0]["betas"]),
                          eps=opt["param_groups"][0]["eps"], weight_decay=opt["param_groups"][0]["weight_decay"],
                          n_state=len(opt["state"]),
                          adam_step=float(next(iter(opt["state"].values()))["step"])))
    if keep_ckpt is not None:                        # the checkpoint FILES the reference wrote, byte for byte
        import shutil
        dest, which = keep_ckpt
        os.makedirs(dest, exist_ok=True)
        for k in which:
            for stem in ("diffusion", "config"):
                shutil.copyfile(os.path.join(out_dir, "checkpoint", f"{stem}_{k}.pt"), os.path.join(dest, f"{stem}_{k}.pt"))
    return dict(module=module, kwargs=net_kw, shapes=shapes, seed=seed, config=cfg, steps=steps, checkpoints=ckpts,
                samples=SAMPLES)


def main():
    out = {}
    scratch = os.path.join(ROOT, "gpurun_out")          # git-ignored scratch space inside the repo
    os.makedirs(scratch, exist_ok=True)
    with tempfile.TemporaryDirectory(dir=scratch) as work:
        for name, (module, net_kw, size, overrides) in RUNS.items():
            out[name] = run(name, module, net_kw, size, overrides, work)
            print(name, [round(s["loss"], 6) for s in out[name]["steps"]], [c["lr"] for c in out[name]["checkpoints"]])
    torch.save(out, os.path.join(HERE, "train_runs.pt"))
    print("wrote", os.path.join(HERE, "train_runs.pt"), os.path.getsize(os.path.join(HERE, "train_runs.pt")), "bytes")


if __name__ == "__main__":
    main()
