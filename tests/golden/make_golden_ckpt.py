"""Generates tests/golden/ref_ckpt/ + tests/golden/ckpt_resume.pt: a checkpoint WRITTEN BY THE UNMODIFIED REFERENCE TRAINER
(`train_diffusion.main()`, CPU, micro U_Net, four steps, checkpoint every step) kept byte for byte -- `diffusion_1.pt`
({"model", "optimizer"}: weights and Adam moments after two optimisation steps) and `config_1.pt` -- together with what the
reference did NEXT: the (x0, t, eps) of steps 2 and 3, their losses and samples of the weights it checkpointed after them.
tests/test_checkpoint_resume_gpu.py loads the files through this repo's `load_checkpoint` / `custom_load_state_dict` /
`FusedAdam.load_state_dict` and must reproduce the reference's next-step loss and weights (the Adam moments
only matter if they were really adopted, which is what ADVICE r1 found broken).

Run in the build container only:  PYTHONDONTWRITEBYTECODE=1 python tests/golden/make_golden_ckpt.py
"""
import os
import sys
import tempfile

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import make_golden_train as mgt  # noqa: E402  (sets up sys.path for /root/reference and the recorders)

import torch  # noqa: E402

NET = dict(num_resnet_blocks=1, time_dim=32, num_layers=2, attn_layers=[1], min_channel=32, max_channel=32)


def main():
    scratch = os.path.join(mgt.ROOT, "gpurun_out")
    os.makedirs(scratch, exist_ok=True)
    dest = os.path.join(HERE, "ref_ckpt")
    with tempfile.TemporaryDirectory(dir=scratch) as work:
        fx = mgt.run("resume", "train_diffusion", dict(NET), 16,
                     dict(noise_scheduler="LINEAR", beta1=5e-3, betaT=9e-3, diffusion_alg="DDIM"), work, keep_ckpt=(dest, [1]))
    # keep only what the resume test needs: the two steps after the kept checkpoint
    out = dict(kwargs=fx["kwargs"], config=fx["config"], samples=fx["samples"], shapes=fx["shapes"], seed=fx["seed"],
               steps=fx["steps"][2:], checkpoints=fx["checkpoints"][2:], loaded_after_steps=2)
    torch.save(out, os.path.join(HERE, "ckpt_resume.pt"))
    for f in sorted(os.listdir(dest)):
        print(f, os.path.getsize(os.path.join(dest, f)), "bytes")
    print("losses", [s["loss"] for s in fx["steps"]], "lr", [c["lr"] for c in fx["checkpoints"]])


if __name__ == "__main__":
    main()
