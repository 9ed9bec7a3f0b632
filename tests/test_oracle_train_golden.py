"""The oracle's train-step restatement against runs of the UNMODIFIED reference trainers (tests/golden/train_runs.pt, written
by tests/golden/make_golden_train.py): `train_diffusion.main()`, `train_noise_cold_diffusion.main()` and
`train_SR_diffusion.main()` and `train_doodle_diffusion.main()`, four optimisation steps each.  The fixture holds what flowed through the reference's degrader
and F.mse_loss at every step and samples of the checkpoints it wrote; replaying the recorded (x0, t, eps) through the oracle
must reproduce every loss, every checkpointed weight and the learning-rate schedule."""
import pytest
import torch
import torch.nn.functional as F

from conftest import load_golden, rel_l2
from oracle import diffusion_oracle as orc
from oracle.weights import synth_state_dict


def _sample(t, n):
    flat = t.detach().float().flatten()
    return flat[(torch.arange(n, dtype=torch.int64) * flat.numel()) // n]


@pytest.mark.parametrize("name", ["base", "cold", "sr", "doodle"])
def test_oracle_reproduces_reference_trainer_run(name):
    fx = load_golden("train_runs.pt")[name]
    cfg, kw = fx["config"], fx["kwargs"]
    sd = {k: v.clone() for k, v in synth_state_dict(fx["shapes"], fx["seed"]).items()}
    if cfg["noise_scheduler"] == "LINEAR":
        sched = ("linear", cfg["beta1"], cfg["betaT"], cfg["max_noise_step"])
    else:
        sched = ("cosine", cfg["max_noise_step"])
    recon = kw.get("image_recon", False)
    moments = {}
    lr = cfg["diffusion_lr"]
    for gstep, (rec, ck) in enumerate(zip(fx["steps"], fx["checkpoints"])):
        calls = rec["degrader_calls"]
        x0, t, eps = calls[0]["img"], calls[0]["steps"], calls[0]["eps"]
        assert t.dtype == torch.int64 and tuple(t.shape) == (x0.shape[0],)
        params = {k: v.clone().requires_grad_(True) for k, v in sd.items() if v.is_floating_point()}
        x_t = orc.q_sample(sched, x0, t, eps)
        assert rel_l2(x_t, calls[0]["out"]) < 1e-6
        if name == "sr":
            # train_SR_diffusion.py:321-366: lr image = area down + area up; same eps, fixed cond_t, target x0 - lr
            assert len(calls) == 2
            low = F.interpolate(F.interpolate(x0, size=(cfg["lr_dim"],) * 2, mode="area"), size=(cfg["sr_dim"],) * 2, mode="area")
            assert torch.equal(calls[1]["img"], low) and torch.equal(calls[1]["eps"], eps)
            assert calls[1]["steps"].tolist() == [cfg["cond_t"]]
            inp = torch.cat((x_t, orc.q_sample(sched, low, calls[1]["steps"], eps)), dim=1)
            target = x0 - low
        elif name == "doodle":
            # train_doodle_diffusion.py:296-315: the condition image is concatenated un-noised; target eps; no labels
            assert len(calls) == 1
            cond_img = rec["net_input"]["x"][:, 3:]
            assert float(cond_img.abs().max()) <= 1.0 and cond_img.shape == x0.shape
            inp = torch.cat((x_t, cond_img), dim=1)
            target = eps
        else:
            assert len(calls) == 1
            inp = x_t
            target = eps if name == "base" else x0
        assert torch.equal(rec["target"], target)
        # what the reference's U_Net actually received: (x_t [, condition]), the per-image timesteps, no labels
        assert rel_l2(inp, rec["net_input"]["x"]) < 1e-6 and torch.equal(rec["net_input"]["t"], t)
        assert rec["net_input"]["cond"] is None
        pred = orc.unet_forward(params, inp, t, None, heads=1, image_recon=recon)
        assert rel_l2(pred.detach(), rec["pred"]) < 1e-5
        loss = F.mse_loss(pred, target)
        assert abs(float(loss.detach()) - rec["loss"]) < 1e-5 * max(1.0, abs(rec["loss"]))
        grads = torch.autograd.grad(loss, list(params.values()), allow_unused=True)
        n_updated = 0
        for (k, p), g in zip(params.items(), grads):
            if g is None:
                continue                                    # y_shift / attention norm: never reached, Adam skips them
            m, v = moments.get(k, (torch.zeros_like(p), torch.zeros_like(p)))
            new, m, v = orc.adam_update(p.detach(), g, m, v, gstep + 1, lr)
            moments[k] = (m, v)
            sd[k] = new
            n_updated += 1
        # Adam exactly as the reference configures it (train_diffusion.py:214-218) ...
        assert ck["betas"] == (0.5, 0.999) and ck["eps"] == 1e-8 and ck["weight_decay"] == 0
        assert ck["n_state"] == n_updated and ck["adam_step"] == gstep + 1
        # ... the weights it checkpointed after this step ...
        got = torch.cat([_sample(sd[k], fx["samples"]) for k in sorted(ck["weights"])])
        want = torch.cat([ck["weights"][k] for k in sorted(ck["weights"])])
        start = torch.cat([_sample(v, fx["samples"]) for k, v in sorted(synth_state_dict(fx["shapes"], fx["seed"]).items())])
        err = rel_l2(got - start, want - start)              # compared on the UPDATE, not the weight
        print(f"{name} step {gstep}: update rel_l2 {err:.2e}")
        assert err < 2e-4, f"step {gstep}"
        # ... and the learning-rate halving (:368-371: after the step, when global_steps % lr_steps == 0 and > 0)
        if gstep % cfg["lr_steps"] == 0 and gstep > 0:
            lr *= 0.5
        assert ck["lr"] == pytest.approx(lr, rel=1e-12)
