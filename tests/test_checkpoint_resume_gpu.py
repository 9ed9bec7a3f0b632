"""Checkpoint interchange with the reference: tests/golden/ref_ckpt/{diffusion_1,config_1}.pt are the files
the UNMODIFIED reference trainer wrote after two optimisation steps (tests/golden/make_golden_ckpt.py).  Loading them through
this repo's `load_checkpoint` + `custom_load_state_dict` + `FusedAdam.load_state_dict` and replaying the reference's NEXT two
steps must give the reference's losses and the weights it checkpointed afterwards -- which only happens when the Adam
moments and step count were really adopted by the fused optimiser (they used to be dropped silently)."""
import os

import pytest
import torch

from conftest import GOLDEN, load_golden, rel_l2

pytestmark = pytest.mark.gpu


def _sample(t, n):
    flat = t.detach().float().flatten().cpu()
    return flat[(torch.arange(n, dtype=torch.int64) * flat.numel()) // n]


@pytest.mark.parametrize("mode", ["eager", "graph"])
def test_resume_from_reference_checkpoint(mode):
    from b200.graph import GraphedTrainStep
    from b200.optim import FusedAdam
    from b200.steps import eps_prediction_step
    from degraders import NoiseDegradation
    from models.U_Net import U_Net
    from utils.utils import load_checkpoint

    fx = load_golden("ckpt_resume.pt")
    cfg = fx["config"]
    dev = torch.device("cuda")
    ok, ckpt = load_checkpoint(os.path.join(GOLDEN, "ref_ckpt", "diffusion_1.pt"), log=lambda *a, **k: None)
    assert ok and set(ckpt) == {"model", "optimizer"}
    ok, cc = load_checkpoint(os.path.join(GOLDEN, "ref_ckpt", "config_1.pt"), log=lambda *a, **k: None)
    assert ok and cc["global_steps"] == 1 and cc["beta_1"] == cfg["beta1"] and cc["beta_T"] == cfg["betaT"]

    net = U_Net(**fx["kwargs"])
    assert list(net.state_dict().keys()) == list(ckpt["model"].keys())          # same names, same registration order
    net.custom_load_state_dict(ckpt["model"])
    net = net.to(dev).train().set_precision("tf32")
    net.engine().grad_layout(dev).flatten_params()
    opt = FusedAdam(net.parameters(), lr=123.0, betas=(0.9, 0.9), capturable=(mode == "graph"))   # overwritten by the checkpoint
    opt.load_state_dict(ckpt["optimizer"])
    assert opt.param_groups[0]["lr"] == cfg["diffusion_lr"] and tuple(opt.param_groups[0]["betas"]) == (0.5, 0.999)
    # the loaded moments are the live ones: views of the flat buffers the kernel updates
    lay = net.engine().layout
    m_flat, v_flat = opt._flat[id(lay)]
    ref_state = ckpt["optimizer"]["state"]
    params = list(net.parameters())
    for idx, st in ref_state.items():
        p = params[idx]
        mine = opt.state[p]
        assert float(mine["step"]) == 2.0
        assert mine["exp_avg"].data_ptr() == lay._shaped(m_flat, p).data_ptr()
        assert torch.equal(mine["exp_avg"].cpu(), st["exp_avg"]) and torch.equal(mine["exp_avg_sq"].cpu(), st["exp_avg_sq"])

    deg = NoiseDegradation(cc["beta_1"], cc["beta_T"], cfg["max_noise_step"], device=dev)
    graphed = GraphedTrainStep(net, deg, opt, kind="eps") if mode == "graph" else None
    start = torch.cat([_sample(v, fx["samples"]) for _, v in sorted(ckpt["model"].items())])
    gstep = fx["loaded_after_steps"]
    for rec, ck in zip(fx["steps"], fx["checkpoints"]):
        call = rec["degrader_calls"][0]
        x0, t, eps = call["img"].to(dev), call["steps"].to(dev), call["eps"].to(dev)
        loss = graphed(x0, t, eps) if graphed is not None else eps_prediction_step(net, deg, opt, x0, t, eps)
        loss = float(loss)
        assert abs(loss - rec["loss"]) < 1e-3 * max(1.0, abs(rec["loss"])), (gstep, loss, rec["loss"])
        sd = net.state_dict()
        got = torch.cat([_sample(sd[k], fx["samples"]) for k in sorted(ck["weights"])])
        want = torch.cat([ck["weights"][k] for k in sorted(ck["weights"])])
        err = rel_l2(got - start, want - start)
        print(f"resume/{mode} step {gstep}: loss {loss:.6f} (ref {rec['loss']:.6f}) update rel_l2 {err:.2e}")
        # with zeroed moments the first update would be ~16x too large (lr (1-b1)/sqrt(1-b2) per element): err >> 1
        assert err < 6e-2, (gstep, err)
        if gstep % cfg["lr_steps"] == 0 and gstep > 0:
            for group in opt.param_groups:
                group["lr"] = group["lr"] * 0.5
        assert ck["lr"] == pytest.approx(opt.param_groups[0]["lr"], rel=1e-12)
        assert float(next(iter(opt.state_dict()["state"].values()))["step"]) == ck["adam_step"]
        gstep += 1

    # and back: a checkpoint written here has the reference's format (keys, param-group options torch.optim.Adam needs)
    out = {"model": net.state_dict(), "optimizer": opt.state_dict()}
    assert list(out["model"].keys()) == list(ckpt["model"].keys())
    assert set(out["optimizer"]["state"].keys()) == set(ref_state.keys())
    theirs = torch.optim.Adam([torch.nn.Parameter(torch.zeros_like(p, device="cpu").contiguous()) for p in net.parameters()])
    theirs.load_state_dict(out["optimizer"])
    assert theirs.param_groups[0]["lr"] == opt.param_groups[0]["lr"]
