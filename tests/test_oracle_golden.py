"""Pins oracle/ (the CPU restatement) to fixtures produced by the unmodified reference (tests/golden/make_golden.py)."""
import pytest
import torch
import torch.nn.functional as F

from conftest import load_golden, rel_l2
from oracle import diffusion_oracle as orc
from oracle.weights import synth_state_dict


@pytest.mark.parametrize("name", ["tiny", "tiny_cond", "gpu_small", "gpu_cond"])
def test_unet_forward_backward_matches_reference(name):
    fx = load_golden(f"unet_{name}.pt")
    kw = fx["kwargs"]
    sd = synth_state_dict(fx["shapes"], fx["seed"])
    for v in sd.values():
        v.requires_grad_(True)
    heads = kw.get("num_heads", 1)
    out = orc.unet_forward(sd, fx["x"], fx["t"], fx["cond"], heads=heads, image_recon=kw.get("image_recon", False))
    assert rel_l2(out, fx["out"]) < 2e-5
    loss = F.mse_loss(out, fx["target"])
    assert abs(float(loss) - fx["loss"]) < 1e-5 * max(1.0, abs(fx["loss"]))
    loss.backward()
    no_grad = sorted(k for k, v in sd.items() if v.grad is None or float(v.grad.abs().sum()) == 0.0)
    assert no_grad == fx["no_grad"]            # y_shift.* and attention norm.* never train
    for pname, g in fx["grads"].items():
        got = sd[pname].grad.flatten()
        assert abs(float(got.norm()) - g["norm"]) <= 2e-4 * g["norm"] + 1e-9, pname
        assert rel_l2(got[:64], g["head"]) < 5e-4 or float(g["head"].norm()) < 1e-9, pname
    with torch.no_grad():
        sdd = {k: v.detach() for k, v in sd.items()}
        out1 = orc.unet_forward(sdd, fx["x"], fx["t"][:1], fx["cond"][0] if fx["cond"] is not None else None,
                                heads=heads, image_recon=kw.get("image_recon", False))
    assert rel_l2(out1, fx["out_t1"]) < 2e-5


def test_unet_default64_forward_matches_reference():
    fx = load_golden("unet_default64.pt")
    sd = synth_state_dict(fx["shapes"], fx["seed"])
    with torch.no_grad():
        out = orc.unet_forward(sd, fx["x"], fx["t"], None)
    assert rel_l2(out, fx["out"]) < 2e-5
    assert len(fx["no_grad"]) == 260


def test_schedules_bit_exact():
    fx = load_golden("schedules.pt")
    beta, alpha, abar = orc.linear_tables(*fx["lin_args"])
    assert torch.equal(beta, fx["lin_beta"]) and torch.equal(alpha, fx["lin_alpha"]) and torch.equal(abar, fx["lin_abar"])
    steps = torch.arange(1, 1001)
    cb, ca, cab = orc.timestep_params(("cosine", 1000), steps)
    assert torch.equal(cb, fx["cos_beta"]) and torch.equal(ca, fx["cos_alpha"]) and torch.equal(cab, fx["cos_abar"])
    lin = ("linear",) + tuple(fx["lin_args"])
    assert torch.equal(orc.q_sample(lin, fx["img"], fx["t"], fx["eps"]), fx["lin_q"])
    assert torch.equal(orc.q_sample(("cosine", 1000), fx["img"], fx["t"], fx["eps"]), fx["cos_q"])
    assert torch.equal(orc.q_sample(("cosine", 1000), fx["img"], fx["t"][1:2], fx["eps"]), fx["cos_q_t1"])


def _stub(calls):
    def net(x, t, labels=None):
        calls.append(int(t[0]))
        out = 0.3 * torch.tanh(x[:, :3]) + 0.001 * t.float()[:, None, None, None] / 10.0
        if x.shape[1] > 3:
            out = out + 0.05 * x[:, 3:6]
        return out
    return net


@pytest.mark.parametrize("sched_name", ["linear", "cosine"])
def test_samplers_match_reference(sched_name):
    fx = load_golden("samplers.pt")
    sched = ("linear", 5e-3, 9e-3, 1000) if sched_name == "linear" else ("cosine", 1000)
    x_T = fx["x_T"]
    for (mn, mx, step) in ((1, 1000, 20), (1, 1000, 100), (250, 1000, 37), (1, 60, 7)):
        ref = fx[f"ddim_{sched_name}_{mn}_{mx}_{step}"]
        calls = []
        torch.manual_seed(2024)
        out = orc.ddim_sample(_stub(calls), sched, x_T.clone(), mn, mx, step)
        assert calls == ref["calls"] == orc.skip_schedule(mn, mx, step)      # bit-exact skip schedule (Q13)
        assert torch.allclose(out, ref["out"], rtol=1e-5, atol=1e-6)
        assert torch.rand(1).item() == ref["rng_after"]                       # same number of draws consumed (Q11)
        refc = fx[f"cold_{sched_name}_{mn}_{mx}_{step}"]
        calls = []
        outc = orc.cold_sample(_stub(calls), sched, x_T.clone(), x_T.clone(), mn, mx, step)
        assert calls == refc["calls"]
        assert torch.allclose(outc, refc["out"], rtol=1e-5, atol=1e-6)
    ref = fx[f"ddpm_{sched_name}_1_40"]
    calls = []
    torch.manual_seed(2025)
    out = orc.ddpm_sample(_stub(calls), sched, x_T.clone(), 1, 40)
    assert calls == ref["calls"]
    assert torch.allclose(out, ref["out"], rtol=1e-5, atol=1e-6)
    assert torch.rand(1).item() == ref["rng_after"]
    ref = fx[f"ddim_cond_{sched_name}"]
    calls = []
    torch.manual_seed(2026)
    out = orc.ddim_sample(_stub(calls), sched, x_T.clone(), 1, 1000, 50, cond_img=fx["cond_img"])
    assert calls == ref["calls"] and len(calls) == 21
    assert torch.allclose(out, ref["out"], rtol=1e-5, atol=1e-6)


def test_ddim_50_is_51_evals():
    assert len(orc.skip_schedule(1, 1000, 20)) == 51
    assert orc.skip_schedule(1, 1000, 20)[-2:] == [20, 1]
