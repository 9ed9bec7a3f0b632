"""GPU parity of the hand-written backward pass: loss and every parameter gradient against the reference's autograd
(fixtures from the unmodified reference), plus a few optimisation steps against the oracle."""
import pytest
import torch
import torch.nn.functional as F

from conftest import load_golden, rel_l2
from oracle.weights import synth_state_dict

pytestmark = pytest.mark.gpu

# gradient tolerances (relative L2 per parameter tensor / on the norm)
TOL = {"tf32": 2e-3, "bf16": 3e-2}


def _build(fx, precision):
    from models.U_Net import U_Net
    net = U_Net(**fx["kwargs"])
    net.load_state_dict(synth_state_dict(fx["shapes"], fx["seed"]))
    return net.cuda().train().set_precision(precision)


@pytest.mark.parametrize("precision", ["tf32", "bf16"])
@pytest.mark.parametrize("name", ["gpu_small", "gpu_cond", "default64"])
def test_unet_backward_matches_reference(name, precision):
    fx = load_golden(f"unet_{name}.pt")
    net = _build(fx, precision)
    cond = fx["cond"].cuda() if fx["cond"] is not None else None
    out = net(fx["x"].cuda(), fx["t"].cuda(), cond)
    assert out.requires_grad
    loss = F.mse_loss(out, fx["target"].cuda())
    loss.backward()
    assert abs(float(loss) - fx["loss"]) < 5e-3 * abs(fx["loss"])
    named = dict(net.named_parameters())
    no_grad = sorted(k for k, p in named.items() if p.grad is None)
    assert no_grad == fx["no_grad"]
    worst = (0.0, None)
    for pname, g in fx["grads"].items():
        got = named[pname].grad.detach().float().cpu().flatten()
        assert torch.isfinite(got).all(), pname
        if g["norm"] < 1e-7:
            continue
        err_norm = abs(float(got.norm()) - g["norm"]) / g["norm"]
        err = err_norm
        if g["full"] is not None:
            err = max(err, rel_l2(got, g["full"].flatten()))
        elif float(g["head"].norm()) > 1e-3 * g["norm"] / max(1.0, (got.numel() / 64) ** 0.5):
            err = max(err, min(rel_l2(got[:64], g["head"]), 10.0) * 0.5)      # 64-element slices are noisier than whole tensors
        if err > worst[0]:
            worst = (err, pname)
        assert err < TOL[precision], (pname, err)
    print(f"{name} {precision}: worst gradient error {worst[0]:.3e} at {worst[1]}")
