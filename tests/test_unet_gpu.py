"""GPU parity of the whole U-Net forward against fixtures produced by the unmodified reference."""
import pytest
import torch

from conftest import load_golden, rel_l2
from oracle.weights import synth_state_dict

pytestmark = pytest.mark.gpu

# fp32-accumulate parity mode (TF32 operands): the north-star bound.  bf16 storage + bf16 operands: stated tolerance.
TOL = {"tf32": 1e-3, "bf16": 1e-2}


def _build(fx, precision):
    from models.U_Net import U_Net
    net = U_Net(**fx["kwargs"])
    net.load_state_dict(synth_state_dict(fx["shapes"], fx["seed"]))
    return net.cuda().eval().set_precision(precision)


@pytest.mark.parametrize("precision", ["tf32", "bf16"])
@pytest.mark.parametrize("name", ["gpu_small", "gpu_cond", "default64"])
def test_unet_forward_matches_reference(name, precision):
    fx = load_golden(f"unet_{name}.pt")
    net = _build(fx, precision)
    cond = fx["cond"].cuda() if fx["cond"] is not None else None
    with torch.no_grad():
        out = net(fx["x"].cuda(), fx["t"].cuda(), cond)
        err = rel_l2(out.cpu(), fx["out"])
        print(f"{name} {precision}: rel_l2 = {err:.3e}")
        assert err < TOL[precision]
        out1 = net(fx["x"].cuda(), fx["t"][:1].cuda(), cond[0] if cond is not None else None)
        assert rel_l2(out1.cpu(), fx["out_t1"]) < TOL[precision]
