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
@pytest.mark.parametrize("name", ["gpu_small", "gpu_cond", "default64", "default128_cond", "default256_sr"])
def test_unet_forward_matches_reference(name, precision):
    """default128_cond / default256_sr: the class-default 610.7 M net at BASELINE configs[2] / configs[3] shapes (labels at
    128x128 -> S = 1024 attention, multi-tile softmax fix-up inside the full net; 6 input channels + tanh at 256x256 -> S =
    4096), fixtures from the unmodified reference."""
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


@pytest.mark.parametrize("name", ["tiny", "tiny_cond"])
def test_narrow_unet_forward_backward_tf32(name):
    """32/64-channel nets (1-2 channels per GroupNorm group: statistics come from the separate pass, not the conv epilogue),
    two attention heads, 6 input channels, tanh output -- forward and every gradient against the reference fixtures."""
    import torch.nn.functional as F
    fx = load_golden(f"unet_{name}.pt")
    net = _build(fx, "tf32").train()
    cond = fx["cond"].cuda() if fx["cond"] is not None else None
    out = net(fx["x"].cuda(), fx["t"].cuda(), cond)
    assert rel_l2(out.detach().cpu(), fx["out"]) < TOL["tf32"]
    loss = F.mse_loss(out, fx["target"].cuda())
    loss.backward()
    named = dict(net.named_parameters())
    got, want = [], []
    for pname, g in fx["grads"].items():
        have = named[pname].grad.detach().float().cpu().flatten()
        if g["full"] is not None:
            got.append(have); want.append(g["full"].flatten())
        else:
            idx = (torch.arange(4096, dtype=torch.int64) * have.numel()) // 4096
            got.append(have[idx]); want.append(g["sample"])
    assert rel_l2(torch.cat(got), torch.cat(want)) < 2e-3


def test_reference_unit_test_shape():
    """The reference's only test (tests/test_u_net_model.py:21-23): default U_Net, 1x3x128x128, t = [1000] -> same shape."""
    from models.U_Net import U_Net
    net = U_Net().cuda().eval()
    with torch.no_grad():
        y = net(torch.randn((1, 3, 128, 128), device="cuda"), torch.tensor([1000], device="cuda"))
    assert tuple(y.shape) == (1, 3, 128, 128) and torch.isfinite(y).all()


@pytest.mark.parametrize("shape", [(3, 32, 64), (1, 96, 32)])
def test_non_square_and_odd_batch_forward(shape):
    """Rectangular images and batch sizes that do not fill an M tile: engine vs the oracle on the same weights (tf32 mode)."""
    from oracle import diffusion_oracle as orc
    fx = load_golden("unet_gpu_small.pt")
    net = _build(fx, "tf32")
    n, h, w = shape
    g = torch.Generator().manual_seed(9)
    x = torch.rand((n, fx["x"].shape[1], h, w), generator=g) * 2 - 1
    t = torch.randint(1, 1000, (n,), generator=g)
    sd = synth_state_dict(fx["shapes"], fx["seed"])
    with torch.no_grad():
        want = orc.unet_forward(sd, x, t, None)
        got = net(x.cuda(), t.cuda()).cpu()
    assert rel_l2(got, want) < TOL["tf32"]
