"""FusedAdam state interchange: save -> load -> step must continue torch.optim.Adam's trajectory, for the
flat-layout path (U_Net parameters) and the per-tensor path; plus the reference's GradScaler protocol around the step
(train_diffusion.py:130, 358-364)."""
import copy

import pytest
import torch
import torch.nn.functional as F

from conftest import load_golden, rel_l2
from oracle.weights import synth_state_dict

pytestmark = pytest.mark.gpu


def _net(fx, dev):
    from models.U_Net import U_Net
    net = U_Net(**fx["kwargs"])
    net.load_state_dict(synth_state_dict(fx["shapes"], fx["seed"]))
    net = net.to(dev).train().set_precision("tf32")
    net.engine().grad_layout(dev).flatten_params()
    return net


@pytest.mark.parametrize("capturable", [False, True])
def test_save_load_step_round_trip_matches_torch_adam(capturable):
    from b200.optim import FusedAdam
    fx = load_golden("unet_gpu_small.pt")
    dev = torch.device("cuda")
    net_a = _net(fx, dev)
    opt_a = FusedAdam(net_a.parameters(), lr=3e-4, betas=(0.5, 0.999), capturable=capturable)
    # torch.optim.Adam on plain copies of the same parameters is the reference trajectory
    ref_params = [torch.nn.Parameter(p.detach().clone().contiguous()) for p in net_a.parameters()]
    opt_r = torch.optim.Adam(ref_params, lr=3e-4, betas=(0.5, 0.999))
    g = torch.Generator(device="cuda").manual_seed(1)

    def give_grads(net, scale):
        lay = net.engine().layout
        lay.flat.zero_()
        grads = []
        for p in net.parameters():
            if id(p) in lay.offsets:
                gr = torch.randn(p.shape, device=dev, generator=g) * scale
                lay.view(p).copy_(gr)
                p.grad = lay.view(p)
                grads.append(gr)
            else:
                p.grad = None
                grads.append(None)
        return grads

    def ref_step(grads):
        for rp, gr in zip(ref_params, grads):
            rp.grad = None if gr is None else gr.clone()
        opt_r.step()

    for k in range(2):
        grads = give_grads(net_a, 0.1 + k)
        opt_a.step()
        ref_step(grads)
    saved = copy.deepcopy({"model": {k: v.detach().cpu().clone() for k, v in net_a.state_dict().items()},
                           "optimizer": opt_a.state_dict()})
    saved["optimizer"] = torch.load(_roundtrip(saved["optimizer"]), map_location="cpu", weights_only=False)

    net_b = _net(fx, dev)
    net_b.custom_load_state_dict(saved["model"])
    opt_b = FusedAdam(net_b.parameters(), lr=1.0, betas=(0.9, 0.9), capturable=capturable)
    opt_b.load_state_dict(saved["optimizer"])
    for k in range(2, 4):
        grads = give_grads(net_b, 0.1 + k)
        if capturable:
            opt_b.sync_lr()
        opt_b.step()
        ref_step(grads)
    worst = 0.0
    for p, rp in zip(net_b.parameters(), ref_params):
        worst = max(worst, float((p.detach() - rp.detach()).abs().max()))
    # dropped moments would put the first resumed update at ~16x lr per element: 5e-3, not 1e-6
    assert worst < 2e-6, worst
    st_b, st_r = opt_b.state_dict()["state"], opt_r.state_dict()["state"]
    assert set(st_b) == set(st_r)
    k0 = next(iter(st_b))
    assert float(st_b[k0]["step"]) == 4.0
    assert torch.allclose(st_b[k0]["exp_avg_sq"].cpu().contiguous(), st_r[k0]["exp_avg_sq"].cpu(), rtol=1e-5, atol=1e-10)


def _roundtrip(obj):
    import io
    buf = io.BytesIO()
    torch.save(obj, buf)
    buf.seek(0)
    return buf


def test_grad_scaler_protocol_is_a_numerical_no_op_and_skips_on_inf():
    from b200.optim import FusedAdam
    from b200.steps import eps_prediction_step, scaled_step
    from degraders import NoiseDegradation
    fx = load_golden("unet_gpu_small.pt")
    dev = torch.device("cuda")
    deg = NoiseDegradation(5e-3, 9e-3, 1000, device=dev)
    g = torch.Generator().manual_seed(5)
    x0 = (torch.rand((2, 3, 32, 32), generator=g) * 2 - 1).to(dev)
    eps = torch.randn((2, 3, 32, 32), generator=g).to(dev)
    t = torch.randint(1, 1000, (2,), generator=g).to(dev)
    net_a, net_b = _net(fx, dev), _net(fx, dev)
    opt_a = FusedAdam(net_a.parameters(), lr=2e-4, betas=(0.5, 0.999))
    opt_b = FusedAdam(net_b.parameters(), lr=2e-4, betas=(0.5, 0.999))
    scaler = torch.amp.GradScaler("cuda")                      # the reference's torch.cuda.amp.GradScaler(), scale 65536
    for _ in range(2):
        la = eps_prediction_step(net_a, deg, opt_a, x0, t, eps)
        lb = scaled_step(net_b, deg, opt_b, scaler, x0, t, eps, kind="eps")
        assert abs(float(la) - float(lb)) < 1e-6 * abs(float(la))
    sd0 = synth_state_dict(fx["shapes"], fx["seed"])
    num = den = 0.0
    for (ka, pa), (kb, pb) in zip(net_a.named_parameters(), net_b.named_parameters()):
        num += float((pa.detach() - pb.detach()).double().pow(2).sum())
        den += float((pa.detach().cpu() - sd0[ka]).double().pow(2).sum())
    # power-of-two scaling is exact; what differs between the two runs is the order of the fp32 atomics in the norm / bias
    # reductions, which Adam's normalisation (update ~ lr * m / sqrt(v)) amplifies for near-zero gradients: compare the UPDATE
    err = (num / den) ** 0.5
    print("scaled vs unscaled update rel-L2", err)
    assert err < 2e-2
    assert scaler.get_scale() == 65536.0
    # an overflowing step is skipped and the scale backs off, as torch documents
    before = [p.detach().clone() for p in net_b.parameters()]
    bad = x0.clone()
    bad[0, 0, 0, 0] = float("inf")
    scaled_step(net_b, deg, opt_b, scaler, bad, t, eps, kind="eps")
    assert all(torch.equal(a, p.detach()) for a, p in zip(before, net_b.parameters()))
    assert scaler.get_scale() == 32768.0
