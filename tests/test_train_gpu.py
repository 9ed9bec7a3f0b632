"""GPU parity of the hand-written backward pass: loss and every parameter gradient against the reference's autograd
(fixtures from the unmodified reference), plus a few optimisation steps against the oracle."""
import pytest
import torch
import torch.nn.functional as F

from conftest import load_golden, rel_l2
from oracle.weights import synth_state_dict

pytestmark = pytest.mark.gpu

# (relative L2 of ALL parameter gradients taken together, worst single parameter tensor).
# tf32 = fp32-accumulate parity mode: the north-star bound of 1e-3 applies to the whole-network gradient.
TOL = {"tf32": (1e-3, 2e-2), "bf16": (2e-2, 1.5e-1)}


def _build(fx, precision):
    from models.U_Net import U_Net
    net = U_Net(**fx["kwargs"])
    net.load_state_dict(synth_state_dict(fx["shapes"], fx["seed"]))
    return net.cuda().train().set_precision(precision)


@pytest.mark.parametrize("precision", ["tf32", "bf16"])
@pytest.mark.parametrize("name", ["gpu_small", "gpu_cond", "default64", "default128_cond", "default256_sr"])
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
    errs, got_all, want_all = [], [], []
    biggest = max(g["norm"] for g in fx["grads"].values())
    for pname, g in fx["grads"].items():
        got = named[pname].grad.detach().float().cpu().flatten()
        assert torch.isfinite(got).all(), pname
        if g["full"] is not None:
            want, have = g["full"].flatten(), got
        else:
            ns = fx.get("n_samples", 4096)
            idx = (torch.arange(ns, dtype=torch.int64) * got.numel()) // ns
            want, have = g["sample"], got[idx]
        got_all.append(have)
        want_all.append(want)
        if g["norm"] < 1e-7:
            continue
        # tensors whose gradient is < 1e-3 of the largest one (e.g. the S = 256, d = 1024 attention output projection of the 256x256
        # net: norm 2e-5 against 6e-2) are compared on 5 x the bound: their error is set by the ABSOLUTE rounding noise
        # of the activations that feed them, which the whole-network figure above already bounds
        scale = 1.0 if g["norm"] >= 1e-3 * biggest else 0.2
        errs.append((scale * max(rel_l2(have, want), abs(float(got.norm()) - g["norm"]) / g["norm"]), pname))
    errs.sort(reverse=True)
    total = rel_l2(torch.cat(got_all), torch.cat(want_all))
    print(f"{name} {precision}: gradient rel-L2 over all parameters = {total:.3e}; worst tensors: "
          + "; ".join(f"{n}={e:.2e}" for e, n in errs[:4]))
    assert total < TOL[precision][0]
    bad = [(n, e) for e, n in errs if e >= TOL[precision][1]]
    assert not bad, bad[:20]


def test_fused_adagn_sums_backward_equals_the_two_pass_backward():
    """Opt-in mode (SDM_B200_FUSE_ADAGN_SUMS=1): the data-gradient GEMMs emit the next AdaGN backward's pass-1 sums from their
    epilogues and the reduce pass is skipped -- same gradients as the default two-pass backward (bf16 mode, rounding only)."""
    fx = load_golden("unet_gpu_cond.pt")
    cond = fx["cond"].cuda()
    grads = []
    for fused in (False, True):
        net = _build(fx, "bf16")
        net.engine().fuse_adagn_sums = fused
        out = net(fx["x"].cuda(), fx["t"].cuda(), cond)
        F.mse_loss(out, fx["target"].cuda()).backward()
        grads.append(net.engine().layout.flat.clone())
    err = rel_l2(grads[1], grads[0])
    print("fused vs two-pass gradient rel-L2:", err)
    assert err < 2e-2


@pytest.mark.parametrize("side_stream", [False, True])
def test_deferred_grouped_weight_gradients_equal_per_layer_launches(side_stream):
    """Small-workload mode of the backward pass: the weight gradients of a module are deferred and run as one grouped launch
    (b2_conv2d_wgrad_batch), optionally on the side stream -- same gradients as one launch per conv.  Also: the unfused forms of the
    dual-output forward convs and of the data gradients from the forward weights.  The default mode is not bitwise repeatable
    (GroupNorm statistics and split-K sums are fp32 atomics, and one flipped bf16 rounding of an activation is 4e-3 of that
    element), so every variant is held against the run-to-run noise of two identical runs; the op-level tests are exact."""
    fx = load_golden("unet_gpu_cond.pt")
    cond = fx["cond"].cuda()
    grads, outs = {}, {}
    for variant in ("reference", "again", "grouped", "unfused"):
        net = _build(fx, "bf16")
        eng = net.engine()
        eng.group_wgrad = variant == "grouped"
        eng.overlap_wgrad = side_stream and variant == "grouped"
        if variant == "unfused":
            eng.fuse_fwd_act = False
            eng.dgrad_from_fwd = False
        out = net(fx["x"].cuda(), fx["t"].cuda(), cond)
        F.mse_loss(out, fx["target"].cuda()).backward()
        torch.cuda.synchronize()
        outs[variant] = out.detach().clone()
        grads[variant] = eng.layout.flat.clone()
    floor_out = rel_l2(outs["again"], outs["reference"])
    floor_grad = rel_l2(grads["again"], grads["reference"])
    e_out = rel_l2(outs["unfused"], outs["reference"])
    e_grp, e_unf = rel_l2(grads["grouped"], grads["reference"]), rel_l2(grads["unfused"], grads["reference"])
    print(f"noise floor: forward {floor_out:.2e}, gradients {floor_grad:.2e}; forward unfused {e_out:.2e}; "
          f"gradients grouped {e_grp:.2e}, unfused {e_unf:.2e}")
    tol_out, tol_grad = max(4 * floor_out, 5e-3), max(4 * floor_grad, 5e-3)     # a wrong kernel is off by O(1), not by rounding
    assert e_out <= tol_out
    assert e_grp <= tol_grad and e_unf <= tol_grad
