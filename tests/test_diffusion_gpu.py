"""GPU parity of q-sample and the three samplers (stub network) against fixtures from the unmodified reference."""
import pytest
import torch

from conftest import load_golden, rel_l2

pytestmark = pytest.mark.gpu


class StubNet(torch.nn.Module):
    def __init__(self):
        super().__init__()
        self.calls = []

    def forward(self, x, t, labels=None):
        self.calls.append(int(t[0]))
        out = 0.3 * torch.tanh(x[:, :3]) + 0.001 * t.float()[:, None, None, None] / 10.0
        if x.shape[1] > 3:
            out = out + 0.05 * x[:, 3:6]
        return out


def _degraders():
    from degraders import CosineNoiseDegradation, NoiseDegradation
    return {"linear": NoiseDegradation(5e-3, 9e-3, 1000, device="cuda"), "cosine": CosineNoiseDegradation(1000)}


def test_qsample_matches_reference():
    fx = load_golden("schedules.pt")
    deg = _degraders()
    img, eps, t = fx["img"].cuda(), fx["eps"].cuda(), fx["t"].cuda()
    assert torch.allclose(deg["linear"](img, t, eps).cpu(), fx["lin_q"], rtol=1e-6, atol=1e-6)
    assert torch.allclose(deg["cosine"](img, t, eps).cpu(), fx["cos_q"], rtol=1e-5, atol=2e-6)
    assert torch.allclose(deg["cosine"](img, t[1:2], eps).cpu(), fx["cos_q_t1"], rtol=1e-5, atol=2e-6)
    torch.manual_seed(3)
    a = deg["linear"](img, t)                      # eps drawn like the reference: randn_like on the current generator
    torch.manual_seed(3)
    b = deg["linear"](img, t, torch.randn_like(img))
    assert torch.equal(a, b)


@pytest.mark.parametrize("sched", ["linear", "cosine"])
def test_samplers_match_reference(sched):
    import diffusion_sampling_algorithms as S
    fx = load_golden("samplers.pt")
    deg = _degraders()[sched]
    x_T = fx["x_T"].cuda()
    quiet = lambda *a, **k: None
    for (mn, mx, step) in ((1, 1000, 20), (1, 1000, 100), (250, 1000, 37), (1, 60, 7)):
        ref = fx[f"ddim_{sched}_{mn}_{mx}_{step}"]
        net = StubNet()
        out = S.ddim_sampling(net, deg, x_T.clone(), min_noise=mn, max_noise=mx, ddim_step_size=step, device="cuda", log=quiet)
        assert net.calls == ref["calls"]
        assert rel_l2(out.cpu(), ref["out"]) < 2e-5, (sched, mn, mx, step)
        refc = fx[f"cold_{sched}_{mn}_{mx}_{step}"]
        net = StubNet()
        outc = S.cold_diffusion_sampling(net, deg, x_T.clone(), x_T.clone(), min_noise=mn, max_noise=mx, skip_step_size=step,
                                         device="cuda", log=quiet)
        assert net.calls == refc["calls"]
        assert rel_l2(outc.cpu(), refc["out"]) < 2e-5, (sched, mn, mx, step)
    ref = fx[f"ddim_cond_{sched}"]
    net = StubNet()
    out = S.ddim_sampling(net, deg, x_T.clone(), min_noise=1, max_noise=1000, ddim_step_size=50, cond_img=fx["cond_img"],
                          device="cuda", log=quiet)
    assert net.calls == ref["calls"]
    assert rel_l2(out.cpu(), ref["out"]) < 2e-5


@pytest.mark.parametrize("sched", ["linear", "cosine"])
def test_ddpm_update_matches_oracle_with_injected_noise(sched):
    """The reference draws z on its own device generator; parity is checked per update with z injected."""
    from b200._lib import call, ptr, stream
    from oracle import diffusion_oracle as orc
    deg = _degraders()[sched]
    osched = ("linear", 5e-3, 9e-3, 1000) if sched == "linear" else ("cosine", 1000)
    g = torch.Generator().manual_seed(7)
    x = torch.randn((2, 3, 8, 8), generator=g)
    e = torch.randn((2, 3, 8, 8), generator=g)
    z = torch.randn((2, 3, 8, 8), generator=g)
    xd, ed, zd = x.cuda(), e.cuda(), z.cuda()
    for step in (1000, 500, 2, 1):
        beta, alpha, abar = deg.host_params(step)
        out = torch.empty_like(xd)
        zz = zd if step > 1 else None
        call("b2_ddpm_step", ptr(xd), ptr(ed), ptr(zz), ptr(out), x.numel(), float(1 / alpha ** 0.5),
             float((1 - alpha) / (1 - abar) ** 0.5), float(beta ** 0.5), 0, 0, 0, 0, stream())
        want = orc.ddpm_update(osched, x, e, step, z if step > 1 else 0)
        assert rel_l2(out.cpu(), want) < 1e-5


def test_ddim_consumes_reference_rng_draws():
    import diffusion_sampling_algorithms as S
    deg = _degraders()["linear"]
    x_T = torch.randn((2, 3, 8, 8), device="cuda")
    torch.manual_seed(42)
    S.ddim_sampling(StubNet(), deg, x_T.clone(), ddim_step_size=100, device="cuda", log=lambda *a, **k: None)
    after = torch.rand(1, device="cuda").item()
    torch.manual_seed(42)
    for _ in range(len(S.skip_schedule(1, 1000, 100)) - 1):
        torch.randn_like(x_T)
    assert torch.rand(1, device="cuda").item() == after


def test_philox_normal_is_shard_invariant_and_normal():
    from b200._lib import call, ptr, stream
    n = 1 << 20
    full = torch.empty(n, device="cuda")
    call("b2_philox_normal", ptr(full), n, 1234, 7, 0, stream())
    half = torch.empty(n // 2, device="cuda")
    call("b2_philox_normal", ptr(half), n // 2, 1234, 7, n // 2, stream())
    assert torch.equal(half, full[n // 2:])
    assert abs(float(full.mean())) < 5e-3 and abs(float(full.std()) - 1.0) < 5e-3
    other = torch.empty(n, device="cuda")
    call("b2_philox_normal", ptr(other), n, 1234, 8, 0, stream())
    assert abs(float((full * other).mean())) < 5e-3


def test_mse_loss_and_grad():
    from b200._lib import call, ptr, stream
    p = torch.randn((4, 3, 16, 16), device="cuda")
    t = torch.randn((4, 3, 16, 16), device="cuda")
    g = torch.empty_like(p)
    loss = torch.empty(1, device="cuda")
    call("b2_mse_loss_grad", ptr(p), ptr(t), ptr(g), ptr(loss), p.numel(), 1.0, stream())
    assert abs(float(loss) - float(torch.nn.functional.mse_loss(p, t))) < 1e-5
    assert torch.allclose(g, 2 * (p - t) / p.numel(), rtol=1e-5, atol=1e-8)


@pytest.mark.parametrize("precision,tol", [("tf32", 2e-3), ("bf16", 5e-2)])
def test_end_to_end_sampling_with_the_real_network(precision, tol):
    """Whole sampling loops (DDIM, cold, DDPM with injected draws) through the sm_100a U-Net against the oracle's loop over
    the oracle's network, linear schedule (1/sqrt(abar_T) = 33: well conditioned, unlike the cosine start).  The bf16 bound
    is the stated end-to-end tolerance of the speed path; graph-replayed inference must agree with eager launches."""
    import diffusion_sampling_algorithms as S
    from degraders import NoiseDegradation
    from models.U_Net import U_Net
    from oracle import diffusion_oracle as orc
    from oracle.weights import synth_state_dict
    fx = load_golden("unet_gpu_small.pt")
    sd = synth_state_dict(fx["shapes"], fx["seed"])
    net = U_Net(**fx["kwargs"])
    net.load_state_dict(sd)
    net = net.cuda().eval().set_precision(precision)
    deg = NoiseDegradation(5e-3, 9e-3, 1000, device="cuda")
    osched = ("linear", 5e-3, 9e-3, 1000)
    onet = lambda a, b, c=None: orc.unet_forward(sd, a, b, None)
    quiet = lambda *a, **k: None
    g = torch.Generator().manual_seed(17)
    x_T = torch.randn((2, 3, 32, 32), generator=g)
    with torch.no_grad():
        want = orc.ddim_sample(onet, osched, x_T, 1, 1000, 250)
        got = S.ddim_sampling(net, deg, x_T.cuda(), min_noise=1, max_noise=1000, ddim_step_size=250, device="cuda", log=quiet)
        assert rel_l2(got.cpu(), want) < tol
        net.cuda_graphs(True)
        replayed = S.ddim_sampling(net, deg, x_T.cuda(), min_noise=1, max_noise=1000, ddim_step_size=250, device="cuda", log=quiet)
        net.cuda_graphs(False)
        assert rel_l2(replayed, got) < (1e-5 if precision == "tf32" else 2e-2)
        want_c = orc.cold_sample(onet, osched, x_T, x_T, 1, 1000, 250)
        got_c = S.cold_diffusion_sampling(net, deg, x_T.cuda(), x_T.cuda(), min_noise=1, max_noise=1000, skip_step_size=250,
                                          device="cuda", log=quiet)
        assert rel_l2(got_c.cpu(), want_c) < tol
