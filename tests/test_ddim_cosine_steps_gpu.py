"""Per-step DDIM parity under the COSINE schedule with a real U-Net, all 51 evaluations of BASELINE configs[1]
(ddim_step_size = 20, T = 1000; reference diffusion_sampling_algorithms.py:66-148, degraders.py:70-94).

End-to-end comparison of the final image is ill-conditioned here (the first update divides by
sqrt(abar_1000) = 4.4e-8, so a 1e-3 network error is amplified 2e7 x and then fed back through 50 more evaluations), so the
run is checked STEP BY STEP, two ways:
  * free-running through the public `ddim_sampling` (51 evaluations): timestep list bit-exact, the fused update kernel's next
    x_t against the oracle's update of the same (x_t, eps_hat) at every step, return path (last x0 estimate, reference :146-148);
  * on-distribution: at every t of the schedule the network on x_t = q(x_t | x_0) against the CPU oracle on the same input."""
import pytest
import torch

from conftest import load_golden, rel_l2
from oracle import diffusion_oracle as orc
from oracle.weights import synth_state_dict

pytestmark = pytest.mark.gpu


class _Probe:
    """Records what the sampler feeds the network and what it gets back."""

    def __init__(self, net):
        self.net, self.calls = net, []

    def eval(self):
        self.net.eval()
        return self

    def __call__(self, x, t, labels=None):
        out = self.net(x, t, labels)
        self.calls.append((x.detach().cpu().clone(), t.detach().cpu().clone(), out.detach().cpu().clone()))
        return out


@pytest.mark.parametrize("name,batch,size", [("gpu_small", 2, 32), ("default64", 1, 64)])
def test_ddim_cosine_free_running_trajectory(name, batch, size):
    """The public sampler end to end, from x_T ~ N(0, I), random-init weights.  The first update multiplies by 1/sqrt(abar_T) =
    2.3e7, so from the second evaluation on the network sees inputs of magnitude 1e6 .. 1e8 (printed below): its attention
    logits are ~1e12 and the softmax is a hard argmax, where a one-ulp difference between TF32 and fp32 picks another winner.
    Network parity is therefore bimodal along this trajectory (7e-4 on most steps, jumps to 1e-2 .. 1e-1 where an argmax
    flips -- measured, printed, not asserted: a trained network never leaves O(1) inputs; see the on-distribution test below).
    What IS exact here and asserted at all 51 steps: the schedule, the fused update arithmetic, the return path."""
    import diffusion_sampling_algorithms as S
    from degraders import CosineNoiseDegradation
    from models.U_Net import U_Net
    fx = load_golden(f"unet_{name}.pt")
    sd = synth_state_dict(fx["shapes"], fx["seed"])
    net = U_Net(**fx["kwargs"])
    net.load_state_dict(sd)
    net = net.cuda().eval().set_precision("tf32")
    probe = _Probe(net)
    sched = ("cosine", 1000)
    x_T = torch.randn((batch, 3, size, size), generator=torch.Generator().manual_seed(77))
    out = S.ddim_sampling(probe, CosineNoiseDegradation(1000), x_T.cuda(), min_noise=1, max_noise=1000, ddim_step_size=20,
                          device="cuda", log=lambda *a, **k: None).cpu()
    steps = orc.skip_schedule(1, 1000, 20)
    assert len(steps) == 51 and [int(c[1][0]) for c in probe.calls] == steps           # bit-exact schedule, t of shape [1]
    assert all(tuple(c[1].shape) == (1,) and c[1].dtype == torch.int64 for c in probe.calls)
    assert torch.equal(probe.calls[0][0], x_T)
    net_errs, worst_upd = [], (-1.0, 0)
    with torch.no_grad():
        for i, (x_in, t, eps_gpu) in enumerate(probe.calls):
            assert torch.isfinite(eps_gpu).all(), steps[i]
            e_net = rel_l2(eps_gpu, orc.unet_forward(sd, x_in, t, None))
            if i + 1 < len(steps):
                want_next = orc.ddim_update(sched, x_in, eps_gpu, steps[i], steps[i + 1], torch.zeros_like(x_in))
                got_next = probe.calls[i + 1][0]
            else:
                want_next = orc.ddim_x0(sched, x_in, eps_gpu, steps[i])            # schedule ends at t = 1: returns the x0 estimate
                got_next = out
            e_upd = rel_l2(got_next, want_next)
            print(f"t={steps[i]:4d} |x_t|={float(x_in.abs().max()):.3e} net rel-L2 {e_net:.2e} update rel-L2 {e_upd:.2e}")
            net_errs.append(e_net)
            worst_upd = max(worst_upd, (e_upd, steps[i]))
    within = sum(e < 1e-3 for e in net_errs)
    print(f"network rel-L2 < 1e-3 on {within}/51 steps; worst update step {worst_upd}")
    assert net_errs[0] < 1e-3                                  # the one evaluation whose input is O(1): x_T itself
    assert worst_upd[0] < 1e-6, worst_upd                      # fp32 elementwise arithmetic on identical inputs: rounding only


@pytest.mark.parametrize("name,batch,size,precision,tol", [("gpu_small", 2, 32, "tf32", 1e-3), ("gpu_small", 2, 32, "bf16", 1e-2),
                                                           ("default64", 1, 64, "tf32", 1e-3)])
def test_ddim_cosine_every_step_on_distribution(name, batch, size, precision, tol):
    """All 51 timesteps of the DDIM-50 cosine schedule with the real network on inputs from the forward process,
    x_t = q(x_t | x_0) at each t of the schedule (|x_t| = O(1): where a trained model's trajectory lives).  One public-API call
    per step, `ddim_sampling(min_noise = t_next, max_noise = t)`: the network is evaluated at t and t_next, the fused kernel
    applies the update.  Per step: network output vs the CPU oracle on the same input within the north-star bound (1e-3 in
    fp32-accumulate mode; 1e-2 stated for bf16), update vs the oracle's arithmetic, timesteps bit-exact."""
    import diffusion_sampling_algorithms as S
    from degraders import CosineNoiseDegradation
    from models.U_Net import U_Net
    fx = load_golden(f"unet_{name}.pt")
    sd = synth_state_dict(fx["shapes"], fx["seed"])
    net = U_Net(**fx["kwargs"])
    net.load_state_dict(sd)
    net = net.cuda().eval().set_precision(precision)
    probe = _Probe(net)
    deg = CosineNoiseDegradation(1000)
    sched = ("cosine", 1000)
    g = torch.Generator().manual_seed(31)
    x0 = torch.rand((batch, 3, size, size), generator=g) * 2 - 1
    eps = torch.randn((batch, 3, size, size), generator=g)
    steps = orc.skip_schedule(1, 1000, 20)
    seen, worst_net, worst_upd = [], (-1.0, 0), (-1.0, 0)
    with torch.no_grad():
        for i in range(len(steps) - 1):
            t, t_next = steps[i], steps[i + 1]
            x_t = orc.q_sample(sched, x0, torch.tensor([t]), eps)
            probe.calls.clear()
            out = S.ddim_sampling(probe, deg, x_t.cuda(), min_noise=t_next, max_noise=t, ddim_step_size=20, device="cuda",
                                  log=lambda *a, **k: None).cpu()
            assert [int(c[1][0]) for c in probe.calls] == [t, t_next]
            x_in, tt, eps_gpu = probe.calls[0]
            assert torch.equal(x_in, x_t)
            e_net = rel_l2(eps_gpu, orc.unet_forward(sd, x_in, tt, None))
            x_next = orc.ddim_update(sched, x_in, eps_gpu, t, t_next, torch.zeros_like(x_in))
            if t_next == 1:                 # last pair: the sampler returns the x0 estimate made at t = 1 from x_1
                x1_in, t1, eps1 = probe.calls[1]
                e_upd = rel_l2(x1_in, x_next)
                e_net = max(e_net, rel_l2(eps1, orc.unet_forward(sd, x1_in, t1, None)))
                assert rel_l2(out, orc.ddim_x0(sched, x1_in, eps1, 1)) < 1e-5
                seen.append(1)
            else:
                e_upd = rel_l2(out, x_next)
            seen.append(t)
            worst_net = max(worst_net, (e_net, t))
            worst_upd = max(worst_upd, (e_upd, t))
    assert sorted(seen, reverse=True) == steps
    print(f"{name} {precision}: worst network step {worst_net}, worst update step {worst_upd}")
    assert worst_net[0] < tol, worst_net
    assert worst_upd[0] < 1e-5, worst_upd
