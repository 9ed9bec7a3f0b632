"""Host-side logic of the schedules and samplers (no kernels): bit-exact tables and skip schedules."""
import torch

from conftest import load_golden


def test_linear_table_bit_exact():
    from degraders import NoiseDegradation
    fx = load_golden("schedules.pt")
    deg = NoiseDegradation(*fx["lin_args"])
    assert torch.equal(deg.beta, fx["lin_beta"]) and torch.equal(deg.alpha, fx["lin_alpha"])
    assert torch.equal(deg.alpha_cumulative_prod, fx["lin_abar"])
    steps = torch.arange(0, 1001)
    b, a, ab = deg.get_timestep_params(steps)
    assert torch.equal(b, fx["lin_beta"]) and torch.equal(a, fx["lin_alpha"]) and torch.equal(ab, fx["lin_abar"])
    hb, ha, hab = deg.host_params([1, 500, 1000])
    assert torch.equal(hab, fx["lin_abar"][[1, 500, 1000]])


def test_cosine_params_bit_exact():
    from degraders import CosineNoiseDegradation
    fx = load_golden("schedules.pt")
    deg = CosineNoiseDegradation(1000)
    b, a, ab = deg.get_timestep_params(torch.arange(1, 1001))
    assert torch.equal(b, fx["cos_beta"]) and torch.equal(a, fx["cos_alpha"]) and torch.equal(ab, fx["cos_abar"])
    assert torch.equal(deg.host_params(1000)[2], fx["cos_abar"][999:1000])


def test_skip_schedule_bit_exact():
    from diffusion_sampling_algorithms import skip_schedule
    fx = load_golden("samplers.pt")
    for sched in ("linear", "cosine"):
        for (mn, mx, step) in ((1, 1000, 20), (1, 1000, 100), (250, 1000, 37), (1, 60, 7)):
            assert skip_schedule(mn, mx, step) == fx[f"ddim_{sched}_{mn}_{mx}_{step}"]["calls"]
            assert skip_schedule(mn, mx, step) == fx[f"cold_{sched}_{mn}_{mx}_{step}"]["calls"]
    assert len(skip_schedule(1, 1000, 20)) == 51        # "DDIM-50" = 51 network evaluations


def test_enums():
    from diffusion_enums import DiffusionAlg, NoiseScheduler
    assert DiffusionAlg.DDPM.value == 0 and DiffusionAlg.DDIM.value == 1
    assert NoiseScheduler.LINEAR.value == 0 and NoiseScheduler.COSINE.value == 1
