"""In-kernel Philox noise for q-sample and the eps-prediction loss (north_star: "q(x_t|x_0) noising, the eps/x0 MSE loss ...
as single fused elementwise kernels with Philox RNG"; reference degraders.py:51-59, train_diffusion.py:310,336-350)."""
import pytest
import torch

from conftest import load_golden
from oracle.weights import synth_state_dict

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("sched", ["linear", "cosine"])
def test_qsample_philox_equals_qsample_of_the_same_draw(sched):
    from degraders import CosineNoiseDegradation, NoiseDegradation, PhiloxNoise
    deg = NoiseDegradation(5e-3, 9e-3, 1000, device="cuda") if sched == "linear" else CosineNoiseDegradation(1000)
    x0 = (torch.rand((6, 3, 32, 32), device="cuda") * 2 - 1)
    t = torch.tensor([1, 10, 250, 500, 999, 1000], device="cuda")
    noise = PhiloxNoise(seed=77, offset=3)
    eps = torch.empty_like(x0)
    x_t = deg.forward_philox(x0, t, noise, eps_out=eps)
    assert torch.equal(x_t, deg(x0, t, eps))                                   # same arithmetic on the stored draw
    assert abs(float(eps.mean())) < 2e-2 and abs(float(eps.std()) - 1.0) < 2e-2
    # the draw is a function of (seed, offset, GLOBAL element index): a shard / another rank sees its slice of the same stream
    per = x0[0].numel()
    half = deg.forward_philox(x0[3:], t[3:], PhiloxNoise(77, 3, None, first_elem=3 * per))
    assert torch.equal(half, x_t[3:])
    other = torch.empty_like(x0)
    deg.forward_philox(x0, t, PhiloxNoise(77, 4), eps_out=other)               # next optimisation step: fresh noise
    assert abs(float((eps * other).mean())) < 2e-2
    # a device-resident draw counter overrides the host value (CUDA-graph replays)
    ctr = torch.tensor([3.0], device="cuda")
    assert torch.equal(deg.forward_philox(x0, t, PhiloxNoise(77, 0, ctr)), x_t)


def test_mse_philox_regenerates_the_target():
    from b200.functional import mse_loss_philox
    from degraders import CosineNoiseDegradation, PhiloxNoise
    noise = PhiloxNoise(seed=5, offset=9, first_elem=4096)
    x0 = torch.zeros((4, 3, 16, 16), device="cuda")
    eps = torch.empty_like(x0)
    CosineNoiseDegradation(1000).forward_philox(x0, torch.tensor([500], device="cuda"), noise, eps_out=eps)
    pred = torch.randn_like(x0)
    loss, grad = mse_loss_philox(pred, noise)
    assert abs(float(loss) - float(torch.nn.functional.mse_loss(pred, eps))) < 1e-5
    assert torch.allclose(grad, 2 * (pred - eps) / pred.numel(), rtol=1e-5, atol=1e-8)


@pytest.mark.parametrize("kind", ["eps", "x0"])
def test_graph_replay_draws_fresh_noise_every_step(kind):
    """The captured step reads its draw number from the optimiser's device-side step count: two replays see different eps
    (different losses on the same batch), and each equals the eager step fed with that step's draw."""
    from b200.graph import GraphedTrainStep
    from b200.optim import FusedAdam
    from b200.steps import eps_prediction_step, x0_prediction_step
    from degraders import NoiseDegradation, PhiloxNoise
    from models.U_Net import U_Net
    fx = load_golden("unet_gpu_small.pt")
    dev = torch.device("cuda")

    def build(capturable):
        net = U_Net(**fx["kwargs"])
        net.load_state_dict(synth_state_dict(fx["shapes"], fx["seed"]))
        net = net.to(dev).train().set_precision("tf32")
        net.engine().grad_layout(dev).flatten_params()
        return net, FusedAdam(net.parameters(), lr=2e-4, betas=(0.5, 0.999), capturable=capturable)

    deg = NoiseDegradation(5e-3, 9e-3, 1000, device=dev)
    g = torch.Generator().manual_seed(2)
    x0 = (torch.rand((2, 3, 32, 32), generator=g) * 2 - 1).to(dev)
    t = torch.randint(1, 1000, (2,), generator=g).to(dev)
    net_g, opt_g = build(True)
    step = GraphedTrainStep(net_g, deg, opt_g, kind=kind, philox_seed=123, philox_first_elem=0)
    graph_losses = [float(step(x0, t, None)) for _ in range(3)]
    net_e, opt_e = build(False)
    fn = eps_prediction_step if kind == "eps" else x0_prediction_step
    eager_losses = [float(fn(net_e, deg, opt_e, x0, t, PhiloxNoise(123, k))) for k in range(3)]
    print(kind, graph_losses, eager_losses)
    assert len({round(v, 6) for v in graph_losses}) == 3                       # fresh noise each replay
    for a, b in zip(graph_losses, eager_losses):
        assert abs(a - b) < 2e-3 * abs(b)


def test_linear_table_rejects_out_of_range_timesteps():
    """the reference's gather raises on t > max_noise_step; here host-side steps raise and device-side steps
    poison the image with NaN (the trainer's NaN check trips) instead of reading outside the table."""
    from degraders import NoiseDegradation
    deg = NoiseDegradation(5e-3, 9e-3, 100, device="cuda")
    x = torch.ones((2, 3, 8, 8), device="cuda")
    with pytest.raises(IndexError):
        deg(x, torch.tensor([101]), torch.zeros_like(x))
    with pytest.raises(IndexError):
        deg.host_params(-1)
    out = deg(x, torch.tensor([100, 101], device="cuda"), torch.zeros_like(x))
    assert torch.isfinite(out[0]).all() and torch.isnan(out[1]).all()
