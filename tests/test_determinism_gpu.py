"""Run-to-run and shard-to-shard reproducibility.

Deterministic mode (b2_set_deterministic / SDM_B200_DETERMINISTIC=1):
  * the weight-gradient GEMM's split-K is ORDERED (per-split partial tiles, last arriver sums them in split order): repeated
    launches on the same operands are bitwise equal -- with the default fp32 atomics they are not;
  * GroupNorm statistics without atomics and no split-K on the forward kernel, so an image's result does not depend on batch
    size -- batch-sharded DDIM sampling from Philox x_T is bitwise identical to the unsharded run."""
import pytest
import torch

from conftest import load_golden
from oracle.weights import synth_state_dict

pytestmark = pytest.mark.gpu


@pytest.fixture
def deterministic():
    import b200
    b200.set_deterministic(True)
    yield
    b200.set_deterministic(False)


@pytest.mark.parametrize("shape", [(4, 64, 64, 128, 128), (32, 16, 16, 256, 512), (2, 8, 8, 512, 512)])
def test_weight_gradient_split_k_is_bitwise_repeatable(deterministic, shape):
    from b200 import ops
    n, h, w, cin, cout = shape
    g = torch.Generator(device="cuda").manual_seed(0)
    x = torch.randn((n, h, w, cin), device="cuda", generator=g).bfloat16()
    dz = torch.randn((n, h, w, cout), device="cuda", generator=g).bfloat16()
    outs = []
    for _ in range(4):
        grad = torch.zeros((cout, 9 * cin), device="cuda")
        ops.conv2d_wgrad(0, x, dz, cout, grad)
        outs.append(grad)
    torch.cuda.synchronize()
    assert all(torch.equal(outs[0], o) for o in outs[1:])
    # and it is still the right gradient: fp32 reference of dW[co][tap][ci] = sum_p dz[p][co] * x[p + tap][ci]
    xf, dzf = x.float().permute(0, 3, 1, 2), dz.float().permute(0, 3, 1, 2)
    want = torch.nn.grad.conv2d_weight(xf, (cout, cin, 3, 3), dzf, padding=1).permute(0, 2, 3, 1).reshape(cout, 9 * cin)
    err = float((outs[0] - want).norm() / want.norm())
    assert err < 2e-3, err


@pytest.mark.parametrize("precision", ["bf16", "tf32"])
def test_forward_is_batch_invariant_in_deterministic_mode(deterministic, precision):
    from models.U_Net import U_Net
    fx = load_golden("unet_gpu_small.pt")
    net = U_Net(**fx["kwargs"])
    net.load_state_dict(synth_state_dict(fx["shapes"], fx["seed"]))
    net = net.cuda().eval().set_precision(precision)
    g = torch.Generator(device="cuda").manual_seed(3)
    x = torch.randn((6, 3, 32, 32), device="cuda", generator=g)
    t = torch.tensor([400], device="cuda")
    with torch.no_grad():
        full = net(x, t)
        again = net(x, t)
        parts = torch.cat([net(x[:2], t), net(x[2:3], t), net(x[3:], t)])
    assert torch.equal(full, again)
    assert torch.equal(full, parts)


def test_sharded_ddim_equals_unsharded_bitwise(deterministic):
    """sampler sharded vs unsharded bit-identical.  x_T comes from Philox keyed on the global element index
    (what every rank of a batch-sharded job would draw for its slice), the net is real, the loop is the public sampler."""
    import diffusion_sampling_algorithms as S
    from b200._lib import call, ptr, stream
    from b200.parallel import shard_range
    from degraders import NoiseDegradation
    from models.U_Net import U_Net
    fx = load_golden("unet_gpu_small.pt")
    net = U_Net(**fx["kwargs"])
    net.load_state_dict(synth_state_dict(fx["shapes"], fx["seed"]))
    net = net.cuda().eval().set_precision("bf16")
    deg = NoiseDegradation(5e-3, 9e-3, 1000, device="cuda")
    total, per = 6, 3 * 32 * 32

    def x_T(lo, hi):
        out = torch.empty((hi - lo, 3, 32, 32), device="cuda")
        call("b2_philox_normal", ptr(out), out.numel(), 2024, 0, lo * per, stream())
        return out

    run = lambda x: S.ddim_sampling(net, deg, x, min_noise=1, max_noise=1000, ddim_step_size=100, device="cuda",
                                    log=lambda *a, **k: None)
    full = run(x_T(0, total))
    for world in (2, 3):
        shards = [run(x_T(*shard_range(total, r, world))) for r in range(world)]
        assert torch.equal(torch.cat(shards), full), world
    assert torch.isfinite(full).all()


def test_sharded_ddpm_noise_follows_the_unsharded_draw():
    """under torchrun every rank used to add the SAME per-step z to different images.  With set_shard the draw is
    made for the whole job and sliced (torch generator) or keyed on the global element index (Philox): the union of the
    shards equals the single-process result, bit for bit (elementwise stub network)."""
    import diffusion_sampling_algorithms as S
    from b200.parallel import shard_range
    from degraders import NoiseDegradation

    class Stub(torch.nn.Module):
        def forward(self, x, t, labels=None):
            return 0.3 * torch.tanh(x) + 1e-4 * t.float()[:, None, None, None]

    deg = NoiseDegradation(5e-3, 9e-3, 1000, device="cuda")
    x_T = torch.randn((5, 3, 8, 8), device="cuda")
    quiet = lambda *a, **k: None
    for philox in (None, 99):
        S.PHILOX_SEED = philox
        try:
            S.set_shard()
            torch.manual_seed(7)
            full = S.ddpm_sampling(Stub(), deg, x_T.clone(), min_noise=1, max_noise=30, device="cuda", log=quiet)
            parts = []
            for r in range(2):
                lo, hi = shard_range(5, r, 2)
                S.set_shard(lo, hi, 5)
                torch.manual_seed(7)
                parts.append(S.ddpm_sampling(Stub(), deg, x_T[lo:hi].clone(), min_noise=1, max_noise=30, device="cuda", log=quiet))
            assert torch.equal(torch.cat(parts), full), philox
        finally:
            S.set_shard()
            S.PHILOX_SEED = None
