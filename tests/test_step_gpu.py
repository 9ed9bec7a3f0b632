"""GPU checks of the optimiser kernel and of whole training steps against the CPU oracle + torch.optim.Adam."""
import pytest
import torch
import torch.nn.functional as F

from conftest import load_golden, rel_l2
from oracle import diffusion_oracle as orc
from oracle.weights import synth_state_dict

pytestmark = pytest.mark.gpu


def test_fused_adam_matches_torch_adam():
    from b200.optim import FusedAdam
    torch.manual_seed(0)
    ps = [torch.randn(s, device="cuda") for s in ((128, 64, 3, 3), (1000,), (64, 64))]
    a = [torch.nn.Parameter(p.clone()) for p in ps]
    b = [torch.nn.Parameter(p.clone()) for p in ps]
    oa = FusedAdam(a, lr=2e-3, betas=(0.5, 0.999))
    ob = torch.optim.Adam(b, lr=2e-3, betas=(0.5, 0.999))
    for step in range(5):
        for x, y in zip(a, b):
            g = torch.randn_like(x) * (0.1 + step)
            x.grad, y.grad = g.clone(), g.clone()
        oa.step()
        ob.step()
    for x, y in zip(a, b):
        assert torch.allclose(x, y, rtol=1e-5, atol=1e-6)
    sa, sb = oa.state_dict()["state"], ob.state_dict()["state"]
    assert set(sa[0].keys()) == {"step", "exp_avg", "exp_avg_sq"} == set(sb[0].keys())
    assert torch.allclose(sa[0]["exp_avg_sq"], sb[0]["exp_avg_sq"], rtol=1e-5, atol=1e-9)


@pytest.mark.parametrize("sched_name", ["linear", "cosine"])
def test_training_steps_match_oracle(sched_name):
    """Three eps-prediction steps (train_diffusion.py:310-366) from identical weights / batches / t / eps."""
    from b200.optim import FusedAdam
    from b200.steps import eps_prediction_step
    from degraders import CosineNoiseDegradation, NoiseDegradation
    from models.U_Net import U_Net
    fx = load_golden("unet_gpu_small.pt")
    sd0 = synth_state_dict(fx["shapes"], fx["seed"])
    net = U_Net(**fx["kwargs"])
    net.load_state_dict(sd0)
    net = net.cuda().train().set_precision("tf32")
    net.engine().grad_layout(torch.device("cuda")).flatten_params()
    opt = FusedAdam(net.parameters(), lr=2e-4, betas=(0.5, 0.999))
    deg = NoiseDegradation(5e-3, 9e-3, 1000, device="cuda") if sched_name == "linear" else CosineNoiseDegradation(1000)
    osched = ("linear", 5e-3, 9e-3, 1000) if sched_name == "linear" else ("cosine", 1000)
    # oracle side: reference-format state dict as leaf tensors + the reference's optimiser (torch Adam)
    sd = {k: v.clone().requires_grad_(True) for k, v in sd0.items()}
    live = [v for k, v in sd.items() if ".y_shift." not in k and not (".attn_layers." in k and ".norm." in k)]
    oopt = torch.optim.Adam(live, lr=2e-4, betas=(0.5, 0.999))
    g = torch.Generator().manual_seed(5)
    losses, olosses = [], []
    for step in range(3):
        x0 = torch.rand((2, 3, 32, 32), generator=g) * 2 - 1
        eps = torch.randn((2, 3, 32, 32), generator=g)
        t = torch.randint(1, 1000, (2,), generator=g)
        loss = eps_prediction_step(net, deg, opt, x0.cuda(), t.cuda(), eps.cuda())
        losses.append(float(loss))
        oopt.zero_grad()
        oloss = orc.train_step_loss(sd, osched, x0, t, eps)
        oloss.backward()
        oopt.step()
        olosses.append(float(oloss))
    print("losses", losses, olosses)
    for a, b in zip(losses, olosses):
        assert abs(a - b) < 2e-3 * abs(b)
    # weights after 3 Adam steps: compare the UPDATE (w - w0), which is what training changes
    named = dict(net.named_parameters())
    num, den = 0.0, 0.0
    for k, v in sd.items():
        if not v.requires_grad or v.grad is None:
            continue
        du = named[k].detach().cpu() - sd0[k]
        dr = v.detach() - sd0[k]
        num += float((du - dr).double().pow(2).sum())
        den += float(dr.double().pow(2).sum())
    err = (num / den) ** 0.5
    print("relative error of the 3-step weight update:", err)
    assert err < 5e-2         # Adam's sign-like first steps amplify tiny gradient differences near zero crossings


def test_graphed_train_step_equals_eager_step():
    """CUDA-graph replay of the whole step (b200/graph.py) must train exactly like the eager kernel sequence."""
    from b200.graph import GraphedTrainStep
    from b200.optim import FusedAdam
    from b200.steps import eps_prediction_step
    from degraders import NoiseDegradation
    from models.U_Net import U_Net
    fx = load_golden("unet_gpu_small.pt")
    sd0 = synth_state_dict(fx["shapes"], fx["seed"])
    nets, opts = [], []
    for capturable in (False, True):
        net = U_Net(**fx["kwargs"])
        net.load_state_dict(sd0)
        net = net.cuda().train().set_precision("tf32")
        net.engine().grad_layout(torch.device("cuda")).flatten_params()
        nets.append(net)
        opts.append(FusedAdam(net.parameters(), lr=2e-4, betas=(0.5, 0.999), capturable=capturable))
    deg = NoiseDegradation(5e-3, 9e-3, 1000, device="cuda")
    graphed = GraphedTrainStep(nets[1], deg, opts[1], kind="eps")
    g = torch.Generator().manual_seed(11)
    for step in range(4):
        x0 = (torch.rand((2, 3, 32, 32), generator=g) * 2 - 1).cuda()
        eps = torch.randn((2, 3, 32, 32), generator=g).cuda()
        t = torch.randint(1, 1000, (2,), generator=g).cuda()
        if step == 2:                       # the reference halves the learning rate on the fly (train_diffusion.py:368-371)
            for o in opts:
                for grp in o.param_groups:
                    grp["lr"] *= 0.5
        l_eager = float(eps_prediction_step(nets[0], deg, opts[0], x0, t, eps))
        l_graph = float(graphed(x0, t, eps))
        assert abs(l_eager - l_graph) < 1e-4 * abs(l_eager), (step, l_eager, l_graph)
    assert graphed.replays == 4
    # compare the accumulated 4-step update (GroupNorm sums use fp32 atomics, so not bit-identical)
    p0 = {k: v.cuda() for k, v in sd0.items()}
    num = den = 0.0
    for (k, pa), (_, pb) in zip(nets[0].named_parameters(), nets[1].named_parameters()):
        da, db = pa.detach() - p0[k], pb.detach() - p0[k]
        num += float((da - db).double().pow(2).sum())
        den += float(da.double().pow(2).sum())
    assert (num / den) ** 0.5 < 2e-2
    st = opts[1].state_dict()["state"]
    assert float(st[0]["step"]) == 4.0


def test_bf16_shadow_weights_equal_packed_weights():
    """FusedAdam's bf16 copy of the channels-last stored weights (+ the transposing data-gradient pack) must train like
    the plain path that re-packs every kernel-layout weight from the fp32 masters (torch.optim.Adam: no shadow)."""
    from b200.optim import FusedAdam
    from b200.steps import eps_prediction_step
    from degraders import NoiseDegradation
    from models.U_Net import U_Net
    fx = load_golden("unet_gpu_small.pt")
    sd0 = synth_state_dict(fx["shapes"], fx["seed"])
    nets, opts = [], []
    for fused in (True, False):
        net = U_Net(**fx["kwargs"])
        net.load_state_dict(sd0)
        net = net.cuda().train().set_precision("bf16")
        net.engine().grad_layout(torch.device("cuda")).flatten_params()
        nets.append(net)
        live = [p for n_, p in net.named_parameters() if ".y_shift." not in n_ and not (".attn_layers." in n_ and ".norm." in n_)]
        opts.append(FusedAdam(net.parameters(), lr=2e-4, betas=(0.5, 0.999)) if fused
                    else torch.optim.Adam(live, lr=2e-4, betas=(0.5, 0.999)))
    deg = NoiseDegradation(5e-3, 9e-3, 1000, device="cuda")
    g = torch.Generator().manual_seed(3)
    for step in range(3):
        x0 = (torch.rand((2, 3, 32, 32), generator=g) * 2 - 1).cuda()
        eps = torch.randn((2, 3, 32, 32), generator=g).cuda()
        t = torch.randint(1, 1000, (2,), generator=g).cuda()
        la = float(eps_prediction_step(nets[0], deg, opts[0], x0, t, eps))
        lb = float(eps_prediction_step(nets[1], deg, opts[1], x0, t, eps))
        assert abs(la - lb) < 3e-3 * abs(lb), (step, la, lb)
    lay = nets[0].engine().layout
    assert lay.shadow is not None and len(lay.cl) > 0
    # the shadow is exactly bf16(master) and the permuted parameter views still have the reference's shapes
    assert torch.equal(lay.shadow, lay.params_flat.to(torch.bfloat16))
    for k, v in nets[0].state_dict().items():
        assert tuple(v.shape) == tuple(sd0[k].shape)
    # eval after training reads the same (current) weights through the cache
    nets[0].eval()
    with torch.no_grad():
        x = (torch.rand((2, 3, 32, 32), generator=g) * 2 - 1).cuda()
        tt = torch.tensor([5, 700]).cuda()
        y_shadow = nets[0](x, tt)
        lay.shadow = None
        nets[0].engine().cache.clear()
        y_packed = nets[0](x, tt)
    # bf16 activations + order-dependent fp32 atomics in the GroupNorm sums: equal up to a few bf16 roundings
    assert rel_l2(y_shadow, y_packed) < 5e-3


@pytest.mark.parametrize("graph", [False, True])
def test_bucketwise_overlapped_adam_equals_single_pass(graph):
    """DataParallel.attach_optimizer: Adam runs bucket by bucket on a second stream during backward (after each bucket's
    all-reduce; world size 1 here) -- same training trajectory as one Adam pass after backward."""
    from b200.graph import GraphedTrainStep
    from b200.optim import FusedAdam
    from b200.parallel import DataParallel
    from b200.steps import eps_prediction_step
    from degraders import NoiseDegradation
    from models.U_Net import U_Net
    fx = load_golden("unet_gpu_small.pt")
    sd0 = synth_state_dict(fx["shapes"], fx["seed"])
    nets, opts, dps = [], [], []
    for overlapped in (False, True):
        net = U_Net(**fx["kwargs"])
        net.load_state_dict(sd0)
        net = net.cuda().train().set_precision("tf32")
        dp = DataParallel(net, bucket_bytes=1 << 20, device=torch.device("cuda"))
        opt = FusedAdam(net.parameters(), lr=2e-4, betas=(0.5, 0.999), capturable=graph)
        if overlapped:
            dp.attach_optimizer(opt)
        nets.append(net); opts.append(opt); dps.append(dp)
    deg = NoiseDegradation(5e-3, 9e-3, 1000, device="cuda")
    steppers = [GraphedTrainStep(n_, deg, o_, kind="eps") if graph else None for n_, o_ in zip(nets, opts)]
    g = torch.Generator().manual_seed(21)
    for step in range(3):
        x0 = (torch.rand((2, 3, 32, 32), generator=g) * 2 - 1).cuda()
        eps = torch.randn((2, 3, 32, 32), generator=g).cuda()
        t = torch.randint(1, 1000, (2,), generator=g).cuda()
        losses = []
        for i in range(2):
            if graph:
                losses.append(float(steppers[i](x0, t, eps)))
            else:
                losses.append(float(eps_prediction_step(nets[i], deg, opts[i], x0, t, eps)))
        assert abs(losses[0] - losses[1]) < 1e-4 * abs(losses[0]), (step, losses)
    assert len(dps[1].launched) > 3                      # several buckets were updated separately
    p0 = {k: v.cuda() for k, v in sd0.items()}
    num = den = 0.0
    for (k, pa), (_, pb) in zip(nets[0].named_parameters(), nets[1].named_parameters()):
        da, db = pa.detach() - p0[k], pb.detach() - p0[k]
        num += float((da - db).double().pow(2).sum())
        den += float(da.double().pow(2).sum())
    assert (num / den) ** 0.5 < 2e-2
    assert float(opts[1].state_dict()["state"][0]["step"]) == 3.0


def test_bf16_gradient_transport_matches_fp32_transport():
    """DataParallel(grad_dtype="bf16"): buckets are cast to bf16 for the all-reduce and consumed as bf16 by the fused Adam
    (moments, weights, arithmetic fp32).  At world size 1 with the transport path forced, one step must equal the fp32-transport
    step up to the bf16 rounding of the gradient (Adam normalises it: the update moves by < 1 %)."""
    from b200.optim import FusedAdam
    from b200.parallel import DataParallel
    from b200.steps import eps_prediction_step
    from degraders import NoiseDegradation
    from models.U_Net import U_Net
    fx = load_golden("unet_gpu_small.pt")
    sd0 = synth_state_dict(fx["shapes"], fx["seed"])
    dev = torch.device("cuda")
    deg = NoiseDegradation(5e-3, 9e-3, 1000, device=dev)
    g = torch.Generator().manual_seed(5)
    x0 = (torch.rand((2, 3, 32, 32), generator=g) * 2 - 1).to(dev)
    eps = torch.randn((2, 3, 32, 32), generator=g).to(dev)
    t = torch.randint(1, 1000, (2,), generator=g).to(dev)
    nets = []
    for dtype in ("fp32", "bf16"):
        net = U_Net(**fx["kwargs"])
        net.load_state_dict(sd0)
        net = net.to(dev).train().set_precision("tf32")
        dp = DataParallel(net, device=dev, grad_dtype=dtype, bucket_bytes=1 << 20, _force_transport=True)
        opt = FusedAdam(net.parameters(), lr=2e-4, betas=(0.5, 0.999), grad_scale=dp.grad_scale)
        dp.attach_optimizer(opt)
        for _ in range(2):
            eps_prediction_step(net, deg, opt, x0, t, eps)
        assert (dp.g16 is not None) == (dtype == "bf16") and len(dp.launched) > 2
        nets.append(net)
    num = den = 0.0
    for (k, pa), (_, pb) in zip(nets[0].named_parameters(), nets[1].named_parameters()):
        num += float((pa.detach() - pb.detach()).double().pow(2).sum())
        den += float((pa.detach().cpu() - sd0[k]).double().pow(2).sum())
    err = (num / den) ** 0.5
    print("bf16 vs fp32 gradient transport, update rel-L2 after 2 steps:", err)
    assert err < 5e-2
