"""CPU checks of round-2 host logic that needs no kernel launch: optimiser state adoption, shard-consistent
sampler noise, per-rank seeds, timestep range validation, the reference's flip draws."""
import pytest
import torch


def _tiny_net():
    from models.U_Net import U_Net
    torch.manual_seed(3)
    return U_Net(num_resnet_blocks=1, num_layers=1, attn_layers=[0], min_channel=32, max_channel=32, time_dim=32)


def test_fused_adam_adopts_a_torch_adam_state_dict():
    """`load_state_dict` of a checkpoint written by torch.optim.Adam (the reference's optimiser): the loaded moments must BE the
    flat buffers the kernels update (views), with the loaded values, step count and hyper-parameters."""
    from b200.optim import FusedAdam
    net = _tiny_net()
    lay = net.engine().grad_layout(torch.device("cpu"))
    lay.flatten_params()
    ref_params = [torch.nn.Parameter(p.detach().clone().contiguous()) for p in net.parameters()]
    ref = torch.optim.Adam(ref_params, lr=3e-4, betas=(0.5, 0.999))
    g = torch.Generator().manual_seed(0)
    live = {i for i, p in enumerate(net.parameters()) if id(p) in lay.offsets}
    for _ in range(3):
        for i, rp in enumerate(ref_params):
            rp.grad = torch.randn(rp.shape, generator=g) if i in live else None
        ref.step()
    sd = ref.state_dict()
    opt = FusedAdam(net.parameters(), lr=1.0, betas=(0.9, 0.9))
    opt.load_state_dict(sd)
    assert opt.param_groups[0]["lr"] == 3e-4 and tuple(opt.param_groups[0]["betas"]) == (0.5, 0.999)
    m_flat, v_flat = opt._flat[id(lay)]
    params = list(net.parameters())
    assert set(sd["state"]) == live
    for idx, st in sd["state"].items():
        p = params[idx]
        mine = opt.state[p]
        assert float(mine["step"]) == 3.0
        assert mine["exp_avg"].data_ptr() == lay._shaped(m_flat, p).data_ptr()          # a VIEW of the flat buffer, not a copy
        assert mine["exp_avg_sq"].data_ptr() == lay._shaped(v_flat, p).data_ptr()
        assert torch.equal(mine["exp_avg"], st["exp_avg"]) and torch.equal(mine["exp_avg_sq"], st["exp_avg_sq"])
    # and back: torch.optim.Adam accepts what FusedAdam saves
    again = torch.optim.Adam([torch.nn.Parameter(torch.zeros_like(p).contiguous()) for p in params])
    again.load_state_dict(opt.state_dict())
    assert again.param_groups[0]["betas"] == (0.5, 0.999) and len(again.state_dict()["state"]) == len(live)


def test_sharded_step_noise_is_a_slice_of_the_unsharded_draw():
    import diffusion_sampling_algorithms as S
    from b200.parallel import shard_range
    x = torch.zeros((7, 3, 4, 4))
    try:
        S.set_shard()
        torch.manual_seed(5)
        full = S._step_noise(x)
        parts = []
        for r in range(3):
            lo, hi = shard_range(7, r, 3)
            S.set_shard(lo, hi, 7)
            torch.manual_seed(5)
            parts.append(S._step_noise(x[lo:hi]))
            assert S._first_elem(x[lo:hi]) == lo * 48
        assert torch.equal(torch.cat(parts), full)
        S.set_shard(0, 7, 7)                     # the whole job is not a shard
        assert S._SHARD is None and S._first_elem(x) == 0
    finally:
        S.set_shard()


def test_rank_seeds_differ_and_dataset_shuffles_agree(tmp_path):
    import json
    from b200.trainer import rank_seed
    from custom_dataset._tables import load_tables
    assert len({rank_seed(11, r) for r in range(8)}) == 8 and rank_seed(11, 0) == 11
    db = {"Labels": {"1": {"labels": ["a", "b"]}},
          "Data": {str(i): {"filename": f"img{i}.png", "a": i % 2, "b": 1} for i in range(1, 30)}}
    path = tmp_path / "db.json"
    path.write_text(json.dumps(db))
    rows, labels = load_tables(str(path))
    assert labels == ["a", "b"] and len(rows) == 29
    from custom_dataset.conditional_img_dataset import ConditionalImgDataset
    a = ConditionalImgDataset(str(path), shuffle_seed=4)
    b = ConditionalImgDataset(str(path), shuffle_seed=4)          # another rank: same row order
    c = ConditionalImgDataset(str(path), shuffle_seed=5)
    assert [r[0] for r in a.dataset] == [r[0] for r in b.dataset] != [r[0] for r in c.dataset]
    assert sorted(r[0] for r in a.dataset) == sorted(f"img{i}.png" for i in range(1, 30))


def test_linear_schedule_rejects_out_of_range_steps_on_the_host():
    from degraders import CosineNoiseDegradation, NoiseDegradation
    deg = NoiseDegradation(5e-3, 9e-3, 100)
    assert float(deg.host_params(100)[2]) > 0
    for bad in (-1, 101, [5, 200]):
        with pytest.raises(IndexError):
            deg.host_params(bad)
    CosineNoiseDegradation(100).host_params(150)          # the closed form is defined for every t, like the reference


def test_flip_flags_follow_torchvision_per_image_draws():
    import torchvision
    from b200.image_io import draw_flip_flags
    x = torch.arange(8 * 3 * 2 * 5, dtype=torch.float32).reshape(8, 3, 2, 5)
    flip = torchvision.transforms.RandomHorizontalFlip(p=0.5)
    torch.manual_seed(21)
    want = torch.stack([flip(x[i]) for i in range(8)])          # train_diffusion.py:312-314
    torch.manual_seed(21)
    flags = draw_flip_flags(8)
    got = torch.where(flags.bool()[:, None, None, None], x.flip(-1), x)
    assert flags.dtype == torch.uint8 and torch.equal(got, want)


def test_flat_layout_aligns_every_tensor_to_128_elements():
    """Round 2b: every tensor of the flat parameter / gradient / moment buffers starts on a 128-element boundary (256 bytes in the
    bf16 copy TMA reads): with the old 8-element granularity each 128-byte weight row straddled two L2 lines.  The bucket ranges,
    the AdaGN front region and the channels-last views must stay consistent with the padded spans."""
    from b200.train_engine import GradLayout
    from models.U_Net import U_Net
    torch.manual_seed(1)
    net = U_Net(num_resnet_blocks=1, num_layers=2, attn_layers=[1], min_channel=64, max_channel=128, time_dim=64, cond_dim=10)
    lay = GradLayout(net, torch.device("cpu"))
    assert GradLayout.ALIGN == 128
    spans = []
    for p in lay.params:
        off = lay.offsets[id(p)]
        assert off % GradLayout.ALIGN == 0
        spans.append((off, off + p.numel()))
        assert lay.view(p).shape == p.shape                       # channels-last stored weights are exposed in the reference's shape
        assert lay.view(p).data_ptr() == lay.flat.data_ptr() + 4 * off
    spans.sort()
    assert all(a[1] <= b[0] for a, b in zip(spans, spans[1:])), "tensors overlap"
    assert lay.total % GradLayout.ALIGN == 0 and spans[-1][1] <= lay.total
    # the front region holds exactly the AdaGN scale Linears, and every other module maps to one contiguous range behind it
    assert lay.front_end == sum(GradLayout._span(p) for p in lay.params if lay.offsets[id(p)] < lay.front_end)
    for blk in list(net.down_layers) + list(net.up_layers) + [net.in_layer, net.middle_layer, net.out_layers]:
        lo, hi = lay.module_range(blk)
        assert lay.front_end <= lo < hi <= lay.total and lo % GradLayout.ALIGN == 0
    # a 3-element tensor (the bias of the last conv) no longer shifts what follows it
    last_bias = net.out_layers[1].conv_layer[0].bias
    assert last_bias.numel() == 3 and GradLayout._span(last_bias) == 128


def test_weight_gradient_batch_rows_describe_the_jobs(monkeypatch):
    """ops.conv2d_wgrad_batch hands b2_conv2d_wgrad_batch n rows of 11 values {mode, x, N, H, W, Cin, ldx, dz, Cout, lddz, grad};
    stride-2 jobs pass the parity planes' image count divided by four, channel slices keep their per-pixel stride."""
    import ctypes
    from b200 import ops
    seen = {}

    def fake_call(name, *args):
        seen["name"] = name
        n, desc = args[0], args[1]
        arr = ctypes.cast(desc, ctypes.POINTER(ctypes.c_longlong))
        seen["rows"] = [[arr[11 * i + j] for j in range(11)] for i in range(n)]

    monkeypatch.setattr(ops, "call", fake_call)
    monkeypatch.setattr(ops, "stream", lambda: None)
    wide = torch.zeros((2, 8, 8, 256), dtype=torch.bfloat16)
    x0, dz0, g0 = wide[..., 128:], torch.zeros((2, 8, 8, 64), dtype=torch.bfloat16), torch.zeros(9 * 64 * 128)
    planes, dz1, g1 = torch.zeros((8, 4, 4, 64), dtype=torch.bfloat16), torch.zeros((2, 4, 4, 128), dtype=torch.bfloat16), torch.zeros(9 * 64 * 128)
    ops.conv2d_wgrad_batch([(0, x0, dz0, 64, g0), (1, planes, dz1, 128, g1)])
    assert seen["name"] == "b2_conv2d_wgrad_batch"
    r0, r1 = seen["rows"]
    assert r0 == [0, x0.data_ptr(), 2, 8, 8, 128, 256, dz0.data_ptr(), 64, 64, g0.data_ptr()]
    assert r1 == [1, planes.data_ptr(), 2, 4, 4, 64, 64, dz1.data_ptr(), 128, 128, g1.data_ptr()]
    ops.conv2d_wgrad_batch([])                                       # nothing to do, nothing launched
    assert len(seen["rows"]) == 2


def test_conv128_autotune_stays_out_of_the_way(monkeypatch):
    """The run-time choice of the 128-channel conv implementation must not run when the user pinned a variant, switched the tuner off,
    asked for deterministic results, uses the TF32 parity mode, or the net's first level is not 128 channels wide."""
    import b200
    from b200 import autotune
    from models.U_Net import U_Net
    net = U_Net(num_resnet_blocks=1, num_layers=1, attn_layers=[], min_channel=128, max_channel=128, time_dim=32)
    narrow = _tiny_net()
    called = []
    monkeypatch.setattr(autotune, "_time", lambda fn, reps=5: called.append(1) or 1.0)
    monkeypatch.delenv("SDM_B200_HALO", raising=False)
    monkeypatch.delenv("SDM_B200_SWAP_AB", raising=False)
    monkeypatch.setenv("SDM_B200_AUTOTUNE", "0")
    assert autotune.tune_conv128(net, 8, 64, 64, "cpu") is None
    monkeypatch.setenv("SDM_B200_AUTOTUNE", "1")
    monkeypatch.setenv("SDM_B200_HALO", "1")
    assert autotune.tune_conv128(net, 8, 64, 64, "cpu") is None
    monkeypatch.delenv("SDM_B200_HALO")
    monkeypatch.setattr(b200, "_DETERMINISTIC", True)
    assert autotune.tune_conv128(net, 8, 64, 64, "cpu") is None
    monkeypatch.setattr(b200, "_DETERMINISTIC", False)
    net.precision = "tf32"
    assert autotune.tune_conv128(net, 8, 64, 64, "cpu") is None
    narrow.precision = "bf16"
    assert autotune.tune_conv128(narrow, 8, 64, 64, "cpu") is None
    net.precision = "bf16"
    if not torch.cuda.is_available():
        assert autotune.tune_conv128(net, 8, 64, 64, "cpu") is None      # no device, nothing to time
    assert not called
