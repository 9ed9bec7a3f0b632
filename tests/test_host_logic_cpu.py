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
