"""Device-side image input / output edges (csrc/image_io.cu, b200/image_io.py) against the host code
they replace: numpy normalisation (custom_dataset/img_dataset.py:26-35), torchvision RandomHorizontalFlip per image
(train_diffusion.py:312-314), make_grid + save_image (utils/utils.py:39-65) and the uint8 cascade hand-off
(generate_sr_images_diffusion.py:106-126).  Byte / bit exact everywhere."""
import os

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def test_u8_to_image_is_bit_identical_to_the_reference_arithmetic():
    from b200.image_io import u8_to_image
    rng = np.random.RandomState(0)
    img = rng.randint(0, 256, (5, 24, 40, 3)).astype(np.uint8)
    img[0, 0, :6, 0] = [0, 1, 127, 128, 254, 255]
    want = torch.from_numpy((img.astype(float) - 127.5) / 127.5).float().permute(0, 3, 1, 2)      # img_dataset.py:29-33
    got = u8_to_image(torch.from_numpy(img).cuda())
    assert got.shape == (5, 3, 24, 40) and torch.equal(got.cpu(), want)
    flags = torch.tensor([1, 0, 1, 1, 0], dtype=torch.uint8)
    flipped = u8_to_image(torch.from_numpy(img).cuda(), flags)
    assert torch.equal(flipped.cpu(), torch.where(flags.bool()[:, None, None, None], want.flip(-1), want))


def test_flip_images_follows_the_reference_draws():
    import torchvision
    from b200.image_io import draw_flip_flags, flip_images
    x = torch.randn((6, 3, 8, 10))
    torch.manual_seed(5)
    flip = torchvision.transforms.RandomHorizontalFlip(p=0.5)
    want = torch.stack([flip(x[i]) for i in range(6)])              # train_diffusion.py:312-314, image by image
    torch.manual_seed(5)
    got = flip_images(x.cuda(), draw_flip_flags(6)).cpu()
    assert torch.equal(got, want)


@pytest.mark.parametrize("n", [1, 4, 7, 12])
def test_image_grid_matches_torchvision_bytes(n):
    import torchvision
    from b200.image_io import image_grid_u8
    x = torch.randn((n, 3, 16, 20)) * 0.8                           # values outside [-1, 1] exercise the clamp
    grid = torchvision.utils.make_grid(x[:, [2, 1, 0]], nrow=5, normalize=True, value_range=(-1, 1))
    want = grid.mul(255).add_(0.5).clamp_(0, 255).permute(1, 2, 0).to(torch.uint8)       # torchvision.utils.save_image
    got = image_grid_u8(x.cuda(), nrow=5, padding=2, value_range=(-1, 1), swap_rb=True).cpu()
    assert got.shape == want.shape and torch.equal(got, want)


def test_plot_sampled_images_writes_the_same_jpeg(tmp_path):
    from utils.utils import plot_sampled_images
    x = torch.rand((7, 3, 32, 32)) * 2 - 1
    quiet = lambda *a, **k: None
    plot_sampled_images(x.cuda(), "device", dest_path=str(tmp_path / "a"), log=quiet)
    plot_sampled_images(x, "host", dest_path=str(tmp_path / "b"), log=quiet)            # the reference's torchvision path
    a = open(os.path.join(tmp_path, "a", "plots", "device.jpg"), "rb").read()
    b = open(os.path.join(tmp_path, "b", "plots", "host.jpg"), "rb").read()
    assert len(a) > 1000 and a == b


def test_cascade_handoff_stays_on_the_device():
    """Samples -> uint8 HWC image -> [-1, 1] input of the next stage, both directions on the GPU, equal to the numpy route."""
    from b200.image_io import image_to_u8, u8_to_image
    x = torch.rand((3, 3, 16, 16)) * 2.4 - 1.2
    u8 = image_to_u8(x.cuda())
    want_u8 = ((x.clamp(-1, 1) + 1) / 2).mul(255).add(0.5).clamp(0, 255).to(torch.uint8).permute(0, 2, 3, 1)
    assert u8.dtype == torch.uint8 and u8.is_cuda and torch.equal(u8.cpu(), want_u8)
    back = u8_to_image(u8)
    want = torch.from_numpy((want_u8.numpy().astype(float) - 127.5) / 127.5).float().permute(0, 3, 1, 2)
    assert torch.equal(back.cpu(), want)
    assert float((back.cpu() - x.clamp(-1, 1)).abs().max()) <= 1.0 / 127.5


def test_device_image_loader_prefetches_and_matches_the_host_pipeline():
    from b200.image_io import DeviceImageLoader, draw_flip_flags
    from custom_dataset.img_dataset import SyntheticImages
    raw = SyntheticImages("synthetic:10x3x16x16:img", raw_uint8=True)
    loader = torch.utils.data.DataLoader(raw, batch_size=4, shuffle=False, num_workers=0)
    torch.manual_seed(11)
    got = list(DeviceImageLoader(loader, "cuda", flip_fn=draw_flip_flags))
    assert [b[0].shape[0] for b in got] == [4, 4, 2] and len(DeviceImageLoader(loader, "cuda")) == 3
    torch.manual_seed(11)
    iter(loader)             # a DataLoader iterator draws its base seed from the CPU generator when it is created: same stream position
    for i, (img, cond) in enumerate(got):
        assert img.is_cuda and img.dtype == torch.float32 and tuple(img.shape[1:]) == (3, 16, 16) and cond.shape == img.shape
        host = [raw[j] for j in range(4 * i, min(4 * i + 4, 10))]
        norm = lambda u: torch.from_numpy((u.numpy().astype(float) - 127.5) / 127.5).float().permute(2, 0, 1)
        flags = draw_flip_flags(len(host))                           # same CPU-generator draws, in hand-out order
        want = torch.stack([norm(h[0]).flip(-1) if f else norm(h[0]) for h, f in zip(host, flags)])
        assert torch.equal(img.cpu(), want)
        assert torch.equal(cond.cpu(), torch.stack([norm(h[1]) for h in host]))      # the condition image is never flipped
