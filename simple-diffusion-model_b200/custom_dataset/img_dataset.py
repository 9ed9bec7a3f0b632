"""Image-folder dataset with the reference's interface (custom_dataset/img_dataset.py:11-39): cv2 BGR images scaled
to [-1, 1], CHW fp32.  `SyntheticImages` is an addition for smoke tests and benchmarks (no files needed)."""
import torch
from torch.utils.data import Dataset


def read_image(path, raw_uint8=False):
    """raw_uint8 (addition): return the decoded uint8 HWC BGR bytes; normalisation then happens on the device
    (b200/image_io.py: 4x fewer bytes over PCIe, bit-identical values)."""
    import cv2
    img = cv2.imread(path)
    if img is None:
        raise Exception(f"Could not read image: {path}")
    if raw_uint8:
        return torch.from_numpy(img)
    return torch.from_numpy((img.astype(float) - 127.5) / 127.5).float().permute(2, 0, 1)


class ImageDataset(Dataset):
    def __init__(self, img_paths=[], return_filepaths=False, raw_uint8=False):
        self.img_paths = img_paths
        self.return_filepaths = return_filepaths
        self.raw_uint8 = raw_uint8

    def __len__(self):
        return len(self.img_paths)

    def __getitem__(self, index):
        path = self.img_paths[index]
        img = read_image(path, self.raw_uint8)
        return (img, path) if self.return_filepaths else img


class SyntheticImages(Dataset):
    """`synthetic:<count>x<C>x<H>x<W>[:<cond_dim>|:img]` -- U(-1,1) images (value range of (img-127.5)/127.5), optional
    multi-hot labels or a conditioning image, generated deterministically per index."""

    def __init__(self, spec, raw_uint8=False):
        self.raw_uint8 = raw_uint8
        parts = spec.split(":")
        self.count, self.c, self.h, self.w = (int(v) for v in parts[1].split("x"))
        extra = parts[2] if len(parts) > 2 else None
        self.cond_dim = int(extra) if extra not in (None, "img") else None
        self.cond_img = extra == "img"

    def get_labels(self):
        return [f"label_{i}" for i in range(self.cond_dim or 0)]

    def __len__(self):
        return self.count

    def __getitem__(self, index):
        g = torch.Generator().manual_seed(1234 + index)

        def image():
            if self.raw_uint8:      # decoded-file format: uint8 HWC
                return torch.randint(0, 256, (self.h, self.w, self.c), generator=g, dtype=torch.uint8)
            return torch.rand((self.c, self.h, self.w), generator=g) * 2 - 1

        img = image()
        if self.cond_dim is not None:
            return img, (torch.rand((self.cond_dim,), generator=g) > 0.7).float()
        if self.cond_img:
            return img, image()
        return img
