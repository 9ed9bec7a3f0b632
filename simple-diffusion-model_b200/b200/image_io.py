"""Device-side image input / output edges.

The reference normalises every image on the host with numpy (custom_dataset/img_dataset.py:26-35), flips image by image
in Python (train_diffusion.py:312-314), ships fp32 tensors over PCIe, and turns samples back into pictures with torchvision
on the host (utils/utils.py:39-65).  Here the bytes cross PCIe as uint8 (4x fewer), and normalisation, flips, the sample
grid and the uint8 cascade hand-off are single kernels (csrc/image_io.cu).  Everything is bit-identical to the host path:
the normalisation is evaluated in double and rounded once, the quantisation reproduces save_image's mul/add/clamp/truncate.
"""
import threading

import torch

from ._lib import B200Error, call, ptr, stream


def _cuda(t, what):
    if not t.is_cuda:
        raise B200Error(f"{what} needs a CUDA tensor: this build has no CPU path")


def u8_to_image(u8_nhwc, flip=None):
    """uint8 [N, H, W, C] (cv2 BGR bytes, CUDA) -> fp32 [N, C, H, W] in [-1, 1]; `flip`: optional bool/uint8 [N] flags."""
    _cuda(u8_nhwc, "u8_to_image")
    if u8_nhwc.dtype != torch.uint8 or u8_nhwc.dim() != 4:
        raise B200Error("u8_to_image expects a uint8 [N, H, W, C] tensor")
    src = u8_nhwc.contiguous()
    n, h, w, c = src.shape
    out = torch.empty((n, c, h, w), dtype=torch.float32, device=src.device)
    flags = None if flip is None else flip.to(device=src.device, dtype=torch.uint8).contiguous()
    call("b2_u8_to_image", ptr(src), ptr(out), ptr(flags), n, h, w, c, stream())
    return out


def flip_images(x, flip):
    """Horizontal flip of the images of an fp32 NCHW batch whose flag is set (RandomHorizontalFlip applied image by image)."""
    _cuda(x, "flip_images")
    xc = x.contiguous().float()
    n, c, h, w = xc.shape
    out = torch.empty_like(xc)
    flags = flip.to(device=xc.device, dtype=torch.uint8).contiguous()
    call("b2_flip_images", ptr(xc), ptr(out), ptr(flags), n, c, h, w, stream())
    return out


def draw_flip_flags(n, p=0.5):
    """The reference's per-image flip decisions: one `torch.rand(1) < p` draw from the CPU generator per image, in order
    (torchvision RandomHorizontalFlip.forward called once per image, train_diffusion.py:312-314)."""
    return torch.tensor([bool(torch.rand(1) < p) for _ in range(n)], dtype=torch.uint8)


def image_to_u8(x, value_range=(-1.0, 1.0)):
    """fp32 [N, C, H, W] (or [C, H, W]) in `value_range` -> uint8 [N, H, W, C] on the device: the uint8-range HWC BGR image the
    super-resolution entry point takes as `lr_img` (generate_sr_images_diffusion.py:106-126), without a host round trip."""
    _cuda(x, "image_to_u8")
    xc = (x if x.dim() == 4 else x.unsqueeze(0)).contiguous().float()
    n, c, h, w = xc.shape
    out = torch.empty((n, h, w, c), dtype=torch.uint8, device=xc.device)
    call("b2_image_to_u8", ptr(xc), ptr(out), n, c, h, w, float(value_range[0]), float(value_range[1]), stream())
    return out


def image_grid_u8(x, nrow=5, padding=2, value_range=(-1.0, 1.0), swap_rb=True):
    """plot_sampled_images' picture (utils/utils.py:39-65) as ONE kernel: BGR -> RGB, torchvision make_grid(nrow, padding,
    normalize=True, value_range) and save_image's quantisation; returns the uint8 [GH, GW, C] picture on the device."""
    _cuda(x, "image_grid_u8")
    xc = x.contiguous().float()
    n, c, h, w = xc.shape
    if n == 1:
        gh, gw = h, w
    else:
        xmaps = min(nrow, n)
        ymaps = (n + xmaps - 1) // xmaps
        gh, gw = ymaps * (h + padding) + padding, xmaps * (w + padding) + padding
    grid = torch.empty((gh, gw, c), dtype=torch.uint8, device=xc.device)
    call("b2_image_grid_u8", ptr(xc), ptr(grid), n, c, h, w, int(nrow), int(padding), 1 if swap_rb else 0, float(value_range[0]),
         float(value_range[1]), stream())
    return grid


class DeviceImageLoader:
    """Double-buffered uint8 input pipeline: wraps an iterable of host batches whose first element is a uint8 [N, H, W, C]
    image tensor (datasets built with `raw_uint8=True`) and yields device batches with the image normalised to fp32
    [N, C, H, W] in [-1, 1].  While the trainer works on batch i, batch i+1 is staged in pinned memory and copied host ->
    device on a side stream; the consumer's stream only waits on that copy's event.  Flips are NOT applied here (their random
    draws must interleave with the trainer's other draws): use `u8_to_image(..., flip=flags)` semantics through
    the `flip_fn` hook, called as flip_fn(n) -> uint8 [n] flags at hand-out time."""

    def __init__(self, batches, device, flip_fn=None, depth=2):
        self.batches, self.device, self.flip_fn, self.depth = batches, torch.device(device), flip_fn, max(1, depth)
        self.copy_stream = torch.cuda.Stream(device=self.device)
        self._pinned = {}

    def __len__(self):
        return len(self.batches)

    def _stage(self, batch, slot):
        """Host batch -> (device tensors, ready event); the uint8 image goes through a re-used pinned staging buffer."""
        items = list(batch) if isinstance(batch, (list, tuple)) else [batch]
        out = []
        with torch.cuda.stream(self.copy_stream):
            for j, t in enumerate(items):
                if not torch.is_tensor(t):
                    out.append(t)
                    continue
                if not t.is_pinned():
                    key = (slot, j, tuple(t.shape), t.dtype)
                    buf = self._pinned.get(key)
                    if buf is None:
                        buf = torch.empty(t.shape, dtype=t.dtype).pin_memory()
                        self._pinned[key] = buf
                    buf.copy_(t)
                    t = buf
                out.append(t.to(self.device, non_blocking=True))
            ev = torch.cuda.Event()
            ev.record(self.copy_stream)
        return out, ev

    def __iter__(self):
        it = iter(self.batches)
        queue, slot = [], 0
        try:
            for _ in range(self.depth):
                queue.append(self._stage(next(it), slot))
                slot = (slot + 1) % (self.depth + 1)
        except StopIteration:
            pass
        while queue:
            tensors, ev = queue.pop(0)
            try:
                queue.append(self._stage(next(it), slot))
                slot = (slot + 1) % (self.depth + 1)
            except StopIteration:
                pass
            cur = torch.cuda.current_stream(self.device)
            cur.wait_event(ev)
            for j, t in enumerate(tensors):
                if not torch.is_tensor(t):
                    continue
                t.record_stream(cur)
                if t.dtype == torch.uint8 and t.dim() == 4:          # image bytes (the doodle flavour carries two of them)
                    flags = self.flip_fn(t.shape[0]) if (self.flip_fn is not None and j == 0) else None
                    tensors[j] = u8_to_image(t, flags)
            yield tensors if len(tensors) > 1 else tensors[0]
