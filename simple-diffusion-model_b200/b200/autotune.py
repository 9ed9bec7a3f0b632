"""Run-time choice between the three implementations of the 128-channel 3x3 convolutions (csrc/conv_api.cu: halo tiles, per-tap
loads, swapped operands).  Which one wins depends on the machine: on the B200s of round 2's first session the halo tiles led
(965 vs 884 per-tap vs 848 swapped TFLOP/s at batch 256, profiles/r02n_conv128_variants.md); on the boxes of the second session the
halo and per-tap forms fell to 705 / 715 TFLOP/s -- everything bound by L2 -> SM traffic slowed there -- while the swapped form kept
its 846 (profiles/r02z13_conv128_variants_today.log).  The captured graphs therefore time the candidates once per workload
signature (a few milliseconds, before the warm-up steps) and set the library's process-wide switches to the winner.
SDM_B200_HALO / SDM_B200_SWAP_AB (explicit choices) and SDM_B200_AUTOTUNE=0 disable the measurement."""
import os

import torch

from . import ops

_CHOICE = {}          # (n, h, w, device index) -> (halo, swap_ab)
VARIANTS = ((1, 0), (0, 0), (0, 1))          # (halo, swap_ab): halo tiles, per-tap loads, swapped operands


def set_option(name, value):
    from . import set_option as _set
    _set(name, value)


def _time(fn, reps=5):
    for _ in range(2):
        fn()
    torch.cuda.synchronize()
    best = float("inf")
    for _ in range(reps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        fn()
        e1.record()
        torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1))
    return best


def tune_conv128(net, n, h, w, device):
    """Times the forward 128 -> 128 channel 3x3 conv of the net's top level (conv + bias + Swish + GroupNorm sums, bf16) in its
    three forms at this workload's shape and selects the fastest for the process.  Returns the (halo, swap_ab) pair in force."""
    if os.environ.get("SDM_B200_AUTOTUNE", "1") == "0" or "SDM_B200_HALO" in os.environ or "SDM_B200_SWAP_AB" in os.environ:
        return None
    from . import is_deterministic
    if getattr(net, "precision", "bf16") != "bf16" or is_deterministic():
        return None          # deterministic mode: one fixed kernel choice, whatever the batch size (sharded == unsharded bitwise)
    first = net.in_layer[1].conv_layer[0]
    if tuple(first.weight.shape[:2]) != (128, 128):
        return None          # only the class-default width has dedicated variants
    if not torch.cuda.is_available() or torch.cuda.is_current_stream_capturing():
        return None
    dev = torch.device(device)
    key = (n, h, w, dev.index if dev.index is not None else torch.cuda.current_device())
    if key not in _CHOICE:
        c = 128
        g = torch.Generator(device=dev).manual_seed(0)
        x = (torch.randn((n, h, w, c), device=dev, generator=g) * 0.5).bfloat16()
        wt = (torch.randn((c, 9 * c), device=dev, generator=g) * 0.02).bfloat16()
        bias = torch.zeros(c, device=dev)
        stats = torch.zeros((n, 32, 2), device=dev)
        y = torch.empty_like(x)
        times = []
        for halo, swap in VARIANTS:
            set_option("halo", halo)
            set_option("swap_ab", swap)
            times.append(_time(lambda: ops.conv2d(0, x, wt, bias, c, act=1, out=y, gn_stats=stats, groups=32)))
        best = min(range(len(VARIANTS)), key=lambda i: times[i])
        # keep the default (halo tiles) unless another form wins by more than timing noise
        if best != 0 and times[best] > 0.97 * times[0]:
            best = 0
        _CHOICE[key] = VARIANTS[best]
        if os.environ.get("SDM_B200_AUTOTUNE_LOG"):
            print(f"[b200.autotune] conv128 N={n} {h}x{w}: halo {times[0]:.3f} ms, per-tap {times[1]:.3f} ms, swapped {times[2]:.3f} ms"
                  f" -> {('halo', 'per-tap', 'swapped')[best]}", flush=True)
    halo, swap = _CHOICE[key]
    set_option("halo", halo)
    set_option("swap_ab", swap)
    return halo, swap


def reset_conv128():
    """Back to the library's default choice (halo tiles) unless the environment pins one: the train step keeps the default -- see
    GraphedTrainStep._capture -- even when an inference graph captured earlier in the process selected another form."""
    if "SDM_B200_HALO" in os.environ or "SDM_B200_SWAP_AB" in os.environ:
        return
    set_option("halo", 1)
    set_option("swap_ab", 0)
