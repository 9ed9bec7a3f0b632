"""Deterministic synthetic weights shared by the golden-fixture generator and the tests -- TEST INFRASTRUCTURE.

A fixture stores only (name, shape) pairs and a seed; both the reference model (in tests/golden/make_golden.py)
and the model under test are loaded with `synth_state_dict(shapes, seed)`, so 600 M-parameter checkpoints never
have to be committed.  Values: weights ~ U(-b, b) with b = 1 / sqrt(fan_in) (the bound of the reference default init), biases ~ U(-0.1, 0.1),
GroupNorm scales 1 + U(-0.2, 0.2) so that every parameter of the path influences the output.
"""
import math

import torch


def synth_state_dict(shapes, seed):
    g = torch.Generator(device="cpu")
    g.manual_seed(int(seed))
    sd = {}
    for name in sorted(shapes):
        shape = tuple(shapes[name])
        u = torch.rand(shape, generator=g, dtype=torch.float32) * 2 - 1
        if len(shape) >= 2:
            fan_in = 1
            for s in shape[1:]:
                fan_in *= s
            if "out_layer.conv_layer.0.weight" in name and name.startswith("up_layers"):
                fan_in = shape[0] * 4      # ConvTranspose2d [Cin, Cout, 4, 4]: 4 taps x Cin feed one output
            sd[name] = u * math.sqrt(1.0 / fan_in)
        elif name.endswith("norm.weight"):
            sd[name] = 1.0 + 0.2 * u
        else:
            sd[name] = 0.1 * u
    return sd
