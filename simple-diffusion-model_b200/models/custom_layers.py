"""Building blocks of the U-Net -- same class names, constructor signatures, parameter names and shapes as the
reference's models/custom_layers.py, so reference checkpoints load unchanged -- but every `forward` runs hand-written
sm_100a kernels (b200/engine.py).  torch.nn.{Conv2d, ConvTranspose2d, Linear, GroupNorm} objects are used purely as
fp32 parameter containers / initialisers; their own forward methods are never called.

Reference behaviours kept on purpose (all verified against the reference):
  * AdaGN's shift re-uses the scale Linear: out = s * GN(x) + s; `y_shift` holds parameters but is dead;
  * AttentionBlock's GroupNorm is never applied; its softmax normalises over the QUERY axis;
  * post-activation norm order Conv -> Swish -> AdaGN; ResidualBlock's shortcut is the identity when C_in == C_out.
"""
from enum import Enum

import torch
import torch.nn as nn


class UNetBlockType(Enum):
    UP = 0
    DOWN = 1


def _standalone(module, fn, x, *args):
    """Runs one block on NCHW fp32 CUDA input through the engine of a scratch host (standalone-block API)."""
    from b200.blocks import run_standalone
    return run_standalone(module, fn, x, *args)


class Swish(nn.Module):
    """x * sigmoid(x).  Inside the network this never runs as its own pass (it is a conv/GEMM epilogue)."""

    def forward(self, x):
        from b200.blocks import swish_standalone
        return swish_standalone(x)


class AdaGN(nn.Module):
    """Adaptive GroupNorm (custom_layers.py:26-45).  In the engine this is ONE streaming pass (`b2_adagn_apply`): the
    GroupNorm sums were already produced by the epilogue of the convolution that wrote `x`, and the `y_scale` Linears of
    every AdaGN in the net are evaluated together as a single small GEMM per forward (engine.Engine._scales)."""
    def __init__(self, emb_dim, out_dim, groups=32):
        super().__init__()
        self.y_scale = nn.Linear(emb_dim, out_dim)
        self.y_shift = nn.Linear(emb_dim, out_dim)     # never used by forward (reference quirk); kept for checkpoints
        self.group_norm = nn.GroupNorm(groups, out_dim)

    def forward(self, x, emb):
        return _standalone(self, "adagn", x, emb)


class ConditionalEmbedding(nn.Module):
    """Sinusoidal timestep features -> 4-layer MLP, plus an optional second MLP over the condition vector whose output is
    added (custom_layers.py:51-98).  Engine: `b2_sinusoid` + `b2_small_gemm` with the Swish fused; a few microseconds."""
    def __init__(self, time_dim, cond_dim=None):
        super().__init__()
        self.time_dim = time_dim
        self.cond_dim = cond_dim

        def mlp(d_in):
            return nn.Sequential(nn.Linear(d_in, time_dim), Swish(), nn.Linear(time_dim, time_dim), Swish(),
                                 nn.Linear(time_dim, time_dim), Swish(), nn.Linear(time_dim, time_dim))

        self.time_layer = mlp(time_dim)
        self.cond_layer = mlp(cond_dim) if cond_dim is not None else None

    def forward(self, t, cond=None):
        return _standalone(self, "embedding", t, cond)


class AttentionBlock(nn.Module):
    """Self-attention over pixels with a residual (custom_layers.py:104-160).  Engine: QKV projection as a tcgen05 GEMM,
    scores + query-axis softmax fused in one kernel epilogue (`b2_attn_scores_softmax`), P.V as a TN GEMM on the
    un-transposed V, output projection with the residual added in its epilogue."""
    def __init__(self, channels, heads=1, d_k=None, groups=32):
        super().__init__()
        if d_k is None:
            d_k = channels
        self.norm = nn.GroupNorm(groups, channels)     # declared, never applied (reference quirk)
        self.projection = nn.Linear(channels, heads * d_k * 3)
        self.output = nn.Linear(heads * d_k, channels)
        self.scale = d_k ** -0.5
        self.heads = heads
        self.d_k = d_k

    def forward(self, x, t=None):
        _ = t
        return _standalone(self, "attention", x)


class UpsampleBlock(nn.Module):
    """ConvTranspose2d(4, stride 2, pad 1) + Swish (custom_layers.py:169-186).  Engine: four parity sub-convolutions, each
    a 2x2-tap implicit GEMM writing its interleaved quarter of the output through a strided TMA-free epilogue."""
    def __init__(self, in_channels, out_channels):
        super().__init__()
        self.conv_layer = nn.Sequential(
            nn.ConvTranspose2d(in_channels, out_channels, kernel_size=4, stride=2, padding=1), Swish())

    def forward(self, x, emb=None):
        _ = emb
        return _standalone(self, "upsample", x)


class DownsampleBlock(nn.Module):
    """Conv2d(3, stride 2, pad 1) + Swish (custom_layers.py:191-208).  Engine: space-to-depth of the input, then a 2x2-tap
    stride-1 implicit GEMM over 4x the channels (weights re-packed once per optimiser step)."""
    def __init__(self, in_channels, out_channels):
        super().__init__()
        self.conv_layer = nn.Sequential(
            nn.Conv2d(in_channels, out_channels, kernel_size=3, stride=2, padding=1), Swish())

    def forward(self, x, emb=None):
        _ = emb
        return _standalone(self, "downsample", x)


class UNet_ConvBlock(nn.Module):
    """Conv 3x3 -> Swish -> AdaGN (custom_layers.py:213-245): the hot layer.  Engine: one persistent tcgen05 implicit-GEMM
    launch (bias + Swish + GroupNorm partial sums in the epilogue) followed by one `b2_adagn_apply` pass."""
    def __init__(self, in_channels, out_channels, use_activation=True, emb_dim=None, groups=32):
        super().__init__()
        layers = [nn.Conv2d(in_channels, out_channels, kernel_size=3, padding=1)]
        if use_activation:
            layers.append(Swish())
        self.conv_layer = nn.Sequential(*layers)
        self.use_activation = use_activation
        if emb_dim is not None:
            self.adagn = AdaGN(emb_dim, out_channels, groups=groups)
        else:
            self.adagn = None

    def forward(self, x, emb=None):
        return _standalone(self, "conv_block", x, emb)


class ResidualBlock(nn.Module):
    """x + block2(block1(x)) (custom_layers.py:251-287); the skip add rides in the second block's AdaGN apply pass."""
    def __init__(self, in_channels, out_channels, use_activation=True, emb_dim=None, groups=32):
        super().__init__()
        self.conv_block_1 = UNet_ConvBlock(in_channels=in_channels, out_channels=out_channels,
                                           use_activation=use_activation, emb_dim=emb_dim, groups=groups)
        self.conv_block_2 = UNet_ConvBlock(in_channels=in_channels, out_channels=out_channels,
                                           use_activation=use_activation, emb_dim=emb_dim, groups=groups)
        if in_channels != out_channels:
            # Unreachable from U_Net (hidden == in everywhere); parameters kept for state_dict parity.
            self.shortcut = nn.Conv2d(in_channels, out_channels, kernel_size=(1, 1))
        else:
            self.shortcut = nn.Identity()

    def forward(self, x, emb=None):
        return _standalone(self, "residual_block", x, emb)


class UNetBlock(nn.Module):
    """`num_resnet_blocks` x (ResidualBlock, AttentionBlock | Identity) then a down- or up-sampling layer
    (custom_layers.py:293-350)."""
    def __init__(self, in_channels, out_channels, emb_dim, num_resnet_blocks=1, use_attn=True, num_heads=1,
                 dim_per_head=None, groups=32, block_type=UNetBlockType.DOWN):
        super().__init__()
        hidden_channels = in_channels
        self.res_layers = nn.ModuleList()
        self.attn_layers = nn.ModuleList()
        for _ in range(num_resnet_blocks):
            # NB: like the reference, `groups` is not forwarded to the residual blocks (they always use 32).
            self.res_layers.append(ResidualBlock(in_channels=hidden_channels, out_channels=hidden_channels, emb_dim=emb_dim))
            self.attn_layers.append(AttentionBlock(channels=hidden_channels, heads=num_heads, d_k=dim_per_head, groups=groups)
                                    if use_attn else nn.Identity())
        if block_type == UNetBlockType.DOWN:
            self.out_layer = DownsampleBlock(in_channels=hidden_channels, out_channels=out_channels)
        elif block_type == UNetBlockType.UP:
            self.out_layer = UpsampleBlock(in_channels=hidden_channels, out_channels=out_channels)

    def forward(self, x, emb=None):
        return _standalone(self, "unet_block", x, emb)
