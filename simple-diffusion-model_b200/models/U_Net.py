"""U_Net with the reference's constructor, parameter tree (state_dict keys/shapes) and call signature
(reference models/U_Net.py:11-25,132,147), executed by the sm_100a engine in b200/engine.py."""
import os

import torch
import torch.nn as nn

from .custom_layers import *  # noqa: F401,F403  (the reference exports the blocks from here too)
from .custom_layers import ConditionalEmbedding, UNet_ConvBlock, UNetBlock, UNetBlockType


class U_Net(nn.Module):
    """Parameter tree of the reference U-Net; none of the sub-modules' `forward`s run when the net is called.

    `forward` hands (x, t, cond) to one engine object (built lazily, owns the activation arena, weight cache and flat
    parameter / gradient buffers):
      * grad disabled  -> `Engine.forward`: NCHW fp32 in -> NHWC bf16 (or fp32/TF32) activations -> ~1.2 k launches of the
        C-ABI kernels, or one CUDA-graph replay after `cuda_graphs(True)`;
      * grad enabled   -> `UNetTrainEngine.forward_train`: same kernels with the activations the hand-written backward
        needs kept on a tape; the returned tensor's autograd node runs that backward and fills `param.grad` (views into one
        flat fp32 buffer, 3x3 weights stored channels-last behind a permuted view).
    The skip connections concatenate along channels, which in NHWC is two column ranges of one buffer: the down path's
    epilogues write straight into the right half of the up path's input, no cat kernel.
    """

    def __init__(self, num_resnet_blocks=5, in_channel=3, out_channel=3, time_dim=64, cond_dim=None, num_layers=5,
                 attn_layers=[2, 3, 4], num_heads=1, dim_per_head=None, groups=32, min_channel=128, max_channel=512,
                 image_recon=False):
        super().__init__()
        if not isinstance(num_layers, int) or not isinstance(attn_layers, list):
            raise TypeError("Invalid type!")
        if num_layers < 1:
            raise ValueError("Invalid num layer value!")
        for a in attn_layers:
            if not isinstance(a, int):
                raise ValueError("Invalid type in attention layer!")
            if a < 0 or a >= num_layers:
                raise ValueError("Invalid Attention Layer values!")

        # Channel plan: double per level, capped at max_channel (reference models/U_Net.py:41-46).
        widths, doubled = [min_channel], min_channel
        for _ in range(num_layers):
            doubled *= 2
            widths.append(min(doubled, max_channel))
        self.image_recon = image_recon
        # "bf16" (tensor-core speed path) or "tf32" (fp32 storage + TF32 MMA: the parity mode)
        self.precision = os.environ.get("SDM_B200_PRECISION", "bf16")

        self.cond_emb = ConditionalEmbedding(time_dim, cond_dim) if time_dim is not None else None
        self.in_layer = nn.Sequential(
            UNet_ConvBlock(in_channel, widths[0], use_activation=True, emb_dim=None),
            UNet_ConvBlock(widths[0], widths[0], use_activation=True, emb_dim=None))
        self.down_layers = nn.ModuleList(
            UNetBlock(in_channels=widths[i], out_channels=widths[i + 1], emb_dim=time_dim, num_resnet_blocks=num_resnet_blocks,
                      use_attn=i in attn_layers, num_heads=num_heads, dim_per_head=dim_per_head, groups=groups,
                      block_type=UNetBlockType.DOWN) for i in range(num_layers))
        self.middle_layer = nn.Sequential(
            UNet_ConvBlock(widths[-1], widths[-1], use_activation=True, emb_dim=None),
            UNet_ConvBlock(widths[-1], widths[-1], use_activation=True, emb_dim=None))
        self.up_layers = nn.ModuleList(
            UNetBlock(in_channels=widths[i + 1] * 2, out_channels=widths[i], emb_dim=time_dim,
                      num_resnet_blocks=num_resnet_blocks, use_attn=i in attn_layers, num_heads=num_heads,
                      dim_per_head=dim_per_head, groups=groups, block_type=UNetBlockType.UP)
            for i in range(num_layers - 1, -1, -1))
        tail = [UNet_ConvBlock(widths[0], widths[0], use_activation=True, emb_dim=None),
                UNet_ConvBlock(widths[0], out_channel, use_activation=False, emb_dim=None)]
        if image_recon:
            tail.append(nn.Tanh())      # parameter-free; fused into the last conv's epilogue by the engine
        self.out_layers = nn.Sequential(*tail)
        self._engine = None
        self._graphed = None

    def set_precision(self, precision):
        """'bf16': bf16 activations/weights, fp32 accumulation (default).  'tf32': fp32 storage, TF32 MMA (parity mode)."""
        if precision not in ("bf16", "tf32"):
            raise ValueError("precision must be 'bf16' or 'tf32'")
        self.precision = precision
        return self

    def cuda_graphs(self, enabled=True):
        """Inference-mode forwards replay a captured CUDA graph per input signature (b200/graph.py) instead of issuing the
        ~1.2 k kernel launches one by one: the samplers call the network 51-1000 times with identical shapes."""
        if enabled and self._graphed is None:
            from b200.graph import GraphedUNet
            self._graphed = GraphedUNet(self)
        elif not enabled:
            self._graphed = None
        return self

    def engine(self):
        """The per-net engine (weights packed lazily on first use; rebuilt caches follow `.to()` / `load_state_dict`)."""
        if self._engine is None:
            from b200.train_engine import UNetTrainEngine
            self._engine = UNetTrainEngine(self)
        return self._engine

    def custom_load_state_dict(self, state_dict):
        """Tolerant loader (reference models/U_Net.py:132-145): unknown / mismatched entries are skipped."""
        own = self.state_dict()
        for name, param in state_dict.items():
            if name not in own:
                print(f"No Layer found: {name}, skipping")
                continue
            if own[name].shape != param.data.shape:
                print(f"Skipped: {name}")
                continue
            if isinstance(param, torch.nn.parameter.Parameter):
                param = param.data
            own[name].copy_(param)

    def forward(self, x, t=None, cond=None):
        eng = self.engine()
        if torch.is_grad_enabled() and any(p.requires_grad for p in self.parameters()):
            return eng.forward_train(x, t, cond)
        if self._graphed is not None and x.is_cuda:
            return self._graphed(x, t, cond).clone()        # the graph's output buffer is overwritten by the next replay
        return eng.forward(x, t, cond)
