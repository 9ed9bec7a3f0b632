"""Every class of models/custom_layers.py runs on its own (the reference exposes them as ordinary nn.Modules):
standalone forward of each block against the oracle's restatement on identical parameters."""
import pytest
import torch

from conftest import rel_l2
from oracle import diffusion_oracle as orc

pytestmark = pytest.mark.gpu
TOL = 1e-2          # bf16 mode


def _sd(mod, prefix="m"):
    return {f"{prefix}.{k}": v.detach().cpu().clone() for k, v in mod.state_dict().items()}


@pytest.fixture(autouse=True)
def _seed():
    torch.manual_seed(11)


def test_conv_and_residual_blocks():
    from models.custom_layers import ResidualBlock, UNet_ConvBlock
    x = torch.randn((2, 128, 8, 8))
    emb = torch.randn((2, 64))
    cb = UNet_ConvBlock(128, 128, emb_dim=64).cuda()
    assert rel_l2(cb(x.cuda(), emb.cuda()).cpu(), orc.conv_block(_sd(cb), "m", x, emb)) < TOL
    plain = UNet_ConvBlock(128, 64, use_activation=False).cuda()
    assert rel_l2(plain(x.cuda()).cpu(), orc.conv_block(_sd(plain), "m", x, None, act=False)) < TOL
    rb = ResidualBlock(128, 128, emb_dim=64).cuda()
    assert rel_l2(rb(x.cuda(), emb.cuda()).cpu(), orc.residual_block(_sd(rb), "m", x, emb)) < TOL


def test_adagn_and_embedding_and_swish():
    from models.custom_layers import AdaGN, ConditionalEmbedding, Swish
    x = torch.randn((3, 64, 4, 4))
    emb = torch.randn((3, 32))
    ag = AdaGN(32, 64).cuda()
    assert rel_l2(ag(x.cuda(), emb.cuda()).cpu(), orc.adagn(_sd(ag), "m", x, emb)) < TOL
    ce = ConditionalEmbedding(64, cond_dim=5).cuda()
    t = torch.tensor([3, 500, 999])
    cond = torch.rand((3, 5))
    sd = {k.replace("m.", "cond_emb."): v for k, v in _sd(ce).items()}
    assert rel_l2(ce(t.cuda(), cond.cuda()).cpu(), orc.cond_embedding(sd, t, cond)) < 1e-4
    assert rel_l2(Swish()(x.cuda()).cpu(), orc.swish(x)) < 1e-5


def test_attention_and_samplers_and_unet_block():
    from models.custom_layers import AttentionBlock, DownsampleBlock, UNetBlock, UNetBlockType, UpsampleBlock
    import torch.nn.functional as F
    x = torch.randn((2, 128, 8, 8))
    emb = torch.randn((2, 64))
    at = AttentionBlock(128, heads=2, d_k=64).cuda()
    assert rel_l2(at(x.cuda()).cpu(), orc.attention_block(_sd(at), "m", x, 2)) < TOL
    up = UpsampleBlock(128, 64).cuda()
    ref = F.conv_transpose2d(x, up.conv_layer[0].weight.cpu(), up.conv_layer[0].bias.cpu(), stride=2, padding=1)
    assert rel_l2(up(x.cuda()).cpu(), orc.swish(ref)) < TOL
    dn = DownsampleBlock(128, 256).cuda()
    ref = F.conv2d(x, dn.conv_layer[0].weight.cpu(), dn.conv_layer[0].bias.cpu(), stride=2, padding=1)
    assert rel_l2(dn(x.cuda()).cpu(), orc.swish(ref)) < TOL
    for kind, cout in ((UNetBlockType.DOWN, 256), (UNetBlockType.UP, 64)):
        ub = UNetBlock(128, cout, emb_dim=64, num_resnet_blocks=2, use_attn=True, num_heads=1, block_type=kind).cuda()
        want = orc.unet_block(_sd(ub), "m", x, emb, 1, kind == UNetBlockType.UP)
        assert rel_l2(ub(x.cuda(), emb.cuda()).cpu(), want) < 2 * TOL


def test_standalone_blocks_refuse_cpu_tensors():
    from b200._lib import B200Error
    from models.custom_layers import UNet_ConvBlock
    with pytest.raises(B200Error):
        UNet_ConvBlock(128, 128)(torch.randn((1, 128, 4, 4)))
