"""Algorithmic FLOP count of one U-Net forward (2 x MAC of conv / conv-transpose / linear / QK^T / PV only), the
basis of every tensor-roofline fraction.  `tensor_core_only` drops the tiny CUDA-core linears
(embedding MLPs, AdaGN scale vectors) so the figure matches exactly what the tcgen05 kernel executes."""


def unet_forward_flops(net, height, width, batch=1, tensor_core_only=False):
    from models.custom_layers import AttentionBlock, UpsampleBlock

    total = 0.0

    def conv(cin, cout, h, w, k=9):
        return 2.0 * h * w * cout * cin * k

    def res_stack(blk, h, w):
        f = 0.0
        for res, attn in zip(blk.res_layers, blk.attn_layers):
            for cb in (res.conv_block_1, res.conv_block_2):
                wt = cb.conv_layer[0].weight
                f += conv(wt.shape[1], wt.shape[0], h, w)
                if cb.adagn is not None and not tensor_core_only:
                    f += 2.0 * cb.adagn.y_scale.weight.numel()
            if isinstance(attn, AttentionBlock):
                p, c, hd = h * w, attn.projection.weight.shape[1], attn.heads * attn.d_k
                f += 2.0 * p * c * 3 * hd + 2.0 * p * hd * c + 4.0 * attn.heads * p * p * attn.d_k
        return f

    h, w = height, width
    for cb in net.in_layer:
        wt = cb.conv_layer[0].weight
        total += conv(wt.shape[1], wt.shape[0], h, w)
    for blk in net.down_layers:
        total += res_stack(blk, h, w)
        wt = blk.out_layer.conv_layer[0].weight
        h, w = h // 2, w // 2
        total += conv(wt.shape[1], wt.shape[0], h, w)
    for cb in net.middle_layer:
        wt = cb.conv_layer[0].weight
        total += conv(wt.shape[1], wt.shape[0], h, w)
    for blk in net.up_layers:
        total += res_stack(blk, h, w)
        wt = blk.out_layer.conv_layer[0].weight          # [Cin][Cout][4][4]
        assert isinstance(blk.out_layer, UpsampleBlock)
        total += 2.0 * h * w * wt.shape[0] * wt.shape[1] * 16
        h, w = h * 2, w * 2
    for cb in net.out_layers:
        if hasattr(cb, "conv_layer"):
            wt = cb.conv_layer[0].weight
            total += conv(wt.shape[1], wt.shape[0], h, w)
    if net.cond_emb is not None and not tensor_core_only:
        for seq in (net.cond_emb.time_layer, net.cond_emb.cond_layer):
            if seq is not None:
                total += sum(2.0 * m.weight.numel() for m in seq if hasattr(m, "weight"))
    return total * batch
