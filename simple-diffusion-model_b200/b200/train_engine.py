"""Training-mode executor: forward with a tape + hand-written backward for the whole U-Net.

The reference relies on torch.autograd over ~1.2 k ATen ops (train_diffusion.py:343-358).  Here one autograd node wraps
the network; its backward replays a tape of layer records in reverse and launches the sm_100a gradient kernels:
  * data gradients of every convolution re-use the forward implicit-GEMM kernel on re-packed weights,
  * weight gradients run on the tcgen05 TN kernel (contraction over pixels, split-K, fp32 atomics),
  * GroupNorm x AdaGN x Swish backward is two streaming passes, bias gradients ride along in those passes,
  * all parameter gradients land in ONE flat fp32 buffer (`grad_flat`); `param.grad` tensors are views of it, which is
    what the data-parallel all-reduce and the fused Adam consume (no per-parameter gradient tensors, no copies).
Parameters that never receive a gradient in the reference (AdaGN.y_shift.*, AttentionBlock.norm.*) are
excluded statically, so no `find_unused_parameters` machinery is needed.
"""
import torch

from . import ops
from ._lib import B200Error, call, ptr, stream
from .engine import UNetEngine


def _is_dead(name):
    return ".y_shift." in name or (".attn_layers." in name and ".norm." in name)


class GradLayout:
    """Flat fp32 gradient buffer in bucket order: AdaGN scale weights, AdaGN scale biases (each contiguous so the
    batched scale-vector GEMM writes them in place), then every other live parameter in reverse execution order
    (out_layers -> up -> middle -> down -> in_layer -> embedding), the order in which backward completes them.

    3x3 Conv2d weights whose channel counts fill whole 128-byte K blocks are STORED channels-last
    ([Cout][3][3][Cin] == the tensor-core kernels' [Cout][tap][Cin] layout); the nn.Parameter keeps the reference's
    shape [Cout][Cin][3][3] as a permuted view, so state_dicts interchange.  With that storage
      * the weight-gradient GEMM writes straight into the flat gradient buffer (no scratch, no unpack pass),
      * the forward weights are a plain bf16 cast of the master copy, which the fused Adam emits in the same pass
        (`shadow`), so no per-step weight packing remains on the forward path."""

    def __init__(self, net, device):
        import torch.nn as nn
        from models.custom_layers import AdaGN
        self.cl = {}          # id(param) -> (Cout, Cin) for channels-last stored 3x3 conv weights
        for m in net.modules():
            if isinstance(m, nn.Conv2d) and tuple(m.kernel_size) == (3, 3) and m.weight.shape[1] % 64 == 0:
                self.cl[id(m.weight)] = (m.weight.shape[0], m.weight.shape[1])
        named = [(n, p) for n, p in net.named_parameters() if not _is_dead(n) and p.requires_grad]
        adagn = [m for m in net.modules() if isinstance(m, AdaGN)]
        ys_w = [m.y_scale.weight for m in adagn]
        ys_b = [m.y_scale.bias for m in adagn]
        special = {id(p) for p in ys_w + ys_b}

        def order_key(item):
            n = item[0]
            if n.startswith("out_layers"):
                return (0, 0)
            if n.startswith("up_layers"):
                return (1, -int(n.split(".")[1]))
            if n.startswith("middle_layer"):
                return (2, 0)
            if n.startswith("down_layers"):
                return (3, -int(n.split(".")[1]))
            if n.startswith("in_layer"):
                return (4, 0)
            return (5, 0)

        rest = sorted([it for it in named if id(it[1]) not in special], key=order_key)
        self.params = ys_w + ys_b + [p for _, p in rest]
        self.offsets, off = {}, 0
        for p in self.params:
            self.offsets[id(p)] = off
            off += self._span(p)
        self.total = off
        self.front_end = sum(self._span(p) for p in ys_w + ys_b)
        self.flat = torch.zeros(self.total, dtype=torch.float32, device=device)
        self.views = {id(p): self._shaped(self.flat, p) for p in self.params}
        self.params_flat = None
        self.shadow = None                     # bf16 copy of params_flat (same offsets), maintained by FusedAdam
        self.shadow_key = {}                   # id(p) -> (p._version, epoch) for which the shadow slice is current
        self.shadow_bulk_epoch = -1            # epoch at which a fused optimiser pass rewrote the whole shadow
        self.epoch = 0                         # bumped by FusedAdam.step(): invalidates cached kernel-layout weights
        for p in self.params:
            p._b2_layout = self
        self.adagn_w_numel = sum(p.numel() for p in ys_w)
        self.adagn_b_numel = sum(p.numel() for p in ys_b)
        # the concatenated y_scale views are dense when no tensor needs padding up to the next ALIGN boundary
        self.dense_adagn = all(p.numel() % self.ALIGN == 0 for p in ys_w + ys_b)

    # Every tensor starts on a 128-element boundary: 256 bytes in the bf16 copy the tensor-core kernels read through TMA, 512 bytes
    # in the fp32 buffers.  With the 8-element (16-byte) granularity of round 1 everything behind the 3-element bias of the last
    # conv sat 16 bytes off: each 128-byte TMA row of a weight tile then straddled two L2 lines.  MEASURED (round 2, 128x128 batch
    # 32): data gradients reading the forward weights in place were 3.5 ms per step slower than through freshly allocated
    # (aligned) transposed copies although the kernels are equally fast on aligned buffers.
    ALIGN = 128

    @classmethod
    def _span(cls, p):
        return (p.numel() + cls.ALIGN - 1) // cls.ALIGN * cls.ALIGN

    def _shaped(self, flat, p):
        """The slice of `flat` that belongs to p, shaped like p (a permuted view for channels-last stored weights)."""
        off = self.offsets[id(p)]
        raw = flat[off:off + p.numel()]
        if id(p) in self.cl:
            cout, cin = self.cl[id(p)]
            return raw.view(cout, 3, 3, cin).permute(0, 3, 1, 2)
        return raw.view(p.shape)

    def view(self, p):
        return self.views[id(p)]

    def raw_grad(self, p):
        """Gradient storage of p as the flat slice (kernel layout for channels-last weights)."""
        off = self.offsets[id(p)]
        return self.flat[off:off + p.numel()]

    def is_cl(self, p):
        return id(p) in self.cl

    def shadow_slice(self, p):
        """Current bf16 copy of p in storage layout, or None when no fused optimiser maintains one.  A slice that went
        stale (load_state_dict, a non-fused optimiser) is refreshed from the fp32 master with one cast."""
        if self.shadow is None or self.params_flat is None:
            return None
        off = self.offsets[id(p)]
        sl = self.shadow[off:off + p.numel()]
        key = self.shadow_key.get(id(p))
        fresh = key is not None and key[0] == p._version and (key[1] == self.epoch or self.shadow_bulk_epoch == self.epoch)
        if not fresh:
            sl.copy_(self.params_flat[off:off + p.numel()])
            self.shadow_key[id(p)] = (p._version, self.epoch)
        return sl

    def stepped(self, shadow_written):
        """A fused optimiser pass changed every parameter (bypassing torch's version counters)."""
        self.epoch += 1
        if shadow_written:
            self.shadow_bulk_epoch = self.epoch
            if not self.shadow_key:
                self.shadow_key = {id(p): (p._version, self.epoch) for p in self.params}

    def ensure_shadow(self):
        if self.shadow is None and self.params_flat is not None:
            self.shadow = torch.empty(self.total, dtype=torch.bfloat16, device=self.flat.device)
            self.shadow.copy_(self.params_flat)
            self.shadow_key = {id(p): (p._version, self.epoch) for p in self.params}
        return self.shadow

    def module_range(self, module):
        """Flat range covered by a module's live parameters, AdaGN scale Linears excluded (they live in the front
        region).  Contiguous by construction of the bucket order."""
        offs = [(self.offsets[id(p)], self.offsets[id(p)] + self._span(p)) for p in module.parameters()
                if id(p) in self.offsets and self.offsets[id(p)] >= self.front_end]
        if not offs:
            return 0, 0
        return min(o[0] for o in offs), max(o[1] for o in offs)

    def flatten_params(self):
        """Re-homes every live parameter into one flat fp32 buffer laid out like the gradients, so the optimiser
        (and a parameter broadcast) is a single pass.  Parameter objects, names and shapes are unchanged."""
        if self.params_flat is None:
            flat = torch.zeros(self.total, dtype=torch.float32, device=self.flat.device)
            self.pviews = {}
            for p in self.params:
                v = self._shaped(flat, p)
                v.copy_(p.data)
                p.data = v
                self.pviews[id(p)] = v
            self.params_flat = flat
        return self.params_flat

    def param_view(self, p):
        return self.pviews[id(p)]


class _UNetFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, engine, x, t, cond, *params):
        out, tape = engine._forward_tape(x, t, cond)
        ctx.engine, ctx.tape = engine, tape
        ctx.n_params = len(params)
        return out

    @staticmethod
    def backward(ctx, dout):
        ctx.engine._backward_tape(ctx.tape, dout)
        ctx.tape = None
        return (None, None, None, None) + (None,) * ctx.n_params      # grads are written into param.grad views directly


class UNetTrainEngine(UNetEngine):
    def __init__(self, net):
        super().__init__(net)
        self.layout = None
        # Weight gradients do not feed the backward chain, so they CAN run on a second stream next to the data-gradient /
        # normalisation kernels of the following layers (fork/join with stream events, kept by CUDA-graph capture).
        # Measured on B200 (tools/bench_train.py, batch 32 @64x64 and batch 16 @128x128): within +-1 % of the single-stream
        # schedule -- both GEMM kernels are persistent 1-CTA-per-SM grids with ~200 KB of shared memory, so two of them
        # cannot share an SM -- hence off by default (SDM_B200_OVERLAP_WGRAD=1 enables it).
        import os
        self.overlap_wgrad = os.environ.get("SDM_B200_OVERLAP_WGRAD", "0") == "1"
        self._side_stream = None
        self._side_busy = False
        self._side_keep = []               # operands of in-flight side-stream kernels (kept alive until the join)
        # SDM_B200_FUSE_ADAGN_SUMS=1: the data-gradient GEMM that produces an AdaGN layer's `dout` also emits that layer's pass-1
        # sums from its epilogue (b2_conv2d_nhwc_colsum), so the backward's reduce pass -- a full read of dout and z -- disappears
        # for every layer whose gradient comes straight from a stride-1 conv (90 of 100).  Implemented, parity-tested and
        # MEASURED SLOWER (profiles/r02x_fused_adagn_sums_ab.log, 128x128 batch 32: 90.6 ms two-pass, 92.8 ms fused everywhere,
        # 92.1 / 91.2 ms fused only for >= 512 / 1024 channels): the extra 62 shuffles + 64 atomics + Swish per 32x32 chunk
        # lengthen the GEMM epilogues by more than the 2 ms of memory-bound reduce passes they remove.  Hence OFF by default.
        self.fuse_adagn_sums = os.environ.get("SDM_B200_FUSE_ADAGN_SUMS", "0") == "1"
        self.fuse_adagn_min_c = int(os.environ.get("SDM_B200_FUSE_ADAGN_MIN_C", "0"))
        # data gradients of the 3x3 stride-1 convs read the forward weights MN-major (SDM_B200_DGRAD_FROM_FWD=0: transposed copies)
        self.dgrad_from_fwd = os.environ.get("SDM_B200_DGRAD_FROM_FWD", "1") == "1"
        # SDM_B200_GROUP_WGRAD=1 (auto for small workloads under GraphedTrainStep): weight gradients of the channels-last stored convs
        # are deferred and run as ONE grouped launch per module (b2_conv2d_wgrad_batch) -- they feed nothing but the optimiser,
        # and at small batches the step is bound by the number of ~16 us dependent launches
        self.group_wgrad = os.environ.get("SDM_B200_GROUP_WGRAD", "0") == "1"
        self._wgrad_jobs = []
        self.attn_bmn = os.environ.get("SDM_B200_ATTN_BMN", "1") == "1"      # attention backward reads dO / Q in place (see _bwd_attention)
        # un-normalised convs store z and Swish(z) from one epilogue (b2_conv2d_nhwc_dual); SDM_B200_FUSE_FWD_ACT=0: separate pass
        self.fuse_fwd_act = os.environ.get("SDM_B200_FUSE_FWD_ACT", "1") == "1"
        self.post_backward = None          # optional callable(layout), runs when every gradient is complete
        self.on_grads_ready = None         # optional callable(layout, lo, hi): flat range [lo, hi) is final (DP buckets)

    # ------------------------------------------------------------------------------------------ public
    def forward_train(self, x, t=None, cond=None):
        params = [p for p in self.net.parameters() if p.requires_grad]
        return _UNetFn.apply(self, x, t, cond, *params)

    def grad_layout(self, device):
        if self.layout is None or self.layout.flat.device != device:
            self.layout = GradLayout(self.net, device)
        return self.layout

    # ------------------------------------------------------------------------------------------ forward with tape
    @torch.no_grad()
    def _forward_tape(self, x, t, cond):
        net = self.net
        if not x.is_cuda:
            raise B200Error("U_Net.forward needs CUDA tensors: this build has no CPU path")
        code = self._code()
        n, cin, hgt, wid = x.shape
        levels = len(net.down_layers)
        if hgt % (1 << levels) or wid % (1 << levels):
            raise B200Error(f"H and W must be divisible by 2**num_layers = {1 << levels}")
        dev = x.device
        self.cache.refresh_all()         # every kernel-layout weight the optimiser invalidated, in two launches
        tape = []
        ctx = {"emb": None, "stats_i": 0, "tape": tape, "n": n}
        if net.cond_emb is None:
            raise B200Error("training needs the timestep embedding (time_dim is not None)")
        if net.cond_emb is not None:
            if t is None:
                raise B200Error("timestep tensor `t` is required")
            emb, emb_rec = self._embedding_train(t.to(dev), cond.to(dev) if cond is not None else None)
            w_all, b_all, off, total = self._adagn_table()
            be = emb.shape[0]
            if be not in (1, n):
                raise B200Error(f"embedding batch {be} does not broadcast over image batch {n}")
            s_all = torch.empty((be, total), dtype=torch.float32, device=dev)
            ops.small_gemm(emb, w_all, be, total, emb.shape[1], emb.shape[1], w_all.shape[1], s_all, total, bias=b_all)
            mods = self._adagn_modules()
            max_groups = max(m.group_norm.num_groups for m in mods)
            ctx.update(emb=emb, s_all=s_all, adagn_off=off, s_bstride=(total if be == n else 0), be=be, total=total,
                       stats=torch.zeros((len(off), n, max_groups, 2), dtype=torch.float32, device=dev),
                       ds_all=torch.zeros((be, total), dtype=torch.float32, device=dev))
            tape.append(("emb", emb_rec, ctx))
        kal = ops.K_ALIGN[code]
        cpad = ((cin + kal - 1) // kal) * kal
        h = ops.nchw_to_nhwc_pad(x, cpad, code)
        tape.append(("mark", net.in_layer))
        first = net.in_layer[0]
        if ops.edge_first_ok(first.conv_layer[0]) and wid % 2 == 0:
            conv0 = first.conv_layer[0]           # forward on the CUDA-core edge kernel; the padded NHWC copy only feeds the wgrad
            c0 = conv0.weight.shape[0]
            z0 = ops.conv_first(x.contiguous().float(), self.cache.get_edge(conv0.weight, "first"), conv0.bias, c0, 0, code)
            a0 = ops.new_act(n, hgt, wid, c0, code, dev)
            ops.act(0, None, z0, a0, None, n * hgt * wid, c0, 0, z0.stride(2), a0.stride(2), code)
            tape.append(("plain", first, h, z0, False))
            h = a0
        else:
            h = self._plain_conv_train(first, h, ctx, need_dx=False)
        h = self._plain_conv_train(net.in_layer[1], h, ctx)
        cats = []
        hh, ww = hgt, wid
        for blk in net.down_layers:
            cout = blk.out_layer.conv_layer[0].weight.shape[0]
            hh, ww = hh // 2, ww // 2
            cat = ops.new_act(n, hh, ww, 2 * cout, code, dev)
            tape.append(("mark", blk))
            h = self._block_train(blk, h, ctx, out=cat[..., cout:])
            tape.append(("skip_out", len(cats)))             # its gradient also arrives through the concat buffer
            cats.append(cat)
        tape.append(("mark", net.middle_layer))
        h = self._plain_conv_train(net.middle_layer[0], h, ctx)
        c_mid = cats[-1].shape[3] // 2
        self._plain_conv_train(net.middle_layer[1], h, ctx, out=cats[-1][..., :c_mid])
        n_up = len(net.up_layers)
        for i, blk in enumerate(net.up_layers):
            level = len(cats) - 1
            cat = cats.pop()
            cout = blk.out_layer.conv_layer[0].weight.shape[1]
            dst = cats[-1][..., :cout] if i + 1 < n_up else None
            tape.append(("mark", blk))
            tape.append(("cat_in", level, cat.shape[3] // 2))   # splits d(cat) into d(x part) and d(skip part)
            h = self._block_train(blk, cat, ctx, out=dst)
        tape.append(("mark", net.out_layers))
        h = self._plain_conv_train(net.out_layers[0], h, ctx)
        last = net.out_layers[1]
        conv = last.conv_layer[0]
        c_out = conv.weight.shape[0]
        y = torch.empty((n, c_out, hgt, wid), dtype=torch.float32, device=dev)
        if ops.edge_last_ok(conv) and wid % 4 == 0:
            ops.conv_last(h, self.cache.get_edge(conv.weight, "last"), conv.bias, c_out, 2 if net.image_recon else 0, y)
        else:
            w = self.cache.get(conv.weight, 0, code, c_out, conv.weight.shape[1], h.shape[3])
            ops.conv2d(0, h, w, conv.bias, c_out, act=(2 if net.image_recon else 0), out_nchw_fp32=y)
        tape.append(("last", last, h, y if net.image_recon else None))
        del ctx["tape"]          # break the ctx <-> tape reference cycle: activations must die by refcount, not by GC
        return y, tape

    def _embedding_train(self, t, cond):
        ce = self.net.cond_emb
        dim = ce.time_dim
        t = t.to(torch.int64).contiguous()
        bt = t.shape[0]
        sin = torch.empty((bt, dim), dtype=torch.float32, device=t.device)
        call("b2_sinusoid_embedding", ptr(t), ptr(sin), bt, dim, stream())

        def mlp(seq, x0, b, d_in):
            lins = [seq[0], seq[2], seq[4], seq[6]]
            acts, pres, h, k = [x0], [], x0, d_in
            for i, lin in enumerate(lins):
                nn_ = lin.weight.shape[0]
                pre = torch.empty((b, nn_), dtype=torch.float32, device=x0.device)
                ops.small_gemm(h, lin.weight, b, nn_, k, k, lin.weight.shape[1], pre, nn_, bias=lin.bias)
                if i < 3:
                    hn = torch.empty_like(pre)
                    call("b2_f32_act", 0, None, ptr(pre), ptr(hn), pre.numel(), stream())
                    pres.append(pre)
                    acts.append(hn)
                    h = hn
                else:
                    h = pre
                k = nn_
            return h, (lins, acts, pres)

        emb, rec_t = mlp(ce.time_layer, sin, bt, dim)
        rec_c = None
        if ce.cond_layer is not None:
            if cond is None:
                raise B200Error("this U_Net was built with cond_dim: `cond` is required")
            c2 = cond.float().reshape(-1, cond.shape[-1]).contiguous()
            cemb, rec_c = mlp(ce.cond_layer, c2, c2.shape[0], c2.shape[1])
            emb = emb + cemb if cemb.shape[0] != emb.shape[0] else emb.add_(cemb)
        return emb, (rec_t, rec_c, bt)

    def _plain_conv_train(self, blk, x, ctx, out=None, need_dx=True):
        """Conv + bias + Swish without normalisation (in/middle/out layers): keeps the pre-activation for backward."""
        conv = blk.conv_layer[0]
        code = ops.code_of(x)
        cout, cin = conv.weight.shape[0], conv.weight.shape[1]
        w = self.cache.get(conv.weight, 0, code, cout, cin, x.shape[3])
        n, hh, ww, _ = x.shape
        if out is None:
            out = ops.new_act(n, hh, ww, cout, code, x.device)
        if self.fuse_fwd_act:
            z = ops.conv2d_dual(0, x, w, conv.bias, cout, out)          # z and Swish(z) from one epilogue
        else:
            z = ops.conv2d(0, x, w, conv.bias, cout, act=0)
            ops.act(0, None, z, out, None, n * hh * ww, cout, 0, z.stride(2), out.stride(2), code)
        ctx["tape"].append(("plain", blk, x, z, need_dx))
        return out

    def _gn_conv_train(self, blk, x, ctx, out=None, residual=None):
        conv = blk.conv_layer[0]
        code = ops.code_of(x)
        cout, cin = conv.weight.shape[0], conv.weight.shape[1]
        w = self.cache.get(conv.weight, 0, code, cout, cin, x.shape[3])
        gn = blk.adagn.group_norm
        idx = ctx["stats_i"]
        ctx["stats_i"] += 1
        stats = ctx["stats"][idx]
        z = ops.conv2d(0, x, w, conv.bias, cout, act=3, gn_stats=stats, groups=gn.num_groups)
        off = ctx["adagn_off"][id(blk.adagn)]
        s = ctx["s_all"][:, off:off + cout]
        y = ops.adagn_apply(z, stats, gn.weight, gn.bias, s, ctx["s_bstride"], out=out, residual=residual,
                            groups=gn.num_groups, eps=gn.eps, pre_swish=True)
        return y, (blk, x, z, stats, off)

    def _block_train(self, blk, x, ctx, out):
        from models.custom_layers import AttentionBlock, UpsampleBlock
        tape = ctx["tape"]
        for res, attn in zip(blk.res_layers, blk.attn_layers):
            a1, rec1 = self._gn_conv_train(res.conv_block_1, x, ctx)
            y, rec2 = self._gn_conv_train(res.conv_block_2, a1, ctx, residual=x)
            tape.append(("res", rec1, rec2))
            x = y
            if isinstance(attn, AttentionBlock):
                x = self._attention_train(attn, x, tape)
        code = ops.code_of(x)
        conv = blk.out_layer.conv_layer[0]
        n, hh, ww, _ = x.shape
        if isinstance(blk.out_layer, UpsampleBlock):
            cin, cout = conv.weight.shape[0], conv.weight.shape[1]
            w = self.cache.get(conv.weight, 2, code, cout, cin, cin)
            if out is None:
                out = ops.new_act(n, 2 * hh, 2 * ww, cout, code, x.device)
            if self.fuse_fwd_act:
                z = ops.conv2d_dual(2, x, w, conv.bias, cout, out)
            else:
                z = ops.conv2d(2, x, w, conv.bias, cout, act=0)
                ops.act(0, None, z, out, None, n * 4 * hh * ww, cout, 0, z.stride(2), out.stride(2), code)
            tape.append(("up", blk.out_layer, x, z))
            return out
        cout, cin = conv.weight.shape[0], conv.weight.shape[1]
        w = self.cache.get(conv.weight, 0, code, cout, cin, cin)
        planes = ops.space_to_depth2(x)
        if self.fuse_fwd_act:
            z = ops.conv2d_dual(1, planes, w, conv.bias, cout, out)
        else:
            z = ops.conv2d(1, planes, w, conv.bias, cout, act=0)
            ops.act(0, None, z, out, None, n * (hh // 2) * (ww // 2), cout, 0, z.stride(2), out.stride(2), code)
        tape.append(("down", blk.out_layer, planes, z, (n, hh, ww, cin)))
        return out

    def _attention_train(self, blk, x, tape):
        saved = {}
        out = self.attention(blk, x, save=saved)
        tape.append(("attn", blk, x, saved))
        return out

    # ------------------------------------------------------------------------------------------ backward
    def _scratch(self, numel, device):
        """Persistent, zeroed fp32 scratch for kernel-layout weight gradients (stream order makes one buffer enough;
        re-using it keeps the caching allocator from churning multi-GB blocks every step)."""
        buf = getattr(self, "_scratch_buf", None)
        if buf is None or buf.numel() < numel or buf.device != device:
            buf = torch.empty((max(numel, 1 << 24),), dtype=torch.float32, device=device)
            self._scratch_buf = buf
        out = buf[:numel]
        out.zero_()
        return out

    def _wgrad_conv(self, mode, conv, x, dz, kind):
        """Weight gradient in kernel layout -> unpacked into the flat gradient view of the parameter."""
        code = ops.code_of(x)
        if mode == 2:
            cin, cout = conv.weight.shape[0], conv.weight.shape[1]
            packed = self._scratch(16 * cout * cin, x.device)
            ops.conv2d_wgrad(2, x, dz, cout, packed)
            call("b2_unpack_weight_grad", 2, ptr(packed), ptr(self.layout.view(conv.weight)), cout, cin, cin, 0, stream())
            return
        cout, cin = conv.weight.shape[0], conv.weight.shape[1]
        cin_pad = x.shape[3]
        if self.layout.is_cl(conv.weight) and cin_pad == cin:
            # channels-last stored weight: the flat gradient slice IS the kernel layout (zeroed at the start of backward)
            grad = self.layout.raw_grad(conv.weight)
            if self.group_wgrad:
                self._wgrad_jobs.append((mode, x, dz, cout, grad))       # launched by _flush_wgrad (keeps x and dz alive until then)
                return
            self._launch_wgrads([(mode, x, dz, cout, grad)])
            return
        packed = self._scratch(cout * 9 * cin_pad, x.device)
        ops.conv2d_wgrad(mode, x, dz, cout, packed)
        call("b2_unpack_weight_grad", 0, ptr(packed), ptr(self.layout.view(conv.weight)), cout, cin, cin_pad, 0, stream())

    def _launch_wgrads(self, jobs):
        """One weight gradient directly, several as a grouped launch; on the side stream when weight gradients overlap."""
        def run():
            if len(jobs) == 1:
                ops.conv2d_wgrad(*jobs[0])
            else:
                ops.conv2d_wgrad_batch(jobs)
        if not self.overlap_wgrad:
            run()
            return
        dev = jobs[0][1].device
        main = torch.cuda.current_stream(dev)
        if self._side_stream is None or self._side_stream.device != dev:
            self._side_stream = torch.cuda.Stream(device=dev)
        side = self._side_stream
        side.wait_stream(main)                      # dz (and the zeroed gradient buffer) are ready
        with torch.cuda.stream(side):
            run()
        self._side_busy = True
        self._side_keep.append(jobs)

    def _flush_wgrad(self):
        if self._wgrad_jobs:
            jobs, self._wgrad_jobs = self._wgrad_jobs, []
            self._launch_wgrads(jobs)

    def _join_side(self):
        """Main stream waits for the weight-gradient stream (before gradients are consumed: all-reduce, optimiser)."""
        if self._side_busy:
            torch.cuda.current_stream(self._side_stream.device).wait_stream(self._side_stream)
            self._side_busy = False
        self._side_keep.clear()

    def _bwd_work(self, numel, device):
        """Zeroed fp32 scratch for one AdaGN backward, carved from an arena that is cleared ONCE per backward pass."""
        arena = getattr(self, "_work_arena", None)
        if arena is None or self._work_used + numel > arena.numel():
            arena = torch.zeros((max(numel, 1 << 20),), dtype=torch.float32, device=device)     # only for standalone block runs
            self._work_arena, self._work_used = arena, 0
        out = arena[self._work_used:self._work_used + numel]
        self._work_used += (numel + 3) // 4 * 4
        return out

    def _dgrad_s1(self, conv, dz, residual=None, out=None, colsum=None):
        code = ops.code_of(dz)
        cout, cin = conv.weight.shape[0], conv.weight.shape[1]
        lay = self.layout
        if self.dgrad_from_fwd and code == ops.BF16 and lay is not None and lay.shadow is not None and lay.is_cl(conv.weight) \
                and dz.shape[3] == cout and cout % 64 == 0 and cin % 64 == 0:
            # the optimiser's bf16 copy [Cout][3][3][Cin] is consumed MN-major with mirrored taps (b2_conv2d_nhwc mode 5): no
            # transposed weight copy, no per-step transpose launch
            return ops.conv2d(5, dz, lay.shadow_slice(conv.weight), None, cin, act=0, residual=residual, out=out, colsum=colsum)
        w = self.cache.get(conv.weight, 1, code, cout, cin, dz.shape[3])
        return ops.conv2d(0, dz, w, None, cin, act=0, residual=residual, out=out, colsum=colsum)

    def _colsum_ok(self, z):
        """Can the GEMM that produces this layer's `dout` also emit its pass-1 sums?  bf16, whole 32-channel chunks, and every
        warp of 32 tile rows inside one image (the epilogue reduces over a warp)."""
        n, hh, ww, c = z.shape
        return self.fuse_adagn_sums and z.dtype == torch.bfloat16 and c % 32 == 0 and hh * ww >= 32 and (hh * ww) % 32 == 0 \
            and c >= self.fuse_adagn_min_c

    def _bwd_gn_conv(self, rec, dout, ctx, residual=None, need_dx=True, sums=None, next_sums=None):
        """sums: this layer's [2][N][C] pass-1 sums, already accumulated by the GEMM that produced `dout` (raw form).
        next_sums: (z, work) of the layer that will consume the data gradient produced here."""
        blk, x, z, stats, off = rec
        conv, gn = blk.conv_layer[0], blk.adagn.group_norm
        code = ops.code_of(z)
        n, hh, ww, c = z.shape
        lay = self.layout
        work = sums if sums is not None else self._bwd_work(2 * n * c, z.device)
        dz = torch.empty((n, hh, ww, c), dtype=z.dtype, device=z.device)
        s = ctx["s_all"][:, off:off + c]
        ds = ctx["ds_all"][:, off:off + c]
        call("b2_adagn_bwd_fused", ptr(dout), dout.stride(2), ptr(z), z.stride(2), ptr(stats), ptr(gn.weight), ptr(gn.bias), ptr(s),
             ctx["s_bstride"], ptr(work), ptr(ds), ctx["total"] if ctx["be"] == n else 0, ptr(lay.view(gn.weight)),
             ptr(lay.view(gn.bias)), ptr(dz), dz.stride(2), ptr(lay.view(conv.bias)), n, hh * ww, c, gn.num_groups,
             float(gn.eps), 1 if sums is not None else 0, code, stream())
        self._wgrad_conv(0, conv, x, dz, 0)
        if not need_dx:
            return None
        colsum = None
        if next_sums is not None:
            zn, wn = next_sums
            nc = zn.shape[0] * zn.shape[3]
            colsum = (zn, wn[:nc], wn[nc:2 * nc])
        return self._dgrad_s1(conv, dz, residual=residual, colsum=colsum)

    def _bwd_act(self, dy, z, dbias_view):
        code = ops.code_of(z)
        n, hh, ww, c = z.shape
        dz = torch.empty((n, hh, ww, c), dtype=z.dtype, device=z.device)
        ops.act(1, dy, z, dz, dbias_view, n * hh * ww, c, dy.stride(2), z.stride(2), c, code)
        return dz

    @torch.no_grad()
    def _backward_tape(self, tape, dout):
        net = self.net
        dev = dout.device
        lay = self.grad_layout(dev)
        call("b2_zero", ptr(lay.flat), lay.flat.numel() * 4, stream())
        need = sum(2 * e[2][2].shape[0] * e[2][2].shape[3] + 2 * e[1][2].shape[0] * e[1][2].shape[3] + 8
                   for e in tape if e[0] == "res")
        self._work_arena, self._work_used = torch.zeros((max(need, 4),), dtype=torch.float32, device=dev), 0
        code = self._code()
        kal = ops.K_ALIGN[code]
        ctx = None
        for entry in tape:
            if entry[0] == "emb":
                ctx = entry[2]
        d = None                     # running gradient w.r.t. the current activation (NHWC, compute dtype)
        d_skips = {}                 # level -> gradient arriving at a skip tensor through the concat buffer
        order = list(reversed(tape))
        pending = None               # pass-1 sums of the NEXT residual block's second AdaGN, filled by the GEMM that produced d
        for pos, entry in enumerate(order):
            kind = entry[0]
            if kind == "last":
                _, blk, h_in, y_tanh = entry
                conv = blk.conv_layer[0]
                g = dout.contiguous().float()
                if y_tanh is not None:
                    g2 = torch.empty_like(g)
                    call("b2_f32_act", 2, ptr(g), ptr(y_tanh), ptr(g2), g.numel(), stream())
                    g = g2
                c_out = conv.weight.shape[0]
                cp = ((c_out + kal - 1) // kal) * kal
                dz = ops.nchw_to_nhwc_pad(g, cp, code)
                n, hh, ww, _ = dz.shape
                dbias = torch.zeros((cp,), dtype=torch.float32, device=dev)
                ops.act(2, dz, None, None, dbias, n * hh * ww, cp, cp, 0, 0, code)
                lay.view(conv.bias).copy_(dbias[:c_out])
                self._wgrad_conv(0, conv, h_in, dz, 0)
                d = self._dgrad_s1(conv, dz)
            elif kind == "plain":
                _, blk, x_in, z, need_dx = entry
                conv = blk.conv_layer[0]
                dz = self._bwd_act(d, z, lay.view(conv.bias))
                self._wgrad_conv(0, conv, x_in, dz, 0)
                d = self._dgrad_s1(conv, dz) if need_dx else None
            elif kind == "up":
                _, layer, x_in, z = entry
                conv = layer.conv_layer[0]
                dz = self._bwd_act(d, z, lay.view(conv.bias))
                self._wgrad_conv(2, conv, x_in, dz, 2)
                cin, cout = conv.weight.shape[0], conv.weight.shape[1]
                w = self.cache.get(conv.weight, 6, code, cout, cin, cout)
                d = ops.conv2d(4, ops.space_to_depth2(dz), w, None, cin, act=0)
            elif kind == "down":
                _, layer, planes, z, in_shape = entry
                conv = layer.conv_layer[0]
                dz = self._bwd_act(d, z, lay.view(conv.bias))
                self._wgrad_conv(1, conv, planes, dz, 0)
                cout, cin = conv.weight.shape[0], conv.weight.shape[1]
                w = self.cache.get(conv.weight, 5, code, cout, cin, cout)
                d = ops.conv2d(3, dz, w, None, cin, act=0)
            elif kind == "skip_out":
                level = entry[1]
                extra = d_skips.pop(level)
                merged = torch.empty(d.shape, dtype=d.dtype, device=dev)
                ops.add(d, extra, merged)
                d = merged
            elif kind == "cat_in":
                _, level, c_half = entry
                d_skips[level] = d[..., c_half:]
                d = d[..., :c_half]
            elif kind == "res":
                _, rec1, rec2 = entry
                z1 = rec1[2]
                work1 = self._bwd_work(2 * z1.shape[0] * z1.shape[3], dev) if self._colsum_ok(z1) else None
                da1 = self._bwd_gn_conv(rec2, d, ctx, sums=pending, next_sums=(z1, work1) if work1 is not None else None)
                # does the block's input gradient go straight into another residual block (only marks in between)?
                nxt = next((e for e in order[pos + 1:] if e[0] != "mark"), None)
                work_next, z_next = None, None
                if nxt is not None and nxt[0] == "res":
                    z_next = nxt[2][2]
                    if self._colsum_ok(z_next) and tuple(z_next.shape) == tuple(rec1[1].shape):
                        work_next = self._bwd_work(2 * z_next.shape[0] * z_next.shape[3], dev)
                d = self._bwd_gn_conv(rec1, da1, ctx, residual=d, sums=work1,
                                      next_sums=(z_next, work_next) if work_next is not None else None)
                pending = work_next
            elif kind == "attn":
                _, blk, x_in, saved = entry
                d = self._bwd_attention(blk, x_in, saved, d)
            elif kind == "emb":
                self._bwd_embedding(entry[1], ctx)
            elif kind == "mark":
                self._flush_wgrad()              # the module's deferred weight gradients, one grouped launch
                if self.on_grads_ready is not None:
                    lo, hi = lay.module_range(entry[1])
                    if hi > lo:
                        self._join_side()
                        self.on_grads_ready(lay, lo, hi)
        self._flush_wgrad()
        self._join_side()
        for p in lay.params:
            v = lay.view(p)
            if p.grad is None or p.grad.data_ptr() == v.data_ptr():
                p.grad = v
            else:
                p.grad.add_(v)
        if self.on_grads_ready is not None:          # AdaGN scale Linears (front of the buffer) + embedding MLPs (tail)
            self.on_grads_ready(lay, 0, lay.front_end)
            lo, hi = lay.module_range(net.cond_emb)
            if hi > lo:
                self.on_grads_ready(lay, lo, hi)
        tape.clear()
        if self.post_backward is not None:
            self.post_backward(lay)

    def _bwd_attention(self, blk, x, saved, dout):
        code = ops.code_of(x)
        n, hh, ww, c = x.shape
        ldx = x.stride(2)
        p_len, heads, d = hh * ww, blk.heads, blk.d_k
        hd, ldq = heads * d, 3 * heads * d
        rows = n * p_len
        dt, dev = x.dtype, x.device
        lay = self.layout
        qkv, pt, o, ldp = saved["qkv"], saved["pt"], saved["o"], saved["ldp"]
        ldd = dout.stride(2)
        # output projection: out = o Wo^T + bo + x
        ops.act(2, dout, None, None, lay.view(blk.output.bias), rows, c, ldd, 0, 0, code)
        ops.gemm_tn(dout, o, c, hd, rows, ldd, hd, lay.view(blk.output.weight), hd, code=code)
        d_o = torch.empty((rows, hd), dtype=dt, device=dev)
        # bf16 with the optimiser's copy of the weights: the Linear data gradients read W [out][in] in place (MN-major B)
        w_o = lay.shadow_slice(blk.output.weight) if (self.attn_bmn and code == ops.BF16 and lay.shadow is not None
                                                       and c % 64 == 0 and hd % 64 == 0 and id(blk.output.weight) in lay.offsets
                                                       and id(blk.projection.weight) in lay.offsets) else None
        if w_o is not None:
            ops.gemm_nt_bmn(dout, w_o, rows, hd, c, ldd, hd, d_o, hd)
        else:
            wo_t = self.cache.get(blk.output.weight, 4, code, c, hd, c)               # [hd][C]
            ops.gemm_nt(dout, wo_t, rows, hd, c, ldd, c, d_o, hd)
        dqkv = torch.empty((rows, ldq), dtype=dt, device=dev)
        pt_s, t_s = (p_len * ldp, heads * p_len * ldp), (d * ldp, heads * d * ldp)      # batch strides of P^T-shaped / transposed tensors
        qkv_s = (3 * d, p_len * ldq)

        def transposed(src, ld_src, src_strides):
            """[P][d] slices -> [d][ldp] (K-major B operand of the NT kernel for contractions over the query index)."""
            out_t = torch.empty((n, heads, d, ldp), dtype=dt, device=dev)
            call("b2_transpose_batched", ptr(src), ld_src, src_strides[0], src_strides[1], ptr(out_t), ldp, t_s[0], t_s[1],
                 p_len, d, heads, n, code, stream())
            return out_t

        # bf16: dO and Q are consumed in place, MN-major (b2_gemm_nt_bmn) -- no transposed copies (60 launches per step before)
        bmn = self.attn_bmn and code == ops.BF16 and d % 64 == 0
        # dV[j][c] = sum_i P^T[j][i] dO[i][c]
        if bmn:
            ops.gemm_nt_bmn(pt, d_o, p_len, d, p_len, ldp, hd, dqkv[:, 2 * d:], ldq, batch=(heads, n), a_strides=pt_s,
                            b_strides=(d, p_len * hd), c_strides=qkv_s)
        else:
            ops.gemm_nt(pt, transposed(d_o, hd, (d, p_len * hd)), p_len, d, p_len, ldp, ldp, dqkv[:, 2 * d:], ldq, batch=(heads, n),
                        a_strides=pt_s, b_strides=t_s, c_strides=qkv_s, code=code)
        # softmax backward needs sum_i P^T dP^T per key, which equals sum_c V[j][c] dV[j][c]: a row dot product
        dot = torch.empty((rows, heads), dtype=torch.float32, device=dev)
        call("b2_rowdot", ptr(qkv[:, 2 * d:]), ldq, ptr(dqkv[:, 2 * d:]), ldq, 3 * d, ptr(dot), rows, heads, d, code, stream())
        # dS^T = scale * P^T .* (V dO^T - dot): GEMM + softmax backward in one kernel, dP never materialised
        dst = torch.empty((n, heads, p_len, ldp), dtype=dt, device=dev)
        call("b2_attn_scores_bwd", ptr(qkv[:, 2 * d:]), ldq, 3 * d, p_len * ldq, ptr(d_o), hd, d, p_len * hd, ptr(pt), ptr(dot),
             ptr(dst), ldp, p_len, d, heads, n, float(blk.scale), code, stream())
        # dQ[i][c] = sum_j dS^T[j][i] K[j][c]  (TN: contraction index = row index of both operands)
        ops.gemm_tn(dst, qkv[:, d:], p_len, d, p_len, ldp, ldq, dqkv, ldq, out_mode=1, batch=(heads, n),
                    a_strides=pt_s, b_strides=qkv_s, c_strides=qkv_s, code=code)
        # dK[j][c] = sum_i dS^T[j][i] Q[i][c]
        if bmn:
            ops.gemm_nt_bmn(dst, qkv, p_len, d, p_len, ldp, ldq, dqkv[:, d:], ldq, batch=(heads, n), a_strides=pt_s,
                            b_strides=qkv_s, c_strides=qkv_s)
        else:
            ops.gemm_nt(dst, transposed(qkv, ldq, qkv_s), p_len, d, p_len, ldp, ldp, dqkv[:, d:], ldq, batch=(heads, n),
                        a_strides=pt_s, b_strides=t_s, c_strides=qkv_s, code=code)
        # input projection: qkv = x Wp^T + bp
        ops.act(2, dqkv, None, None, lay.view(blk.projection.bias), rows, ldq, ldq, 0, 0, code)
        ops.gemm_tn(dqkv, x, ldq, c, rows, ldq, ldx, lay.view(blk.projection.weight), c, code=code)
        dx = torch.empty((n, hh, ww, c), dtype=dt, device=dev)
        if w_o is not None:
            ops.gemm_nt_bmn(dqkv, lay.shadow_slice(blk.projection.weight), rows, c, ldq, ldq, c, dx, c, residual=dout, ldr=ldd)
        else:
            wp_t = self.cache.get(blk.projection.weight, 4, code, ldq, c, ldq)          # [C][3hd]
            ops.gemm_nt(dqkv, wp_t, rows, c, ldq, ldq, ldq, dx, c, residual=dout, ldr=ldd)
        return dx

    def _ones(self, n, device):
        """[n, 1] fp32 ones (column sums as a small GEMM), cached: nothing is allocated or filled inside a captured step."""
        cache = getattr(self, "_ones_cache", None)
        if cache is None:
            cache = self._ones_cache = {}
        key = (n, str(device))
        if key not in cache:
            if torch.cuda.is_current_stream_capturing():
                return torch.ones((n, 1), dtype=torch.float32, device=device)
            cache[key] = torch.ones((n, 1), dtype=torch.float32, device=device)
        return cache[key]

    def _bwd_embedding(self, rec, ctx):
        (rec_t, rec_c, bt) = rec
        lay = self.layout
        emb, ds_all, total, be = ctx["emb"], ctx["ds_all"], ctx["total"], ctx["be"]
        w_all, b_all, off, _ = self._adagn_table()
        dim = emb.shape[1]
        mods = self._adagn_modules()
        dev = emb.device
        # gradients of all AdaGN scale Linears in one shot: dW_all = ds_all^T emb, db_all = colsum(ds_all)
        if lay.dense_adagn:
            dw_all = lay.flat[:lay.adagn_w_numel].view(total, dim)
            db_all = lay.flat[lay.adagn_w_numel:lay.adagn_w_numel + lay.adagn_b_numel]
        else:
            dw_all = torch.zeros((total, dim), dtype=torch.float32, device=dev)
            db_all = torch.zeros((total,), dtype=torch.float32, device=dev)
        ops.small_gemm(ds_all, emb, total, dim, be, total, dim, dw_all, dim, ta=1, tb=1, accumulate=True)
        # bias gradient = column sums of ds_all: the same kernel against a vector of ones (no ATen reduce + add in the step)
        ops.small_gemm(ds_all, self._ones(be, dev), total, 1, be, total, 1, db_all, 1, ta=1, tb=1, accumulate=True)
        if not lay.dense_adagn:
            o = 0
            for m in mods:
                cch = m.y_scale.weight.shape[0]
                lay.view(m.y_scale.weight).copy_(dw_all[o:o + cch])
                lay.view(m.y_scale.bias).copy_(db_all[o:o + cch])
                o += cch
        demb = torch.empty((be, dim), dtype=torch.float32, device=dev)
        ops.small_gemm(ds_all, w_all, be, dim, total, total, dim, demb, dim, tb=1)

        def mlp_bwd(rec_m, dy):
            lins, acts, pres = rec_m
            b = dy.shape[0]
            for i in (3, 2, 1, 0):
                lin, h_in = lins[i], acts[i]
                n_out, n_in = lin.weight.shape
                ops.small_gemm(dy, h_in, n_out, n_in, b, n_out, n_in, lay.view(lin.weight), n_in, ta=1, tb=1, accumulate=True)
                ops.small_gemm(dy, self._ones(b, dev), n_out, 1, b, n_out, 1, lay.view(lin.bias), 1, ta=1, tb=1, accumulate=True)
                if i == 0:
                    break
                dh = torch.empty((b, n_in), dtype=torch.float32, device=dev)
                ops.small_gemm(dy, lin.weight, b, n_in, n_out, n_out, n_in, dh, n_in, tb=1)
                dpre = torch.empty_like(dh)
                call("b2_f32_act", 1, ptr(dh), ptr(pres[i - 1]), ptr(dpre), dh.numel(), stream())
                dy = dpre

        mlp_bwd(rec_t, demb if demb.shape[0] == bt else demb.sum(dim=0, keepdim=True))
        if rec_c is not None:
            bc = rec_c[1][0].shape[0]
            mlp_bwd(rec_c, demb if demb.shape[0] == bc else demb.sum(dim=0, keepdim=True))
