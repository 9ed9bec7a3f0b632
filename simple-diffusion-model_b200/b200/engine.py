"""U-Net executor: walks the module tree of models/U_Net.py and issues the sm_100a kernels.

Design (B200-first, not a translation of the reference's eager graph):
  * activations live NHWC in the compute dtype (bf16, or fp32 for the TF32 parity mode); NCHW fp32 exists only at the
    two network edges, and both edges are fused (input pad/transpose kernel; final conv epilogue writes NCHW fp32);
  * every `torch.cat` of the reference (models/U_Net.py:168) is zero-copy: the down-sampler and the producer of `x`
    write straight into the two channel halves of one pre-allocated buffer;
  * Swish, bias, GroupNorm statistics and residual adds never run as separate passes: they are conv/GEMM epilogues or
    the single AdaGN-apply pass;
  * all 2 x (#res blocks) AdaGN scale vectors of a forward come from ONE small GEMM over a concatenated weight.
Parameters stay fp32 in the reference's shapes (state_dict compatible); kernel-layout copies are cached here and
refreshed when a parameter's version counter changes.
"""
import torch

from . import ops
from ._lib import B200Error, call, ptr, stream


def _version_key(params):
    return tuple((p.data_ptr(), p._version, getattr(getattr(p, "_b2_layout", None), "epoch", 0)) for p in params)


class WeightCache:
    def __init__(self):
        self._packed = {}
        self._log = {}            # key -> request: every derived layout a step asked for (refresh_all re-derives them in bulk)
        self._tables = {}         # job-table device tensors keyed by the pointers they describe
        self.bulk_transposes = False

    def get(self, param, kind, code, cout, cin, cin_pad):
        key = (id(param), kind, code, cin_pad)
        lay = getattr(param, "_b2_layout", None)      # fused optimiser updates bypass torch's version counter
        if lay is not None and code == ops.BF16 and lay.shadow is not None:
            # training with the fused optimiser: its bf16 copy already is the kernel layout of channels-last stored
            # 3x3 weights (kind 0) and of Linear weights (kind 3); the stride-1 data-gradient layout is one transpose away
            if (kind == 0 and lay.is_cl(param) and cin_pad == cin) or (kind == 3 and cin_pad == cin and id(param) in lay.offsets):
                sl = lay.shadow_slice(param)
                return sl.view(cout, -1)
            if kind == 1 and lay.is_cl(param) and cin_pad == cout and cout % 64 == 0:
                ver = (param.data_ptr(), param._version, lay.epoch)
                hit = self._packed.get(key)
                self._log.setdefault(key, (param, kind, code, cout, cin, cin_pad, "tr"))
                if hit is None or hit[0] != ver:
                    out = hit[1] if hit is not None else torch.empty((cin, 9 * cout), dtype=torch.bfloat16, device=param.device)
                    call("b2_transpose_weight_cl", ptr(lay.shadow_slice(param)), ptr(out), cout, cin, stream())
                    hit = (ver, out)
                    self._packed[key] = hit
                return hit[1]
            # Linear data-gradient weights (a transpose) and both ConvTranspose layouts come from the bf16 copy as well,
            # through shared-memory tiles: the generic fp32 gather reads one 32-byte sector per element for these layouts
            # (1.3 ms per train step when it handled them)
            derived = None
            if self._from_shadow(param, lay, kind, code, cout, cin, cin_pad):
                if kind == 4:
                    derived = ((cin, cout), lambda src, out: call("b2_transpose_linear_weight", ptr(src), ptr(out), cout, cin, stream()))
                elif kind == 2:
                    derived = ((4 * cout, 4 * cin), lambda src, out: call("b2_pack_convt_bf16", ptr(src), ptr(out), None, cin, cout, stream()))
                else:
                    derived = ((cin, 16 * cout), lambda src, out: call("b2_pack_convt_bf16", ptr(src), None, ptr(out), cin, cout, stream()))
            if derived is not None:
                ver = (param.data_ptr(), param._version, lay.epoch)
                hit = self._packed.get(key)
                if hit is None or hit[0] != ver or tuple(hit[1].shape) != derived[0]:
                    out = hit[1] if hit is not None and tuple(hit[1].shape) == derived[0] else \
                        torch.empty(derived[0], dtype=torch.bfloat16, device=param.device)
                    derived[1](lay.shadow_slice(param), out)
                    hit = (ver, out)
                    self._packed[key] = hit
                return hit[1]
        ver = (param.data_ptr(), param._version, lay.epoch if lay is not None else 0)
        hit = self._packed.get(key)
        if lay is not None:
            self._log.setdefault(key, (param, kind, code, cout, cin, cin_pad, "pack"))
        if hit is None or hit[0] != ver:
            hit = (ver, ops.pack_weight(kind, param, cout, cin, cin_pad, code))
            self._packed[key] = hit
        return hit[1]

    @staticmethod
    def _from_shadow(param, lay, kind, code, cout, cin, cin_pad):
        """True when get() derives this layout from the optimiser's bf16 copy with a tiled kernel (so refresh_all leaves it)."""
        if lay is None or code != ops.BF16 or lay.shadow is None or id(param) not in lay.offsets or lay.is_cl(param):
            return False
        if kind == 4:
            return cin_pad == cout and cout % 64 == 0 and cin % 64 == 0
        return kind in (2, 6) and cin % 32 == 0 and cout % 32 == 0

    def _table(self, name, rows, device):
        """Device int64 job table, cached by content (pointers are stable across steps, so steady state uploads nothing --
        and nothing is uploaded inside a CUDA-graph capture, whose eager warm-up steps build the tables first)."""
        key = (name, tuple(rows))
        t = self._tables.get(key)
        if t is None:
            if torch.cuda.is_current_stream_capturing():
                return None
            if len(self._tables) > 16:
                self._tables.clear()
            t = torch.tensor(rows, dtype=torch.int64, device=device).reshape(-1)
            self._tables[key] = t
        return t

    def refresh_all(self):
        """Re-derives every stale kernel-layout weight a previous step asked for with ONE launch per family (bf16 transposes
        of the channels-last shadow; fp32 -> kernel-layout packs) instead of ~180 few-microsecond launches per train step."""
        tr, packs = [], {}
        pending = []
        stale = []
        for key, (param, kind, code, cout, cin, cin_pad, family) in self._log.items():
            lay = getattr(param, "_b2_layout", None)
            if lay is None:
                continue
            ver = (param.data_ptr(), param._version, lay.epoch)
            hit = self._packed.get(key)
            if hit is not None and hit[0] == ver:
                continue
            if family == "tr":
                # measured: producing the 1.2 GB of data-gradient transposes up front costs more than it saves -- made just
                # in time in the backward pass each one is still in L2 when its conv reads it -- so they stay lazy
                if lay.shadow is None or not self.bulk_transposes:
                    continue
                out = hit[1] if hit is not None else torch.empty((cin, 9 * cout), dtype=torch.bfloat16, device=param.device)
                src = lay.shadow_slice(param)
                tiles = (cin // 64) * (cout // 64) * 9
                tr.append((src.data_ptr(), out.data_ptr(), cout, cin, tiles))
                pending.append((key, ver, out))
            else:
                if not param.is_contiguous():
                    continue                      # channels-last stored parameter read through a permuted view: stays lazy
                if self._from_shadow(param, lay, kind, code, cout, cin, cin_pad):
                    continue                      # made just in time from the bf16 copy (get)
                if code == ops.BF16 and lay.shadow is not None and cin_pad == cin and id(param) in lay.offsets and \
                        ((kind == 0 and lay.is_cl(param)) or kind == 3):
                    # get() serves this layout straight from the optimiser's bf16 copy; the key was logged by the very first
                    # forward, before that copy existed.  Re-packing it here cost one 0.48 ms launch per train step (all attention
                    # Linear weights, round 2b) for buffers nobody read.
                    stale.append(key)
                    continue
                shape = {0: (cout, 9 * cin_pad), 1: (cin, 9 * cin_pad), 2: (4 * cout, 4 * cin), 3: (cout, cin_pad), 4: (cin, cin_pad),
                         5: (4 * cin, 4 * cout), 6: (cin, 16 * cout)}[kind]
                out = hit[1] if hit is not None and tuple(hit[1].shape) == shape else \
                    torch.empty(shape, dtype=ops.TORCH_DTYPE[code], device=param.device)
                packs.setdefault(code, []).append((param.data_ptr(), out.data_ptr(), kind, cout, cin, cin_pad, out.numel()))
                pending.append((key, ver, out))
        for key in stale:
            self._log.pop(key, None)
            self._packed.pop(key, None)
        if not pending:
            return
        dev = pending[0][2].device
        ok = True
        if tr:
            rows, start = [], 0
            for src, out, cout, cin, tiles in tr:
                rows.append((src, out, cout, cin, start, tiles))
                start += tiles
            table = self._table("tr", rows, dev)
            if table is None:
                ok = False
            else:
                call("b2_transpose_weight_cl_multi", ptr(table), len(rows), start, stream())
        import os
        if os.environ.get("SDM_B200_DEBUG_PACKS") and packs:
            for code, jobs in packs.items():
                print("[b200] bulk weight packs:", [(kind, cout, cin, numel) for _, _, kind, cout, cin, _, numel in jobs], flush=True)
        for code, jobs in packs.items():
            rows, start = [], 0
            for w, out, kind, cout, cin, k_pad, numel in jobs:
                rows.append((w, out, kind, cout, cin, k_pad, start, numel))
                start += numel
            table = self._table(("pack", code), rows, dev)
            if table is None:
                ok = False
            else:
                call("b2_pack_weight_multi", ptr(table), len(rows), start, code, stream())
        if ok:
            for key, ver, out in pending:
                self._packed[key] = (ver, out)

    def get_edge(self, param, which):
        """fp32 weight layouts of the CUDA-core edge convs: "first" -> [Cin*9][Cout], "last" -> [9][Cin][4] (tiny tensors,
        re-derived with torch ops when the parameter changes)."""
        key = (id(param), which)
        lay = getattr(param, "_b2_layout", None)
        ver = (param.data_ptr(), param._version, lay.epoch if lay is not None else 0)
        hit = self._packed.get(key)
        if hit is None or hit[0] != ver:
            w = param.detach().float()
            cout, cin = w.shape[0], w.shape[1]
            if which == "first":
                t = w.permute(1, 2, 3, 0).reshape(cin * 9, cout).contiguous()
            else:
                t = torch.zeros((9, cin, 4), dtype=torch.float32, device=w.device)
                t[:, :, :cout] = w.permute(2, 3, 1, 0).reshape(9, cin, cout)
            hit = (ver, t)
            self._packed[key] = hit
        return hit[1]

    def clear(self):
        self._packed.clear()


class UNetEngine:
    def __init__(self, net):
        self.net = net
        self.cache = WeightCache()
        self._adagn_key = None
        self._adagn_w = None
        self._adagn_b = None
        self._adagn_off = None

    # ------------------------------------------------------------------------------------------ helpers
    def _code(self):
        return ops.TF32 if self.net.precision == "tf32" else ops.BF16

    def _adagn_modules(self):
        from models.custom_layers import AdaGN
        return [m for m in self.net.modules() if isinstance(m, AdaGN)]

    def _adagn_table(self):
        """Concatenated y_scale weights of every AdaGN: one GEMM yields all scale vectors (custom_layers.py:38-42)."""
        mods = self._adagn_modules()
        params = [p for m in mods for p in (m.y_scale.weight, m.y_scale.bias)]
        key = _version_key(params)
        if key != self._adagn_key:
            self._adagn_w = torch.cat([m.y_scale.weight.detach().float() for m in mods], dim=0).contiguous()
            self._adagn_b = torch.cat([m.y_scale.bias.detach().float() for m in mods], dim=0).contiguous()
            off, o = {}, 0
            for m in mods:
                off[id(m)] = o
                o += m.y_scale.weight.shape[0]
            self._adagn_off, self._adagn_total, self._adagn_key = off, o, key
        return self._adagn_w, self._adagn_b, self._adagn_off, self._adagn_total

    def _mlp(self, seq, x, b, dim_in, out, accumulate_last=False):
        """Linear/Swish x3 + Linear on CUDA cores (tiny: custom_layers.py:59-77)."""
        lins = [seq[0], seq[2], seq[4], seq[6]]
        h = x
        k = dim_in
        for i, lin in enumerate(lins):
            n = lin.weight.shape[0]
            dst = out if i == 3 else torch.empty((b, n), dtype=torch.float32, device=x.device)
            ops.small_gemm(h, lin.weight, b, n, k, k, lin.weight.shape[1], dst, n, bias=lin.bias, act=1 if i < 3 else 0,
                           accumulate=(i == 3 and accumulate_last))
            h, k = dst, n
        return h

    def embedding(self, t, cond):
        ce = self.net.cond_emb
        dim = ce.time_dim
        t = t.to(torch.int64).contiguous()
        bt = t.shape[0]
        sin = torch.empty((bt, dim), dtype=torch.float32, device=t.device)
        call("b2_sinusoid_embedding", ptr(t), ptr(sin), bt, dim, stream())
        emb = torch.empty((bt, dim), dtype=torch.float32, device=t.device)
        self._mlp(ce.time_layer, sin, bt, dim, emb)
        if ce.cond_layer is not None:
            if cond is None:
                raise B200Error("this U_Net was built with cond_dim: `cond` is required")
            c2 = cond.float().reshape(-1, cond.shape[-1]).contiguous()
            bc = c2.shape[0]
            if bc == bt:
                # the last cond Linear accumulates straight into the time embedding (emb = time + cond, :96-97)
                self._mlp(ce.cond_layer, c2, bc, c2.shape[1], emb, accumulate_last=True)
            else:
                cemb = torch.empty((bc, dim), dtype=torch.float32, device=t.device)
                self._mlp(ce.cond_layer, c2, bc, c2.shape[1], cemb)
                emb = emb + cemb              # batch-broadcast corner (t of shape [1] with batched labels)
        return emb

    # ------------------------------------------------------------------------------------------ layers
    def conv_block(self, blk, x, ctx, out=None, residual=None, out_nchw=None, act_override=None):
        """UNet_ConvBlock (custom_layers.py:240-245): conv + bias + Swish (+GN stats) then AdaGN apply (+residual)."""
        conv = blk.conv_layer[0]
        code = ops.code_of(x)
        cout, cin = conv.weight.shape[0], conv.weight.shape[1]
        cin_pad = x.shape[3]
        w = self.cache.get(conv.weight, 0, code, cout, cin, cin_pad)
        act = 1 if blk.use_activation else 0
        if act_override is not None:
            act = act_override
        has_gn = ctx["emb"] is not None and blk.adagn is not None
        if not has_gn:
            return ops.conv2d(0, x, w, conv.bias, cout, act=act, out=out, residual=residual, out_nchw_fp32=out_nchw)
        n = x.shape[0]
        groups = blk.adagn.group_norm.num_groups
        stats = ctx["stats"][ctx["stats_i"]]
        ctx["stats_i"] += 1
        y = ops.conv2d(0, x, w, conv.bias, cout, act=act, gn_stats=stats, groups=groups)
        off = ctx["adagn_off"][id(blk.adagn)]
        s = ctx["s_all"][:, off:off + cout]
        gn = blk.adagn.group_norm
        return ops.adagn_apply(y, stats, gn.weight, gn.bias, s, ctx["s_bstride"], out=out, residual=residual,
                               groups=groups, eps=gn.eps)

    def residual_block(self, blk, x, ctx, out=None):
        h = self.conv_block(blk.conv_block_1, x, ctx)
        return self.conv_block(blk.conv_block_2, h, ctx, out=out, residual=x)

    def attention(self, blk, x, out=None, save=None):
        """AttentionBlock (custom_layers.py:127-163): softmax over the query axis, no norm, residual add."""
        code = ops.code_of(x)
        n, hh, ww, c = x.shape
        ldx = x.stride(2)
        p_len, heads, d = hh * ww, blk.heads, blk.d_k
        dt, dev = x.dtype, x.device
        kal = ops.K_ALIGN[code]
        wp = self.cache.get(blk.projection.weight, 3, code, 3 * heads * d, c, c)
        wo = self.cache.get(blk.output.weight, 3, code, c, heads * d, heads * d)
        qkv = torch.empty((n * p_len, 3 * heads * d), dtype=dt, device=dev)
        ops.gemm_nt(x, wp, n * p_len, 3 * heads * d, c, ldx, c, qkv, 3 * heads * d, bias=blk.projection.bias)
        ldq = 3 * heads * d
        # P^T[n][h][j][i] = softmax_i(scale * q_i . k_j): scores and the query-axis softmax in ONE tensor-core kernel
        # (S^T = K Q^T, one thread per key row in TMEM) -- the fp32 score matrix of the reference never exists
        ldp = ((p_len + 7) // 8) * 8
        pt = torch.empty((n, heads, p_len, ldp), dtype=dt, device=dev)
        tiles = (p_len + 255) // 256
        work = torch.empty((2 * n * heads * p_len * tiles,), dtype=torch.float32, device=dev) if tiles > 1 else None
        call("b2_attn_scores_softmax", ptr(qkv[:, d:]), ptr(qkv), ldq, 3 * d, p_len * ldq, ptr(pt), ldp, p_len, d, heads, n,
             float(blk.scale), ptr(work), code, stream())
        # O[i][c] = sum_j P^T[j][i] V[j][c]: both operands have the contraction index as their row index, which is the
        # TN kernel's native form -- V is consumed in place inside qkv, no transpose
        o = torch.empty((n * p_len, heads * d), dtype=dt, device=dev)
        ops.gemm_tn(pt, qkv[:, 2 * d:], p_len, d, p_len, ldp, ldq, o, heads * d, out_mode=1, batch=(heads, n),
                    a_strides=(p_len * ldp, heads * p_len * ldp), b_strides=(3 * d, p_len * ldq),
                    c_strides=(d, p_len * heads * d), code=code)
        if out is None:
            out = torch.empty((n, hh, ww, c), dtype=dt, device=dev)
        ops.gemm_nt(o, wo, n * p_len, c, heads * d, heads * d, heads * d, out, out.stride(2), bias=blk.output.bias,
                    residual=x, ldr=ldx)
        if save is not None:
            save.update(qkv=qkv, pt=pt, o=o, ldp=ldp)
        return out

    def unet_block(self, blk, x, ctx, out):
        """UNetBlock (custom_layers.py:336-341); `out` receives the sampler output (may be a channel slice)."""
        from models.custom_layers import AttentionBlock, UpsampleBlock
        for res, attn in zip(blk.res_layers, blk.attn_layers):
            x = self.residual_block(res, x, ctx)
            if isinstance(attn, AttentionBlock):
                x = self.attention(attn, x)
        return self.resample(blk.out_layer, x, out)

    def resample(self, layer, x, out=None):
        """UpsampleBlock (ConvTranspose 4x4/s2 + Swish, custom_layers.py:169-185) / DownsampleBlock (Conv 3x3/s2 + Swish,
        :191-207); `out` may be a channel slice of a concat buffer."""
        from models.custom_layers import UpsampleBlock
        code = ops.code_of(x)
        conv = layer.conv_layer[0]
        if isinstance(layer, UpsampleBlock):
            cin, cout = conv.weight.shape[0], conv.weight.shape[1]
            w = self.cache.get(conv.weight, 2, code, cout, cin, cin)
            return ops.conv2d(2, x, w, conv.bias, cout, act=1, out=out)
        cout, cin = conv.weight.shape[0], conv.weight.shape[1]
        w = self.cache.get(conv.weight, 0, code, cout, cin, cin)
        planes = ops.space_to_depth2(x)
        return ops.conv2d(1, planes, w, conv.bias, cout, act=1, out=out)

    # ------------------------------------------------------------------------------------------ whole network
    @torch.no_grad()
    def forward(self, x, t=None, cond=None):
        """models/U_Net.py:147-173.  x fp32 NCHW CUDA -> fp32 NCHW CUDA."""
        net = self.net
        if not x.is_cuda:
            raise B200Error("U_Net.forward needs CUDA tensors: this build has no CPU path")
        code = self._code()
        n, cin, hgt, wid = x.shape
        levels = len(net.down_layers)
        if hgt % (1 << levels) or wid % (1 << levels):
            raise B200Error(f"H and W must be divisible by 2**num_layers = {1 << levels}")
        ctx = {"emb": None, "stats_i": 0}
        if net.cond_emb is not None:
            if t is None:
                raise B200Error("timestep tensor `t` is required")
            emb = self.embedding(t.to(x.device), cond.to(x.device) if cond is not None else None)
            w_all, b_all, off, total = self._adagn_table()
            be = emb.shape[0]
            if be not in (1, n):
                raise B200Error(f"embedding batch {be} does not broadcast over image batch {n}")
            s_all = torch.empty((be, total), dtype=torch.float32, device=x.device)
            ops.small_gemm(emb, w_all, be, total, emb.shape[1], emb.shape[1], w_all.shape[1], s_all, total, bias=b_all)
            n_adagn = len(off)
            max_groups = max(m.group_norm.num_groups for m in self._adagn_modules())
            ctx.update(emb=emb, s_all=s_all, adagn_off=off, s_bstride=(total if be == n else 0),
                       stats=torch.zeros((n_adagn, n, max_groups, 2), dtype=torch.float32, device=x.device))
        first = net.in_layer[0]
        if ops.edge_first_ok(first.conv_layer[0]) and first.adagn is None and first.use_activation and wid % 2 == 0:
            conv0 = first.conv_layer[0]           # 3/6 input channels: CUDA-core kernel straight from the fp32 NCHW image
            h = ops.conv_first(x.contiguous().float(), self.cache.get_edge(conv0.weight, "first"), conv0.bias,
                               conv0.weight.shape[0], 1, code)
        else:
            cpad = ((cin + ops.K_ALIGN[code] - 1) // ops.K_ALIGN[code]) * ops.K_ALIGN[code]
            h = ops.nchw_to_nhwc_pad(x, cpad, code)
            h = self.conv_block(first, h, ctx)
        h = self.conv_block(net.in_layer[1], h, ctx)
        cats = []
        hh, ww = hgt, wid
        for blk in net.down_layers:
            cout = blk.out_layer.conv_layer[0].weight.shape[0]
            hh, ww = hh // 2, ww // 2
            cat = ops.new_act(n, hh, ww, 2 * cout, code, x.device)
            h = self.unet_block(blk, h, ctx, out=cat[..., cout:])       # skip half of the future concat
            cats.append(cat)
        h = self.conv_block(net.middle_layer[0], h, ctx)
        c_mid = cats[-1].shape[3] // 2
        self.conv_block(net.middle_layer[1], h, ctx, out=cats[-1][..., :c_mid])
        n_up = len(net.up_layers)
        for i, blk in enumerate(net.up_layers):
            cat = cats.pop()
            cout = blk.out_layer.conv_layer[0].weight.shape[1]
            if i + 1 < n_up:
                dst = cats[-1][..., :cout]
            else:
                dst = None
            h = self.unet_block(blk, cat, ctx, out=dst)
        h = self.conv_block(net.out_layers[0], h, ctx)
        last = net.out_layers[1]
        c_out = last.conv_layer[0].weight.shape[0]
        y = torch.empty((n, c_out, hgt, wid), dtype=torch.float32, device=x.device)
        if ops.edge_last_ok(last.conv_layer[0]) and last.adagn is None and wid % 4 == 0:
            conv_l = last.conv_layer[0]
            ops.conv_last(h, self.cache.get_edge(conv_l.weight, "last"), conv_l.bias, c_out, 2 if net.image_recon else 0, y)
        else:
            self.conv_block(last, h, ctx, out_nchw=y, act_override=(2 if net.image_recon else 0))
        return y
