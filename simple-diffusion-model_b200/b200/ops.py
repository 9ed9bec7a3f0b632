"""Thin Python wrappers over the C ABI.  Activations are torch CUDA tensors laid out NHWC ([N, H, W, C] with
unit channel stride); a tensor may be a channel slice of a wider buffer -- the per-pixel stride is passed down as `ld`.
dtype codes: 0 = bf16 storage / kind::f16 MMA, 1 = fp32 storage / kind::tf32 MMA."""
import torch

from ._lib import B200Error, call, ptr, stream

BF16, TF32 = 0, 1
TORCH_DTYPE = {BF16: torch.bfloat16, TF32: torch.float32}
K_ALIGN = {BF16: 64, TF32: 32}          # channels per 128-byte K block


def code_of(t):
    if t.dtype == torch.bfloat16:
        return BF16
    if t.dtype == torch.float32:
        return TF32
    raise B200Error(f"unsupported activation dtype {t.dtype}")


def _nhwc(t):
    if t.dim() != 4 or t.stride(3) != 1:
        raise B200Error("expected an NHWC tensor with unit channel stride")
    n, h, w, c = t.shape
    ld = t.stride(2)
    if (h > 1 and t.stride(1) != w * ld) or (n > 1 and t.stride(0) != h * w * ld):
        raise B200Error("NHWC tensor must be dense in N, H, W (only the channel dim may be a slice)")
    return n, h, w, c, ld


def new_act(n, h, w, c, code, device):
    return torch.empty((n, h, w, c), dtype=TORCH_DTYPE[code], device=device)


def conv2d(mode, x, wpacked, bias, cout, act=0, out=None, residual=None, gn_stats=None, groups=32, out_nchw_fp32=None, colsum=None):
    """mode 0: 3x3/s1; 1: 3x3/s2 (x = parity planes [4*N, H/2, W/2, C]); 2: transposed 4x4/s2;
    3: data gradient of mode 1 (x = dz); 4: data gradient of mode 2 (x = parity planes of dz)."""
    code = code_of(x)
    n, h, w, cin, ldx = _nhwc(x)
    if mode in (1, 4):          # the A operand is a stack of 4 parity planes
        n //= 4
    oh, ow = (2 * h, 2 * w) if mode in (2, 3) else (h, w)
    if out_nchw_fp32 is not None:
        y, ldy, out_mode = out_nchw_fp32, 0, 1
    else:
        if out is None:
            out = new_act(n, oh, ow, cout, code, x.device)
        y, out_mode = out, 0
        ldy = _nhwc(out)[4]
    ldr = _nhwc(residual)[4] if residual is not None else 0
    if colsum is not None:
        # (z, s1, s2): the epilogue also accumulates sum_p y and sum_p y * swish(z) per (image, channel) -- the pass-1 sums of
        # the AdaGN backward that consumes y (b2_conv2d_nhwc_colsum)
        cz, s1, s2 = colsum
        call("b2_conv2d_nhwc_colsum", mode, ptr(x), n, h, w, cin, ldx, ptr(wpacked), ptr(bias), cout, ptr(y), ldy, act,
             ptr(residual), ldr, ptr(gn_stats), groups if gn_stats is not None else 0, out_mode, code, stream(),
             ptr(cz), _nhwc(cz)[4], ptr(s1), ptr(s2))
        return y
    call("b2_conv2d_nhwc", mode, ptr(x), n, h, w, cin, ldx, ptr(wpacked), ptr(bias), cout, ptr(y), ldy, act,
         ptr(residual), ldr, ptr(gn_stats), groups if gn_stats is not None else 0, out_mode, code, stream())
    return y


def conv2d_dual(mode, x, wpacked, bias, cout, out_act):
    """Forward conv (modes 0..2) + bias that stores the pre-activation z AND Swish(z) (into `out_act`, which may be a channel slice
    of a concat buffer) from one epilogue.  Returns z."""
    code = code_of(x)
    n, h, w, cin, ldx = _nhwc(x)
    if mode == 1:
        n //= 4
    oh, ow = (2 * h, 2 * w) if mode == 2 else (h, w)
    z = new_act(n, oh, ow, cout, code, x.device)
    if tuple(out_act.shape) != (n, oh, ow, cout) or out_act.dtype != z.dtype:
        raise B200Error(f"conv2d_dual: out_act {tuple(out_act.shape)} does not match the conv output {(n, oh, ow, cout)}")
    call("b2_conv2d_nhwc_dual", mode, ptr(x), n, h, w, cin, ldx, ptr(wpacked), ptr(bias), cout, ptr(z), _nhwc(z)[4],
         ptr(out_act), _nhwc(out_act)[4], code, stream())
    return z


def gemm_nt(a, b, m, ncols, k, lda, ldb, out, ldc, bias=None, alpha=1.0, act=0, residual=None, ldr=0, out_fp32=False,
            batch=(1, 1), a_strides=(0, 0), b_strides=(0, 0), c_strides=(0, 0), code=None):
    code = code_of(a) if code is None else code
    call("b2_gemm_nt", ptr(a), lda, a_strides[0], a_strides[1], ptr(b), ldb, b_strides[0], b_strides[1], ptr(out), ldc,
         c_strides[0], c_strides[1], m, ncols, k, batch[0], batch[1], ptr(bias), float(alpha), act, ptr(residual), ldr,
         1 if out_fp32 else 0, code, stream())
    return out


def gemm_nt_bmn(a, b, m, ncols, k, lda, ldb, out, ldc, alpha=1.0, batch=(1, 1), a_strides=(0, 0), b_strides=(0, 0),
                c_strides=(0, 0), residual=None, ldr=0):
    """out = alpha * a @ b (+ residual) with b given un-transposed ([k][ncols] rows of stride ldb per batch entry): bf16,
    ncols % 64 == 0."""
    call("b2_gemm_nt_bmn", ptr(a), lda, a_strides[0], a_strides[1], ptr(b), ldb, b_strides[0], b_strides[1], ptr(out), ldc,
         c_strides[0], c_strides[1], m, ncols, k, batch[0], batch[1], float(alpha), ptr(residual), ldr, code_of(a), stream())
    return out


def nchw_to_nhwc_pad(x, cpad, code):
    n, c, h, w = x.shape
    x = x.contiguous().float()
    y = new_act(n, h, w, cpad, code, x.device)
    call("b2_nchw_to_nhwc_pad", ptr(x), ptr(y), n, c, h, w, cpad, code, stream())
    return y


def nhwc_to_nchw(x):
    n, h, w, c, ld = _nhwc(x)
    y = torch.empty((n, c, h, w), dtype=torch.float32, device=x.device)
    call("b2_nhwc_to_nchw", ptr(x), ld, ptr(y), n, c, h, w, code_of(x), stream())
    return y


def space_to_depth2(x):
    n, h, w, c, ld = _nhwc(x)
    planes = torch.empty((4 * n, h // 2, w // 2, c), dtype=x.dtype, device=x.device)
    call("b2_space_to_depth2", ptr(x), ld, ptr(planes), n, h, w, c, code_of(x), stream())
    return planes


def pack_weight(kind, w, cout, cin, k_pad, code):
    """fp32 parameter -> kernel layout (kinds documented at b2_pack_weight in include/sdm_b200.h)."""
    shape = {0: (cout, 9 * k_pad), 1: (cin, 9 * k_pad), 2: (4 * cout, 4 * cin), 3: (cout, k_pad), 4: (cin, k_pad),
             5: (4 * cin, 4 * cout), 6: (cin, 16 * cout)}[kind]
    out = torch.empty(shape, dtype=TORCH_DTYPE[code], device=w.device)
    call("b2_pack_weight", kind, ptr(w.detach().contiguous()), ptr(out), cout, cin, k_pad, code, stream())
    return out


def adagn_apply(y, stats, gamma, beta, s, s_bstride, out=None, residual=None, groups=32, eps=1e-5, pre_swish=False):
    n, h, w, c, ldy = _nhwc(y)
    if out is None:
        out = torch.empty((n, h, w, c), dtype=y.dtype, device=y.device)
    ldo = _nhwc(out)[4]
    ldr = _nhwc(residual)[4] if residual is not None else 0
    call("b2_adagn_apply", ptr(y), ldy, ptr(stats), ptr(gamma), ptr(beta), ptr(s), s_bstride, ptr(residual), ldr, ptr(out),
         ldo, n, h * w, c, groups, float(eps), 1 if pre_swish else 0, code_of(y), stream())
    return out


def small_gemm(a, b, m, n, k, lda, ldb, out, ldc, ta=0, tb=0, bias=None, act=0, accumulate=False):
    call("b2_small_gemm", ptr(a), lda, ta, ptr(b), ldb, tb, ptr(out), ldc, m, n, k, ptr(bias), act, 1 if accumulate else 0, stream())
    return out


def conv2d_wgrad(mode, x, dz, cout, grad_packed):
    """Accumulates the weight gradient (kernel layout, fp32) of conv mode 0/1/2 into the zeroed `grad_packed`."""
    code = code_of(x)
    n, h, w, cin, ldx = _nhwc(x)
    if mode == 1:
        n //= 4
    lddz = _nhwc(dz)[4]
    call("b2_conv2d_wgrad", mode, ptr(x), n, h, w, cin, ldx, ptr(dz), cout, lddz, ptr(grad_packed), code, stream())


def conv2d_wgrad_batch(jobs):
    """jobs: list of (mode, x, dz, cout, grad_packed) as for conv2d_wgrad; one grouped launch per N-tile width (b2_conv2d_wgrad_batch)."""
    if not jobs:
        return
    import ctypes
    code = code_of(jobs[0][1])
    rows = []
    for mode, x, dz, cout, grad in jobs:
        if code_of(x) != code:
            raise B200Error("conv2d_wgrad_batch: mixed activation dtypes")
        n, h, w, cin, ldx = _nhwc(x)
        if mode == 1:
            n //= 4
        rows += [mode, x.data_ptr(), n, h, w, cin, ldx, dz.data_ptr(), cout, _nhwc(dz)[4], grad.data_ptr()]
    desc = (ctypes.c_longlong * len(rows))(*rows)
    call("b2_conv2d_wgrad_batch", len(jobs), ctypes.cast(desc, ctypes.c_void_p), code, stream())


def gemm_tn(a, b, m, ncols, k, lda, ldb, out, ldc, alpha=1.0, out_mode=0, batch=(1, 1), a_strides=(0, 0), b_strides=(0, 0),
            c_strides=(0, 0), code=None):
    code = code_of(a) if code is None else code
    call("b2_gemm_tn", ptr(a), lda, a_strides[0], a_strides[1], ptr(b), ldb, b_strides[0], b_strides[1], ptr(out), ldc,
         c_strides[0], c_strides[1], m, ncols, k, batch[0], batch[1], float(alpha), out_mode, code, stream())
    return out


def act(mode, a, z, out, dbias, rows, c, lda, ldz, ldo, code):
    call("b2_act", mode, ptr(a), lda, ptr(z), ldz, ptr(out), ldo, ptr(dbias), rows, c, code, stream())
    return out


def add(a, b, out):
    n, h, w, c, lda = _nhwc(a)
    call("b2_add", ptr(a), lda, ptr(b), _nhwc(b)[4], ptr(out), _nhwc(out)[4], n * h * w, c, code_of(a), stream())
    return out


def conv_first(x_nchw, w_kc, bias, cout, act, code):
    """First conv of the network on CUDA cores: fp32 NCHW image (3 or 6 channels) -> NHWC activation, bias (+ Swish)."""
    n, cin, h, w = x_nchw.shape
    y = new_act(n, h, w, cout, code, x_nchw.device)
    call("b2_conv3x3_first", ptr(x_nchw), ptr(w_kc), ptr(bias), ptr(y), cout, n, cin, h, w, cout, act, code, stream())
    return y


def conv_last(x, w_tc4, bias, cout, act, out_nchw):
    """Last conv of the network (<= 4 output channels) on CUDA cores: NHWC activation -> fp32 NCHW, bias (+ tanh)."""
    n, h, w, cin, ldx = _nhwc(x)
    call("b2_conv3x3_last", ptr(x), ldx, ptr(w_tc4), ptr(bias), ptr(out_nchw), n, h, w, cin, cout, act, code_of(x), stream())
    return out_nchw


def edge_first_ok(conv):
    return conv.weight.shape[1] in (3, 6) and conv.weight.shape[0] % 16 == 0 and conv.weight.shape[0] <= 512


def edge_last_ok(conv):
    """The CUDA-core last conv measured 0.73 ms vs 0.50 ms for the (column-padded) tensor-core path at batch 256 x 64 x 64,
    so the engines keep the tensor-core kernel unless SDM_B200_EDGE_LAST=1."""
    import os
    if os.environ.get("SDM_B200_EDGE_LAST", "0") != "1":
        return False
    return conv.weight.shape[0] <= 4 and conv.weight.shape[1] % 8 == 0 and conv.weight.shape[1] <= 1024
