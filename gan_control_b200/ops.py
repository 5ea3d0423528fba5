"""Differentiable StyleGAN2 operators on top of libb200gan (kernels.py).

Public functions keep the reference signatures (``gan_model.py:39-50`` and the upstream
``conv2d_gradfix`` names the README points at):

    upfirdn2d(input, kernel, up=1, down=1, pad=(0, 0))
    fused_leaky_relu(input, bias, negative_slope=0.2, scale=2 ** 0.5)
    conv2d(input, weight, bias=None, stride=1, padding=0)          # conv2d_gradfix.conv2d
    conv_transpose2d(input, weight, bias=None, stride=1, padding=0)

Every op is a ``torch.autograd.Function`` whose backward is written with the SAME small set of
Functions (a conv "gather", its weight gradient, upfirdn2d, the bias-act gradient, a GEMM), so the
set is closed under differentiation: R1 (``generator_trainer.py:713-719``) and path-length
(``:601-614``) double-backward run through the CUDA kernels too.

Tensors are logical NCHW like the reference; physically the kernels want NHWC, so 4-D activations
travel in ``torch.channels_last`` (a permuted view, no copy once the network is in that format).
"""
import math

import torch
from torch.autograd import Function

from . import kernels as K

SQRT2 = math.sqrt(2.0)


# ---------------------------------------------------------------------------------------------
# layout helpers
# ---------------------------------------------------------------------------------------------
def _nhwc(x):
    """logical (N,C,H,W) -> contiguous (N,H,W,C) (a view when x is channels_last)."""
    xp = x.permute(0, 2, 3, 1)
    return xp if xp.is_contiguous() else xp.contiguous()


def _nchw(y):
    """contiguous (N,H,W,C) -> logical (N,C,H,W) with channels_last strides."""
    return y.permute(0, 3, 1, 2)


def up32(t):
    """at least fp32: bf16 -> fp32; fp32 / fp64 (CPU algebra tests) unchanged"""
    return t if t.dtype in (torch.float32, torch.float64) else t.float()


def _is_nhwc(x):
    return x.ndim == 4 and x.permute(0, 2, 3, 1).is_contiguous()


# ---------------------------------------------------------------------------------------------
# regulariser backward mode
# ---------------------------------------------------------------------------------------------
_DATA_GRADS_ONLY = False


class data_grads_only:
    """Context for the regularisers' INNER backward: `autograd.grad(pred.sum(), real_img, create_graph=True)`
    (R1, generator_trainer.py:713-716) and `autograd.grad((img*noise).sum(), latents, create_graph=True)`
    (path length, :606-608) only ask for the gradient w.r.t. an INPUT, but a custom Function cannot see that: its
    needs_input_grad is fixed at forward time, so it would also compute (and record for double backward) every
    parameter gradient -- a full set of weight-gradient convolutions and bias / noise reductions that autograd.grad
    then throws away.  Inside this context the Functions skip gradients of quantities that depend on parameters
    only (shared convolution weights, biases, noise strengths); gradients that can reach the input -- activations,
    per-sample (style-modulated) weights, demodulation coefficients, styles -- are computed as usual.  The result
    of the enclosed autograd.grad call is unchanged.  A module-level flag: backward runs on autograd's own thread."""

    def __enter__(self):
        global _DATA_GRADS_ONLY
        self.prev, _DATA_GRADS_ONLY = _DATA_GRADS_ONLY, True

    def __exit__(self, *exc):
        global _DATA_GRADS_ONLY
        _DATA_GRADS_ONLY = self.prev


def _param_grads():
    return not _DATA_GRADS_ONLY


_FIRST_ORDER = False


class first_order:
    """Context for passes whose backward is FIRST-ORDER ONLY (the plain discriminator / generator steps,
    generator_trainer.py:407-436, 645-667): layers may then use `mod_conv`, whose weight (de)modulation, convolution,
    epilogue and all their gradients are single hand-written kernels but whose backward is not itself differentiable.
    The regularisation steps (R1, path length: `create_graph=True`) stay outside and run the closed op algebra.
    Without autograd (inference, the generator pass of the discriminator step) the fused ops are used automatically."""

    def __enter__(self):
        global _FIRST_ORDER
        self.prev, _FIRST_ORDER = _FIRST_ORDER, True

    def __exit__(self, *exc):
        global _FIRST_ORDER
        _FIRST_ORDER = self.prev


def fused_prep():
    return _FIRST_ORDER or not torch.is_grad_enabled()


# ---------------------------------------------------------------------------------------------
# per-(sample, channel) scale and dot product: the differentiable pieces of the epilogue backward
# ---------------------------------------------------------------------------------------------
class _RowScale(Function):
    """y[n,c,h,w] = x[n,c,h,w] * d[n,c] in one pass (d fp32); closed under differentiation with _RowDot."""

    @staticmethod
    def forward(ctx, x, d):
        ctx.save_for_backward(x, d)
        return _nchw(K.bias_act_fwd(_nhwc(x), None, d, None, None, 1.0, 1.0))

    @staticmethod
    def backward(ctx, g):
        x, d = ctx.saved_tensors
        gx = _RowScale.apply(g, d) if ctx.needs_input_grad[0] else None
        gd = _RowDot.apply(g, x).to(d.dtype) if ctx.needs_input_grad[1] else None
        return gx, gd


class _RowDot(Function):
    """out[n,c] = sum_{h,w} a[n,c,h,w] * b[n,c,h,w], accumulated and returned in fp32 (one pass over a and b)."""

    @staticmethod
    def forward(ctx, a, b):
        ctx.save_for_backward(a, b)
        return K.reduce_nhwc(_nhwc(a), _nhwc(b), per_channel=False, per_sample_channel=True)[1]

    @staticmethod
    def backward(ctx, g):
        a, b = ctx.saved_tensors
        ga = _RowScale.apply(b, g).to(a.dtype) if ctx.needs_input_grad[0] else None
        gb = _RowScale.apply(a, g).to(b.dtype) if ctx.needs_input_grad[1] else None
        return ga, gb


def _epilogue_grads_graph(gy, y, z, d, noise, noise_w, bias, need_d, need_noise, need_nw, need_b, slope, gain):
    """Differentiable (create_graph) backward of y = gain*lrelu(z*d + nw*noise + bias) given the recomputed z:
    returns (gconv, gd, gnoise, gnw, gb).  Full-size work goes through _BiasActGrad / _RowScale / _RowDot (one
    kernel each); parameter-only gradients are skipped inside `data_grads_only`."""
    gz = _BiasActGrad.apply(gy, y, slope, gain)
    gconv = gz if d is None else _RowScale.apply(gz, d)
    gd = gnoise = gnw = gb = None
    if d is not None and need_d:
        gd = _RowDot.apply(gz, z).to(d.dtype)
    if noise is not None and need_noise:
        gnoise = (up32(gz).sum(1, keepdim=True) * up32(noise_w)).to(noise.dtype)
    if _param_grads():
        if noise is not None and need_nw:
            gnw = (up32(gz).sum(1, keepdim=True) * up32(noise)).sum().reshape(noise_w.shape).to(noise_w.dtype)
        if bias is not None and need_b:
            gb = up32(gz).sum((0, 2, 3)).to(bias.dtype)
    return gconv, gd, gnoise, gnw, gb


# ---------------------------------------------------------------------------------------------
# upfirdn2d                                  (gan_model.py:45-50 / pytorch_upfirdn2d.py:9-51)
# ---------------------------------------------------------------------------------------------
class _UpFirDn2d(Function):
    @staticmethod
    def forward(ctx, x, taps, up, down, pad0_y, pad0_x, out_h, out_w, flip):
        n, c, h, w = x.shape
        ctx.cfg = (up, down, pad0_y, pad0_x, h, w, flip, taps.shape[0], taps.shape[1])
        ctx.save_for_backward(taps)
        if x.is_contiguous() and not (_is_nhwc(x) and c > 1):
            # planar NCHW == NHWC with one channel: no layout change, no copy
            y = K.upfirdn2d(x.reshape(n * c, h, w, 1), taps, up, down, pad0_y, pad0_x, out_h, out_w, flip)
            return y.view(n, c, out_h, out_w)
        return _nchw(K.upfirdn2d(_nhwc(x), taps, up, down, pad0_y, pad0_x, out_h, out_w, flip))

    @staticmethod
    def backward(ctx, gy):
        taps, = ctx.saved_tensors
        up, down, p0y, p0x, h, w, flip, kh, kw = ctx.cfg
        # adjoint = the same op with up<->down, reversed taps, pad0' = k-1-pad0
        gx = _UpFirDn2d.apply(gy, taps, down, up, kh - 1 - p0y, kw - 1 - p0x, h, w, not flip)
        return gx, None, None, None, None, None, None, None, None


def upfirdn2d(input, kernel, up=1, down=1, pad=(0, 0)):
    """Drop-in for ``gan_model.upfirdn2d`` (gm.py:45-50)."""
    n, c, h, w = input.shape
    kh, kw = kernel.shape
    out_h = (h * up + pad[0] + pad[1] - kh) // down + 1
    out_w = (w * up + pad[0] + pad[1] - kw) // down + 1
    return _UpFirDn2d.apply(input, kernel, up, down, pad[0], pad[0], out_h, out_w, True)


# ---------------------------------------------------------------------------------------------
# bias + leaky-ReLU                                                    (gan_model.py:25-41)
# ---------------------------------------------------------------------------------------------
def _planar(x):
    """bias_act operand layout: NHWC when the 4-D tensor is channels_last, else planar."""
    return not (_is_nhwc(x) and x.shape[1] > 1)


class _BiasActGrad(Function):
    """gx = gy * gain * (y > 0 ? 1 : slope) -- linear in gy, so it is its own derivative."""

    @staticmethod
    def forward(ctx, gy, y, slope, gain):
        ctx.cfg = (slope, gain)
        ctx.save_for_backward(y)
        if _planar(y):
            return K.bias_act_bwd(gy.contiguous(), y.contiguous(), None, slope, gain, planar=True)
        return _nchw(K.bias_act_bwd(_nhwc(gy), _nhwc(y), None, slope, gain))

    @staticmethod
    def backward(ctx, ggx):
        y, = ctx.saved_tensors
        return _BiasActGrad.apply(ggx, y, *ctx.cfg), None, None, None


class _BiasAct(Function):
    """y = gain * lrelu(x + bias[c]); channel axis = 1."""

    @staticmethod
    def forward(ctx, x, bias, slope, gain):
        ctx.cfg = (slope, gain)
        if _planar(x):
            y = K.bias_act_fwd(x.contiguous(), bias, slope=slope, gain=gain, planar=True)
        else:
            y = _nchw(K.bias_act_fwd(_nhwc(x), bias, slope=slope, gain=gain))
        ctx.save_for_backward(y)
        ctx.bias_dtype = None if bias is None else bias.dtype
        return y

    @staticmethod
    def backward(ctx, gy):
        y, = ctx.saved_tensors
        gx = _BiasActGrad.apply(gy, y, *ctx.cfg)
        gb = None
        if ctx.bias_dtype is not None and ctx.needs_input_grad[1] and _param_grads():
            dims = [d for d in range(gx.ndim) if d != 1]
            gb = up32(gx).sum(dims).to(ctx.bias_dtype)
        return gx, gb, None, None


def fused_leaky_relu(input, bias, negative_slope=0.2, scale=SQRT2):
    """Drop-in for ``gan_model.fused_leaky_relu`` (gm.py:39-41)."""
    return _BiasAct.apply(input, bias, negative_slope, scale)


class _ModEpilogue(Function):
    """StyledConv tail in one pass: y = gain*lrelu(x*d[b,c] + nw*noise[b,1,h,w] + bias[c])
    (demod scale of gm.py:288-289 applied to the activation, NoiseInjection gm.py:340-345,
    FusedLeakyReLU gm.py:32-35).  d / noise / bias may be None."""

    @staticmethod
    def forward(ctx, x, d, noise, noise_w, bias, slope, gain):
        ctx.cfg = (slope, gain)
        y = _nchw(K.bias_act_fwd(_nhwc(x), bias, d, noise, noise_w, slope, gain))
        ctx.save_for_backward(x, d, noise, noise_w, y)
        ctx.bias_dtype = None if bias is None else bias.dtype
        return y

    @staticmethod
    def backward(ctx, gy):
        x, d, noise, noise_w, y = ctx.saved_tensors
        slope, gain = ctx.cfg
        need = ctx.needs_input_grad
        gx = gd = gnoise = gnw = gb = None
        if torch.is_grad_enabled():
            # create_graph=True (R1 / path-length): differentiable kernels
            bias_like = None if ctx.bias_dtype is None else torch.empty(0, dtype=ctx.bias_dtype, device=gy.device)
            gx, gd, gnoise, gnw, gb = _epilogue_grads_graph(gy, y, x, d, noise, noise_w, bias_like, need[1], need[2],
                                                            need[3], need[4], slope, gain)
        else:
            gz = _BiasActGrad.apply(gy, y, slope, gain)          # d loss / d (pre-activation)
            # first-order only: ONE fused pass (kernels.epilogue_bwd) instead of four
            gconv, gd_, gb_, gnw_ = K.epilogue_bwd(_nhwc(gy), _nhwc(y), d, noise, noise_w if noise is not None else None,
                                                   None, slope, gain, want_gd=d is not None and need[1],
                                                   want_gb=ctx.bias_dtype is not None and need[4],
                                                   want_gnw=noise is not None and need[3])
            # (bias is not needed to form z*d here because x = z is saved: use the exact product)
            if need[0]:
                gx = _nchw(gconv)
            if d is not None and need[1]:
                gd = K.reduce_nhwc(_nhwc(gz), _nhwc(x), per_channel=False, per_sample_channel=True)[1].to(d.dtype)
            if noise is not None and need[2]:
                gnoise = (up32(gz).sum(1, keepdim=True) * up32(noise_w)).to(noise.dtype)
            if gnw_ is not None:
                gnw = gnw_.reshape(noise_w.shape).to(noise_w.dtype)
            if gb_ is not None:
                gb = gb_.to(ctx.bias_dtype)
        return gx, gd, gnoise, gnw, gb, None, None


def mod_epilogue(x, d=None, noise=None, noise_w=None, bias=None, slope=0.2, gain=SQRT2):
    return _ModEpilogue.apply(x, d, noise, noise_w, bias, slope, gain)


class _FirEpilogue(Function):
    """y = gain*lrelu(upfirdn2d(x)*d[b,c] + nw*noise + bias[c]) in ONE kernel: the upsampling StyledConv's
    Blur (gm.py:307) with its demodulation scale / NoiseInjection / FusedLeakyReLU tail applied to the filter
    output in registers, so the blurred tensor never reaches HBM before its activation.
    Backward: first order = fused epilogue backward (from the saved output) + the FIR adjoint; under
    create_graph the filter output is recomputed and the differentiable formulas are used."""

    @staticmethod
    def forward(ctx, x, taps, d, noise, noise_w, bias, up, down, pad0, out_h, out_w, slope, gain):
        y = _nchw(K.upfirdn2d(_nhwc(x), taps, up, down, pad0, pad0, out_h, out_w, True,
                              epilogue=(bias, d, noise, noise_w, slope, gain)))
        ctx.cfg = (up, down, pad0, x.shape[2], x.shape[3], out_h, out_w, slope, gain)
        ctx.save_for_backward(x, taps, d, noise, noise_w, bias, y)
        return y

    @staticmethod
    def backward(ctx, gy):
        x, taps, d, noise, noise_w, bias, y = ctx.saved_tensors
        up, down, pad0, h, wd, oh, ow, slope, gain = ctx.cfg
        need = ctx.needs_input_grad
        kh, kw = taps.shape
        gx = gd = gnoise = gnw = gb = None
        if torch.is_grad_enabled():
            z = None
            if d is not None and need[2]:
                z = _UpFirDn2d.apply(x, taps, up, down, pad0, pad0, oh, ow, True)      # recompute, differentiable
            gfir, gd, gnoise, gnw, gb = _epilogue_grads_graph(gy, y, z, d, noise, noise_w, bias, need[2], need[3],
                                                              need[4], need[5], slope, gain)
        else:
            gfir, gd_, gb_, gnw_ = K.epilogue_bwd(_nhwc(gy), _nhwc(y), d, noise, noise_w if noise is not None else None,
                                                  bias, slope, gain, want_gd=d is not None and need[2],
                                                  want_gb=bias is not None and need[5],
                                                  want_gnw=noise is not None and need[4])
            gfir = _nchw(gfir)
            if gd_ is not None:
                gd = gd_.to(d.dtype)
            if gb_ is not None:
                gb = gb_.to(bias.dtype)
            if gnw_ is not None:
                gnw = gnw_.reshape(noise_w.shape).to(noise_w.dtype)
            if noise is not None and need[3]:
                gz = _BiasActGrad.apply(gy, y, slope, gain)
                gnoise = (up32(gz).sum(1, keepdim=True) * up32(noise_w)).to(noise.dtype)
        if need[0]:
            gx = _UpFirDn2d.apply(gfir, taps, down, up, kh - 1 - pad0, kw - 1 - pad0, h, wd, False)
        return gx, None, gd, gnoise, gnw, gb, None, None, None, None, None, None, None


def fir_epilogue(x, kernel, pad, d=None, noise=None, noise_w=None, bias=None, up=1, down=1, slope=0.2, gain=SQRT2):
    """upfirdn2d(x, kernel, up, down, pad) followed by the StyledConv tail, fused (see _FirEpilogue)."""
    n, c, h, w = x.shape
    kh, kw = kernel.shape
    out_h = (h * up + pad[0] + pad[1] - kh) // down + 1
    out_w = (w * up + pad[0] + pad[1] - kw) // down + 1
    return _FirEpilogue.apply(x, kernel, d, noise, noise_w, bias, up, down, pad[0], out_h, out_w, slope, gain)


# ---------------------------------------------------------------------------------------------
# convolution family
# ---------------------------------------------------------------------------------------------
def _kernel_layout(w, dtype):
    """(Bw,OC,IC,KH,KW) parameter layout -> (Bw,KH,KW,OC,IC) K-major operand in `dtype`, contiguous.
    One strided-read / cast / dense-write kernel (`.to(dtype).contiguous()` on the permuted view is two)."""
    v = w.detach().permute(0, 3, 4, 1, 2)
    if v.dtype == dtype and v.is_contiguous():
        return v                                               # already a view of a K-major operand
    return torch.empty(v.shape, dtype=dtype, device=w.device).copy_(v)


def _flip_t(w):
    """weights of the adjoint convolution: taps reversed, OC <-> IC"""
    return w.flip(3, 4).transpose(1, 2)


class _ConvGather(Function):
    """y[b,o,oy,ox] = sum_{i,ky,kx} z[b,i,oy*down+ky-pad0,ox*down+kx-pad0] * w[wb,o,i,ky,kx]
    with z = x zero-upsampled by `up` (include/b200gan.h).  w: (Bw,OC,IC,KH,KW), Bw in {1,B}.
    pack_in / pack_out: the convolution runs between space-to-depth VIEWS of plain tensors (up = down = 1;
    include/b200gan.h "packed"): x is then the plain (B,IC/4,2H,2W) tensor and / or y the plain (B,OC/4,2OH,2OW)
    tensor, while w, out_h, out_w are in view terms.  The adjoint swaps the two flags."""

    @staticmethod
    def forward(ctx, x, w, up, down, pad0, out_h, out_w, pack_in=False, pack_out=False, param_weight=False):
        h, wd = (x.shape[2] // 2, x.shape[3] // 2) if pack_in else (x.shape[2], x.shape[3])
        ctx.cfg = (up, down, pad0, h, wd, pack_in, pack_out, param_weight)
        ctx.save_for_backward(x, w)
        y = K.conv_fwd(_nhwc(x), _kernel_layout(w, x.dtype), out_h, out_w, up, down, pad0, pack_in=pack_in,
                       pack_out=pack_out)
        return _nchw(y)

    @staticmethod
    def backward(ctx, gy):
        x, w = ctx.saved_tensors
        up, down, pad0, h, wd, pack_in, pack_out, param_weight = ctx.cfg
        kh, kw = w.shape[3], w.shape[4]
        gx = gw = None
        if ctx.needs_input_grad[0]:
            gx = _ConvGather.apply(gy, _flip_t(w), down, up, kh - 1 - pad0, h, wd, pack_out, pack_in, param_weight)
        # only a weight the CALLER declared parameter-only may be skipped (never inferred from its shape: at batch 1 a
        # style-modulated per-sample weight is (1,OC,IC,k,k) too, and its gradient carries the path latent -> style -> w)
        if ctx.needs_input_grad[1] and (not param_weight or _param_grads()):
            gw = _ConvWgrad.apply(x, gy, up, down, pad0, kh, kw, w.shape[0] > 1, pack_in, pack_out).to(w.dtype)
        return gx, gw, None, None, None, None, None, None, None, None


class _ConvWgrad(Function):
    """gw[wb,o,i,ky,kx] = sum_{b,oy,ox} gy[b,o,oy,ox] * z[b,i,oy*down+ky-pad0,ox*down+kx-pad0]
    (fp32 result; summed over the batch unless per_sample); pack_x / pack_gy as in _ConvGather."""

    @staticmethod
    def forward(ctx, x, gy, up, down, pad0, kh, kw, per_sample, pack_x=False, pack_gy=False):
        h, wd = (x.shape[2] // 2, x.shape[3] // 2) if pack_x else (x.shape[2], x.shape[3])
        oh, ow = (gy.shape[2] // 2, gy.shape[3] // 2) if pack_gy else (gy.shape[2], gy.shape[3])
        ctx.cfg = (up, down, pad0, kh, kw, h, wd, oh, ow, pack_x, pack_gy)
        ctx.save_for_backward(x, gy)
        gw = K.conv_wgrad(_nhwc(x), _nhwc(gy), kh, kw, up, down, pad0, per_sample, pack_x=pack_x, pack_gy=pack_gy)
        return gw.permute(0, 3, 4, 1, 2)                       # (Bw,OC,IC,KH,KW) view

    @staticmethod
    def backward(ctx, ggw):
        x, gy = ctx.saved_tensors
        up, down, pad0, kh, kw, h, wd, oh, ow, pack_x, pack_gy = ctx.cfg
        gx = ggy = None
        if ctx.needs_input_grad[0]:
            gx = _ConvGather.apply(gy, _flip_t(ggw), down, up, kh - 1 - pad0, h, wd, pack_gy, pack_x)
        if ctx.needs_input_grad[1]:
            ggy = _ConvGather.apply(x, ggw, up, down, pad0, oh, ow, pack_x, pack_gy)
        return gx, ggy, None, None, None, None, None, None, None, None


class _ConvEpilogue(Function):
    """y = gain*lrelu(conv(x, w)*d[b,o] + nw*noise + bias[o]) in ONE kernel (the convolution's epilogue;
    the conv output never reaches HBM).  StyledConv without upsampling (gm.py:402-408) and the
    discriminator's ConvLayer (EqualConv2d + FusedLeakyReLU, gm.py:872-888); with pack_out / pack_in also the
    upsampling StyledConv and the downsampling ConvLayer (composite FIR (*) conv weights, see composite_up/down).
    Backward: first order = fused epilogue backward (from the saved output) + the conv gradients;
    under create_graph the conv output is recomputed and the differentiable formulas are used."""

    @staticmethod
    def forward(ctx, x, w, d, noise, noise_w, bias, up, down, pad0, out_h, out_w, slope, gain, pack_in=False,
                pack_out=False, param_weight=False):
        y = _nchw(K.conv_fwd(_nhwc(x), _kernel_layout(w, x.dtype), out_h, out_w, up, down, pad0, bias, d, noise,
                             noise_w, slope, gain, pack_in=pack_in, pack_out=pack_out))
        h, wd = (x.shape[2] // 2, x.shape[3] // 2) if pack_in else (x.shape[2], x.shape[3])
        ctx.cfg = (up, down, pad0, h, wd, out_h, out_w, slope, gain, pack_in, pack_out, param_weight)
        ctx.save_for_backward(x, w, d, noise, noise_w, bias, y)
        return y

    @staticmethod
    def backward(ctx, gy):
        x, w, d, noise, noise_w, bias, y = ctx.saved_tensors
        up, down, pad0, h, wd, oh, ow, slope, gain, pack_in, pack_out, param_weight = ctx.cfg
        need = ctx.needs_input_grad
        kh, kw = w.shape[3], w.shape[4]
        gx = gw = gd = gnoise = gnw = gb = None
        if torch.is_grad_enabled():
            z = None
            if d is not None and need[2]:
                z = _ConvGather.apply(x, w, up, down, pad0, oh, ow, pack_in, pack_out, param_weight)   # recompute, differentiable
            gconv, gd, gnoise, gnw, gb = _epilogue_grads_graph(gy, y, z, d, noise, noise_w, bias, need[2], need[3],
                                                               need[4], need[5], slope, gain)
        else:
            gconv, gd_, gb_, gnw_ = K.epilogue_bwd(_nhwc(gy), _nhwc(y), d, noise, noise_w if noise is not None else None,
                                                   bias, slope, gain, want_gd=d is not None and need[2],
                                                   want_gb=bias is not None and need[5],
                                                   want_gnw=noise is not None and need[4])
            gconv = _nchw(gconv)
            if gd_ is not None:
                gd = gd_.to(d.dtype)
            if gb_ is not None:
                gb = gb_.to(bias.dtype)
            if gnw_ is not None:
                gnw = gnw_.reshape(noise_w.shape).to(noise_w.dtype)
            if noise is not None and need[3]:
                gz = _BiasActGrad.apply(gy, y, slope, gain)
                gnoise = (up32(gz).sum(1, keepdim=True) * up32(noise_w)).to(noise.dtype)
        if need[0]:
            gx = _ConvGather.apply(gconv, _flip_t(w), down, up, kh - 1 - pad0, h, wd, pack_out, pack_in, param_weight)
        if need[1] and (not param_weight or _param_grads()):      # skipped only when the caller declared it parameter-only
            gw = _ConvWgrad.apply(x, gconv, up, down, pad0, kh, kw, w.shape[0] > 1, pack_in, pack_out).to(w.dtype)
        return gx, gw, gd, gnoise, gnw, gb, None, None, None, None, None, None, None, None, None, None


_ONES = {}


def _ones_row(ic, like):
    key = (ic, like.device)
    if key not in _ONES:
        _ONES[key] = torch.ones(1, ic, dtype=torch.float32, device=like.device)
    return _ONES[key]


def _channel_sums(t):
    """sum over (n, h, w) of a contiguous NHWC tensor, fp32.  Narrow tensors (C < 16, the RGB side) are folded so that
    g consecutive pixels form g*C channels: the reduction kernel then runs with full lanes instead of C of 32."""
    n, h, w, c = t.shape
    hw = h * w
    if c >= 16:
        return K.reduce_nhwc(t, None, per_channel=True)[0]
    g = 32
    while hw % g:
        g //= 2
    return K.reduce_nhwc(t.reshape(n, hw // g, 1, g * c), None, per_channel=True)[0].view(g, c).sum(0)


class _ModConv(Function):
    """y = gain*lrelu(conv(x, w_eff)*d_ext[b,o] + nw*noise + bias[o]),  w_eff[b] = scale*weight*s[b]*demod[b]
    -- a whole ModulatedConv2d / StyledConv / ToRGB / ConvLayer with its weight path (gm.py:284-289, 152-160):
    forward = `modweight_fwd` (per-sample weights straight into the K-major operand, demodulation folded in) + ONE
    convolution kernel with the fused epilogue; backward = fused epilogue backward, data-gradient convolution on the
    adjoint operand written by the same forward kernel, weight-gradient convolution, `modweight_bwd` (style and
    parameter gradients incl. the demodulation terms).  FIRST ORDER ONLY (see `first_order`).
    s = None: shared weights (EqualConv2d);  demod folds rsqrt(sum w^2) into the weights (weight-modulated form);
    d_ext: demodulation applied to the OUTPUT instead (activation-modulated form, modules.ModulatedConv2d)."""

    @staticmethod
    def forward(ctx, x, s, weight, d_ext, noise, noise_w, bias, scale, demod, flip, up, down, pad0, out_h, out_w, slope, gain,
                defer_gate=False, in_gate=None):
        oc, ic, kh, kw = weight.shape[-4:]
        assert not defer_gate or (d_ext is None and noise is None), 'a deferred gate needs a plain bias + leaky-ReLU epilogue'
        ctx.gates = (defer_gate, in_gate)
        w4 = weight.reshape(oc, ic, kh, kw)
        sk = s if s is not None else _ones_row(ic, x)
        has_ep = bias is not None or d_ext is not None or noise is not None or slope != 1.0 or gain != 1.0
        wk, wkt, d = K.modweight_fwd(w4, sk, scale, demod, flip, x.dtype, want_adjoint=ctx.needs_input_grad[0])
        bias_k = None if bias is None else bias.reshape(-1)
        y = K.conv_fwd(_nhwc(x), wk, out_h, out_w, up, down, pad0, bias_k, d_ext, noise, noise_w, slope, gain)
        ctx.cfg = (scale, demod, flip, up, down, pad0, x.shape[2], x.shape[3], slope, gain, has_ep)
        ctx.save_for_backward(x, s, weight, d_ext, noise, noise_w, bias, y if has_ep else None, wkt, d)
        return _nchw(y)

    @staticmethod
    def backward(ctx, gy):
        if torch.is_grad_enabled():
            raise RuntimeError('mod_conv is first-order only: run create_graph=True passes (R1 / path length) outside '
                               '`ops.first_order()` so that the layers use the closed, twice-differentiable op algebra')
        x, s, weight, d_ext, noise, noise_w, bias, y, wkt, d = ctx.saved_tensors
        scale, demod, flip, up, down, pad0, h, wd, slope, gain, has_ep = ctx.cfg
        need = ctx.needs_input_grad
        oc, ic, kh, kw = weight.shape[-4:]
        gx = gs = gw = gd = gnoise = gnw = gb = None
        gyn = _nhwc(gy)
        defer_gate, in_gate = ctx.gates
        if has_ep and defer_gate:
            # the consumers of y applied this layer's activation backward in their data-gradient epilogues (`in_gate`):
            # gy IS the gradient of the convolution output; only the bias reduction is left
            gconv = gyn
            if bias is not None and need[6]:
                gb = K.reduce_nhwc(gconv, None, per_channel=True)[0].reshape(bias.shape).to(bias.dtype)
        elif has_ep and d_ext is None and noise is None and slope == 1.0 and gain == 1.0:
            # bias-only epilogue (ToRGB, gm.py:424-428): the gradient passes through unchanged, only the bias sum is left.
            # (The generic epilogue backward copied the 3-channel tensor with 3 of 32 lanes active: 1.3 ms per G step.)
            gconv = gyn
            if bias is not None and need[6]:
                gb = _channel_sums(gconv).reshape(bias.shape).to(bias.dtype)
        elif has_ep:
            gconv, gd_, gb_, gnw_ = K.epilogue_bwd(gyn, y, d_ext, noise, noise_w if noise is not None else None,
                                                   None if bias is None else bias.reshape(-1), slope, gain,
                                                   want_gd=d_ext is not None and need[3], want_gb=bias is not None and need[6],
                                                   want_gnw=noise is not None and need[5])
            if gd_ is not None:
                gd = gd_.to(d_ext.dtype)
            if gb_ is not None:
                gb = gb_.reshape(bias.shape).to(bias.dtype)
            if gnw_ is not None:
                gnw = gnw_.reshape(noise_w.shape).to(noise_w.dtype)
            if noise is not None and need[4]:
                gz = K.bias_act_bwd(gyn, y, None, slope, gain)
                gnoise = (up32(gz).sum(-1).unsqueeze(1) * up32(noise_w)).to(noise.dtype)
        else:
            gconv = gyn
        if need[0]:
            if in_gate is None:
                gx = _nchw(K.conv_fwd(gconv, wkt, h, wd, down, up, kh - 1 - pad0))
            else:   # x is the output of a deferred-gate layer: its activation backward rides in this epilogue
                gx = _nchw(K.conv_fwd(gconv, wkt, h, wd, down, up, kh - 1 - pad0, slope=in_gate[0], gain=in_gate[1], gate=_nhwc(x)))
        if (need[1] and s is not None) or need[2]:
            gwk = K.conv_wgrad(_nhwc(x), gconv, kh, kw, up, down, pad0, s is not None)
            sk = s if s is not None else _ones_row(ic, x)
            gs_, gw_ = K.modweight_bwd(gwk, weight.reshape(oc, ic, kh, kw), sk, d, scale, demod, flip,
                                       want_gs=need[1] and s is not None, want_gw=need[2])
            if gs_ is not None:
                gs = gs_.to(s.dtype)
            if gw_ is not None:
                gw = gw_.reshape(weight.shape).to(weight.dtype)
        return gx, gs, gw, gd, gnoise, gnw, gb, None, None, None, None, None, None, None, None, None, None, None, None


def mod_conv(x, s, weight, d_ext=None, noise=None, noise_w=None, bias=None, scale=1.0, demod=False, flip=False, up=1, down=1,
             pad0=0, out_hw=None, slope=1.0, gain=1.0, defer_gate=False, in_gate=None):
    """See _ModConv.  weight: the parameter, (OC,IC,KH,KW) or (1,OC,IC,KH,KW); s: (B,IC) styles or None.
    defer_gate: EVERY consumer of the result promises to apply this layer's leaky-ReLU backward itself (`in_gate=(slope,
    gain)` on its side, from the saved output = its own input), so this layer's backward receives the gradient of the
    convolution output directly and the separate activation-backward pass disappears."""
    if out_hw is None:
        kh, kw = weight.shape[-2:]
        zh, zw = (x.shape[2] - 1) * up + 1, (x.shape[3] - 1) * up + 1
        out_hw = ((zh + 2 * pad0 - kh) // down + 1, (zw + 2 * pad0 - kw) // down + 1)
    return _ModConv.apply(x, s, weight, d_ext, noise, noise_w, bias, float(scale), bool(demod), bool(flip), up, down, pad0,
                          out_hw[0], out_hw[1], float(slope), float(gain), bool(defer_gate), in_gate)


class _ResBlockFn(Function):
    """The discriminator's ResBlock (gm.py:893-922) as ONE autograd node with a hand-scheduled first-order backward:

        y1  = sqrt2*lrelu(conv3x3(x, W1) + b1)                                          conv1   (gm.py:901)
        y2  = lrelu(conv3x3_s2(blur(y1), W2) + b2)          [* sqrt2 / sqrt2]            conv2   (gm.py:902, 857-872)
        out = conv1x1_s2(blur(x), Ws) / sqrt2 + y2                                       skip + residual (gm.py:907-920)

    What the node buys over the per-layer graph: the residual sum rides in the skip convolution's epilogue (`addend`);
    in the backward pass the two gradient contributions to x are summed in conv1's data-gradient epilogue instead of by
    autograd, and if x itself is the output of a deferred-gate layer (from_rgb) its activation backward rides there too
    (measured at 32 ch @1024^2: +0.37 ms on the convolution against 0.46 + 0.62 ms for the two separate passes,
    profiles/r02_side_inputs.md).  FIRST ORDER ONLY (see `first_order`)."""

    @staticmethod
    def forward(ctx, x, w1, b1, w2, b2, ws, taps, toeplitz, cfg):
        scale1, scale2, scale_s, slope, gain, fuse_down, pad_blur2, pad_blur_s, in_gate = cfg
        xn = _nhwc(x)
        bsz, h, wd, c = xn.shape
        oc = w2.shape[0]
        need_x = ctx.needs_input_grad[0]
        ones_c = _ones_row(c, x)
        inv = 1.0 / SQRT2
        wk1, wk1t, _ = K.modweight_fwd(w1, ones_c, scale1, False, False, x.dtype, want_adjoint=need_x)
        y1 = K.conv_fwd(xn, wk1, h, wd, 1, 1, 1, b1, None, None, None, slope, gain)
        t = None
        if fuse_down:
            w2c = composite_down((w2.detach() * scale2).unsqueeze(0), toeplitz)        # (1, OC, 4C, 3, 3)
            wk2 = _kernel_layout(w2c, x.dtype)
            y2 = K.conv_fwd(y1, wk2, h // 2, wd // 2, 1, 1, 1, b2, None, None, None, slope, gain * inv, pack_in=True)
            wk2t = _kernel_layout(_flip_t(w2c), x.dtype)
        else:
            t = K.upfirdn2d(y1, taps, 1, 1, pad_blur2, pad_blur2, h + 1, wd + 1, True)
            wk2, wk2t, _ = K.modweight_fwd(w2, ones_c, scale2, False, False, x.dtype, want_adjoint=True)
            y2 = K.conv_fwd(t, wk2, h // 2, wd // 2, 1, 2, 0, b2, None, None, None, slope, gain * inv)
        xs = K.upfirdn2d(xn, taps, 1, 2, pad_blur_s, pad_blur_s, h // 2, wd // 2, True)
        wks, wkst, _ = K.modweight_fwd(ws, ones_c, scale_s * inv, False, False, x.dtype, want_adjoint=need_x)
        out = K.conv_fwd(xs, wks, h // 2, wd // 2, 1, 1, 0, addend=y2)
        ctx.cfg = cfg
        ctx.save_for_backward(x, w1, b1, w2, b2, ws, taps, toeplitz, y1, y2, xs, t, wk1t, wk2t, wkst)
        return _nchw(out)

    @staticmethod
    def backward(ctx, g_out):
        if torch.is_grad_enabled():
            raise RuntimeError('the fused ResBlock is first-order only (see ops.first_order)')
        x, w1, b1, w2, b2, ws, taps, toeplitz, y1, y2, xs, t, wk1t, wk2t, wkst = ctx.saved_tensors
        scale1, scale2, scale_s, slope, gain, fuse_down, pad_blur2, pad_blur_s, in_gate = ctx.cfg
        need = ctx.needs_input_grad
        xn, g = _nhwc(x), _nhwc(g_out)
        bsz, h, wd, c = xn.shape
        oc = w2.shape[0]
        inv = 1.0 / SQRT2
        ones_c = _ones_row(c, x)
        gx = gw1 = gb1 = gw2 = gb2 = gws = None
        # conv2's activation backward (its output feeds the residual sum, so nobody else could do it) + bias gradient
        gz2, _, gb2_, _ = K.epilogue_bwd(g, y2, None, None, None, b2, slope, gain * inv, want_gd=False, want_gb=need[4],
                                         want_gnw=False)
        if gb2_ is not None:
            gb2 = gb2_.to(b2.dtype)
        # skip branch: 1x1 data gradient, then the adjoint of the decimating FIR
        gskip = None
        if need[0]:
            gxs = K.conv_fwd(g, wkst, h // 2, wd // 2, 1, 1, 0)
            kh = taps.shape[0]
            gskip = K.upfirdn2d(gxs, taps, 2, 1, kh - 1 - pad_blur_s, kh - 1 - pad_blur_s, h, wd, False)
        if need[5]:
            gws = K.modweight_bwd(K.conv_wgrad(xs, g, 1, 1, 1, 1, 0, False), ws, ones_c, None, scale_s * inv, False, False,
                                  want_gs=False)[1].to(ws.dtype)
        # conv2: data gradient (carrying conv1's activation backward where one kernel produces it) and weight gradient
        if fuse_down:
            # (measured, profiles/r02_side_inputs.md: in the depth-to-space epilogue a gate costs +0.48 ms at 32 ch @1024^2,
            # more than the separate activation-backward pass, 0.63 ms incl. the bias reduction it would still need)
            gy1 = K.conv_fwd(gz2, wk2t, h // 2, wd // 2, 1, 1, 1, pack_out=True)
            gz1, _, gb1_, _ = K.epilogue_bwd(gy1, y1, None, None, None, b1, slope, gain, want_gd=False, want_gb=need[2],
                                             want_gnw=False)
            if gb1_ is not None:
                gb1 = gb1_.to(b1.dtype)
            if need[3]:
                gw2c = K.conv_wgrad(y1, gz2, 3, 3, 1, 1, 1, False, pack_x=True).permute(0, 3, 4, 1, 2)
                with torch.enable_grad():
                    wq = w2.detach().requires_grad_(True)
                    gw2, = torch.autograd.grad(composite_down((wq * scale2).unsqueeze(0), toeplitz), wq, gw2c.to(wq.dtype))
        else:
            kh = taps.shape[0]
            gt = K.conv_fwd(gz2, wk2t, h + 1, wd + 1, 2, 1, 2)                         # adjoint of the stride-2 convolution
            gy1 = K.upfirdn2d(gt, taps, 1, 1, kh - 1 - pad_blur2, kh - 1 - pad_blur2, h, wd, False)
            gz1, _, gb1_, _ = K.epilogue_bwd(gy1, y1, None, None, None, b1, slope, gain, want_gd=False, want_gb=need[2],
                                             want_gnw=False)
            if gb1_ is not None:
                gb1 = gb1_.to(b1.dtype)
            if need[3]:
                gw2 = K.modweight_bwd(K.conv_wgrad(t, gz2, 3, 3, 1, 2, 0, False), w2, ones_c, None, scale2, False, False,
                                      want_gs=False)[1].to(w2.dtype)
        # conv1: data gradient + the skip branch's contribution (+ the producer's activation backward), weight gradient
        if need[0]:
            if in_gate is None:
                gx = _nchw(K.conv_fwd(gz1, wk1t, h, wd, 1, 1, 1, addend=gskip))
            else:
                gx = _nchw(K.conv_fwd(gz1, wk1t, h, wd, 1, 1, 1, slope=in_gate[0], gain=in_gate[1], addend=gskip, gate=xn))
        if need[1]:
            gw1 = K.modweight_bwd(K.conv_wgrad(xn, gz1, 3, 3, 1, 1, 1, False), w1, ones_c, None, scale1, False, False,
                                  want_gs=False)[1].to(w1.dtype)
        return gx, gw1, gb1, gw2, gb2, gws, None, None, None


def res_block(x, w1, b1, w2, b2, ws, taps, toeplitz, scale1, scale2, scale_s, slope=0.2, gain=SQRT2, fuse_down=False,
              pad_blur2=2, pad_blur_s=1, in_gate=None):
    """See _ResBlockFn."""
    return _ResBlockFn.apply(x, w1, b1, w2, b2, ws, taps, toeplitz,
                             (float(scale1), float(scale2), float(scale_s), float(slope), float(gain), bool(fuse_down),
                              int(pad_blur2), int(pad_blur_s), in_gate))


def _view_hw(x, w, up, down, pad0, pack_in):
    h, wd = (x.shape[2] // 2, x.shape[3] // 2) if pack_in else (x.shape[2], x.shape[3])
    zh, zw = (h - 1) * up + 1, (wd - 1) * up + 1
    return (zh + 2 * pad0 - w.shape[3]) // down + 1, (zw + 2 * pad0 - w.shape[4]) // down + 1


def conv_epilogue(x, w, d=None, noise=None, noise_w=None, bias=None, up=1, down=1, pad0=0, out_hw=None,
                  slope=0.2, gain=SQRT2, pack_in=False, pack_out=False, param_weight=False):
    """conv_gather fused with the demod-scale / noise / bias / leaky-ReLU epilogue; `w` (Bw,OC,IC,KH,KW).
    param_weight: `w` is a function of parameters only (see conv_gather)."""
    if out_hw is None:
        out_hw = _view_hw(x, w, up, down, pad0, pack_in)
    return _ConvEpilogue.apply(x, w, d, noise, noise_w, bias, up, down, pad0, out_hw[0], out_hw[1], slope, gain,
                               pack_in, pack_out, param_weight)


def conv_gather(x, w, up=1, down=1, pad0=0, out_hw=None, pack_in=False, pack_out=False, param_weight=False):
    """General form; `w` (Bw,OC,IC,KH,KW).  Default output extent = 'valid' over the padded,
    zero-upsampled input with symmetric padding pad0 (in view terms when packed).
    param_weight=True declares that `w` depends on parameters only (an EqualConv2d weight, the shared weight of the
    activation-modulated form): inside `data_grads_only` its gradient is then skipped.  A style-modulated weight must
    keep the default -- its gradient is part of d(output)/d(latent) -- whatever its batch dimension is."""
    if out_hw is None:
        out_hw = _view_hw(x, w, up, down, pad0, pack_in)
    return _ConvGather.apply(x, w, up, down, pad0, out_hw[0], out_hw[1], pack_in, pack_out, param_weight)


# ---------------------------------------------------------------------------------------------
# composite FIR (*) convolution weights for the packed (space-to-depth view) convolutions
# ---------------------------------------------------------------------------------------------
def fir_toeplitz(kernel, flip):
    """(9, 36) matrix T with  full_conv2d(w3x3, g).reshape(36) == w3x3.reshape(9) @ T,  g = kernel (flipped when
    asked): the TRUE (not correlated) full 2-D convolution of a 3x3 weight with the 4x4 FIR."""
    g = kernel.flip(0, 1) if flip else kernel
    assert g.shape == (4, 4)
    t = g.new_zeros(3, 3, 6, 6)
    for a in range(3):
        for b in range(3):
            t[a, b, a:a + 4, b:b + 4] = g
    return t.reshape(9, 36)


def _composite6(w, toeplitz):
    bw, oc, ic = w.shape[:3]
    c = _Gemm.apply(w.reshape(-1, 9), toeplitz.to(w.dtype), False, False, 1.0)
    return c.reshape(bw, oc, ic, 3, 2, 3, 2)                   # 6x6 taps as (m_y, p_y, m_x, p_x), n = 2*m + p


def composite_up(w, toeplitz):
    """Transposed stride-2 3x3 convolution followed by the 4x4 FIR (gm.py:295-307) as ONE 3x3 stride-1 convolution
    whose output is read depth-to-space (four output phases):
        out[2y+py, 2x+px, o] = sum_{ky,kx,i} x[y+ky-1, x+kx-1, i] * C[o, i, 4-2ky+py, 4-2kx+px],   C = fir (*) w
    w (Bw,OC,IC,3,3) -> (Bw,4*OC,IC,3,3) with output channel (py*2+px)*OC + o; toeplitz = fir_toeplitz(fir, False)."""
    bw, oc, ic = w.shape[:3]
    c = _composite6(w, toeplitz).flip(3, 5)                    # m -> ky = 2 - m
    return c.permute(0, 4, 6, 1, 2, 3, 5).reshape(bw, 4 * oc, ic, 3, 3)


def composite_down(w, toeplitz):
    """4x4 FIR (pad 2,2) followed by a stride-2 3x3 convolution (gm.py:857-872) as ONE 3x3 stride-1 convolution on
    the space-to-depth view of the input:
        out[y, x, o] = sum_{ky,kx,py,px,i} x[2(y+ky-1)+py, 2(x+kx-1)+px, i] * C[o, i, 2ky+py, 2kx+px],  C = flip(fir) (*) w
    w (Bw,OC,IC,3,3) -> (Bw,OC,4*IC,3,3) with input channel (py*2+px)*IC + i; toeplitz = fir_toeplitz(fir, True)."""
    bw, oc, ic = w.shape[:3]
    c = _composite6(w, toeplitz)
    return c.permute(0, 1, 4, 6, 2, 3, 5).reshape(bw, oc, 4 * ic, 3, 3)


def _only_defaults(who, **kw):
    for k, (v, default) in kw.items():
        if v != default:
            raise NotImplementedError(f'{who}: {k}={v!r} is not used on the gan-control path and is not built')


def conv2d(input, weight, bias=None, stride=1, padding=0, dilation=1, groups=1):
    """``conv2d_gradfix.conv2d`` / ``F.conv2d`` (upstream signature; groups=1, dilation=1 only).  weight (OC,IC,KH,KW),
    or (B,OC,IC,KH,KW) for per-sample weights (the reference's groups=batch trick, gm.py:326-329)."""
    _only_defaults('conv2d', dilation=(dilation, 1), groups=(groups, 1))
    w = weight if weight.ndim == 5 else weight.unsqueeze(0)
    # upstream `no_weight_gradients` semantics for an ordinary (OC,IC,KH,KW) weight; per-sample weights are data
    y = conv_gather(input, w, 1, stride, padding, param_weight=weight.ndim == 4)
    if bias is not None:
        y = y + bias.view(1, -1, 1, 1).to(y.dtype)
    return y


def conv_transpose2d(input, weight, bias=None, stride=1, padding=0, output_padding=0, groups=1, dilation=1):
    """``conv2d_gradfix.conv_transpose2d`` / ``F.conv_transpose2d`` (upstream signature; groups=1, dilation=1,
    output_padding=0 only).  weight (IC,OC,KH,KW) like torch, or (B,IC,OC,KH,KW)."""
    _only_defaults('conv_transpose2d', output_padding=(output_padding, 0), groups=(groups, 1), dilation=(dilation, 1))
    w = weight if weight.ndim == 5 else weight.unsqueeze(0)
    w = w.flip(3, 4).transpose(1, 2)                            # -> gather form (Bw,OC,IC,KH,KW)
    kh, kw = w.shape[3], w.shape[4]
    h, wd = input.shape[2], input.shape[3]
    out_hw = ((h - 1) * stride - 2 * padding + kh, (wd - 1) * stride - 2 * padding + kw)
    y = _ConvGather.apply(input, w, stride, 1, kh - 1 - padding, out_hw[0], out_hw[1], False, False, weight.ndim == 4)
    if bias is not None:
        y = y + bias.view(1, -1, 1, 1).to(y.dtype)
    return y


# ---------------------------------------------------------------------------------------------
# ADA augmentation core                                   (trainers/non_leaking.py:341-357, 373-383)
# ---------------------------------------------------------------------------------------------
class _AffineColor(Function):
    """Bilinear warp by a per-sample affine map of pixel coordinates + per-sample colour matrix, one kernel each way
    (include/b200gan.h b200gan_affine_color_fwd / _bwd).  Linear in x, so the backward is the adjoint kernel and the
    double backward is the forward again; the matrices carry no gradient."""

    @staticmethod
    def forward(ctx, x, mat, color, out_h, out_w):
        ctx.save_for_backward(mat, color)
        ctx.cfg = (x.shape[2], x.shape[3], out_h, out_w, x.dtype)
        return K.affine_color_fwd(x, mat, color, out_h, out_w)

    @staticmethod
    def backward(ctx, gy):
        mat, color = ctx.saved_tensors
        h, w, oh, ow, dtype = ctx.cfg
        return _AffineColorAdjoint.apply(gy, mat, color, h, w).to(dtype), None, None, None, None


class _AffineColorAdjoint(Function):
    @staticmethod
    def forward(ctx, gy, mat, color, in_h, in_w):
        ctx.save_for_backward(mat, color)
        ctx.cfg = (gy.shape[2], gy.shape[3], gy.dtype)
        return K.affine_color_bwd(gy, mat, color, in_h, in_w)

    @staticmethod
    def backward(ctx, ggx):
        mat, color = ctx.saved_tensors
        oh, ow, dtype = ctx.cfg
        return _AffineColor.apply(ggx.to(dtype), mat, color, oh, ow), None, None, None, None


def affine_color(x, mat, color, out_hw):
    """x (N,C,H,W), C <= 4; mat (N,6) float64: output pixel (ox, oy) samples x at (m0 ox + m1 oy + m2, m3 ox + m4 oy + m5)
    bilinearly (zero outside); color (N,C,C+1) float32 rows (matrix | offset) or None."""
    return _AffineColor.apply(x, mat, color, out_hw[0], out_hw[1])


# ---------------------------------------------------------------------------------------------
# dense layers                                                         (gan_model.py:189-197)
# ---------------------------------------------------------------------------------------------
class _Gemm(Function):
    """alpha * op(a) @ op(b), fp32, closed under differentiation."""

    @staticmethod
    def forward(ctx, a, b, trans_a, trans_b, alpha):
        ctx.cfg = (trans_a, trans_b, alpha)
        ctx.save_for_backward(a, b)
        return K.gemm_f32(a.contiguous(), b.contiguous(), trans_a, trans_b, alpha)

    @staticmethod
    def backward(ctx, gc):
        a, b = ctx.saved_tensors
        ta, tb, alpha = ctx.cfg
        ga = gb = None
        if ctx.needs_input_grad[0]:
            # C = op(a) op(b):  d op(a) = gc op(b)^T
            ga = _Gemm.apply(b, gc, tb, True, alpha) if ta else _Gemm.apply(gc, b, False, not tb, alpha)
        if ctx.needs_input_grad[1]:
            gb = _Gemm.apply(gc, a, True, ta, alpha) if tb else _Gemm.apply(a, gc, not ta, False, alpha)
        return (None if ga is None else ga.to(a.dtype)), (None if gb is None else gb.to(b.dtype)), None, None, None


class _EqualLinear(Function):
    """y = act(scale * x @ w.T + bias * bias_mul) in one kernel; backward via _Gemm/_BiasActGrad."""

    @staticmethod
    def forward(ctx, x, w, bias, scale, bias_mul, act):
        y = K.linear_fwd(x.contiguous(), w, bias, scale, bias_mul, act)
        ctx.cfg = (scale, bias_mul, act)
        ctx.has_bias = bias is not None
        ctx.save_for_backward(x, w, y)
        return y

    @staticmethod
    def backward(ctx, gy):
        x, w, y = ctx.saved_tensors
        scale, bias_mul, act = ctx.cfg
        gz = _BiasActGrad.apply(gy, y, 0.2, SQRT2) if act else gy
        gz32 = up32(gz)
        gx = gw = gb = None
        if ctx.needs_input_grad[0]:
            gx = _Gemm.apply(gz32, w, False, False, scale).to(x.dtype)
        if ctx.needs_input_grad[1] and _param_grads():
            gw = _Gemm.apply(gz32, up32(x), True, False, scale).to(w.dtype)
        if ctx.has_bias and ctx.needs_input_grad[2] and _param_grads():
            gb = gz32.sum(0) * bias_mul
        return gx, gw, gb, None, None, None


def equal_linear(x, weight, bias, scale, lr_mul=1.0, activation=False):
    """``EqualLinear.forward`` (gm.py:189-197) on a 2-D input."""
    return _EqualLinear.apply(x, weight, bias, scale, lr_mul, bool(activation))


# ---------------------------------------------------------------------------------------------
# mapping network in one persistent kernel (forward / no-grad)       (gan_model.py:633-642, 489-502)
# ---------------------------------------------------------------------------------------------
def build_fc_table(groups, device):
    """groups: [(lo, hi, [EqualLinear, ...]), ...] in latent order.  Returns the device-resident layer table
    of `b200gan_mapping_fwd` (layer-major FcLayer structs) plus its geometry.  Layer 0 reads its slice
    [lo, hi) of z, the last layer writes its slice of w, hidden activations are packed group after group."""
    n_groups, n_layers = len(groups), len(groups[0][2])
    assert n_layers >= 1 and all(len(g[2]) == n_layers for g in groups)
    table = (K.FcLayer * (n_groups * n_layers))()
    keep, row_width, out_width, in_offs = [], 0, 0, None
    keep_params, scales = [], []
    for l in range(n_layers):
        off = 0
        for gi, (lo, hi, layers) in enumerate(groups):
            lin = layers[l]
            w, b = lin.weight.detach(), lin.bias.detach()
            assert w.is_contiguous() and b.is_contiguous() and w.dtype == torch.float32
            keep += [w, b]
            keep_params += [lin.weight, lin.bias]
            scales.append((float(lin.scale), float(lin.lr_mul)))
            e = table[l * n_groups + gi]
            e.w, e.bias = w.data_ptr(), b.data_ptr()
            e.in_dim, e.out_dim = w.shape[1], w.shape[0]
            e.in_off = lo if l == 0 else in_offs[gi]
            e.out_off = lo if l == n_layers - 1 else off
            e.scale, e.bias_mul = float(lin.scale), float(lin.lr_mul)
            off += w.shape[0]
            row_width = max(row_width, e.in_off + e.in_dim, e.out_off + e.out_dim)
            if l == n_layers - 1:
                out_width = max(out_width, e.out_off + e.out_dim)
        in_offs = [table[l * n_groups + gi].out_off for gi in range(n_groups)]
    raw = torch.frombuffer(bytearray(bytes(table)), dtype=torch.uint8).to(device)
    return {'table': raw, 'n_groups': n_groups, 'n_layers': n_layers, 'row_width': row_width, 'out_width': out_width,
            'key': tuple(t.data_ptr() for t in keep), 'keep': keep, 'keep_params': keep_params,
            'geom': {'n_layers': n_layers, 'slices': [(lo, hi) for lo, hi, _ in groups], 'scales': scales}}


def mapping_forward(tbl, z, normalize=True):
    """Whole mapping network / MultiFcStack / FcStack forward in ONE cooperative kernel (no autograd)."""
    acts = K.mapping_fwd(z, tbl['table'], tbl['n_groups'], tbl['n_layers'], tbl['row_width'], normalize,
                         linear=tbl.get('linear', False))
    return acts[tbl['n_layers'], :, :tbl['out_width']]


def build_linear_batch_table(items, in_width, device):
    """items: [(EqualLinear without activation, input column offset)]: a batch of independent linear layers evaluated by ONE
    launch of the mapping kernels (flag `linear`): layer j reads columns [off_j, off_j + in_dim_j) of the (B, in_width)
    input and writes its out_dim_j outputs after those of layer j - 1.  Same dictionary as `build_fc_table`."""
    table = (K.FcLayer * len(items))()
    keep, keep_params, scales, slices, off = [], [], [], [], 0
    for j, (lin, in_off) in enumerate(items):
        w, b = lin.weight.detach(), lin.bias.detach()
        assert w.is_contiguous() and b.is_contiguous() and w.dtype == torch.float32 and not lin.activation
        keep += [w, b]
        keep_params += [lin.weight, lin.bias]
        scales.append((float(lin.scale), float(lin.lr_mul)))
        slices.append((in_off, in_off + w.shape[1]))
        e = table[j]
        e.w, e.bias = w.data_ptr(), b.data_ptr()
        e.in_dim, e.out_dim, e.in_off, e.out_off = w.shape[1], w.shape[0], in_off, off
        e.scale, e.bias_mul = float(lin.scale), float(lin.lr_mul)
        off += w.shape[0]
    raw = torch.frombuffer(bytearray(bytes(table)), dtype=torch.uint8).to(device)
    return {'table': raw, 'n_groups': len(items), 'n_layers': 1, 'row_width': max(in_width, off), 'out_width': off,
            'key': tuple(t.data_ptr() for t in keep), 'keep': keep, 'keep_params': keep_params, 'linear': True,
            'widths': [lin.weight.shape[0] for lin, _ in items],
            'geom': {'n_layers': 1, 'slices': slices, 'scales': scales, 'linear': True}}


def _grad_table(tbl, device):
    """flat gradient buffer for every weight / bias of the table + the device array of pointers into it (built once per
    table: a CUDA-graph capture must not contain its H2D copy)"""
    if 'grad_table' not in tbl:
        params = tbl['keep']                                   # [w, b, w, b, ...] in table (layer-major) order
        flat = torch.zeros(sum(p.numel() for p in params), dtype=torch.float32, device=device)
        gt = (K.FcLayerGrad * (len(params) // 2))()
        views, off = [], 0
        for j in range(0, len(params), 2):
            w, b = params[j], params[j + 1]
            gt[j // 2].gw = flat.data_ptr() + 4 * off
            views.append((off, w.shape))
            off += w.numel()
            gt[j // 2].gb = flat.data_ptr() + 4 * off
            views.append((off, b.shape))
            off += b.numel()
        tbl['grad_flat'], tbl['grad_views'] = flat, views
        tbl['grad_table'] = torch.frombuffer(bytearray(bytes(gt)), dtype=torch.uint8).to(device)
    return tbl['grad_table']


def _mapping_reference(z, params, geom, normalize):
    """the same network with the differentiable per-layer ops (used only for a create_graph backward)"""
    outs, j = [], 0
    n_layers = geom['n_layers']
    cols = []
    for gi, (lo, hi) in enumerate(geom['slices']):
        x = z[:, lo:hi]
        if normalize:
            x = x * torch.rsqrt(torch.mean(x * x, dim=1, keepdim=True) + 1e-8)
        cols.append(x)
    for l in range(n_layers):
        for gi in range(len(cols)):
            w, b = params[2 * (l * len(cols) + gi)], params[2 * (l * len(cols) + gi) + 1]
            scale, lr_mul = geom['scales'][l * len(cols) + gi]
            cols[gi] = equal_linear(cols[gi], w, b, scale, lr_mul, not geom.get('linear', False))
    return torch.cat(cols, dim=1)


class _MappingFn(Function):
    """z -> w through `b200gan_mapping_fwd`, gradients through `b200gan_mapping_bwd`: the whole mapping network
    (gm.py:633-642), split-FC MultiFcStack (gm.py:489-502) or controller FcStack is one kernel each way under autograd
    (the per-layer path costs 8..56 launches forward and ~6 per layer backward)."""

    @staticmethod
    def forward(ctx, z, tbl, normalize, *params):
        acts = K.mapping_fwd(z, tbl['table'], tbl['n_groups'], tbl['n_layers'], tbl['row_width'], normalize,
                             linear=tbl.get('linear', False))
        ctx.tbl, ctx.normalize = tbl, normalize
        ctx.save_for_backward(z, acts, *params)
        return acts[tbl['n_layers'], :, :tbl['out_width']].clone()

    @staticmethod
    def backward(ctx, g):
        z, acts, *params = ctx.saved_tensors
        tbl, normalize = ctx.tbl, ctx.normalize
        need_z = ctx.needs_input_grad[0]
        if torch.is_grad_enabled():
            # double backward: recompute with the closed op algebra (never hit by the training steps: the path-length
            # regulariser differentiates w.r.t. the latents, i.e. ABOVE the mapping network)
            with torch.enable_grad():
                w = _mapping_reference(z, params, tbl['geom'], normalize)
                ins = [t for t in [z] + list(params) if t.requires_grad]
                grads = iter(torch.autograd.grad(w, ins, g, create_graph=True, allow_unused=True))
            return tuple([next(grads) if z.requires_grad else None, None, None] +
                         [next(grads) if p.requires_grad else None for p in params])
        gt = _grad_table(tbl, z.device)
        dz = K.mapping_bwd(z, acts, g.contiguous().float(), tbl['table'], gt, tbl['n_groups'], tbl['n_layers'], tbl['row_width'],
                           normalize, want_dz=need_z, linear=tbl.get('linear', False))
        flat = tbl['grad_flat'].clone()        # the shared buffer is overwritten by the next backward of this table
        outs = []
        for i, (off, shape) in enumerate(tbl['grad_views']):
            outs.append(flat[off:off + math.prod(shape)].view(shape) if ctx.needs_input_grad[3 + i] else None)
        return (dz, None, None) + tuple(outs)


def mapping_apply(tbl, z, normalize=True):
    """differentiable `mapping_forward`"""
    return _MappingFn.apply(z, tbl, normalize, *tbl['keep_params'])
