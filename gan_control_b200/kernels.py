"""Thin ctypes binding of libb200gan.so (include/b200gan.h) on torch CUDA tensors.

Every function here launches hand-written sm_100a kernels and nothing else: there is no CPU
path and no PyTorch fallback -- a missing library or a non-CUDA tensor raises.  torch is used
only for device memory (outputs come from the caching allocator) and the current stream.

Layout convention at this level: activations are physical NHWC, i.e. contiguous
``(N, H, W, C)`` tensors; ``ops.py`` owns the logical-NCHW <-> NHWC views.
"""
import ctypes
import os

import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, 'libb200gan.so')
_lib = None

F32, BF16, F16 = 0, 1, 2
_c = ctypes
_vp, _i, _i64, _f = _c.c_void_p, _c.c_int, _c.c_int64, _c.c_float


class FcLayer(ctypes.Structure):
    _fields_ = [('w', _vp), ('bias', _vp), ('in_dim', _i), ('out_dim', _i), ('in_off', _i),
                ('out_off', _i), ('scale', _f), ('bias_mul', _f)]


class ConvEpilogue(ctypes.Structure):
    _fields_ = [('bias', _vp), ('rowscale', _vp), ('noise', _vp), ('noise_w', _vp), ('slope', _f), ('gain', _f),
                ('addend', _vp), ('gate', _vp)]


class FcLayerGrad(ctypes.Structure):
    _fields_ = [('gw', _vp), ('gb', _vp)]


_SIGNATURES = {
    'b200gan_version': ([], _i),
    'b200gan_last_error': ([], _c.c_char_p),
    'b200gan_launch_count': ([], _c.c_uint64),
    'b200gan_upfirdn2d': ([_vp, _vp, _vp, _i] + [_i] * 6 + [_i] * 7 + [_f, _vp], _i),
    'b200gan_upfirdn2d_act': ([_vp, _vp, _vp, _i] + [_i] * 6 + [_i] * 7 + [_f, _vp, _vp, _vp, _vp, _f, _f, _vp], _i),
    'b200gan_bias_act_fwd': ([_vp, _vp, _vp, _vp, _vp, _vp, _i, _i64, _i64, _i64, _i64, _f, _f, _vp], _i),
    'b200gan_bias_act_bwd': ([_vp, _vp, _vp, _vp, _i, _i64, _i64, _i64, _i64, _f, _f, _vp], _i),
    'b200gan_epilogue_bwd': ([_vp] * 10 + [_i, _i64, _i64, _i64, _f, _f, _vp], _i),
    'b200gan_reduce_nhwc': ([_vp, _vp, _vp, _vp, _vp, _i, _i64, _i64, _i64, _vp], _i),
    'b200gan_conv_fwd': ([_vp, _vp, _vp, _i] + [_i] * 13 + [_vp, _vp, _vp, _vp, _f, _f, _vp], _i),
    'b200gan_conv_fwd_packed': ([_vp, _vp, _vp, _i] + [_i] * 13 + [_vp, _vp, _vp, _vp, _f, _f, _vp], _i),
    'b200gan_conv_fwd_ex': ([_vp, _vp, _vp, _i] + [_i] * 15 + [_vp, _vp], _i),
    'b200gan_conv_wgrad_packed': ([_vp, _vp, _vp, _i] + [_i] * 13 + [_vp], _i),
    'b200gan_set_conv_engine': ([_i], _i),
    'b200gan_engine_launches': ([_i], _c.c_uint64),
    'b200gan_last_conv_engine': ([], _i),
    'b200gan_conv_wgrad': ([_vp, _vp, _vp, _i] + [_i] * 13 + [_vp], _i),
    'b200gan_modweight_fwd': ([_vp] * 5 + [_i] * 6 + [_f, _i, _i, _vp], _i),
    'b200gan_modweight_bwd': ([_vp] * 7 + [_i] * 5 + [_f, _i, _i, _vp], _i),
    'b200gan_linear_fwd': ([_vp, _vp, _vp, _vp, _i, _i, _i, _i, _f, _f, _i, _vp], _i),
    'b200gan_gemm_f32': ([_vp, _vp, _vp] + [_i] * 8 + [_f, _f, _vp], _i),
    'b200gan_mapping_fwd': ([_vp, _vp, _vp] + [_i] * 6 + [_vp], _i),
    'b200gan_mapping_bwd': ([_vp] * 6 + [_i] * 6 + [_vp], _i),
    'b200gan_adam_ema': ([_vp, _vp, _vp, _vp, _vp, _i64] + [_f] * 4 + [_vp, _f, _f, _vp], _i),
    'b200gan_affine_color_fwd': ([_vp, _vp, _vp, _vp, _i] + [_i] * 6 + [_vp, _vp, _vp], _i),
    'b200gan_affine_color_bwd': ([_vp, _vp, _vp, _vp, _i] + [_i] * 6 + [_vp, _vp], _i),
}


def lib():
    """Load libb200gan.so (built by `python -m gan_control_b200.build`); fail loudly if absent."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise RuntimeError(
                f'{LIB_PATH} is missing: the sm_100a kernels are the only implementation of this '
                f'package (no CPU / PyTorch fallback). Build it with `python -m gan_control_b200.build`.')
        handle = ctypes.CDLL(LIB_PATH)
        for name, (args, res) in _SIGNATURES.items():
            fn = getattr(handle, name)          # AttributeError if the library lacks a declared symbol
            fn.argtypes, fn.restype = args, res
        _lib = handle
    return _lib


def exported_symbols():
    return sorted(_SIGNATURES)


def set_conv_engine(engine):
    """0 = auto (tcgen05 when eligible), 1 = CUDA-core engine only. Returns the previous setting."""
    return int(lib().b200gan_set_conv_engine(int(engine)))


def launch_count():
    return int(lib().b200gan_launch_count())


ENGINES = ('fwd_simt', 'fwd_pointwise', 'fwd_umma', 'fwd_halo', 'wgrad_simt', 'wgrad_pointwise', 'wgrad_umma', 'wgrad_halo')


def engine_launches():
    """{engine name: convolution calls it served so far} (include/b200gan.h B200GAN_ENGINE_*); `fwd_umma`, `fwd_halo`,
    `wgrad_umma`, `wgrad_halo` are the tcgen05 kernels."""
    return {name: int(lib().b200gan_engine_launches(i)) for i, name in enumerate(ENGINES)}


def last_conv_engine():
    i = int(lib().b200gan_last_conv_engine())
    return None if i < 0 else ENGINES[i]


def _check(rc, what):
    if rc != 0:
        raise RuntimeError(f'libb200gan {what} failed (code {rc}): {lib().b200gan_last_error().decode()}')


def _dt(t):
    if t.dtype == torch.float32:
        return F32
    if t.dtype == torch.bfloat16:
        return BF16
    if t.dtype == torch.float16:
        return F16
    raise TypeError(f'libb200gan supports float32 / bfloat16 / float16 activations, got {t.dtype}')


def _cuda(*ts):
    for t in ts:
        if t is not None and not t.is_cuda:
            raise RuntimeError('gan_control_b200 kernels run on CUDA (sm_100a) tensors only; got a '
                               f'{t.device} tensor. There is no CPU fallback.')


def _ptr(t):
    return None if t is None else ctypes.c_void_p(t.data_ptr())


def _stream():
    return ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)


def _f32c(t):
    """small fp32 side inputs (bias, rowscale, taps) as contiguous fp32"""
    if t is None:
        return None
    return t.detach().to(torch.float32).contiguous()


# ---------------------------------------------------------------------------------------------
def upfirdn2d(x, taps, up, down, pad0_y, pad0_x, out_h, out_w, flip, gain=1.0, epilogue=None):
    """x: contiguous (N,H,W,C).  Returns contiguous (N,out_h,out_w,C).
    epilogue = (bias, rowscale, noise, noise_w, slope, act_gain): the StyledConv tail fused on the filter output."""
    _cuda(x, taps)
    assert x.ndim == 4 and x.is_contiguous()
    n, h, w, c = x.shape
    taps = _f32c(taps)
    y = torch.empty((n, out_h, out_w, c), dtype=x.dtype, device=x.device)
    if y.numel() == 0:
        return y
    with torch.cuda.device(x.device):
        if epilogue is None:
            _check(lib().b200gan_upfirdn2d(_ptr(x), _ptr(y), _ptr(taps), _dt(x), n, h, w, c, out_h, out_w,
                                           taps.shape[0], taps.shape[1], up, down, pad0_y, pad0_x, int(flip),
                                           float(gain), _stream()), 'upfirdn2d')
        else:
            bias, rowscale, noise, noise_w, slope, act_gain = epilogue
            _cuda(bias, rowscale, noise, noise_w)
            bias, rowscale, noise_w = _f32c(bias), _f32c(rowscale), _f32c(noise_w)
            if noise is not None:
                noise = noise.detach().to(x.dtype).contiguous()
                assert noise.numel() == n * out_h * out_w
            _check(lib().b200gan_upfirdn2d_act(_ptr(x), _ptr(y), _ptr(taps), _dt(x), n, h, w, c, out_h, out_w,
                                               taps.shape[0], taps.shape[1], up, down, pad0_y, pad0_x, int(flip),
                                               float(gain), _ptr(bias), _ptr(rowscale), _ptr(noise), _ptr(noise_w),
                                               float(slope), float(act_gain), _stream()), 'upfirdn2d_act')
    return y


def _ew_shape(x, planar):
    """(n, hw, c, inner) of a bias_act operand: NHWC (N,...,C) or planar NCHW (N,C,...)."""
    if planar:
        n, c = x.shape[0], x.shape[1]
        return n, 1, c, max(1, x.numel() // max(1, n * c))
    n, c = x.shape[0], x.shape[-1]
    return n, max(1, x.numel() // max(1, n * c)), c, 1


def bias_act_fwd(x, bias=None, rowscale=None, noise=None, noise_w=None, slope=0.2, gain=2 ** 0.5, planar=False):
    """y = gain*lrelu(x*rowscale[n,c] + noise_w*noise[n,pix] + bias[c]).  x contiguous; channel axis
    last (NHWC) or axis 1 (planar=True)."""
    _cuda(x, bias, rowscale, noise, noise_w)
    assert x.is_contiguous()
    n, hw, c, inner = _ew_shape(x, planar)
    bias, rowscale, noise_w = _f32c(bias), _f32c(rowscale), _f32c(noise_w)
    if noise is not None:
        noise = noise.detach().to(x.dtype).contiguous()
        assert noise.numel() == n * hw * inner
    y = torch.empty_like(x)
    if x.numel() == 0:
        return y
    with torch.cuda.device(x.device):
        _check(lib().b200gan_bias_act_fwd(_ptr(x), _ptr(y), _ptr(bias), _ptr(rowscale), _ptr(noise), _ptr(noise_w),
                                          _dt(x), n, hw, c, inner, float(slope), float(gain), _stream()), 'bias_act_fwd')
    return y


def bias_act_bwd(gy, y, rowscale=None, slope=0.2, gain=2 ** 0.5, planar=False):
    """gx = gy * gain * (y > 0 ? 1 : slope) * rowscale[n,c]."""
    _cuda(gy, y, rowscale)
    assert gy.is_contiguous() and y.is_contiguous() and gy.shape == y.shape and gy.dtype == y.dtype
    n, hw, c, inner = _ew_shape(gy, planar)
    rowscale = _f32c(rowscale)
    gx = torch.empty_like(gy)
    if gy.numel() == 0:
        return gx
    with torch.cuda.device(gy.device):
        _check(lib().b200gan_bias_act_bwd(_ptr(gy), _ptr(y), _ptr(gx), _ptr(rowscale), _dt(gy), n, hw, c, inner,
                                          float(slope), float(gain), _stream()), 'bias_act_bwd')
    return gx


def epilogue_bwd(gy, y, rowscale=None, noise=None, noise_w=None, bias=None, slope=0.2, gain=2 ** 0.5,
                 want_gd=True, want_gb=True, want_gnw=True):
    """Fused backward of y = gain*lrelu(z*rowscale + noise_w*noise + bias) from the saved output (NHWC).
    Returns (gconv, gd (N,C)|None, gb (C,)|None, gnw (1,)|None), reductions in fp32."""
    _cuda(gy, y, rowscale, noise, noise_w, bias)
    assert gy.is_contiguous() and y.is_contiguous() and gy.shape == y.shape and gy.dtype == y.dtype
    n, c = gy.shape[0], gy.shape[-1]
    hw = max(1, gy.numel() // max(1, n * c))
    rowscale, noise_w, bias = _f32c(rowscale), _f32c(noise_w), _f32c(bias)
    if noise is not None:
        noise = noise.detach().to(gy.dtype).contiguous()
    gconv = torch.empty_like(gy)
    dev = gy.device
    gd = torch.zeros(n, c, dtype=torch.float32, device=dev) if (want_gd and rowscale is not None) else None
    gb = torch.zeros(c, dtype=torch.float32, device=dev) if want_gb else None
    gnw = torch.zeros(1, dtype=torch.float32, device=dev) if (want_gnw and noise is not None) else None
    if gy.numel():
        with torch.cuda.device(dev):
            _check(lib().b200gan_epilogue_bwd(_ptr(gy), _ptr(y), _ptr(gconv), _ptr(rowscale), _ptr(noise), _ptr(noise_w),
                                              _ptr(bias), _ptr(gd), _ptr(gb), _ptr(gnw), _dt(gy), n, hw, c, float(slope),
                                              float(gain), _stream()), 'epilogue_bwd')
    return gconv, gd, gb, gnw


def reduce_nhwc(a, b=None, per_channel=True, per_sample_channel=False, pixw=None):
    """sums of a*b*pixw[n,pix] over pixels: (C,) and/or (N,C), fp32.  a, b contiguous (N,...,C)."""
    _cuda(a, b, pixw)
    if pixw is not None:
        pixw = pixw.detach().to(a.dtype).contiguous()
        assert pixw.numel() * a.shape[-1] == a.numel()
    assert a.is_contiguous() and (b is None or (b.is_contiguous() and b.shape == a.shape and b.dtype == a.dtype))
    n, c = a.shape[0], a.shape[-1]
    hw = max(1, a.numel() // max(1, n * c))
    out_c = torch.zeros(c, dtype=torch.float32, device=a.device) if per_channel else None
    out_nc = torch.zeros(n, c, dtype=torch.float32, device=a.device) if per_sample_channel else None
    if a.numel():
        with torch.cuda.device(a.device):
            _check(lib().b200gan_reduce_nhwc(_ptr(a), _ptr(b), _ptr(pixw), _ptr(out_c), _ptr(out_nc), _dt(a), n, hw, c,
                                             _stream()), 'reduce_nhwc')
    return out_c, out_nc


def conv_fwd(x, w, out_h, out_w, up=1, down=1, pad0=0, bias=None, rowscale=None, noise=None, noise_w=None,
             slope=1.0, gain=1.0, pack_in=False, pack_out=False, addend=None, gate=None):
    """x: (B,H,W,IC) contiguous; w: (Bw,KH,KW,OC,IC) contiguous, same dtype; -> (B,out_h,out_w,OC).
    pack_in / pack_out: the convolution runs between space-to-depth views (include/b200gan.h): x is then the plain
    (B,2H,2W,IC/4) tensor, and / or the result is the plain (B,2*out_h,2*out_w,OC/4) tensor; out_h, out_w and the
    channel counts of `w` are the LOGICAL (view) sizes.
    addend / gate: tensors of the OUTPUT's shape and dtype (include/b200gan.h b200gan_conv_epilogue): `addend` is added to
    the accumulator first; with `gate` (the saved output of the producer layer) the call is a data gradient that also
    applies that layer's activation backward: (acc + addend) * rowscale * gain * (gate > 0 ? 1 : slope)."""
    _cuda(x, w, bias, rowscale, noise, noise_w, addend, gate)
    assert x.ndim == 4 and w.ndim == 5 and x.is_contiguous() and w.is_contiguous() and x.dtype == w.dtype
    b, h, wd, ic = x.shape
    if pack_in:
        assert up == 1 and down == 1 and h % 2 == 0 and wd % 2 == 0
        h, wd, ic = h // 2, wd // 2, ic * 4
    bw, kh, kw, oc, ic2 = w.shape
    assert ic2 == ic and bw in (1, b), (x.shape, w.shape)
    bias, rowscale, noise_w = _f32c(bias), _f32c(rowscale), _f32c(noise_w)
    if pack_out:
        assert up == 1 and down == 1 and oc % 4 == 0
        y = torch.empty((b, 2 * out_h, 2 * out_w, oc // 4), dtype=x.dtype, device=x.device)
    else:
        y = torch.empty((b, out_h, out_w, oc), dtype=x.dtype, device=x.device)
    if noise is not None:
        noise = noise.detach().to(x.dtype).contiguous()
        assert noise.numel() == b * y.shape[1] * y.shape[2]
    if y.numel() == 0:
        return y
    if addend is not None or gate is not None:
        for t in (addend, gate):
            assert t is None or (t.shape == y.shape and t.dtype == y.dtype and t.is_contiguous()), 'side input != output shape'
        ep = ConvEpilogue(_ptr(bias), _ptr(rowscale), _ptr(noise), _ptr(noise_w), float(slope), float(gain), _ptr(addend), _ptr(gate))
        with torch.cuda.device(x.device):
            _check(lib().b200gan_conv_fwd_ex(_ptr(x), _ptr(w), _ptr(y), _dt(x), b, h, wd, ic, out_h, out_w, oc, kh, kw, up, down,
                                             pad0, int(bw > 1), int(pack_in), int(pack_out), ctypes.byref(ep), _stream()),
                   'conv_fwd_ex')
        return y
    with torch.cuda.device(x.device):
        if pack_in or pack_out:
            _check(lib().b200gan_conv_fwd_packed(_ptr(x), _ptr(w), _ptr(y), _dt(x), b, h, wd, ic, out_h, out_w, oc, kh, kw,
                                                 pad0, int(bw > 1), int(pack_in), int(pack_out),
                                                 _ptr(bias), _ptr(rowscale), _ptr(noise), _ptr(noise_w), float(slope),
                                                 float(gain), _stream()), 'conv_fwd_packed')
        else:
            _check(lib().b200gan_conv_fwd(_ptr(x), _ptr(w), _ptr(y), _dt(x), b, h, wd, ic, out_h, out_w, oc, kh, kw,
                                          up, down, pad0, int(bw > 1),
                                          _ptr(bias), _ptr(rowscale), _ptr(noise), _ptr(noise_w), float(slope),
                                          float(gain), _stream()), 'conv_fwd')
    return y


def conv_wgrad(x, gy, kh, kw, up=1, down=1, pad0=0, per_sample=False, pack_x=False, pack_gy=False):
    """x: (B,H,W,IC), gy: (B,OH,OW,OC) contiguous -> fp32 (Bw,KH,KW,OC,IC) (logical channel counts when an
    operand is read through its space-to-depth view, see conv_fwd)."""
    _cuda(x, gy)
    assert x.is_contiguous() and gy.is_contiguous() and x.dtype == gy.dtype
    b, h, wd, ic = x.shape
    b2, oh, ow, oc = gy.shape
    assert b2 == b
    if pack_x:
        assert h % 2 == 0 and wd % 2 == 0
        h, wd, ic = h // 2, wd // 2, ic * 4
    if pack_gy:
        assert oh % 2 == 0 and ow % 2 == 0
        oh, ow, oc = oh // 2, ow // 2, oc * 4
    gw = torch.zeros((b if per_sample else 1, kh, kw, oc, ic), dtype=torch.float32, device=x.device)
    if x.numel() and gy.numel():
        with torch.cuda.device(x.device):
            if pack_x or pack_gy:
                assert up == 1 and down == 1
                _check(lib().b200gan_conv_wgrad_packed(_ptr(x), _ptr(gy), _ptr(gw), _dt(x), b, h, wd, ic, oh, ow, oc, kh, kw,
                                                       pad0, int(per_sample), int(pack_x), int(pack_gy), _stream()),
                       'conv_wgrad_packed')
            else:
                _check(lib().b200gan_conv_wgrad(_ptr(x), _ptr(gy), _ptr(gw), _dt(x), b, h, wd, ic, oh, ow, oc, kh, kw, up,
                                                down, pad0, int(per_sample), _stream()), 'conv_wgrad')
    return gw


def modweight_fwd(weight, s, scale, demodulate, flip, dtype, want_adjoint=False):
    """ModulatedConv2d's per-sample weights in one kernel (include/b200gan.h): weight (OC,IC,KH,KW) fp32, s (B,IC) fp32 ->
    wk (B,KH,KW,OC,IC) `dtype` (K-major operand of conv_fwd), wk_adjoint (B,KH,KW,IC,OC) or None, d (B,OC) fp32 or None."""
    _cuda(weight, s)
    weight, s = _f32c(weight), _f32c(s)
    oc, ic, kh, kw = weight.shape
    b = s.shape[0]
    assert s.shape[1] == ic
    wk = torch.empty((b, kh, kw, oc, ic), dtype=dtype, device=s.device)
    wkt = torch.empty((b, kh, kw, ic, oc), dtype=dtype, device=s.device) if want_adjoint else None
    d = torch.empty((b, oc), dtype=torch.float32, device=s.device) if demodulate else None
    if b:
        with torch.cuda.device(s.device):
            _check(lib().b200gan_modweight_fwd(_ptr(weight), _ptr(s), _ptr(d), _ptr(wk), _ptr(wkt), _dt(wk), b, oc, ic, kh, kw,
                                               float(scale), int(bool(demodulate)), int(bool(flip)), _stream()), 'modweight_fwd')
    return wk, wkt, d


def modweight_bwd(g, weight, s, d, scale, demodulate, flip, want_gs=True, want_gw=True):
    """First-order backward of modweight_fwd: g (B,KH,KW,OC,IC) fp32 -> (gs (B,IC) | None, gweight (OC,IC,KH,KW) | None)."""
    _cuda(g, weight, s, d)
    assert g.dtype == torch.float32 and g.is_contiguous()
    weight, s, d = _f32c(weight), _f32c(s), _f32c(d)
    oc, ic, kh, kw = weight.shape
    b = s.shape[0]
    assert g.shape == (b, kh, kw, oc, ic)
    need_gs = want_gs or (want_gw and demodulate)
    gs = torch.zeros((b, ic), dtype=torch.float32, device=s.device) if need_gs else None
    e = torch.empty((b, oc), dtype=torch.float32, device=s.device) if demodulate else None
    gw = torch.empty((oc, ic, kh, kw), dtype=torch.float32, device=s.device) if want_gw else None
    if b:
        with torch.cuda.device(s.device):
            _check(lib().b200gan_modweight_bwd(_ptr(g), _ptr(weight), _ptr(s), _ptr(d), _ptr(gs), _ptr(e), _ptr(gw), b, oc, ic,
                                               kh, kw, float(scale), int(bool(demodulate)), int(bool(flip)), _stream()),
                   'modweight_bwd')
    elif gw is not None:
        gw.zero_()
    return (gs if want_gs else None), gw


def linear_fwd(x, w, bias, scale, bias_mul, act):
    """y = act(scale * x @ w.T + bias*bias_mul); x (M,K) fp32/bf16 contiguous, w (N,K) fp32."""
    _cuda(x, w, bias)
    assert x.ndim == 2 and x.is_contiguous()
    w, bias = _f32c(w), _f32c(bias)
    m, k = x.shape
    n = w.shape[0]
    assert w.shape[1] == k
    y = torch.empty((m, n), dtype=x.dtype, device=x.device)
    if m:
        with torch.cuda.device(x.device):
            _check(lib().b200gan_linear_fwd(_ptr(x), _ptr(w), _ptr(bias), _ptr(y), _dt(x), m, n, k, float(scale),
                                            float(bias_mul), int(act), _stream()), 'linear_fwd')
    return y


def gemm_f32(a, b, trans_a, trans_b, alpha=1.0):
    """alpha * op(a) @ op(b) in fp32; op = transpose when the flag is set. 2-D contiguous inputs."""
    _cuda(a, b)
    a, b = _f32c(a), _f32c(b)
    m, k = (a.shape[1], a.shape[0]) if trans_a else a.shape
    k2, n = (b.shape[1], b.shape[0]) if trans_b else b.shape
    assert k == k2, (a.shape, b.shape, trans_a, trans_b)
    c = torch.empty((m, n), dtype=torch.float32, device=a.device)
    if m and n:
        if k == 0:
            return c.zero_()
        with torch.cuda.device(a.device):
            # kernel convention: B(k,n) = b[n*ldb + k] when its trans_b flag = 1 ("NT")
            _check(lib().b200gan_gemm_f32(_ptr(a), _ptr(b), _ptr(c), m, n, k, a.shape[1], b.shape[1], n,
                                          int(trans_a), int(trans_b), float(alpha), 0.0, _stream()), 'gemm_f32')
    return c


def mapping_fwd(z, layer_table, n_groups, n_layers, row_width, normalize, linear=False):
    """Persistent mapping-network kernel.  z (B, z_dim) fp32; layer_table: uint8 CUDA tensor holding
    n_layers*n_groups `FcLayer` structs.  Returns acts (n_layers+1, B, row_width) fp32.
    linear: the layers have no activation (include/b200gan.h: flag bit 1) -- a batch of plain EqualLinear layers."""
    _cuda(z, layer_table)
    z = _f32c(z)
    batch, z_dim = z.shape
    acts = torch.zeros((n_layers + 1, batch, row_width), dtype=torch.float32, device=z.device)
    if batch:
        with torch.cuda.device(z.device):
            _check(lib().b200gan_mapping_fwd(_ptr(z), _ptr(acts), _ptr(layer_table), n_groups, n_layers, batch, z_dim,
                                             row_width, int(bool(normalize)) | (2 if linear else 0), _stream()), 'mapping_fwd')
    return acts


def mapping_bwd(z, acts, g_out, layer_table, grad_table, n_groups, n_layers, row_width, normalize, want_dz=False, linear=False):
    """Backward of mapping_fwd in one cooperative kernel.  g_out (B, out_width) = dL/d(acts[n_layers][:, :out_width]); the
    parameter gradients are WRITTEN where `grad_table` (uint8 CUDA tensor of n_layers*n_groups FcLayerGrad) points.
    Returns dz (B, z_dim) or None."""
    _cuda(z, acts, g_out, layer_table, grad_table)
    z = _f32c(z)
    batch, z_dim = z.shape
    gbuf = torch.zeros((2, batch, row_width), dtype=torch.float32, device=z.device)
    gbuf[0, :, :g_out.shape[1]] = g_out
    dz = torch.zeros_like(z) if want_dz else None          # columns no layer reads keep a zero gradient
    if batch:
        with torch.cuda.device(z.device):
            _check(lib().b200gan_mapping_bwd(_ptr(z), _ptr(acts), _ptr(gbuf), _ptr(layer_table), _ptr(grad_table), _ptr(dz),
                                             n_groups, n_layers, batch, z_dim, row_width,
                                             int(bool(normalize)) | (2 if linear else 0), _stream()), 'mapping_bwd')
    return dz


def adam_ema(p, g, m, v, ema, lr, beta1, beta2, eps, bias_corr, ema_decay=0.0, grad_scale=1.0):
    """In-place Adam step (torch.optim.Adam semantics) on flat fp32 buffers, fused with EMA.
    bias_corr: device float32[2] = (1 - beta1**t, 1 - beta2**t)."""
    _cuda(p, g, m, v, ema, bias_corr)
    for t in (p, g, m, v) + ((ema,) if ema is not None else ()):
        assert t.dtype == torch.float32 and t.is_contiguous() and t.numel() == p.numel()
    assert bias_corr.dtype == torch.float32 and bias_corr.numel() == 2 and bias_corr.is_contiguous()
    with torch.cuda.device(p.device):
        _check(lib().b200gan_adam_ema(_ptr(p), _ptr(g), _ptr(m), _ptr(v), _ptr(ema), p.numel(), float(lr), float(beta1),
                                      float(beta2), float(eps), _ptr(bias_corr), float(ema_decay),
                                      float(grad_scale), _stream()), 'adam_ema')


def _strides4(t):
    return (ctypes.c_int64 * 4)(*t.stride())


def affine_color_fwd(x, mat, color, out_h, out_w):
    """ADA warp + colour (include/b200gan.h): x (N,C,H,W) in ANY layout (strides are passed), C <= 4; mat (N,6) float64 source-
    pixel affine map; color (N,C,C+1) float32 or None.  Returns (N,C,out_h,out_w) in x's dtype and memory format."""
    _cuda(x, mat, color)
    assert x.ndim == 4 and x.shape[1] <= 4 and mat.dtype == torch.float64 and mat.shape == (x.shape[0], 6) and mat.is_contiguous()
    n, c, h, w = x.shape
    if color is not None:
        assert color.dtype == torch.float32 and color.shape == (n, c, c + 1) and color.is_contiguous()
    fmt = torch.channels_last if (c > 1 and x.is_contiguous(memory_format=torch.channels_last) and not x.is_contiguous()) \
        else torch.contiguous_format
    y = torch.empty((n, c, out_h, out_w), dtype=x.dtype, device=x.device, memory_format=fmt)
    if y.numel():
        with torch.cuda.device(x.device):
            _check(lib().b200gan_affine_color_fwd(_ptr(x), _ptr(y), _ptr(mat), _ptr(color), _dt(x), n, c, h, w, out_h, out_w,
                                                  _strides4(x), _strides4(y), _stream()), 'affine_color_fwd')
    return y


def affine_color_bwd(gy, mat, color, in_h, in_w):
    """adjoint of affine_color_fwd w.r.t. the image: gy (N,C,out_h,out_w) any layout -> fp32 planar (N,C,in_h,in_w)"""
    _cuda(gy, mat, color)
    n, c, oh, ow = gy.shape
    gx = torch.zeros((n, c, in_h, in_w), dtype=torch.float32, device=gy.device)
    if gy.numel():
        with torch.cuda.device(gy.device):
            _check(lib().b200gan_affine_color_bwd(_ptr(gy), _ptr(gx), _ptr(mat), _ptr(color), _dt(gy), n, c, in_h, in_w, oh, ow,
                                                  _strides4(gy), _stream()), 'affine_color_bwd')
    return gx
