"""CPU stand-ins for every entry point of `gan_control_b200/kernels.py` (TEST INFRASTRUCTURE).

Each function restates the *contract* of one libb200gan kernel (include/b200gan.h) with plain
torch-CPU arithmetic in the dtype it is given (fp64 allowed).  Two uses, both in `tests/` only:

* CPU suite: `tests/conftest.py::cpu_kernels` monkeypatches these over `kernels.*` so that the
  host-side autograd algebra of `ops.py` / `modules.py` (backward and double-backward formulas,
  layouts, geometry) is verified against the reference-generated goldens without a GPU;
* GPU suite: each CUDA kernel is compared against its stand-in on random inputs.

The product never imports this module.
"""
import torch
import torch.nn.functional as F

# every `gan_control_b200.kernels` entry point that has a stand-in here (what the CPU tests monkeypatch)
STAND_INS = ['upfirdn2d', 'bias_act_fwd', 'bias_act_bwd', 'epilogue_bwd', 'reduce_nhwc', 'conv_fwd', 'conv_wgrad', 'linear_fwd',
             'gemm_f32', 'adam_ema', 'launch_count', 'modweight_fwd', 'modweight_bwd', 'affine_color_fwd', 'affine_color_bwd']


def _gather_pad(z, pad0_y, pad0_x, need_h, need_w):
    """zero-extend / crop so that index 0 of the result is z-index -pad0 and the extent is need_*"""
    zh, zw = z.shape[-2:]
    z = F.pad(z, [max(pad0_x, 0), max(0, need_w - pad0_x - zw), max(pad0_y, 0), max(0, need_h - pad0_y - zh)])
    z = z[..., max(-pad0_y, 0):, max(-pad0_x, 0):]
    return z[..., :need_h, :need_w]


def _zero_upsample(x, up):
    if up == 1:
        return x
    n, c, h, w = x.shape
    z = x.new_zeros(n, c, (h - 1) * up + 1, (w - 1) * up + 1)
    z[:, :, ::up, ::up] = x
    return z


def upfirdn2d(x, taps, up, down, pad0_y, pad0_x, out_h, out_w, flip, gain=1.0, epilogue=None):
    xn = x.permute(0, 3, 1, 2)
    c = xn.shape[1]
    kh, kw = taps.shape
    z = _gather_pad(_zero_upsample(xn, up), pad0_y, pad0_x, (out_h - 1) * down + kh, (out_w - 1) * down + kw)
    f = (taps.flip(0, 1) if flip else taps).to(x.dtype) * gain
    y = F.conv2d(z, f[None, None].expand(c, 1, kh, kw), groups=c, stride=down)
    y = y.permute(0, 2, 3, 1).contiguous()
    if epilogue is not None:
        bias, rowscale, noise, noise_w, slope, act_gain = epilogue
        y = bias_act_fwd(y, bias, rowscale, noise, noise_w, slope, act_gain)
    return y


def _bcast(v, x, planar, kind):
    """reshape a per-channel (C,), per-sample-channel (N,C) or per-pixel (N,pix) side input"""
    n = x.shape[0]
    if planar:
        c = x.shape[1]
        tail = (1,) * (x.ndim - 2)
        if kind == 'c':
            return v.reshape((1, c) + tail)
        if kind == 'nc':
            return v.reshape((n, c) + tail)
        return v.reshape((n, 1) + tuple(x.shape[2:]))
    c = x.shape[-1]
    mid = (1,) * (x.ndim - 2)
    if kind == 'c':
        return v.reshape((1,) + mid + (c,))
    if kind == 'nc':
        return v.reshape((n,) + mid + (c,))
    return v.reshape((n,) + tuple(x.shape[1:-1]) + (1,))


def bias_act_fwd(x, bias=None, rowscale=None, noise=None, noise_w=None, slope=0.2, gain=2 ** 0.5, planar=False):
    v = x
    if rowscale is not None:
        v = v * _bcast(rowscale.to(x.dtype), x, planar, 'nc')
    if noise is not None:
        v = v + noise_w.to(x.dtype).reshape(()) * _bcast(noise.to(x.dtype).contiguous(), x, planar, 'pix')
    if bias is not None:
        v = v + _bcast(bias.to(x.dtype), x, planar, 'c')
    return (gain * torch.where(v > 0, v, v * slope)).contiguous()


def bias_act_bwd(gy, y, rowscale=None, slope=0.2, gain=2 ** 0.5, planar=False):
    g = gy * gain * torch.where(y > 0, torch.ones_like(y), torch.full_like(y, slope))
    if rowscale is not None:
        g = g * _bcast(rowscale.to(gy.dtype), gy, planar, 'nc')
    return g.contiguous()


def epilogue_bwd(gy, y, rowscale=None, noise=None, noise_w=None, bias=None, slope=0.2, gain=2 ** 0.5,
                 want_gd=True, want_gb=True, want_gnw=True):
    hi = torch.float64
    n, c = gy.shape[0], gy.shape[-1]
    g = gy.to(hi) * gain * torch.where(y > 0, 1.0, slope).to(hi)
    d = _bcast(rowscale.to(hi), gy, False, 'nc') if rowscale is not None else torch.ones((), dtype=hi)
    u = torch.where(y > 0, y.to(hi) / gain, y.to(hi) / (gain * slope))
    nz = noise_w.to(hi).reshape(()) * _bcast(noise.to(hi).contiguous(), gy, False, 'pix') if noise is not None else 0.0
    bv = _bcast(bias.to(hi), gy, False, 'c') if bias is not None else 0.0
    z = (u - nz - bv) / d
    out_t = torch.float64 if gy.dtype == torch.float64 else torch.float32
    gconv = (g * d).to(gy.dtype).contiguous()
    gd = (g * z).reshape(n, -1, c).sum(1).to(out_t) if (want_gd and rowscale is not None) else None
    gb = g.reshape(-1, c).sum(0).to(out_t) if want_gb else None
    gnw = None
    if want_gnw and noise is not None:
        gnw = (g * _bcast(noise.to(hi).contiguous(), gy, False, 'pix')).sum().reshape(1).to(out_t)
    return gconv, gd, gb, gnw


def reduce_nhwc(a, b=None, per_channel=True, per_sample_channel=False, pixw=None):
    v = a.double() if a.dtype != torch.float64 else a
    if b is not None:
        v = v * b.to(v.dtype)
    if pixw is not None:
        v = v * pixw.to(v.dtype).reshape(tuple(a.shape[:-1]) + (1,))
    n, c = a.shape[0], a.shape[-1]
    v = v.reshape(n, -1, c)
    out_t = torch.float64 if a.dtype == torch.float64 else torch.float32
    return (v.sum((0, 1)).to(out_t) if per_channel else None,
            v.sum(1).to(out_t) if per_sample_channel else None)


def _s2d(x):
    """plain NHWC (B,2H,2W,C) -> its space-to-depth view (B,H,W,4C), channel = (py*2+px)*C + c"""
    b, h2, w2, c = x.shape
    return x.reshape(b, h2 // 2, 2, w2 // 2, 2, c).permute(0, 1, 3, 2, 4, 5).reshape(b, h2 // 2, w2 // 2, 4 * c)


def _d2s(y):
    """(B,H,W,4C) -> plain NHWC (B,2H,2W,C): inverse of _s2d"""
    b, h, w, c4 = y.shape
    c = c4 // 4
    return y.reshape(b, h, w, 2, 2, c).permute(0, 1, 3, 2, 4, 5).reshape(b, 2 * h, 2 * w, c).contiguous()


def conv_fwd(x, w, out_h, out_w, up=1, down=1, pad0=0, bias=None, rowscale=None, noise=None, noise_w=None,
             slope=1.0, gain=1.0, pack_in=False, pack_out=False, addend=None, gate=None):
    if pack_in:
        x = _s2d(x)
    b = x.shape[0]
    bw, kh, kw, oc, ic = w.shape
    xn = x.permute(0, 3, 1, 2)
    z = _gather_pad(_zero_upsample(xn, up), pad0, pad0, (out_h - 1) * down + kh, (out_w - 1) * down + kw)
    wt = w.permute(0, 3, 4, 1, 2).contiguous()                         # (Bw, OC, IC, KH, KW)
    if bw == 1:
        y = F.conv2d(z, wt[0], stride=down)
    else:
        y = torch.cat([F.conv2d(z[i:i + 1], wt[i], stride=down) for i in range(b)], 0)
    y = y.permute(0, 2, 3, 1).contiguous()
    if pack_out:
        y = _d2s(y)
    if addend is not None:
        y = y + addend.to(y.dtype)
    if gate is not None:              # backward mode: the producer's activation gradient from its saved output
        return bias_act_bwd(y, gate.to(y.dtype), rowscale, slope, gain)
    if bias is not None or rowscale is not None or noise is not None or slope != 1.0 or gain != 1.0:
        y = bias_act_fwd(y, bias, rowscale, noise, noise_w, slope, gain)
    return y


def conv_wgrad(x, gy, kh, kw, up=1, down=1, pad0=0, per_sample=False, pack_x=False, pack_gy=False):
    if pack_x:
        x = _s2d(x)
    if pack_gy:
        gy = _s2d(gy)
    b, _, _, ic = x.shape
    oc = gy.shape[-1]
    with torch.enable_grad():
        w = x.new_zeros((b if per_sample else 1, kh, kw, oc, ic), requires_grad=True)
        y = conv_fwd(x.detach(), w, gy.shape[1], gy.shape[2], up, down, pad0)
        gw, = torch.autograd.grad(y, w, gy.detach())
    return gw if x.dtype == torch.float64 else gw.float()


def _modweight(weight, s, scale, demodulate, flip):
    """(B,OC,IC,KH,KW) per-sample weights and d, in the dtype of `weight` (gm.py:284-289)"""
    w = scale * weight.unsqueeze(0) * s.to(weight.dtype)[:, None, :, None, None]
    d = None
    if demodulate:
        d = torch.rsqrt(w.pow(2).sum((2, 3, 4)) + 1e-8)
        w = w * d[:, :, None, None, None]
    if flip:
        w = w.flip(3, 4)
    return w, d


def modweight_fwd(weight, s, scale, demodulate, flip, dtype, want_adjoint=False):
    w, d = _modweight(weight.detach(), s.detach(), scale, demodulate, flip)
    wk = w.permute(0, 3, 4, 1, 2).to(dtype).contiguous()
    wkt = w.flip(3, 4).transpose(1, 2).permute(0, 3, 4, 1, 2).to(dtype).contiguous() if want_adjoint else None
    return wk, wkt, d


def modweight_bwd(g, weight, s, d, scale, demodulate, flip, want_gs=True, want_gw=True):
    with torch.enable_grad():
        wq, sq = weight.detach().clone().requires_grad_(True), s.detach().clone().requires_grad_(True)
        w, _ = _modweight(wq, sq, scale, demodulate, flip)
        gw, gs = torch.autograd.grad(w.permute(0, 3, 4, 1, 2), (wq, sq), g.to(w.dtype))
    return (gs if want_gs else None), (gw if want_gw else None)


def linear_fwd(x, w, bias, scale, bias_mul, act):
    y = (x @ w.to(x.dtype).t()) * scale
    if bias is not None:
        y = y + bias.to(x.dtype) * bias_mul
    if act:
        y = (2 ** 0.5) * torch.where(y > 0, y, 0.2 * y)
    return y


def gemm_f32(a, b, trans_a, trans_b, alpha=1.0):
    a2 = a.t() if trans_a else a
    b2 = b.t() if trans_b else b
    return (alpha * (a2 @ b2)).contiguous()


def adam_ema(p, g, m, v, ema, lr, beta1, beta2, eps, bias_corr, ema_decay=0.0, grad_scale=1.0):
    gi = g * grad_scale
    m.mul_(beta1).add_(gi, alpha=1 - beta1)
    v.mul_(beta2).addcmul_(gi, gi, value=1 - beta2)
    bc1, bc2 = float(bias_corr[0]), float(bias_corr[1])
    p.addcdiv_(m, v.sqrt() / (bc2 ** 0.5) + eps, value=-lr / bc1)
    if ema is not None:
        ema.mul_(ema_decay).add_(p, alpha=1 - ema_decay)


def launch_count():
    return 0


def _affine_color(x, mat, color, out_h, out_w):
    """contract of b200gan_affine_color_fwd, differentiable in x: bilinear sample (zeros outside) of x at the source PIXEL
    coordinates mat @ (ox, oy, 1), then the per-sample colour matrix | offset"""
    n, c, h, w = x.shape
    dt = x.dtype if x.dtype == torch.float64 else torch.float32
    ox = torch.arange(out_w, dtype=torch.float64).view(1, 1, out_w)
    oy = torch.arange(out_h, dtype=torch.float64).view(1, out_h, 1)
    m = mat.double().view(n, 6, 1, 1)
    sx = m[:, 0] * ox + m[:, 1] * oy + m[:, 2]
    sy = m[:, 3] * ox + m[:, 4] * oy + m[:, 5]
    x0, y0 = torch.floor(sx), torch.floor(sy)
    fx, fy = (sx - x0).to(dt), (sy - y0).to(dt)
    xp = F.pad(x.to(dt), (1, 1, 1, 1))                              # one ring of zeros: every out-of-range tap lands in it
    out = 0
    for dy, dx, wt in [(0, 0, (1 - fx) * (1 - fy)), (0, 1, fx * (1 - fy)), (1, 0, (1 - fx) * fy), (1, 1, fx * fy)]:
        xi = (x0 + dx).clamp(-1, w).long() + 1
        yi = (y0 + dy).clamp(-1, h).long() + 1
        idx = (yi * (w + 2) + xi).view(n, 1, -1).expand(n, c, -1)
        out = out + xp.reshape(n, c, -1).gather(2, idx).view(n, c, out_h, out_w) * wt.unsqueeze(1)
    if color is not None:
        cm = color.to(dt)
        out = torch.einsum('noc,nchw->nohw', cm[:, :, :c], out) + cm[:, :, c].view(n, c, 1, 1)
    return out


def affine_color_fwd(x, mat, color, out_h, out_w):
    return _affine_color(x, mat, color, out_h, out_w).to(x.dtype)


def affine_color_bwd(gy, mat, color, in_h, in_w):
    n, c = gy.shape[:2]
    dt = torch.float64 if gy.dtype == torch.float64 else torch.float32
    with torch.enable_grad():                                     # (called from inside an autograd Function's forward)
        x = torch.zeros(n, c, in_h, in_w, dtype=dt, requires_grad=True)
        y = _affine_color(x, mat, color, gy.shape[2], gy.shape[3])
        gx, = torch.autograd.grad(y, x, gy.detach().to(dt))
    return gx
