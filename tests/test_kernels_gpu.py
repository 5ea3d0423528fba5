"""Each libb200gan kernel against its contract stand-in (oracle/kernels_ref.py, evaluated in fp64
on the CPU) on seeded random inputs: fp32 storage to ~1e-5, bf16 storage to bf16 rounding."""
import numpy as np
import pytest
import torch

from gan_control_b200 import kernels as K
from oracle import kernels_ref as R

pytestmark = pytest.mark.gpu
DTYPES = [torch.float32, torch.bfloat16]
DTYPES16 = [torch.float32, torch.bfloat16, torch.float16]


def rnd(seed, *shape):
    return torch.from_numpy(np.random.default_rng(seed).standard_normal(shape))


def tol(dt):
    # references are evaluated on the exact (already rounded) device inputs, so a 16-bit result differs from them by its
    # own output rounding (2^-9 for bf16, 2^-11 for fp16) plus fp32 accumulation noise
    return {torch.float32: 2e-5, torch.bfloat16: 4e-3, torch.float16: 1e-3}[dt]


TCGEN05 = {'fwd_umma', 'fwd_halo', 'wgrad_umma', 'wgrad_halo'}


class engines:
    """`with engines('fwd_halo'):` -- every convolution call inside must have been served by one of the named engines
    (libb200gan's per-engine counters, include/b200gan.h): a silent fallback to the CUDA-core kernel fails the test."""

    def __init__(self, *allowed):
        self.allowed = set(allowed)

    def __enter__(self):
        self.before = K.engine_launches()
        return self

    def __exit__(self, *exc):
        if exc[0] is not None:
            return False
        after = K.engine_launches()
        used = {k: after[k] - self.before[k] for k in after if after[k] != self.before[k]}
        assert used, 'no convolution was launched'
        assert set(used) <= self.allowed, f'expected engines {sorted(self.allowed)}, ran {used}'
        return False


def prep(t, dt):
    """device tensor in dt and the fp64 CPU value it actually holds"""
    d = t.to(dt).cuda()
    return d, d.cpu().double()


def close(out, ref, dt, what=''):
    err = float((out.detach().cpu().double() - ref).abs().max() / ref.abs().max().clamp_min(1e-30))
    assert err < tol(dt), f'{what}: rel err {err:.3e}'


@pytest.mark.parametrize('dt', DTYPES)
@pytest.mark.parametrize('cfg', [
    # n, h, w, c, kh, up, down, pad0, out_h, out_w, flip
    (2, 9, 9, 8, 4, 1, 1, 1, 8, 8, True), (2, 6, 7, 3, 4, 2, 1, 2, 12, 14, True), (3, 8, 8, 16, 4, 1, 1, 2, 9, 9, False),
    (2, 21, 21, 1, 12, 1, 2, 0, 5, 5, True), (1, 10, 10, 4, 12, 2, 1, 0, 9, 9, True), (2, 9, 7, 5, 4, 1, 1, -1, 7, 5, True),
    (2, 33, 33, 32, 4, 1, 1, 1, 32, 32, True), (2, 5, 5, 24, 4, 1, 2, 1, 2, 2, False),
    (2, 37, 41, 64, 4, 1, 1, 2, 38, 42, False), (1, 20, 50, 128, 3, 1, 1, 1, 20, 50, True), (2, 65, 65, 96, 4, 1, 1, 1, 64, 64, True),
    (1, 130, 70, 32, 4, 1, 1, 1, 129, 69, True),
    # decimating / interpolating TMA kernels (ResBlock.skip and its adjoint): even / odd pads, ragged edges, 3 taps
    (2, 64, 64, 32, 4, 1, 2, 1, 32, 32, True), (1, 66, 70, 64, 4, 1, 2, 1, 33, 35, False), (2, 40, 40, 96, 4, 1, 2, 2, 21, 21, True),
    (1, 37, 45, 32, 3, 1, 2, 0, 18, 22, True), (2, 32, 32, 32, 4, 2, 1, 2, 64, 64, False), (1, 33, 35, 64, 4, 2, 1, 2, 66, 70, True),
    (2, 20, 24, 128, 4, 2, 1, 1, 38, 46, False), (1, 17, 19, 32, 3, 2, 1, 1, 33, 37, True), (1, 16, 16, 32, 4, 2, 1, 3, 35, 35, True),
])
def test_upfirdn2d(cfg, dt):
    n, h, w, c, k, up, down, pad0, oh, ow, flip = cfg
    x, xr = prep(rnd(1, n, h, w, c), dt)
    taps = rnd(2, k, k).float()
    y = K.upfirdn2d(x, taps.cuda(), up, down, pad0, pad0, oh, ow, flip, 1.5)
    close(y, R.upfirdn2d(xr, taps.double(), up, down, pad0, pad0, oh, ow, flip, 1.5), dt)


@pytest.mark.parametrize('dt', DTYPES)
@pytest.mark.parametrize('cfg', [
    # n, h, w, c, kh, up, down, pad0, out_h, out_w   (generic, blur_rows and TMA kernels; decimating / interpolating)
    (2, 9, 9, 8, 4, 1, 1, 1, 8, 8), (2, 33, 33, 32, 4, 1, 1, 1, 32, 32), (2, 65, 65, 64, 4, 1, 1, 1, 64, 64),
    (1, 129, 69, 96, 4, 1, 1, 1, 128, 68), (2, 16, 16, 24, 4, 1, 2, 1, 8, 8), (2, 8, 8, 16, 4, 2, 1, 2, 16, 16),
    (3, 5, 5, 5, 4, 1, 1, 1, 4, 4),
])
def test_upfirdn2d_act(cfg, dt):
    """FIR with the StyledConv tail fused on its output == FIR stand-in followed by the bias-act stand-in"""
    n, h, w, c, k, up, down, pad0, oh, ow = cfg
    x, xr = prep(rnd(1, n, h, w, c), dt)
    taps = rnd(2, k, k).float()
    bias, rs, nw = rnd(4, c).float(), (rnd(5, n, c).abs() + 0.5).float(), torch.tensor([0.7])
    noise, noiser = prep(rnd(6, n, oh * ow), dt)
    y = K.upfirdn2d(x, taps.cuda(), up, down, pad0, pad0, oh, ow, True, 1.5,
                    epilogue=(bias.cuda(), rs.cuda(), noise, nw.cuda(), 0.2, 2 ** 0.5))
    yr = R.upfirdn2d(xr, taps.double(), up, down, pad0, pad0, oh, ow, True, 1.5,
                     epilogue=(bias.double(), rs.double(), noiser, nw.double(), 0.2, 2 ** 0.5))
    close(y, yr, dt)
    y2 = K.upfirdn2d(x, taps.cuda(), up, down, pad0, pad0, oh, ow, True, 1.5, epilogue=(None, None, None, None, 0.2, 2.0))
    close(y2, R.upfirdn2d(xr, taps.double(), up, down, pad0, pad0, oh, ow, True, 1.5, epilogue=(None, None, None, None, 0.2, 2.0)), dt)


@pytest.mark.parametrize('dt', DTYPES)
@pytest.mark.parametrize('shape,planar', [((3, 5, 7, 16), False), ((2, 4, 4, 3), False), ((2, 5, 6, 6), True), ((4, 24), True)])
def test_bias_act(shape, planar, dt):
    x, xr = prep(rnd(3, *shape), dt)
    c = shape[1] if planar else shape[-1]
    n = shape[0]
    bias, rs, nw = rnd(4, c).float(), (rnd(5, n, c).abs() + 0.5).float(), torch.tensor([0.7])
    npix = int(np.prod(shape)) // (n * c)
    noise, noiser = prep(rnd(6, n, npix), dt)
    y = K.bias_act_fwd(x, bias.cuda(), rs.cuda(), noise, nw.cuda(), 0.2, 2 ** 0.5, planar)
    yr = R.bias_act_fwd(xr, bias.double(), rs.double(), noiser, nw.double(), 0.2, 2 ** 0.5, planar)
    close(y, yr, dt, 'fwd')
    gy, gyr = prep(rnd(7, *shape), dt)
    ysaved = y.cpu().double()
    close(K.bias_act_bwd(gy, y, rs.cuda(), 0.2, 2 ** 0.5, planar), R.bias_act_bwd(gyr, ysaved, rs.double(), 0.2, 2 ** 0.5, planar), dt, 'bwd')
    close(K.bias_act_fwd(x, None, None, None, None, 1.0, 1.0, planar), xr, dt, 'identity')


@pytest.mark.parametrize('dt', DTYPES)
@pytest.mark.parametrize('shape', [(3, 9, 9, 16), (2, 64, 64, 40), (1, 3, 3, 513), (16, 4, 4, 512)])
def test_reduce_nhwc(shape, dt):
    a, ar = prep(rnd(8, *shape), dt)
    b, br = prep(rnd(9, *shape), dt)
    pw, pwr = prep(rnd(10, *shape[:-1]), dt)
    oc, onc = K.reduce_nhwc(a, b, True, True)
    rc, rnc = R.reduce_nhwc(ar, br, True, True)
    scale = float(rnc.abs().max())
    assert float((oc.cpu().double() - rc).abs().max()) < 1e-4 * max(scale, float(rc.abs().max()))
    assert float((onc.cpu().double() - rnc).abs().max()) < 1e-4 * scale
    oc2, _ = K.reduce_nhwc(a, None, True, False, pixw=pw)
    rc2, _ = R.reduce_nhwc(ar, None, True, False, pixw=pwr)
    assert float((oc2.cpu().double() - rc2).abs().max()) < 1e-4 * float(rc2.abs().max() + 1)


@pytest.mark.parametrize('dt', DTYPES)
@pytest.mark.parametrize('shape', [(3, 9, 9, 16), (2, 32, 32, 64), (16, 4, 4, 512)])
def test_epilogue_bwd(shape, dt):
    n, h, w, c = shape
    gy, gyr = prep(rnd(60, *shape), dt)
    y, yr = prep(rnd(61, *shape), dt)
    d, bias, nw = (rnd(62, n, c).abs() + 0.5).float(), rnd(63, c).float(), torch.tensor([0.4])
    noise, noiser = prep(rnd(64, n, h, w), dt)
    out = K.epilogue_bwd(gy, y, d.cuda(), noise, nw.cuda(), bias.cuda(), 0.2, 2 ** 0.5)
    ref = R.epilogue_bwd(gyr, yr, d.double(), noiser, nw.double(), bias.double(), 0.2, 2 ** 0.5)
    close(out[0], ref[0].double(), dt, 'gconv')
    for o, r_, name in zip(out[1:], ref[1:], ['gd', 'gb', 'gnw']):
        err = float((o.cpu().double() - r_.double()).abs().max() / r_.double().abs().max().clamp_min(1e-12))
        assert err < 2e-4, f'{name}: {err:.2e}'
    out2 = K.epilogue_bwd(gy, y, None, None, None, bias.cuda(), 0.2, 2 ** 0.5)
    ref2 = R.epilogue_bwd(gyr, yr, None, None, None, bias.double(), 0.2, 2 ** 0.5)
    close(out2[0], ref2[0].double(), dt, 'gconv (bias only)')
    assert out2[1] is None and out2[3] is None
    assert float((out2[2].cpu().double() - ref2[2].double()).abs().max() / ref2[2].double().abs().max()) < 2e-4


@pytest.mark.parametrize('dt', DTYPES)
def test_blur_separable_taps(dt):
    x, xr = prep(rnd(70, 2, 40, 40, 32), dt)
    k1 = torch.tensor([1., 3., 3., 1.])
    taps = torch.outer(k1, k1) / 64 * 4
    y = K.upfirdn2d(x, taps.cuda(), 1, 1, 1, 1, 39, 39, True)
    close(y, R.upfirdn2d(xr, taps.double(), 1, 1, 1, 1, 39, 39, True), dt)
    y = K.upfirdn2d(x, taps.cuda(), 1, 1, 2, 2, 41, 41, False)
    close(y, R.upfirdn2d(xr, taps.double(), 1, 1, 2, 2, 41, 41, False), dt)


CONV_CASES = [
    (2, 32, 32, 32, 3, 1, 1, 1, 0, True), (2, 16, 16, 3, 64, 1, 1, 1, 0, False), (3, 8, 8, 512, 3, 1, 1, 1, 0, True),
    (2, 16, 16, 3, 512, 1, 1, 1, 0, True), (2, 9, 9, 40, 2, 1, 1, 1, 0, False),
    # register-weight pointwise kernels: every lanes-per-pixel / vectors-per-lane variant, ragged pixel counts
    (2, 9, 9, 128, 3, 1, 1, 1, 0, True), (1, 11, 7, 256, 3, 1, 1, 1, 0, True), (2, 13, 13, 1024, 3, 1, 1, 1, 0, False),
    (3, 5, 5, 64, 4, 1, 1, 1, 0, True), (2, 5, 5, 4, 128, 1, 1, 1, 0, True), (1, 33, 31, 3, 32, 1, 1, 1, 0, False),
    (2, 7, 9, 2, 256, 1, 1, 1, 0, True), (2, 1, 1, 512, 3, 1, 1, 1, 0, True),
    # b, h, w, ic, oc, k, up, down, pad0, per_sample
    (2, 8, 8, 16, 32, 3, 1, 1, 1, False), (3, 7, 9, 8, 12, 3, 1, 1, 1, True), (2, 6, 6, 8, 6, 3, 2, 1, 2, True),
    (2, 9, 9, 16, 8, 3, 1, 2, 0, False), (2, 8, 8, 3, 32, 1, 1, 1, 0, False), (2, 4, 4, 513, 64, 3, 1, 1, 1, False),
    (4, 16, 16, 64, 3, 1, 1, 1, 0, True), (2, 7, 7, 8, 8, 1, 1, 2, 0, False), (1, 12, 12, 24, 40, 3, 1, 1, 1, False),
    (2, 5, 5, 32, 16, 3, 2, 1, 2, False), (2, 17, 17, 16, 16, 3, 1, 2, 0, True), (3, 4, 4, 128, 128, 3, 1, 1, 1, False),
]


def conv_out_hw(h, w, k, up, down, pad0):
    zh, zw = (h - 1) * up + 1, (w - 1) * up + 1
    return (zh + 2 * pad0 - k) // down + 1, (zw + 2 * pad0 - k) // down + 1


@pytest.mark.parametrize('dt', DTYPES16)
@pytest.mark.parametrize('case', CONV_CASES)
def test_conv_fwd_and_wgrad(case, dt):
    b, h, w, ic, oc, k, up, down, pad0, ps = case
    oh, ow = conv_out_hw(h, w, k, up, down, pad0)
    x, xr = prep(rnd(11, b, h, w, ic), dt)
    wt, wr = prep(rnd(12, b if ps else 1, k, k, oc, ic) / (ic * k * k) ** 0.5, dt)
    y = K.conv_fwd(x, wt, oh, ow, up, down, pad0)
    close(y, R.conv_fwd(xr, wr, oh, ow, up, down, pad0), dt, 'fwd')
    # fused epilogue
    bias, rs, nw = rnd(13, oc).float(), (rnd(14, b, oc).abs() + 0.5).float(), torch.tensor([0.3])
    noise, noiser = prep(rnd(15, b, oh, ow), dt)
    y2 = K.conv_fwd(x, wt, oh, ow, up, down, pad0, bias.cuda(), rs.cuda(), noise, nw.cuda(), 0.2, 2 ** 0.5)
    close(y2, R.conv_fwd(xr, wr, oh, ow, up, down, pad0, bias.double(), rs.double(), noiser, nw.double(), 0.2, 2 ** 0.5), dt, 'fwd+epilogue')
    gy, gyr = prep(rnd(16, b, oh, ow, oc), dt)
    gw = K.conv_wgrad(x, gy, k, k, up, down, pad0, ps)
    gwr = R.conv_wgrad(xr, gyr, k, k, up, down, pad0, ps)
    err = float((gw.cpu().double() - gwr).abs().max() / gwr.abs().max())
    assert err < 1e-4, f'wgrad rel err {err:.2e}'


@pytest.mark.parametrize('dt', DTYPES)
def test_linear_and_gemm(dt):
    x, xr = prep(rnd(20, 7, 96), dt)
    w, b = rnd(21, 40, 96).float(), rnd(22, 40).float()
    for act in [0, 1]:
        close(K.linear_fwd(x, w.cuda(), b.cuda(), 0.05, 0.3, act), R.linear_fwd(xr, w.double(), b.double(), 0.05, 0.3, act), dt)
    if dt == torch.float32:
        for ta in [False, True]:
            for tb in [False, True]:
                a = rnd(23, *((33, 17) if ta else (17, 33))).float()
                bb = rnd(24, *((50, 33) if tb else (33, 50))).float()
                close(K.gemm_f32(a.cuda(), bb.cuda(), ta, tb, 0.7), R.gemm_f32(a.double(), bb.double(), ta, tb, 0.7), dt, f'gemm {ta}{tb}')
        # split-K path: skinny M, long K
        a, bb = rnd(25, 16, 8192).float(), rnd(26, 512, 8192).float()
        close(K.gemm_f32(a.cuda(), bb.cuda(), False, True, 1.0), R.gemm_f32(a.double(), bb.double(), False, True, 1.0), dt, 'split-k')


def test_adam_ema():
    p, g = rnd(30, 1000).float(), rnd(31, 1000).float()
    m, v, ema = torch.zeros(1000), torch.zeros(1000), p.clone()
    dev = [t.clone().cuda() for t in (p, g, m, v, ema)]
    ref = [t.clone().double() for t in (p, g, m, v, ema)]
    for step in (1, 2, 3):
        bc = torch.tensor([1 - 0.0 ** step, 1 - 0.99 ** step], dtype=torch.float32)
        K.adam_ema(*dev, lr=0.002, beta1=0.0, beta2=0.99, eps=1e-8, bias_corr=bc.cuda(), ema_decay=0.998)
        R.adam_ema(*ref, lr=0.002, beta1=0.0, beta2=0.99, eps=1e-8, bias_corr=bc.double(), ema_decay=0.998)
    for d, r in zip(dev, ref):
        assert float((d.cpu().double() - r).abs().max()) < 1e-5
    # agrees with torch.optim.Adam
    pt = torch.nn.Parameter(p.clone().cuda())
    opt = torch.optim.Adam([pt], lr=0.002, betas=(0.0, 0.99))
    for _ in range(3):
        pt.grad = g.cuda()
        opt.step()
    assert float((pt.detach() - dev[0]).abs().max()) < 1e-5


# ---- tcgen05 implicit-GEMM engine ----------------------------------------------------------------
UMMA_CASES = [
    # b, h, w, ic, oc, k, up, down, pad0, per_sample
    (2, 16, 16, 64, 64, 3, 1, 1, 1, False), (4, 8, 8, 128, 128, 3, 1, 1, 1, False), (16, 4, 4, 512, 512, 3, 1, 1, 1, False),
    (2, 32, 32, 32, 32, 3, 1, 1, 1, False), (2, 64, 64, 64, 32, 3, 1, 1, 1, False), (3, 16, 16, 64, 64, 3, 1, 1, 1, True),
    (2, 32, 32, 32, 32, 3, 1, 1, 1, True), (2, 16, 16, 64, 64, 1, 1, 1, 0, False), (2, 16, 16, 96, 48, 3, 1, 1, 1, False),
    (4, 8, 8, 512, 512, 3, 1, 1, 1, False), (2, 8, 8, 64, 64, 3, 2, 1, 2, False), (2, 8, 8, 64, 64, 3, 2, 1, 2, True),
    (2, 4, 4, 128, 64, 3, 2, 1, 2, False), (2, 17, 17, 64, 128, 3, 1, 2, 0, False), (2, 33, 33, 32, 64, 3, 1, 2, 0, False),
    (2, 15, 15, 64, 128, 1, 1, 2, 0, False), (2, 8, 8, 128, 64, 1, 2, 1, 0, False), (5, 16, 16, 64, 256, 3, 1, 1, 1, False),
    (1, 128, 128, 64, 64, 3, 1, 1, 1, False), (2, 64, 64, 32, 32, 3, 2, 1, 2, True), (2, 129, 129, 32, 32, 3, 1, 2, 0, True),
]


@pytest.mark.parametrize('dt', [torch.bfloat16, torch.float16])
@pytest.mark.parametrize('case', UMMA_CASES)
def test_conv_fwd_umma(case, dt):
    """the tcgen05 engine against the fp64 contract stand-in, and against the CUDA-core engine"""
    b, h, w, ic, oc, k, up, down, pad0, ps = case
    if up == 2:
        oh, ow = (h - 1) * 2 + k - 2 * (k - 1 - pad0), (w - 1) * 2 + k - 2 * (k - 1 - pad0)
    else:
        oh, ow = conv_out_hw(h, w, k, up, down, pad0)
    x, xr = prep(rnd(41, b, h, w, ic), dt)
    wt, wr = prep(rnd(42, b if ps else 1, k, k, oc, ic) / (ic * k * k) ** 0.5, dt)
    bias, rs, nw = rnd(43, oc).float(), (rnd(44, b, oc).abs() + 0.5).float(), torch.tensor([0.3])
    noise, noiser = prep(rnd(45, b, oh, ow), dt)
    ref = R.conv_fwd(xr, wr, oh, ow, up, down, pad0)
    ref_ep = R.conv_fwd(xr, wr, oh, ow, up, down, pad0, bias.double(), rs.double(), noiser, nw.double(), 0.2, 2 ** 0.5)
    with engines('fwd_umma', 'fwd_halo'):
        y = K.conv_fwd(x, wt, oh, ow, up, down, pad0)
        y_ep = K.conv_fwd(x, wt, oh, ow, up, down, pad0, bias.cuda(), rs.cuda(), noise, nw.cuda(), 0.2, 2 ** 0.5)
    torch.cuda.synchronize()
    prev = K.set_conv_engine(2)
    try:
        with engines('fwd_umma'):
            y_gen = K.conv_fwd(x, wt, oh, ow, up, down, pad0)
        close(y_gen, ref, dt, 'general umma fwd')
    finally:
        K.set_conv_engine(prev)
    prev = K.set_conv_engine(1)
    try:
        with engines('fwd_simt'):
            y_simt = K.conv_fwd(x, wt, oh, ow, up, down, pad0)
    finally:
        K.set_conv_engine(prev)
    close(y, ref, dt, 'umma fwd')
    close(y_ep, ref_ep, dt, 'umma fwd+epilogue')
    close(y, y_simt.cpu().double(), dt, 'umma vs simt')


WGRAD_UMMA_CASES = UMMA_CASES[:8] + UMMA_CASES[9:] + [
    (2, 16, 16, 256, 128, 3, 1, 1, 1, False), (2, 16, 16, 32, 128, 3, 1, 1, 1, False), (3, 32, 32, 64, 32, 3, 2, 1, 2, False),
    (2, 65, 65, 64, 64, 3, 1, 2, 0, False), (16, 8, 8, 512, 512, 3, 1, 1, 1, False),
]


@pytest.mark.parametrize('dt', [torch.bfloat16, torch.float16])
@pytest.mark.parametrize('case', WGRAD_UMMA_CASES)
def test_conv_wgrad_umma(case, dt):
    b, h, w, ic, oc, k, up, down, pad0, ps = case
    if up == 2:
        oh, ow = (h - 1) * 2 + k - 2 * (k - 1 - pad0), (w - 1) * 2 + k - 2 * (k - 1 - pad0)
    else:
        oh, ow = conv_out_hw(h, w, k, up, down, pad0)
    x, xr = prep(rnd(51, b, h, w, ic), dt)
    gy, gyr = prep(rnd(52, b, oh, ow, oc), dt)
    with engines('wgrad_umma', 'wgrad_halo'):
        gw = K.conv_wgrad(x, gy, k, k, up, down, pad0, ps)
    torch.cuda.synchronize()
    prev = K.set_conv_engine(1)
    try:
        with engines('wgrad_simt'):
            gw_simt = K.conv_wgrad(x, gy, k, k, up, down, pad0, ps)
    finally:
        K.set_conv_engine(prev)
    scale = float(gw_simt.abs().max())
    err_simt = float((gw - gw_simt).abs().max()) / scale
    assert err_simt < 2e-4, f'umma vs simt wgrad rel err {err_simt:.2e}'
    if b * oh * ow * ic * oc <= 2 ** 27:
        gwr = R.conv_wgrad(xr, gyr, k, k, up, down, pad0, ps)
        err = float((gw.cpu().double() - gwr).abs().max() / gwr.abs().max())
        assert err < 2e-4, f'wgrad rel err {err:.2e}'


HALO_CASES = [
    # b, h, w, ic, oc, k, per_sample       (3x3 pad 1 / 1x1 pad 0, stride 1, <= 64 channels)
    (2, 32, 32, 64, 64, 3, False), (3, 64, 64, 64, 64, 3, True), (2, 32, 32, 32, 32, 3, True), (2, 48, 40, 32, 64, 3, False),
    (2, 32, 32, 64, 32, 3, True), (2, 64, 64, 64, 64, 1, False), (2, 40, 24, 32, 32, 1, True), (1, 128, 128, 64, 64, 3, False),
    (5, 17, 9, 64, 16, 3, True),
    # protocol stress for the multi-issuer kernel: a per-sample weight reload every 1 / 2 / 4 tiles, more samples than SMs
    (150, 16, 8, 32, 32, 3, True), (160, 16, 16, 64, 64, 3, True), (75, 32, 16, 64, 32, 3, True), (301, 16, 8, 32, 64, 1, True),
]


@pytest.mark.parametrize('dt', [torch.bfloat16, torch.float16])
@pytest.mark.parametrize('case', HALO_CASES)
def test_conv_fwd_halo(case, dt):
    """halo-reuse tcgen05 variant vs the general tcgen05 kernel (engine 2) and the fp64 stand-in"""
    b, h, w, ic, oc, k, ps = case
    pad0 = k // 2
    x, xr = prep(rnd(81, b, h, w, ic), dt)
    wt, wr = prep(rnd(82, b if ps else 1, k, k, oc, ic) / (ic * k * k) ** 0.5, dt)
    bias, rs, nw = rnd(83, oc).float(), (rnd(84, b, oc).abs() + 0.5).float(), torch.tensor([0.3])
    noise, noiser = prep(rnd(85, b, h, w), dt)
    with engines('fwd_halo'):
        y = K.conv_fwd(x, wt, h, w, 1, 1, pad0)
        y_ep = K.conv_fwd(x, wt, h, w, 1, 1, pad0, bias.cuda(), rs.cuda(), noise, nw.cuda(), 0.2, 2 ** 0.5)
    torch.cuda.synchronize()
    prev = K.set_conv_engine(2)
    try:
        with engines('fwd_umma'):
            y_gen = K.conv_fwd(x, wt, h, w, 1, 1, pad0)
    finally:
        K.set_conv_engine(prev)
    assert float((y.float() - y_gen.float()).abs().max()) <= 2e-2 * float(y_gen.float().abs().max()), 'halo vs general'
    close(y, R.conv_fwd(xr, wr, h, w, 1, 1, pad0), dt, 'halo fwd')
    close(y_ep, R.conv_fwd(xr, wr, h, w, 1, 1, pad0, bias.double(), rs.double(), noiser, nw.double(), 0.2, 2 ** 0.5), dt, 'halo fwd+ep')


@pytest.mark.parametrize('dt', [torch.bfloat16, torch.float16])
@pytest.mark.parametrize('case', [c for c in HALO_CASES if c[4] in (32, 64)] + [(16, 64, 64, 32, 32, 3, False)])
def test_conv_wgrad_halo(case, dt):
    """halo-reuse weight-gradient kernel vs the general tcgen05 wgrad (engine 2) and the fp64 stand-in"""
    b, h, w, ic, oc, k, ps = case
    pad0 = k // 2
    x, xr = prep(rnd(91, b, h, w, ic), dt)
    gy, gyr = prep(rnd(92, b, h, w, oc), dt)
    with engines('wgrad_halo'):
        gw = K.conv_wgrad(x, gy, k, k, 1, 1, pad0, ps)
    torch.cuda.synchronize()
    prev = K.set_conv_engine(2)
    try:
        with engines('wgrad_umma'):
            gw_gen = K.conv_wgrad(x, gy, k, k, 1, 1, pad0, ps)
    finally:
        K.set_conv_engine(prev)
    scale = float(gw_gen.abs().max())
    assert float((gw - gw_gen).abs().max()) / scale < 2e-4, 'halo vs general wgrad'
    gwr = R.conv_wgrad(xr, gyr, k, k, 1, 1, pad0, ps)
    assert float((gw.cpu().double() - gwr).abs().max() / gwr.abs().max()) < 2e-4


# ---- convolutions between space-to-depth views (fused FIR resampling, include/b200gan.h "packed") -------------
PACKED_CASES = [
    # b, h, w (view extent), ic, oc (view channels), per_sample, pack_in, pack_out
    (2, 32, 32, 64, 128, True, False, True),      # upsampling StyledConv 64 -> 32 (1024^2 layer)
    (2, 32, 24, 128, 64, True, True, False),      # ... its data gradient
    (2, 32, 32, 128, 64, False, True, False),     # downsampling ConvLayer 32 -> 64 (res1024.conv2)
    (3, 48, 40, 64, 128, False, False, True),     # ... its data gradient, ragged tiles
    (2, 32, 32, 64, 64, False, True, False), (2, 32, 32, 32, 64, True, False, True), (1, 64, 64, 128, 128, False, True, True),
    (2, 8, 8, 16, 32, False, True, True), (2, 6, 10, 32, 16, True, False, True),      # small: CUDA-core engine
    # protocol stress: per-sample weight reload every 1 / 2 tiles with 2 activation stages (<= 2 issuers) and 2 chunks (1 issuer)
    (150, 16, 8, 64, 128, True, False, True), (80, 16, 16, 128, 64, True, True, False),
]


@pytest.mark.parametrize('dt', [torch.bfloat16, torch.float16])
@pytest.mark.parametrize('case', PACKED_CASES)
def test_conv_packed(case, dt):
    """packed forward (+ fused epilogue) and weight gradient on every engine that takes the shape, against the
    fp64 stand-in (which materialises the space-to-depth views)"""
    b, h, w, ic, oc, ps, pin, pout = case
    x, xr = prep(rnd(101, b, 2 * h, 2 * w, ic // 4) if pin else rnd(101, b, h, w, ic), dt)
    wt, wr = prep(rnd(102, b if ps else 1, 3, 3, oc, ic) / (ic * 9) ** 0.5, dt)
    ocp = oc // 4 if pout else oc
    oh, ow = (2 * h, 2 * w) if pout else (h, w)
    bias, rs, nw = rnd(103, ocp).float(), (rnd(104, b, ocp).abs() + 0.5).float(), torch.tensor([0.3])
    noise, noiser = prep(rnd(105, b, oh, ow), dt)
    gy, gyr = prep(rnd(106, b, oh, ow, ocp), dt)
    ref = R.conv_fwd(xr, wr, h, w, 1, 1, 1, pack_in=pin, pack_out=pout)
    ref_ep = R.conv_fwd(xr, wr, h, w, 1, 1, 1, bias.double(), rs.double(), noiser, nw.double(), 0.2, 2 ** 0.5,
                        pack_in=pin, pack_out=pout)
    gwr = R.conv_wgrad(xr, gyr, 3, 3, 1, 1, 1, ps, pack_x=pin, pack_gy=pout)
    small = h < 16 or w < 8 or min(ic, oc) < 32          # below the halo kernel's tile / channel minimum: CUDA cores
    for engine in (0, 1):
        prev = K.set_conv_engine(engine)
        if engine == 1 or small:
            want = ('fwd_simt', 'wgrad_simt')
        elif pin and pout:          # both sides packed (not a layer of the path): 9 x 128 x 128 resident weights exceed shared memory
            want = ('fwd_simt', 'wgrad_halo')
        else:
            want = ('fwd_halo', 'wgrad_halo')
        try:
            with engines(*want):
                y = K.conv_fwd(x, wt, h, w, 1, 1, 1, pack_in=pin, pack_out=pout)
                y_ep = K.conv_fwd(x, wt, h, w, 1, 1, 1, bias.cuda(), rs.cuda(), noise, nw.cuda(), 0.2, 2 ** 0.5,
                                  pack_in=pin, pack_out=pout)
                gw = K.conv_wgrad(x, gy, 3, 3, 1, 1, 1, ps, pack_x=pin, pack_gy=pout)
            torch.cuda.synchronize()
        finally:
            K.set_conv_engine(prev)
        assert y.shape == ref.shape
        close(y, ref, dt, f'packed fwd engine {engine}')
        close(y_ep, ref_ep, dt, f'packed fwd+epilogue engine {engine}')
        err = float((gw.cpu().double() - gwr).abs().max() / gwr.abs().max())
        assert err < 2e-4, f'packed wgrad engine {engine} rel err {err:.2e}'


# ---- weight (de)modulation kernels ------------------------------------------------------------------------------
@pytest.mark.parametrize('dt', DTYPES)
@pytest.mark.parametrize('cfg', [
    # b, oc, ic, k, demod, flip
    (3, 5, 4, 3, True, False), (2, 32, 32, 3, True, True), (16, 64, 64, 3, True, False), (4, 3, 32, 1, False, False),
    (1, 64, 32, 3, False, False), (2, 128, 256, 3, True, False), (1, 64, 3, 1, False, False), (2, 40, 24, 3, True, True),
])
def test_modweight(cfg, dt):
    b, oc, ic, k, demod, flip = cfg
    w = rnd(120, oc, ic, k, k).float()
    s = (rnd(121, b, ic) * 0.5 + 1.0).float()
    scale = 1.0 / (ic * k * k) ** 0.5
    wk, wkt, d = K.modweight_fwd(w.cuda(), s.cuda(), scale, demod, flip, dt, want_adjoint=True)
    rk, rkt, rd = R.modweight_fwd(w.double(), s.double(), scale, demod, flip, torch.float64, want_adjoint=True)
    assert wk.dtype == dt and wk.shape == rk.shape and wkt.shape == rkt.shape
    close(wk, rk, dt, 'wk')
    close(wkt, rkt, dt, 'wk adjoint')
    if demod:
        assert float((d.cpu().double() - rd).abs().max() / rd.abs().max()) < 1e-5
    else:
        assert d is None
    wk2, wkt2, _ = K.modweight_fwd(w.cuda(), s.cuda(), scale, demod, flip, dt, want_adjoint=False)
    assert wkt2 is None and torch.equal(wk2, wk)
    g = rnd(122, b, k, k, oc, ic).float()
    gs, gw = K.modweight_bwd(g.cuda(), w.cuda(), s.cuda(), d, scale, demod, flip)
    rgs, rgw = R.modweight_bwd(g.double(), w.double(), s.double(), rd, scale, demod, flip)
    assert float((gs.cpu().double() - rgs).abs().max() / rgs.abs().max()) < 2e-5, 'gs'
    assert float((gw.cpu().double() - rgw).abs().max() / rgw.abs().max()) < 2e-5, 'gweight'
    gs2, gw2 = K.modweight_bwd(g.cuda(), w.cuda(), s.cuda(), d, scale, demod, flip, want_gs=False, want_gw=True)
    assert gs2 is None and float((gw2 - gw).abs().max()) <= 1e-6 * float(gw.abs().max())


# ---- output-shaped side inputs of the convolution epilogue: addend and (backward mode) gate -----------------------
SIDE_CASES = [
    # b, h, w, ic, oc, k, up, down, pad0, per_sample, pack_in, pack_out, engines allowed under auto
    (2, 32, 32, 32, 32, 3, 1, 1, 1, True, False, False, ('fwd_halo',)),
    (2, 32, 32, 64, 64, 3, 1, 1, 1, False, False, False, ('fwd_halo',)),
    (2, 48, 40, 32, 64, 1, 1, 1, 0, False, False, False, ('fwd_halo',)),
    (2, 32, 24, 128, 64, 3, 1, 1, 1, True, True, False, ('fwd_halo',)),      # fused-up data gradient (pack_in)
    (3, 48, 40, 64, 128, 3, 1, 1, 1, False, False, True, ('fwd_halo',)),     # fused-down data gradient (pack_out)
    (4, 16, 16, 128, 128, 3, 1, 1, 1, False, False, False, ('fwd_umma',)),
    (2, 8, 8, 256, 128, 3, 2, 1, 2, False, False, False, ('fwd_umma',)),     # transposed conv (data gradient of a stride-2 conv)
    (2, 33, 33, 64, 128, 3, 1, 2, 0, False, False, False, ('fwd_umma',)),
    (2, 16, 16, 3, 64, 1, 1, 1, 0, True, False, False, ('fwd_pointwise',)),  # ToRGB data gradient
    (2, 16, 16, 64, 3, 1, 1, 1, 0, True, False, False, ('fwd_pointwise',)),
    (3, 7, 9, 8, 12, 3, 1, 1, 1, True, False, False, ('fwd_simt',)),
]


@pytest.mark.parametrize('dt', DTYPES)
@pytest.mark.parametrize('case', SIDE_CASES)
def test_conv_epilogue_addend_and_gate(case, dt):
    b, h, w, ic, oc, k, up, down, pad0, ps, pin, pout, allowed = case
    if up == 2:
        oh, ow = (h - 1) * 2 + k - 2 * (k - 1 - pad0), (w - 1) * 2 + k - 2 * (k - 1 - pad0)
    else:
        oh, ow = conv_out_hw(h, w, k, up, down, pad0)
    x, xr = prep(rnd(131, b, 2 * h, 2 * w, ic // 4) if pin else rnd(131, b, h, w, ic), dt)
    wt, wr = prep(rnd(132, b if ps else 1, k, k, oc, ic) / (ic * k * k) ** 0.5, dt)
    yshape = (b, 2 * oh, 2 * ow, oc // 4) if pout else (b, oh, ow, oc)
    ocp = yshape[-1]
    add, addr = prep(rnd(133, *yshape), dt)
    gate, gater = prep(rnd(134, *yshape), dt)
    rs, bias = (rnd(135, b, ocp).abs() + 0.5).float(), rnd(136, ocp).float()
    kw = dict(pack_in=pin, pack_out=pout)
    want = allowed if dt == torch.bfloat16 else (('fwd_pointwise',) if 'fwd_pointwise' in allowed else ('fwd_simt',))
    with engines(*want):
        y_add = K.conv_fwd(x, wt, oh, ow, up, down, pad0, bias.cuda(), None, None, None, 0.2, 2 ** 0.5, addend=add, **kw)
        y_gate = K.conv_fwd(x, wt, oh, ow, up, down, pad0, None, rs.cuda(), None, None, 0.2, 2 ** 0.5, addend=add, gate=gate, **kw)
        y_gate2 = K.conv_fwd(x, wt, oh, ow, up, down, pad0, None, None, None, None, 0.2, 2 ** 0.5, gate=gate, **kw)
    close(y_add, R.conv_fwd(xr, wr, oh, ow, up, down, pad0, bias.double(), None, None, None, 0.2, 2 ** 0.5, addend=addr, **kw), dt, 'addend')
    close(y_gate, R.conv_fwd(xr, wr, oh, ow, up, down, pad0, None, rs.double(), None, None, 0.2, 2 ** 0.5, addend=addr, gate=gater, **kw),
          dt, 'addend + gate')
    close(y_gate2, R.conv_fwd(xr, wr, oh, ow, up, down, pad0, None, None, None, None, 0.2, 2 ** 0.5, gate=gater, **kw), dt, 'gate')


def test_protocol_stress_multi_issuer_wgrad_and_sub_tiles():
    """Barrier-protocol stress of the round-2 pipeline changes over many small shapes (a mismatch shows as a device-side
    barrier timeout / wrong values): the halo weight gradient with one issuing warp per accumulator group against the
    CUDA-core engine, and the general forward engine with two 64-channel sub-tiles per stage against one (bit for bit)."""
    import os
    torch.manual_seed(0)
    for b in (1, 3, 8, 17):
        for h, w in ((16, 8), (48, 40), (64, 24), (130, 70)):
            for ch in (32, 64):
                for k in (1, 3):
                    for ps in (False, True):
                        x = torch.randn(b, h, w, ch, device='cuda').bfloat16()
                        gy = torch.randn(b, h, w, ch, device='cuda').bfloat16()
                        with engines('wgrad_halo'):
                            gw = K.conv_wgrad(x, gy, k, k, 1, 1, k // 2, ps)
                        prev = K.set_conv_engine(1)
                        try:
                            ref = K.conv_wgrad(x, gy, k, k, 1, 1, k // 2, ps)
                        finally:
                            K.set_conv_engine(prev)
                        assert float((gw - ref).abs().max() / ref.abs().max()) < 2e-3, (b, h, w, ch, k, ps)
    for b in (1, 8):
        for h in (16, 33, 64):
            for ic, oc in ((128, 64), (128, 128), (256, 32), (384, 128)):
                for up, down, pad0 in ((1, 1, 1), (2, 1, 2), (1, 2, 0)):
                    hh = h + 1 if (down == 2 and h % 2 == 0) else h
                    oh = 2 * hh + 1 if up == 2 else (hh + 2 * pad0 - 3) // down + 1
                    x = torch.randn(b, hh, hh, ic, device='cuda').bfloat16()
                    wt = (torch.randn(1, 3, 3, oc, ic, device='cuda') / (3 * ic ** 0.5)).bfloat16()
                    with engines('fwd_umma'):
                        y = K.conv_fwd(x, wt, oh, oh, up, down, pad0)
                    os.environ['B200GAN_UMMA_SUB'] = '1'
                    try:
                        y1 = K.conv_fwd(x, wt, oh, oh, up, down, pad0)
                    finally:
                        os.environ.pop('B200GAN_UMMA_SUB', None)
                    assert torch.equal(y, y1), (b, hh, ic, oc, up, down)
    torch.cuda.synchronize()


def test_general_wgrad_issuer_counts_agree():
    """The general weight-gradient kernel with two and three MMA-issuing warps (B200GAN_WGRAD_GENERAL_ISSUERS, opt-in)
    against the shipped single issuer, on the ring geometries where skipped barrier phases once hung a block (oc = 64 with
    ic >= 128: five x stages, eight groups per K step; the 4 / 2 / 2 / 1-group phases of the transposed conv) --
    profiles/r02_wgrad_general_issuers.md.  Differences are the order of the split-K fp32 atomics."""
    import os
    torch.manual_seed(0)
    env = 'B200GAN_WGRAD_GENERAL_ISSUERS'
    for b in (1, 8):
        for h in (16, 33, 64):
            for ic, oc in ((128, 64), (256, 64), (128, 128), (64, 128), (512, 512)):
                for k, up, down, pad0 in ((3, 1, 1, 1), (3, 2, 1, 2), (3, 1, 2, 0), (1, 1, 1, 0)):
                    hh = h + 1 if (down == 2 and h % 2 == 0) else h
                    oh = 2 * hh + 1 if up == 2 else (hh + 2 * pad0 - k) // down + 1
                    for ps in (False, True):
                        x = torch.randn(b, hh, hh, ic, device='cuda').bfloat16()
                        gy = torch.randn(b, oh, oh, oc, device='cuda').bfloat16()
                        with engines('wgrad_umma'):
                            ref = K.conv_wgrad(x, gy, k, k, up, down, pad0, ps)
                        for iss in ('2', '3'):
                            os.environ[env] = iss
                            try:
                                got = K.conv_wgrad(x, gy, k, k, up, down, pad0, ps)
                            finally:
                                os.environ.pop(env, None)
                            assert float((got - ref).abs().max() / ref.abs().max()) < 1e-4, (b, hh, ic, oc, k, up, down, ps, iss)
    torch.cuda.synchronize()
