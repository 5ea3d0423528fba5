"""Host-side logic of gan_control_b200 (autograd formulas, layouts, geometry, module wiring,
state_dict layout) checked on CPU in fp64 against the reference-generated goldens.  The CUDA
kernels are replaced by their contract stand-ins (fixture `cpu_kernels`); the kernels themselves
are verified against the same stand-ins in tests/test_kernels_gpu.py."""
import math

import numpy as np
import pytest
import torch

from gan_control_b200 import modules as M
from gan_control_b200 import ops
from oracle import params as P
from golden_io import Fixture, reduce_like_golden, max_rel

F64 = torch.float64
UP_CASES = ['g_upblur', 'rgb_skip', 'd_conv2_blur', 'd_skip_blur', 'ada_up', 'ada_down', 'down_module', 'negpad', 'rect']


def rnd(seed, *shape, dtype=F64):
    return torch.from_numpy(np.random.default_rng(seed).standard_normal(shape)).to(dtype)


@pytest.mark.parametrize('layout', ['nchw', 'channels_last'])
@pytest.mark.parametrize('name', UP_CASES)
def test_upfirdn2d(cpu_kernels, name, layout):
    fx = Fixture('upfirdn2d')
    up, down, p0, p1 = [int(v) for v in fx.np(name + '.cfg')]
    x = fx.t(name + '.x')
    if layout == 'channels_last':
        x = x.contiguous(memory_format=torch.channels_last)
    x.requires_grad_(True)
    y = ops.upfirdn2d(x, fx.t(name + '.k'), up, down, (p0, p1))
    assert max_rel(y, fx.t(name + '.y')) < 1e-12
    gx, = torch.autograd.grad(y, x, fx.t(name + '.gy'), create_graph=True)
    assert max_rel(gx, fx.t(name + '.gx')) < 1e-12
    # second derivative: the op is linear, so d/d(gy) <gx, v> = upfirdn2d(v)
    v = rnd(1, *gx.shape)
    gy = fx.t(name + '.gy').requires_grad_(True)
    gx2, = torch.autograd.grad(y, x, gy, create_graph=True)
    ggy, = torch.autograd.grad((gx2 * v).sum(), gy)
    assert max_rel(ggy, ops.upfirdn2d(v, fx.t(name + '.k'), up, down, (p0, p1))) < 1e-12


def test_fused_leaky_relu(cpu_kernels):
    fx = Fixture('bias_act')
    for c in ['c0', 'c1']:
        for cl in [False, True]:
            x = fx.t(c + '.x')
            if cl and x.ndim == 4:
                x = x.contiguous(memory_format=torch.channels_last)
            x.requires_grad_(True)
            b = fx.t(c + '.b').requires_grad_(True)
            y = ops.fused_leaky_relu(x, b)
            assert max_rel(y, fx.t(c + '.y')) < 1e-13
            gx, gb = torch.autograd.grad(y, (x, b), fx.t(c + '.gy'))
            assert max_rel(gx, fx.t(c + '.gx')) < 1e-13 and max_rel(gb, fx.t(c + '.gb')) < 1e-13
    m = M.FusedLeakyReLU(5).double()
    m.bias.data.copy_(fx.t('mod.b'))
    assert max_rel(m(fx.t('mod.x')), fx.t('mod.y')) < 1e-13


def test_equal_linear(cpu_kernels):
    fx = Fixture('equal_linear')
    for c in ['c0', 'c1', 'c2']:
        lr_mul, act = fx.np(c + '.cfg')
        w = fx.t(c + '.w')
        m = M.EqualLinear(w.shape[1], w.shape[0], lr_mul=float(lr_mul), activation='fused_lrelu' if act else None).double()
        m.weight.data.copy_(w)
        m.bias.data.copy_(fx.t(c + '.b'))
        x = fx.t(c + '.x').requires_grad_(True)
        y = m(x)
        assert max_rel(y, fx.t(c + '.y')) < 1e-12
        g = torch.autograd.grad(y, (x, m.weight, m.bias), fx.t(c + '.gy'))
        for gi, n in zip(g, ['gx', 'gw', 'gb']):
            assert max_rel(gi, fx.t(f'{c}.{n}')) < 1e-12, n


@pytest.mark.parametrize('form', ['weight', 'activation'])
@pytest.mark.parametrize('name', ['plain3', 'up3', 'rgb1', 'plain3_b1'])
def test_modulated_conv(cpu_kernels, name, form):
    fx = Fixture('modconv')
    ic, oc, k, demod, up, h, b, sdim = [int(v) for v in fx.np(name + '.cfg')]
    m = M.ModulatedConv2d(ic, oc, k, sdim, demodulate=bool(demod), upsample=bool(up), conv_transpose=True).double()
    m.form = form
    m.weight.data.copy_(fx.t(name + '.w'))
    m.modulation.weight.data.copy_(fx.t(name + '.mw'))
    m.modulation.bias.data.copy_(fx.t(name + '.mb'))
    x = fx.t(name + '.x').contiguous(memory_format=torch.channels_last).requires_grad_(True)
    s = fx.t(name + '.s').requires_grad_(True)
    y = m(x, s)
    assert max_rel(y, fx.t(name + '.y')) < 1e-12
    ps = (x, s, m.weight, m.modulation.weight, m.modulation.bias)
    g = torch.autograd.grad(y, ps, fx.t(name + '.gy'), create_graph=True)
    for gi, n in zip(g, ['gx', 'gs', 'gw', 'gmw', 'gmb']):
        assert max_rel(gi, fx.t(f'{name}.{n}')) < 1e-11, n
    pl = g[1].pow(2).sum()
    assert max_rel(pl, fx.t(name + '.pl')) < 1e-11
    gg = torch.autograd.grad(pl, (x, s, m.weight, m.modulation.weight), allow_unused=True)
    for gi, n in zip(gg, ['pl_gx', 'pl_gs', 'pl_gw', 'pl_gmw']):
        if gi is None:
            continue
        assert max_rel(gi, fx.t(f'{name}.{n}')) < 1e-10, n
    # first-order only (no graph) takes the fused-kernel branches of the backward passes
    g2 = torch.autograd.grad(m(x, s), ps, fx.t(name + '.gy'))
    for gi, n in zip(g2, ['gx', 'gs', 'gw', 'gmw', 'gmb']):
        assert max_rel(gi, fx.t(f'{name}.{n}')) < 1e-11, n


def _load(module, sd):
    missing, unexpected = module.load_state_dict(sd, strict=True)
    return module


def test_styled_conv_and_to_rgb(cpu_kernels):
    fx = Fixture('modconv')
    for up in [0, 1]:
        n = f'styled_up{up}'
        m = M.StyledConv(8, 6, 3, 16, upsample=bool(up), conv_transpose=True).double()
        _load(m, fx.sub(n + '.sd.'))
        for form in ['weight', 'activation']:
            m.conv.form = form
            y = m(fx.t(n + '.x'), fx.t(n + '.s'), noise=fx.t(n + '.noise'))
            assert max_rel(y, fx.t(n + '.y')) < 1e-12
    m = M.ToRGB(8, 16, conv_transpose=True).double()
    _load(m, fx.sub('torgb.sd.'))
    assert max_rel(m(fx.t('torgb.x'), fx.t('torgb.s'), fx.t('torgb.skip')), fx.t('torgb.y')) < 1e-12
    assert max_rel(m(fx.t('torgb.x'), fx.t('torgb.s')), fx.t('torgb.y_noskip')) < 1e-12


def test_fcstack(cpu_kernels):
    fx = Fixture('fcstack')
    n_mlp, din, mid, dout, seed = [int(v) for v in fx.np('cfg')]
    m = M.FcStack(0.01, n_mlp, din, mid, dout).double()
    _load(m, P.seeded_state_dict(P.fc_stack_shapes(n_mlp, din, mid, dout), seed, dtype=F64))
    assert max_rel(m(fx.t('x')), fx.t('y')) < 1e-12


GROUPS = [('id', 0, 24), ('pose', 24, 40), ('other', 40, 64)]


def _fc_config(groups):
    return M.FcConfig([g[0] for g in groups],
                      {n: {'latent_place': [lo, hi], 'latent_size': hi - lo} for n, lo, hi in groups})


@pytest.mark.parametrize('name,fcg', [('g16', None), ('g16split', GROUPS)])
def test_generator(cpu_kernels, name, fcg):
    fx = Fixture('networks')
    size, sdim, n_mlp, seed = [int(v) for v in fx.np(name + '.cfg')]
    shapes = P.generator_shapes(size, sdim, n_mlp, 2, fcg)
    g = M.Generator(size, sdim, n_mlp, channel_multiplier=2, conv_transpose=True, split_fc=fcg is not None,
                    fc_config=_fc_config(fcg) if fcg else None, act_dtype=F64).double()
    assert {k: tuple(v.shape) for k, v in g.state_dict().items()} == {k: tuple(v) for k, v in shapes.items()}
    _load(g, P.seeded_state_dict(shapes, seed, dtype=F64))
    z, z2 = rnd(seed * 10, 2, sdim), rnd(seed * 10 + 1, 2, sdim)
    nl = g.num_layers
    noise = [rnd(seed * 100 + i, 2, 1, 2 ** ((i + 5) // 2), 2 ** ((i + 5) // 2)) for i in range(nl)]
    tol = 1e-10
    img, lat = g([z], noise=noise, return_latents=True)
    assert max_rel(lat, fx.t(f'{name}.f64.latent')) < tol
    assert max_rel(img, fx.t(f'{name}.f64.img')) < tol
    assert max_rel(g([z, z2], noise=noise, inject_index=3)[0], fx.t(f'{name}.f64.img_mix')) < tol
    mean_w = g.style(rnd(seed * 10 + 2, 16, sdim)).mean(0, keepdim=True)
    assert max_rel(mean_w, fx.t(f'{name}.f64.mean_w')) < tol
    assert max_rel(g([z], noise=noise, truncation=0.7, truncation_latent=mean_w)[0], fx.t(f'{name}.f64.img_trunc')) < tol
    assert max_rel(g([z], randomize_noise=False)[0], fx.t(f'{name}.f64.img_fixed_noise')) < tol
    # W+ input path (the notebook feeds gen_batch's latent_w back, gm.py:757-760)
    assert max_rel(g([lat.detach()], input_is_latent=True, noise=noise)[0], fx.t(f'{name}.f64.img')) < tol
    # path-length regulariser: double backward through every op
    from oracle import stylegan2_oracle as O
    img, lat = g([z], noise=noise, return_latents=True)
    torch.manual_seed(seed)
    pl_noise = torch.randn(img.shape, dtype=F64)     # the reference draws randn_like on an NCHW-contiguous image
    pen, mean, lengths = O.g_path_regularize(img, lat, 0.0, pl_noise=pl_noise)
    assert max_rel(pen, fx.t(name + '.pl.penalty')) < 1e-9
    assert max_rel(lengths, fx.t(name + '.pl.lengths')) < 1e-9
    g.zero_grad()
    pen.backward()
    params = dict(g.named_parameters())
    for k in fx.keys(name + '.pl.g.'):
        key = k[len(name) + 6:]
        assert max_rel(reduce_like_golden(params[key].grad), fx.t(k)) < 1e-8, key
    g.zero_grad()
    img, _ = g([z], noise=noise)
    (img * fx.t(name + '.bw.cot')).sum().backward()
    for k in fx.keys(name + '.bw.g.'):
        key = k[len(name) + 6:]
        assert max_rel(reduce_like_golden(params[key].grad), fx.t(k)) < 1e-8, key


def test_discriminator(cpu_kernels):
    from oracle import stylegan2_oracle as O
    fx = Fixture('networks')
    shapes = P.discriminator_shapes(16, 2)
    d = M.Discriminator(16, channel_multiplier=2, act_dtype=F64).double()
    assert {k: tuple(v.shape) for k, v in d.state_dict().items()} == {k: tuple(v) for k, v in shapes.items()}
    _load(d, P.seeded_state_dict(shapes, 21, dtype=F64))
    x = rnd(210, 8, 3, 16, 16).requires_grad_(True)
    pred, _ = d(x)
    assert max_rel(pred, fx.t('d16.f64.pred')) < 1e-10
    r1 = O.d_r1_loss(pred, x)
    assert max_rel(r1, fx.t('d16.r1')) < 1e-9
    d.zero_grad()
    (0.5 * r1 * 16 + 0 * pred[0]).sum().backward()
    params = dict(d.named_parameters())
    for k in fx.keys('d16.r1.g.'):
        key = k[len('d16.r1.g.'):]
        assert max_rel(reduce_like_golden(params[key].grad), fx.t(k)) < 1e-8, key
    d.zero_grad()
    loss = O.d_logistic_loss(d(x.detach())[0], d(rnd(211, 8, 3, 16, 16))[0])
    assert max_rel(loss, fx.t('d16.dloss')) < 1e-10
    loss.backward()
    for k in fx.keys('d16.dl.g.'):
        key = k[len('d16.dl.g.'):]
        assert max_rel(reduce_like_golden(params[key].grad), fx.t(k)) < 1e-8, key


@pytest.mark.parametrize('batch', [1, 2])
@pytest.mark.parametrize('form', ['weight', 'activation', 'auto'])
def test_data_grads_only_matches_plain_grad(cpu_kernels, batch, form):
    """inside `data_grads_only` the path-length inner gradient d(img.n)/d(latent) must equal the
    plain autograd.grad at EVERY batch size.  At batch 1 a style-modulated weight is (1,OC,IC,k,k) like a shared one;
    its gradient carries latent -> style -> weight and must not be skipped."""
    size, sdim = 16, 32
    torch.manual_seed(3)
    g = M.Generator(size, sdim, 2, channel_multiplier=2, conv_transpose=True).double()
    for m in g.modules():
        if isinstance(m, M.ModulatedConv2d):
            m.form = form
        if isinstance(m, M.NoiseInjection):
            m.weight.data.fill_(0.3)
    z = torch.randn(batch, sdim, dtype=F64)
    noise = [torch.randn(batch, 1, 2 ** ((i + 5) // 2), 2 ** ((i + 5) // 2), dtype=F64) for i in range(g.num_layers)]
    pl = torch.randn(batch, 3, size, size, dtype=F64)
    img, latents = g([z], return_latents=True, noise=noise)
    plain, = torch.autograd.grad((img * pl).sum(), latents, create_graph=True)
    img, latents = g([z], return_latents=True, noise=noise)
    with ops.data_grads_only():
        only, = torch.autograd.grad((img * pl).sum(), latents, create_graph=True)
    assert plain.abs().max() > 0
    assert max_rel(only, plain) < 1e-11
    # and the double backward through it (what the regulariser trains on)
    w = g.convs[0].conv.weight
    gg_plain, = torch.autograd.grad(plain.pow(2).sum(), w)
    gg_only, = torch.autograd.grad(only.pow(2).sum(), w)
    assert max_rel(gg_only, gg_plain) < 1e-10
