"""`ops.mod_conv` (weight path + convolution + epilogue as single kernels, first-order backward) against the closed
op algebra that the reference goldens pin (tests/test_host_algebra_cpu.py): same outputs and the same first-order
gradients for every way the modules use it -- weight- and activation-modulated StyledConv, the transposed
(upsampling) convolution, ToRGB with its bias, the discriminator's ConvLayers / skip convolutions.  CPU, fp64, kernel
stand-ins (the CUDA kernels are checked against the same stand-ins in tests/test_kernels_gpu.py)."""
import pytest
import torch

from gan_control_b200 import modules as M
from gan_control_b200 import ops
from golden_io import max_rel

F64 = torch.float64


def _randomize(net):
    for m in net.modules():
        if isinstance(m, M.NoiseInjection):
            m.weight.data.fill_(0.3)
        if isinstance(m, M.FusedLeakyReLU):
            m.bias.data.normal_()
        if isinstance(m, M.ToRGB):
            m.bias.data.normal_()


def _run(g, d, z, noise, fused):
    g.zero_grad()
    d.zero_grad()
    z = z.clone().requires_grad_(True)
    with (ops.first_order() if fused else torch.enable_grad()):
        img, _ = g([z], noise=noise)
        pred, _ = d(img)
        torch.nn.functional.softplus(-pred).mean().backward()
    grads = {k: v.grad.clone() for k, v in list(g.named_parameters()) + [('d.' + k, v) for k, v in d.named_parameters()]
             if v.grad is not None}
    return img.detach(), pred.detach(), z.grad.clone(), grads


@pytest.mark.parametrize('form', ['weight', 'activation', 'auto'])
def test_fused_networks_match_unfused(cpu_kernels, form):
    torch.manual_seed(0)
    size, sdim = 32, 32
    g = M.Generator(size, sdim, 2, channel_multiplier=2, conv_transpose=True, act_dtype=F64).double()
    d = M.Discriminator(size, channel_multiplier=2, act_dtype=F64).double()
    _randomize(g)
    _randomize(d)
    for m in g.modules():
        if isinstance(m, M.ModulatedConv2d):
            m.form = form
    z = torch.randn(2, sdim, dtype=F64)
    noise = [torch.randn(2, 1, 2 ** ((i + 5) // 2), 2 ** ((i + 5) // 2), dtype=F64) for i in range(g.num_layers)]
    calls = {'n': 0}
    real = cpu_kernels.modweight_fwd

    def counting(*a, **k):
        calls['n'] += 1
        return real(*a, **k)
    from gan_control_b200 import kernels
    kernels.modweight_fwd = counting
    try:
        img0, pred0, gz0, g0 = _run(g, d, z, noise, False)
        assert calls['n'] == 0
        img1, pred1, gz1, g1 = _run(g, d, z, noise, True)
        assert calls['n'] > 20                                   # every conv of G and D went through mod_conv
    finally:
        kernels.modweight_fwd = real
    assert max_rel(img1, img0) < 1e-12 and max_rel(pred1, pred0) < 1e-12 and max_rel(gz1, gz0) < 1e-8
    assert set(g0) == set(g1)
    for k in g0:
        assert max_rel(g1[k], g0[k]) < 1e-7, k


def test_no_grad_forward_uses_the_fused_path_and_double_backward_is_refused(cpu_kernels):
    torch.manual_seed(1)
    m = M.StyledConv(8, 8, 3, 16, conv_transpose=True).double()
    m.conv.form = 'weight'
    x = torch.randn(2, 8, 8, 8, dtype=F64).contiguous(memory_format=torch.channels_last)
    s = torch.randn(2, 16, dtype=F64)
    nz = torch.randn(2, 1, 8, 8, dtype=F64)
    with torch.no_grad():
        y_fast = m(x, s, noise=nz)
    y_ref = m(x, s, noise=nz)
    assert max_rel(y_fast, y_ref) < 1e-12
    xg = x.clone().requires_grad_(True)
    with ops.first_order():
        y = m(xg, s, noise=nz)
        with pytest.raises(RuntimeError, match='first-order only'):
            torch.autograd.grad(y.sum(), xg, create_graph=True)


@pytest.mark.parametrize('cfg', [dict(demod=True, flip=False), dict(demod=True, flip=True), dict(demod=False, flip=False)])
def test_modweight_standin_matches_reference_formula(cpu_kernels, cfg):
    """the stand-in (the contract of b200gan_modweight_fwd / _bwd) == gm.py:284-289 written out, incl. gradients"""
    torch.manual_seed(2)
    b, oc, ic, k = 3, 5, 4, 3
    w = torch.randn(oc, ic, k, k, dtype=F64, requires_grad=True)
    s = torch.randn(b, ic, dtype=F64, requires_grad=True)
    scale = 0.37
    wk, wkt, d = cpu_kernels.modweight_fwd(w, s, scale, cfg['demod'], cfg['flip'], F64, want_adjoint=True)
    weight = scale * w.unsqueeze(0) * s.view(b, 1, ic, 1, 1)                      # gm.py:285
    if cfg['demod']:
        demod = torch.rsqrt(weight.pow(2).sum([2, 3, 4]) + 1e-8)                  # gm.py:288
        weight = weight * demod.view(b, oc, 1, 1, 1)
        assert max_rel(d, demod) < 1e-13
    if cfg['flip']:
        weight = weight.flip(3, 4)
    assert max_rel(wk, weight.permute(0, 3, 4, 1, 2)) < 1e-13
    assert max_rel(wkt, ops._flip_t(weight).permute(0, 3, 4, 1, 2)) < 1e-13
    g = torch.randn(b, k, k, oc, ic, dtype=F64)
    gw_ref, gs_ref = torch.autograd.grad(weight.permute(0, 3, 4, 1, 2), (w, s), g)
    gs, gw = cpu_kernels.modweight_bwd(g, w.detach(), s.detach(), d, scale, cfg['demod'], cfg['flip'])
    assert max_rel(gs, gs_ref) < 1e-12 and max_rel(gw, gw_ref) < 1e-12


def test_fused_resblock_with_single_pass_downsampling(cpu_kernels):
    """ResBlock(32 -> 64): conv2 is the single-pass FIR (*) stride-2 form (space-to-depth view), so in the fused block
    conv1's activation backward rides in conv2's data-gradient epilogue and from_rgb's in conv1's: discriminator output
    and every parameter / input gradient equal the per-layer op algebra."""
    torch.manual_seed(3)
    d = M.Discriminator(128, channel_multiplier=0.25, act_dtype=F64).double()
    _randomize(d)
    assert d.convs[1].conv2.fuse_down and not d.convs[2].conv2.fuse_down
    x0 = torch.randn(2, 3, 128, 128, dtype=F64)
    out = []
    for fused in (False, True):
        d.zero_grad()
        x = x0.clone().requires_grad_(True)
        with (ops.first_order() if fused else torch.enable_grad()):
            pred, _ = d(x)
            torch.nn.functional.softplus(-pred).mean().backward()
        out.append((pred.detach(), x.grad.clone(), {k: v.grad.clone() for k, v in d.named_parameters()}))
    (p0, gx0, g0), (p1, gx1, g1) = out
    assert max_rel(p1, p0) < 1e-12 and max_rel(gx1, gx0) < 1e-8
    for k in g0:
        assert max_rel(g1[k], g0[k]) < 1e-8, k
    # frozen discriminator (the generator step): only the input gradient, no parameter gradients
    for p in d.parameters():
        p.requires_grad_(False)
    x = x0.clone().requires_grad_(True)
    with ops.first_order():
        torch.nn.functional.softplus(-d(x)[0]).mean().backward()
    assert max_rel(x.grad, gx0) < 1e-8
