"""The fused single-pass forms of the resampling convolutions (composite FIR (*) conv weights between
space-to-depth views, ops.composite_up / composite_down) against the two-pass forms that the reference-golden tests
pin (tests/test_host_algebra_cpu.py): outputs, first-order gradients and the path-length / R1 style double backward,
in fp64 with the kernel contract stand-ins."""
import pytest
import torch

from gan_control_b200 import modules as M, ops

F64 = torch.float64


def max_rel(a, b):
    a, b = a.detach(), b.detach()
    return float((a - b).abs().max() / b.abs().max().clamp_min(1e-300))


def _styled_up(seed, ic, oc, sdim, fused):
    torch.manual_seed(seed)
    m = M.StyledConv(ic, oc, 3, sdim, upsample=True, conv_transpose=True).double()
    m.conv.form = 'weight'
    m.noise.weight.data.fill_(0.3)
    m.activate.bias.data.normal_()
    if not fused:
        m.conv.FUSE_UP_MAX_IN_CHANNELS = 0
    return m


@pytest.mark.parametrize('ic,oc', [(16, 8), (8, 4)])
def test_fused_up_styled_conv(cpu_kernels, ic, oc):
    sdim, b, h = 12, 2, 6
    torch.manual_seed(5)
    x0 = torch.randn(b, ic, h, h, dtype=F64).contiguous(memory_format=torch.channels_last)
    s0 = torch.randn(b, sdim, dtype=F64)
    noise = torch.randn(b, 1, 2 * h, 2 * h, dtype=F64)
    gy = torch.randn(b, oc, 2 * h, 2 * h, dtype=F64)
    res = []
    for fused in (True, False):
        m = _styled_up(3, ic, oc, sdim, fused)
        assert m.conv.fuses_up(h, h) == fused
        x, s = x0.clone().requires_grad_(True), s0.clone().requires_grad_(True)
        ps = (x, s, m.conv.weight, m.conv.modulation.weight, m.noise.weight, m.activate.bias)
        y = m(x, s, noise=noise)
        g = torch.autograd.grad(y, ps, gy, create_graph=True)
        pl = g[1].pow(2).sum() + g[0].pow(2).sum()                          # path-length / R1 style second-order objective
        gg = torch.autograd.grad(pl, (x, s, m.conv.weight, m.conv.modulation.weight))
        g1 = torch.autograd.grad(m(x, s, noise=noise), ps, gy)             # first-order only: fused backward branches
        res.append((y, g, gg, g1))
    (y, g, gg, g1), (yr, gr, ggr, g1r) = res
    assert max_rel(y, yr) < 1e-12
    for a, r in zip(g, gr):
        assert max_rel(a, r) < 1e-11
    for a, r in zip(gg, ggr):
        assert max_rel(a, r) < 1e-10
    for a, r in zip(g1, g1r):
        assert max_rel(a, r) < 1e-11


@pytest.mark.parametrize('ic,oc,h', [(8, 16, 8), (4, 12, 6)])
def test_fused_down_conv_layer(cpu_kernels, ic, oc, h):
    torch.manual_seed(7)
    x0 = torch.randn(2, ic, h, h, dtype=F64).contiguous(memory_format=torch.channels_last)
    gy = torch.randn(2, oc, h // 2, h // 2, dtype=F64)
    res = []
    for fused in (True, False):
        torch.manual_seed(11)
        m = M.ConvLayer(ic, oc, 3, downsample=True).double()
        m[2].bias.data.normal_()
        assert m.fuse_down
        m.fuse_down = fused
        x = x0.clone().requires_grad_(True)
        ps = (x, m[1].weight, m[2].bias)
        y = m(x, out_scale=0.7)
        g = torch.autograd.grad(y, ps, gy, create_graph=True)
        r1 = g[0].pow(2).sum()
        gg = torch.autograd.grad(r1, (x, m[1].weight))
        with ops.data_grads_only():
            gx_only = torch.autograd.grad(m(x, out_scale=0.7), x, gy, create_graph=True)[0]
        g1 = torch.autograd.grad(m(x, out_scale=0.7), ps, gy)
        res.append((y, g, gg, g1, gx_only))
    (y, g, gg, g1, gxo), (yr, gr, ggr, g1r, gxor) = res
    assert max_rel(y, yr) < 1e-12
    for a, r in zip(g + (gxo,), gr + (gxor,)):
        assert max_rel(a, r) < 1e-11
    for a, r in zip(gg, ggr):
        assert max_rel(a, r) < 1e-10
    for a, r in zip(g1, g1r):
        assert max_rel(a, r) < 1e-11


def test_composite_weights_shapes(cpu_kernels):
    k = M.make_kernel([1, 3, 3, 1])
    w = torch.randn(2, 4, 8, 3, 3, dtype=F64)
    assert ops.composite_up(w, ops.fir_toeplitz(k.double() * 4, False)).shape == (2, 16, 8, 3, 3)
    assert ops.composite_down(w, ops.fir_toeplitz(k.double(), True)).shape == (2, 4, 32, 3, 3)


@pytest.mark.parametrize('dtype', [torch.float32, torch.bfloat16, torch.float64])
def test_kernel_layout(dtype):
    """the K-major weight operand: values, dtype and dense layout (kernels.conv_fwd asserts contiguity)"""
    w = torch.randn(3, 5, 4, 3, 3, dtype=torch.float32, requires_grad=True)
    for src in (w, w.flip(3, 4).transpose(1, 2)):
        k = ops._kernel_layout(src, dtype)
        ref = src.detach().permute(0, 3, 4, 1, 2).to(dtype).contiguous()
        assert k.dtype == dtype and k.is_contiguous() and not k.requires_grad and k.shape == ref.shape
        assert torch.equal(k, ref)
