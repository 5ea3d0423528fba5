"""ADA augmentation on the GPU: the fused warp + colour kernel against its contract stand-in, and `augment` end to end
against goldens produced by the unmodified reference (`oracle/make_golden_ada.py`)."""
import math

import numpy as np
import pytest
import torch

from gan_control_b200 import augment as A
from gan_control_b200 import kernels as K
from golden_io import Fixture, max_rel, rel_err
from oracle import kernels_ref as R

pytestmark = pytest.mark.gpu
DEV = 'cuda'


def _mats(n, h, w, oh, ow, seed):
    """rotations / scales / shifts around the image centre, some sampling far outside (zero padding)"""
    rng = np.random.default_rng(seed)
    out = []
    for i in range(n):
        th, s = rng.uniform(-math.pi, math.pi), rng.uniform(0.5, 1.6)
        a, b = s * math.cos(th), -s * math.sin(th)
        tx, ty = rng.uniform(-0.2, 0.2) * w, rng.uniform(-0.2, 0.2) * h
        cx, cy, ocx, ocy = (w - 1) / 2, (h - 1) / 2, (ow - 1) / 2, (oh - 1) / 2
        out.append([a, b, cx + tx - a * ocx - b * ocy, -b, a, cy + ty + b * ocx - a * ocy])
    out[0] = [1.0, 0.0, 0.0, 0.0, 1.0, 0.0]                       # identity (integer coordinates: weights exactly 1 / 0)
    return torch.tensor(out, dtype=torch.float64)


@pytest.mark.parametrize('dt', [torch.float32, torch.bfloat16, torch.float16])
@pytest.mark.parametrize('layout', ['planar', 'channels_last'])
@pytest.mark.parametrize('shape', [(3, 3, 37, 29, 41, 33), (2, 3, 64, 64, 64, 64), (2, 1, 16, 20, 9, 50), (1, 4, 8, 8, 8, 8)])
def test_affine_color_kernel(shape, layout, dt):
    n, c, h, w, oh, ow = shape
    rng = np.random.default_rng(5)
    x = torch.from_numpy(rng.standard_normal((n, c, h, w))).to(dt)
    gy = torch.from_numpy(rng.standard_normal((n, c, oh, ow))).to(dt)
    mat = _mats(n, h, w, oh, ow, 6)
    color = torch.from_numpy(rng.standard_normal((n, c, c + 1))).float()
    fmt = torch.channels_last if layout == 'channels_last' else torch.contiguous_format
    xd, gyd = x.to(DEV).contiguous(memory_format=fmt), gy.to(DEV).contiguous(memory_format=fmt)
    tol = {torch.float32: 2e-5, torch.bfloat16: 4e-3, torch.float16: 1e-3}[dt]
    for col in (color, None):
        y = K.affine_color_fwd(xd, mat.to(DEV), None if col is None else col.to(DEV), oh, ow)
        assert y.shape == (n, c, oh, ow) and y.dtype == dt
        want = R._affine_color(x.double(), mat, None if col is None else col.double(), oh, ow)
        assert float((y.cpu().double() - want).abs().max() / want.abs().max()) < tol
        gx = K.affine_color_bwd(gyd, mat.to(DEV), None if col is None else col.to(DEV), h, w)
        gwant = R.affine_color_bwd(gy.double(), mat, None if col is None else col.double(), h, w)
        assert gx.dtype == torch.float32 and float((gx.cpu().double() - gwant).abs().max() / gwant.abs().max()) < 2e-5
    y_id = K.affine_color_fwd(xd[:1], mat[:1].to(DEV), None, min(h, oh), min(w, ow))
    assert torch.equal(y_id.cpu(), x[:1, :, :min(h, oh), :min(w, ow)])            # identity map: exact copy


@pytest.mark.parametrize('name', ['a', 'b', 'c'])
def test_augment_matches_reference_goldens_on_gpu(name):
    fx = Fixture('ada')
    img = fx.t(name + '.img', torch.float32, DEV).requires_grad_(True)
    G, C = fx.t(name + '.G', torch.float32), fx.t(name + '.C', torch.float32)
    p = float(fx.np(name + '.cfg')[1])
    n0 = K.launch_count()
    y, _ = A.augment(img, p, (G, C))
    launches = K.launch_count() - n0
    assert 3 <= launches <= 5, launches                        # interpolating FIR, warp + colour, decimating FIR
    e = max_rel(y, fx.t(name + '.y'))
    gx, = torch.autograd.grad((y * fx.t(name + '.cot', torch.float32, DEV)).sum(), img)
    eg = max_rel(gx, fx.t(name + '.gx'))
    print(f'ADA augment {name}: fp32 image max-rel {e:.2e}, input gradient max-rel {eg:.2e} vs the reference')
    assert e < 1e-3 and eg < 1e-3                                 # north_star tolerance
    # bf16 channels-last (what the generator hands to the discriminator in the throughput configuration)
    yb, _ = A.augment(img.detach().to(torch.bfloat16).contiguous(memory_format=torch.channels_last), p, (G, C))
    assert yb.dtype == torch.bfloat16 and rel_err(yb, fx.t(name + '.y')) < 2e-2


def test_augment_draws_its_own_transforms_on_gpu():
    torch.manual_seed(3)
    img = torch.randn(4, 3, 64, 64, device=DEV)
    y, (G, C) = A.augment(img, 0.7)
    assert y.shape == img.shape and G.shape == (4, 3, 3) and C.shape == (4, 4, 4) and torch.isfinite(y).all()
    y2, _ = A.augment(img, 0.7, (G, C))
    assert torch.equal(y, y2)


def test_train_step_with_ada_on_gpu():
    """the throughput configuration (bf16, channels-last fake images, fp32 NCHW real images) through the augmented eager
    step, incl. the regularisation iteration; the controller adapts p from the real predictions"""
    import copy
    from gan_control_b200 import modules as M
    from gan_control_b200.train_step import GanTrainStep
    torch.manual_seed(0)
    g = M.Generator(64, 64, 3, channel_multiplier=0.5, conv_transpose=True, act_dtype=torch.bfloat16).to(DEV)
    d = M.Discriminator(64, channel_multiplier=0.5, act_dtype=torch.bfloat16).to(DEV)
    ctl = A.AdaptiveP(p=0.0, ada_target=0.6, ada_length=2000)
    step = GanTrainStep(g, d, copy.deepcopy(g), batch=64, latent_size=64, ada=ctl)
    real = torch.randn(64, 3, 64, 64, device=DEV).clamp_(-1, 1)
    ps = []
    for i in range(5):
        d_loss, g_loss = step.train_step(i, real, regularize=True)
        ps.append(ctl.p)
        assert math.isfinite(float(d_loss)) and math.isfinite(float(g_loss))
    # 64 predictions per iteration: the first update comes with the 4th, by +-(0.6 / 2000) * 256, clipped at 0
    assert ps[:3] == [0.0, 0.0, 0.0] and (ps[3] == 0.0 or abs(ps[3] - 0.6 / 2000 * 256) < 1e-9)
    assert ctl.count == 64 and -1.0 <= ctl.r_t <= 1.0
    assert all(torch.isfinite(p.grad).all() for p in g.parameters() if p.grad is not None)
