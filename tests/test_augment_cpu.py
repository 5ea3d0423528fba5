"""ADA augmentation: `gan_control_b200.augment` against the reference's `non_leaking.augment`.
CPU: the host logic (transform sampling, padding / grid geometry, colour folding) with the kernels replaced by their
contract stand-ins, against goldens the unmodified reference produced (`oracle/make_golden_ada.py`), and -- where the
reference tree is present -- draw-for-draw equality of the sampled transforms."""
import numpy as np
import pytest
import torch

from golden_io import Fixture, max_rel
from oracle.ref_import import available

CASES = ['a', 'b', 'c']


@pytest.mark.parametrize('name', CASES)
def test_augment_matches_reference_goldens(cpu_kernels, name):
    from gan_control_b200 import augment as A
    fx = Fixture('ada')
    img = fx.t(name + '.img').double().requires_grad_(True)
    G, C = fx.t(name + '.G', torch.float32), fx.t(name + '.C', torch.float32)
    p = float(fx.np(name + '.cfg')[1])
    y, (G2, C2) = A.augment(img, p, (G, C))
    assert y.shape == img.shape and G2 is G and C2 is C
    # the reference evaluates its sampling grid from an fp32 linspace: coordinates agree to ~1e-5 pixel
    assert max_rel(y, fx.t(name + '.y')) < 1e-4
    gx, = torch.autograd.grad((y * fx.t(name + '.cot').double()).sum(), img)
    assert max_rel(gx, fx.t(name + '.gx')) < 1e-4
    ya, _ = A.random_apply_affine(img.detach(), p, G)
    assert max_rel(ya, fx.t(name + '.y_affine')) < 1e-4
    yc, _ = A.random_apply_color(img.detach(), p, C)
    assert max_rel(yc, fx.t(name + '.y_color')) < 1e-6


def test_seeded_draws_reproduce_the_golden_transforms():
    """same seed -> the same G and C as the reference drew (the draws are requested in the reference's order)"""
    from gan_control_b200 import augment as A
    fx = Fixture('ada')
    for name in CASES:
        seed, p = fx.np(name + '.cfg')
        img = fx.t(name + '.img')
        torch.manual_seed(int(seed))
        G = A._first_admissible_affine(img, float(p), (len(A.SYM6) + 1) // 2)
        C = A.sample_color(float(p), img.shape[0])
        assert torch.allclose(G, fx.t(name + '.G', torch.float32), rtol=0, atol=1e-6), name
        assert torch.allclose(C, fx.t(name + '.C', torch.float32), rtol=0, atol=1e-6), name


@pytest.mark.skipif(not available(), reason='reference tree not present on this machine')
@pytest.mark.parametrize('p', [0.3, 1.0])
def test_samplers_draw_for_draw_with_the_reference(p):
    from gan_control_b200 import augment as A
    from oracle.make_golden_ada import load_non_leaking
    from oracle.ref_import import import_reference
    gm, _ = import_reference()
    nl = load_non_leaking(gm)
    for seed in range(5):
        torch.manual_seed(seed)
        g_ref, c_ref = nl.sample_affine(p, 6, 48, 32), nl.sample_color(p, 6)
        torch.manual_seed(seed)
        g_new, c_new = A.sample_affine(p, 6, 48, 32), A.sample_color(p, 6)
        assert torch.equal(g_ref, g_new) and torch.equal(c_ref[:, :3], c_new[:, :3])
        assert nl.get_padding(torch.inverse(g_ref), 48, 32) == A._padding_for(torch.inverse(g_new), 48, 32)


def test_adaptive_p_controller():
    """generator_trainer.py:669-687 restated: r_t over >= 256 predictions, p moves by target / length per image"""
    from gan_control_b200.augment import AdaptiveP
    ctl = AdaptiveP(p=0.0, ada_target=0.6, ada_length=1000)
    pred = torch.ones(64, 1)
    for i in range(3):
        assert ctl.update(pred) == 0.0                       # fewer than 256 predictions: unchanged
    p = ctl.update(pred)                                     # 256 predictions, r_t = 1 > target: p += step * 256
    assert abs(p - 0.6 / 1000 * 256) < 1e-12 and ctl.r_t == 1.0 and ctl.count == 0
    for i in range(4):
        p = ctl.update(-pred)
    assert p == 0.0 and ctl.r_t == -1.0                      # clipped at 0
    fixed = AdaptiveP(p=0.25)
    for i in range(8):
        assert fixed.update(pred) == 0.25                    # a configured p > 0 is not adapted


def test_train_step_with_ada(cpu_kernels):
    """generator_trainer.py:421-422, 647-653, 669-687 in `GanTrainStep`: the real and the fake batch of the discriminator
    step and the fake batch of the generator step go through `augment` with the controller's current p (the regularisers
    do not); the controller sees the real predictions; CUDA-graph capture refuses."""
    import copy
    import gan_control_b200.augment as A
    from gan_control_b200 import modules as M
    from gan_control_b200.train_step import GanTrainStep
    from oracle import params as P
    size, sdim, batch = 32, 16, 4
    f64 = torch.float64
    g = M.Generator(size, sdim, 2, channel_multiplier=0.25, conv_transpose=True, act_dtype=f64).double()
    g.load_state_dict(P.seeded_state_dict(P.generator_shapes(size, sdim, 2, 0.25), 5, dtype=f64))
    d = M.Discriminator(size, channel_multiplier=0.25, act_dtype=f64).double()
    d.load_state_dict(P.seeded_state_dict(P.discriminator_shapes(size, 0.25), 6, dtype=f64))
    real = torch.from_numpy(np.random.default_rng(3).standard_normal((batch, 3, size, size))).clamp(-1, 1)
    step = GanTrainStep(g, d, copy.deepcopy(g), batch=batch, latent_size=sdim, ada=0.5)
    assert isinstance(step.ada, A.AdaptiveP) and step.ada.p == 0.5 and not step.ada.adaptive
    calls, orig = [], A.augment

    def spy(img, p, transform_matrix=(None, None)):
        out = orig(img, p, transform_matrix)
        calls.append((tuple(img.shape), p, img.requires_grad, torch.is_grad_enabled()))
        return out
    A.augment = spy
    try:
        torch.manual_seed(12)
        g0, d0 = step.g_arena.data.clone(), step.d_arena.data.clone()
        d_loss, g_loss = step.train_step(0, real, regularize=True)          # iteration 0: both regularisers run as well
    finally:
        A.augment = orig
    shape = (batch, 3, size, size)
    # D step: real + fake without autograd; G step: fake with the generator's graph attached; R1 / path length: none
    assert calls == [(shape, 0.5, False, False), (shape, 0.5, False, False), (shape, 0.5, True, True)]
    assert np.isfinite(float(d_loss)) and np.isfinite(float(g_loss)) and step.ada.count == batch
    assert float((step.g_arena.data - g0).abs().max()) > 0 and float((step.d_arena.data - d0).abs().max()) > 0
    with pytest.raises(RuntimeError, match='eagerly'):
        step.capture(shape)
