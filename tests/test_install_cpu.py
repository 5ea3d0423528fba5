"""`gan_control_b200.install()` on the UNMODIFIED reference (build container only: needs /root/reference):
the reference's own Generator / Discriminator classes, built after install(), run on the new operators and
reproduce the goldens; state_dict layout unchanged."""
import numpy as np
import pytest
import torch

from oracle import params as P
from oracle.ref_import import available, import_reference
from golden_io import Fixture, max_rel

pytestmark = pytest.mark.skipif(not available(), reason='reference tree not present on this machine')


def rnd(seed, *shape):
    return torch.from_numpy(np.random.default_rng(seed).standard_normal(shape))


def test_install_patches_reference(cpu_kernels):
    import sys
    gm, _ = import_reference()
    saved = {k: getattr(gm, k) for k in dir(gm) if not k.startswith('__')}
    try:
        import gan_control_b200
        from gan_control_b200 import modules as M
        gan_control_b200.install(gm)
        assert gm.StyledConv is M.StyledConv and 'gan_control.models.op' in sys.modules
        assert sys.modules['gan_control.models.op'].upfirdn2d is gm.upfirdn2d
        fx = Fixture('networks')
        size, sdim, n_mlp, seed = [int(v) for v in fx.np('g16.cfg')]
        g = gm.Generator(size, sdim, n_mlp, channel_multiplier=2, conv_transpose=True).double()   # the REFERENCE class
        assert isinstance(g.conv1, M.StyledConv) and isinstance(g.style[1], M.EqualLinear)
        shapes = P.generator_shapes(size, sdim, n_mlp, 2)
        assert {k: tuple(v.shape) for k, v in g.state_dict().items()} == {k: tuple(v) for k, v in shapes.items()}
        g.load_state_dict(P.seeded_state_dict(shapes, seed, dtype=torch.float64))
        z = rnd(seed * 10, 2, sdim)
        noise = [rnd(seed * 100 + i, 2, 1, 2 ** ((i + 5) // 2), 2 ** ((i + 5) // 2)) for i in range(g.num_layers)]
        img, _ = g([z], noise=noise)
        assert max_rel(img, fx.t('g16.f64.img')) < 1e-10
        d = gm.Discriminator(16, channel_multiplier=2).double()
        d.load_state_dict(P.seeded_state_dict(P.discriminator_shapes(16, 2), 21, dtype=torch.float64))
        pred, _ = d(rnd(210, 8, 3, 16, 16))
        assert max_rel(pred, fx.t('d16.f64.pred')) < 1e-10
    finally:
        for k, v in saved.items():
            setattr(gm, k, v)
        sys.modules.pop('gan_control.models.op', None)
        sys.modules.pop('gan_control.models.op.conv2d_gradfix', None)


def test_parameter_order_matches_reference():
    """`optim.Adam(model.parameters())` numbers parameters by registration order (gt.py:161-173): the optimiser entries
    of a checkpoint interchange only if the new modules register their parameters in the reference's order."""
    gm, _ = import_reference()
    from gan_control_b200 import modules as M
    for size in (16, 64):
        ref_g = gm.Generator(size, 32, 2, channel_multiplier=2, conv_transpose=True)
        new_g = M.Generator(size, 32, 2, channel_multiplier=2, conv_transpose=True)
        assert [(n, tuple(p.shape)) for n, p in ref_g.named_parameters()] == [(n, tuple(p.shape)) for n, p in new_g.named_parameters()]
        assert list(ref_g.state_dict()) == list(new_g.state_dict())
        ref_d, new_d = gm.Discriminator(size, channel_multiplier=2), M.Discriminator(size, channel_multiplier=2)
        assert [(n, tuple(p.shape)) for n, p in ref_d.named_parameters()] == [(n, tuple(p.shape)) for n, p in new_d.named_parameters()]


def test_ada_augment_runs_on_the_registered_op(cpu_kernels):
    """`trainers/non_leaking.py` imports `upfirdn2d` from `gan_control.models.op` (:6), the package
    install() registers; its 12x12-tap up=2 / down=2 calls (:338, :359) then run on the new operator and give the same
    augmented images as the reference's own pure-PyTorch upfirdn2d."""
    import importlib.util
    import os
    import sys
    from oracle.ref_import import REF_SRC
    gm, _ = import_reference()
    saved = {k: getattr(gm, k) for k in dir(gm) if not k.startswith('__')}
    ref_upfirdn2d = gm.upfirdn2d
    try:
        import gan_control_b200
        gan_control_b200.install(gm)
        spec = importlib.util.spec_from_file_location('ref_non_leaking_real', os.path.join(REF_SRC, 'gan_control', 'trainers', 'non_leaking.py'))
        nl = importlib.util.module_from_spec(spec)
        spec.loader.exec_module(nl)                      # `from gan_control.models.op import upfirdn2d` resolves to the new op
        assert nl.upfirdn2d is sys.modules['gan_control.models.op'].upfirdn2d and nl.upfirdn2d is not ref_upfirdn2d
        img = rnd(400, 2, 3, 32, 32)
        torch.manual_seed(7)
        out_new, _ = nl.augment(img, 1.0)
        nl.upfirdn2d = ref_upfirdn2d
        torch.manual_seed(7)
        out_ref, _ = nl.augment(img, 1.0)
        assert out_new.shape == out_ref.shape == img.shape
        assert max_rel(out_new, out_ref) < 1e-9
    finally:
        for k, v in saved.items():
            setattr(gm, k, v)
        sys.modules.pop('gan_control.models.op', None)
        sys.modules.pop('gan_control.models.op.conv2d_gradfix', None)


@pytest.mark.parametrize('name', ['ffhq', 'metfaces', 'afhq'])
def test_from_config_matches_reference_construction(name):
    """SURVEY.md 8(f) row 3, config wire format: `GanTrainStep.from_config` on the reference's shipped configs builds the
    networks `GeneratorTrainer.init_models_and_optim` builds (gt.py:120-157: same state_dict keys and shapes as the
    reference classes with the reference's own MiniBatchUtils fc_config) and the optimiser / regulariser settings of
    gt.py:158-173 (resolution reduced to keep the test light)."""
    import json
    import os
    from oracle.ref_import import REF_SRC
    from gan_control_b200.train_step import GanTrainStep
    gm, _ = import_reference()
    from gan_control.utils.mini_batch_multi_split_utils import MiniBatchUtils
    cfg = json.load(open(os.path.join(REF_SRC, 'gan_control', 'configs', name + '.json')))
    cfg['model_config']['size'] = 32
    mc, tc = cfg['model_config'], cfg['training_config']
    with pytest.raises(NotImplementedError, match='does not build'):     # attribute losses / ADA are not silently dropped
        GanTrainStep.from_config(cfg, device='cpu', world_size=4, act_dtype=torch.float32)
    with pytest.warns(UserWarning, match='IGNORED config terms'):
        step = GanTrainStep.from_config(cfg, device='cpu', world_size=4, act_dtype=torch.float32, vanilla_only=True)
    assert step.effective_objective['ignored_config_terms']
    # ADA is built (gan_control_b200.augment): enabled configs get the controller with the config's p / target / length
    aug = tc.get('augment') or {}
    assert (step.ada is not None) == bool(aug.get('enabled'))
    assert not any('augment' in t for t in step.effective_objective['ignored_config_terms'])
    if step.ada is not None:
        assert step.ada.target == aug['ada_target'] and ('ada_augment' in step.effective_objective['terms'])
    bu = MiniBatchUtils(tc['mini_batch'], tc['sub_groups_dict'], total_batch=tc['batch'])
    ref_g = gm.Generator(mc['size'], mc['latent_size'], mc['n_mlp'], channel_multiplier=mc['channel_multiplier'],
                         out_channels=mc['img_channels'], split_fc=mc['split_fc'], marge_fc=mc['marge_fc'],
                         fc_config=bu.get_fc_config(), conv_transpose=mc['conv_transpose'], noise_mode=mc['g_noise_mode'])
    ref_d = gm.Discriminator(mc['size'], channel_multiplier=mc['channel_multiplier'], in_channels=mc['img_channels'])
    shapes = lambda m: [(k, tuple(v.shape)) for k, v in m.state_dict().items()]
    assert shapes(step.g) == shapes(ref_g) and shapes(step.g_ema) == shapes(ref_g) and shapes(step.d) == shapes(ref_d)
    assert [n for n, _ in step.g.named_parameters()] == [n for n, _ in ref_g.named_parameters()]
    assert step.batch == tc['batch'] // 4 and step.global_batch == tc['batch']
    g_ratio, d_ratio = tc['g_reg_every'] / (tc['g_reg_every'] + 1), tc['d_reg_every'] / (tc['d_reg_every'] + 1)
    assert step.g_optim.lr == tc['lr_g'] * g_ratio and step.g_optim.betas == (0 ** g_ratio, 0.99 ** g_ratio)
    assert step.d_optim.lr == tc['lr_d'] * d_ratio and step.d_optim.betas == (0 ** d_ratio, 0.99 ** d_ratio)
    assert (step.r1, step.d_reg_every, step.g_reg_every, step.path_regularize, step.path_batch_shrink, step.mixing) == \
        (tc['r1'], tc['d_reg_every'], tc['g_reg_every'], tc['path_regularize'], tc['path_batch_shrink'], tc['mixing'])
    assert step.accum == 0.5 ** (tc['batch'] / tc['g_moving_average'])
    # g_ema starts as a copy of g (accumulate(g_ema, g, 0), gt.py:156)
    for (k, a), (_, b) in zip(step.g.named_parameters(), step.g_ema.named_parameters()):
        assert torch.equal(a, b), k


def test_conv2d_gradfix_surface(cpu_kernels):
    """upstream `op/conv2d_gradfix.py` surface (README.md:88-89): both functions with upstream's full signatures (the
    unused dilation / groups / output_padding only at their defaults), `no_weight_gradients()`, and the values of
    F.conv2d / F.conv_transpose2d incl. gradients."""
    import torch.nn.functional as F
    from gan_control_b200.install import op_package
    op, gradfix = op_package()
    assert op.conv2d_gradfix is gradfix and gradfix.enabled is True
    x = rnd(500, 2, 6, 9, 9).requires_grad_(True)
    w = rnd(501, 4, 6, 3, 3).requires_grad_(True)
    b = rnd(502, 4).requires_grad_(True)
    y = gradfix.conv2d(x, w, bias=b, stride=1, padding=1, dilation=1, groups=1)
    ref = F.conv2d(x, w, b, stride=1, padding=1)
    assert max_rel(y, ref) < 1e-12
    for a, r in zip(torch.autograd.grad(y.sum() + (y * y).sum(), (x, w, b)), torch.autograd.grad(ref.sum() + (ref * ref).sum(), (x, w, b))):
        assert max_rel(a, r) < 1e-11
    wt = rnd(503, 6, 4, 3, 3).requires_grad_(True)
    yt = gradfix.conv_transpose2d(x, wt, bias=None, stride=2, padding=0, output_padding=0, groups=1, dilation=1)
    reft = F.conv_transpose2d(x, wt, None, stride=2, padding=0)
    assert yt.shape == reft.shape and max_rel(yt, reft) < 1e-12
    with gradfix.no_weight_gradients():
        gx, = torch.autograd.grad(gradfix.conv2d(x, w, padding=1).sum(), x)
    assert max_rel(gx, torch.autograd.grad(F.conv2d(x, w, padding=1).sum(), x)[0]) < 1e-12
    for bad in (dict(dilation=2), dict(groups=2)):
        with pytest.raises(NotImplementedError):
            gradfix.conv2d(x, w, **bad)


def test_install_augment_replaces_the_reference_functions(cpu_kernels):
    """`install_augment` on the real `trainers/non_leaking.py`: its `augment` then IS the fused one and, drawing from the
    same seed, returns what the reference's own pipeline returns"""
    import gan_control_b200
    from gan_control_b200 import augment as A
    from oracle.make_golden_ada import load_non_leaking
    gm, _ = import_reference()
    ref = load_non_leaking(gm)                     # pristine copy for the expected values
    nl = load_non_leaking(gm)
    gan_control_b200.install_augment(nl)
    assert nl.augment is A.augment and nl.random_apply_affine is A.random_apply_affine
    img = rnd(600, 2, 3, 32, 32)
    torch.manual_seed(21)
    want, (g_ref, c_ref) = ref.augment(img, 0.9)
    torch.manual_seed(21)
    got, (g_new, c_new) = nl.augment(img, 0.9)
    assert torch.equal(g_ref, g_new) and torch.equal(c_ref[:, :3], c_new[:, :3])
    assert max_rel(got, want) < 1e-4
