"""Attach the B200 path to an unmodified gan-control checkout.

The reference cannot be switched in place (`gan_model.py:19-23`: `FUSED = True` raises), so the hook
is the one the reference itself implies: module-level names that every layer resolves at call /
construction time (`gm.py:87,127,192,400,885`) and the CUDA-op package location
`gan_control.models.op` (`trainers/non_leaking.py:6`).

    import gan_control_b200
    gan_control_b200.install()                    # before building Generator / Discriminator
    from gan_control.trainers.generator_trainer import GeneratorTrainer   # unchanged from here on

`install()` (1) registers `gan_control.models.op` exporting `upfirdn2d, FusedLeakyReLU,
fused_leaky_relu, conv2d_gradfix`, (2) replaces the operator names inside
`gan_control.models.gan_model`, (3) if `gan_control.trainers.non_leaking` is (or later gets) imported,
`install_augment()` swaps its `augment / random_apply_affine / random_apply_color` for the fused ones
(`gan_control_b200.augment`: same signatures, same random draws).  `state_dict` keys and shapes are unchanged, so checkpoints written by
either side load in the other.
"""
import sys
import types

from . import modules, ops

PATCHED_FUNCTIONS = ('upfirdn2d', 'fused_leaky_relu')
PATCHED_CLASSES = ('FusedLeakyReLU', 'ScaledLeakyReLU', 'PixelNorm', 'Upsample', 'Downsample', 'Blur', 'EqualConv2d',
                   'EqualLinear', 'ModulatedConv2d', 'NoiseInjection', 'ConstantInput', 'StyledConv', 'ToRGB',
                   'ConvLayer', 'ResBlock')


def op_package():
    """The module the reference expects at `gan_control.models.op` (and upstream's `op/`)."""
    op = types.ModuleType('gan_control.models.op')
    op.upfirdn2d = ops.upfirdn2d
    op.fused_leaky_relu = ops.fused_leaky_relu
    op.FusedLeakyReLU = modules.FusedLeakyReLU
    gradfix = types.ModuleType('gan_control.models.op.conv2d_gradfix')
    gradfix.conv2d = ops.conv2d
    gradfix.conv_transpose2d = ops.conv_transpose2d
    # upstream's switches: `with conv2d_gradfix.no_weight_gradients():` around the path-length inner backward
    gradfix.no_weight_gradients = ops.data_grads_only
    gradfix.enabled = True
    gradfix.weight_gradients_disabled = False
    op.conv2d_gradfix = gradfix
    return op, gradfix


def install(gan_model=None):
    """Patch `gan_control.models.gan_model` (imported here unless passed in). Returns the module."""
    op, gradfix = op_package()
    sys.modules['gan_control.models.op'] = op
    sys.modules['gan_control.models.op.conv2d_gradfix'] = gradfix
    if gan_model is None:
        import gan_control.models.gan_model as gan_model
    gan_model.upfirdn2d = ops.upfirdn2d
    gan_model.fused_leaky_relu = ops.fused_leaky_relu
    for name in PATCHED_CLASSES:
        setattr(gan_model, name, getattr(modules, name))
    gan_model.B200GAN_INSTALLED = True
    nl = sys.modules.get('gan_control.trainers.non_leaking')
    if nl is not None and hasattr(nl, 'random_apply_affine'):       # the real module (not a test stub) is loaded
        install_augment(nl)
    return gan_model


def install_augment(non_leaking=None):
    """Replace the image work of `gan_control.trainers.non_leaking` (ADA, :314-392) by the three-kernel path of
    `gan_control_b200.augment`; the samplers keep drawing on the host.  `generator_trainer` binds `augment` by name at
    import (`from ...non_leaking import augment`, gt.py:28), so call this before importing the trainer, or pass the
    trainer module to have its binding replaced as well."""
    from . import augment as A
    if non_leaking is None:
        import gan_control.trainers.non_leaking as non_leaking
    for name in ('augment', 'random_apply_affine', 'random_apply_color'):
        setattr(non_leaking, name, getattr(A, name))
    trainer = sys.modules.get('gan_control.trainers.generator_trainer')
    if trainer is not None and hasattr(trainer, 'augment'):
        trainer.augment = A.augment
    return non_leaking
