"""Import the UNMODIFIED reference (`/root/reference/src`) on CPU.

TEST INFRASTRUCTURE ONLY, and only usable in the build container: the GPU box
has no `/root/reference`.  Used by `oracle/make_golden.py` (fixture generation)
and by `tests/test_oracle_vs_reference.py` (skipped when the tree is absent).

Three modules of the reference do not import in this image;
they are stubbed in `sys.modules` so that
`gan_control.trainers.generator_trainer` (the loss functions) can be imported.
"""
import os
import sys
import types

REF_SRC = os.environ.get('GAN_CONTROL_REF', '/root/reference/src')


def available():
    return os.path.isfile(os.path.join(REF_SRC, 'gan_control', 'models', 'gan_model.py'))


def _stub(name, **attrs):
    m = types.ModuleType(name)
    m.__dict__.update(attrs)
    sys.modules[name] = m
    return m


def import_reference():
    """Returns (gan_model module, GeneratorTrainer class or None)."""
    if not available():
        raise ImportError(f'reference tree not found at {REF_SRC}')
    if REF_SRC not in sys.path:
        sys.path.insert(0, REF_SRC)
    import logging
    logging.disable(logging.INFO)
    import gan_control.models.gan_model as gm
    trainer = None
    try:
        if 'gan_control.trainers.non_leaking' not in sys.modules:
            _stub('gan_control.trainers.non_leaking', augment=lambda img, p: (img, None))
        if 'gan_control.fid_utils.calc_inception' not in sys.modules:
            _stub('gan_control.fid_utils.calc_inception', load_patched_inception_v3=lambda: None)
        if 'gan_control.evaluation.tracker' not in sys.modules:
            _stub('gan_control.evaluation.tracker', Tracker=object)
        from gan_control.trainers.generator_trainer import GeneratorTrainer
        trainer = GeneratorTrainer
    except Exception:  # the trainer is optional for fixture generation
        trainer = None
    return gm, trainer
