"""Generate `tests/golden/ada.npz`: the UNMODIFIED reference's `trainers/non_leaking.augment` on CPU (its own pure-PyTorch
`upfirdn2d`, F.grid_sample, apply_color), fp64 images, for seeded inputs -- the transforms it drew (G, C), the augmented
images and the gradient w.r.t. the input image.  Build container only (needs /root/reference):

    python oracle/make_golden_ada.py

TEST INFRASTRUCTURE: nothing here is product code."""
import importlib.util
import os
import sys
import types

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))

from oracle.ref_import import REF_SRC, import_reference   # noqa: E402

OUT = os.path.join(os.path.dirname(HERE), 'tests', 'golden')


def load_non_leaking(gm):
    """`non_leaking.py` imports `upfirdn2d` from the CUDA-op package the reference does not ship (:6): give it the
    reference's own FUSED=False implementation (gan_model.py:45-50)."""
    op = types.ModuleType('gan_control.models.op')
    op.upfirdn2d = gm.upfirdn2d
    saved = sys.modules.get('gan_control.models.op')
    sys.modules['gan_control.models.op'] = op
    try:
        spec = importlib.util.spec_from_file_location('ref_non_leaking_golden', os.path.join(REF_SRC, 'gan_control', 'trainers', 'non_leaking.py'))
        nl = importlib.util.module_from_spec(spec)
        spec.loader.exec_module(nl)
    finally:
        if saved is None:
            sys.modules.pop('gan_control.models.op', None)
        else:
            sys.modules['gan_control.models.op'] = saved
    return nl


CASES = [('a', 7, (2, 3, 32, 32), 1.0), ('b', 11, (3, 3, 40, 24), 0.8), ('c', 23, (2, 3, 48, 48), 0.5)]


def main():
    gm, _ = import_reference()
    nl = load_non_leaking(gm)
    out = {}
    for name, seed, shape, p in CASES:
        rng = np.random.default_rng(seed)
        img = torch.from_numpy(rng.standard_normal(shape).astype(np.float32)).double().requires_grad_(True)   # fp32-exact inputs
        cot = torch.from_numpy(rng.standard_normal(shape).astype(np.float32)).double()
        torch.manual_seed(seed)
        y, (G, C) = nl.augment(img, p)
        assert y.shape == img.shape
        gx, = torch.autograd.grad((y * cot).sum(), img)
        out.update({f'{name}.cfg': np.array([seed, p]), f'{name}.img': img.detach().numpy(), f'{name}.cot': cot.numpy(),
                    f'{name}.G': G.numpy(), f'{name}.C': C.numpy(), f'{name}.y': y.detach().numpy(), f'{name}.gx': gx.numpy()})
        # the colour stage alone and the geometric stage alone (explicit matrices)
        yc, _ = nl.random_apply_color(img.detach(), p, C)
        ya, _ = nl.random_apply_affine(img.detach(), p, G)
        out.update({f'{name}.y_color': yc.numpy(), f'{name}.y_affine': ya.numpy()})
    os.makedirs(OUT, exist_ok=True)
    path = os.path.join(OUT, 'ada.npz')
    # images as fp32 (inputs are fp32-exact; outputs are compared at >= 1e-5), matrices as drawn (fp32)
    np.savez_compressed(path, **{k: (v.astype(np.float32) if v.dtype == np.float64 and not k.endswith('.cfg') else v) for k, v in out.items()})
    print(f'ada: {len(out)} arrays, {os.path.getsize(path) / 1024:.1f} KiB')


if __name__ == '__main__':
    main()
