"""Shape tables and seeded parameter sets for the oracle (TEST INFRASTRUCTURE).

The reference ships no trained weights, and default init zeroes every noise
strength and bias, which would hide those terms from a parity
test.  ``seeded_state_dict`` therefore draws *every* tensor from a numpy
``default_rng`` stream keyed on (seed, key index): fixtures only need to store
seeds and outputs, and both the reference run (``make_golden.py``) and the
tests rebuild bit-identical parameters.

Shape tables restate ``Generator.__init__`` (gm:505-617, channels gm:552-563),
``create_fc_stack`` (gm:658-681) and ``Discriminator.__init__`` (gm:925-988).
``make_golden.py`` asserts they equal the reference ``state_dict`` exactly.
"""
import math

import numpy as np
import torch


def channel_table(channel_multiplier=2):
    return {4: 512, 8: 512, 16: 512, 32: 512,
            64: int(256 * channel_multiplier), 128: int(128 * channel_multiplier),
            256: int(64 * channel_multiplier), 512: int(32 * channel_multiplier),
            1024: int(16 * channel_multiplier)}


def generator_shapes(size, style_dim=512, n_mlp=8, channel_multiplier=2, fc_groups=None, mid_dim=256):
    ch = channel_table(channel_multiplier)
    log_size = int(math.log2(size))
    s = {}
    if fc_groups is None:
        for i in range(1, n_mlp + 1):
            s[f'style.{i}.weight'] = (style_dim, style_dim)
            s[f'style.{i}.bias'] = (style_dim,)
    else:
        for name, lo, hi in fc_groups:
            g = hi - lo
            for i in range(1, n_mlp + 1):
                d0 = g if i == 1 else mid_dim
                d1 = g if i == n_mlp else mid_dim
                s[f'style.{name}.{i}.weight'] = (d1, d0)
                s[f'style.{name}.{i}.bias'] = (d1,)
    s['input.input'] = (1, ch[4], 4, 4)

    def styled(prefix, ic, oc, up):
        s[prefix + 'conv.weight'] = (1, oc, ic, 3, 3)
        s[prefix + 'conv.modulation.weight'] = (ic, style_dim)
        s[prefix + 'conv.modulation.bias'] = (ic,)
        if up:
            s[prefix + 'conv.blur.kernel'] = (4, 4)
        s[prefix + 'noise.weight'] = (1,)
        s[prefix + 'activate.bias'] = (oc,)

    def rgb(prefix, ic, up):
        if up:
            s[prefix + 'upsample.kernel'] = (4, 4)
        s[prefix + 'conv.weight'] = (1, 3, ic, 1, 1)
        s[prefix + 'conv.modulation.weight'] = (ic, style_dim)
        s[prefix + 'conv.modulation.bias'] = (ic,)
        s[prefix + 'bias'] = (1, 3, 1, 1)

    styled('conv1.', ch[4], ch[4], False)
    rgb('to_rgb1.', ch[4], False)
    ic = ch[4]
    for r, i in enumerate(range(3, log_size + 1)):
        oc = ch[2 ** i]
        styled(f'convs.{2 * r}.', ic, oc, True)
        styled(f'convs.{2 * r + 1}.', oc, oc, False)
        rgb(f'to_rgbs.{r}.', oc, True)
        ic = oc
    for layer in range(2 * (log_size - 2) + 1):
        res = 2 ** ((layer + 5) // 2)
        s[f'noises.noise_{layer}'] = (1, 1, res, res)
    return s


def discriminator_shapes(size, channel_multiplier=2, in_channels=3):
    ch = channel_table(channel_multiplier)
    log_size = int(math.log2(size))
    s = {'convs.0.0.weight': (ch[size], in_channels, 1, 1), 'convs.0.1.bias': (ch[size],)}
    ic = ch[size]
    for j, i in enumerate(range(log_size, 2, -1), start=1):
        oc = ch[2 ** (i - 1)]
        p = f'convs.{j}.'
        s[p + 'conv1.0.weight'] = (ic, ic, 3, 3)
        s[p + 'conv1.1.bias'] = (ic,)
        s[p + 'conv2.0.kernel'] = (4, 4)
        s[p + 'conv2.1.weight'] = (oc, ic, 3, 3)
        s[p + 'conv2.2.bias'] = (oc,)
        s[p + 'skip.0.kernel'] = (4, 4)
        s[p + 'skip.1.weight'] = (oc, ic, 1, 1)
        ic = oc
    s['final_conv.0.weight'] = (ch[4], ic + 1, 3, 3)
    s['final_conv.1.bias'] = (ch[4],)
    s['final_linear.0.weight'] = (ch[4], ch[4] * 16)
    s['final_linear.0.bias'] = (ch[4],)
    s['final_linear.1.weight'] = (1, ch[4])
    s['final_linear.1.bias'] = (1,)
    return s


def fc_stack_shapes(n_mlp, in_dim, mid_dim, out_dim, prefix='fc_stack.'):
    """controller ``FcStack`` (controller_model.py:24-43)."""
    s = {}
    for i in range(n_mlp):
        d0 = in_dim if i == 0 else mid_dim
        d1 = out_dim if i == n_mlp - 1 else mid_dim
        s[f'{prefix}{i}.weight'] = (d1, d0)
        s[f'{prefix}{i}.bias'] = (d1,)
    return s


def _fir4(gain):
    k = np.outer([1., 3., 3., 1.], [1., 3., 3., 1.])
    return (k / k.sum() * gain).astype(np.float32)


def seeded_state_dict(shapes, seed, lr_mlp=0.01, dtype=torch.float32):
    """Deterministic, everywhere-non-trivial parameters for a shape table."""
    out = {}
    for idx, key in enumerate(sorted(shapes)):
        shape = shapes[key]
        if key.endswith('kernel'):
            gain = 4.0 if ('blur' in key or 'upsample' in key) else 1.0   # gm:76,119-120
            out[key] = torch.from_numpy(_fir4(gain)).to(dtype)
            continue
        rng = np.random.default_rng([seed, idx])
        n = torch.from_numpy(rng.standard_normal(shape).astype(np.float32))
        mapping = key.startswith('style.') or key.startswith('fc_stack.')
        if key.endswith('modulation.bias'):
            v = 1.0 + 0.1 * n                      # bias_init=1 (gm:271)
        elif key.endswith('noise.weight'):
            v = 0.1 + 0.05 * n                     # zero at init; make it matter
        elif key.endswith('bias'):
            v = (0.1 / lr_mlp if mapping else 0.1) * n
        elif key.endswith('weight'):
            v = n / lr_mlp if mapping else n       # randn / lr_mul (gm:176)
        else:                                      # input.input, noises.*
            v = n
        out[key] = v.to(dtype)
    return out
