"""Readers for tests/golden/*.npz (written by oracle/make_golden.py)."""
import os

import numpy as np
import torch

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), 'golden')


class Fixture:
    def __init__(self, name):
        self.z = np.load(os.path.join(GOLDEN, name + '.npz'))

    def keys(self, prefix=''):
        return [k for k in self.z.files if k.startswith(prefix)]

    def np(self, key):
        return self.z[key]

    def t(self, key, dtype=None, device=None):
        t = torch.from_numpy(np.array(self.z[key]))
        if dtype is not None and t.is_floating_point():
            t = t.to(dtype)
        return t.to(device) if device is not None else t

    def sub(self, prefix, dtype=None, device=None):
        """All arrays under `prefix` as a dict with the prefix stripped."""
        return {k[len(prefix):]: self.t(k, dtype, device) for k in self.keys(prefix)}


def reduce_like_golden(t, limit=8192):
    """Same reduction `make_golden.subsample` applied to big gradient tensors."""
    a = t.detach().double().cpu().reshape(-1)
    if a.numel() <= limit:
        return a
    return torch.cat([a[:: a.numel() // 4096][:4096], a.pow(2).sum().sqrt()[None]])


def rel_err(a, b):
    a = a.detach().double().cpu().reshape(-1)
    b = b.detach().double().cpu().reshape(-1)
    return float((a - b).norm() / b.norm().clamp_min(1e-30))


def max_rel(a, b):
    a = a.detach().double().cpu().reshape(-1)
    b = b.detach().double().cpu().reshape(-1)
    return float((a - b).abs().max() / b.abs().max().clamp_min(1e-30))
