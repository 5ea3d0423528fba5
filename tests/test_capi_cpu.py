"""The C-ABI library loads on a CPU-only box and exports every symbol include/b200gan.h declares;
the product refuses to compute without CUDA (no fallback)."""
import ctypes
import os
import re

import pytest
import torch

from gan_control_b200 import kernels

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
    src = open(os.path.join(ROOT, 'include', 'b200gan.h')).read()
    src = re.sub(r'/\*.*?\*/', '', src, flags=re.S)
    return sorted(set(re.findall(r'\b(b200gan_\w+)\s*\(', src)))


def test_library_exports_every_declared_symbol():
    import __graft_entry__
    __graft_entry__.build()
    handle = ctypes.CDLL(kernels.LIB_PATH)
    names = declared_symbols()
    assert len(names) >= 13
    for n in names:
        assert hasattr(handle, n), f'{n} declared in include/b200gan.h but not exported'
    assert sorted(kernels.exported_symbols()) == names
    assert kernels.lib().b200gan_version() >= 100


def test_no_cpu_fallback():
    from gan_control_b200 import ops
    x = torch.randn(1, 2, 4, 4)
    k = torch.ones(2, 2)
    with pytest.raises(RuntimeError, match='CUDA'):
        ops.upfirdn2d(x, k)
    with pytest.raises(RuntimeError, match='CUDA'):
        ops.fused_leaky_relu(x, torch.zeros(2))


def test_ctypes_signatures_match_header():
    """every binding in kernels._SIGNATURES has as many arguments as the prototype in include/b200gan.h"""
    src = open(os.path.join(ROOT, 'include', 'b200gan.h')).read()
    src = re.sub(r'/\*.*?\*/', '', src, flags=re.S)
    protos = {m.group(1): m.group(2) for m in re.finditer(r'\b(b200gan_\w+)\s*\(([^)]*)\)\s*;', src)}
    for name, (argtypes, _) in kernels._SIGNATURES.items():
        args = protos[name].strip()
        n = 0 if args in ('', 'void') else len(args.split(','))
        assert n == len(argtypes), f'{name}: header has {n} parameters, binding {len(argtypes)}'
