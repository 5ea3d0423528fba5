"""General tcgen05 engine: one vs two 64-channel sub-tiles per pipeline stage (B200GAN_UMMA_SUB=1 forces one), timed alone
(CUDA events, L2 flushed), batch 16, bf16, shared weights."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import __graft_entry__  # noqa: E402

__graft_entry__.build()
from gan_control_b200 import kernels as K  # noqa: E402

dev = 'cuda'
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)


def timed(fn, reps=8):
    for _ in range(3):
        fn()
    ts = []
    for _ in range(reps):
        flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        fn()
        e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    return sorted(ts)[len(ts) // 2]


cases = [  # name, h, ic, oc, up, down, pad0, out_h
    ('conv3x3 128->128 @256^2', 256, 128, 128, 1, 1, 1, 256),
    ('transposed 128->64 256^2 -> 513^2 (up512)', 256, 128, 64, 2, 1, 2, 513),
    ('transposed 256->128 128^2 -> 257^2 (up256)', 128, 256, 128, 2, 1, 2, 257),
    ('stride-2 64->128 513^2 -> 256^2 (res512.conv2)', 513, 64, 128, 1, 2, 0, 256),
    ('stride-2 128->256 257^2 -> 128^2 (res256.conv2)', 257, 128, 256, 1, 2, 0, 128),
    ('conv3x3 256->256 @128^2', 128, 256, 256, 1, 1, 1, 128),
]
for name, h, ic, oc, up, down, pad0, oh in cases:
    x = torch.randn(16, h, h, ic, device=dev).bfloat16()
    w = (torch.randn(1, 3, 3, oc, ic, device=dev) / (3 * ic ** 0.5)).bfloat16()
    res = {}
    for sub in ('1', '2'):
        if sub == '1':
            os.environ['B200GAN_UMMA_SUB'] = '1'
        else:
            os.environ.pop('B200GAN_UMMA_SUB', None)
        res[sub] = timed(lambda: K.conv_fwd(x, w, oh, oh, up, down, pad0))
        res['y' + sub] = K.conv_fwd(x, w, oh, oh, up, down, pad0)
    same = torch.equal(res['y1'], res['y2'])
    print(f'{name}: one sub-tile {res["1"]:.3f} ms, two {res["2"]:.3f} ms, identical={same}, engine {K.last_conv_engine()}')
    del x, w
