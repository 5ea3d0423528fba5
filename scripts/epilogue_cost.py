"""What each fused-epilogue operand costs in the small-channel tcgen05 convolution (timed alone, CUDA events, L2 flushed).
    python scripts/epilogue_cost.py > gpurun_out/epilogue_cost.txt"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import __graft_entry__  # noqa: E402

__graft_entry__.build()
from gan_control_b200 import kernels as K  # noqa: E402

flush = torch.empty(256 << 20, dtype=torch.uint8, device='cuda')


def timeit(fn, reps=10):
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
    return sum(ts) / len(ts)


b = 16
for res, ch in ((1024, 32), (512, 64)):
    x = torch.randn(b, res, res, ch, device='cuda').bfloat16()
    w = (torch.randn(b, 3, 3, ch, ch, device='cuda') / (3 * ch ** 0.5)).bfloat16()
    d = torch.rand(b, ch, device='cuda') + 0.5
    bias = torch.randn(ch, device='cuda')
    noise = torch.randn(b, res, res, device='cuda').bfloat16()
    nw = torch.full((1,), 0.1, device='cuda')
    variants = {'bare': (None, None, None, None, 1.0, 1.0), 'lrelu only': (None, None, None, None, 0.2, 1.41),
                'bias': (bias, None, None, None, 0.2, 1.41), 'rowscale': (None, d, None, None, 0.2, 1.41),
                'noise': (None, None, noise, nw, 0.2, 1.41), 'bias+noise': (bias, None, noise, nw, 0.2, 1.41),
                'all': (bias, d, noise, nw, 0.2, 1.41)}
    for name, (bi, rs, nz, nww, slope, gain) in variants.items():
        ms = timeit(lambda: K.conv_fwd(x, w, res, res, 1, 1, 1, bi, rs, nz, nww, slope, gain))
        print(f'conv3x3 {ch}->{ch} @{res} batch {b} per-sample w, epilogue {name:12s}: {ms:.3f} ms  '
              f'{2 * x.numel() * 2 / ms / 1e6:.0f} GB/s  [{K.last_conv_engine()}]', flush=True)
