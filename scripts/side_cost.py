"""What the output-shaped side inputs of the convolution epilogue (addend / gate, include/b200gan.h) cost, against the
separate passes they replace (elementwise add, epilogue_bwd).  Timed alone, CUDA events, L2 flushed.
    python scripts/side_cost.py > gpurun_out/side_cost.txt"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import __graft_entry__  # noqa: E402

__graft_entry__.build()
from gan_control_b200 import kernels as K  # noqa: E402

flush = torch.empty(256 << 20, dtype=torch.uint8, device='cuda')
bf = torch.bfloat16


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
cases = [('conv3x3 32->32 @1024 (halo)', 1024, 32, 32, False, False), ('conv3x3 64->64 @512 (halo)', 512, 64, 64, False, False),
         ('conv3x3 128->128 @256 (umma)', 256, 128, 128, False, False),
         ('fused-down data gradient 64 -> 4x32 @512 view, pack_out (halo)', 512, 64, 128, False, True)]
for name, res, ic, oc, pin, pout in cases:
    x = torch.randn(b, res, res, ic, device='cuda').to(bf)
    w = (torch.randn(1, 3, 3, oc, ic, device='cuda') / (3 * ic ** 0.5)).to(bf)
    yshape = (b, 2 * res, 2 * res, oc // 4) if pout else (b, res, res, oc)
    add = torch.randn(*yshape, device='cuda').to(bf)
    gate = torch.randn(*yshape, device='cuda').to(bf)
    kw = dict(pack_in=pin, pack_out=pout)
    t0 = timeit(lambda: K.conv_fwd(x, w, res, res, 1, 1, 1, **kw))
    t1 = timeit(lambda: K.conv_fwd(x, w, res, res, 1, 1, 1, addend=add, **kw))
    t2 = timeit(lambda: K.conv_fwd(x, w, res, res, 1, 1, 1, slope=0.2, gain=1.41, gate=gate, **kw))
    t3 = timeit(lambda: K.conv_fwd(x, w, res, res, 1, 1, 1, slope=0.2, gain=1.41, addend=add, gate=gate, **kw))
    y = K.conv_fwd(x, w, res, res, 1, 1, 1, **kw)
    ta = timeit(lambda: torch.add(y, add))
    te = timeit(lambda: K.epilogue_bwd(y, gate, None, None, None, None, 0.2, 1.41, want_gd=False, want_gb=True, want_gnw=False))
    tr = timeit(lambda: K.reduce_nhwc(y, None, per_channel=True))
    print(f'{name}: plain {t0:.3f} | +addend {t1:.3f} | +gate {t2:.3f} | +both {t3:.3f} ms   [separate passes: add {ta:.3f}, '
          f'epilogue_bwd {te:.3f}, bias reduction {tr:.3f} ms]  engine {K.last_conv_engine()}', flush=True)
    del x, w, add, gate, y
