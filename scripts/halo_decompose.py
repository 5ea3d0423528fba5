"""(Needs a library built with -DB200GAN_HALO_DEBUG=1: the switches are compiled out of the product build.)
Which stage bounds conv_fwd_halo?  Times the kernel with stages switched off through B200GAN_HALO_DEBUG
(1 no MMA, 2 no TMA loads, 4 no global stores, 8 no tcgen05.ld): results are garbage, only the time matters."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import __graft_entry__  # noqa: E402

__graft_entry__.build()
from gan_control_b200 import kernels as K  # noqa: E402

B, bf, dev = 16, torch.bfloat16, 'cuda'
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
shapes = {'c32 32->32@1024 ps': (1024, 32, 32, True, False, False), 'c64 64->64@512 ps': (512, 64, 64, True, False, False)}


def timeit(fn, reps=5):
    fn()
    torch.cuda.synchronize()
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


for name, (h, ic, oc, ps, pin, pout) in shapes.items():
    x = torch.randn(B, 2 * h, 2 * h, ic // 4, device=dev).to(bf) if pin else torch.randn(B, h, h, ic, device=dev).to(bf)
    w = (torch.randn(B if ps else 1, 3, 3, oc, ic, device=dev) / (ic * 9) ** 0.5).to(bf)
    row = []
    for mask in (0, 1, 2, 4, 12, 3, 5, 6, 7, 14, 15):
        os.environ['B200GAN_HALO_DEBUG'] = str(mask)
        row.append((mask, timeit(lambda: K.conv_fwd(x, w, h, h, 1, 1, 1, pack_in=pin, pack_out=pout))))
    os.environ['B200GAN_HALO_DEBUG'] = '0'
    print(f'{name:28s} ' + '  '.join(f'[{m}] {t:.3f}' for m, t in row), flush=True)
