"""Launch a handful of representative convolution kernels once each (for `ncu --set full`)."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import __graft_entry__  # noqa: E402

__graft_entry__.build()
from gan_control_b200 import kernels as K  # noqa: E402

engine = int(sys.argv[1]) if len(sys.argv) > 1 else 0
which = sys.argv[2] if len(sys.argv) > 2 else 'all'
K.set_conv_engine(engine)
B, bf, dev = 16, torch.bfloat16, 'cuda'
shapes = {'c32': (1024, 32, 32), 'c64': (512, 64, 64), 'c256': (128, 256, 256), 'c512': (64, 512, 512)}
for name, (h, ic, oc) in shapes.items():
    if which != 'all' and which != name:
        continue
    x = torch.randn(B, h, h, ic, device=dev).to(bf)
    w = (torch.randn(1, 3, 3, oc, ic, device=dev) / (3 * ic ** 0.5)).to(bf)
    gy = torch.randn(B, h, h, oc, device=dev).to(bf)
    for _ in range(2):
        y = K.conv_fwd(x, w, h, h, 1, 1, 1)
        gw = K.conv_wgrad(x, gy, 3, 3, 1, 1, 1, False)
    torch.cuda.synchronize()
    del x, w, gy, y, gw
