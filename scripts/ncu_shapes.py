"""Launch one convolution shape twice (warm + profiled) for `ncu --set full --launch-skip 1 -c 1`.
    python scripts/ncu_shapes.py {c32|c64|up64|down32} [wgrad]"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import __graft_entry__  # noqa: E402

__graft_entry__.build()
from gan_control_b200 import kernels as K  # noqa: E402

which = sys.argv[1]
wgrad = len(sys.argv) > 2 and sys.argv[2] == 'wgrad'
B, bf, dev = 16, torch.bfloat16, 'cuda'
# name: (h, ic, oc, k, up, down, pad0, out)
shapes = {'c32': (1024, 32, 32, 3, 1, 1, 1, 1024), 'c64': (512, 64, 64, 3, 1, 1, 1, 512),
          'up64': (512, 64, 32, 3, 2, 1, 2, 1025), 'down32': (1025, 32, 64, 3, 1, 2, 0, 512)}
h, ic, oc, k, up, down, pad0, oh = shapes[which]
x = torch.randn(B, h, h, ic, device=dev).to(bf)
w = (torch.randn(1, k, k, oc, ic, device=dev) / (k * ic ** 0.5)).to(bf)
gy = torch.randn(B, oh, oh, oc, device=dev).to(bf)
for _ in range(2):
    if wgrad:
        K.conv_wgrad(x, gy, k, k, up, down, pad0, False)
    else:
        K.conv_fwd(x, w, oh, oh, up, down, pad0)
torch.cuda.synchronize()
