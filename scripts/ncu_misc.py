"""Launch the bandwidth-bound kernels once at the 1024^2 shapes (for `ncu --set full`)."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import __graft_entry__  # noqa: E402

__graft_entry__.build()
from gan_control_b200 import kernels as K  # noqa: E402

B, bf, dev = 16, torch.bfloat16, 'cuda'
taps = (torch.outer(torch.tensor([1., 3, 3, 1]), torch.tensor([1., 3, 3, 1])) / 64 * 4).to(dev)
x = torch.randn(B, 1025, 1025, 32, device=dev).to(bf)
for _ in range(2):
    y = K.upfirdn2d(x, taps, 1, 1, 1, 1, 1024, 1024, True)
torch.cuda.synchronize()
