"""Launch the RGB-side 1x1 kernels once at the 1024^2 / 512^2 shapes (for `ncu --set full -k regex:pw_small`)."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import __graft_entry__  # noqa: E402

__graft_entry__.build()
from gan_control_b200 import kernels as K  # noqa: E402

B, bf, dev = 16, torch.bfloat16, 'cuda'
x = torch.randn(B, 1024, 1024, 32, device=dev).to(bf)
w_rgb = torch.randn(B, 1, 1, 3, 32, device=dev).to(bf)
bias3 = torch.randn(3, device=dev)
img = torch.randn(B, 1024, 1024, 3, device=dev).to(bf)
w_in = torch.randn(1, 1, 1, 32, 3, device=dev).to(bf)
b32 = torch.randn(32, device=dev)
for _ in range(2):
    K.conv_fwd(x, w_rgb, 1024, 1024, 1, 1, 0, bias3, None, None, None, 1.0, 1.0)               # ToRGB forward
    K.conv_fwd(img, w_in, 1024, 1024, 1, 1, 0, b32, None, None, None, 0.2, 2 ** 0.5)           # from_rgb forward
torch.cuda.synchronize()
