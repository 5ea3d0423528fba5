"""Launch the step's heaviest kernels once each at the FFHQ-1024 batch-16 shapes (for `ncu --set full`):
halo conv 32->32 @1024^2 with the StyledConv epilogue, conv 64->64 @512^2 (bench.py's roofline kernel), the blur,
the fused epilogue backward."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import __graft_entry__  # noqa: E402

__graft_entry__.build()
from gan_control_b200 import kernels as K  # noqa: E402

B, bf, dev = 16, torch.bfloat16, 'cuda'
for h, c in ((1024, 32), (512, 64)):
    x = torch.randn(B, h, h, c, device=dev).to(bf)
    w = (torch.randn(B, 3, 3, c, c, device=dev) / (3 * c ** 0.5)).to(bf)
    d = torch.rand(B, c, device=dev) + 0.5
    noise = torch.randn(B, h, h, device=dev).to(bf)
    nw = torch.full((1,), 0.1, device=dev)
    bias = torch.randn(c, device=dev) * 0.1
    for _ in range(2):
        y = K.conv_fwd(x, w, h, h, 1, 1, 1, bias, d, noise, nw, 0.2, 2 ** 0.5)
        y2 = K.conv_fwd(x, w[:1].contiguous(), h, h, 1, 1, 1)
    gy = torch.randn_like(y)
    for _ in range(2):
        K.epilogue_bwd(gy, y, d, noise, nw, bias, 0.2, 2 ** 0.5)
    torch.cuda.synchronize()
    del x, w, y, y2, gy, noise
taps = (torch.outer(torch.tensor([1., 3, 3, 1]), torch.tensor([1., 3, 3, 1])) / 64 * 4).to(dev)
x = torch.randn(B, 1025, 1025, 32, device=dev).to(bf)
for _ in range(2):
    y = K.upfirdn2d(x, taps, 1, 1, 1, 1, 1024, 1024, True)
torch.cuda.synchronize()
