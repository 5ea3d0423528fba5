"""Launch the layer the metric is quoted on exactly as the bench's `roofline` entry times it (conv3x3 32 -> 32 @1024^2,
batch 16, bf16, per-sample demodulated weights, noise + bias + leaky-ReLU epilogue), and its weight gradient, once each
(for `ncu --set full -k regex:halo`)."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import __graft_entry__  # noqa: E402

__graft_entry__.build()
from gan_control_b200 import kernels as K  # noqa: E402

B, bf, dev, res, ch = 16, torch.bfloat16, 'cuda', 1024, 32
x = torch.randn(B, res, res, ch, device=dev).to(bf)
w = (torch.randn(B, 3, 3, ch, ch, device=dev) / (3 * ch ** 0.5)).to(bf)
gy = torch.randn(B, res, res, ch, device=dev).to(bf)
bias = torch.randn(ch, device=dev)
noise = torch.randn(B, res, res, device=dev).to(bf)
nw = torch.full((1,), 0.1, device=dev)
for _ in range(2):
    K.conv_fwd(x, w, res, res, 1, 1, 1, bias, None, noise, nw, 0.2, 2 ** 0.5)
    K.conv_wgrad(x, gy, 3, 3, 1, 1, 1, True)
torch.cuda.synchronize()
