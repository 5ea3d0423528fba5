"""Launch one convolution shape twice (warm + profiled) for `ncu --set full --launch-skip 1 -c 1`.
    python scripts/ncu_shapes.py {c32|c64|up64|down32|upfused|downfused} [wgrad]"""
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
          'torgb': (1024, 32, 3, 1, 1, 1, 0, 1024), 'fromrgb': (1024, 3, 32, 1, 1, 1, 0, 1024),
          'c128': (256, 128, 128, 3, 1, 1, 1, 256), 'c256': (128, 256, 256, 3, 1, 1, 1, 128),
          'up64': (512, 64, 32, 3, 2, 1, 2, 1025), 'down32': (1025, 32, 64, 3, 1, 2, 0, 512)}
h, ic, oc, k, up, down, pad0, oh = shapes.get(which, shapes['c32'])
x = torch.randn(B, h, h, ic, device=dev).to(bf)
w = (torch.randn(1, k, k, oc, ic, device=dev) / (k * ic ** 0.5)).to(bf)
gy = torch.randn(B, oh, oh, oc, device=dev).to(bf)
packed = {'upfused': (512, 64, 128, True, False, True), 'downfused': (512, 128, 64, False, True, False)}
if which in packed:
    h, ic, oc, ps, pin, pout = packed[which]
    x = torch.randn(B, 2 * h, 2 * h, ic // 4, device=dev).to(bf) if pin else torch.randn(B, h, h, ic, device=dev).to(bf)
    w = (torch.randn(B if ps else 1, 3, 3, oc, ic, device=dev) / (ic * 9) ** 0.5).to(bf)
    gy = torch.randn(B, 2 * h, 2 * h, oc // 4, device=dev).to(bf) if pout else torch.randn(B, h, h, oc, device=dev).to(bf)
    for _ in range(2):
        if wgrad:
            K.conv_wgrad(x, gy, 3, 3, 1, 1, 1, ps, pack_x=pin, pack_gy=pout)
        else:
            K.conv_fwd(x, w, h, h, 1, 1, 1, pack_in=pin, pack_out=pout)
    torch.cuda.synchronize()
    sys.exit(0)
ep = len(sys.argv) > 2 and sys.argv[2] == 'ep'        # as a StyledConv: per-sample weights + demod / noise / bias / lrelu epilogue
if ep:
    w = (torch.randn(B, k, k, oc, ic, device=dev) / (k * ic ** 0.5)).to(bf)
    d, bias = torch.rand(B, oc, device=dev) + 0.5, torch.randn(oc, device=dev)
    noise, nw = torch.randn(B, oh, oh, device=dev).to(bf), torch.full((1,), 0.1, device=dev)
for _ in range(2):
    if wgrad:
        K.conv_wgrad(x, gy, k, k, up, down, pad0, False)
    elif ep:
        K.conv_fwd(x, w, oh, oh, up, down, pad0, bias, d, noise, nw, 0.2, 2 ** 0.5)
    else:
        K.conv_fwd(x, w, oh, oh, up, down, pad0)
torch.cuda.synchronize()
