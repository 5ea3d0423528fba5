"""Protocol stress of the multi-issuer halo weight-gradient kernel and the two-sub-tile general forward engine: many
shapes / batches / weight modes against the CUDA-core engine (a hang shows as the caller's timeout)."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import __graft_entry__  # noqa: E402

__graft_entry__.build()
from gan_control_b200 import kernels as K  # noqa: E402

dev = 'cuda'
torch.manual_seed(0)
n = 0
worst = 0.0
for b in (1, 2, 3, 5, 8, 16, 33):
    for h, w in ((16, 8), (32, 32), (48, 40), (64, 24), (130, 70), (256, 256)):
        if b * h * w > 33 * 130 * 70 and not (b <= 8):
            continue
        for ch in (32, 64):
            for k in (1, 3):
                for ps in (False, True):
                    x = torch.randn(b, h, w, ch, device=dev).bfloat16()
                    gy = torch.randn(b, h, w, ch, device=dev).bfloat16()
                    gw = K.conv_wgrad(x, gy, k, k, 1, 1, k // 2, ps)
                    eng = K.last_conv_engine()
                    prev = K.set_conv_engine(1)
                    ref = K.conv_wgrad(x, gy, k, k, 1, 1, k // 2, ps)
                    K.set_conv_engine(prev)
                    err = float((gw - ref).abs().max() / ref.abs().max())
                    worst = max(worst, err)
                    assert err < 2e-3, (b, h, w, ch, k, ps, err, eng)
                    n += 1
print(f'wgrad: {n} cases, worst rel diff vs the CUDA-core engine {worst:.1e}', flush=True)
n, worst = 0, 0.0
for b in (1, 3, 8, 16):
    for h in (16, 33, 64, 128):
        for ic, oc in ((128, 64), (128, 128), (256, 128), (256, 32), (512, 64), (384, 128)):
            for up, down, pad0 in ((1, 1, 1), (2, 1, 2), (1, 2, 0)):
                if down == 2 and h % 2 == 0:
                    hh = h + 1
                else:
                    hh = h
                oh = (hh - 1) * 2 + 3 - 2 * (3 - 1 - pad0) if up == 2 else ((hh - 1) + 1 + 2 * pad0 - 3) // down + 1
                x = torch.randn(b, hh, hh, ic, device=dev).bfloat16()
                wt = (torch.randn(1, 3, 3, oc, ic, device=dev) / (3 * ic ** 0.5)).bfloat16()
                y = K.conv_fwd(x, wt, oh, oh, up, down, pad0)
                eng = K.last_conv_engine()
                os.environ['B200GAN_UMMA_SUB'] = '1'
                y1 = K.conv_fwd(x, wt, oh, oh, up, down, pad0)
                os.environ.pop('B200GAN_UMMA_SUB')
                assert torch.equal(y, y1), (b, hh, ic, oc, up, down, eng)
                n += 1
torch.cuda.synchronize()
print(f'general forward engine: {n} cases, two sub-tiles per stage == one, bit for bit')
