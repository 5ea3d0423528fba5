"""Weight-gradient kernels: one MMA-issuing warp vs one per accumulator group (B200GAN_WGRAD_ISSUERS; honoured by the HALO
kernel -- the multi-issuer build of the general kernel was reverted, see profiles/r02_summary.md, so its rows show no
difference), timed alone
(CUDA events, L2 flushed), batch 16, bf16."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import __graft_entry__  # noqa: E402

__graft_entry__.build()
from gan_control_b200 import kernels as K  # noqa: E402

dev = 'cuda'
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)


def timed(fn, reps=8):
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
    return sorted(ts)[len(ts) // 2]


for res, ch, ps in [(1024, 32, True), (1024, 32, False), (512, 64, True), (512, 64, False),
                    (256, 128, True), (256, 128, False), (128, 256, True), (64, 512, False)]:      # last four: general engine
    x = torch.randn(16, res, res, ch, device=dev).bfloat16()
    gy = torch.randn(16, res, res, ch, device=dev).bfloat16()
    out = {}
    for iss in ('1', '3'):
        os.environ['B200GAN_WGRAD_ISSUERS'] = iss
        out[iss] = timed(lambda: K.conv_wgrad(x, gy, 3, 3, 1, 1, 1, ps))
        ref = K.conv_wgrad(x, gy, 3, 3, 1, 1, 1, ps)
        out['g' + iss] = ref
    err = float((out['g3'] - out['g1']).abs().max() / out['g1'].abs().max())
    print(f'conv3x3 {ch}->{ch} @{res}^2 wgrad per_sample={ps}: 1 issuer {out["1"]:.3f} ms, 3 issuers {out["3"]:.3f} ms, '
          f'max-rel diff {err:.1e}, engine {K.last_conv_engine()}')
    del x, gy
os.environ.pop('B200GAN_WGRAD_ISSUERS', None)
