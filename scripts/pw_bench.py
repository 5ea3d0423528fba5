"""1x1 RGB-side convolutions (ToRGB forward / data gradient, from_rgb forward) timed alone at the FFHQ-1024 shapes,
batch 16, bf16, CUDA events, L2 flushed between launches.  B200GAN_PW_MODE=1 disables the mma.sync kernels, 2 forces the generic shared-memory-weight kernels.
    python scripts/pw_bench.py"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import __graft_entry__  # noqa: E402

__graft_entry__.build()
from gan_control_b200 import kernels as K  # noqa: E402

dev = torch.device('cuda')
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
B = 16


def timed(fn, reps=5):
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


print(f"B200GAN_PW_MODE = {os.environ.get('B200GAN_PW_MODE', '0')}")
print('| layer | ms | GB moved | TB/s |')
print('|---|---|---|---|')
for res, c in [(1024, 32), (512, 64), (256, 128), (128, 256), (64, 512)]:
    x = torch.randn(B, res, res, c, device=dev).bfloat16()
    w_rgb = torch.randn(B, 1, 1, 3, c, device=dev).bfloat16()
    bias3 = torch.randn(3, device=dev)
    ms = timed(lambda: K.conv_fwd(x, w_rgb, res, res, 1, 1, 0, bias3, None, None, None, 1.0, 1.0))
    gb = B * res * res * (c + 3) * 2 / 1e9
    print(f'| ToRGB fwd {c}->3 @{res}^2 | {ms:.3f} | {gb:.2f} | {gb / ms:.2f} |')
    g3 = torch.randn(B, res, res, 3, device=dev).bfloat16()
    w_t = torch.randn(B, 1, 1, c, 3, device=dev).bfloat16()
    ms = timed(lambda: K.conv_fwd(g3, w_t, res, res, 1, 1, 0))
    print(f'| ToRGB dgrad 3->{c} @{res}^2 | {ms:.3f} | {gb:.2f} | {gb / ms:.2f} |')
    ms = timed(lambda: K.conv_wgrad(x, g3, 1, 1, 1, 1, 0, True))
    print(f'| ToRGB wgrad {c}->3 @{res}^2 (per-sample) | {ms:.3f} | {gb:.2f} | {gb / ms:.2f} |')
    del x, g3
img = torch.randn(B, 1024, 1024, 3, device=dev).bfloat16()
w_in = torch.randn(1, 1, 1, 32, 3, device=dev).bfloat16()
b32 = torch.randn(32, device=dev)
ms = timed(lambda: K.conv_fwd(img, w_in, 1024, 1024, 1, 1, 0, b32, None, None, None, 0.2, 2 ** 0.5))
gb = B * 1024 * 1024 * 35 * 2 / 1e9
print(f'| from_rgb fwd 3->32 @1024^2 + bias + lrelu | {ms:.3f} | {gb:.2f} | {gb / ms:.2f} |')
y0 = torch.randn(B, 1024, 1024, 32, device=dev).bfloat16()
ms = timed(lambda: K.conv_wgrad(img, y0, 1, 1, 1, 1, 0, False))
print(f'| from_rgb wgrad 3->32 @1024^2 (shared) | {ms:.3f} | {gb:.2f} | {gb / ms:.2f} |')
ms = timed(lambda: K.conv_fwd(img, w_in, 1024, 1024, 1, 1, 0, None, None, None, None, 0.2, 2 ** 0.5, gate=y0))
gb = B * 1024 * 1024 * 67 * 2 / 1e9
print(f'| 3->32 @1024^2 + gate | {ms:.3f} | {gb:.2f} | {gb / ms:.2f} |')
