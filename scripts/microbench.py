"""Per-kernel timings at the FFHQ-1024 layer shapes (CUDA events, inputs > L2 or L2 flushed between
launches), with achieved TFLOP/s / GB/s against MEASURED_PEAKS.json.   python scripts/microbench.py [filter]"""
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import __graft_entry__  # noqa: E402

__graft_entry__.build()
from gan_control_b200 import kernels as K  # noqa: E402

flt = sys.argv[1] if len(sys.argv) > 1 else ''
peaks = json.load(open(os.path.join(ROOT, 'MEASURED_PEAKS.json'))) if os.path.exists(os.path.join(ROOT, 'MEASURED_PEAKS.json')) else {'hbm_gbs': 6650, 'bf16_tflops': 1590}
dev = 'cuda'
bf = torch.bfloat16
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)


def timeit(fn, reps=5):
    fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(reps):
        flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        fn()
        e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    ts.sort()
    return ts[len(ts) // 2]


def report(name, ms, flops=0, bytes_=0):
    tf = flops / ms / 1e9 if flops else 0
    gbs = bytes_ / ms / 1e6 if bytes_ else 0
    print(f'{name:58s} {ms:8.3f} ms  {tf:8.1f} TF/s ({100 * tf / peaks["bf16_tflops"]:5.1f}%)  {gbs:8.0f} GB/s ({100 * gbs / peaks["hbm_gbs"]:5.1f}%)', flush=True)


B = 16
convs = [  # name, h, ic, oc, k, up, down, pad0, per_sample
    ('conv3x3 s1 32->32 @1024 ps', 1024, 32, 32, 3, 1, 1, 1, True), ('conv3x3 s1 64->64 @512 ps', 512, 64, 64, 3, 1, 1, 1, True),
    ('conv3x3 s1 64->64 @512 shared', 512, 64, 64, 3, 1, 1, 1, False), ('conv3x3 s1 128->128 @256 ps', 256, 128, 128, 3, 1, 1, 1, True),
    ('conv3x3 s1 256->256 @128 ps', 128, 256, 256, 3, 1, 1, 1, True), ('conv3x3 s1 512->512 @64', 64, 512, 512, 3, 1, 1, 1, False),
    ('conv3x3 s1 512->512 @32', 32, 512, 512, 3, 1, 1, 1, False), ('conv3x3 s1 512->512 @8', 8, 512, 512, 3, 1, 1, 1, False),
    ('convT up2 64->32 @512->1025 ps', 512, 64, 32, 3, 2, 1, 2, True), ('convT up2 512->512 @32->65', 32, 512, 512, 3, 2, 1, 2, False),
    ('conv3x3 s2 32->64 @1025->512', 1025, 32, 64, 3, 1, 2, 0, False), ('conv3x3 s2 256->512 @129->64', 129, 256, 512, 3, 1, 2, 0, False),
    ('conv1x1 s2 32->64 @1023->512', 1023, 32, 64, 1, 1, 2, 0, False), ('torgb 1x1 32->3 @1024 ps', 1024, 32, 3, 1, 1, 1, 0, True),
    ('fromrgb 1x1 3->32 @1024', 1024, 3, 32, 1, 1, 1, 0, False),
]
for name, h, ic, oc, k, up, down, pad0, ps in convs:
    if flt and flt not in name and flt != 'conv':
        continue
    if up == 2:
        oh = (h - 1) * 2 + k - 2 * (k - 1 - pad0)
    else:
        oh = ((h - 1) * up + 1 + 2 * pad0 - k) // down + 1
    x = torch.randn(B, h, h, ic, device=dev).to(bf)
    w = (torch.randn(B if ps else 1, k, k, oc, ic, device=dev) / (ic * k * k) ** 0.5).to(bf)
    fl = 2.0 * B * (h * h if up == 2 else oh * oh) * ic * oc * k * k
    by = 2.0 * B * (h * h * ic + oh * oh * oc)
    report('fwd   ' + name, timeit(lambda: K.conv_fwd(x, w, oh, oh, up, down, pad0)), fl, by)
    gy = torch.randn(B, oh, oh, oc, device=dev).to(bf)
    report('wgrad ' + name, timeit(lambda: K.conv_wgrad(x, gy, k, k, up, down, pad0, ps)), fl, by)
    del x, w, gy

# fused resampling layers: 3x3 convolutions between space-to-depth views (composite FIR (*) conv weights)
packed = [  # name, view h, view ic, view oc, per_sample, pack_in, pack_out
    ('up 64->32 @512->1024 fused (pack_out) ps', 512, 64, 128, True, False, True),
    ('up 64->32 dgrad (pack_in) ps', 512, 128, 64, True, True, False),
    ('down 32->64 @1024->512 fused (pack_in)', 512, 128, 64, False, True, False),
    ('down 32->64 dgrad (pack_out)', 512, 64, 128, False, False, True),
]
for name, h, ic, oc, ps, pin, pout in packed:
    if flt and flt not in name and flt not in ('conv', 'packed'):
        continue
    x = torch.randn(B, 2 * h, 2 * h, ic // 4, device=dev).to(bf) if pin else torch.randn(B, h, h, ic, device=dev).to(bf)
    w = (torch.randn(B if ps else 1, 3, 3, oc, ic, device=dev) / (ic * 9) ** 0.5).to(bf)
    gy = torch.randn(B, 2 * h, 2 * h, oc // 4, device=dev).to(bf) if pout else torch.randn(B, h, h, oc, device=dev).to(bf)
    fl = 2.0 * B * h * h * ic * oc * 9
    by = 2.0 * B * h * h * (ic + oc)
    report('fwd   ' + name, timeit(lambda: K.conv_fwd(x, w, h, h, 1, 1, 1, pack_in=pin, pack_out=pout)), fl, by)
    report('wgrad ' + name, timeit(lambda: K.conv_wgrad(x, gy, 3, 3, 1, 1, 1, ps, pack_x=pin, pack_gy=pout)), fl, by)
    del x, w, gy

if not flt or flt in 'blur':
    taps = (torch.outer(torch.tensor([1., 3, 3, 1]), torch.tensor([1., 3, 3, 1])) / 64 * 4).to(dev)
    for h, c in [(1025, 32), (1024, 32), (513, 64), (257, 128), (65, 512)]:
        x = torch.randn(B, h, h, c, device=dev).to(bf)
        report(f'blur 4x4 pad1 {c}ch @{h}', timeit(lambda: K.upfirdn2d(x, taps, 1, 1, 1, 1, h - 1, h - 1, True)), 0, 2.0 * B * c * (h * h + (h - 1) ** 2))
        del x
    for h, c in [(1024, 32), (512, 64), (256, 128)]:
        # ResBlock.skip: blur decimated by 2 (forward) and its adjoint (zero-insert x2, reversed taps)
        x = torch.randn(B, h, h, c, device=dev).to(bf)
        g = torch.randn(B, h // 2, h // 2, c, device=dev).to(bf)
        t1 = taps / 4
        report(f'blur-down 4x4 pad1 down2 {c}ch @{h}->{h // 2}', timeit(lambda: K.upfirdn2d(x, t1, 1, 2, 1, 1, h // 2, h // 2, True)), 0, 2.0 * B * c * (h * h + (h // 2) ** 2))
        report(f'blur-down adjoint up2 {c}ch @{h // 2}->{h}', timeit(lambda: K.upfirdn2d(g, t1, 2, 1, 2, 2, h, h, False)), 0, 2.0 * B * c * (h * h + (h // 2) ** 2))
        del x, g
    x = torch.randn(B, 512, 512, 3, device=dev).to(bf)
    report('skip upsample x2 3ch @512->1024', timeit(lambda: K.upfirdn2d(x, taps, 2, 1, 2, 2, 1024, 1024, True)), 0, 2.0 * B * 3 * (512 * 512 + 1024 * 1024))

if not flt or flt in 'epilogue':
    for h, c in [(1024, 32), (512, 64), (64, 512)]:
        x = torch.randn(B, h, h, c, device=dev).to(bf)
        y = torch.randn(B, h, h, c, device=dev).to(bf)
        d = torch.rand(B, c, device=dev) + 0.5
        nz = torch.randn(B, h, h, device=dev).to(bf)
        nw = torch.tensor([0.1], device=dev)
        bias = torch.randn(c, device=dev)
        by = 2.0 * B * h * h * c
        report(f'bias_act_fwd {c}ch @{h}', timeit(lambda: K.bias_act_fwd(x, bias, d, nz, nw)), 0, 2 * by)
        report(f'epilogue_bwd {c}ch @{h}', timeit(lambda: K.epilogue_bwd(x, y, d, nz, nw, bias)), 0, 3 * by)
        report(f'reduce_nhwc(a*b) {c}ch @{h}', timeit(lambda: K.reduce_nhwc(x, y, True, True)), 0, 2 * by)
        del x, y
