"""Where do the fused first-order path's bf16 gradients of cancellation-heavy sums (ToRGB biases, noise strengths) lose
accuracy against the unfused bf16 path?  Splits the G -> D chain at the image: (1) dL/d(image) of D alone, (2) G's
parameter gradients for a FIXED image gradient.  fp32 unfused is the truth.
    python scripts/diag_fused_bf16.py"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import __graft_entry__  # noqa: E402

__graft_entry__.build()
from gan_control_b200 import modules as M, ops  # noqa: E402

DEV = 'cuda'
size, sdim = 64, 64


def rel(a, b):
    return float((a.double() - b.double()).norm() / b.double().norm().clamp_min(1e-30))


nets = {}
for dt in (torch.float32, torch.bfloat16):
    torch.manual_seed(5)
    g = M.Generator(size, sdim, 3, channel_multiplier=2, conv_transpose=True, act_dtype=dt).to(DEV)
    d = M.Discriminator(size, channel_multiplier=2, act_dtype=dt).to(DEV)
    for m in list(g.modules()) + list(d.modules()):
        if isinstance(m, M.NoiseInjection):
            m.weight.data.fill_(0.3)
        if isinstance(m, (M.FusedLeakyReLU,)):
            m.bias.data.normal_(std=0.3)
        if isinstance(m, M.ToRGB):
            m.bias.data.normal_(std=0.3)
    nets[dt] = (g, d)
torch.manual_seed(7)
z = torch.randn(4, sdim, device=DEV)
noise = [torch.randn(4, 1, 2 ** ((i + 5) // 2), 2 ** ((i + 5) // 2), device=DEV) for i in range(nets[torch.float32][0].num_layers)]
with torch.no_grad():
    img32, _ = nets[torch.float32][0]([z], noise=noise)


def d_image_grad(d, img, fused):
    x = img.detach().clone().requires_grad_(True)
    with (ops.first_order() if fused else torch.enable_grad()):
        pred, _ = d(x)
        torch.nn.functional.softplus(-pred).mean().backward()
    return x.grad.float()


print('--- (1) D alone: dL/d(image) for the same fp32 image')
ref = d_image_grad(nets[torch.float32][1], img32, False)
for dt, fused in [(torch.float32, True), (torch.bfloat16, False), (torch.bfloat16, True)]:
    gi = d_image_grad(nets[dt][1], img32, fused)
    dc, dc_ref = gi.sum((0, 2, 3)), ref.sum((0, 2, 3))
    print(f'{str(dt):16s} fused={fused!s:5s}: rel-L2 {rel(gi, ref):.3e}; per-channel DC (sum over pixels) {dc.tolist()} vs {dc_ref.tolist()}, '
          f'DC rel {rel(dc, dc_ref):.3e}; sum|g| {float(ref.abs().sum()):.3e}')


def g_param_grads(g, gimg, fused):
    g.zero_grad()
    with (ops.first_order() if fused else torch.enable_grad()):
        img, _ = g([z], noise=noise)
        img.backward(gimg.to(img.dtype))
    return {k: v.grad.float().clone() for k, v in g.named_parameters() if v.grad is not None}


print('--- (2) G alone: parameter gradients for the same (fp32 D) image gradient')
gref = g_param_grads(nets[torch.float32][0], ref, False)
res = {}
for dt, fused in [(torch.float32, True), (torch.bfloat16, False), (torch.bfloat16, True)]:
    res[(dt, fused)] = g_param_grads(nets[dt][0], ref, fused)
for k in gref:
    if float(gref[k].abs().max()) == 0:
        continue
    e = [rel(res[key][k], gref[k]) for key in res]
    flag = '  <--' if e[2] > max(2.5 * e[1], 5e-2) else ''
    if gref[k].numel() <= 4 or flag:
        print(f'{k:28s} numel {gref[k].numel():7d}: fp32 fused {e[0]:.2e}, bf16 unfused {e[1]:.2e}, bf16 fused {e[2]:.2e}{flag}')
