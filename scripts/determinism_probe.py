"""Run the fp32 smoke G + D forward / backward several times on identical inputs and report which parameter gradients
differ between runs (atomic-order noise is ~1e-7; anything above ~1e-5 points at a race or an uninitialised read)."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import __graft_entry__ as G  # noqa: E402

G.build()
from gan_control_b200 import modules as M  # noqa: E402
from oracle import params as P  # noqa: E402

dev = torch.device('cuda:0')
size, sdim, n_mlp, cm, batch, seed = 16, 64, 3, 2, 4, 11
sd = P.seeded_state_dict(P.generator_shapes(size, sdim, n_mlp, cm), seed)
g = M.Generator(size, sdim, n_mlp, channel_multiplier=cm, conv_transpose=True).to(dev)
g.load_state_dict(sd)
d = M.Discriminator(size, channel_multiplier=cm).to(dev)
d.load_state_dict(P.seeded_state_dict(P.discriminator_shapes(size, cm), seed + 10))
torch.manual_seed(seed)
z = torch.randn(batch, sdim).to(dev)
noise = [torch.randn(batch, 1, 2 ** ((i + 5) // 2), 2 ** ((i + 5) // 2)).to(dev) for i in range(g.num_layers)]
runs = []
for i in range(5):
    g.zero_grad()
    d.zero_grad()
    img, _ = g([z], noise=noise)
    img.retain_grad()
    pred, _ = d(img)
    torch.nn.functional.softplus(-pred).mean().backward()
    torch.cuda.synchronize()
    grads = {'g.' + k: v.grad.clone() for k, v in g.named_parameters() if v.grad is not None}
    grads.update({'d.' + k: v.grad.clone() for k, v in d.named_parameters() if v.grad is not None})
    grads['img'] = img.detach().clone()
    grads['pred'] = pred.detach().clone()
    grads['dL/dimg'] = img.grad.clone()
    runs.append(grads)
worst = []
for k in runs[0]:
    ref = runs[0][k].double()
    dev_ = max(float((r[k].double() - ref).norm() / ref.norm().clamp_min(1e-30)) for r in runs[1:])
    worst.append((dev_, k))
worst.sort(reverse=True)
for dv, k in worst[:25]:
    print(f'{dv:.3e}  {k}')
print('...')
for dv, k in worst:
    if k in ('img', 'pred', 'dL/dimg'):
        print(f'{dv:.3e}  {k}')

print('--- D alone on ONE fixed image tensor')
fixed = runs[0]['img'].clone()
dr = []
for i in range(5):
    d.zero_grad()
    x = fixed.clone().requires_grad_(True)
    pred, _ = d(x)
    torch.nn.functional.softplus(-pred).mean().backward()
    torch.cuda.synchronize()
    gr = {'d.' + k: v.grad.clone() for k, v in d.named_parameters() if v.grad is not None}
    gr['pred'] = pred.detach().clone()
    gr['dL/dimg'] = x.grad.clone()
    dr.append(gr)
w2 = []
for k in dr[0]:
    ref = dr[0][k].double()
    w2.append((max(float((r[k].double() - ref).norm() / ref.norm().clamp_min(1e-30)) for r in dr[1:]), k,
               all(torch.equal(r[k], dr[0][k]) for r in dr[1:])))
w2.sort(reverse=True)
for dv, k, same in w2[:12]:
    print(f'{dv:.3e}  {k}  bitwise-equal={same}')
for dv, k, same in w2:
    if k in ('pred', 'dL/dimg'):
        print(f'{dv:.3e}  {k}  bitwise-equal={same}')

print('--- D layer by layer: forward activations of two passes over the fixed image')
acts = []
for i in range(2):
    cur, outs = fixed.clone(), []
    with torch.no_grad():
        x = cur.to(memory_format=torch.channels_last)
        for blk in d.convs:
            x = blk(x)
            outs.append(x.clone())
    acts.append(outs)
for j, (a, b) in enumerate(zip(*acts)):
    print(f'block {j}: max abs diff {float((a - b).abs().max()):.3e}, bitwise-equal={torch.equal(a, b)}')
