"""Generate `tests/golden/*.npz` by running the UNMODIFIED reference on CPU.

Run in the build container only (needs /root/reference):

    python oracle/make_golden.py

Every fixture stores seeded inputs, the parameters (or the seed they are
rebuilt from) and the outputs / gradients the reference `FUSED = False` path
produced (fp32 and fp64).  Nothing here is product code.
"""
import math
import os
import random
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))

from oracle import params as P            # noqa: E402
from oracle.ref_import import import_reference   # noqa: E402

OUT = os.path.join(os.path.dirname(HERE), 'tests', 'golden')


def npy(t):
    return t.detach().cpu().numpy()


def rnd(seed, *shape, dtype=torch.float64):
    rng = np.random.default_rng(seed)
    return torch.from_numpy(rng.standard_normal(shape)).to(dtype)


def subsample(a, limit=8192):
    """Big tensors are stored as (strided subsample, L2 norm); `tests/golden_io.py`
    applies the same reduction to the tensor under test."""
    a = np.asarray(a)
    if a.size <= limit:
        return a
    flat = a.reshape(-1)
    return np.concatenate([flat[:: flat.size // 4096][:4096], [np.sqrt((flat.astype(np.float64) ** 2).sum())]])


def save(name, d):
    os.makedirs(OUT, exist_ok=True)
    path = os.path.join(OUT, name + '.npz')
    arrays = {}
    for k, v in d.items():
        if v is None:
            continue
        a = npy(v) if torch.is_tensor(v) else np.asarray(v)
        is_grad = '.g.' in k
        arrays[k] = subsample(a) if is_grad else a
    np.savez_compressed(path, **arrays)
    print(f'{name}: {len(d)} arrays, {os.path.getsize(path) / 1024:.1f} KiB')


def gen_upfirdn2d(gm):
    d = {}
    k4 = gm.make_kernel([1, 3, 3, 1]).double()
    rng = np.random.default_rng(7)
    k12 = torch.from_numpy(np.outer(rng.standard_normal(12), rng.standard_normal(12)))
    cases = {   # name: (kernel, up, down, pad, in_h, in_w)   -- SURVEY.md Appendix A.3 call sites
        'g_upblur': (k4 * 4, 1, 1, (1, 1), 9, 9),
        'rgb_skip': (k4 * 4, 2, 1, (2, 1), 6, 6),
        'd_conv2_blur': (k4, 1, 1, (2, 2), 8, 8),
        'd_skip_blur': (k4, 1, 1, (1, 1), 8, 8),
        'ada_up': (k12, 2, 1, (0, 0), 10, 10),
        'ada_down': (k12, 1, 2, (0, 0), 21, 21),
        'down_module': (k4, 1, 2, (1, 1), 8, 8),
        'negpad': (k4, 1, 1, (-1, 2), 9, 7),
        'rect': (k4 * 4, 2, 1, (2, 1), 5, 7),
    }
    for i, (name, (k, up, down, pad, h, w)) in enumerate(cases.items()):
        x = rnd(100 + i, 2, 3, h, w).requires_grad_(True)
        y = gm.upfirdn2d(x, k, up=up, down=down, pad=pad)
        gy = rnd(200 + i, *y.shape)
        gx, = torch.autograd.grad(y, x, gy)
        d.update({f'{name}.x': x, f'{name}.k': k, f'{name}.cfg': np.array([up, down, pad[0], pad[1]]),
                  f'{name}.y': y, f'{name}.gy': gy, f'{name}.gx': gx})
    save('upfirdn2d', d)


def gen_bias_act(gm):
    d = {}
    for i, shape in enumerate([(2, 5, 4, 4), (3, 7)]):
        x = rnd(300 + i, *shape).requires_grad_(True)
        b = rnd(310 + i, shape[1]).requires_grad_(True)
        y = gm.fused_leaky_relu(x, b)
        gy = rnd(320 + i, *shape)
        gx, gb = torch.autograd.grad(y, (x, b), gy)
        d.update({f'c{i}.x': x, f'c{i}.b': b, f'c{i}.y': y, f'c{i}.gy': gy, f'c{i}.gx': gx, f'c{i}.gb': gb})
    m = gm.FusedLeakyReLU(5).double()
    with torch.no_grad():
        m.bias.copy_(rnd(330, 5))
    x = rnd(331, 2, 5, 3, 3)
    d.update({'mod.x': x, 'mod.b': m.bias, 'mod.y': m(x)})
    save('bias_act', d)


def gen_linear(gm):
    d = {}
    for i, (din, dout, lr_mul, act, bias_init) in enumerate(
            [(16, 12, 0.01, 'fused_lrelu', 0), (16, 8, 1, None, 1), (3, 24, 0.01, 'fused_lrelu', 0)]):
        torch.manual_seed(400 + i)
        m = gm.EqualLinear(din, dout, bias_init=bias_init, lr_mul=lr_mul, activation=act).double()
        with torch.no_grad():
            m.bias.add_(rnd(410 + i, dout) * (0.1 / lr_mul))
        x = rnd(420 + i, 5, din).requires_grad_(True)
        y = m(x)
        gy = rnd(430 + i, *y.shape)
        gx, gw, gb = torch.autograd.grad(y, (x, m.weight, m.bias), gy)
        d.update({f'c{i}.x': x, f'c{i}.w': m.weight, f'c{i}.b': m.bias, f'c{i}.cfg': np.array([lr_mul, 1.0 if act else 0.0]),
                  f'c{i}.y': y, f'c{i}.gy': gy, f'c{i}.gx': gx, f'c{i}.gw': gw, f'c{i}.gb': gb})
    save('equal_linear', d)


def gen_modconv(gm):
    """ModulatedConv2d / StyledConv / ToRGB incl. first- and second-order grads."""
    d = {}
    cases = {   # name: (ic, oc, k, demod, up, h)
        'plain3': (8, 12, 3, True, False, 6),
        'up3': (8, 6, 3, True, True, 5),
        'rgb1': (8, 3, 1, False, False, 6),
        'plain3_b1': (4, 4, 3, True, False, 4),
    }
    sdim = 16
    for i, (name, (ic, oc, k, demod, up, h)) in enumerate(cases.items()):
        torch.manual_seed(500 + i)
        m = gm.ModulatedConv2d(ic, oc, k, sdim, demodulate=demod, upsample=up, conv_transpose=True).double()
        with torch.no_grad():
            m.modulation.bias.add_(rnd(510 + i, ic) * 0.1)
        b = 1 if name.endswith('b1') else 3
        x = rnd(520 + i, b, ic, h, h).requires_grad_(True)
        s = rnd(530 + i, b, sdim).requires_grad_(True)
        y = m(x, s)
        gy = rnd(540 + i, *y.shape)
        ps = (x, s, m.weight, m.modulation.weight, m.modulation.bias)
        g = torch.autograd.grad(y, ps, gy, create_graph=True)
        # path-length style second order: d/dθ || d<y,gy>/ds ||²
        pl = g[1].pow(2).sum()
        gg = torch.autograd.grad(pl, (x, s, m.weight, m.modulation.weight), allow_unused=True)
        d.update({f'{name}.cfg': np.array([ic, oc, k, int(demod), int(up), h, b, sdim]),
                  f'{name}.x': x, f'{name}.s': s, f'{name}.w': m.weight, f'{name}.mw': m.modulation.weight,
                  f'{name}.mb': m.modulation.bias, f'{name}.y': y, f'{name}.gy': gy,
                  f'{name}.gx': g[0], f'{name}.gs': g[1], f'{name}.gw': g[2], f'{name}.gmw': g[3], f'{name}.gmb': g[4],
                  f'{name}.pl': pl, f'{name}.pl_gx': gg[0], f'{name}.pl_gs': gg[1], f'{name}.pl_gw': gg[2], f'{name}.pl_gmw': gg[3]})
        if up:
            d[f'{name}.blur'] = m.blur.kernel
    # StyledConv (noise + bias + act) and ToRGB (bias + upsampled skip)
    for i, up in enumerate([False, True]):
        torch.manual_seed(600 + i)
        m = gm.StyledConv(8, 6, 3, sdim, upsample=up, conv_transpose=True).double()
        with torch.no_grad():
            m.noise.weight.fill_(0.3)
            m.activate.bias.copy_(rnd(610 + i, 6) * 0.2)
        x = rnd(620 + i, 2, 8, 5, 5)
        s = rnd(630 + i, 2, sdim)
        ho = 10 if up else 5
        nz = rnd(640 + i, 2, 1, ho, ho)
        name = f'styled_up{int(up)}'
        sd = {k: v for k, v in m.state_dict().items()}
        d.update({f'{name}.sd.{k}': v for k, v in sd.items()})
        d.update({f'{name}.x': x, f'{name}.s': s, f'{name}.noise': nz, f'{name}.y': m(x, s, noise=nz)})
    torch.manual_seed(650)
    m = gm.ToRGB(8, sdim, conv_transpose=True).double()
    with torch.no_grad():
        m.bias.copy_(rnd(651, 1, 3, 1, 1) * 0.2)
    x, s, skip = rnd(652, 2, 8, 8, 8), rnd(653, 2, sdim), rnd(654, 2, 3, 4, 4)
    d.update({f'torgb.sd.{k}': v for k, v in m.state_dict().items()})
    d.update({'torgb.x': x, 'torgb.s': s, 'torgb.skip': skip, 'torgb.y': m(x, s, skip), 'torgb.y_noskip': m(x, s)})
    save('modconv', d)


class FcCfg:
    def __init__(self, groups):
        self.in_order_group_names = [g[0] for g in groups]
        self.groups = {n: {'latent_place': [lo, hi], 'latent_size': hi - lo} for n, lo, hi in groups}


def load_seeded(module, shapes, seed, dtype):
    sd_ref = module.state_dict()
    assert set(sd_ref) == set(shapes), (sorted(set(sd_ref) ^ set(shapes)))
    for k, v in sd_ref.items():
        assert tuple(v.shape) == tuple(shapes[k]), (k, v.shape, shapes[k])
    sd = P.seeded_state_dict(shapes, seed, dtype=dtype)
    for k in sd:
        if k.endswith('kernel'):
            assert torch.allclose(sd[k].float(), sd_ref[k].float(), atol=1e-7), k
    module.load_state_dict(sd)
    return sd


def gen_networks(gm, trainer):
    d = {}
    # ---- generators: vanilla + split-FC, size 16, fp64 and fp32 ------------
    groups = [('id', 0, 24), ('pose', 24, 40), ('other', 40, 64)]
    for name, size, sdim, n_mlp, fcg, seed in [('g16', 16, 64, 3, None, 11), ('g16split', 16, 64, 4, groups, 12)]:
        for dt, tag in [(torch.float64, 'f64'), (torch.float32, 'f32')]:
            g = gm.Generator(size, sdim, n_mlp, channel_multiplier=2, conv_transpose=True,
                             split_fc=fcg is not None, fc_config=FcCfg(fcg) if fcg else None).to(dt)
            shapes = P.generator_shapes(size, sdim, n_mlp, 2, fcg)
            load_seeded(g, shapes, seed, dt)
            z = rnd(seed * 10, 2, sdim, dtype=dt)
            z2 = rnd(seed * 10 + 1, 2, sdim, dtype=dt)
            noise = [rnd(seed * 100 + i, 2, 1, 2 ** ((i + 5) // 2), 2 ** ((i + 5) // 2), dtype=dt) for i in range(g.num_layers)]
            img, lat = g([z], noise=noise, return_latents=True)
            img_mix, _ = g([z, z2], noise=noise, inject_index=3)
            mean_w = g.style(rnd(seed * 10 + 2, 16, sdim, dtype=dt)).mean(0, keepdim=True)
            img_trunc, _ = g([z], noise=noise, truncation=0.7, truncation_latent=mean_w)
            img_fixed, _ = g([z], randomize_noise=False)
            d.update({f'{name}.{tag}.img': img, f'{name}.{tag}.latent': lat, f'{name}.{tag}.img_mix': img_mix,
                      f'{name}.{tag}.img_trunc': img_trunc, f'{name}.{tag}.mean_w': mean_w,
                      f'{name}.{tag}.img_fixed_noise': img_fixed})
            if dt == torch.float64:
                d[f'{name}.cfg'] = np.array([size, sdim, n_mlp, seed])
                # path-length regulariser (gt:601-614) with the trainer's own function
                torch.manual_seed(seed)
                img2, lat2 = g([z], noise=noise, return_latents=True)
                pen, mean, lengths = trainer.g_path_regularize(img2, lat2, 0.0)
                g.zero_grad()
                pen.backward()
                d.update({f'{name}.pl.penalty': pen, f'{name}.pl.mean': mean, f'{name}.pl.lengths': lengths,
                          f'{name}.pl.g.conv1.conv.weight': g.conv1.conv.weight.grad,
                          f'{name}.pl.g.convs.0.conv.weight': g.convs[0].conv.weight.grad,
                          f'{name}.pl.g.convs.1.conv.modulation.weight': g.convs[1].conv.modulation.weight.grad,
                          f'{name}.pl.g.to_rgbs.0.conv.weight': g.to_rgbs[0].conv.weight.grad,
                          f'{name}.pl.g.convs.2.noise.weight': g.convs[2].noise.weight.grad,
                          f'{name}.pl.g.input.input': g.input.input.grad})
                first = [k for k in shapes if k.startswith('style.') and k.endswith('.weight')][0]
                mod = g
                for part in first.split('.')[:-1]:
                    mod = getattr(mod, part) if not part.isdigit() else mod[int(part)]
                d[f'{name}.pl.g.{first}'] = mod.weight.grad
                # plain G loss grads through D-free objective
                g.zero_grad()
                img3, _ = g([z], noise=noise)
                cot = rnd(seed * 10 + 5, *img3.shape)
                (img3 * cot).sum().backward()
                d.update({f'{name}.bw.cot': cot,
                          f'{name}.bw.g.conv1.conv.weight': g.conv1.conv.weight.grad,
                          f'{name}.bw.g.convs.3.conv.weight': g.convs[3].conv.weight.grad,
                          f'{name}.bw.g.convs.2.activate.bias': g.convs[2].activate.bias.grad,
                          f'{name}.bw.g.convs.2.noise.weight': g.convs[2].noise.weight.grad,
                          f'{name}.bw.g.to_rgbs.1.bias': g.to_rgbs[1].bias.grad,
                          f'{name}.bw.g.to_rgbs.1.conv.modulation.bias': g.to_rgbs[1].conv.modulation.bias.grad,
                          f'{name}.bw.g.{first}': mod.weight.grad})
    # ---- discriminator size 16, batch 8 (two stddev groups) ----------------
    for dt, tag in [(torch.float64, 'f64'), (torch.float32, 'f32')]:
        dnet = gm.Discriminator(16, channel_multiplier=2).to(dt)
        shapes = P.discriminator_shapes(16, 2)
        load_seeded(dnet, shapes, 21, dt)
        x = rnd(210, 8, 3, 16, 16, dtype=dt).requires_grad_(True)
        pred, _ = dnet(x)
        d[f'd16.{tag}.pred'] = pred
        if dt == torch.float64:
            r1 = trainer.d_r1_loss(None, pred, x)
            dnet.zero_grad()
            (0.5 * r1 * 16 + 0 * pred[0]).sum().backward()
            d.update({'d16.r1': r1,
                      'd16.r1.g.convs.0.0.weight': dnet.convs[0][0].weight.grad,
                      'd16.r1.g.convs.1.conv1.0.weight': dnet.convs[1].conv1[0].weight.grad,
                      'd16.r1.g.convs.1.conv2.1.weight': dnet.convs[1].conv2[1].weight.grad,
                      'd16.r1.g.convs.1.conv2.2.bias': dnet.convs[1].conv2[2].bias.grad,
                      'd16.r1.g.convs.2.skip.1.weight': dnet.convs[2].skip[1].weight.grad,
                      'd16.r1.g.final_conv.0.weight': dnet.final_conv[0].weight.grad,
                      'd16.r1.g.final_linear.0.weight': dnet.final_linear[0].weight.grad})
            dnet.zero_grad()
            pred2, _ = dnet(x.detach())
            fake2, _ = dnet(rnd(211, 8, 3, 16, 16))
            loss = trainer.d_logistic_loss(pred2, fake2)
            loss.backward()
            d.update({'d16.dloss': loss, 'd16.gloss': trainer.g_nonsaturating_loss(fake2),
                      'd16.dl.g.convs.0.0.weight': dnet.convs[0][0].weight.grad,
                      'd16.dl.g.convs.1.conv1.1.bias': dnet.convs[1].conv1[1].bias.grad,
                      'd16.dl.g.convs.2.conv2.1.weight': dnet.convs[2].conv2[1].weight.grad,
                      'd16.dl.g.final_linear.1.weight': dnet.final_linear[1].weight.grad})
    save('networks', d)


def gen_config1(gm):
    """BASELINE.json configs[0]: Generator(256) batch 1 fp32 CPU, seeded params."""
    g = gm.Generator(256, 512, 8, channel_multiplier=2, conv_transpose=True)
    load_seeded(g, P.generator_shapes(256, 512, 8, 2), 31, torch.float32)
    g.eval()
    z = rnd(310, 1, 512, dtype=torch.float32)
    with torch.no_grad():
        img, _ = g([z], randomize_noise=False)
    save('config1_g256', {'z': z, 'img': img.half(), 'img_mean': img.double().mean(), 'img_std': img.double().std(),
                          'img_row': img[0, :, 128, :].clone()})


def gen_fcstack(gm):
    sys.path.insert(0, '/root/reference/src')
    from gan_control.models.controller_model import FcStack
    m = FcStack(0.01, 4, 3, 32, 24).double()
    shapes = P.fc_stack_shapes(4, 3, 32, 24)
    load_seeded(m, shapes, 41, torch.float64)
    x = rnd(410, 7, 3)
    save('fcstack', {'x': x, 'y': m(x), 'cfg': np.array([4, 3, 32, 24, 41])})


if __name__ == '__main__':
    torch.set_num_threads(8)
    random.seed(0)
    gm, trainer = import_reference()
    assert trainer is not None, 'trainer import failed'
    gen_upfirdn2d(gm)
    gen_bias_act(gm)
    gen_linear(gm)
    gen_modconv(gm)
    gen_fcstack(gm)
    gen_networks(gm, trainer)
    gen_config1(gm)


def gen_config5_controls():
    """BASELINE.json configs[4] inputs: the 1000 `orientation` control vectors and `latents_w` rows of the
    reference's data fixture resources/ffhq_1K_attributes_samples_df.pkl (derived arrays only)."""
    import pandas as pd
    df = pd.read_pickle('/root/reference/resources/ffhq_1K_attributes_samples_df.pkl')
    ori = np.stack(df['orientation'].values).astype(np.float32)
    lw = np.stack(df['latents_w'].values).astype(np.float16)
    np.savez_compressed(os.path.join(OUT, 'config5_controls.npz'), orientation=ori, latents_w=lw)
