"""CPU oracle for the StyleGAN2 G/D hot path of amazon-science/gan-control.

TEST INFRASTRUCTURE ONLY.  This file is a plain-PyTorch-on-CPU restatement of
the reference's ``FUSED = False`` arithmetic.  Only ``tests/``,
``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` / ``--impl
reference`` legs may import it, and only as the checker (or as the timed CPU
baseline) -- never as the product path.  The product (``gan_control_b200``)
fails loudly when its CUDA library is missing; it never routes through here.

Parity pinning: the reference holds NO tests, golden vectors or KATs for this
path.  The oracle is instead pinned against outputs of the
reference itself run in the build container: ``oracle/make_golden.py`` imports
``/root/reference/src/gan_control/models/gan_model.py`` on CPU, runs it on
seeded inputs and writes ``tests/golden/*.npz``; ``tests/test_oracle_golden.py``
replays every fixture through this file.

All citations are ``file:line`` relative to ``/root/reference/src/gan_control``;
``gm`` = ``models/gan_model.py``, ``up`` = ``models/pytorch_upfirdn2d.py``,
``gt`` = ``trainers/generator_trainer.py``, ``tu`` = ``trainers/utils.py``.

Everything is functional: networks are evaluated straight from a
``state_dict``-shaped mapping (reference key layout), so reference checkpoints
and the product's checkpoints are both valid inputs.
"""
import math

import torch
import torch.nn.functional as F

SQRT2 = math.sqrt(2.0)


# ---------------------------------------------------------------------------
# a1  upfirdn2d                                                   (up:9-51)
# ---------------------------------------------------------------------------
def fir_kernel(taps, gain=1.0, dtype=torch.float32):
    """Separable -> 2-D normalised FIR taps (gm:60-68); ``gain`` is the
    ``factor**2`` of gm:76 / gm:119-120."""
    k = torch.as_tensor(taps, dtype=dtype)
    if k.ndim == 1:
        k = torch.outer(k, k)
    return k / k.sum() * gain


def upfirdn2d(x, kernel, up=1, down=1, pad=(0, 0)):
    """gm:45-50 -> up:9-51.  Per channel: zero-insert (sample at i*up, zeros
    after, up:19-21), pad / crop (up:23-31), TRUE convolution with ``kernel``
    (flip + correlation, up:37-38), keep every ``down``-th sample (up:46)."""
    n, c, h, w = x.shape
    p0, p1 = pad
    kh, kw = kernel.shape
    z = x.new_zeros(n, c, h * up, w * up)
    z[:, :, ::up, ::up] = x
    z = F.pad(z, [max(p0, 0), max(p1, 0), max(p0, 0), max(p1, 0)])
    z = z[:, :, max(-p0, 0): z.shape[2] - max(-p1, 0), max(-p0, 0): z.shape[3] - max(-p1, 0)]
    flipped = torch.flip(kernel, [0, 1]).to(x.dtype)[None, None].expand(c, 1, kh, kw)
    y = F.conv2d(z, flipped, groups=c)
    return y[:, :, ::down, ::down]


# ---------------------------------------------------------------------------
# a2  fused_leaky_relu                                            (gm:25-41)
# ---------------------------------------------------------------------------
def fused_leaky_relu(x, bias, negative_slope=0.2, scale=SQRT2):
    shape = (1, -1) + (1,) * (x.ndim - 2)
    return scale * F.leaky_relu(x + bias.view(shape), negative_slope)


# ---------------------------------------------------------------------------
# a3/a4  EqualLinear, PixelNorm, mapping stacks        (gm:52-57,171-202,489-502)
# ---------------------------------------------------------------------------
def pixel_norm(x):
    return x * torch.rsqrt(torch.mean(x * x, dim=1, keepdim=True) + 1e-8)


def equal_linear(x, weight, bias, lr_mul=1.0, activation=False):
    """gm:189-197: scale = lr_mul / sqrt(in_dim); bias enters as bias*lr_mul;
    the activated form has gain sqrt(2)."""
    scale = lr_mul / math.sqrt(weight.shape[1])
    y = F.linear(x, weight * scale)
    if activation:
        return fused_leaky_relu(y, bias * lr_mul)
    return y + bias * lr_mul if bias is not None else y


def _stack_indices(sd, prefix):
    idx = set()
    for k in sd:
        if k.startswith(prefix) and k.endswith('.weight'):
            rest = k[len(prefix):].split('.')
            if len(rest) == 2 and rest[0].isdigit():
                idx.add(int(rest[0]))
    return sorted(idx)


def fc_stack(sd, prefix, x, lr_mul=0.01, normalize=True):
    """A ``Sequential(PixelNorm, EqualLinear*n)`` (gm:633-642, 658-681) or, with
    ``normalize=False``, the controller ``FcStack`` (controller_model.py:24-43)."""
    if normalize:
        x = pixel_norm(x)
    for i in _stack_indices(sd, prefix):
        x = equal_linear(x, sd[f'{prefix}{i}.weight'], sd[f'{prefix}{i}.bias'], lr_mul, True)
    return x


def mapping_network(sd, z, fc_groups=None, lr_mul=0.01):
    """``Generator.style``.  ``fc_groups`` = None for the vanilla stack
    (state_dict keys ``style.<k>.*``) or an ordered list of
    ``(group_name, lo, hi)`` for ``MultiFcStack`` (gm:489-502; keys
    ``style.<group>.<k>.*``; PixelNorm is per slice)."""
    if fc_groups is None:
        return fc_stack(sd, 'style.', z, lr_mul)
    outs = [fc_stack(sd, f'style.{name}.', z[:, lo:hi], lr_mul) for name, lo, hi in fc_groups]
    return torch.cat(outs, dim=1)


# ---------------------------------------------------------------------------
# a5  ModulatedConv2d                                            (gm:281-331)
# ---------------------------------------------------------------------------
def modulated_conv2d(x, style, weight, mod_weight, mod_bias, demodulate=True,
                     upsample=False, blur=None):
    """Reference formulation: per-sample weights + grouped convolution.
    ``weight`` is the ``(1, OC, IC, k, k)`` parameter, ``blur`` the
    ``conv.blur.kernel`` buffer (already carries gain 4, gm:119-120)."""
    b, ic, h, w = x.shape
    _, oc, _, k, _ = weight.shape
    s = equal_linear(style, mod_weight, mod_bias)                       # gm:284
    wmod = weight * (1.0 / math.sqrt(ic * k * k)) * s.view(b, 1, ic, 1, 1)   # gm:285
    if demodulate:
        wmod = wmod * torch.rsqrt(wmod.pow(2).sum([2, 3, 4], keepdim=True) + 1e-8)  # gm:288
    if upsample:
        wt = wmod.transpose(1, 2).reshape(b * ic, oc, k, k)             # gm:301-303
        y = F.conv_transpose2d(x.reshape(1, b * ic, h, w), wt, stride=2, padding=0, groups=b)
        y = y.view(b, oc, y.shape[2], y.shape[3])
        # gm:245-249: pad = (1, 1) for k=3, 4-tap blur
        p = (blur.shape[0] - 2) - (k - 1)
        return upfirdn2d(y, blur, pad=((p + 1) // 2 + 1, p // 2 + 1))
    y = F.conv2d(x.reshape(1, b * ic, h, w), wmod.view(b * oc, ic, k, k), padding=k // 2, groups=b)
    return y.view(b, oc, y.shape[2], y.shape[3])


def styled_conv(sd, prefix, x, style, noise, upsample):
    """gm:402-408: conv -> noise injection (gm:340-345) -> FusedLeakyReLU."""
    y = modulated_conv2d(x, style, sd[prefix + 'conv.weight'],
                         sd[prefix + 'conv.modulation.weight'], sd[prefix + 'conv.modulation.bias'],
                         True, upsample, sd.get(prefix + 'conv.blur.kernel'))
    if noise is None:
        noise = torch.randn(y.shape[0], 1, y.shape[2], y.shape[3], dtype=y.dtype, device=y.device)
    y = y + sd[prefix + 'noise.weight'] * noise
    return fused_leaky_relu(y, sd[prefix + 'activate.bias'])


def to_rgb(sd, prefix, x, style, skip=None):
    """gm:424-435: 1x1 modconv without demod + bias, + Upsample(skip) with
    pad (2, 1) (gm:79-84) and kernel*4."""
    y = modulated_conv2d(x, style, sd[prefix + 'conv.weight'],
                         sd[prefix + 'conv.modulation.weight'], sd[prefix + 'conv.modulation.bias'],
                         demodulate=False)
    y = y + sd[prefix + 'bias']
    if skip is not None:
        y = y + upfirdn2d(skip, sd[prefix + 'upsample.kernel'], up=2, pad=(2, 1))
    return y


# ---------------------------------------------------------------------------
# a10  Generator.forward                                         (gm:709-801)
# ---------------------------------------------------------------------------
def generator_forward(sd, styles, size, fc_groups=None, input_is_latent=False,
                      noise=None, inject_index=None, truncation=1.0,
                      truncation_latent=None, return_latents=False):
    """``styles`` is a list of (B,512) tensors (or one (B,n_latent,512) W+).
    ``noise``: list of num_layers tensors, or None for fresh normal noise."""
    log_size = int(math.log2(size))
    n_latent = 2 * log_size - 2
    num_layers = 2 * (log_size - 2) + 1
    if not input_is_latent:
        styles = [mapping_network(sd, s, fc_groups) for s in styles]
    if noise is None:
        noise = [None] * num_layers
    if truncation < 1:
        styles = [truncation_latent + truncation * (s - truncation_latent) for s in styles]
    if len(styles) < 2:
        latent = styles[0] if styles[0].ndim == 3 else styles[0].unsqueeze(1).repeat(1, n_latent, 1)
    else:
        assert inject_index is not None, 'oracle wants an explicit inject_index (gm:764 draws it)'
        latent = torch.cat([styles[0].unsqueeze(1).repeat(1, inject_index, 1),
                            styles[1].unsqueeze(1).repeat(1, n_latent - inject_index, 1)], 1)
    b = latent.shape[0]
    out = sd['input.input'].repeat(b, 1, 1, 1)                           # gm:354-358
    out = styled_conv(sd, 'conv1.', out, latent[:, 0], noise[0], False)
    skip = to_rgb(sd, 'to_rgb1.', out, latent[:, 1])
    i = 1
    for r in range(log_size - 2):
        out = styled_conv(sd, f'convs.{2 * r}.', out, latent[:, i], noise[2 * r + 1], True)
        out = styled_conv(sd, f'convs.{2 * r + 1}.', out, latent[:, i + 1], noise[2 * r + 2], False)
        skip = to_rgb(sd, f'to_rgbs.{r}.', out, latent[:, i + 2], skip)
        i += 2
    return (skip, latent) if return_latents else skip


# ---------------------------------------------------------------------------
# a11/a12  Discriminator                                   (gm:132-168,844-1016)
# ---------------------------------------------------------------------------
def equal_conv2d(x, weight, bias=None, stride=1, padding=0):
    oc, ic, k, _ = weight.shape
    return F.conv2d(x, weight * (1.0 / math.sqrt(ic * k * k)), bias, stride, padding)


def conv_layer(sd, prefix, x, k, downsample=False, activate=True, blur=None):
    """gm:844-890.  Sequential indices: [Blur], EqualConv2d, [FusedLeakyReLU]."""
    i = 0
    if downsample:
        p = (blur.shape[0] - 2) + (k - 1)                                # gm:859-861
        x = upfirdn2d(x, blur, pad=((p + 1) // 2, p // 2))
        i = 1
    y = equal_conv2d(x, sd[f'{prefix}{i}.weight'], sd.get(f'{prefix}{i}.bias'),
                     2 if downsample else 1, 0 if downsample else k // 2)
    if activate:
        y = fused_leaky_relu(y, sd[f'{prefix}{i + 1}.bias'])
    return y


def minibatch_stddev(x, group_size=4):
    """gm:1003-1012: strided groups, biased variance, +1e-8, one feature."""
    b, c, h, w = x.shape
    g = min(b, group_size)
    y = x.view(g, -1, 1, c, h, w)
    y = torch.sqrt(y.var(0, unbiased=False) + 1e-8)
    y = y.mean([2, 3, 4], keepdim=True).squeeze(2)
    return torch.cat([x, y.repeat(g, 1, h, w)], 1)


def discriminator_forward(sd, x, size):
    log_size = int(math.log2(size))
    blur = fir_kernel([1, 3, 3, 1], dtype=x.dtype)
    out = conv_layer(sd, 'convs.0.', x, 1)
    for j in range(1, log_size - 1):                                     # ResBlocks gm:907-922
        p = f'convs.{j}.'
        y = conv_layer(sd, p + 'conv1.', out, 3)
        y = conv_layer(sd, p + 'conv2.', y, 3, downsample=True, blur=sd.get(p + 'conv2.0.kernel', blur))
        s = conv_layer(sd, p + 'skip.', out, 1, downsample=True, activate=False,
                       blur=sd.get(p + 'skip.0.kernel', blur))
        out = (y + s) / SQRT2
    out = minibatch_stddev(out)
    out = conv_layer(sd, 'final_conv.', out, 3)
    out = out.reshape(out.shape[0], -1)
    out = equal_linear(out, sd['final_linear.0.weight'], sd['final_linear.0.bias'], 1.0, True)
    return equal_linear(out, sd['final_linear.1.weight'], sd['final_linear.1.bias'])


# ---------------------------------------------------------------------------
# a13  losses / regularisers                      (gt:563-566,601-624,690-719)
# ---------------------------------------------------------------------------
def g_nonsaturating_loss(fake_pred):
    return F.softplus(-fake_pred).mean()


def d_logistic_loss(real_pred, fake_pred):
    return F.softplus(-real_pred).mean() + F.softplus(fake_pred).mean()


def d_r1_loss(real_pred, real_img):
    grad, = torch.autograd.grad(real_pred.sum(), real_img, create_graph=True)
    return grad.pow(2).reshape(grad.shape[0], -1).sum(1).mean()


def path_lengths_from_grad(grad, mean_path_length, decay=0.01):
    """gt:617-624."""
    lengths = torch.sqrt(grad.pow(2).sum(2).mean(1))
    mean = mean_path_length + decay * (lengths.mean() - mean_path_length)
    penalty = (lengths - mean).pow(2).mean()
    return penalty, mean.detach(), lengths


def g_path_regularize(fake_img, latents, mean_path_length, pl_noise=None, decay=0.01):
    """gt:601-614 / gm:803-811.  ``pl_noise`` overrides randn_like for tests."""
    if pl_noise is None:
        pl_noise = torch.randn_like(fake_img)
    pl_noise = pl_noise / math.sqrt(fake_img.shape[2] * fake_img.shape[3])
    grad, = torch.autograd.grad((fake_img * pl_noise).sum(), latents, create_graph=True)
    return path_lengths_from_grad(grad, mean_path_length, decay)


def ema_accumulate(ema_sd, sd, decay):
    """tu:8-12 (parameters only; buffers are not averaged)."""
    for k, v in ema_sd.items():
        v.mul_(decay).add_(sd[k].detach(), alpha=1 - decay)


def lazy_adam_hparams(lr, reg_every):
    """gt:158-173: lr*c, betas=(0**c, 0.99**c), c = k/(k+1)."""
    c = reg_every / (reg_every + 1)
    return lr * c, (0 ** c, 0.99 ** c)
