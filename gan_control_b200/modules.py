"""StyleGAN2 layers and networks with the reference's constructor / forward signatures and
``state_dict`` layout (``gan_control/models/gan_model.py``), computed by libb200gan.

What is kept identical to the reference: class names, constructor arguments, sub-module and
parameter names / shapes (so checkpoints interchange), forward signatures and
semantics.  What is different: how the work is issued.

* ``ModulatedConv2d`` never materialises the reference's 5-D per-sample weight + grouped
  convolution when that is wasteful.  Two algebraically identical forms:
  "weight-modulated" ``d * conv(x, W*s)`` with per-sample weights (cheap when the image is large
  and the weight small) and "activation-modulated" ``d * conv(x*s, W)`` with one shared weight
  (cheap when the weight is large: its gradient is then a single batch-summed GEMM).
* ``StyledConv`` fuses demodulation scale, noise, bias and leaky-ReLU into one epilogue pass.
* activations run channels-last in ``act_dtype`` (fp32 for parity runs, bf16 for throughput);
  parameters stay fp32.
"""
import math
import random

import torch
from torch import nn

from . import ops

SQRT2 = math.sqrt(2.0)


def make_kernel(k):
    """gm.py:60-68"""
    k = torch.tensor(k, dtype=torch.float32)
    if k.ndim == 1:
        k = k[None, :] * k[:, None]
    return k / k.sum()


class PixelNorm(nn.Module):                                                   # gm.py:52-57
    def forward(self, input):
        return input * torch.rsqrt(torch.mean(input * input, dim=1, keepdim=True) + 1e-8)


class Upsample(nn.Module):                                                    # gm.py:71-89
    def __init__(self, kernel, factor=2):
        super().__init__()
        self.factor = factor
        self.register_buffer('kernel', make_kernel(kernel) * (factor ** 2))
        p = self.kernel.shape[0] - factor
        self.pad = ((p + 1) // 2 + factor - 1, p // 2)

    def forward(self, input):
        return ops.upfirdn2d(input, self.kernel, up=self.factor, down=1, pad=self.pad)


class Downsample(nn.Module):                                                  # gm.py:92-110
    def __init__(self, kernel, factor=2):
        super().__init__()
        self.factor = factor
        self.register_buffer('kernel', make_kernel(kernel))
        p = self.kernel.shape[0] - factor
        self.pad = ((p + 1) // 2, p // 2)

    def forward(self, input):
        return ops.upfirdn2d(input, self.kernel, up=1, down=self.factor, pad=self.pad)


class Blur(nn.Module):                                                        # gm.py:113-129
    def __init__(self, kernel, pad, upsample_factor=1):
        super().__init__()
        kernel = make_kernel(kernel)
        if upsample_factor > 1:
            kernel = kernel * (upsample_factor ** 2)
        self.register_buffer('kernel', kernel)
        self.pad = pad

    def forward(self, input):
        return ops.upfirdn2d(input, self.kernel, pad=self.pad)


class FusedLeakyReLU(nn.Module):                                              # gm.py:25-36
    def __init__(self, channel, negative_slope=0.2, scale=2 ** 0.5):
        super().__init__()
        self.bias = nn.Parameter(torch.zeros(channel))
        self.negative_slope = negative_slope
        self.scale = scale

    def forward(self, input):
        return ops.fused_leaky_relu(input, self.bias, self.negative_slope, self.scale)


class ScaledLeakyReLU(nn.Module):                                             # gm.py:205-214
    def __init__(self, negative_slope=0.2):
        super().__init__()
        self.negative_slope = negative_slope

    def forward(self, input):
        return ops.fused_leaky_relu(input, None, self.negative_slope, SQRT2)


class EqualConv2d(nn.Module):                                                 # gm.py:132-168
    def __init__(self, in_channel, out_channel, kernel_size, stride=1, padding=0, bias=True):
        super().__init__()
        self.weight = nn.Parameter(torch.randn(out_channel, in_channel, kernel_size, kernel_size))
        self.scale = 1 / math.sqrt(in_channel * kernel_size ** 2)
        self.stride = stride
        self.padding = padding
        self.bias = nn.Parameter(torch.zeros(out_channel)) if bias else None

    def forward(self, input):
        return ops.conv2d(input, self.weight * self.scale, self.bias, self.stride, self.padding)

    def __repr__(self):
        return (f'{self.__class__.__name__}({self.weight.shape[1]}, {self.weight.shape[0]},'
                f' {self.weight.shape[2]}, stride={self.stride}, padding={self.padding})')


class EqualLinear(nn.Module):                                                 # gm.py:171-202
    def __init__(self, in_dim, out_dim, bias=True, bias_init=0, lr_mul=1, activation=None):
        super().__init__()
        self.weight = nn.Parameter(torch.randn(out_dim, in_dim).div_(lr_mul))
        self.bias = nn.Parameter(torch.zeros(out_dim).fill_(bias_init)) if bias else None
        self.activation = activation
        self.scale = (1 / math.sqrt(in_dim)) * lr_mul
        self.lr_mul = lr_mul

    def forward(self, input):
        lead = input.shape[:-1]
        x = ops.up32(input.reshape(-1, input.shape[-1]))
        y = ops.equal_linear(x, self.weight, self.bias, self.scale, self.lr_mul, bool(self.activation))
        return y.reshape(*lead, y.shape[-1])

    def __repr__(self):
        return f'{self.__class__.__name__}({self.weight.shape[1]}, {self.weight.shape[0]})'


class ModulatedConv2d(nn.Module):                                             # gm.py:217-331
    def __init__(self, in_channel, out_channel, kernel_size, style_dim, demodulate=True, upsample=False,
                 downsample=False, blur_kernel=[1, 3, 3, 1], conv_transpose=False, overwrite_padding=None):
        super().__init__()
        if not conv_transpose:                                                # gm.py:232-233
            raise ValueError('conv_transpose is %s' % str(conv_transpose))
        if downsample or overwrite_padding is not None:
            raise NotImplementedError('downsample / overwrite_padding ("896" mode) are dead code in every '
                                      'shipped config (SURVEY.md F6) and are not built')
        self.eps = 1e-8
        self.kernel_size = kernel_size
        self.in_channel = in_channel
        self.out_channel = out_channel
        self.upsample = upsample
        self.downsample = downsample
        self.conv_transpose = conv_transpose
        if upsample:
            factor = 2
            p = (len(blur_kernel) - factor) - (kernel_size - 1)
            self.blur = Blur(blur_kernel, pad=((p + 1) // 2 + factor - 1, p // 2 + 1), upsample_factor=factor)
            if kernel_size == 3 and len(blur_kernel) == 4:
                # not a parameter and not in the state_dict: the FIR as a (9 -> 36)-tap Toeplitz map (ops.composite_up)
                self.register_buffer('fir_toeplitz', ops.fir_toeplitz(self.blur.kernel, False), persistent=False)
        self.scale = 1 / math.sqrt(in_channel * kernel_size ** 2)
        self.padding = kernel_size // 2
        self.weight = nn.Parameter(torch.randn(1, out_channel, in_channel, kernel_size, kernel_size))
        self.modulation = EqualLinear(style_dim, in_channel, bias_init=1)
        self.demodulate = demodulate
        self.form = 'auto'          # 'auto' | 'weight' | 'activation'  (see module docstring)

    def __repr__(self):
        return (f'{self.__class__.__name__}({self.in_channel}, {self.out_channel}, {self.kernel_size}, '
                f'upsample={self.upsample}, downsample={self.downsample})')

    def _weight_form(self, height, width):
        if self.form != 'auto':
            return self.form == 'weight'
        # per-sample weights cost B*OC*IC*k^2 elements, activation scaling costs B*IC*H*W
        return self.out_channel * self.kernel_size ** 2 <= height * width

    def operands(self, input, style):
        """(x, wk, d, shared): the convolution operands of one of the two equivalent forms (module docstring), the
        demodulation coefficients d still to be applied to the OUTPUT (None when demodulate=False or when they are
        already folded into per-sample weights) and whether wk is the shared, parameter-only weight
        of the activation-modulated form (ops.conv_gather `param_weight`; the weight-modulated wk carries the style
        even when the batch is 1)."""
        batch, in_channel, height, width = input.shape
        s = self._style(style)                                                # (B, IC) fp32, gm.py:284
        w = self.weight[0] * self.scale                                       # (OC, IC, k, k)
        d = None
        if self.demodulate:                                                   # gm.py:287-289
            wsq = w.pow(2).sum((2, 3))                                        # (OC, IC)
            d = torch.rsqrt(ops._Gemm.apply(s * s, wsq, False, True, 1.0) + 1e-8)
        if self._weight_form(height, width):
            wk = w.unsqueeze(0) * s.view(batch, 1, in_channel, 1, 1)          # (B, OC, IC, k, k)
            if d is not None:
                # demodulation folded into the per-sample weights, as the reference does (gm.py:289): the B*OC*IC*k^2
                # product is tiny next to the activation, and the convolution then needs no per-(sample, channel) scale
                # in its epilogue -- under create_graph (path length) that removes the recomputation of the convolution
                # output for d's gradient and two full-size passes (scale, dot) per layer and per differentiation order
                wk = wk * d.view(batch, self.out_channel, 1, 1, 1)
                d = None
            return input, wk, d, False
        wk = w.unsqueeze(0)
        x = input * s.view(batch, in_channel, 1, 1).to(input.dtype)
        return x, wk, d, True

    _s_pre = None       # (B, IC) modulation computed for all layers at once by Generator.synthesis, consumed by the next call

    def _style(self, style):
        """s = modulation(style) (gm.py:284), or the value the generator computed for all layers in one launch"""
        if self._s_pre is not None:
            s, self._s_pre = self._s_pre, None
            return s
        return self.modulation(style)

    def demod_coeff(self, s):
        """d[b,o] = rsqrt(sum_{i,k} (scale*W*s)^2 + 1e-8) (gm.py:287-289) as a (B,IC)x(IC,OC) product"""
        w = self.weight[0] * self.scale
        return torch.rsqrt(ops._Gemm.apply(s * s, w.pow(2).sum((2, 3)), False, True, 1.0) + 1e-8)

    def fused(self, input, style, noise=None, noise_w=None, bias=None, slope=1.0, gain=1.0, transpose=False):
        """The layer through `ops.mod_conv` (first-order passes only, ops.first_order): weight path, convolution and
        epilogue in two kernels.  transpose=True: the stride-2 transposed convolution of the upsampling layer
        (gm.py:301-306) WITHOUT its blur (the caller fuses that with the epilogue); then no epilogue here.
        Returns (y, d): d is None unless the demodulation is still to be applied by the caller (activation-modulated
        transposed convolution)."""
        batch, in_channel, height, width = input.shape
        s = self._style(style)                                                # (B, IC) fp32, gm.py:284
        k = self.kernel_size
        geom = dict(flip=True, up=2, pad0=k - 1, out_hw=((height - 1) * 2 + k, (width - 1) * 2 + k)) if transpose \
            else dict(pad0=self.padding)
        if self._weight_form(height, width):
            # per-sample weights with the demodulation folded in (what the reference itself convolves with, gm.py:289)
            y = ops.mod_conv(input, s, self.weight, None, noise, noise_w, bias, self.scale, self.demodulate, slope=slope,
                             gain=gain, **geom)
            return y, None
        x = input * s.view(batch, in_channel, 1, 1).to(input.dtype)
        d = self.demod_coeff(s) if self.demodulate else None
        if transpose:
            return ops.mod_conv(x, None, self.weight, scale=self.scale, **geom), d
        return ops.mod_conv(x, None, self.weight, d, noise, noise_w, bias, self.scale, slope=slope, gain=gain, **geom), None

    # Above this many input channels the upsampling layer is bound by the tensor pipe, not by HBM, and the fused
    # single-pass form (4x the MMA work of the transposed convolution, no (2H+1)^2 intermediate) stops paying.
    FUSE_UP_MAX_IN_CHANNELS = 64

    def fuses_up(self, height, width):
        return (self.upsample and hasattr(self, 'fir_toeplitz') and self.in_channel <= self.FUSE_UP_MAX_IN_CHANNELS
                and self.in_channel % 8 == 0 and self.out_channel % 4 == 0 and self._weight_form(height, width))

    def raw(self, input, style, blur=True):
        """Un-demodulated convolution and the demodulation coefficients: (z, d) with
        ModulatedConv2d(x, style) == z * d[:, :, None, None].  blur=False leaves the upsampling layer's Blur
        (gm.py:307) to the caller (StyledConv fuses it with its epilogue)."""
        height, width = input.shape[2], input.shape[3]
        x, wk, d, shared = self.operands(input, style)
        k = self.kernel_size
        if self.upsample:
            # conv_transpose2d(stride 2, padding 0) in gather form (gm.py:301-306) ...
            z = ops.conv_gather(x, wk.flip(3, 4), up=2, down=1, pad0=k - 1,
                                out_hw=((height - 1) * 2 + k, (width - 1) * 2 + k), param_weight=shared)
            if blur:
                z = self.blur(z)                                              # ... then Blur (gm.py:307)
        else:
            z = ops.conv_gather(x, wk, 1, 1, self.padding, param_weight=shared)
        return z, d

    def forward(self, input, style):
        z, d = self.raw(input, style)
        return z if d is None else ops.mod_epilogue(z, d, slope=1.0, gain=1.0)


class NoiseInjection(nn.Module):                                              # gm.py:334-345
    def __init__(self):
        super().__init__()
        self.weight = nn.Parameter(torch.zeros(1))

    def forward(self, image, noise=None):
        if noise is None:
            batch, _, height, width = image.shape
            noise = image.new_empty(batch, 1, height, width).normal_()
        return image + (self.weight * ops.up32(noise)).to(image.dtype)


class ConstantInput(nn.Module):                                               # gm.py:348-358
    def __init__(self, channel, size=4):
        super().__init__()
        self.input = nn.Parameter(torch.randn(1, channel, size, size))

    def forward(self, input):
        return self.input.repeat(input.shape[0], 1, 1, 1)


class StyledConv(nn.Module):                                                  # gm.py:361-408
    def __init__(self, in_channel, out_channel, kernel_size, style_dim, upsample=False,
                 blur_kernel=[1, 3, 3, 1], demodulate=True, conv_transpose=False, overwrite_padding=None,
                 noise_mode='normal'):
        super().__init__()
        if noise_mode not in ('normal', 'same_for_same_id'):
            raise NotImplementedError(f'noise_mode={noise_mode!r} is not used by any shipped config')
        self.conv = ModulatedConv2d(in_channel, out_channel, kernel_size, style_dim, upsample=upsample,
                                    blur_kernel=blur_kernel, demodulate=demodulate,
                                    conv_transpose=conv_transpose, overwrite_padding=overwrite_padding)
        self.noise_mode = noise_mode
        self.noise = NoiseInjection()
        self.activate = FusedLeakyReLU(out_channel)

    def forward(self, input, style, noise=None):
        conv = self.conv
        if conv.upsample and conv.fuses_up(input.shape[2], input.shape[3]):
            # the WHOLE upsampling StyledConv in one kernel: transposed stride-2 conv, 4x4 FIR, demodulation, noise,
            # bias, leaky-ReLU (gm.py:295-307, 340-345, 32-35) = a 3x3 convolution with composite FIR (*) conv
            # weights whose accumulator tile is stored depth-to-space through the fused epilogue
            x, wk, d, shared = conv.operands(input, style)
            oh, ow = input.shape[2] * 2, input.shape[3] * 2
            if noise is None:
                noise = input.new_empty(input.shape[0], 1, oh, ow).normal_()
            return ops.conv_epilogue(x, ops.composite_up(wk, conv.fir_toeplitz), d, noise, self.noise.weight,
                                     self.activate.bias, 1, 1, 1, out_hw=(input.shape[2], input.shape[3]),
                                     slope=self.activate.negative_slope, gain=self.activate.scale, pack_out=True,
                                     param_weight=shared)
        if ops.fused_prep() and not conv.upsample:
            if noise is None:
                noise = input.new_empty(input.shape[0], 1, input.shape[2], input.shape[3]).normal_()
            return conv.fused(input, style, noise, self.noise.weight, self.activate.bias, self.activate.negative_slope,
                              self.activate.scale)[0]
        if conv.upsample:
            # transposed conv -> [blur + demod scale + noise + bias + leaky-ReLU*sqrt(2)] in one pass
            z, d = conv.fused(input, style, transpose=True) if ops.fused_prep() else conv.raw(input, style, blur=False)
            oh, ow = input.shape[2] * 2, input.shape[3] * 2
            if noise is None:
                noise = z.new_empty(z.shape[0], 1, oh, ow).normal_()
            return ops.fir_epilogue(z, conv.blur.kernel, conv.blur.pad, d, noise, self.noise.weight, self.activate.bias,
                                    slope=self.activate.negative_slope, gain=self.activate.scale)
        # plain layer: the whole StyledConv is ONE kernel (epilogue fused into the convolution)
        x, wk, d, shared = conv.operands(input, style)
        if noise is None:
            noise = input.new_empty(input.shape[0], 1, input.shape[2], input.shape[3]).normal_()
        return ops.conv_epilogue(x, wk, d, noise, self.noise.weight, self.activate.bias, 1, 1, conv.padding,
                                 slope=self.activate.negative_slope, gain=self.activate.scale, param_weight=shared)


class ToRGB(nn.Module):                                                       # gm.py:411-435
    def __init__(self, in_channel, style_dim, upsample=True, blur_kernel=[1, 3, 3, 1], out_channels=3,
                 conv_transpose=False, overwrite_negative_padding=None):
        super().__init__()
        if overwrite_negative_padding is not None:
            raise NotImplementedError('"896" mode is not built (SURVEY.md §8 a10)')
        if upsample:
            self.upsample = Upsample(blur_kernel)
        self.conv = ModulatedConv2d(in_channel, out_channels, 1, style_dim, demodulate=False,
                                    conv_transpose=conv_transpose)
        self.bias = nn.Parameter(torch.zeros(1, out_channels, 1, 1))

    def forward(self, input, style, skip=None):
        if ops.fused_prep():
            out = self.conv.fused(input, style, bias=self.bias)[0]            # 1x1 modulated conv + bias in one kernel
        else:
            out = self.conv(input, style)
            out = out + self.bias.to(out.dtype)
        if skip is not None:
            out = out + self.upsample(skip)
        return out


class MultiFcStack(nn.Module):                                                # gm.py:489-502
    def __init__(self, fc_dict, fc_config):
        super().__init__()
        self.fc_config = fc_config
        for group_name in fc_config.in_order_group_names:
            setattr(self, group_name, fc_dict[group_name])

    def forward(self, x):
        outs = []
        for name in self.fc_config.in_order_group_names:
            lo, hi = self.fc_config.groups[name]['latent_place']
            outs.append(getattr(self, name)(x[:, lo:hi]))
        return torch.cat(outs, dim=1)


class FcConfig:                                    # utils/mini_batch_multi_split_utils.py:13-16
    def __init__(self, in_order_group_names, groups):
        self.in_order_group_names = in_order_group_names
        self.groups = groups

    @classmethod
    def from_sub_groups_dict(cls, sub_groups_dict):
        """`MiniBatchUtils.get_fc_config` (mini_batch_multi_split_utils.py:46-55,103-115): groups in
        order of their first latent index."""
        names = sorted(sub_groups_dict, key=lambda n: sub_groups_dict[n]['place_in_latent'][0])
        groups = {n: {'latent_place': list(sub_groups_dict[n]['place_in_latent']),
                      'latent_size': sub_groups_dict[n]['place_in_latent'][1] - sub_groups_dict[n]['place_in_latent'][0]}
                  for n in names}
        return cls(names, groups)


def _cached_linear_batch(module, items, in_width, device):
    """layer table of `ops.build_linear_batch_table`, rebuilt only when a parameter moved (see _cached_fc_table)"""
    key = tuple(p.data_ptr() for lin, _ in items for p in (lin.weight, lin.bias)) + (in_width,)
    tbl = module.__dict__.get('_mod_table')
    if tbl is None or tbl['key'] != key or tbl['table'].device != device:
        tbl = ops.build_linear_batch_table(items, in_width, device)
        tbl['key'] = key
        module.__dict__['_mod_table'] = tbl
    return tbl


def _cached_fc_table(module, groups_fn, device):
    """The kernel's layer table holds raw parameter pointers: rebuild only when a parameter moved (the H2D
    copy of a fresh table must not happen inside CUDA-graph capture; warm-up builds it)."""
    groups = groups_fn()
    key = tuple(p.data_ptr() for _, _, layers in groups for lin in layers for p in (lin.weight, lin.bias))
    tbl = module.__dict__.get('_fc_table')
    if tbl is None or tbl['key'] != key or tbl['table'].device != device:
        tbl = ops.build_fc_table(groups, device)
        tbl['key'] = key        # group-major, as computed above (build_fc_table lists its tensors layer-major: with more than
        #                         one group the two orders differ and the table was rebuilt -- an H2D copy -- on every call)
        module.__dict__['_fc_table'] = tbl
    return tbl


def _channels(channel_multiplier):
    return {4: 512, 8: 512, 16: 512, 32: 512, 64: int(256 * channel_multiplier),
            128: int(128 * channel_multiplier), 256: int(64 * channel_multiplier),
            512: int(32 * channel_multiplier), 1024: int(16 * channel_multiplier)}


class Generator(nn.Module):                                                   # gm.py:505-811
    def __init__(self, size, style_dim, n_mlp, channel_multiplier=2, blur_kernel=[1, 3, 3, 1], lr_mlp=0.01,
                 out_channels=3, vae=False, bottleneck_size=256, split_fc=False, marge_fc=False, fc_config=None,
                 conv_transpose=False, model_mode='normal', noise_mode='normal', act_dtype=torch.float32):
        super().__init__()
        if vae or marge_fc or model_mode != 'normal':
            raise NotImplementedError('vae / marge_fc / "896" modes are not used by any shipped config '
                                      '(SURVEY.md §8 a10) and are not built')
        self.noise_mode = noise_mode
        self.model_mode = model_mode
        self.size = size
        self.vae = vae
        self.out_channels = out_channels
        self.fc_config = fc_config
        self.style_dim = style_dim
        self.act_dtype = act_dtype
        if split_fc:
            self.style = self.make_fc_stacks_using_fc_config(fc_config, lr_mlp, n_mlp)
        else:
            self.style = self.create_regular_fc_stack(lr_mlp, n_mlp, style_dim)
        self.channels = _channels(channel_multiplier)
        self.input = ConstantInput(self.channels[4])
        self.conv1 = StyledConv(self.channels[4], self.channels[4], 3, style_dim, blur_kernel=blur_kernel,
                                conv_transpose=conv_transpose, noise_mode=noise_mode)
        self.to_rgb1 = ToRGB(self.channels[4], style_dim, upsample=False, out_channels=out_channels,
                             conv_transpose=conv_transpose)
        self.log_size = int(math.log(size, 2))
        self.num_layers = (self.log_size - 2) * 2 + 1
        self.convs = nn.ModuleList()
        self.upsamples = nn.ModuleList()
        self.to_rgbs = nn.ModuleList()
        self.noises = nn.Module()
        for layer_idx in range(self.num_layers):
            res = 2 ** ((layer_idx + 5) // 2)
            self.noises.register_buffer(f'noise_{layer_idx}', torch.randn(1, 1, res, res))
        in_channel = self.channels[4]
        for i in range(3, self.log_size + 1):
            out_channel = self.channels[2 ** i]
            self.convs.append(StyledConv(in_channel, out_channel, 3, style_dim, upsample=True,
                                         blur_kernel=blur_kernel, conv_transpose=conv_transpose,
                                         noise_mode=noise_mode))
            self.convs.append(StyledConv(out_channel, out_channel, 3, style_dim, blur_kernel=blur_kernel,
                                         conv_transpose=conv_transpose))
            self.to_rgbs.append(ToRGB(out_channel, style_dim, out_channels=out_channels,
                                      conv_transpose=conv_transpose))
            in_channel = out_channel
        self.n_latent = self.log_size * 2 - 2

    # -- mapping-network builders (gm.py:619-681) ------------------------------------------------
    def make_fc_stacks_using_fc_config(self, fc_config, lr_mlp, n_mlp):
        stacks = {name: self.create_fc_stack(lr_mlp, n_mlp, fc_config.groups[name]['latent_size'], mid_dim=256)
                  for name in fc_config.in_order_group_names}
        return MultiFcStack(stacks, fc_config)

    def create_regular_fc_stack(self, lr_mlp, n_mlp, style_dim):
        return nn.Sequential(PixelNorm(), *[EqualLinear(style_dim, style_dim, lr_mul=lr_mlp, activation='fused_lrelu')
                                            for _ in range(n_mlp)])

    @staticmethod
    def create_fc_stack(lr_mlp, n_mlp, style_dim, mid_dim=None):
        layers = [PixelNorm()]
        for i in range(n_mlp):
            d0 = style_dim if i == 0 else mid_dim
            d1 = style_dim if i == n_mlp - 1 else mid_dim
            layers.append(EqualLinear(d0, d1, lr_mul=lr_mlp, activation='fused_lrelu'))
        return nn.Sequential(*layers)

    def load_transfer_learning_model(self, transfer_learning_model, load_only_main=True):   # gm.py:645-656
        missing, unexpected = self.load_state_dict(transfer_learning_model.state_dict(), strict=False)
        if (missing or unexpected) and not load_only_main:
            self.load_state_dict(transfer_learning_model.state_dict())
        for key in list(missing) + list(unexpected):
            if key.split('.')[0] != 'style':
                raise ValueError('key:%s is part of main network' % key)

    def make_noise(self, batch_size=1, device=None):                                          # gm.py:683-696
        device = self.input.input.device if device is None else device
        noises = [torch.randn(batch_size, 1, 4, 4, device=device)]
        for i in range(3, self.log_size + 1):
            noises += [torch.randn(batch_size, 1, 2 ** i, 2 ** i, device=device) for _ in range(2)]
        return noises

    def mean_latent(self, n_latent):                                                          # gm.py:698-704
        latent_in = torch.randn(n_latent, self.style_dim, device=self.input.input.device)
        return self.style(latent_in).mean(0, keepdim=True)

    def get_latent(self, input):
        return self.map_styles(input)

    def _mapping_groups(self):
        """[(lo, hi, [EqualLinear...])] of the mapping network in latent order (vanilla: one group)."""
        if isinstance(self.style, MultiFcStack):
            cfg = self.style.fc_config
            return [(cfg.groups[n]['latent_place'][0], cfg.groups[n]['latent_place'][1],
                     [m for m in getattr(self.style, n) if isinstance(m, EqualLinear)]) for n in cfg.in_order_group_names]
        return [(0, self.style_dim, [m for m in self.style if isinstance(m, EqualLinear)])]

    def map_styles(self, z):
        """z -> w: the whole mapping network is ONE persistent kernel (and one more for its backward, ops._MappingFn)"""
        if z.is_cuda and z.ndim == 2 and z.dtype == torch.float32:
            tbl = _cached_fc_table(self, self._mapping_groups, z.device)
            if not torch.is_grad_enabled():
                return ops.mapping_forward(tbl, z.contiguous(), normalize=True)
            return ops.mapping_apply(tbl, z.contiguous(), normalize=True)
        return self.style(z)

    def forward(self, styles, return_latents=False, inject_index=None, truncation=1, truncation_latent=None,
                input_is_latent=False, noise=None, randomize_noise=True, return_grad=False):   # gm.py:709-801
        if not input_is_latent:
            styles = [self.map_styles(s) for s in styles]
        if noise is None:
            if randomize_noise:
                noise = [None] * self.num_layers
            else:
                noise = [getattr(self.noises, f'noise_{i}') for i in range(self.num_layers)]
        if truncation < 1:
            styles = [truncation_latent + truncation * (s - truncation_latent) for s in styles]
        if len(styles) < 2:
            inject_index = self.n_latent
            latent = styles[0] if styles[0].ndim == 3 else styles[0].unsqueeze(1).repeat(1, inject_index, 1)
        else:
            if inject_index is None:
                inject_index = random.randint(1, self.n_latent - 1)
            latent = torch.cat([styles[0].unsqueeze(1).repeat(1, inject_index, 1),
                                styles[1].unsqueeze(1).repeat(1, self.n_latent - inject_index, 1)], 1)
        image = self.synthesis(latent, noise)
        if return_grad:
            return image, self.g_path_regularize_grad(image, latent)
        if return_latents:
            return image, latent
        return image, None

    def _modulated_layers(self):
        """[(ModulatedConv2d, latent index)] in the order `synthesis` calls them (gm.py:779-793)"""
        items = [(self.conv1.conv, 0), (self.to_rgb1.conv, 1)]
        i = 1
        for conv1, conv2, to_rgb in zip(self.convs[::2], self.convs[1::2], self.to_rgbs):
            items += [(conv1.conv, i), (conv2.conv, i + 1), (to_rgb.conv, i + 2)]
            i += 2
        return items

    def _precompute_modulations(self, latent):
        """All style modulations `s = EqualLinear(w_l)` (gm.py:245, 284: one per StyledConv / ToRGB, 26 at 1024^2) in ONE
        launch of the mapping kernel (batch-of-linear-layers mode) instead of one skinny GEMM per layer -- and one launch for
        their backward instead of ~5 per layer.  First-order / no-grad passes on CUDA; each layer picks its slice up in
        `ModulatedConv2d._style`."""
        if not (ops.fused_prep() and latent.is_cuda and latent.dtype == torch.float32 and latent.ndim == 3
                and latent.shape[1] == self.n_latent):
            return
        layers = self._modulated_layers()
        items = [(m.modulation, k * self.style_dim) for m, k in layers]
        tbl = _cached_linear_batch(self, items, self.n_latent * self.style_dim, latent.device)
        flat = latent.reshape(latent.shape[0], -1).contiguous()
        s_all = ops.mapping_apply(tbl, flat, normalize=False) if torch.is_grad_enabled() \
            else ops.mapping_forward(tbl, flat, normalize=False)
        for (m, _), s in zip(layers, s_all.split(tbl['widths'], dim=1)):          # split: ONE cat in the backward pass
            m._s_pre = s

    def synthesis(self, latent, noise):
        self._precompute_modulations(latent)
        try:
            return self._synthesis(latent, noise)
        finally:
            for m, _ in self._modulated_layers():
                m._s_pre = None

    def _synthesis(self, latent, noise):
        out = self.input(latent).to(dtype=self.act_dtype, memory_format=torch.channels_last)
        out = self.conv1(out, latent[:, 0], noise=self._noise(noise[0], out))
        skip = self.to_rgb1(out, latent[:, 1])
        i = 1
        for conv1, conv2, noise1, noise2, to_rgb in zip(self.convs[::2], self.convs[1::2], noise[1::2],
                                                        noise[2::2], self.to_rgbs):
            out = conv1(out, latent[:, i], noise=self._noise(noise1, out))
            out = conv2(out, latent[:, i + 1], noise=self._noise(noise2, out))
            skip = to_rgb(out, latent[:, i + 2], skip)
            i += 2
        return skip

    @staticmethod
    def _noise(noise, like):
        """stored noise buffers are (1,1,H,W): broadcast over the batch like `image + w*noise` does"""
        if noise is not None and noise.shape[0] != like.shape[0]:
            noise = noise.expand(like.shape[0], -1, -1, -1)
        return noise

    @staticmethod
    def g_path_regularize_grad(fake_img, latents, dim_1_shape=1):                             # gm.py:803-811
        noise = torch.randn_like(fake_img) / math.sqrt(fake_img.shape[2] * fake_img.shape[3] * dim_1_shape)
        with ops.data_grads_only():
            grad, = torch.autograd.grad(outputs=(fake_img * noise).sum(), inputs=latents, create_graph=True)
        return grad


class ConvLayer(nn.Sequential):                                               # gm.py:844-890
    def __init__(self, in_channel, out_channel, kernel_size, downsample=False, blur_kernel=[1, 3, 3, 1],
                 bias=True, activate=True):
        layers = []
        if downsample:
            factor = 2
            p = (len(blur_kernel) - factor) + (kernel_size - 1)
            layers.append(Blur(blur_kernel, pad=((p + 1) // 2, p // 2)))
            stride, self.padding = 2, 0
        else:
            stride, self.padding = 1, kernel_size // 2
        layers.append(EqualConv2d(in_channel, out_channel, kernel_size, padding=self.padding, stride=stride,
                                  bias=bias and not activate))
        if activate:
            layers.append(FusedLeakyReLU(out_channel) if bias else ScaledLeakyReLU(0.2))
        super().__init__(*layers)
        # Blur -> 3x3 stride-2 conv as ONE 3x3 convolution on the space-to-depth view of the input (ops.composite_down):
        # only where the layer is HBM-bound (4x the MMA work, no blurred (H+1)^2 intermediate)
        self.fuse_down = (downsample and kernel_size == 3 and len(blur_kernel) == 4 and activate
                          and in_channel <= self.FUSE_DOWN_MAX_IN_CHANNELS and in_channel % 4 == 0)
        if self.fuse_down:
            self.register_buffer('fir_toeplitz', ops.fir_toeplitz(layers[0].kernel, True), persistent=False)

    FUSE_DOWN_MAX_IN_CHANNELS = 32

    def forward(self, input, out_scale=1.0, defer_gate=False):
        """[Blur] -> conv (+ bias + leaky-ReLU fused into the convolution's epilogue).  `out_scale`
        multiplies the result (ResBlock folds its 1/sqrt(2) into both branches this way).  defer_gate: see ops.mod_conv
        (only honoured on the fused first-order path; the caller must then pass `in_gate` to every consumer)."""
        mods = list(self)
        x = input
        stride = None
        if self.fuse_down and input.shape[2] % 2 == 0 and input.shape[3] % 2 == 0:
            blur, conv, act = mods
            w = ops.composite_down((conv.weight * conv.scale).unsqueeze(0), self.fir_toeplitz)
            bias = act.bias if isinstance(act, FusedLeakyReLU) else None
            gain = act.scale if isinstance(act, FusedLeakyReLU) else SQRT2
            y = ops.conv_epilogue(x, w, None, None, None, bias, 1, 1, 1, slope=act.negative_slope, gain=gain * out_scale,
                                  pack_in=True, param_weight=True)
            return _plain_if_tiny(y)
        if isinstance(mods[0], Blur):
            blur, conv = mods[0], mods[1]
            if conv.weight.shape[2] == 1 and conv.stride == 2 and conv.padding == 0:
                # Blur -> 1x1 stride-2 conv (ResBlock.skip, gm.py:907-909) only ever reads the blur at even
                # positions: decimate inside the FIR (upfirdn2d down=2, a quarter of the outputs) and run the
                # 1x1 conv at stride 1 on the small map.  Same arithmetic, same taps, same result.
                x = ops.upfirdn2d(x, blur.kernel, up=1, down=2, pad=blur.pad)
                stride = 1
            else:
                x = blur(x)
            mods = mods[1:]
        conv = mods[0]
        stride = conv.stride if stride is None else stride
        if ops.fused_prep():
            # weight scaling + layout, convolution, bias + leaky-ReLU and all gradients as single kernels (ops.mod_conv)
            if len(mods) == 1:
                bias = None if conv.bias is None else conv.bias * out_scale
                return ops.mod_conv(x, None, conv.weight, bias=bias, scale=conv.scale * out_scale, down=stride, pad0=conv.padding)
            act = mods[1]
            bias = act.bias if isinstance(act, FusedLeakyReLU) else None
            gain = act.scale if isinstance(act, FusedLeakyReLU) else SQRT2
            return _plain_if_tiny(ops.mod_conv(x, None, conv.weight, bias=bias, scale=conv.scale, down=stride, pad0=conv.padding,
                                               slope=act.negative_slope, gain=gain * out_scale, defer_gate=defer_gate))
        if len(mods) == 1:                                   # no activation (ResBlock.skip): scale the weights
            w = (conv.weight * (conv.scale * out_scale)).unsqueeze(0)
            y = ops.conv_gather(x, w, 1, stride, conv.padding, param_weight=True)
            return y if conv.bias is None else y + (conv.bias * out_scale).view(1, -1, 1, 1).to(y.dtype)
        w = (conv.weight * conv.scale).unsqueeze(0)
        act = mods[1]
        bias = act.bias if isinstance(act, FusedLeakyReLU) else None
        gain = act.scale if isinstance(act, FusedLeakyReLU) else SQRT2
        y = ops.conv_epilogue(x, w, None, None, None, bias, 1, stride, conv.padding,
                              slope=act.negative_slope, gain=gain * out_scale, param_weight=True)
        return _plain_if_tiny(y)


def _plain_if_tiny(y):
    """The reference Discriminator `.view()`s its last (4x4) feature maps (gm.py:1006,1014), which needs the
    default NCHW-contiguous layout; at <= 16 pixels the copy is free."""
    if y.shape[2] * y.shape[3] <= 16 and not y.is_contiguous():
        return y.contiguous()
    return y


class ResBlock(nn.Module):                                                    # gm.py:893-922
    def __init__(self, in_channel, out_channel, blur_kernel=[1, 3, 3, 1], overwrite_padding=None):
        super().__init__()
        if overwrite_padding is not None:
            raise NotImplementedError('"896" mode is not built')
        self.conv1 = ConvLayer(in_channel, in_channel, 3)
        self.conv2 = ConvLayer(in_channel, out_channel, 3, downsample=True)
        self.skip = ConvLayer(in_channel, out_channel, 1, downsample=True, activate=False, bias=False)

    def fusable(self, input):
        """the whole block as one autograd node (ops.res_block): first-order passes, standard 3x3 / 4-tap geometry"""
        c1, c2, sk = self.conv1, self.conv2, self.skip
        return (ops.fused_prep() and input.shape[2] % 2 == 0 and input.shape[3] % 2 == 0 and input.shape[2] >= 4
                and len(c1) == 2 and isinstance(c1[1], FusedLeakyReLU) and len(c2) == 3 and isinstance(c2[2], FusedLeakyReLU)
                and len(sk) == 2 and sk[1].bias is None and c2[0].kernel.shape == (4, 4)
                and c1[1].negative_slope == c2[2].negative_slope and c1[1].scale == c2[2].scale)

    def forward(self, input, in_gate=None):
        """in_gate=(slope, gain): `input` is the output of a layer that deferred its activation backward to us
        (ops.mod_conv `defer_gate`); only legal when the block is `fusable`."""
        if self.fusable(input):
            c1, c2, sk = self.conv1, self.conv2, self.skip
            fuse_down = c2.fuse_down
            return _plain_if_tiny(ops.res_block(
                input, c1[0].weight, c1[1].bias, c2[1].weight, c2[2].bias, sk[1].weight, c2[0].kernel,
                c2.fir_toeplitz if fuse_down else c2[0].kernel, c1[0].scale, c2[1].scale, sk[1].scale,
                slope=c1[1].negative_slope, gain=c1[1].scale, fuse_down=fuse_down, pad_blur2=c2[0].pad[0], pad_blur_s=sk[0].pad[0],
                in_gate=in_gate))
        assert in_gate is None, 'a deferred activation gate reached a ResBlock that cannot apply it'
        # (conv2(conv1(x)) + skip(x)) / sqrt(2) (gm.py:920) with the scale folded into both branches:
        # one elementwise pass instead of two
        out = self.conv2(self.conv1(input), out_scale=1 / SQRT2)
        return _plain_if_tiny(out + self.skip(input, out_scale=1 / SQRT2))


class Discriminator(nn.Module):                                               # gm.py:925-1016
    def __init__(self, size, channel_multiplier=2, blur_kernel=[1, 3, 3, 1], in_channels=3, verification=False,
                 verification_res_split=None, model_mode=None, act_dtype=torch.float32):
        super().__init__()
        if verification or model_mode == '896':
            raise NotImplementedError('the verification head / "896" mode are never enabled by the trainer '
                                      '(generator_trainer.py:146-150) and are not built')
        self.model_mode = model_mode
        self.verification = verification
        self.act_dtype = act_dtype
        channels = _channels(channel_multiplier)
        convs = [ConvLayer(in_channels, channels[size], 1)]
        log_size = int(math.log(size, 2))
        in_channel = channels[size]
        for i in range(log_size, 2, -1):
            out_channel = channels[2 ** (i - 1)]
            convs.append(ResBlock(in_channel, out_channel, blur_kernel))
            in_channel = out_channel
        self.convs = nn.Sequential(*convs)
        self.convs_adv = nn.Sequential()
        self.convs_verification = nn.Sequential()
        self.stddev_group = 4
        self.stddev_feat = 1
        self.final_conv = ConvLayer(in_channel + 1, channels[4], 3)
        self.final_linear = nn.Sequential(EqualLinear(channels[4] * 4 * 4, channels[4], activation='fused_lrelu'),
                                          EqualLinear(channels[4], 1))

    def forward(self, input, stddev_chunks=1):
        """stddev_chunks > 1: `input` is that many independent batches concatenated (the discriminator step's fake and
        real batch, gt.py:655-656).  Every layer is per-sample except the minibatch-stddev statistic, which is then taken
        within each chunk, so the result equals separate calls -- with half the launches and better-filled small layers."""
        x = input.to(dtype=self.act_dtype, memory_format=torch.channels_last)
        blocks = list(self.convs)
        if (ops.fused_prep() and torch.is_grad_enabled() and len(blocks) > 1 and isinstance(blocks[1], ResBlock)
                and isinstance(blocks[0], ConvLayer) and len(blocks[0]) == 2 and isinstance(blocks[0][1], FusedLeakyReLU)):
            # from_rgb defers its leaky-ReLU backward to the first ResBlock, whose conv1 data-gradient epilogue applies it
            # together with the sum of the two gradient contributions (ops._ResBlockFn)
            y0 = blocks[0](x, defer_gate=True)
            if blocks[1].fusable(y0):
                out = blocks[1](y0, in_gate=(blocks[0][1].negative_slope, blocks[0][1].scale))
            else:                                                    # (cannot happen for the standard sizes; stay correct)
                out = blocks[1](blocks[0](x))
            for blk in blocks[2:]:
                out = blk(out)
        else:
            out = self.convs(x)
        return self._forward_split(out, self.final_conv, self.final_linear, stddev_chunks), None

    def _minibatch_stddev(self, out):                                         # gm.py:1004-1012
        batch, channel, height, width = out.shape
        group = min(batch, self.stddev_group)
        stddev = ops.up32(out).reshape(group, -1, self.stddev_feat, channel // self.stddev_feat, height, width)
        stddev = torch.sqrt(stddev.var(0, unbiased=False) + 1e-8)
        stddev = stddev.mean([2, 3, 4], keepdim=True).squeeze(2)
        return stddev.repeat(group, 1, height, width).to(out.dtype)

    def _forward_split(self, out, final_conv, final_linear, stddev_chunks=1):     # gm.py:1003-1016
        batch = out.shape[0]
        if stddev_chunks > 1:
            assert batch % stddev_chunks == 0
            stddev = torch.cat([self._minibatch_stddev(o) for o in out.chunk(stddev_chunks)])
        else:
            stddev = self._minibatch_stddev(out)
        out = self._final_conv_split(out, stddev, final_conv)
        out = out.reshape(batch, -1)
        return final_linear(out)

    @staticmethod
    def _final_conv_split(out, stddev, final_conv):
        """final_conv(cat([out, stddev], 1)) (gm.py:1013-1014) without the concatenation: a convolution is a sum over
        input channels, so the 512 feature channels go through the tensor-core engine (513 channels are not a
        multiple of its 8-channel vectors) and the single stddev channel is added as its own 9-tap convolution;
        bias + leaky-ReLU run on the sum.  The parameter keeps the reference shape (512, 513, 3, 3)."""
        mods = list(final_conv)
        conv, act = mods[0], mods[1]
        channel = out.shape[1]
        w = conv.weight * conv.scale
        z = ops.conv_gather(out, w[:, :channel].unsqueeze(0), 1, conv.stride, conv.padding, param_weight=True)
        z = z + ops.conv_gather(stddev, w[:, channel:].unsqueeze(0), 1, conv.stride, conv.padding, param_weight=True)
        y = ops.fused_leaky_relu(z, act.bias, act.negative_slope, act.scale)
        return _plain_if_tiny(y)


class FcStack(nn.Module):                                    # models/controller_model.py:13-52
    def __init__(self, lr_mlp, n_mlp, in_dim, mid_dim, out_dim):
        super().__init__()
        self.lr_mlp, self.n_mlp, self.in_dim, self.mid_dim, self.out_dim = lr_mlp, n_mlp, in_dim, mid_dim, out_dim
        layers = []
        for i in range(n_mlp):
            d0 = in_dim if i == 0 else mid_dim
            d1 = out_dim if i == n_mlp - 1 else mid_dim
            layers.append(EqualLinear(d0, d1, lr_mul=lr_mlp, activation='fused_lrelu'))
        self.fc_stack = nn.Sequential(*layers)

    def forward(self, x):
        if x.is_cuda and x.ndim == 2 and x.dtype == torch.float32:
            tbl = _cached_fc_table(self, lambda: [(0, self.in_dim, list(self.fc_stack))], x.device)
            if not torch.is_grad_enabled():
                return ops.mapping_forward(tbl, x.contiguous(), normalize=False)
            return ops.mapping_apply(tbl, x.contiguous(), normalize=False)
        return self.fc_stack(x)
