"""The reference's G+D training iteration restated on the oracle (TEST / BASELINE INFRASTRUCTURE ONLY).

`generator_trainer.py:343-369` (one `discriminator_update` + one `generator_update`) for the vanilla objective, with
the reference's own gradient accumulation over mini-batches (`mini_batch < batch`, gt.py:361,411-436,645-667):

    discriminator_step            gt.py:645-667   fake = G(z) without graph; d_logistic_loss; backward per mini-batch
    discriminator_regularize_step gt.py:697-711   every d_reg_every:  r1/2 * R1 * d_reg_every
    generator_step                gt.py:407-436   g_nonsaturating_loss; backward per mini-batch
    generator_regularize_step     gt.py:568-599   every g_reg_every at batch // path_batch_shrink
    optimisers                    gt.py:158-173   lazy-regularisation Adam;  EMA  trainers/utils.py:8-12

It evaluates the networks with `oracle/stylegan2_oracle.py` (plain PyTorch ops = the reference's FUSED=False path),
on whatever device / dtype the state dicts live on:
  * CPU fp32, all host threads      -> `bench.py --impl reference` and `cpu_baseline` (kind "port")
  * CUDA fp32 / bf16 autocast       -> `bench.py`'s `gpu_library_baseline`: the same arithmetic through cuDNN / cuBLAS,
                                       the stand-in for the unavailable FUSED=True comparator.
Only bench.py and tests/ import this module; the product never does.
"""
import contextlib

import torch

from . import params as P
from . import stylegan2_oracle as O


class ReferenceStep:
    def __init__(self, size, batch, mini_batch=None, device='cpu', dtype=torch.float32, autocast=None, seed=1,
                 style_dim=512, n_mlp=8, channel_multiplier=2, r1=1.0, d_reg_every=16, g_reg_every=4,
                 path_regularize=2.0, path_batch_shrink=2, lr=0.002, g_moving_average=10000):
        self.size, self.batch, self.mini_batch = size, batch, mini_batch or batch
        assert batch % self.mini_batch == 0
        self.device, self.dtype, self.autocast = torch.device(device), dtype, autocast
        self.style_dim = style_dim
        self.r1, self.d_reg_every, self.g_reg_every = r1, d_reg_every, g_reg_every
        self.path_regularize, self.path_batch_shrink = path_regularize, path_batch_shrink

        def net(shapes, s):
            return {k: v.to(self.device, dtype).requires_grad_(not k.endswith('kernel') and not k.startswith('noises.'))
                    for k, v in P.seeded_state_dict(shapes, s).items()}
        self.sd_g = net(P.generator_shapes(size, style_dim, n_mlp, channel_multiplier), seed)
        self.sd_d = net(P.discriminator_shapes(size, channel_multiplier), seed + 1)
        self.gp = {k: v for k, v in self.sd_g.items() if v.requires_grad}
        self.dp = {k: v for k, v in self.sd_d.items() if v.requires_grad}
        self.ema = {k: v.detach().clone() for k, v in self.gp.items()}
        lr_g, b_g = O.lazy_adam_hparams(lr, g_reg_every)
        lr_d, b_d = O.lazy_adam_hparams(lr, d_reg_every)
        self.g_opt = torch.optim.Adam(list(self.gp.values()), lr=lr_g, betas=b_g)
        self.d_opt = torch.optim.Adam(list(self.dp.values()), lr=lr_d, betas=b_d)
        self.accum = 0.5 ** (batch / g_moving_average)
        self.mean_path_length = torch.zeros((), device=self.device, dtype=dtype)
        self.losses = {}

    def _ctx(self):
        if self.autocast is None:
            return contextlib.nullcontext()
        return torch.autocast(self.device.type, dtype=self.autocast)

    def _z(self, n):
        return torch.randn(n, self.style_dim, device=self.device, dtype=self.dtype)

    def _g(self, z, **kw):
        with self._ctx():
            return O.generator_forward(self.sd_g, [z], self.size, **kw)

    def _d(self, x):
        with self._ctx():
            return O.discriminator_forward(self.sd_d, x, self.size).float()

    def discriminator_step(self, real):
        self.d_opt.zero_grad(set_to_none=True)
        for lo in range(0, self.batch, self.mini_batch):
            mb = real[lo:lo + self.mini_batch]
            with torch.no_grad():
                fake = self._g(self._z(len(mb)))
            d_loss = O.d_logistic_loss(self._d(mb), self._d(fake)) / len(mb)          # gt.py:656 `div_(len(mini_real_img))`
            d_loss.backward(inputs=list(self.dp.values()))
        self.d_opt.step()
        self.losses['d'] = d_loss.detach()

    def discriminator_regularize_step(self, real):
        self.d_opt.zero_grad(set_to_none=True)
        for lo in range(0, self.batch, self.mini_batch):
            x = real[lo:lo + self.mini_batch].detach().clone().requires_grad_(True)
            pred = self._d(x)
            r1 = O.d_r1_loss(pred, x)
            r1 = r1 / (self.batch // self.mini_batch)                      # gt.py:704 `div_(len(mini_real_inputs))`
            (self.r1 / 2 * r1 * self.d_reg_every + 0 * pred[0]).sum().backward(inputs=list(self.dp.values()))
        self.dp['final_linear.1.bias'].grad = None                       # set_grad_none, gt.py:708
        self.d_opt.step()
        self.losses['r1'] = r1.detach()

    def generator_step(self):
        self.g_opt.zero_grad(set_to_none=True)
        for lo in range(0, self.batch, self.mini_batch):
            n = min(self.mini_batch, self.batch - lo)
            fake = self._g(self._z(n))
            g_loss = O.g_nonsaturating_loss(self._d(fake)) / (self.batch // self.mini_batch)     # gt.py:428
            g_loss.backward(inputs=list(self.gp.values()))
        self.g_opt.step()
        self.losses['g'] = g_loss.detach()

    def generator_regularize_step(self):
        self.g_opt.zero_grad(set_to_none=True)
        path_batch = max(1, self.batch // self.path_batch_shrink)
        mini = min(self.mini_batch, path_batch)
        for lo in range(0, path_batch, mini):
            n = min(mini, path_batch - lo)
            fake, lat = self._g(self._z(n), return_latents=True)
            pen, self.mean_path_length, _ = O.g_path_regularize(fake.float(), lat, self.mean_path_length)
            pen = pen / -(-path_batch // mini)                             # gt.py:586 `div_(len(mini_noise_inputs))`
            (self.path_regularize * self.g_reg_every * pen + 0 * fake[0, 0, 0, 0]).backward(inputs=list(self.gp.values()))
        for k in self.gp:
            if k.startswith('to_rgb') and k.endswith('.bias') and 'modulation' not in k:
                self.gp[k].grad = None                                   # set_grad_none, gt.py:594
        self.g_opt.step()
        self.losses['path'] = pen.detach()

    def train_step(self, i, real, regularize=True):
        """iteration i of gt.py:343-353 on the batch `real` (batch, 3, size, size)"""
        self.discriminator_step(real)
        if regularize and i % self.d_reg_every == 0:
            self.discriminator_regularize_step(real)
        self.generator_step()
        if regularize and i % self.g_reg_every == 0:
            self.generator_regularize_step()
        O.ema_accumulate(self.ema, self.sd_g, self.accum)
        return self.losses['d'], self.losses['g']
