"""The G+D training step of gan-control on the B200 path.

Host-side mirror of ``GeneratorTrainer``'s step functions for the vanilla (adversarial + R1 +
path-length) objective, with the same names and arithmetic:

    discriminator_step            generator_trainer.py:645-667   (d_logistic_loss :690-695)
    discriminator_regularize_step generator_trainer.py:697-719   (d_r1_loss :713-719)
    generator_step                generator_trainer.py:407-436   (g_nonsaturating_loss :563-566)
    generator_regularize_step     generator_trainer.py:568-624   (g_path_regularize :601-614)
    optimiser setup               generator_trainer.py:158-173   (lazy-regularisation Adam)
    accumulate (EMA)              trainers/utils.py:8-12, decay generator_trainer.py:332

What is B200-native here:
  * one process per GPU; the reference's single-process ``nn.DataParallel`` (gt.py:195-199:
    replicate + scatter + gather every forward) is replaced by replicas that stay bit-identical
    and exchange ONLY gradients: each network's parameters and gradients live in two flat fp32
    arenas, gradients are all-reduced (NCCL over NVLink/NVSwitch) straight out of the arena in
    buckets as backward produces them, no flatten/unflatten copies;
  * Adam + the g_ema accumulation are one fused kernel over the arena (b200gan_adam_ema);
  * no per-step ``.item()`` host syncs (gt.py:429,659,705): losses stay on the device.
"""
import math

import torch
import torch.distributed as dist
import torch.nn.functional as F

from . import kernels as K
from .ops import data_grads_only, first_order, up32


# ---------------------------------------------------------------------------------------------
# losses (gt.py:563-566, 690-695, 713-719, 601-624)
# ---------------------------------------------------------------------------------------------
def g_nonsaturating_loss(fake_pred):
    return F.softplus(-fake_pred).mean()


def d_logistic_loss(real_pred, fake_pred):
    return F.softplus(-real_pred).mean() + F.softplus(fake_pred).mean()


def d_r1_loss(real_pred, real_img):
    with data_grads_only():
        grad_real, = torch.autograd.grad(outputs=real_pred.sum(), inputs=real_img, create_graph=True)
    return up32(grad_real).pow(2).reshape(grad_real.shape[0], -1).sum(1).mean()


def g_path_regularize(fake_img, latents, mean_path_length, decay=0.01, pl_noise=None, all_reduce_mean=None):
    if pl_noise is None:
        pl_noise = torch.randn_like(fake_img)
    pl_noise = pl_noise / math.sqrt(fake_img.shape[2] * fake_img.shape[3])
    with data_grads_only():
        grad, = torch.autograd.grad(outputs=(fake_img * pl_noise).sum(), inputs=latents, create_graph=True)
    path_lengths = torch.sqrt(grad.pow(2).sum(2).mean(1))
    local_mean = path_lengths.mean()
    if all_reduce_mean is None:
        path_mean = mean_path_length + decay * (local_mean - mean_path_length)
        path_penalty = (path_lengths - path_mean).pow(2).mean()
    else:
        # The reference computes the running mean AND the penalty on the gathered full batch
        # (gt.py:579,621-623); path_mean stays attached to the graph there, so every sample's length
        # also receives d(penalty)/d(path_mean) * decay / B.  With the batch sharded over replicas
        # that cross-sample term is restored exactly from the all-reduced mean:
        global_mean = all_reduce_mean(local_mean.detach())
        path_mean = mean_path_length + decay * (global_mean - mean_path_length)          # a constant here
        dpen_dmean = -2.0 * (global_mean - path_mean)
        path_penalty = (path_lengths - path_mean).pow(2).mean() + dpen_dmean * decay * (local_mean - local_mean.detach())
    return path_penalty, path_mean.detach(), path_lengths


# ---------------------------------------------------------------------------------------------
# flat parameter / gradient arenas
# ---------------------------------------------------------------------------------------------
class ParamArena:
    """All parameters of a module as views into ONE fp32 buffer, gradients into another.

    ``tail`` names parameters placed last so that a regularisation step can leave them out of the
    optimiser (the reference's ``set_grad_none(..., none_*_grads)``, gt.py:594,708)."""

    ALIGN = 32    # elements; keeps every view 128-byte aligned

    def __init__(self, module, tail=()):
        named = [(n, p) for n, p in module.named_parameters()]
        head = [(n, p) for n, p in named if n not in tail]
        last = [(n, p) for n, p in named if n in tail]
        self.names, self.params, self.offsets = [], [], []
        off = 0
        for group in (head, last):
            if group is last:
                self.split = off
            for n, p in group:
                self.names.append(n)
                self.params.append(p)
                self.offsets.append(off)
                off += -(-p.numel() // self.ALIGN) * self.ALIGN
        if not last:
            self.split = off
        self.numel = off
        dev, dtype = named[0][1].device, named[0][1].dtype     # fp32 (fp64 only in the CPU algebra tests)
        self.data = torch.zeros(off, dtype=dtype, device=dev)
        self.grad = torch.zeros(off, dtype=dtype, device=dev)
        for p, o in zip(self.params, self.offsets):
            assert p.dtype == dtype
            view = self.data[o:o + p.numel()].view(p.shape)
            view.copy_(p.data)
            p.data = view
            p.grad = self.grad[o:o + p.numel()].view(p.shape)

    def zero_grad(self):
        self.grad.zero_()

    def check_views(self):
        """backward must accumulate in place; a replaced .grad would silently drop out of the arena"""
        for p, o in zip(self.params, self.offsets):
            assert p.grad is not None and p.grad.data_ptr() == self.grad.data_ptr() + self.grad.element_size() * o, \
                'a parameter gradient left the flat arena'
            assert p.data_ptr() == self.data.data_ptr() + self.data.element_size() * o, 'a parameter left the flat arena'


class ArenaAdam:
    """torch.optim.Adam semantics (eps 1e-8, no weight decay) on an arena, fused with EMA.

    The step counters live on the device (the bias corrections 1 - beta^t are recomputed there), so
    the optimiser step is safe to capture in a CUDA graph and replay."""

    def __init__(self, arena, lr, betas, ema_arena=None):
        self.arena, self.lr, self.betas, self.ema = arena, lr, betas, ema_arena
        self.m = torch.zeros_like(arena.data)
        self.v = torch.zeros_like(arena.data)
        dev = arena.data.device
        # [head, tail] ranges step separately (see ParamArena.tail)
        self.t = torch.zeros(2, dtype=torch.float64, device=dev)
        self.log_betas = torch.tensor([math.log(b) if b > 0 else -math.inf for b in betas], dtype=torch.float64, device=dev)

    def _bias_corr(self, which):
        self.t[which] += 1
        # 1 - beta^t  (beta = 0 -> exp(-inf * t) = 0 -> correction 1, like 0 ** t for t >= 1)
        return (1.0 - torch.exp(self.log_betas * self.t[which])).to(self.arena.data.dtype)

    def step(self, skip_tail=False, ema_decay=None, grad_scale=1.0, buckets=None):
        """One optimiser step.  With `buckets` (a GradBuckets whose all-reduces are in flight) the arena is stepped
        bucket by bucket, each as soon as ITS all-reduce has completed, so the update of the early buckets overlaps the
        reduction of the late ones instead of waiting for all of them (the gradients that backward produces last --
        the large 4x4..64x64 weights -- no longer gate a monolithic step)."""
        a = self.arena
        corr = {0: self._bias_corr(0) if a.split > 0 else None}
        tail = not skip_tail and a.split < a.numel
        if tail:
            corr[1] = self._bias_corr(1)
        if buckets is not None and buckets.enabled:
            ranges = [(lo, hi, 0 if hi <= a.split else 1, work) for lo, hi, work in buckets.drain()]
        else:
            ranges = [(0, a.split, 0, None), (a.split, a.numel, 1, None)]
        for lo, hi, which, work in ranges:
            if work is not None:
                work.wait()                            # this stream waits for this bucket's all-reduce only
            if hi <= lo or (which == 1 and not tail):
                continue
            ema = None
            if self.ema is not None and ema_decay is not None:
                ema = self.ema.data[lo:hi]
            K.adam_ema(a.data[lo:hi], a.grad[lo:hi], self.m[lo:hi], self.v[lo:hi], ema, self.lr, self.betas[0],
                       self.betas[1], 1e-8, corr[which], ema_decay if ema is not None else 0.0, grad_scale)

    # -- torch.optim.Adam wire format (the reference checkpoints `g_optim` / `d_optim`, gt.py:852-865, 192-193) ----
    def _arena_index(self, module):
        """arena slot of every parameter, in `module.parameters()` order (the order `optim.Adam(model.parameters())`
        numbers them, gt.py:161-173)"""
        slot = {id(p): j for j, p in enumerate(self.arena.params)}
        return [slot[id(p)] for p in module.parameters()]

    def state_dict(self, module):
        """`torch.optim.Adam(module.parameters(), lr, betas).state_dict()` of the equivalent optimiser: per-parameter
        `step` / `exp_avg` / `exp_avg_sq` (a parameter of the tail range has stepped only in the non-regularised
        iterations, exactly like a parameter whose gradient `set_grad_none` dropped, gt.py:594,708)."""
        a = self.arena
        order = self._arena_index(module)
        probe = torch.optim.Adam([torch.zeros(1)], lr=self.lr, betas=self.betas)       # the installed torch's group keys
        group = dict(probe.state_dict()['param_groups'][0], params=list(range(len(order))))
        t = self.t.tolist()
        state = {}
        for i, j in enumerate(order):
            lo, n = a.offsets[j], a.params[j].numel()
            steps = t[1] if lo >= a.split else t[0]
            if steps <= 0:
                continue                                   # torch creates the state at a parameter's first step
            shape = a.params[j].shape
            state[i] = {'step': torch.tensor(float(steps)),
                        'exp_avg': self.m[lo:lo + n].view(shape).detach().clone(),
                        'exp_avg_sq': self.v[lo:lo + n].view(shape).detach().clone()}
        return {'state': state, 'param_groups': [group]}

    def load_state_dict(self, module, sd):
        a = self.arena
        order = self._arena_index(module)
        group = sd['param_groups'][0]
        assert len(group['params']) == len(order), 'optimizer state is for a different parameter list'
        self.lr, self.betas = group['lr'], tuple(group['betas'])
        self.log_betas.copy_(torch.tensor([math.log(b) if b > 0 else -math.inf for b in self.betas], dtype=torch.float64))
        self.m.zero_()
        self.v.zero_()
        steps = [0.0, 0.0]
        for i, j in enumerate(order):
            st = sd['state'].get(group['params'][i])
            if st is None:
                continue
            lo, n = a.offsets[j], a.params[j].numel()
            self.m[lo:lo + n].copy_(st['exp_avg'].reshape(-1))
            self.v[lo:lo + n].copy_(st['exp_avg_sq'].reshape(-1))
            which = 1 if lo >= a.split else 0
            steps[which] = max(steps[which], float(st['step']))
        self.t.copy_(torch.tensor(steps, dtype=torch.float64))


class GradBuckets:
    """Bucketed gradient all-reduce out of a ParamArena, launched as backward fills buckets.

    Buckets are contiguous arena ranges, numbered from the END of the arena (backward produces gradients roughly in
    reverse registration order) and never straddle `arena.split` (the optimiser steps the two sides separately).
    Collectives are issued strictly in bucket-index order on every rank -- bucket b is launched only once buckets
    0..b-1 have been -- so ranks always pair collectives of the same range even if their autograd graphs complete
    parameters in different orders (e.g. style mixing drawing a different number of latents per rank)."""

    def __init__(self, arena, world_size, bucket_mb=8, group=None):
        self.arena, self.world, self.group = arena, world_size, group
        self.enabled = world_size > 1
        self.pending = []
        if not self.enabled:
            return
        cap = max(1, int(bucket_mb * (1 << 20)) // arena.data.element_size())
        self.bucket_of, self.bucket_range, self.bucket_size = {}, [], []
        hi, members, idx = arena.numel, 0, 0
        n = len(arena.params)
        for i in reversed(range(n)):
            lo_edge = arena.offsets[i]
            self.bucket_of[i] = idx
            members += 1
            at_split = lo_edge == arena.split and i > 0          # close the bucket at the head / tail boundary
            if hi - lo_edge >= cap or i == 0 or at_split:
                self.bucket_range.append((lo_edge, hi))
                self.bucket_size.append(members)
                hi, members, idx = lo_edge, 0, idx + 1
        for i, p in enumerate(arena.params):
            p.register_post_accumulate_grad_hook(self._make_hook(i))
        self.active = False
        self.begin()
        self.active = False

    def _make_hook(self, i):
        def hook(param):
            if not self.active:
                return
            b = self.bucket_of[i]
            self.count[b] += 1
            if self.count[b] == self.bucket_size[b]:
                self.ready[b] = True
                self._launch_ready()
        return hook

    def _launch_ready(self):
        while self.next < len(self.bucket_range) and self.ready[self.next]:
            lo, hi = self.bucket_range[self.next]
            work = dist.all_reduce(self.arena.grad[lo:hi], op=dist.ReduceOp.SUM, group=self.group, async_op=True)
            self.pending.append((lo, hi, work))
            self.next += 1

    def begin(self):
        """arm the hooks for one backward pass"""
        if not self.enabled:
            return
        self.active = True
        nb = len(self.bucket_range)
        self.count, self.ready, self.next, self.pending = [0] * nb, [False] * nb, 0, []

    def finish(self):
        """after backward: launch (still in index order) whatever did not fill -- parameters without a gradient this
        pass.  Gradients are SUMS over ranks (the optimiser divides by world).  The collectives stay in flight: the
        optimiser consumes them bucket by bucket through `drain`."""
        if not self.enabled:
            return
        self.active = False
        self.ready = [True] * len(self.bucket_range)
        self._launch_ready()

    def drain(self):
        """(lo, hi, work) of this pass in launch order; the caller waits on each `work` before touching its range"""
        out, self.pending = self.pending, []
        return out

    def wait_all(self):
        for _, _, w in self.drain():
            w.wait()


# ---------------------------------------------------------------------------------------------
def unbuilt_objective_terms(config):
    """What a gan-control config asks for that this step does NOT do.  The reference applies, when `model_config.vanilla`
    is false, every enabled `training_config.*_loss` (embedding / orientation / expression / age / hair / ..., gt.py:205-
    300, 431-434) on batches arranged by `MiniBatchUtils.re_arrange_z`; independently it applies `d_every` (gt.py:351)
    and a transfer-learning initialisation (gt.py:134-143).  This path builds the adversarial + R1 + path-length
    objective, with ADA augmentation (`augment.enabled`, gt.py:421-422, 647-653: `gan_control_b200.augment`) when the
    config asks for it (SURVEY.md section 8 scope), so a config that relies on any of the above would silently train
    something else."""
    mc, tc = config['model_config'], config['training_config']
    out = []
    if not mc.get('vanilla', False):
        out += [f'{k} (attribute loss)' for k, v in tc.items() if k.endswith('_loss') and isinstance(v, dict)
                and v.get('enabled')]
    if tc.get('d_every', 1) != 1:
        out.append(f"d_every={tc['d_every']}")
    if (tc.get('transfer_learning_model') or {}).get('enabled'):
        out.append('transfer_learning_model')
    return out


class GanTrainStep:
    """One process = one GPU = one replica.  ``batch`` is the per-replica batch."""

    def __init__(self, generator, discriminator, g_ema=None, batch=16, lr_g=0.002, lr_d=0.002, r1=1.0,
                 d_reg_every=16, g_reg_every=4, path_regularize=2.0, path_batch_shrink=2, mixing=0.0,
                 g_moving_average=10000, latent_size=512, world_size=1, bucket_mb=8, global_batch=None, ada=None):
        self.g, self.d, self.g_ema = generator, discriminator, g_ema
        self.batch, self.world = batch, world_size
        self.global_batch = global_batch or batch * world_size
        self.r1, self.d_reg_every, self.g_reg_every = r1, d_reg_every, g_reg_every
        self.path_regularize, self.path_batch_shrink, self.mixing = path_regularize, path_batch_shrink, mixing
        self.latent_size = latent_size
        # (a discriminator class without the `stddev_chunks` argument -- e.g. the reference's own -- gets the two separate calls)
        import inspect
        self._d_takes_chunks = 'stddev_chunks' in inspect.signature(discriminator.forward).parameters
        # ADA (gt.py:333-337): None, a fixed probability, or an `augment.AdaptiveP` controller
        if ada is not None and not hasattr(ada, 'update'):
            from .augment import AdaptiveP
            ada = AdaptiveP(p=float(ada))
        self.ada = ada
        self.device = next(generator.parameters()).device
        self.mean_path_length = torch.zeros((), device=self.device, dtype=next(generator.parameters()).dtype)
        self.accum = 0.5 ** (self.global_batch / g_moving_average)                       # gt.py:332
        # parameters that get no gradient from the regularisers (gt.py:301-327 dry_run finds them):
        g_tail = [n for n, _ in generator.named_parameters() if n.startswith('to_rgb') and n.endswith('.bias')
                  and 'modulation' not in n]
        d_tail = ['final_linear.1.bias']
        self.g_arena = ParamArena(generator, tail=g_tail)
        self.d_arena = ParamArena(discriminator, tail=d_tail)
        self.ema_arena = ParamArena(g_ema, tail=g_tail) if g_ema is not None else None
        if self.ema_arena is not None:
            self.ema_arena.data.copy_(self.g_arena.data)                                 # accumulate(g_ema, g, 0)
        g_ratio = g_reg_every / (g_reg_every + 1)                                        # gt.py:158-173
        d_ratio = d_reg_every / (d_reg_every + 1)
        self.g_optim = ArenaAdam(self.g_arena, lr_g * g_ratio, (0 ** g_ratio, 0.99 ** g_ratio), self.ema_arena)
        self.d_optim = ArenaAdam(self.d_arena, lr_d * d_ratio, (0 ** d_ratio, 0.99 ** d_ratio))
        self.g_buckets = GradBuckets(self.g_arena, world_size, bucket_mb)
        self.d_buckets = GradBuckets(self.d_arena, world_size, bucket_mb)
        self.stats = {}
        self.graphs = None

    # -- construction from the reference's config files (configs/*.json, `args.json` of a run) -----------------------
    @classmethod
    def from_config(cls, config, device='cuda', world_size=1, act_dtype=torch.bfloat16, vanilla_only=False, **overrides):
        """Networks and step exactly as `GeneratorTrainer.init_models_and_optim` builds them (gt.py:120-173) from the
        `model_config` / `training_config` sections of a gan-control config (a dict, or the path of `configs/ffhq.json`
        or of a run's `args.json`): Generator / g_ema / Discriminator constructor arguments, split-FC layout from
        `sub_groups_dict`, lazy-regularisation Adam, R1 / path-length weights and cadence, EMA horizon, style mixing.
        `training_config['batch']` is the GLOBAL batch (the reference scatters it over its DataParallel replicas); each
        of the `world_size` processes gets batch / world_size.  Resumes from `ckpt_config` when enabled (gt.py:175-193).
        A config that enables objective terms this path does not build (`unbuilt_objective_terms`) raises unless
        `vanilla_only=True` says the adversarial + R1 + path-length objective alone is what is wanted; the effective
        objective is recorded in `step.effective_objective` (and in the run's args.json by `train`)."""
        import json
        import warnings
        from . import modules as M
        if not isinstance(config, dict):
            with open(config) as f:
                config = json.load(f)
        mc, tc = config['model_config'], config['training_config']
        dropped = unbuilt_objective_terms(config)
        if dropped and not vanilla_only:
            raise NotImplementedError(
                'this config enables training terms the B200 path does not build: ' + ', '.join(dropped) + '. It would '
                'train the vanilla objective (adversarial + R1 + path-length) only; pass vanilla_only=True '
                '(`--vanilla-only`) to do that deliberately.')
        if dropped:
            warnings.warn('gan_control_b200: training the vanilla objective only; IGNORED config terms: ' + ', '.join(dropped))
        if mc.get('marge_fc') or mc.get('vae'):
            raise NotImplementedError('marge_fc / vae are not used by any shipped config and are not built')
        if tc.get('mini_batch', tc['batch']) != tc['batch']:
            raise NotImplementedError('gradient accumulation over mini-batches (mini_batch < batch) is not built: one '
                                      'process per GPU holds batch / world_size samples')
        groups = None if mc.get('vanilla') or not mc.get('split_fc') else tc['sub_groups_dict']
        fc_config = M.FcConfig.from_sub_groups_dict(groups) if groups else None

        def generator():
            return M.Generator(mc['size'], mc['latent_size'], mc['n_mlp'], channel_multiplier=mc['channel_multiplier'],
                               out_channels=mc['img_channels'], split_fc=bool(groups), fc_config=fc_config,
                               conv_transpose=mc['conv_transpose'], noise_mode=mc.get('g_noise_mode', 'normal'),
                               act_dtype=act_dtype).to(device)
        g, g_ema = generator(), generator()
        g_ema.eval()
        d = M.Discriminator(mc['size'], channel_multiplier=mc['channel_multiplier'], in_channels=mc['img_channels'],
                            act_dtype=act_dtype).to(device)
        if tc['batch'] % world_size:
            raise ValueError('batch %d is not divisible by world size %d' % (tc['batch'], world_size))
        kw = dict(batch=tc['batch'] // world_size, lr_g=tc['lr_g'], lr_d=tc['lr_d'], r1=tc['r1'], d_reg_every=tc['d_reg_every'],
                  g_reg_every=tc['g_reg_every'], path_regularize=tc['path_regularize'],
                  path_batch_shrink=tc['path_batch_shrink'], mixing=tc['mixing'], g_moving_average=tc['g_moving_average'],
                  latent_size=mc['latent_size'], world_size=world_size, global_batch=tc['batch'])
        aug = tc.get('augment') or {}
        if aug.get('enabled'):
            from .augment import AdaptiveP
            kw['ada'] = AdaptiveP(p=aug.get('p', 0), ada_target=aug.get('ada_target', 0.6), ada_length=aug.get('ada_length', 500000))
        kw.update(overrides)
        step = cls(g, d, g_ema, **kw)
        step.config = config
        step.effective_objective = {'terms': ['adversarial (non-saturating / logistic)', 'r1', 'path_length'] +
                                             (['ada_augment'] if step.ada is not None else []),
                                    'ignored_config_terms': dropped}
        ck = config.get('ckpt_config') or {}
        if ck.get('enabled'):
            step.load_checkpoint(ck['ckpt'])
        return step

    # -- helpers ------------------------------------------------------------------------------
    @staticmethod
    def requires_grad(model, flag=True):                                                 # tu:13-15
        for p in model.parameters():
            p.requires_grad_(flag)

    def mixing_noise(self, batch):                                                       # tu:19-23
        import random
        if self.mixing > 0 and random.random() < self.mixing:
            return list(torch.randn(2, batch, self.latent_size, device=self.device).unbind(0))
        return [torch.randn(batch, self.latent_size, device=self.device)]

    def _styles(self, batch, noise=None):
        """What the generator is called with: (styles, kwargs).  Explicit `noise` (a list of z, tests) is passed through.
        Otherwise style mixing (tu:19-23 + gm.py:762-769) is drawn ON THE DEVICE: both latents are always mapped, the
        mixing coin and the crossover index are device scalars and every layer selects its latent with a mask -- the same
        distribution as the reference's host-side `random.random() < mixing` / `random.randint(1, n_latent - 1)`, but
        with a fixed launch sequence, so the step can be captured in a CUDA graph and every rank issues the same
        kernels / collectives whatever it draws."""
        if noise is not None:
            return noise, {}
        dtype = self.mean_path_length.dtype                    # fp32 (fp64 only in the CPU algebra tests)
        if self.mixing <= 0:
            return [torch.randn(batch, self.latent_size, device=self.device, dtype=dtype)], {}
        z = torch.randn(2, batch, self.latent_size, device=self.device, dtype=dtype)
        w1, w2 = self.g.map_styles(z[0]), self.g.map_styles(z[1])
        n = self.g.n_latent
        mix = torch.rand((), device=self.device) < self.mixing
        index = torch.randint(1, n, (), device=self.device)                           # uniform on 1 .. n_latent - 1
        first = (torch.arange(n, device=self.device) < index) | ~mix
        latent = torch.where(first.view(1, n, 1), w1.unsqueeze(1), w2.unsqueeze(1))
        return [latent], {'input_is_latent': True}

    def _mean_over_ranks(self, t):
        if self.world > 1:
            t = t.clone()
            dist.all_reduce(t)
            t /= self.world
        return t

    # -- discriminator ------------------------------------------------------------------------
    def discriminator_step(self, real_img, noise=None):
        self.requires_grad(self.g, False)
        self.requires_grad(self.d, True)
        self.d_arena.zero_grad()
        with torch.no_grad():
            styles, kw = self._styles(self.batch, noise)
            fake_img, _ = self.g(styles, **kw)
            if self.ada is not None:                 # gt.py:651-653: both batches, independent draws
                from .augment import augment
                real_img, _ = augment(real_img, self.ada.p)
                fake_img, _ = augment(fake_img, self.ada.p)
        with first_order():                          # plain step: no double backward -> fused single-kernel layers
            if fake_img.shape == real_img.shape and self._d_takes_chunks:
                # D(fake), D(real) (gt.py:655-656) as ONE pass over the concatenated batch: identical values (the
                # minibatch-stddev statistic stays per batch), half the launches, one weight-gradient reduction
                both = torch.cat([fake_img, real_img.to(dtype=fake_img.dtype, memory_format=torch.channels_last)])
                pred, _ = self.d(both, stddev_chunks=2)
                fake_pred, real_pred = pred.chunk(2)
            else:
                fake_pred, _ = self.d(fake_img)
                real_pred, _ = self.d(real_img)
            d_loss = d_logistic_loss(real_pred, fake_pred)
            d_loss = d_loss / self.global_batch      # gt.py:656 `d_loss.div_(len(mini_real_img))`, kept
            self.d_buckets.begin()
            d_loss.backward()
            self.d_buckets.finish()
        self.d_optim.step(grad_scale=1.0 / self.world, buckets=self.d_buckets)
        if self.ada is not None:
            self.ada.update(real_pred)               # r_t statistic and the adaptive probability (gt.py:669-687; host sync)
        self.stats['d_loss'] = d_loss.detach()
        return d_loss.detach()

    def discriminator_regularize_step(self, real_img):
        self.requires_grad(self.g, False)
        self.requires_grad(self.d, True)
        self.d_arena.zero_grad()
        real_img = real_img.detach().requires_grad_(True)
        real_pred, _ = self.d(real_img)
        r1_loss = d_r1_loss(real_pred, real_img)
        self.d_buckets.begin()
        (self.r1 / 2 * r1_loss * self.d_reg_every + 0 * real_pred[0]).sum().backward()   # gt.py:706
        self.d_buckets.finish()
        self.d_optim.step(skip_tail=True, grad_scale=1.0 / self.world, buckets=self.d_buckets)                    # set_grad_none gt.py:708
        self.stats['r1_loss'] = r1_loss.detach()
        return r1_loss.detach()

    # -- generator ----------------------------------------------------------------------------
    def generator_step(self, noise=None, ema=True):
        self.requires_grad(self.g, True)
        self.requires_grad(self.d, False)
        self.g_arena.zero_grad()
        with first_order():
            styles, kw = self._styles(self.batch, noise)
            fake_img, _ = self.g(styles, **kw)
            if self.ada is not None:                 # gt.py:421-422: the generator is trained through the augmentation
                from .augment import augment
                fake_img, _ = augment(fake_img, self.ada.p)
            fake_pred, _ = self.d(fake_img)
            g_loss = g_nonsaturating_loss(fake_pred)
            self.g_buckets.begin()
            g_loss.backward()
            self.g_buckets.finish()
        self.g_optim.step(ema_decay=self.accum if ema else None, grad_scale=1.0 / self.world, buckets=self.g_buckets)
        self.stats['g_loss'] = g_loss.detach()
        return g_loss.detach()

    def generator_regularize_step(self, noise=None, pl_noise=None):
        self.requires_grad(self.g, True)
        self.requires_grad(self.d, False)
        self.g_arena.zero_grad()
        path_batch = max(1, self.batch // self.path_batch_shrink)
        styles, kw = self._styles(path_batch, noise)
        fake_img, latents = self.g(styles, return_latents=True, **kw)
        path_loss, new_mean, path_lengths = g_path_regularize(
            fake_img, latents, self.mean_path_length, pl_noise=pl_noise,
            all_reduce_mean=self._mean_over_ranks if self.world > 1 else None)
        self.mean_path_length.copy_(new_mean)          # in place: the tensor is static under CUDA graphs
        weighted = self.path_regularize * self.g_reg_every * path_loss                     # gt.py:587-590
        if self.path_batch_shrink:
            weighted = weighted + 0 * up32(fake_img[0, 0, 0, 0])
        self.g_buckets.begin()
        weighted.backward()
        self.g_buckets.finish()
        self.g_optim.step(skip_tail=True, grad_scale=1.0 / self.world, buckets=self.g_buckets)                    # set_grad_none gt.py:594
        self.stats['path_loss'] = path_loss.detach()
        return path_loss.detach()

    # -- checkpoints in the reference's wire format (gt.py:852-865 save_nets, :175-193 resume) ---------------------
    def checkpoint(self):
        """{'g', 'd', 'g_ema', 'g_optim', 'd_optim'} as `GeneratorTrainer.save_nets` writes them: module state_dicts
        (same keys / shapes as the reference's, tests/test_host_algebra_cpu.py) and `torch.optim.Adam` state_dicts, so
        the file resumes either implementation.  The path-length running mean is not part of the reference's
        checkpoint (it restarts from 0 there as well, gt.py:330); it is stored under an extra key."""
        cpu = lambda sd: {k: v.detach().cpu().clone() for k, v in sd.items()}
        out = {'g': cpu(self.g.state_dict()), 'd': cpu(self.d.state_dict()),
               'g_optim': self.g_optim.state_dict(self.g), 'd_optim': self.d_optim.state_dict(self.d)}
        if self.g_ema is not None:
            out['g_ema'] = cpu(self.g_ema.state_dict())
        for opt in ('g_optim', 'd_optim'):
            for st in out[opt]['state'].values():
                st['exp_avg'], st['exp_avg_sq'] = st['exp_avg'].cpu(), st['exp_avg_sq'].cpu()
        out['b200gan_mean_path_length'] = self.mean_path_length.detach().cpu().clone()
        return out

    def save_nets(self, i, save_dir, best_fid=False):
        """`<save_dir>/checkpoint/<iteration, 6 digits>.pt` (gt.py:852-854)"""
        import os
        os.makedirs(os.path.join(save_dir, 'checkpoint'), exist_ok=True)
        path = os.path.join(save_dir, 'checkpoint', 'best_fid.pt' if best_fid else f'{str(i).zfill(6)}.pt')
        torch.save(self.checkpoint(), path)
        return path

    def load_checkpoint(self, ckpt):
        """resume from a reference-format checkpoint (a dict or a path), gt.py:179-193.  Parameters are copied INTO the
        arena views (the modules keep pointing at the flat buffers; captured CUDA graphs stay valid)."""
        if not isinstance(ckpt, dict):
            ckpt = torch.load(ckpt, map_location='cpu')
        self.g.load_state_dict(ckpt['g'])
        self.d.load_state_dict(ckpt['d'])
        if self.g_ema is not None and 'g_ema' in ckpt:
            self.g_ema.load_state_dict(ckpt['g_ema'])
        self.g_arena.check_views()
        self.d_arena.check_views()
        if 'g_optim' in ckpt:
            self.g_optim.load_state_dict(self.g, ckpt['g_optim'])
        if 'd_optim' in ckpt:
            self.d_optim.load_state_dict(self.d, ckpt['d_optim'])
        if 'b200gan_mean_path_length' in ckpt:
            self.mean_path_length.copy_(ckpt['b200gan_mean_path_length'])
        return self

    # -- one iteration (gt.py:343-353: discriminator_update, generator_update) ------------------------
    def train_step(self, i, real_img, regularize=True):
        d_loss = self.discriminator_step(real_img)
        if regularize and i % self.d_reg_every == 0:
            self.discriminator_regularize_step(real_img)
        do_g_reg = regularize and i % self.g_reg_every == 0
        g_loss = self.generator_step(ema=not do_g_reg)
        if do_g_reg:
            self.generator_regularize_step()
            # accumulate() runs once per iteration after both G updates (gt.py:362-369)
            if self.ema_arena is not None:
                self.ema_arena.data.mul_(self.accum).add_(self.g_arena.data, alpha=1 - self.accum)
        return d_loss, g_loss


    # -- CUDA graphs: the whole step is launch-bound on the host (thousands of small kernels), so the
    #    four step variants are captured once and replayed --------------------------------------------
    def capture(self, real_shape, warmup=2):
        """Capture discriminator_step / generator_step (+ their regularised variants) into CUDA graphs
        reading from a static image buffer.  Style mixing is drawn on the device (`_styles`), so it is captured too."""
        if self.ada is not None:
            raise RuntimeError('ADA augmentation draws its transforms (and the padding they need) on the host every step: '
                               'run `train_step` eagerly, CUDA-graph capture is for the un-augmented step')
        self.static_real = torch.zeros(real_shape, device=self.device)
        snapshot = self._snapshot()          # the warm-up runs are real optimiser steps on a dummy batch: undone below
        side = torch.cuda.Stream()
        side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side):
            for _ in range(warmup):
                self._variant('d')
                self._variant('d_reg')
                self._variant('g')
                self._variant('g_reg')
        torch.cuda.current_stream().wait_stream(side)
        torch.cuda.synchronize()
        self.graphs, self.graph_launches, self.replayed_launches, self.graph_out = {}, {}, 0, {}
        pool = None
        for name in ('d', 'd_reg', 'g', 'g_reg'):
            graph = torch.cuda.CUDAGraph()
            n0 = K.launch_count()
            with torch.cuda.graph(graph, pool=pool):
                self._variant(name)
            self.graph_launches[name] = K.launch_count() - n0      # libb200gan kernels inside this graph
            pool = graph.pool()
            self.graphs[name] = graph
            # this variant's loss tensors: the references keep their memory out of the later captures that share the pool
            # (a dropped one is reused there, and `stats` alone only remembers the LAST variant captured)
            self.graph_out[name] = dict(self.stats)
        self._restore(snapshot)
        torch.cuda.synchronize()
        return self

    def _state_tensors(self):
        """everything a step mutates that outlives it: parameters, Adam moments and step counts, EMA, path-length mean"""
        ts = [self.g_arena.data, self.d_arena.data, self.g_optim.m, self.g_optim.v, self.g_optim.t,
              self.d_optim.m, self.d_optim.v, self.d_optim.t, self.mean_path_length]
        if self.ema_arena is not None:
            ts.append(self.ema_arena.data)
        return ts

    def _snapshot(self):
        return [t.detach().clone() for t in self._state_tensors()]

    def _restore(self, snapshot):
        """in place: the tensors are what the modules' parameters view and what captured graphs point at"""
        with torch.no_grad():
            for t, s in zip(self._state_tensors(), snapshot):
                t.copy_(s)

    def _replay(self, name):
        self.graphs[name].replay()
        self.replayed_launches += self.graph_launches[name]

    def _variant(self, name):
        if name == 'd':
            self.discriminator_step(self.static_real)
        elif name == 'd_reg':
            self.discriminator_regularize_step(self.static_real)
        elif name == 'g':
            self.generator_step(ema=True)
        else:   # generator step without EMA, path-length step, then the iteration's single EMA update
            self.generator_step(ema=False)
            self.generator_regularize_step()
            if self.ema_arena is not None:
                self.ema_arena.data.mul_(self.accum).add_(self.g_arena.data, alpha=1 - self.accum)

    def train_step_graphed(self, i, real_img, regularize=True):
        self.static_real.copy_(real_img, non_blocking=True)
        self._replay('d')
        if regularize and i % self.d_reg_every == 0:
            self._replay('d_reg')
        g_name = 'g_reg' if regularize and i % self.g_reg_every == 0 else 'g'
        self._replay(g_name)
        return self.graph_out['d']['d_loss'], self.graph_out[g_name]['g_loss']
