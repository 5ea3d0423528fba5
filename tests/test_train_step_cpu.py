"""The training step (gan_control_b200/train_step.py) against a restatement of the reference's step
functions built from the oracle + torch.optim.Adam, on CPU with the kernel stand-ins: single
replica, and two gloo replicas vs the same global batch on one."""
import copy
import os

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from gan_control_b200 import modules as M
from gan_control_b200.train_step import GanTrainStep
from oracle import params as P
from oracle import stylegan2_oracle as O

SIZE, SDIM, NMLP = 8, 32, 2


F64 = torch.float64


def rnd(seed, *shape):
    return torch.from_numpy(np.random.default_rng(seed).standard_normal(shape))      # fp64: Adam(beta1=0) amplifies rounding noise


class FixedNoiseG(M.Generator):
    """deterministic per-layer noise so that two implementations see the same draw"""
    def forward(self, styles, **kw):
        if kw.get('noise') is None:
            b = styles[0].shape[0]
            kw['noise'] = [self.fixed_noise[i][:b] for i in range(self.num_layers)]
        return super().forward(styles, **kw)


def build(batch):
    g = FixedNoiseG(SIZE, SDIM, NMLP, channel_multiplier=2, conv_transpose=True, act_dtype=F64).double()
    g.load_state_dict(P.seeded_state_dict(P.generator_shapes(SIZE, SDIM, NMLP, 2), 5, dtype=F64))
    g.fixed_noise = [rnd(50 + i, batch, 1, 2 ** ((i + 5) // 2), 2 ** ((i + 5) // 2)) for i in range(g.num_layers)]
    g_ema = copy.deepcopy(g)
    d = M.Discriminator(SIZE, channel_multiplier=2, act_dtype=F64).double()
    d.load_state_dict(P.seeded_state_dict(P.discriminator_shapes(SIZE, 2), 6, dtype=F64))
    return g, g_ema, d


def oracle_run(batch, real, zs, pl_noise, iters, return_optim=False):
    """generator_trainer.py:343-369,407-436,568-599,645-711 restated on the oracle."""
    sd_g = {k: v.clone().requires_grad_(not k.endswith('kernel') and not k.startswith('noises.'))
            for k, v in P.seeded_state_dict(P.generator_shapes(SIZE, SDIM, NMLP, 2), 5, dtype=F64).items()}
    sd_d = {k: v.clone().requires_grad_(not k.endswith('kernel')) for k, v in P.seeded_state_dict(P.discriminator_shapes(SIZE, 2), 6, dtype=F64).items()}
    ema = {k: v.detach().clone() for k, v in sd_g.items() if v.requires_grad}
    noise = [rnd(50 + i, batch, 1, 2 ** ((i + 5) // 2), 2 ** ((i + 5) // 2)) for i in range(2 * (int(np.log2(SIZE)) - 2) + 1)]
    gp = {k: v for k, v in sd_g.items() if v.requires_grad}
    dp = {k: v for k, v in sd_d.items() if v.requires_grad}
    lr_g, b_g = O.lazy_adam_hparams(0.002, 4)
    lr_d, b_d = O.lazy_adam_hparams(0.002, 16)
    g_opt = torch.optim.Adam(list(gp.values()), lr=lr_g, betas=b_g)
    d_opt = torch.optim.Adam(list(dp.values()), lr=lr_d, betas=b_d)
    accum = 0.5 ** (batch / 10000)
    mean_pl = torch.zeros((), dtype=F64)
    for i in iters:
        z_d, z_g, z_pl = zs[i]
        with torch.no_grad():
            fake = O.generator_forward(sd_g, [z_d], SIZE, noise=noise)
        d_loss = O.d_logistic_loss(O.discriminator_forward(sd_d, real, SIZE), O.discriminator_forward(sd_d, fake, SIZE)) / batch
        d_opt.zero_grad(set_to_none=True)
        d_loss.backward(inputs=list(dp.values()))
        d_opt.step()
        if i % 16 == 0:
            x = real.clone().requires_grad_(True)
            pred = O.discriminator_forward(sd_d, x, SIZE)
            r1 = O.d_r1_loss(pred, x)
            d_opt.zero_grad(set_to_none=True)
            (1.0 / 2 * r1 * 16 + 0 * pred[0]).sum().backward(inputs=list(dp.values()))
            dp['final_linear.1.bias'].grad = None                       # set_grad_none, gt.py:708
            d_opt.step()
        fake = O.generator_forward(sd_g, [z_g], SIZE, noise=noise)
        g_loss = O.g_nonsaturating_loss(O.discriminator_forward(sd_d, fake, SIZE))
        g_opt.zero_grad(set_to_none=True)
        g_loss.backward(inputs=list(gp.values()))
        g_opt.step()
        if i % 4 == 0:
            pb = max(1, batch // 2)
            fake, lat = O.generator_forward(sd_g, [z_pl], SIZE, noise=[n[:pb] for n in noise], return_latents=True)
            pen, mean_pl, _ = O.g_path_regularize(fake, lat, mean_pl, pl_noise=pl_noise)
            g_opt.zero_grad(set_to_none=True)
            (2.0 * 4 * pen + 0 * fake[0, 0, 0, 0]).backward(inputs=list(gp.values()))
            for k in gp:
                if k.startswith('to_rgb') and k.endswith('.bias') and 'modulation' not in k:
                    gp[k].grad = None                                   # set_grad_none, gt.py:594
            g_opt.step()
        O.ema_accumulate(ema, sd_g, accum)
    if return_optim:
        return sd_g, sd_d, ema, (g_opt, gp), (d_opt, dp)
    return sd_g, sd_d, ema


def product_run(batch, real, zs, pl_noise, iters, world=1, rank=0, resume=None, return_step=False):
    g, g_ema, d = build(batch * world)
    if world > 1:   # replica r owns samples r::world of the global batch (strided like minibatch-stddev groups)
        g.fixed_noise = [n[rank::world] for n in g.fixed_noise]
        real = real[rank::world]
    step = GanTrainStep(g, d, g_ema, batch=batch, latent_size=SDIM, world_size=world, bucket_mb=1)
    if resume is not None:
        step.load_checkpoint(resume)
    for i in iters:
        z_d, z_g, z_pl = [z[rank::world] if world > 1 else z for z in zs[i]]
        step.discriminator_step(real, [z_d])
        if i % 16 == 0:
            step.discriminator_regularize_step(real)
        do_reg = i % 4 == 0
        step.generator_step([z_g], ema=not do_reg)
        if do_reg:
            pn = pl_noise[rank::world] if world > 1 else pl_noise
            step.generator_regularize_step([z_pl], pl_noise=pn)
            step.ema_arena.data.mul_(step.accum).add_(step.g_arena.data, alpha=1 - step.accum)
        step.g_arena.check_views()
        step.d_arena.check_views()
    if return_step:
        return g, d, g_ema, step
    return g, d, g_ema


def make_inputs(batch):
    real = rnd(70, batch, 3, SIZE, SIZE).clamp_(-1, 1)
    zs = {i: (rnd(80 + i, batch, SDIM), rnd(90 + i, batch, SDIM), rnd(100 + i, max(1, batch // 2), SDIM)) for i in range(0, 3)}
    pl_noise = rnd(110, max(1, batch // 2), 3, SIZE, SIZE)
    return real, zs, pl_noise


def compare(module, sd_ref, tol):
    worst = 0.0
    for k, p in module.named_parameters():
        ref = sd_ref[k].detach()
        err = float((p.detach() - ref).abs().max() / ref.abs().max().clamp_min(1e-12))
        worst = max(worst, err)
        assert err < tol, (k, err)
    return worst


def test_single_replica_matches_reference_step(cpu_kernels):
    batch = 4
    real, zs, pl_noise = make_inputs(batch)
    iters = [0, 1]          # iteration 0 runs both regularisers, 1 is a plain step
    sd_g, sd_d, ema = oracle_run(batch, real, zs, pl_noise, iters)
    g, d, g_ema = product_run(batch, real, zs, pl_noise, iters)
    compare(g, sd_g, 1e-7)
    compare(d, sd_d, 1e-7)
    compare(g_ema, ema, 1e-7)


def _worker(rank, world, port, out):
    os.environ.update(MASTER_ADDR='127.0.0.1', MASTER_PORT=str(port))
    dist.init_process_group('gloo', rank=rank, world_size=world)
    torch.set_num_threads(2)
    from gan_control_b200 import kernels
    from oracle import kernels_ref
    for name in kernels_ref.STAND_INS:
        setattr(kernels, name, getattr(kernels_ref, name))
    real, zs, pl_noise = make_inputs(4 * world)
    g, d, g_ema = product_run(4, real, zs, pl_noise, [0, 1], world=world, rank=rank)
    if rank == 0:
        torch.save({'g': g.state_dict(), 'd': d.state_dict(), 'ema': g_ema.state_dict()}, out)
    dist.barrier()
    dist.destroy_process_group()


def test_two_replicas_match_one(cpu_kernels, tmp_path):
    """world_size 2 over gloo: gradient all-reduce out of the flat arenas (+ the path-length mean
    exchange) reproduces the single-replica run on the same global batch of 8."""
    world, batch = 2, 4
    out = str(tmp_path / 'ddp.pt')
    port = 29500 + os.getpid() % 2000
    mp.spawn(_worker, args=(world, port, out), nprocs=world, join=True)
    ddp = torch.load(out)
    real, zs, pl_noise = make_inputs(batch * world)
    # single replica on the global batch, samples ordered so that the strided minibatch-stddev groups
    # (gm.py:1005-1011) coincide with the replicas' local groups
    g, d, g_ema = product_run(batch * world, real, zs, pl_noise, [0, 1])
    compare(g, ddp['g'], 1e-9)
    compare(d, ddp['d'], 1e-9)
    compare(g_ema, ddp['ema'], 1e-9)


def test_checkpoint_wire_format(cpu_kernels, tmp_path):
    """`GanTrainStep.checkpoint()` is the reference's {g, d, g_ema, g_optim, d_optim} file (gt.py:852-865): the optimiser
    entries equal the state of the torch.optim.Adam that the restated reference step drove (per-parameter step counts
    included: parameters dropped by set_grad_none in the regularisation steps have stepped once less), torch.optim.Adam
    loads them, and resuming from the file continues exactly like the uninterrupted run."""
    batch = 4
    real, zs, pl_noise = make_inputs(batch)
    sd_g, sd_d, ema, (g_opt, gp), (d_opt, dp) = oracle_run(batch, real, zs, pl_noise, [0], return_optim=True)
    g, d, g_ema, step = product_run(batch, real, zs, pl_noise, [0], return_step=True)
    path = step.save_nets(0, str(tmp_path))
    assert path.endswith('checkpoint/000000.pt')
    ckpt = torch.load(path)
    assert set(ckpt) >= {'g', 'd', 'g_ema', 'g_optim', 'd_optim'}
    for module, key, (opt, params) in [(g, 'g_optim', (g_opt, gp)), (d, 'd_optim', (d_opt, dp))]:
        osd = ckpt[key]
        names = [n for n, _ in module.named_parameters()]
        assert osd['param_groups'][0]['params'] == list(range(len(names)))
        assert osd['param_groups'][0]['lr'] == opt.param_groups[0]['lr'] and tuple(osd['param_groups'][0]['betas']) == tuple(opt.param_groups[0]['betas'])
        for i, n in enumerate(names):
            ref = opt.state[params[n]]
            mine = osd['state'][i]
            assert float(mine['step']) == float(ref['step']), n
            for k in ('exp_avg', 'exp_avg_sq'):
                err = float((mine[k] - ref[k]).abs().max() / ref[k].abs().max().clamp_min(1e-30))
                assert err < 1e-7, (n, k, err)
        steps = {n: float(osd['state'][i]['step']) for i, n in enumerate(names)}
        assert max(steps.values()) == 2.0 and min(steps.values()) == 1.0           # iteration 0: plain + regularised step
        # a real torch.optim.Adam over the same parameter list accepts the entry
        torch.optim.Adam(list(module.parameters()), lr=1.0).load_state_dict(osd)
    # resume: iteration 1 from the file == iterations 0, 1 uninterrupted
    g2, d2, g_ema2 = product_run(batch, real, zs, pl_noise, [1], resume=path)
    g1, d1, g_ema1 = product_run(batch, real, zs, pl_noise, [0, 1])
    for a, b in [(g2, g1), (d2, d1), (g_ema2, g_ema1)]:
        compare(a, b.state_dict(), 1e-12)


def test_bucket_launch_order_is_fixed(monkeypatch):
    """collectives must be issued in bucket-index order whatever order backward fills the buckets in
    (torch DDP's rule), and no bucket may straddle the head / tail split of the arena."""
    import gan_control_b200.train_step as TS
    torch.manual_seed(0)
    net = torch.nn.Sequential(*[torch.nn.Linear(64, 64) for _ in range(6)])
    arena = TS.ParamArena(net, tail=['5.bias'])
    launched = []

    class Work:
        def wait(self):
            pass

    def fake_all_reduce(t, op=None, group=None, async_op=False):
        launched.append((t.data_ptr() - arena.grad.data_ptr()) // 4)
        return Work()
    monkeypatch.setattr(TS.dist, 'all_reduce', fake_all_reduce)
    buckets = TS.GradBuckets(arena, world_size=2, bucket_mb=64 * 64 * 4 / (1 << 20))
    assert len(buckets.bucket_range) >= 6
    assert all(not (lo < arena.split < hi) for lo, hi in buckets.bucket_range)
    order = [lo for lo, _ in buckets.bucket_range]
    for perm_seed in range(3):
        launched.clear()
        buckets.begin()
        idx = torch.randperm(len(arena.params), generator=torch.Generator().manual_seed(perm_seed)).tolist()
        for i in idx[:-2]:                          # two parameters never get a gradient this pass
            buckets._make_hook(i)(arena.params[i])
            assert launched == order[:len(launched)]
        buckets.finish()
        assert launched == order
        ranges = [(lo, hi) for lo, hi, _ in buckets.drain()]
        assert ranges == buckets.bucket_range


def _worker_mixing(rank, world, port, out):
    os.environ.update(MASTER_ADDR='127.0.0.1', MASTER_PORT=str(port))
    dist.init_process_group('gloo', rank=rank, world_size=world)
    torch.set_num_threads(2)
    from gan_control_b200 import kernels
    from oracle import kernels_ref
    for name in kernels_ref.STAND_INS:
        setattr(kernels, name, getattr(kernels_ref, name))
    g, g_ema, d = build(4)
    step = GanTrainStep(g, d, g_ema, batch=4, latent_size=SDIM, world_size=world, bucket_mb=0.05, mixing=0.9)
    real = rnd(70 + rank, 4, 3, SIZE, SIZE).clamp_(-1, 1)
    for i in range(2):
        # ranks deliberately disagree on the mixing coin: rank 0 maps one latent, rank 1 two (tu:19-23)
        zs = [rnd(200 + 10 * i + rank, 4, SDIM)] if rank == 0 else [rnd(300 + i, 4, SDIM), rnd(400 + i, 4, SDIM)]
        step.discriminator_step(real, zs)
        step.generator_step(zs)
        step.generator_regularize_step(zs[:1] if rank == 0 else [z[:2] for z in zs])
    torch.save({'g': g.state_dict(), 'd': d.state_dict()}, out + str(rank))
    dist.barrier()
    dist.destroy_process_group()


def test_two_replicas_with_style_mixing_stay_in_sync(cpu_kernels, tmp_path):
    """world_size 2, mixing > 0 with the ranks taking DIFFERENT mixing branches in the same step: the bucketed
    all-reduces still pair up (fixed launch order) and the replicas end bit-identical."""
    out = str(tmp_path / 'mix')
    port = 31500 + os.getpid() % 2000
    mp.spawn(_worker_mixing, args=(2, port, out), nprocs=2, join=True)
    a, b = torch.load(out + '0'), torch.load(out + '1')
    for net in ('g', 'd'):
        for k in a[net]:
            assert torch.equal(a[net][k], b[net][k]), (net, k)


@pytest.mark.parametrize('mixing,seed', [(1.0, 3), (1.0, 4), (0.5, 5), (0.5, 6), (0.5, 7)])
def test_device_side_style_mixing_equals_reference_mixing(cpu_kernels, mixing, seed):
    """`GanTrainStep._styles` draws the mixing coin and the crossover index as device scalars and selects per layer with a
    mask (so the step is CUDA-graph capturable): the image equals `Generator([z1, z2], inject_index=index)` resp.
    `Generator([z1])` (gm.py:754-769, tu:19-23) for the same draws."""
    g, g_ema, d = build(4)
    step = GanTrainStep(g, d, g_ema, batch=4, latent_size=SDIM, mixing=mixing)
    torch.manual_seed(seed)
    styles, kw = step._styles(4)
    torch.manual_seed(seed)
    z = torch.randn(2, 4, SDIM, dtype=F64)
    mix = bool(torch.rand(()) < mixing)
    index = int(torch.randint(1, g.n_latent, ()))
    assert kw == {'input_is_latent': True} and styles[0].shape == (4, g.n_latent, SDIM)
    img, _ = g(styles, **kw)
    ref, _ = g([z[0], z[1]], inject_index=index) if mix else g([z[0]])
    assert float((img - ref).abs().max()) < 1e-12
    assert 1 <= index <= g.n_latent - 1
