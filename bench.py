#!/usr/bin/env python
"""bench.py -- images/sec of the FFHQ-1024 G+D train step on 1/2/4/8 B200 (BASELINE.json metric).

    python bench.py --gpus N --steps K --warmup W            # this repo's CUDA path
    python bench.py --impl reference --steps K --warmup W    # the reference's CPU path (oracle port), really timed

A "step" is one `discriminator_update` + one `generator_update` (generator_trainer.py:351-353) at
BASELINE.json configs[1]: Generator(1024)/Discriminator(1024), channel_multiplier 2, per-GPU batch
16, bf16 activations / fp32 accumulate and parameters, synthetic images, random-init weights, with
the lazy regularisers at their real cadence (R1 every 16, path-length every 4 iterations).
Rank 0 prints ONE JSON line (see DESIGN.md "Measurement").
"""
import argparse
import gc
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = 'images/sec FFHQ-1024 G+D train step'
UNIT = 'images/s'
# algorithmic work per image, BASELINE.md section 2 (G_f, D_f GFLOP forward; min fused traffic MB forward, bf16)
G_D_GFLOP = {1024: (148.5, 153.3), 512: (119.3, 123.2), 256: (90.2, 93.1)}
G_D_MB = {1024: (598.0, 728.0), 512: (290.0, 354.0), 256: (138.0, 168.0)}
REF_BUDGET_S = 240.0          # wall-clock budget of the whole `--impl reference` run (K + W bounded samples)


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument('--gpus', type=int, default=1)
    ap.add_argument('--steps', type=int, default=16)
    ap.add_argument('--warmup', type=int, default=3)
    ap.add_argument('--impl', default='b200', choices=['b200', 'reference'])
    ap.add_argument('--size', type=int, default=1024)
    ap.add_argument('--batch', type=int, default=16, help='per-GPU batch')
    ap.add_argument('--dtype', default='bf16', choices=['bf16', 'fp32'])
    ap.add_argument('--no-reg', action='store_true', help='plain steps only (no R1 / path-length)')
    ap.add_argument('--cpu-sample-size', type=int, default=None, help='resolution of the CPU-baseline sample')
    ap.add_argument('--skip-cpu-baseline', action='store_true')
    ap.add_argument('--skip-roofline', action='store_true')
    ap.add_argument('--skip-library-baseline', action='store_true', help='no cuDNN/cuBLAS comparator (gpu_library_baseline)')
    ap.add_argument('--skip-extra', action='store_true', help='no extra_configs (BASELINE.json configs[2..4] per-GPU workloads)')
    ap.add_argument('--no-graph', action='store_true', help='launch every kernel from Python instead of replaying CUDA graphs')
    ap.add_argument('--config', type=int, default=1, choices=[1, 5],
                    help='1: the train-step metric (default).  5: BASELINE.json configs[4] inference half -- controller '
                         'FcStack sweep over the 1000 control vectors + gen_batch_by_controls at 512x512')
    ap.add_argument('--ncu-range', action='store_true',
                    help='bracket the timed region with cudaProfilerStart/Stop (for `ncu --profile-from-start off`)')
    return ap.parse_args()


# ---------------------------------------------------------------------------------------------
class ClockSampler(threading.Thread):
    """nvidia-smi clocks / throttle reasons DURING the timed region (B200_PROFILING.md)."""
    Q = ('clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,'
         'clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap')

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index, self.rows, self.stop_flag = index, [], threading.Event()

    def run(self):
        while not self.stop_flag.is_set():
            try:
                out = subprocess.run(['nvidia-smi', '-i', str(self.index), f'--query-gpu={self.Q}', '--format=csv,noheader,nounits'],
                                     capture_output=True, text=True, timeout=5).stdout.strip()
                if out:
                    self.rows.append([c.strip() for c in out.split(',')])
            except Exception:
                pass
            self.stop_flag.wait(0.2)

    def summary(self):
        self.stop_flag.set()
        self.join(timeout=3)
        sm = sorted(float(r[0]) for r in self.rows if r and r[0].replace('.', '').isdigit())
        if not sm:
            return {'sm_mhz': None, 'sm_max_mhz': None, 'reasons': [], 'samples': 0}
        names = ['hw_slowdown', 'hw_thermal_slowdown', 'sw_thermal_slowdown', 'sw_power_cap']
        reasons = [n for j, n in enumerate(names) if any(len(r) > 3 + j and r[3 + j].lower().startswith('active') for r in self.rows)]
        return {'sm_mhz': sm[len(sm) // 2], 'sm_max_mhz': float(self.rows[0][1]), 'reasons': reasons, 'samples': len(sm),
                'power_w_max': max(float(r[2]) for r in self.rows if r[2].replace('.', '').isdigit())}


def measured_peaks():
    p = os.path.join(ROOT, 'MEASURED_PEAKS.json')
    if os.path.exists(p):
        d = json.load(open(p))
        return d, 'measured'
    return {'hbm_gbs': 6650.0, 'bf16_tflops': 1590.0, 'bf16_tflops_sustained': 1400.0}, 'fallback'


def workload_config(size, reg, batch, world):
    """`config` of the JSON line: identical for this repo's arm and for the reference arm"""
    return {'workload': f'FFHQ-{size} G+D train step (BASELINE.json configs[1]): Generator({size})+Discriminator({size}) '
                        f'channel_multiplier 2, random init, lazy regularisers at real cadence '
                        f'(R1 /16, path-length /4)' + ('' if reg else ' DISABLED'),
            'global_batch': batch * world, 'per_gpu_batch': batch, 'parallelism': f'dp{world}',
            'l2': 'inputs larger than L2 (each step streams >10 GB of activations; no explicit flush)'}


def reg_counts(first_iter, n_steps, reg=True):
    its = range(first_iter, first_iter + n_steps)
    return {'r1_steps': sum(1 for i in its if reg and i % 16 == 0), 'path_length_steps': sum(1 for i in its if reg and i % 4 == 0)}


# ---------------------------------------------------------------------------------------------
# the reference's CPU path, really timed (oracle port; `oracle/reference_step.py`)
# ---------------------------------------------------------------------------------------------
def cpu_reference_steps(size, n_timed, n_warm, threads, budget_s, log=None):
    """`n_warm` + `n_timed` REAL training iterations of the reference arithmetic on the host cores: each is one
    discriminator_step + one generator_step (forward, backward, Adam, EMA; gt.py:343-369) on a mini-batch of ONE image
    -- the reference's own gradient-accumulation granularity (mini_batch < batch, gt.py:361,411-436) -- without the lazy
    regularisers (they would add ~15 %: leaving them out favours the reference).  If the first iteration shows that the
    run cannot finish inside `budget_s` at `size`, the sample drops to 256x256 (and says so): its per-image cost is
    LOWER than the benchmark resolution's, again in the reference's favour.  Returns (images/s, ms per iteration, text)."""
    import torch
    from oracle.reference_step import ReferenceStep
    torch.set_num_threads(threads)
    torch.manual_seed(0)
    t_start = time.perf_counter()

    def make(sz):
        return ReferenceStep(sz, 1, 1), torch.randn(1, 3, sz, sz).clamp_(-1, 1)
    rs, real = make(size)
    t0 = time.perf_counter()
    rs.train_step(1, real, regularize=False)
    first = time.perf_counter() - t0
    note = ''
    remaining = n_warm + n_timed - 1
    if size > 256 and first * max(1, remaining) > budget_s - (time.perf_counter() - t_start):
        note = (f' (one iteration at {size}x{size} took {first:.1f} s: {n_warm}+{n_timed} of them do not fit the '
                f'{budget_s:.0f} s budget, so the sample runs at 256x256 -- cheaper per image, in the reference\'s favour)')
        size = 256
        rs, real = make(size)
        rs.train_step(1, real, regularize=False)
    for _ in range(max(0, n_warm - 1)):
        rs.train_step(1, real, regularize=False)
    ts = []
    for k in range(n_timed):
        t0 = time.perf_counter()
        rs.train_step(1 + k, real, regularize=False)
        ts.append(time.perf_counter() - t0)
        if log:
            log(f'reference iteration {k}: {ts[-1]:.2f} s')
    t = sum(ts) / len(ts)
    sample = (f'{n_timed} timed + {n_warm} warm-up REAL training iterations (D step + G step: forward, backward, Adam, EMA; no '
              f'regulariser steps) of the oracle port of gan_model.py / generator_trainer.py, fp32, mini-batch 1 at '
              f'{size}x{size}, {t:.2f} s per iteration on {threads} threads' + note)
    return 1.0 / t, t * 1e3, sample, size


def run_reference(args):
    """--impl reference: rank 0 times the reference's CPU implementation; other ranks exit quietly."""
    rank = int(os.environ.get('RANK', '0'))
    if rank != 0:
        return
    threads = os.cpu_count() or 1
    size = args.cpu_sample_size or args.size
    v, ms, sample, used = cpu_reference_steps(size, max(1, args.steps), args.warmup, threads, REF_BUDGET_S)
    line = {'impl': 'reference', 'metric': METRIC, 'value': v, 'unit': UNIT, 'n_gpus': args.gpus, 'steps': args.steps,
            'warmup': args.warmup, 'ms_per_step': ms, 'higher_is_better': True, 'scaling': 'weak',
            'vs_baseline': None, 'dtype': 'f32', 'data': 'synthetic', 'measured': True,
            'config': workload_config(args.size, not args.no_reg, args.batch, args.gpus),
            'sample_resolution': used, 'sample_batch': 1,
            'cpu_baseline': {'value': v, 'unit': UNIT, 'cores': threads, 'kind': 'port', 'sample': sample},
            'e2e': {'value': v, 'unit': UNIT, 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0}, 'gpu_launches': 0}
    print(json.dumps(line), flush=True)


# ---------------------------------------------------------------------------------------------
# per-layer roofline of the convolution kernels, timed alone WITH their fused epilogue operands
# ---------------------------------------------------------------------------------------------
# dram__bytes_read.sum + dram__bytes_write.sum of ONE launch at batch 16, from the `ncu --set full` captures summarised in
# profiles/r01_ncu_kernels.md / profiles/r02_headline.md (None where the layer has not been captured)
NCU_TRAFFIC_BYTES = {'conv3x3 256->256 @128x128 batch 16': 135.4e6 + 85.79e6,
                     'conv3x3 128->128 @256x256 batch 16': 268.762e6 + 220.351e6,
                     'conv3x3 64->64 @512x512 batch 16': 537.110e6 + 487.299e6,
                     # round 2, the kernel as the bench times it (noise + bias + leaky-ReLU, demodulated weights): profiles/r02_headline.md
                     'conv3x3 32->32 @1024x1024 batch 16': 1.107661e9 + 1.032646e9}
NCU_TRAFFIC_BYTES_WGRAD = {'conv3x3 32->32 @1024x1024 batch 16': 2.284250e9 + 0.005649e9}


def _time_kernel(torch, fn, flush, reps=10, warm=3):
    for _ in range(warm):
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
    return sum(ts) / len(ts)


def conv_roofline(torch, K, dtype, size, batch, peaks):
    """Every 3x3 layer class of the resolution, timed alone with CUDA events on the launching stream (inputs >> L2, L2
    flushed between launches) as the StyledConv it is on the path: per-sample modulated + demodulated weights and the
    fused noise / bias / leaky-ReLU epilogue operands.  Forward and weight gradient.  Headline = the layer the
    metric is quoted on, 32 -> 32 at full resolution: HBM-bound (input read once + output written once) with its
    tensor-pipe fraction beside it.  Returns (headline, all layers)."""
    flush = torch.empty(256 << 20, dtype=torch.uint8, device='cuda')
    layers = [(size, 32), (size // 2, 64), (size // 4, 128), (size // 8, 256)]
    hbm, tf = peaks.get('hbm_gbs', 6650.0), peaks.get('bf16_tflops', 1590.0)
    out = []
    for res, ch in layers:
        x = torch.randn(batch, res, res, ch, device='cuda').to(dtype)
        w = (torch.randn(batch, 3, 3, ch, ch, device='cuda') / (3 * ch ** 0.5)).to(dtype)     # per-sample (modulated) weights
        gy = torch.randn(batch, res, res, ch, device='cuda').to(dtype)
        bias = torch.randn(ch, device='cuda')
        noise = torch.randn(batch, res, res, device='cuda').to(dtype)
        nw = torch.full((1,), 0.1, device='cuda')
        flops = 2.0 * batch * res * res * ch * ch * 9
        nbytes = 2.0 * batch * res * res * ch * x.element_size()
        # (the demodulation coefficients are folded into the per-sample weights by `modweight_fwd` / `ModulatedConv2d.operands`
        # for every layer of this list, gm.py:289 -- the epilogue carries noise, bias and the activation)
        for kind, fn in (('fwd', lambda: K.conv_fwd(x, w, res, res, 1, 1, 1, bias, None, noise, nw, 0.2, 2 ** 0.5)),
                         ('wgrad', lambda: K.conv_wgrad(x, gy, 3, 3, 1, 1, 1, True))):
            ms = _time_kernel(torch, fn, flush)
            what = f'conv3x3 {ch}->{ch} @{res}x{res} batch {batch}'
            gbs, tfs = nbytes / (ms * 1e-3) / 1e9, flops / (ms * 1e-3) / 1e12
            bound = 'hbm' if gbs / hbm >= tfs / tf else 'tensor'
            row = {'bound': bound, 'achieved': gbs if bound == 'hbm' else tfs, 'peak': hbm if bound == 'hbm' else tf,
                   'unit': 'GB/s' if bound == 'hbm' else 'TFLOP/s', 'frac': max(gbs / hbm, tfs / tf),
                   'traffic': (NCU_TRAFFIC_BYTES if kind == 'fwd' else NCU_TRAFFIC_BYTES_WGRAD).get(what) if batch == 16 else None,
                   'kernel': what + (' + fused epilogue (noise, bias, lrelu), per-sample demodulated weights' if kind == 'fwd'
                                     else ' weight gradient (per-sample)'),
                   'pass': kind, 'kernel_ms': ms, 'algorithmic_flops': flops, 'algorithmic_bytes': nbytes,
                   'hbm_gbs': gbs, 'frac_hbm': gbs / hbm, 'tensor_tflops': tfs, 'frac_tensor': tfs / tf,
                   'engine': K.last_conv_engine()}
            out.append(row)
        del x, w, gy, noise
    return out[0], out


# ---------------------------------------------------------------------------------------------
# one measured configuration of this repo's path
# ---------------------------------------------------------------------------------------------
def measure(torch, dist, args, dev, rank, world, size, batch, steps, warmup, fc_groups=None, r1=1.0, lr=0.002, mixing=0.0,
            reg=True, with_e2e=True, sampler_index=None, ncu_range=False):
    """Build G / g_ema / D for (size, batch), capture the step variants, time `steps` iterations starting at iteration 1
    (device-resident input) and -- same window, same cadence -- the host-fed end-to-end variant.  Everything is released
    before returning."""
    from gan_control_b200 import kernels as K, modules as M
    from gan_control_b200.train_step import GanTrainStep
    act = torch.bfloat16 if args.dtype == 'bf16' else torch.float32
    torch.manual_seed(1234)          # identical init on every rank (replicas stay in sync by construction)
    fc = M.FcConfig.from_sub_groups_dict(fc_groups) if fc_groups else None

    def gen():
        return M.Generator(size, 512, 8, channel_multiplier=2, conv_transpose=True, act_dtype=act, split_fc=fc is not None,
                           fc_config=fc).to(dev)
    g, g_ema = gen(), gen()
    d = M.Discriminator(size, channel_multiplier=2, act_dtype=act).to(dev)
    step = GanTrainStep(g, d, g_ema, batch=batch, world_size=world, r1=r1, lr_g=lr, lr_d=lr, mixing=mixing)
    torch.manual_seed(1000 + rank)   # independent latents / noise / data per replica
    host_real = torch.randn(batch, 3, size, size).clamp_(-1, 1).pin_memory()
    dev_real = host_real.to(dev)
    use_graph = not args.no_graph

    def sync():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(n_steps, e2e, first_iter):
        sync()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        n0 = K.launch_count() + getattr(step, 'replayed_launches', 0)
        e0.record()
        sink = 0.0
        for it in range(n_steps):
            i = first_iter + it
            if e2e and not use_graph:
                real = host_real.to(dev, non_blocking=True)      # this step's inputs from pinned host memory
            elif e2e:
                real = host_real                                 # copied H2D into the graph's static buffer
            else:
                real = dev_real
            d_loss, g_loss = run_step(i, real, regularize=reg)
            if e2e:
                sink += float(torch.stack([d_loss.float(), g_loss.float()]).cpu().sum())   # D2H read of the step's result
        e1.record()
        sync()
        ms = e0.elapsed_time(e1)
        if world > 1:
            t = torch.tensor([ms], device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms = float(t)
        return ms / n_steps, K.launch_count() + getattr(step, 'replayed_launches', 0) - n0

    if use_graph:
        step.capture(tuple(dev_real.shape))      # runs every step variant twice on a side stream, then captures
    run_step = step.train_step_graphed if use_graph else step.train_step
    for it in range(warmup):                     # warm-up (one of each regulariser so every kernel / allocation exists)
        run_step(it * 4, dev_real, regularize=reg and it == 0)
    sampler = ClockSampler(sampler_index) if sampler_index is not None else None
    if sampler:
        sampler.start()
    if ncu_range:
        torch.cuda.synchronize()
        torch.cuda.profiler.start()
    engines0 = K.engine_launches()
    ms_step, launches = timed(steps, False, 1)
    if ncu_range:
        torch.cuda.profiler.stop()
    clocks = sampler.summary() if sampler else None
    engines = {k: v - engines0[k] for k, v in K.engine_launches().items()} if not use_graph else None
    out = {'ms_per_step': ms_step, 'value': batch * world / (ms_step * 1e-3), 'gpu_launches': launches, 'clocks': clocks,
           'timed_window': dict(reg_counts(1, steps, reg), first_iteration=1, iterations=steps), 'cuda_graphs': use_graph,
           'graph_launches': dict(step.graph_launches) if use_graph else None, 'conv_engine_calls': engines}
    if with_e2e:
        ms_e2e, _ = timed(steps, True, 1)        # same iteration window (same regulariser steps) as `value`
        out['e2e'] = {'value': batch * world / (ms_e2e * 1e-3), 'unit': UNIT, 'ms_per_step': ms_e2e,
                      'h2d_bytes_per_step': host_real.numel() * 4, 'd2h_bytes_per_step': 8,
                      'timed_window': out['timed_window']}
    sync()
    if world > 1:
        # captured graphs hold NCCL collectives: destroying them mid-run can hang (see _finish) -- keep them to the end
        _KEEP.append((step, g, g_ema, d, host_real, dev_real))
    del step, g, g_ema, d, run_step, host_real, dev_real
    gc.collect()
    torch.cuda.empty_cache()
    return out


_KEEP = []


def gpu_library_baseline(torch, size, batch, dev, log):
    """The reference's FUSED=False arithmetic (oracle port = the same PyTorch ops as gan_model.py) on THIS GPU through
    cuDNN / cuBLAS: grouped `conv2d(groups=batch)` with materialised per-sample weights, `conv_transpose2d`, the 6-pass
    `upfirdn2d_native`, separate noise / bias / leaky-ReLU passes, stock autograd, torch.optim.Adam -- the stand-in for the
    FUSED=True comparator that is not in the reference tree.  Plain iterations (D step + G step) at the
    benchmark batch, accumulated over mini-batches the way the reference does when a batch does not fit (gt.py:361)."""
    from oracle.reference_step import ReferenceStep
    out = {'what': 'reference FUSED=False arithmetic (oracle/reference_step.py) on this GPU via cuDNN/cuBLAS; plain '
                   'iterations (no regulariser steps), 1 warm-up + 2 timed, CUDA events', 'unit': UNIT}
    for name, autocast in (('fp32', None), ('bf16_autocast', torch.bfloat16)):
        mini = batch                  # the whole batch at once when it fits; else the reference's mini-batch accumulation
        while True:
            try:
                torch.manual_seed(0)
                rs = ReferenceStep(size, batch, mini, device=dev, autocast=autocast)
                real = torch.randn(batch, 3, size, size, device=dev).clamp_(-1, 1)
                rs.train_step(1, real, regularize=False)
                torch.cuda.synchronize()
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record()
                for k in range(2):
                    rs.train_step(2 + k, real, regularize=False)
                e1.record()
                torch.cuda.synchronize()
                ms = e0.elapsed_time(e1) / 2
                out[name] = {'value': batch / (ms * 1e-3), 'ms_per_step': ms, 'batch': batch, 'mini_batch': mini,
                             'peak_mem_gb': torch.cuda.max_memory_allocated() / 2 ** 30}
                log(f'gpu_library_baseline {name}: {out[name]}')
                break
            except torch.cuda.OutOfMemoryError:
                mini //= 2
                if mini < 1:
                    out[name] = {'unavailable': 'out of memory at mini-batch 1'}
                    break
            finally:
                rs = real = None
                gc.collect()
                torch.cuda.empty_cache()
                torch.cuda.reset_peak_memory_stats()
    return out


# configs/metfaces.json `sub_groups_dict` (place_in_latent) and configs/afhq.json (192 / 192 / 128)
METFACES_GROUPS = {'id': {'place_in_latent': [0, 128]}, 'expression': {'place_in_latent': [128, 192]},
                   'orientation': {'place_in_latent': [192, 256]}, 'age': {'place_in_latent': [256, 320]},
                   'style': {'place_in_latent': [320, 448]}, 'other': {'place_in_latent': [448, 512]}}
AFHQ_GROUPS = {'a': {'place_in_latent': [0, 192]}, 'b': {'place_in_latent': [192, 384]}, 'c': {'place_in_latent': [384, 512]}}


def run_b200(args):
    import torch
    import torch.distributed as dist
    rank = int(os.environ.get('RANK', '0'))
    local_rank = int(os.environ.get('LOCAL_RANK', '0'))
    world = int(os.environ.get('WORLD_SIZE', '1'))
    assert world == args.gpus, f'--gpus {args.gpus} but WORLD_SIZE={world}'
    assert torch.cuda.is_available(), 'bench.py needs a CUDA device (no CPU fallback)'
    torch.cuda.set_device(local_rank)
    dev = torch.device('cuda', local_rank)
    if world > 1:
        dist.init_process_group('nccl', device_id=dev)
    import __graft_entry__
    if rank == 0:
        __graft_entry__.build()
    if world > 1:
        dist.barrier()
    from gan_control_b200 import kernels as K
    log = (lambda m: print(m, file=sys.stderr, flush=True)) if rank == 0 else (lambda m: None)
    size, batch, reg = args.size, args.batch, not args.no_reg
    main = measure(torch, dist, args, dev, rank, world, size, batch, args.steps, args.warmup, reg=reg,
                   sampler_index=local_rank if rank == 0 else None, ncu_range=args.ncu_range)
    log(f'main: {main["value"]:.1f} img/s, e2e {main["e2e"]["value"]:.1f}')
    extra = []
    if not args.skip_extra and not args.ncu_range:
        # the per-GPU workloads of the other BASELINE.json configurations, same run, same timing method
        for name, kw in (
                ('configs[2]: FFHQ-1024 DDP, 4 images / GPU, R1 + path-length at cadence', dict(size=1024, batch=4)),
                ('configs[3]: MetFaces-1024 layout (split-FC 128/64/64/64/128/64, r1 2), style mixing 0.9 drawn on the device '
                 'and captured in the CUDA graphs, 4 images / GPU', dict(size=1024, batch=4, fc_groups=METFACES_GROUPS, r1=2.0, mixing=0.9)),
                ('configs[4]: AFHQ-512 layout (split-FC 192/192/128, r1 0.5, lr 0.0025), 8 images / GPU',
                 dict(size=512, batch=8, fc_groups=AFHQ_GROUPS, r1=0.5, lr=0.0025))):
            m = measure(torch, dist, args, dev, rank, world, steps=16, warmup=2, with_e2e=False, **kw)
            extra.append({'config': name, 'global_batch': kw['batch'] * world, 'per_gpu_batch': kw['batch'], 'n_gpus': world,
                          'value': m['value'], 'unit': UNIT, 'ms_per_step': m['ms_per_step'], 'steps': 16,
                          'timed_window': m['timed_window'], 'gpu_launches': m['gpu_launches']})
            log(f'extra {name}: {m["value"]:.1f} img/s')
    if rank != 0:
        _finish(world)
        return
    peaks, peak_kind = measured_peaks()
    value = main['value']
    g_f, d_f = G_D_GFLOP.get(size, (0, 0))
    g_mb, d_mb = G_D_MB.get(size, (0, 0))
    tflops = (4 * g_f + 8 * d_f) * value / 1e3
    gbs = (4 * g_mb + 8 * d_mb) * value / 1e3
    line = {
        'metric': METRIC, 'value': value, 'unit': UNIT, 'n_gpus': world, 'steps': args.steps, 'warmup': args.warmup,
        'ms_per_step': main['ms_per_step'], 'higher_is_better': True, 'scaling': 'weak', 'vs_baseline': None,
        'dtype': 'bf16' if args.dtype == 'bf16' else 'f32', 'data': 'synthetic',
        'config': workload_config(size, reg, batch, world), 'cuda_graphs': main['cuda_graphs'],
        'timed_window': main['timed_window'], 'clocks': main['clocks'], 'e2e': main['e2e'],
        'gpu_launches': main['gpu_launches'], 'graph_launches': main['graph_launches'],
        'conv_engine_calls': main['conv_engine_calls'],
        # the whole step against the machine: plain-step algorithmic work (4 G_f + 8 D_f; the
        # regulariser steps inside the window add real work that is NOT counted, so this is a lower bound)
        'roofline_step': {'tflops': tflops, 'frac_sustained': tflops / peaks.get('bf16_tflops_sustained', 1400.0),
                          'gbs': gbs, 'frac_hbm': gbs / peaks.get('hbm_gbs', 6650.0),
                          'algorithmic_gflop_per_image': 4 * g_f + 8 * d_f, 'algorithmic_mb_per_image': 4 * g_mb + 8 * d_mb,
                          'peak_source': peak_kind},
        'step_tflops': tflops,
    }
    if extra:
        line['extra_configs'] = extra
    if not args.skip_roofline:
        import torch as _t
        head, layers = conv_roofline(torch, K, _t.bfloat16 if args.dtype == 'bf16' else _t.float32, size, batch, peaks)
        line['roofline'] = dict(head, peak_source=peak_kind + ' (burst, kernel timed alone)')
        line['roofline_layers'] = layers
    if not args.skip_library_baseline and world == 1 and not args.ncu_range:
        line['gpu_library_baseline'] = gpu_library_baseline(torch, size, batch, dev, log)
        for k in ('fp32', 'bf16_autocast'):
            v = line['gpu_library_baseline'].get(k, {}).get('value')
            if v:
                line['gpu_library_baseline'][k]['speedup_of_this_repo'] = value / v
    if not args.skip_cpu_baseline and world == 1:        # the CPU baseline is reported by the single-GPU run only
        threads = os.cpu_count() or 1
        csize = args.cpu_sample_size or min(size, 256)       # bounded: the full-resolution probe is the reference arm's job
        v, ms, sample, used = cpu_reference_steps(csize, 3, 1, threads, 60.0)
        line['cpu_baseline'] = {'value': v, 'unit': UNIT, 'cores': threads, 'kind': 'port', 'sample': sample,
                                'sample_resolution': used}
    print(json.dumps(line), flush=True)
    _finish(world)


def _finish(world):
    """Multi-rank runs leave through os._exit: tearing NCCL communicators down while captured CUDA graphs still
    reference them can hang interpreter shutdown (observed on 2x B200: JSON printed, process never exited)."""
    if world > 1:
        import torch
        torch.cuda.synchronize()
        sys.stdout.flush()
        sys.stderr.flush()
        os._exit(0)


def run_config5(args):
    """BASELINE.json configs[4], inference half: AFHQ layout (3 groups 192/192/128, split-FC mapping) at 512^2, controller
    FcStack(lr_mlp=0.01, n_mlp=4, in_dim=3, mid_dim=512, out_dim=192) (configs/controller_configs/afhq/
    default_w_latent_controller.json) swept over the 1000 `orientation` control vectors of the reference's fixture
    resources/ffhq_1K_attributes_samples_df.pkl (committed as tests/golden/config5_controls.npz), its `latents_w` as the
    base w.  Two numbers: latents/s of the controller alone, images/s of `gen_batch_by_controls` (CUDA-graphed synthesis)."""
    import numpy as np
    import torch
    import __graft_entry__
    __graft_entry__.build()
    from gan_control_b200 import modules as M
    from gan_control_b200.inference import Controller
    groups = {'id': {'place_in_latent': [0, 192]}, 'orientation': {'place_in_latent': [192, 384]}, 'other': {'place_in_latent': [384, 512]}}
    fx = np.load(os.path.join(ROOT, 'tests', 'golden', 'config5_controls.npz'))
    ori = torch.from_numpy(fx['orientation']).cuda()
    w_base = torch.from_numpy(fx['latents_w'].astype(np.float32)).cuda()
    torch.manual_seed(0)
    g = M.Generator(512, 512, 8, channel_multiplier=2, conv_transpose=True, split_fc=True,
                    fc_config=M.FcConfig.from_sub_groups_dict(groups), act_dtype=torch.bfloat16)
    out = {}
    for graphs in (True, False):
        c = Controller(generator=g, sub_groups_dict=groups, fc_controls={'orientation': M.FcStack(0.01, 4, 3, 512, 192)},
                       cuda_graphs=graphs)
        bs = 50

        def sweep():
            for i in range(0, 1000, bs):
                c.gen_batch_by_controls(latent=w_base[i:i + bs], input_is_latent=True, normalize=False, orientation=ori[i:i + bs])

        def timed(fn, reps):
            for _ in range(max(3, args.warmup)):
                fn()
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(reps):
                fn()
            e1.record()
            torch.cuda.synchronize()
            return e0.elapsed_time(e1) / reps
        with torch.no_grad():
            ms_ctl = timed(lambda: c.fc_controls['orientation'](ori), 20)
            ms_gen = timed(sweep, max(1, args.steps // 8))
        out['cuda_graphs' if graphs else 'eager'] = {'controller_latents_per_s': 1000 / ms_ctl * 1e3,
                                                     'gen_batch_by_controls_images_per_s': 1000 / ms_gen * 1e3,
                                                     'ms_per_1000_images': ms_gen}
    best = out['cuda_graphs']
    print(json.dumps({'metric': 'images/sec gen_batch_by_controls 512x512 over the 1K control vectors (BASELINE.json configs[4])',
                      'value': best['gen_batch_by_controls_images_per_s'], 'unit': UNIT, 'n_gpus': 1, 'higher_is_better': True,
                      'dtype': 'bf16', 'data': 'reference fixture ffhq_1K_attributes_samples_df.pkl (orientation, latents_w); '
                      'random-init weights', 'config': {'workload': 'configs[4] controller EqualLinear inference sweep (1K latents), '
                                                                    'batches of 50'}, 'detail': out}), flush=True)


if __name__ == '__main__':
    a = parse()
    if a.config == 5 and a.impl != 'reference':
        run_config5(a)
    elif a.impl == 'reference':
        run_reference(a)
    else:
        run_b200(a)
