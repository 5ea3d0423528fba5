#!/usr/bin/env python
"""bench.py -- images/sec of the FFHQ-1024 G+D train step on 1/2/4/8 B200 (BASELINE.json metric).

    python bench.py --gpus N --steps K --warmup W            # this repo's CUDA path
    python bench.py --impl reference --steps K --warmup W    # the reference's CPU path (oracle port)

A "step" is one `discriminator_update` + one `generator_update` (generator_trainer.py:351-353) at
BASELINE.json configs[1]: Generator(1024)/Discriminator(1024), channel_multiplier 2, per-GPU batch
16, bf16 activations / fp32 accumulate and parameters, synthetic images, random-init weights, with
the lazy regularisers at their real cadence (R1 every 16, path-length every 4 iterations).
Rank 0 prints ONE JSON line (see DESIGN.md "Measurement").
"""
import argparse
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
# algorithmic work per image of the plain step, BASELINE.md section 2: 4*G_f + 8*D_f
GFLOP_PER_IMG = {1024: 4 * 148.5 + 8 * 153.3, 512: 4 * 119.3 + 8 * 123.2, 256: 4 * 90.2 + 8 * 93.1}


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
    ap.add_argument('--no-graph', action='store_true', help='launch every kernel from Python instead of replaying CUDA graphs')
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
    """`config` of the JSON line: the same for this repo's arm and for the reference arm"""
    return {'workload': f'FFHQ-{size} G+D train step (BASELINE.json configs[1]): Generator({size})+Discriminator({size}) '
                        f'channel_multiplier 2, random init, lazy regularisers at real cadence '
                        f'(R1 /16, path-length /4)' + ('' if reg else ' DISABLED'),
            'global_batch': batch * world, 'per_gpu_batch': batch, 'parallelism': f'dp{world}'}


# ---------------------------------------------------------------------------------------------
def cpu_reference_sample(size, threads, reps=1):
    """Bounded sample of the reference's CPU path (oracle port of gan_model.py, fp32, FUSED=False arithmetic):
    ONE generator forward + ONE discriminator forward on 1 image at `size`.  The full plain G+D step costs
    (4*G_f + 8*D_f) / (G_f + D_f) times the FLOPs of this sample (forward, data-gradient
    and weight-gradient passes cost one forward each), so images/s = 1 / (t_sample * that ratio).  A whole
    1024^2 step on the box's host cores takes minutes (measured: 259 s on 128 cores), hence the sample."""
    import torch
    from oracle import params as P, stylegan2_oracle as O
    torch.set_num_threads(threads)
    sd_g = P.seeded_state_dict(P.generator_shapes(size, 512, 8, 2), 1)
    sd_d = P.seeded_state_dict(P.discriminator_shapes(size, 2), 2)
    z = torch.randn(1, 512)
    ts = []
    with torch.no_grad():
        for _ in range(reps):
            t0 = time.perf_counter()
            img = O.generator_forward(sd_g, [z], size)
            O.discriminator_forward(sd_d, img, size)
            ts.append(time.perf_counter() - t0)
    t = min(ts)
    g_f, d_f = {1024: (148.5, 153.3), 512: (119.3, 123.2), 256: (90.2, 93.1)}.get(size, (148.5, 153.3))
    ratio = (4 * g_f + 8 * d_f) / (g_f + d_f)
    return 1.0 / (t * ratio), t, ratio


def run_reference(args):
    """--impl reference: the reference's CPU path on this box's host cores (oracle port; a Python reference
    cannot travel to the GPU box).  Each of the K steps is one bounded sample (see cpu_reference_sample)."""
    rank = int(os.environ.get('RANK', '0'))
    if rank != 0:
        return
    threads = os.cpu_count() or 1
    size = args.cpu_sample_size or args.size
    reps = max(1, min(args.steps, 2))
    v, t, ratio = cpu_reference_sample(size, threads, reps=reps)
    sample = (f'G forward + D forward on 1 image at {size}x{size}, fp32, {t:.1f} s (best of {reps}); scaled to the plain '
              f'G+D step by its FLOP ratio {ratio:.2f}')
    line = {'impl': 'reference', 'metric': METRIC, 'value': v, 'unit': UNIT, 'n_gpus': args.gpus, 'steps': args.steps,
            'warmup': args.warmup, 'ms_per_step': t * ratio * 1e3, 'higher_is_better': True, 'scaling': 'weak',
            'vs_baseline': None, 'dtype': 'f32', 'data': 'synthetic',
            'config': dict(workload_config(args.size, not args.no_reg, args.batch, args.gpus),
                           reference_arm='reference FUSED=False arithmetic (oracle port) on the host CPU; every step is a bounded '
                                         'sample of this workload, see cpu_baseline.sample'),
            'cpu_baseline': {'value': v, 'unit': UNIT, 'cores': threads, 'kind': 'port', 'sample': sample},
            'e2e': {'value': v, 'unit': UNIT, 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0}, 'gpu_launches': 0}
    print(json.dumps(line), flush=True)


# ---------------------------------------------------------------------------------------------
# dram__bytes_read.sum + dram__bytes_write.sum of ONE launch at batch 16, from the `ncu --set full` captures summarised in
# profiles/r01_ncu_kernels.md (None where the layer has not been captured)
NCU_TRAFFIC_BYTES = {'conv3x3 256->256 @128x128 batch 16': 135.4e6 + 85.79e6,
                     'conv3x3 128->128 @256x256 batch 16': 268.762e6 + 220.351e6,
                     'conv3x3 64->64 @512x512 batch 16': 537.110e6 + 487.299e6,
                     'conv3x3 32->32 @1024x1024 batch 16': 1.073810e9 + 1.025126e9}


def conv_roofline(torch, K, dtype, size, batch, peaks):
    """The dominant kernels = the 3x3 convolutions (conv_fwd_umma_kernel has the largest share of the step,
    conv_fwd_halo_kernel the second, profiles/).  Each layer class of the resolution is timed alone with CUDA events
    on the launching stream (inputs >> L2, L2 flushed between launches) against the roofline that bounds it:
    tensor pipe (algorithmic 2*MAC FLOPs) for >= 64 channels, HBM (input read once + output written once) for the
    32-channel full-resolution layer.  Returns (headline, all layers)."""
    flush = torch.empty(256 << 20, dtype=torch.uint8, device='cuda')
    layers = [(size // 8, 256, 'tensor'), (size // 4, 128, 'tensor'), (size // 2, 64, 'tensor'), (size, 32, 'hbm')]
    out = []
    for res, ch, bound in layers:
        x = torch.randn(batch, res, res, ch, device='cuda').to(dtype)
        w = (torch.randn(batch, 3, 3, ch, ch, device='cuda') / (3 * ch ** 0.5)).to(dtype)     # per-sample (modulated) weights
        flops = 2.0 * batch * res * res * ch * ch * 9
        nbytes = 2.0 * batch * res * res * ch * x.element_size()
        for _ in range(3):
            K.conv_fwd(x, w, res, res, 1, 1, 1)
        ts = []
        for _ in range(10):
            flush.zero_()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            K.conv_fwd(x, w, res, res, 1, 1, 1)
            e1.record()
            torch.cuda.synchronize()
            ts.append(e0.elapsed_time(e1))
        ms = sum(ts) / len(ts)
        what = f'conv3x3 {ch}->{ch} @{res}x{res} batch {batch}'
        if bound == 'tensor':
            ach, peak, unit = flops / (ms * 1e-3) / 1e12, peaks.get('bf16_tflops', 1590.0), 'TFLOP/s'
        else:
            ach, peak, unit = nbytes / (ms * 1e-3) / 1e9, peaks.get('hbm_gbs', 6650.0), 'GB/s'
        out.append({'bound': bound, 'achieved': ach, 'peak': peak, 'unit': unit, 'frac': ach / peak,
                    'traffic': NCU_TRAFFIC_BYTES.get(what) if batch == 16 else None, 'kernel': what, 'kernel_ms': ms,
                    'algorithmic_flops': flops, 'algorithmic_bytes': nbytes})
        del x, w
    return out[0], out


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
    from gan_control_b200 import kernels as K, modules as M
    from gan_control_b200.train_step import GanTrainStep
    act = torch.bfloat16 if args.dtype == 'bf16' else torch.float32
    size, batch = args.size, args.batch
    torch.manual_seed(1234)          # identical init on every rank (replicas stay in sync by construction)
    g = M.Generator(size, 512, 8, channel_multiplier=2, conv_transpose=True, act_dtype=act).to(dev)
    g_ema = M.Generator(size, 512, 8, channel_multiplier=2, conv_transpose=True, act_dtype=act).to(dev)
    d = M.Discriminator(size, channel_multiplier=2, act_dtype=act).to(dev)
    step = GanTrainStep(g, d, g_ema, batch=batch, world_size=world)
    torch.manual_seed(1000 + rank)   # independent latents / noise / data per replica
    host_real = torch.randn(batch, 3, size, size).clamp_(-1, 1).pin_memory()
    dev_real = host_real.to(dev)
    reg = not args.no_reg

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

    use_graph = not args.no_graph
    if use_graph:
        step.capture(tuple(dev_real.shape))      # runs every step variant twice on a side stream, then captures
    run_step = step.train_step_graphed if use_graph else step.train_step
    # warm-up (also runs one of each regulariser so every kernel / allocation exists)
    for it in range(args.warmup):
        run_step(it * 4, dev_real, regularize=reg and it == 0)
    sampler = ClockSampler(local_rank) if rank == 0 else None
    if sampler:
        sampler.start()
    if args.ncu_range:
        torch.cuda.synchronize()
        torch.cuda.profiler.start()
    ms_step, launches = timed(args.steps, False, 1)
    if args.ncu_range:
        torch.cuda.profiler.stop()
    clocks = sampler.summary() if sampler else None
    ms_e2e, _ = timed(max(1, args.steps // 2), True, 1)
    ms_plain, _ = timed(max(1, min(args.steps, 4)), False, 1) if False else (None, None)
    global_batch = batch * world
    value = global_batch / (ms_step * 1e-3)
    e2e_value = global_batch / (ms_e2e * 1e-3)
    if rank != 0:
        _finish(world)
        return
    peaks, peak_kind = measured_peaks()
    line = {
        'metric': METRIC, 'value': value, 'unit': UNIT, 'n_gpus': world, 'steps': args.steps, 'warmup': args.warmup,
        'ms_per_step': ms_step, 'higher_is_better': True, 'scaling': 'weak', 'vs_baseline': None,
        'dtype': 'bf16' if act == torch.bfloat16 else 'f32', 'data': 'synthetic',
        'config': dict(workload_config(size, reg, batch, world), cuda_graphs=use_graph,
                       l2='inputs larger than L2 (each step streams >10 GB of activations; no explicit flush)'),
        'clocks': clocks,
        'e2e': {'value': e2e_value, 'unit': UNIT, 'ms_per_step': ms_e2e,
                'h2d_bytes_per_step': host_real.numel() * 4, 'd2h_bytes_per_step': 8},
        'gpu_launches': launches,
        'step_tflops': GFLOP_PER_IMG.get(size, 0) * value / 1e3,
    }
    if not args.skip_roofline:
        head, layers = conv_roofline(torch, K, act, size, batch, peaks)
        line['roofline'] = dict(head, peak_source=peak_kind + ' (burst, kernel timed alone)')
        line['roofline_layers'] = layers
    if not args.skip_cpu_baseline and world == 1:        # the CPU baseline is reported by the single-GPU run only
        threads = os.cpu_count() or 1
        csize = args.cpu_sample_size or size
        v, t, ratio = cpu_reference_sample(csize, threads)
        line['cpu_baseline'] = {'value': v, 'unit': UNIT, 'cores': threads, 'kind': 'port',
                                'sample': f'G forward + D forward on 1 image at {csize}x{csize}, fp32, {t:.1f} s; scaled to the '
                                          f'plain G+D step by its FLOP ratio {ratio:.2f}'}
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


if __name__ == '__main__':
    a = parse()
    if a.impl == 'reference':
        run_reference(a)
    else:
        run_b200(a)
