"""Training loop on the B200 path, the counterpart of ``train_generator.py`` / ``GeneratorTrainer.train``
(generator_trainer.py:329-353) for the vanilla objective (adversarial + R1 + path-length; the attribute losses,
evaluation and dataset plumbing of gan-control are outside the hot path):

    torchrun --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 -m gan_control_b200.train \
        --config_path configs/ffhq.json --save_dir results/ffhq --iter 1000

One process per GPU; the config's ``batch`` is the global batch.  Images come from ``data`` (any iterator of
``(B_local, 3, size, size)`` tensors in [-1, 1]); without one, synthetic images are used (there are no datasets in this
environment).  Checkpoints are written in the reference's format every ``save_nets_interval`` iterations and resume with
``ckpt_config`` / ``--ckpt`` exactly like gt.py:175-193 (iteration number taken from the file name).
"""
import argparse
import json
import os

import torch
import torch.distributed as dist

from .train_step import GanTrainStep


def synthetic_images(batch, size, channels, device, seed=0):
    gen = torch.Generator(device='cpu').manual_seed(seed)
    while True:
        yield torch.randn(batch, channels, size, size, generator=gen).clamp_(-1, 1).to(device, non_blocking=True)


def train(config, save_dir=None, iters=None, data=None, device='cuda', act_dtype=torch.bfloat16, use_graphs=None, log_every=100,
          log=print, vanilla_only=False):
    """Runs iterations ``start_iter .. iter`` of the config (gt.py:340-353) and returns the GanTrainStep."""
    if not isinstance(config, dict):
        with open(config) as f:
            config = json.load(f)
    world = dist.get_world_size() if dist.is_available() and dist.is_initialized() else 1
    rank = dist.get_rank() if world > 1 else 0
    mc, tc = config['model_config'], config['training_config']
    torch.manual_seed(1234)                                   # identical initialisation on every replica
    step = GanTrainStep.from_config(config, device=device, world_size=world, act_dtype=act_dtype, vanilla_only=vanilla_only)
    if rank == 0 and step.effective_objective['ignored_config_terms']:
        log('WARNING: vanilla objective only; ignored config terms: ' + ', '.join(step.effective_objective['ignored_config_terms']))
    start = int(tc.get('start_iter', 0))
    ck = config.get('ckpt_config') or {}
    if ck.get('enabled'):                                     # gt.py:181-185
        try:
            start = int(os.path.splitext(os.path.basename(ck['ckpt']))[0])
        except ValueError:
            pass
    total = int(tc['iter']) if iters is None else start + int(iters)
    torch.manual_seed(1000 + rank + 7919 * start)             # independent latents / noise per replica
    if data is None:
        data = synthetic_images(step.batch, mc['size'], mc['img_channels'], device, seed=rank)
    if use_graphs is None:
        use_graphs = torch.device(device).type == 'cuda' and step.ada is None
    if use_graphs:
        step.capture((step.batch, mc['img_channels'], mc['size'], mc['size']))
    run = step.train_step_graphed if use_graphs else step.train_step
    save_every = int(tc.get('save_nets_interval', 0) or 0)
    for i in range(start, total):
        d_loss, g_loss = run(i, next(data))
        if rank == 0 and log_every and i % log_every == 0:
            log(f'iter {i}: d_loss {float(d_loss) * step.global_batch:.4f} g_loss {float(g_loss):.4f}')
        if rank == 0 and save_dir and save_every and i > start and i % save_every == 0:
            step.save_nets(i, save_dir)
    if rank == 0 and save_dir:
        step.save_nets(total, save_dir)
        with open(os.path.join(save_dir, 'args.json'), 'w') as f:          # read back by inference.Inference.retrieve_model
            json.dump(dict(config, b200gan_effective_objective=step.effective_objective), f, indent=1)
    return step


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--config_path', type=str, required=True)
    ap.add_argument('--save_dir', type=str, default=None)
    ap.add_argument('--iter', type=int, default=None, help='iterations to run from the start / resume point (default: the config\'s)')
    ap.add_argument('--ckpt', type=str, default=None, help='resume from this reference-format checkpoint')
    ap.add_argument('--dtype', default='bf16', choices=['bf16', 'fp32'])
    ap.add_argument('--vanilla-only', action='store_true',
                    help='train adversarial + R1 + path-length only even if the config enables attribute losses / ADA / '
                         'd_every / transfer learning (which this path does not build); without it such a config is rejected')
    args = ap.parse_args()
    with open(args.config_path) as f:
        config = json.load(f)
    if args.ckpt:
        config['ckpt_config'] = {'enabled': True, 'ckpt': args.ckpt}
    world = int(os.environ.get('WORLD_SIZE', '1'))
    local_rank = int(os.environ.get('LOCAL_RANK', '0'))
    torch.cuda.set_device(local_rank)
    if world > 1:
        dist.init_process_group('nccl', device_id=torch.device('cuda', local_rank))
    train(config, save_dir=args.save_dir, iters=args.iter, device=f'cuda:{local_rank}',
          act_dtype=torch.bfloat16 if args.dtype == 'bf16' else torch.float32, vanilla_only=args.vanilla_only)
    if world > 1:
        torch.cuda.synchronize()
        os._exit(0)                                           # see bench.py::_finish


if __name__ == '__main__':
    main()
