"""Per-kernel device-time table of the train step (torch.profiler / CUPTI), for optimisation work.
    python scripts/profile_step.py [--size 1024] [--batch 16] [--reg]"""
import argparse
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import __graft_entry__  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument('--size', type=int, default=1024)
ap.add_argument('--batch', type=int, default=16)
ap.add_argument('--reg', action='store_true')
ap.add_argument('--rows', type=int, default=45)
args = ap.parse_args()
__graft_entry__.build()
from gan_control_b200 import modules as M  # noqa: E402
from gan_control_b200.train_step import GanTrainStep  # noqa: E402

dev = torch.device('cuda')
act = torch.bfloat16
g = M.Generator(args.size, 512, 8, channel_multiplier=2, conv_transpose=True, act_dtype=act).to(dev)
g_ema = M.Generator(args.size, 512, 8, channel_multiplier=2, conv_transpose=True, act_dtype=act).to(dev)
d = M.Discriminator(args.size, channel_multiplier=2, act_dtype=act).to(dev)
step = GanTrainStep(g, d, g_ema, batch=args.batch)
real = torch.randn(args.batch, 3, args.size, args.size, device=dev).clamp_(-1, 1)
for i in range(2):
    step.train_step(1, real, regularize=False)
if args.reg:
    step.train_step(0, real, regularize=True)
torch.cuda.synchronize()
from torch.profiler import ProfilerActivity, profile  # noqa: E402
with profile(activities=[ProfilerActivity.CPU, ProfilerActivity.CUDA]) as prof:
    if args.reg:
        step.train_step(0, real, regularize=True)
    else:
        step.train_step(1, real, regularize=False)
        step.train_step(1, real, regularize=False)
    torch.cuda.synchronize()
print(prof.key_averages().table(sort_by='cuda_time_total', row_limit=args.rows, max_name_column_width=70))
