"""Device time of each captured step variant (d, d_reg, g, g_reg) and the kernel table of one of them.
    python scripts/time_variants.py [variant-to-profile ...]"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import __graft_entry__  # noqa: E402

__graft_entry__.build()
from gan_control_b200 import modules as M  # noqa: E402
from gan_control_b200.train_step import GanTrainStep  # noqa: E402

dev = torch.device('cuda')
act = torch.bfloat16
size, batch = 1024, 16
g = M.Generator(size, 512, 8, channel_multiplier=2, conv_transpose=True, act_dtype=act).to(dev)
g_ema = M.Generator(size, 512, 8, channel_multiplier=2, conv_transpose=True, act_dtype=act).to(dev)
d = M.Discriminator(size, channel_multiplier=2, act_dtype=act).to(dev)
step = GanTrainStep(g, d, g_ema, batch=batch)
real = torch.randn(batch, 3, size, size, device=dev).clamp_(-1, 1)
step.capture(real.shape)
step.static_real.copy_(real)
for name in ('d', 'd_reg', 'g', 'g_reg'):
    step.graphs[name].replay()
    torch.cuda.synchronize()
    ts = []
    for _ in range(4):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        step.graphs[name].replay()
        e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    print(f'{name:6s} {sorted(ts)[1]:8.2f} ms   ({step.graph_launches[name]} libb200gan launches)', flush=True)
for which in sys.argv[1:]:
    from torch.profiler import ProfilerActivity, profile
    with profile(activities=[ProfilerActivity.CUDA]) as prof:
        step.graphs[which].replay()
        torch.cuda.synchronize()
    print(f'==== kernel table of one replay of the {which!r} graph ====')
    print(prof.key_averages().table(sort_by='cuda_time_total', row_limit=45, max_name_column_width=90))
