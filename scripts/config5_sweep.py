"""BASELINE.json configs[4], inference half: AFHQ layout (3 groups 192/192/128, split-FC mapping) at 512^2,
controller FcStack(lr_mlp=0.01, n_mlp=4, in_dim=3, mid_dim=512, out_dim=192) swept over the 1000 orientation
control vectors of the reference's data fixture; reports latents/s for the controller alone (one persistent
kernel per batch) and for `gen_batch_by_controls` (controller + generator forward from w)."""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import __graft_entry__  # noqa: E402

__graft_entry__.build()
from gan_control_b200 import modules as M  # noqa: E402
from gan_control_b200.inference import Controller  # noqa: E402

groups = {'id': {'place_in_latent': [0, 192]}, 'orientation': {'place_in_latent': [192, 384]}, 'other': {'place_in_latent': [384, 512]}}
fx = np.load(os.path.join(ROOT, 'tests', 'golden', 'config5_controls.npz'))
ori = torch.from_numpy(fx['orientation']).cuda()
w_base = torch.from_numpy(fx['latents_w'].astype(np.float32)).cuda()
torch.manual_seed(0)
g = M.Generator(512, 512, 8, channel_multiplier=2, conv_transpose=True, split_fc=True,
                fc_config=M.FcConfig.from_sub_groups_dict(groups), act_dtype=torch.bfloat16)
ctl = M.FcStack(0.01, 4, 3, 512, 192)
c = Controller(generator=g, sub_groups_dict=groups, fc_controls={'orientation': ctl})


def timed(fn, reps=3):
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
    ms_ctl = timed(lambda: c.fc_controls['orientation'](ori), reps=20)
    bs = 50
    def sweep():
        for i in range(0, 1000, bs):
            c.gen_batch_by_controls(latent=w_base[i:i + bs], input_is_latent=True, normalize=False, orientation=ori[i:i + bs])
    ms_gen = timed(sweep, reps=2)
print(f'controller FcStack on 1000 control vectors: {ms_ctl:.3f} ms -> {1000 / ms_ctl * 1e3:.0f} latents/s')
print(f'gen_batch_by_controls 512^2, 1000 latents in batches of {bs}: {ms_gen:.1f} ms -> {1000 / ms_gen * 1e3:.1f} images/s')
