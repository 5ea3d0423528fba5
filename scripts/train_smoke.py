"""GPU smoke of the training loop: a 64^2 split-FC config (afhq group layout) trains 8 iterations through CUDA graphs, writes
reference-format checkpoints, resumes 4 more, and the run directory loads through the inference front end."""
import json
import os
import sys
import tempfile

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import __graft_entry__  # noqa: E402

__graft_entry__.build()
from gan_control_b200 import train as T  # noqa: E402
from gan_control_b200.inference import Inference  # noqa: E402

groups = {'dog_id': {'place_in_latent': [0, 192]}, 'orientation': {'place_in_latent': [192, 384]}, 'other': {'place_in_latent': [384, 512]}}
cfg = {'model_config': {'vanilla': False, 'img_channels': 3, 'split_fc': True, 'marge_fc': False, 'latent_size': 512, 'size': 64, 'n_mlp': 8,
                        'channel_multiplier': 2.0, 'conv_transpose': True, 'g_noise_mode': 'normal'},
       'training_config': {'iter': 8, 'start_iter': 0, 'batch': 8, 'mini_batch': 8, 'sub_groups_dict': groups, 'r1': 0.5, 'g_reg_every': 4,
                           'd_reg_every': 16, 'lr_g': 0.0025, 'lr_d': 0.0025, 'g_moving_average': 10000, 'path_regularize': 2,
                           'path_batch_shrink': 2, 'mixing': 0, 'save_nets_interval': 4},
       'ckpt_config': {'enabled': False, 'ckpt': 'no_ckpt'}}
d = tempfile.mkdtemp()
step = T.train(json.loads(json.dumps(cfg)), save_dir=d, log_every=4)
print(sorted(os.listdir(os.path.join(d, 'checkpoint'))), 'graphs:', step.graphs is not None)
cfg['ckpt_config'] = {'enabled': True, 'ckpt': os.path.join(d, 'checkpoint', '000008.pt')}
step2 = T.train(cfg, save_dir=d, iters=4, log_every=2)
inf = Inference(d, device='cuda')
img = inf.gen_batch(batch_size=2)[0]
print('resumed to', sorted(os.listdir(os.path.join(d, 'checkpoint')))[-1], 'inference image', tuple(img.shape), bool(torch.isfinite(img.float()).all()))
assert float(step2.g_optim.t[0]) == float(step.g_optim.t[0]) + 5      # iterations 8..11: 4 plain + 1 path-length step
print('train smoke ok')
