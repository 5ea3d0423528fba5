"""Inference / Controller front end (reference `inference/inference.py`, `inference/controller.py`) on the
product modules vs the oracle, CPU + kernel stand-ins, incl. the reference's on-disk layout."""
import json
import os

import numpy as np
import torch

from gan_control_b200 import modules as M
from gan_control_b200.inference import Controller, Inference
from oracle import params as P
from oracle import stylegan2_oracle as O
from golden_io import max_rel

GROUPS = {'id': {'place_in_latent': [0, 24]}, 'pose': {'place_in_latent': [24, 40]}, 'other': {'place_in_latent': [40, 64]}}
FCG = [('id', 0, 24), ('pose', 24, 40), ('other', 40, 64)]
SIZE, SDIM, NMLP = 8, 64, 2


def rnd(seed, *shape):
    return torch.from_numpy(np.random.default_rng(seed).standard_normal(shape)).float()


def make_generator():
    g = M.Generator(SIZE, SDIM, NMLP, channel_multiplier=2, conv_transpose=True, split_fc=True,
                    fc_config=M.FcConfig.from_sub_groups_dict(GROUPS))
    sd = P.seeded_state_dict(P.generator_shapes(SIZE, SDIM, NMLP, 2, FCG), 77)
    g.load_state_dict(sd)
    return g, sd


def test_gen_batch_and_controls(cpu_kernels, tmp_path):
    g, sd = make_generator()
    ctl = M.FcStack(0.01, 3, 3, 32, 16)
    ctl_sd = P.seeded_state_dict(P.fc_stack_shapes(3, 3, 32, 16), 78)
    ctl.load_state_dict(ctl_sd)
    c = Controller(generator=g, sub_groups_dict=GROUPS, latent_size=SDIM, device='cpu', act_dtype=torch.float32,
                   fc_controls={'pose': ctl})
    assert c.sub_group_names == ['id', 'pose', 'other']
    z = rnd(1, 3, SDIM)
    img, lat, lat_w = c.gen_batch(latent=z.clone(), normalize=False)
    noise = c.expend_noise(c.noise, 3)
    ref = O.generator_forward(sd, [z], SIZE, FCG, noise=noise)
    assert max_rel(img, ref) < 1e-5 and lat_w.shape == (3, 2 * 3 - 2, SDIM)
    # truncation toward per-group mean latents (inference.py:73-87)
    c.calc_mean_w_latents(n_batches=2, batch=64)
    img_t, _, _ = c.gen_batch(latent=z.clone(), normalize=True, truncation=0.7)
    w = O.mapping_network(sd, z, FCG)
    w_t = c.mean_w_latent + 0.7 * (w - c.mean_w_latent)
    ref_t = O.generator_forward(sd, [w_t], SIZE, FCG, noise=c.expend_noise(c.noise, 3), input_is_latent=True)
    assert max_rel(img_t, ref_t.mul(0.5).add(0.5).clamp(0, 1)) < 1e-4
    # controls: the FcStack output replaces the group's slice of w (controller.py:30-54,60-71)
    pose = rnd(2, 3, 3)
    noise_c = [n.clone() for n in c.noise]
    img_c, _, w_c = c.gen_batch_by_controls(latent=w.clone(), input_is_latent=True, normalize=False, pose=pose)
    w_ref = w.clone()
    w_ref[:, 24:40] = O.fc_stack(ctl_sd, 'fc_stack.', pose, normalize=False)
    assert max_rel(w_c, w_ref) < 1e-5
    assert max_rel(img_c, O.generator_forward(sd, [w_ref], SIZE, FCG, noise=c.expend_noise(c.noise, 3), input_is_latent=True)) < 1e-5
    assert max_rel(c.get_group_w_latent(w_c, 'pose'), w_ref[:, 24:40]) < 1e-5
    # W+ tensors (the notebook feeds gen_batch's latent_w back)
    img_p, _, _ = c.gen_batch(latent=lat_w.clone(), input_is_latent=True, normalize=False, static_noise=True)
    assert max_rel(img_p, O.generator_forward(sd, [z], SIZE, FCG, noise=c.expend_noise(c.noise, 3))) < 1e-5

    # the reference's directory layout: <dir>/generator/{args.json,checkpoint/NNNNNN.pt}, <dir>/<group>.../
    root = tmp_path / 'controller'
    gdir = root / 'generator'
    os.makedirs(gdir / 'checkpoint')
    os.makedirs(root / 'pose_controller' / 'checkpoint')
    cfg = {'model_config': {'vanilla': False, 'img_channels': 3, 'split_fc': True, 'latent_size': SDIM, 'size': SIZE, 'n_mlp': NMLP,
                            'channel_multiplier': 2, 'conv_transpose': True, 'g_noise_mode': 'normal'},
           'training_config': {'sub_groups_dict': GROUPS, 'mini_batch': 4, 'batch': 4}}
    json.dump(cfg, open(gdir / 'args.json', 'w'))
    torch.save({'g': sd, 'g_ema': sd, 'd': {}}, gdir / 'checkpoint' / '000100.pt')
    json.dump({'model_config': {'lr_mlp': 0.01, 'n_mlp': 3, 'in_dim': 3, 'mid_dim': 32}}, open(root / 'pose_controller' / 'args.json', 'w'))
    torch.save({'controller': ctl_sd}, root / 'pose_controller' / 'checkpoint' / '000010.pt')
    c2 = Controller(str(root), device='cpu', act_dtype=torch.float32)
    assert c2.ckpt_iter == '000100' and list(c2.fc_controls) == ['pose']
    c2.noise = noise_c
    img2, _, _ = c2.gen_batch_by_controls(latent=w.clone(), input_is_latent=True, normalize=False, pose=pose)
    assert max_rel(img2, img_c) < 1e-6


def test_fc_table_is_cached_for_split_fc():
    """The persistent mapping kernel's layer table must be built once (its H2D copy cannot run inside CUDA-graph capture):
    with several latent groups the cache key used to be compared in a different order than it was stored in, and the
    table was rebuilt on every call."""
    from gan_control_b200 import modules as M
    groups = {'a': {'place_in_latent': [0, 8]}, 'b': {'place_in_latent': [8, 24]}, 'c': {'place_in_latent': [24, 32]}}
    g = M.Generator(8, 32, 3, channel_multiplier=2, split_fc=True, fc_config=M.FcConfig.from_sub_groups_dict(groups), conv_transpose=True)
    t1 = M._cached_fc_table(g, g._mapping_groups, torch.device('cpu'))
    t2 = M._cached_fc_table(g, g._mapping_groups, torch.device('cpu'))
    assert t1 is t2 and t1['n_groups'] == 3 and t1['n_layers'] == 3
    g.style.a[1].weight.data = g.style.a[1].weight.data.clone()          # a parameter moved: rebuilt
    assert M._cached_fc_table(g, g._mapping_groups, torch.device('cpu')) is not t1
